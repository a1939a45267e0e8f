// Flat-buffer AdamW with fused unscale / clip / skip and fused sum-of-squares (grad norm).
//
// The optimizer owns one flat shard of the parameter buffer (ZeRO-1 layout), so one launch updates
// everything; no multi-tensor-apply lists.  `scale_ptr` (device float) multiplies the gradient — it
// carries 1/loss_scale * clip_coef computed on the device, and `skip_ptr` (device int) suppresses the
// update when an inf/nan was found, so the whole step runs without a host sync.  Replaces apex
// FusedAdam + Megatron's clip/unscale passes (reference: backend/megatron.py:421-497).
//
// State precision is a template parameter: fp32 states (+ optional fp32 master weights) or bf16
// states with stochastic rounding of the bf16 parameter (master-free mode for memory-tight layouts).
#include <type_traits>

#include "adam_math.cuh"

namespace {
using namespace rbadam;

template <typename TP, typename TG, typename TS, bool kMaster, bool kStochastic>
__global__ void __launch_bounds__(256) adamw_kernel(TP* __restrict__ p, const TG* __restrict__ g, TS* __restrict__ m,
                                                    TS* __restrict__ v, float* __restrict__ master, int64_t n, float lr,
                                                    float b1, float b2, float eps, float wd, float bc1, float bc2,
                                                    const float* __restrict__ scale_ptr, const int* __restrict__ skip_ptr,
                                                    uint32_t seed) {
  if (skip_ptr != nullptr && *skip_ptr != 0) return;
  const float gscale = scale_ptr ? *scale_ptr : 1.f;
  const float inv_bc1 = 1.f / bc1, inv_bc2 = 1.f / bc2;
  constexpr int V = 8;
  const int64_t nvec = n / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    rb::Pack<TP, V> pp = reinterpret_cast<rb::Pack<TP, V>*>(p)[i];
    rb::Pack<TG, V> gg = reinterpret_cast<const rb::Pack<TG, V>*>(g)[i];
    rb::Pack<TS, V> mm = reinterpret_cast<rb::Pack<TS, V>*>(m)[i];
    rb::Pack<TS, V> vv = reinterpret_cast<rb::Pack<TS, V>*>(v)[i];
    rb::Pack<float, V> ms;
    if constexpr (kMaster) ms = reinterpret_cast<rb::Pack<float, V>*>(master)[i];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float grad = rb::to_f(gg.v[k]) * gscale;
      float w = kMaster ? ms.v[k] : rb::to_f(pp.v[k]);
      float mk = b1 * rb::to_f(mm.v[k]) + (1.f - b1) * grad;
      float vk = b2 * rb::to_f(vv.v[k]) + (1.f - b2) * grad * grad;
      const float upd = __fdividef(mk * inv_bc1, sqrt_approx(vk * inv_bc2) + eps) + wd * w;  // IEEE div/sqrt made this kernel ALU-bound
      w -= lr * upd;
      mm.v[k] = rb::from_f<TS>(mk);
      vv.v[k] = rb::from_f<TS>(vk);
      if constexpr (kMaster) ms.v[k] = w;
      if constexpr (kStochastic) {
        pp.v[k] = sr_bf16(w, hash32(seed ^ (uint32_t)(i * V + k)) );
      } else {
        pp.v[k] = rb::from_f<TP>(w);
      }
    }
    reinterpret_cast<rb::Pack<TP, V>*>(p)[i] = pp;
    reinterpret_cast<rb::Pack<TS, V>*>(m)[i] = mm;
    reinterpret_cast<rb::Pack<TS, V>*>(v)[i] = vv;
    if constexpr (kMaster) reinterpret_cast<rb::Pack<float, V>*>(master)[i] = ms;
  }
  // tail (n % 8) handled by thread 0 of block 0
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t j = nvec * V; j < n; ++j) {
      const float grad = rb::to_f(g[j]) * gscale;
      float w = kMaster ? master[j] : rb::to_f(p[j]);
      float mk = b1 * rb::to_f(m[j]) + (1.f - b1) * grad;
      float vk = b2 * rb::to_f(v[j]) + (1.f - b2) * grad * grad;
      w -= lr * ((mk / bc1) / (sqrtf(vk / bc2) + eps) + wd * w);
      m[j] = rb::from_f<TS>(mk);
      v[j] = rb::from_f<TS>(vk);
      if constexpr (kMaster) master[j] = w;
      if constexpr (kStochastic) p[j] = sr_bf16(w, hash32(seed ^ (uint32_t)j));
      else p[j] = rb::from_f<TP>(w);
    }
  }
}

// out[0] += sum g^2 ; out[1] += count of non-finite values
template <typename TG>
__global__ void __launch_bounds__(256) sumsq_kernel(const TG* __restrict__ g, int64_t n, float* __restrict__ out) {
  __shared__ float red[32];
  constexpr int V = 16 / sizeof(TG);
  float acc = 0.f, bad = 0.f;
  const int64_t nvec = n / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    rb::Pack<TG, V> gg = reinterpret_cast<const rb::Pack<TG, V>*>(g)[i];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float x = rb::to_f(gg.v[k]);
      acc = fmaf(x, x, acc);
      bad += (isfinite(x) ? 0.f : 1.f);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t j = nvec * V; j < n; ++j) { const float x = rb::to_f(g[j]); acc = fmaf(x, x, acc); bad += isfinite(x) ? 0.f : 1.f; }
  acc = rb::block_reduce<false>(acc, red);
  bad = rb::block_reduce<false>(bad, red);
  if (threadIdx.x == 0) { atomicAdd(out, acc); if (bad != 0.f) atomicAdd(out + 1, bad); }
}

template <typename TP, typename TG, typename TS>
void launch_adam(void* p, const void* g, void* m, void* v, float* master, int64_t n, float lr, float b1, float b2,
                 float eps, float wd, float bc1, float bc2, const float* scale_ptr, const int* skip_ptr, bool stochastic,
                 uint32_t seed, cudaStream_t s) {
  const int grid = rb::kNumSMs * 8;
#define RB_ADAM(MASTER, SR) adamw_kernel<TP, TG, TS, MASTER, SR><<<grid, 256, 0, s>>>( \
    (TP*)p, (const TG*)g, (TS*)m, (TS*)v, master, n, lr, b1, b2, eps, wd, bc1, bc2, scale_ptr, skip_ptr, seed)
  if (master) RB_ADAM(true, false);
  else if constexpr (std::is_same<TP, __nv_bfloat16>::value) { if (stochastic) RB_ADAM(false, true); else RB_ADAM(false, false); }
  else RB_ADAM(false, false);
#undef RB_ADAM
}

}  // namespace

extern "C" {

// dtype codes: 0 = fp32, 1 = bf16
int rb_adamw(void* p, int p_dt, const void* g, int g_dt, void* m, void* v, int s_dt, float* master, int64_t n, float lr,
             float b1, float b2, float eps, float wd, int step, const float* scale_ptr, const int* skip_ptr,
             int stochastic, uint32_t seed, cudaStream_t s) {
  if (n == 0) return 0;
  const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  const int key = p_dt * 4 + g_dt * 2 + s_dt;
#define RB_CASE(K, TP, TG, TS) case K: launch_adam<TP, TG, TS>(p, g, m, v, master, n, lr, b1, b2, eps, wd, bc1, bc2, scale_ptr, skip_ptr, stochastic != 0, seed, s); break;
  switch (key) {
    RB_CASE(0, float, float, float)
    RB_CASE(2, float, __nv_bfloat16, float)
    RB_CASE(4, __nv_bfloat16, float, float)
    RB_CASE(5, __nv_bfloat16, float, __nv_bfloat16)
    RB_CASE(6, __nv_bfloat16, __nv_bfloat16, float)
    RB_CASE(7, __nv_bfloat16, __nv_bfloat16, __nv_bfloat16)
    default: return -1;
  }
#undef RB_CASE
  return 0;
}

int rb_sumsq(const void* g, int g_dt, int64_t n, float* out2, cudaStream_t s) {
  if (n == 0) return 0;
  const int grid = rb::kNumSMs * 4;
  if (g_dt == 0) sumsq_kernel<float><<<grid, 256, 0, s>>>((const float*)g, n, out2);
  else if (g_dt == 1) sumsq_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)g, n, out2);
  else return -1;
  return 0;
}

}  // extern "C"
