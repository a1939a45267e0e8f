// Log-prob gather / cross-entropy over a logits tile, forward and in-place backward.
//
// The LM head is applied chunk-by-chunk over tokens (ops/lm_head.py); this kernel turns one chunk of
// logits [rows, V] into log p(label) and the log-sum-exp with a single read (online softmax), and the
// backward overwrites the same buffer with d logits so the full [T, V] fp32 softmax of the reference
// (utils/functional.py:165-211, modules.py:1050-1167) never exists.  Optional features, all fused:
// temperature, a bit-packed "filtered by top-k/top-p at generation time" mask, and a vocab-parallel
// mode where the kernel emits (max, sumexp, label logit) partials for a cross-rank combine.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

RB_DEVICE bool masked_out(const uint8_t* __restrict__ mrow, int j) { return mrow && ((mrow[j >> 3] >> (j & 7)) & 1); }

RB_DEVICE void online_combine(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  if (mn == -INFINITY) { s = 0.f; m = mn; return; }
  s = s * __expf(m - mn) + s2 * __expf(m2 - mn);
  m = mn;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) logprob_fwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels,
                                                               const uint8_t* __restrict__ mask, int64_t mask_stride,
                                                               float* __restrict__ logp, float* __restrict__ lse_out,
                                                               float* __restrict__ max_out, float* __restrict__ sum_out,
                                                               float* __restrict__ tgt_out, int V, int64_t row_stride,
                                                               float inv_temp, int vocab_start) {
  __shared__ float red_m[32], red_s[32];
  const int64_t row = blockIdx.x;
  const T* lr = logits + row * row_stride;
  const uint8_t* mrow = mask ? mask + row * mask_stride : nullptr;
  float m = -INFINITY, s = 0.f;
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = (V % VEC == 0) && (row_stride % VEC == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  if (vec_ok) {
    const int nvec = V / VEC;
    for (int i = threadIdx.x; i < nvec; i += kThreads) {
      rb::Pack<T, VEC> p = reinterpret_cast<const rb::Pack<T, VEC>*>(lr)[i];
      float x[VEC];
      float cm = -INFINITY;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        x[k] = masked_out(mrow, i * VEC + k) ? -INFINITY : rb::to_f(p.v[k]) * inv_temp;
        cm = fmaxf(cm, x[k]);
      }
      if (cm > m) { s *= __expf(m - cm); m = cm; }
      if (m != -INFINITY) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) s += __expf(x[k] - m);
      }
    }
  } else {
    for (int j = threadIdx.x; j < V; j += kThreads) {
      const float x = masked_out(mrow, j) ? -INFINITY : rb::to_f(lr[j]) * inv_temp;
      if (x > m) { s *= __expf(m - x); m = x; }
      if (m != -INFINITY) s += __expf(x - m);
    }
  }
  // warp then block combine
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    online_combine(m, s, m2, s2);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red_m[warp] = m; red_s[warp] = s; }
  __syncthreads();
  if (warp == 0) {
    m = lane < kThreads / 32 ? red_m[lane] : -INFINITY;
    s = lane < kThreads / 32 ? red_s[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      online_combine(m, s, m2, s2);
    }
    if (lane == 0) {
      const int64_t lab = labels[row] - vocab_start;
      const bool own = lab >= 0 && lab < V;
      float tgt = own ? (masked_out(mrow, (int)lab) ? -INFINITY : rb::to_f(lr[lab]) * inv_temp) : 0.f;
      if (max_out) {  // vocab-parallel partials
        max_out[row] = m; sum_out[row] = s; tgt_out[row] = own ? tgt : 0.f;
      } else {
        const float lse = m + __logf(s);
        lse_out[row] = lse;
        logp[row] = tgt - lse;
      }
    }
  }
}

// logits <- dlogp[row] * inv_temp * (onehot(label) - softmax)   (in place)
template <typename T>
__global__ void __launch_bounds__(kThreads) logprob_bwd_kernel(T* __restrict__ logits, const int64_t* __restrict__ labels,
                                                               const uint8_t* __restrict__ mask, int64_t mask_stride,
                                                               const float* __restrict__ lse, const float* __restrict__ dlogp,
                                                               int V, int64_t row_stride, float inv_temp, int vocab_start) {
  const int64_t row = blockIdx.x;
  T* lr = logits + row * row_stride;
  const uint8_t* mrow = mask ? mask + row * mask_stride : nullptr;
  const float l = lse[row];
  const float gscale = dlogp[row] * inv_temp;
  const int lab = (int)(labels[row] - vocab_start);
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = (V % VEC == 0) && (row_stride % VEC == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  if (vec_ok) {
    const int nvec = V / VEC;
    for (int i = threadIdx.x; i < nvec; i += kThreads) {
      rb::Pack<T, VEC> p = reinterpret_cast<rb::Pack<T, VEC>*>(lr)[i];
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const int j = i * VEC + k;
        float g = 0.f;
        if (!masked_out(mrow, j)) g = gscale * ((j == lab ? 1.f : 0.f) - __expf(rb::to_f(p.v[k]) * inv_temp - l));
        p.v[k] = rb::from_f<T>(g);
      }
      reinterpret_cast<rb::Pack<T, VEC>*>(lr)[i] = p;
    }
  } else {
    for (int j = threadIdx.x; j < V; j += kThreads) {
      float g = 0.f;
      if (!masked_out(mrow, j)) g = gscale * ((j == lab ? 1.f : 0.f) - __expf(rb::to_f(lr[j]) * inv_temp - l));
      lr[j] = rb::from_f<T>(g);
    }
  }
}

}  // namespace

extern "C" {

int rb_logprob_fwd(const void* logits, const int64_t* labels, const uint8_t* mask, int64_t mask_stride, float* logp,
                   float* lse, float* max_out, float* sum_out, float* tgt_out, int64_t rows, int V, int64_t row_stride,
                   float inv_temp, int vocab_start, int dt, cudaStream_t s) {
  if (rows == 0) return 0;
#define RB_L(T) logprob_fwd_kernel<T><<<(unsigned)rows, kThreads, 0, s>>>((const T*)logits, labels, mask, mask_stride, logp, lse, max_out, sum_out, tgt_out, V, row_stride, inv_temp, vocab_start)
  if (dt == 0) RB_L(float); else if (dt == 1) RB_L(__nv_bfloat16); else if (dt == 2) RB_L(__half); else return -1;
#undef RB_L
  return 0;
}

int rb_logprob_bwd(void* logits, const int64_t* labels, const uint8_t* mask, int64_t mask_stride, const float* lse,
                   const float* dlogp, int64_t rows, int V, int64_t row_stride, float inv_temp, int vocab_start, int dt,
                   cudaStream_t s) {
  if (rows == 0) return 0;
#define RB_L(T) logprob_bwd_kernel<T><<<(unsigned)rows, kThreads, 0, s>>>((T*)logits, labels, mask, mask_stride, lse, dlogp, V, row_stride, inv_temp, vocab_start)
  if (dt == 0) RB_L(float); else if (dt == 1) RB_L(__nv_bfloat16); else if (dt == 2) RB_L(__half); else return -1;
#undef RB_L
  return 0;
}

}  // extern "C"
