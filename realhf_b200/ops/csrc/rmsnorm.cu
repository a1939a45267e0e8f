// RMSNorm forward / backward (LLaMA and Gemma "1+w" flavours) with an optional fused residual add.
//
// Forward: one CTA (128 threads) per token row; the row stays in registers between the
// sum-of-squares pass and the scale pass, so HBM traffic is exactly read-x (+read-residual) and
// write-y (+write-residual).  Backward: persistent CTAs walk rows, produce dx and keep the weight
// gradient in registers; per-CTA partials are reduced by a second tiny kernel.
// Replaces the eager fp32-upcast RMSNorm of the reference (modules/mlp.py:425-467).
#include "common.cuh"

namespace {

constexpr int kThreads = 128;

template <typename T, bool kResidual, int kMaxVec>
__global__ void __launch_bounds__(kThreads) rmsnorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res_in,
                                                               const T* __restrict__ w, T* __restrict__ y,
                                                               T* __restrict__ res_out, float* __restrict__ rstd_out,
                                                               int H, float eps, float w_offset) {
  __shared__ float red[32];
  constexpr int V = 8;
  const int64_t row = blockIdx.x;
  const int nvec = H / V;
  rb::pdl_trigger();  // successors may start launching (and prefetching weights) while this kernel still waits below
  rb::pdl_wait();
  const rb::Pack<T, V>* xr = reinterpret_cast<const rb::Pack<T, V>*>(x + row * H);
  const rb::Pack<T, V>* rr = kResidual ? reinterpret_cast<const rb::Pack<T, V>*>(res_in + row * H) : nullptr;
  float vals[kMaxVec][V];
  float ss = 0.f;
#pragma unroll
  for (int it = 0; it < kMaxVec; ++it) {
    const int i = threadIdx.x + it * kThreads;
    if (i < nvec) {
      rb::Pack<T, V> a = xr[i];
      if constexpr (kResidual) {
        rb::Pack<T, V> b = rr[i];
        rb::Pack<T, V> o;
#pragma unroll
        for (int k = 0; k < V; ++k) {
          // round the residual stream to T first so the normalised value matches what is stored
          o.v[k] = rb::from_f<T>(rb::to_f(a.v[k]) + rb::to_f(b.v[k]));
          vals[it][k] = rb::to_f(o.v[k]);
        }
        reinterpret_cast<rb::Pack<T, V>*>(res_out + row * H)[i] = o;
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) vals[it][k] = rb::to_f(a.v[k]);
      }
#pragma unroll
      for (int k = 0; k < V; ++k) ss = fmaf(vals[it][k], vals[it][k], ss);
    }
  }
  ss = rb::block_reduce<false>(ss, red);
  const float rstd = rsqrtf(ss / (float)H + eps);
  if (threadIdx.x == 0 && rstd_out) rstd_out[row] = rstd;
  const rb::Pack<T, V>* wr = reinterpret_cast<const rb::Pack<T, V>*>(w);
#pragma unroll
  for (int it = 0; it < kMaxVec; ++it) {
    const int i = threadIdx.x + it * kThreads;
    if (i < nvec) {
      rb::Pack<T, V> ww = wr[i], o;
#pragma unroll
      for (int k = 0; k < V; ++k) o.v[k] = rb::from_f<T>(vals[it][k] * rstd * (rb::to_f(ww.v[k]) + w_offset));
      reinterpret_cast<rb::Pack<T, V>*>(y + row * H)[i] = o;
    }
  }
}

// dx = rstd * (g - xhat * mean(g*xhat)),  g = dy * (w + off);  dw += dy * xhat
template <typename T, int kMaxVec>
__global__ void __launch_bounds__(2 * kThreads) rmsnorm_bwd_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                               const T* __restrict__ dy, const float* __restrict__ rstd,
                                                               T* __restrict__ dx, float* __restrict__ dw_partial,
                                                               int64_t T_rows, int H, float w_offset,
                                                               const T* __restrict__ dres) {
  __shared__ float red[32];
  constexpr int V = 8;
  const int nvec = H / V;
  float dw[kMaxVec][V];
  float wv[kMaxVec][V];
#pragma unroll
  for (int it = 0; it < kMaxVec; ++it) {
    const int i = threadIdx.x + it * 2 * kThreads;
#pragma unroll
    for (int k = 0; k < V; ++k) dw[it][k] = 0.f;
    if (i < nvec) {
      rb::Pack<T, V> ww = reinterpret_cast<const rb::Pack<T, V>*>(w)[i];
#pragma unroll
      for (int k = 0; k < V; ++k) wv[it][k] = rb::to_f(ww.v[k]) + w_offset;
    }
  }
  for (int64_t row = blockIdx.x; row < T_rows; row += gridDim.x) {
    const float rs = rstd[row];
    float xh[kMaxVec][V], g[kMaxVec][V];
    float dot = 0.f;
#pragma unroll
    for (int it = 0; it < kMaxVec; ++it) {
      const int i = threadIdx.x + it * 2 * kThreads;
      if (i < nvec) {
        rb::Pack<T, V> a = reinterpret_cast<const rb::Pack<T, V>*>(x + row * H)[i];
        rb::Pack<T, V> d = reinterpret_cast<const rb::Pack<T, V>*>(dy + row * H)[i];
#pragma unroll
        for (int k = 0; k < V; ++k) {
          xh[it][k] = rb::to_f(a.v[k]) * rs;
          const float dyk = rb::to_f(d.v[k]);
          g[it][k] = dyk * wv[it][k];
          dw[it][k] = fmaf(dyk, xh[it][k], dw[it][k]);
          dot = fmaf(g[it][k], xh[it][k], dot);
        }
      }
    }
    dot = rb::block_reduce<false>(dot, red) / (float)H;
#pragma unroll
    for (int it = 0; it < kMaxVec; ++it) {
      const int i = threadIdx.x + it * 2 * kThreads;
      if (i < nvec) {
        rb::Pack<T, V> o;
        if (dres != nullptr) {  // fused residual add in training: the gradient arriving on the residual stream joins here
          rb::Pack<T, V> e = reinterpret_cast<const rb::Pack<T, V>*>(dres + row * H)[i];
#pragma unroll
          for (int k = 0; k < V; ++k) o.v[k] = rb::from_f<T>(rs * (g[it][k] - xh[it][k] * dot) + rb::to_f(e.v[k]));
        } else {
#pragma unroll
          for (int k = 0; k < V; ++k) o.v[k] = rb::from_f<T>(rs * (g[it][k] - xh[it][k] * dot));
        }
        reinterpret_cast<rb::Pack<T, V>*>(dx + row * H)[i] = o;
      }
    }
  }
#pragma unroll
  for (int it = 0; it < kMaxVec; ++it) {
    const int i = threadIdx.x + it * 2 * kThreads;
    if (i < nvec) {
#pragma unroll
      for (int k = 0; k < V; ++k) dw_partial[(int64_t)blockIdx.x * H + i * V + k] = dw[it][k];
    }
  }
}

template <typename T>
__global__ void colsum_kernel(const float* __restrict__ partial, T* __restrict__ out, int n_part, int H) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  float acc = 0.f;
  for (int p = 0; p < n_part; ++p) acc += partial[(int64_t)p * H + c];
  out[c] = rb::from_f<T>(acc);
}

}  // namespace

extern "C" {

// dt: 0 fp32, 1 bf16, 2 fp16.  H must be a multiple of 8 and <= 8192.
int rb_rmsnorm_fwd(const void* x, const void* res_in, const void* w, void* y, void* res_out, float* rstd, int64_t rows,
                   int H, float eps, float w_offset, int dt, cudaStream_t s) {
  if (rows == 0) return 0;
  if (H % 8 != 0 || H > kThreads * 8 * 8) return -1;
#define RB_L2(T, NV)                                                                                               \
  if (res_in) rb::launch_pdl(rmsnorm_fwd_kernel<T, true, NV>, dim3((unsigned)rows), dim3(kThreads), 0, s,                \
      (const T*)x, (const T*)res_in, (const T*)w, (T*)y, (T*)res_out, rstd, H, eps, w_offset);                         \
  else rb::launch_pdl(rmsnorm_fwd_kernel<T, false, NV>, dim3((unsigned)rows), dim3(kThreads), 0, s, (const T*)x,       \
                      (const T*)nullptr, (const T*)w, (T*)y, (T*)nullptr, rstd, H, eps, w_offset);
#define RB_L(T)                                                            \
  { const int nv = RB_CEIL_DIV(H, kThreads * 8);                            \
    if (nv <= 1) { RB_L2(T, 1) } else if (nv <= 2) { RB_L2(T, 2) } else if (nv <= 4) { RB_L2(T, 4) } else { RB_L2(T, 8) } }
  if (dt == 0) { RB_L(float) } else if (dt == 1) { RB_L(__nv_bfloat16) } else if (dt == 2) { RB_L(__half) } else return -1;
#undef RB_L
#undef RB_L2
  return 0;
}

int rb_rmsnorm_bwd_num_partials() { return rb::kNumSMs * 4; }

int rb_rmsnorm_bwd(const void* x, const void* w, const void* dy, const float* rstd, void* dx, float* dw_partial,
                   void* dw, int64_t rows, int H, float w_offset, int dt, const void* dres, cudaStream_t s) {
  if (rows == 0) return 0;
  if (H % 8 != 0 || H > kThreads * 8 * 8) return -1;
  const int grid = (int)(rows < rb::kNumSMs * 4 ? rows : rb::kNumSMs * 4);
#define RB_L2(T, NV)                                                                                               \
  rmsnorm_bwd_kernel<T, NV><<<grid, 2 * kThreads, 0, s>>>((const T*)x, (const T*)w, (const T*)dy, rstd, (T*)dx,        \
                                                          dw_partial, rows, H, w_offset, (const T*)dres);
#define RB_L(T)                                                            \
  { const int nv = RB_CEIL_DIV(H, 2 * kThreads * 8);                        \
    if (nv <= 1) { RB_L2(T, 1) } else if (nv <= 2) { RB_L2(T, 2) } else { RB_L2(T, 4) }                            \
    colsum_kernel<T><<<RB_CEIL_DIV(H, 256), 256, 0, s>>>(dw_partial, (T*)dw, grid, H); }
  if (dt == 0) { RB_L(float) } else if (dt == 1) { RB_L(__nv_bfloat16) } else if (dt == 2) { RB_L(__half) } else return -1;
#undef RB_L
#undef RB_L2
  return 0;
}

}  // extern "C"
