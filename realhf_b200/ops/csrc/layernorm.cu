// LayerNorm forward / backward (GPT-2 family: mean subtraction, weight and bias), same structure as rmsnorm.cu: the forward
// keeps the row in registers between the statistics and the scale pass (one read of x, one write of y); the backward walks
// rows with persistent CTAs, keeps the weight / bias gradients in registers and reduces per-CTA partials in a second kernel.
// Replaces `F.layer_norm` (reference: torch.nn.LayerNorm in modules/mlp.py:56-62).
//
// STATUS: compiled for sm_100a, not yet run on hardware: opt-in (`REAL_LAYERNORM=native`), test gated by REAL_TEST_EXPERIMENTAL=1.
#include "common.cuh"

namespace {

constexpr int kLnThreads = 128;

template <typename T, int kMaxVec>
__global__ void __launch_bounds__(kLnThreads) layernorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                                   const T* __restrict__ b, T* __restrict__ y,
                                                                   float* __restrict__ mean_out, float* __restrict__ rstd_out, int H,
                                                                   float eps) {
  __shared__ float red[32];
  constexpr int V = 8;
  const int64_t row = blockIdx.x;
  const int nvec = H / V;
  const rb::Pack<T, V>* xr = reinterpret_cast<const rb::Pack<T, V>*>(x + row * H);
  float vals[kMaxVec][V];
  float s = 0.f;
#pragma unroll
  for (int it = 0; it < kMaxVec; ++it) {
    const int i = threadIdx.x + it * kLnThreads;
    if (i < nvec) {
      rb::Pack<T, V> a = xr[i];
#pragma unroll
      for (int k = 0; k < V; ++k) {
        vals[it][k] = rb::to_f(a.v[k]);
        s += vals[it][k];
      }
    }
  }
  const float mean = rb::block_reduce<false>(s, red) / (float)H;
  float ss = 0.f;  // second pass over registers: centred sum of squares (no catastrophic cancellation)
#pragma unroll
  for (int it = 0; it < kMaxVec; ++it) {
    const int i = threadIdx.x + it * kLnThreads;
    if (i < nvec) {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float d = vals[it][k] - mean;
        ss = fmaf(d, d, ss);
      }
    }
  }
  const float rstd = rsqrtf(rb::block_reduce<false>(ss, red) / (float)H + eps);
  if (threadIdx.x == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const rb::Pack<T, V>* wr = reinterpret_cast<const rb::Pack<T, V>*>(w);
  const rb::Pack<T, V>* br = reinterpret_cast<const rb::Pack<T, V>*>(b);
#pragma unroll
  for (int it = 0; it < kMaxVec; ++it) {
    const int i = threadIdx.x + it * kLnThreads;
    if (i < nvec) {
      rb::Pack<T, V> ww = wr[i], o;
      rb::Pack<T, V> bb;
      if (b != nullptr) bb = br[i];
#pragma unroll
      for (int k = 0; k < V; ++k)
        o.v[k] = rb::from_f<T>((vals[it][k] - mean) * rstd * rb::to_f(ww.v[k]) + (b != nullptr ? rb::to_f(bb.v[k]) : 0.f));
      reinterpret_cast<rb::Pack<T, V>*>(y + row * H)[i] = o;
    }
  }
}

// xhat = (x - mean) * rstd, g = dy * w:  dx = rstd * (g - mean(g) - xhat * mean(g * xhat));  dw += dy * xhat;  db += dy
template <typename T, int kMaxVec>
__global__ void __launch_bounds__(2 * kLnThreads) layernorm_bwd_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                                       const T* __restrict__ dy, const float* __restrict__ mean,
                                                                       const float* __restrict__ rstd, T* __restrict__ dx,
                                                                       float* __restrict__ dw_partial, float* __restrict__ db_partial,
                                                                       int64_t T_rows, int H) {
  __shared__ float red[32];
  constexpr int V = 8;
  const int nvec = H / V;
  float dw[kMaxVec][V], db[kMaxVec][V], wv[kMaxVec][V];
#pragma unroll
  for (int it = 0; it < kMaxVec; ++it) {
    const int i = threadIdx.x + it * 2 * kLnThreads;
#pragma unroll
    for (int k = 0; k < V; ++k) dw[it][k] = db[it][k] = 0.f;
    if (i < nvec) {
      rb::Pack<T, V> ww = reinterpret_cast<const rb::Pack<T, V>*>(w)[i];
#pragma unroll
      for (int k = 0; k < V; ++k) wv[it][k] = rb::to_f(ww.v[k]);
    }
  }
  for (int64_t row = blockIdx.x; row < T_rows; row += gridDim.x) {
    const float mu = mean[row], rs = rstd[row];
    float xh[kMaxVec][V], g[kMaxVec][V];
    float sum_g = 0.f, dot = 0.f;
#pragma unroll
    for (int it = 0; it < kMaxVec; ++it) {
      const int i = threadIdx.x + it * 2 * kLnThreads;
      if (i < nvec) {
        rb::Pack<T, V> a = reinterpret_cast<const rb::Pack<T, V>*>(x + row * H)[i];
        rb::Pack<T, V> d = reinterpret_cast<const rb::Pack<T, V>*>(dy + row * H)[i];
#pragma unroll
        for (int k = 0; k < V; ++k) {
          xh[it][k] = (rb::to_f(a.v[k]) - mu) * rs;
          const float dyk = rb::to_f(d.v[k]);
          g[it][k] = dyk * wv[it][k];
          dw[it][k] = fmaf(dyk, xh[it][k], dw[it][k]);
          db[it][k] += dyk;
          sum_g += g[it][k];
          dot = fmaf(g[it][k], xh[it][k], dot);
        }
      }
    }
    sum_g = rb::block_reduce<false>(sum_g, red) / (float)H;
    dot = rb::block_reduce<false>(dot, red) / (float)H;
#pragma unroll
    for (int it = 0; it < kMaxVec; ++it) {
      const int i = threadIdx.x + it * 2 * kLnThreads;
      if (i < nvec) {
        rb::Pack<T, V> o;
#pragma unroll
        for (int k = 0; k < V; ++k) o.v[k] = rb::from_f<T>(rs * (g[it][k] - sum_g - xh[it][k] * dot));
        reinterpret_cast<rb::Pack<T, V>*>(dx + row * H)[i] = o;
      }
    }
  }
#pragma unroll
  for (int it = 0; it < kMaxVec; ++it) {
    const int i = threadIdx.x + it * 2 * kLnThreads;
    if (i < nvec) {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        dw_partial[(int64_t)blockIdx.x * H + i * V + k] = dw[it][k];
        db_partial[(int64_t)blockIdx.x * H + i * V + k] = db[it][k];
      }
    }
  }
}

template <typename T>
__global__ void ln_colsum_kernel(const float* __restrict__ partial, T* __restrict__ out, int n_part, int H) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  float acc = 0.f;
  for (int p = 0; p < n_part; ++p) acc += partial[(int64_t)p * H + c];
  out[c] = rb::from_f<T>(acc);
}

}  // namespace

extern "C" {

// dt: 0 fp32, 1 bf16, 2 fp16.  H must be a multiple of 8 and <= 8192.  `b` may be null.
int rb_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, int64_t rows, int H, float eps,
                     int dt, cudaStream_t s) {
  if (rows == 0) return 0;
  if (H % 8 != 0 || H > kLnThreads * 8 * 8) return -1;
#define RB_L2(T, NV) \
  layernorm_fwd_kernel<T, NV><<<(unsigned)rows, kLnThreads, 0, s>>>((const T*)x, (const T*)w, (const T*)b, (T*)y, mean, rstd, H, eps);
#define RB_L(T)                                                            \
  { const int nv = RB_CEIL_DIV(H, kLnThreads * 8);                          \
    if (nv <= 1) { RB_L2(T, 1) } else if (nv <= 2) { RB_L2(T, 2) } else if (nv <= 4) { RB_L2(T, 4) } else { RB_L2(T, 8) } }
  if (dt == 0) { RB_L(float) } else if (dt == 1) { RB_L(__nv_bfloat16) } else if (dt == 2) { RB_L(__half) } else return -1;
#undef RB_L
#undef RB_L2
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int rb_layernorm_bwd_num_partials() { return rb::kNumSMs * 4; }

// dw_partial / db_partial: fp32 [num_partials, H] scratch; dw / db: [H] in the parameter dtype (db may be null).
int rb_layernorm_bwd(const void* x, const void* w, const void* dy, const float* mean, const float* rstd, void* dx, float* dw_partial,
                     float* db_partial, void* dw, void* db, int64_t rows, int H, int dt, cudaStream_t s) {
  if (rows == 0) return 0;
  if (H % 8 != 0 || H > kLnThreads * 8 * 8) return -1;
  const int grid = (int)(rows < rb::kNumSMs * 4 ? rows : rb::kNumSMs * 4);
#define RB_L2(T, NV)                                                                                                        \
  layernorm_bwd_kernel<T, NV><<<grid, 2 * kLnThreads, 0, s>>>((const T*)x, (const T*)w, (const T*)dy, mean, rstd, (T*)dx, dw_partial, \
                                                              db_partial, rows, H);
#define RB_L(T)                                                                                      \
  { const int nv = RB_CEIL_DIV(H, 2 * kLnThreads * 8);                                                \
    if (nv <= 1) { RB_L2(T, 1) } else if (nv <= 2) { RB_L2(T, 2) } else { RB_L2(T, 4) }              \
    ln_colsum_kernel<T><<<RB_CEIL_DIV(H, 256), 256, 0, s>>>(dw_partial, (T*)dw, grid, H);           \
    if (db) ln_colsum_kernel<T><<<RB_CEIL_DIV(H, 256), 256, 0, s>>>(db_partial, (T*)db, grid, H); }
  if (dt == 0) { RB_L(float) } else if (dt == 1) { RB_L(__nv_bfloat16) } else if (dt == 2) { RB_L(__half) } else return -1;
#undef RB_L
#undef RB_L2
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
