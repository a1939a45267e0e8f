// Row-wise e4m3 quantisation for the W8A8 generation path.
//
//   q[m, k] = sat_e4m3(x[m, k] / s[m]),   s[m] = max_k |x[m, k]| / 448   (1 when the row is all zero)
//
// One block per row; the row is held in registers between the abs-max pass and the conversion (rows of the decode path
// are at most a few 10^4 elements), so the input is read once.  Three producers share the quantising tail:
//   * plain rows          (weights: rows = output channels, once per generation call; attention output every step)
//   * gated activation    x = [gate | up] -> act(gate) * up -> e4m3       (input of the down projection)
//   * residual add + RMSNorm -> e4m3                                       (input of the qkv / gate|up projections)
// so that a W8A8 decode layer has one kernel more than the bf16 layer (the quantiser after attention), not four.  The fused
// producers round their value to the model dtype before quantising: bit-identical to running the unfused kernels
// (`rmsnorm.cu`, `elementwise.cu`) followed by the plain quantiser.  All kernels are PDL-chained.
// The reference has no quantised generation path; this is a B200 addition (tcgen05.mma.kind::f8f6f4 consumes the bytes).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxVec = 8;  // 8-element vectors per thread held in registers: K <= 256 * 8 * 8 = 16384

// IEEE division (not a multiplication by an approximate reciprocal, and immune to --use_fast_math): x / s is exact surprisingly
// often -- x and the row maximum are both bf16 values, s = amax / 448 -- and exact ties must round the way every other
// implementation of the same rule rounds them (round-to-nearest-even on the exact quotient), or 3% of the elements of such a row
// land one e4m3 step away from the PyTorch reference
RB_DEVICE uint2 cvt8_e4m3(const float* v, float s) {
  uint32_t w[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const __nv_fp8x2_storage_t lo =
        __nv_cvt_float2_to_fp8x2(make_float2(__fdiv_rn(v[4 * h], s), __fdiv_rn(v[4 * h + 1], s)), __NV_SATFINITE, __NV_E4M3);
    const __nv_fp8x2_storage_t hi =
        __nv_cvt_float2_to_fp8x2(make_float2(__fdiv_rn(v[4 * h + 2], s), __fdiv_rn(v[4 * h + 3], s)), __NV_SATFINITE, __NV_E4M3);
    w[h] = (uint32_t)lo | ((uint32_t)hi << 16);
  }
  return make_uint2(w[0], w[1]);
}

RB_DEVICE float act_fwd(float g, int kind) {  // same expressions as elementwise.cu
  if (kind == 0) return g / (1.f + __expf(-g));
  const float u = 0.7978845608028654f * (g + 0.044715f * g * g * g);
  return 0.5f * g * (1.f + tanhf(u));
}

// Common tail: vals[j][e] of this thread's vectors -> scale[row], q row.
template <int NV>
RB_DEVICE void quant_tail(float (&vals)[NV][8], int nvec, uint8_t* qr, float* scale, int row, float* red) {
  float amax = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if ((int)threadIdx.x + j * kThreads < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(vals[j][e]));
    }
  }
  amax = rb::block_reduce<true>(amax, red);
  const float s = amax > 0.f ? __fdiv_rn(amax, 448.f) : 1.f;
  if (threadIdx.x == 0) scale[row] = s;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = threadIdx.x + j * kThreads;
    if (i < nvec) *reinterpret_cast<uint2*>(qr + (int64_t)i * 8) = cvt8_e4m3(vals[j], s);
  }
}

// kGlu: x rows are [gate (K) | up (K)], the quantised value is act(gate) * up rounded to T.
template <typename T, bool kGlu, int NV>
__global__ void __launch_bounds__(kThreads) quant_rows_e4m3_kernel(const T* __restrict__ x, uint8_t* __restrict__ q,
                                                                   float* __restrict__ scale, int K, int64_t ld_x, int64_t ld_q, int act) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const T* xr = x + (int64_t)row * ld_x;
  const int nvec = K >> 3;
  rb::pdl_trigger();
  rb::pdl_wait();
  float vals[NV][8];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = threadIdx.x + j * kThreads;
    if (i < nvec) {
      const rb::Pack<T, 8> a = *reinterpret_cast<const rb::Pack<T, 8>*>(xr + (int64_t)i * 8);
      if constexpr (kGlu) {
        const rb::Pack<T, 8> u = *reinterpret_cast<const rb::Pack<T, 8>*>(xr + K + (int64_t)i * 8);
#pragma unroll
        for (int e = 0; e < 8; ++e) vals[j][e] = rb::to_f(rb::from_f<T>(act_fwd(rb::to_f(a.v[e]), act) * rb::to_f(u.v[e])));
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) vals[j][e] = rb::to_f(a.v[e]);
      }
    }
  }
  quant_tail<NV>(vals, nvec, q + (int64_t)row * ld_q, scale, row, red);
}

// (x [+ res_in]) -> res_out (the new residual stream, rounded to T), RMSNorm of it with weight w (+ w_offset) rounded to T,
// quantised.  Same arithmetic as rmsnorm_fwd_kernel.
template <typename T, bool kResidual, int NV>
__global__ void __launch_bounds__(kThreads) add_rmsnorm_quant_kernel(const T* __restrict__ x, const T* __restrict__ res_in,
                                                                     const T* __restrict__ w, T* __restrict__ res_out,
                                                                     uint8_t* __restrict__ q, float* __restrict__ scale, int H,
                                                                     int64_t ld_q, float eps, float w_offset) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const int nvec = H >> 3;
  rb::pdl_trigger();
  rb::pdl_wait();
  float vals[NV][8];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = threadIdx.x + j * kThreads;
    if (i < nvec) {
      const rb::Pack<T, 8> a = reinterpret_cast<const rb::Pack<T, 8>*>(x + row * H)[i];
      if constexpr (kResidual) {
        const rb::Pack<T, 8> b = reinterpret_cast<const rb::Pack<T, 8>*>(res_in + row * H)[i];
        rb::Pack<T, 8> o;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          o.v[e] = rb::from_f<T>(rb::to_f(a.v[e]) + rb::to_f(b.v[e]));
          vals[j][e] = rb::to_f(o.v[e]);
        }
        reinterpret_cast<rb::Pack<T, 8>*>(res_out + row * H)[i] = o;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) vals[j][e] = rb::to_f(a.v[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) ss = fmaf(vals[j][e], vals[j][e], ss);
    }
  }
  ss = rb::block_reduce<false>(ss, red);
  const float rstd = rsqrtf(ss / (float)H + eps);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = threadIdx.x + j * kThreads;
    if (i < nvec) {
      const rb::Pack<T, 8> ww = reinterpret_cast<const rb::Pack<T, 8>*>(w)[i];
#pragma unroll
      for (int e = 0; e < 8; ++e) vals[j][e] = rb::to_f(rb::from_f<T>(vals[j][e] * rstd * (rb::to_f(ww.v[e]) + w_offset)));
    }
  }
  quant_tail<NV>(vals, nvec, q + row * ld_q, scale, (int)row, red);
}

template <typename T, bool kGlu>
int launch_quant(const void* x, void* q, float* scale, int M, int K, int64_t ld_x, int64_t ld_q, int act, cudaStream_t s) {
  const int nv = RB_CEIL_DIV(K >> 3, kThreads);
  cudaError_t e;
#define RB_Q(NV)                                                                                                         \
  e = rb::launch_pdl(quant_rows_e4m3_kernel<T, kGlu, NV>, dim3(M), dim3(kThreads), 0, s, (const T*)x, (uint8_t*)q, scale, K, ld_x, \
                     ld_q, act);
  if (nv <= 1) { RB_Q(1) } else if (nv <= 2) { RB_Q(2) } else if (nv <= 4) { RB_Q(4) } else if (nv <= 6) { RB_Q(6) } else { RB_Q(8) }
#undef RB_Q
  return e == cudaSuccess ? 0 : -100 - (int)e;
}

}  // namespace

extern "C" {

// x [M, K] (dt: 1 bf16, 2 fp16, 0 fp32; row pitch ld_x elements, 16-byte aligned rows), q [M, K] bytes (pitch ld_q, 8-byte
// aligned rows), scale [M] fp32.  K % 8 == 0, K <= 16384.
int rb_quant_rows_e4m3(const void* x, void* q, float* scale, int M, int K, int64_t ld_x, int64_t ld_q, int dt, cudaStream_t s) {
  if (M <= 0 || K <= 0) return 0;
  if ((K & 7) || K > kThreads * kMaxVec * 8 || (ld_q & 7) || (ld_x & 7)) return -1;
  if (dt == 1) return launch_quant<__nv_bfloat16, false>(x, q, scale, M, K, ld_x, ld_q, 0, s);
  if (dt == 2) return launch_quant<__half, false>(x, q, scale, M, K, ld_x, ld_q, 0, s);
  if (dt == 0) return launch_quant<float, false>(x, q, scale, M, K, ld_x, ld_q, 0, s);
  return -3;
}

// gu [M, 2F] = [gate | up] -> q [M, F] = e4m3(act(gate) * up).  act: 0 silu, 1 gelu(tanh).
int rb_gated_act_quant_e4m3(const void* gu, void* q, float* scale, int M, int F, int64_t ld_x, int64_t ld_q, int act, int dt,
                            cudaStream_t s) {
  if (M <= 0 || F <= 0) return 0;
  if ((F & 7) || F > kThreads * kMaxVec * 8 || (ld_q & 7) || (ld_x & 7)) return -1;
  if (dt == 1) return launch_quant<__nv_bfloat16, true>(gu, q, scale, M, F, ld_x, ld_q, act, s);
  if (dt == 2) return launch_quant<__half, true>(gu, q, scale, M, F, ld_x, ld_q, act, s);
  return -3;
}

// Contiguous [rows, H]: res_out = x + res_in (when res_in != nullptr), q / scale = e4m3(rmsnorm(res_out) * (w + w_offset)).
int rb_add_rmsnorm_quant_e4m3(const void* x, const void* res_in, const void* w, void* res_out, void* q, float* scale, int64_t rows,
                              int H, int64_t ld_q, float eps, float w_offset, int dt, cudaStream_t s) {
  if (rows <= 0) return 0;
  if ((H & 7) || H > kThreads * 4 * 8 || (ld_q & 7)) return -1;
  const int nv = RB_CEIL_DIV(H >> 3, kThreads);
  cudaError_t e;
#define RB_N2(T, NV)                                                                                                             \
  if (res_in) e = rb::launch_pdl(add_rmsnorm_quant_kernel<T, true, NV>, dim3((unsigned)rows), dim3(kThreads), 0, s, (const T*)x,     \
                                 (const T*)res_in, (const T*)w, (T*)res_out, (uint8_t*)q, scale, H, ld_q, eps, w_offset);            \
  else e = rb::launch_pdl(add_rmsnorm_quant_kernel<T, false, NV>, dim3((unsigned)rows), dim3(kThreads), 0, s, (const T*)x,           \
                          (const T*)nullptr, (const T*)w, (T*)nullptr, (uint8_t*)q, scale, H, ld_q, eps, w_offset);
#define RB_N(T) { if (nv <= 1) { RB_N2(T, 1) } else if (nv <= 2) { RB_N2(T, 2) } else { RB_N2(T, 4) } }
  if (dt == 1) RB_N(__nv_bfloat16) else if (dt == 2) RB_N(__half) else return -3;
#undef RB_N
#undef RB_N2
  return e == cudaSuccess ? 0 : -100 - (int)e;
}

}  // extern "C"
