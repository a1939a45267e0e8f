// Row-wise e4m3 quantisation for the W8A8 generation path.
//
//   q[m, k] = sat_e4m3(x[m, k] / s[m]),   s[m] = max_k |x[m, k]| / 448   (1 when the row is all zero)
//
// One block per row; the row is held in registers between the abs-max pass and the conversion (rows of the decode path
// are at most a few 10^4 elements), so x is read once.  Used for weights (rows = output channels, once per generation
// call) and for activations (rows = tokens, every decode step, PDL-chained between the producing kernel and the GEMM).
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

template <typename T>
__global__ void __launch_bounds__(kThreads) quant_rows_e4m3_kernel(const T* __restrict__ x, uint8_t* __restrict__ q,
                                                                   float* __restrict__ scale, int K, int64_t ld_x, int64_t ld_q) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const T* xr = x + (int64_t)row * ld_x;
  uint8_t* qr = q + (int64_t)row * ld_q;
  const int nvec = K >> 3;
  rb::pdl_trigger();
  rb::pdl_wait();
  rb::Pack<T, 8> v[kMaxVec];
  float amax = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = threadIdx.x + j * kThreads;
    if (i < nvec) {
      v[j] = *reinterpret_cast<const rb::Pack<T, 8>*>(xr + (int64_t)i * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(rb::to_f(v[j].v[e])));
    }
  }
  amax = rb::block_reduce<true>(amax, red);
  const float s = amax > 0.f ? amax * (1.f / 448.f) : 1.f;
  const float inv = 1.f / s;
  if (threadIdx.x == 0) scale[row] = s;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = threadIdx.x + j * kThreads;
    if (i < nvec) {
      uint32_t w[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const __nv_fp8x2_storage_t lo = __nv_cvt_float2_to_fp8x2(
            make_float2(rb::to_f(v[j].v[4 * h]) * inv, rb::to_f(v[j].v[4 * h + 1]) * inv), __NV_SATFINITE, __NV_E4M3);
        const __nv_fp8x2_storage_t hi = __nv_cvt_float2_to_fp8x2(
            make_float2(rb::to_f(v[j].v[4 * h + 2]) * inv, rb::to_f(v[j].v[4 * h + 3]) * inv), __NV_SATFINITE, __NV_E4M3);
        w[h] = (uint32_t)lo | ((uint32_t)hi << 16);
      }
      *reinterpret_cast<uint2*>(qr + (int64_t)i * 8) = make_uint2(w[0], w[1]);
    }
  }
}

}  // namespace

// x [M, K] (dt: 1 bf16, 2 fp16, 0 fp32; row pitch ld_x elements, 16-byte aligned rows), q [M, K] bytes (pitch ld_q, 8-byte
// aligned rows), scale [M] fp32.  K % 8 == 0, K <= 16384.
extern "C" int rb_quant_rows_e4m3(const void* x, void* q, float* scale, int M, int K, int64_t ld_x, int64_t ld_q, int dt,
                                  cudaStream_t s) {
  if (M <= 0 || K <= 0) return 0;
  if ((K & 7) || K > kThreads * kMaxVec * 8 || (ld_q & 7)) return -1;
  cudaError_t e;
  if (dt == 1) {
    if (ld_x & 7) return -2;
    e = rb::launch_pdl(quant_rows_e4m3_kernel<__nv_bfloat16>, dim3(M), dim3(kThreads), 0, s, (const __nv_bfloat16*)x, (uint8_t*)q, scale, K,
                       ld_x, ld_q);
  } else if (dt == 2) {
    if (ld_x & 7) return -2;
    e = rb::launch_pdl(quant_rows_e4m3_kernel<__half>, dim3(M), dim3(kThreads), 0, s, (const __half*)x, (uint8_t*)q, scale, K, ld_x, ld_q);
  } else if (dt == 0) {
    if (ld_x & 7) return -2;
    e = rb::launch_pdl(quant_rows_e4m3_kernel<float>, dim3(M), dim3(kThreads), 0, s, (const float*)x, (uint8_t*)q, scale, K, ld_x, ld_q);
  } else {
    return -3;
  }
  return e == cudaSuccess ? 0 : -100 - (int)e;
}
