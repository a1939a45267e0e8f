// Scalar pieces of the AdamW update shared by the plain flat kernel (adam.cu) and the NVLS-fused
// reduce-scatter / Adam / all-gather kernels (nvls.cu).
#pragma once
#include "common.cuh"

namespace rbadam {

RB_DEVICE uint32_t hash32(uint32_t x) {  // lowbias32
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// Round fp32 -> bf16 stochastically using 16 random bits.
RB_DEVICE __nv_bfloat16 sr_bf16(float x, uint32_t rnd16) {
  uint32_t u = __float_as_uint(x);
  if ((u & 0x7f800000u) != 0x7f800000u) u += (rnd16 & 0xffffu);
  return __ushort_as_bfloat16((unsigned short)(u >> 16));
}

RB_DEVICE float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One AdamW element update.  Returns the new fp32 weight; m / v are updated in place.
RB_DEVICE float adam_elem(float w, float grad, float& mk, float& vk, float lr, float b1, float b2, float eps, float wd,
                          float inv_bc1, float inv_bc2) {
  mk = b1 * mk + (1.f - b1) * grad;
  vk = b2 * vk + (1.f - b2) * grad * grad;
  const float upd = __fdividef(mk * inv_bc1, sqrt_approx(vk * inv_bc2) + eps) + wd * w;
  return w - lr * upd;
}

}  // namespace rbadam
