// Shared device helpers for the sm_100a kernels in this directory.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#define RB_DEVICE __device__ __forceinline__
#define RB_CEIL_DIV(a, b) (((a) + (b)-1) / (b))

namespace rb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

RB_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
RB_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reductions through one float per warp of shared memory. `red` holds >= 32 floats.
template <bool kMax>
RB_DEVICE float block_reduce(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = kMax ? warp_max(v) : warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (lane < nwarp) ? red[lane] : (kMax ? -INFINITY : 0.f);
  r = kMax ? warp_max(r) : warp_sum(r);
  return r;
}

RB_DEVICE float to_f(float x) { return x; }
RB_DEVICE float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
RB_DEVICE float to_f(__half x) { return __half2float(x); }
template <typename T> RB_DEVICE T from_f(float x);
template <> RB_DEVICE float from_f<float>(float x) { return x; }
template <> RB_DEVICE __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <> RB_DEVICE __half from_f<__half>(float x) { return __float2half_rn(x); }

// 16-byte streaming load/store (bypass L1 for single-use data).
RB_DEVICE int4 ld_stream(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
RB_DEVICE void st_stream(void* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename T, int N> struct alignas(sizeof(T) * N) Pack { T v[N]; };

// Programmatic dependent launch (PDL).  A kernel launched with `launch_pdl` may become resident while its predecessor in
// the stream is still running (hiding launch latency and its own prologue); it must execute `pdl_wait()` before it
// touches any global memory the predecessor may read or write.  `pdl_trigger()` lets the successor start launching.
// Both are no-ops for kernels launched the ordinary way, so kernels can call them unconditionally.
RB_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
RB_DEVICE void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Process-wide switch (`REAL_PDL=0` / `rb_set_pdl`).  Measured on the LLaMA-7B generation MFC inside CUDA graphs with the
// trigger issued before the wait in the small kernels and the weight prefetch in the small-M GEMM: -4.5% time at 16
// sequences per GPU, neutral at 128 (attention-bound).
extern "C" int rb_get_pdl();
inline bool pdl_enabled() { return rb_get_pdl() != 0; }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace rb
