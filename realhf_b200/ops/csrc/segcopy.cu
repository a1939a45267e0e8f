// Segment copy: the layout-transforming copy behind parameter reallocation.
//
// A "plan" is a list of segments (src_byte_off, dst_byte_off) with cumulative byte lengths.  One
// launch moves every segment; `dst` may be a pointer into a *peer GPU's* memory (mapped through
// CUDA IPC / VMM), in which case the kernel is the realloc transport itself: it writes the
// destination layout directly over NVLink with 16-byte stores, so there is no pack kernel, no
// NCCL broadcast and no unpack kernel (reference: interval_op.cu:12-104 + real_llm_api.py:706-758).
// The same kernel implements slice_intervals (dst packed) and set_intervals (src packed).
//
// Work is split in the packed byte space, so load balance does not depend on segment sizes:
// every CTA takes 16 KiB tiles, each thread locates its segment with one binary search per tile
// and then walks forward.
#include "common.cuh"

namespace {

constexpr int kTileBytes = 16384;
constexpr int kThreads = 256;

RB_DEVICE int find_segment(const int64_t* __restrict__ cum, int n, int64_t pos) {
  int lo = 0, hi = n;  // cum has n+1 entries; find seg with cum[seg] <= pos < cum[seg+1]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (cum[mid] <= pos) lo = mid; else hi = mid;
  }
  return lo;
}

// scale != nullptr turns the copy into an EMA merge dst = eta*src + (1-eta)*dst on bf16 (ref-EMA realloc).
template <bool kEma>
__global__ void __launch_bounds__(kThreads) segcopy_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                           const int64_t* __restrict__ src_off,
                                                           const int64_t* __restrict__ dst_off,
                                                           const int64_t* __restrict__ cum, int n_seg, int64_t total,
                                                           float eta) {
  const int64_t n_tiles = (total + kTileBytes - 1) / kTileBytes;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t t0 = tile * kTileBytes;
    const int64_t t1 = min(t0 + (int64_t)kTileBytes, total);
    int64_t pos = t0 + (int64_t)threadIdx.x * 16;
    if (pos >= t1) continue;
    int seg = find_segment(cum, n_seg, pos);
    for (; pos < t1; pos += (int64_t)kThreads * 16) {
      while (cum[seg + 1] <= pos) ++seg;
      // this thread owns packed bytes [pos, pos+16) which may straddle segments
      int64_t p = pos;
      const int64_t pend = min(pos + 16, t1);
      int s = seg;
      while (p < pend) {
        while (cum[s + 1] <= p) ++s;
        const int64_t in_seg = p - cum[s];
        const int64_t run = min(pend, cum[s + 1]) - p;
        const uint8_t* sp = src + src_off[s] + in_seg;
        uint8_t* dp = dst + dst_off[s] + in_seg;
        if (run == 16 && ((reinterpret_cast<uintptr_t>(sp) | reinterpret_cast<uintptr_t>(dp)) & 15) == 0) {
          int4 v = rb::ld_stream(sp);
          if constexpr (kEma) {
            int4 o = *reinterpret_cast<const int4*>(dp);
            __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&v);
            __nv_bfloat162* b = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float2 x = __bfloat1622float2(a[i]), y = __bfloat1622float2(b[i]);
              a[i] = __floats2bfloat162_rn(eta * x.x + (1.f - eta) * y.x, eta * x.y + (1.f - eta) * y.y);
            }
          }
          rb::st_stream(dp, v);
        } else if (((reinterpret_cast<uintptr_t>(sp) | reinterpret_cast<uintptr_t>(dp) | run) & 1) == 0) {
          for (int64_t i = 0; i < run; i += 2) {
            if constexpr (kEma) {
              float x = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(sp + i));
              float y = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(dp + i));
              *reinterpret_cast<__nv_bfloat16*>(dp + i) = __float2bfloat16_rn(eta * x + (1.f - eta) * y);
            } else {
              *reinterpret_cast<uint16_t*>(dp + i) = *reinterpret_cast<const uint16_t*>(sp + i);
            }
          }
        } else {
          for (int64_t i = 0; i < run; ++i) dp[i] = sp[i];  // EMA is only defined for 2-byte elements
        }
        p += run;
      }
    }
  }
}

}  // namespace

extern "C" void rb_segment_copy(const void* src, void* dst, const int64_t* src_off, const int64_t* dst_off,
                                const int64_t* cum, int n_seg, int64_t total_bytes, float eta, int use_ema,
                                cudaStream_t stream) {
  if (total_bytes == 0 || n_seg == 0) return;
  const int64_t n_tiles = (total_bytes + kTileBytes - 1) / kTileBytes;
  const int64_t cap = (int64_t)rb::kNumSMs * 8;
  const int grid = (int)(n_tiles < cap ? n_tiles : cap);
  if (use_ema)
    segcopy_kernel<true><<<grid, kThreads, 0, stream>>>((const uint8_t*)src, (uint8_t*)dst, src_off, dst_off, cum, n_seg,
                                                        total_bytes, eta);
  else
    segcopy_kernel<false><<<grid, kThreads, 0, stream>>>((const uint8_t*)src, (uint8_t*)dst, src_off, dst_off, cum,
                                                         n_seg, total_bytes, eta);
}
