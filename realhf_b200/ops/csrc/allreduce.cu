// Peer-memory collectives over NVLink 5 / NVSwitch for sm_100a: barrier, one-shot and two-shot all-reduce,
// reduce of peer-written partial tiles (the tail of the fused GEMM->reduce-scatter path).
//
// Every rank owns a *symmetric buffer* (same size on all ranks, mapped into every process through CUDA IPC) made of
// a data region and a signal pad.  Kernels take the table of peer base pointers.  Synchronisation is flag based:
// block b of rank r writes an epoch number into slot [b][r] of every peer's pad with a system-scope release store
// and spins with acquire loads on its own pad; epochs come from a per-block counter that lives in the pad and is
// advanced by the kernel itself, so the kernels contain no host-provided step number and are legal inside CUDA
// graphs — what the reference's (disabled) vLLM-derived custom_all_reduce was meant to provide
// (csrc/custom_all_reduce/custom_all_reduce.cuh:128-238, mappings.py:21-27 "may randomly get stuck").
#include "comm_common.cuh"

namespace {
using namespace rbcomm;

template <typename T> struct Acc8 {
  float v[8];
  RB_DEVICE void zero() {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
  RB_DEVICE void add(const int4& x) {
    const T* p = reinterpret_cast<const T*>(&x);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += rb::to_f(p[i]);
  }
  RB_DEVICE int4 pack() const {
    int4 o;
    T* p = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = rb::from_f<T>(v[i]);
    return o;
  }
};
template <> struct Acc8<float> {
  float v[4];
  RB_DEVICE void zero() { v[0] = v[1] = v[2] = v[3] = 0.f; }
  RB_DEVICE void add(const int4& x) {
    const float* p = reinterpret_cast<const float*>(&x);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] += p[i];
  }
  RB_DEVICE int4 pack() const {
    int4 o;
    float* p = reinterpret_cast<float*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = v[i];
    return o;
  }
};

__global__ void __launch_bounds__(kThreads) barrier_kernel(Peers P, int rank, int world) { block_barrier(P, rank, world); }

// One-shot: every rank reads all peers' data regions and reduces.  `in` is first staged into this rank's data region.
template <typename T>
__global__ void __launch_bounds__(kThreads) allreduce_1shot_kernel(Peers P, const int4* __restrict__ in, int4* __restrict__ out,
                                                                   int64_t nvec, int rank, int world) {
  int4* mine = reinterpret_cast<int4*>(P.data[rank]);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = i0; i < nvec; i += stride) mine[i] = in[i];
  block_barrier(P, rank, world);  // all stagings visible (same-index blocks cover the same elements on every rank)
  for (int64_t i = i0; i < nvec; i += stride) {
    Acc8<T> acc;
    acc.zero();
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r) {
      if (r < world) acc.add(reinterpret_cast<const int4*>(P.data[(rank + r) % world])[i]);
    }
    out[i] = acc.pack();
  }
  block_barrier(P, rank, world);  // nobody restages before every peer finished reading
}

// One-shot over data that ALREADY lives in the symmetric buffers (the producing GEMM wrote its partial output there):
// one barrier, then every rank sums the `world` copies at byte offset `off`.  No staging copy and no trailing barrier:
// callers alternate between two regions, and a rank can only re-write region A in call k+2 after passing the barrier
// of call k+1, which every peer reaches only after it finished reading A in call k.  All loads of a thread are issued
// before the first add (the loop is bound by NVLink round trips).
template <typename T, int kUnroll>
__global__ void __launch_bounds__(kThreads) allreduce_symm_kernel(Peers P, int64_t off, int4* __restrict__ out, int64_t nvec, int rank,
                                                                  int world) {
  block_barrier(P, rank, world);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t base = i0; base < nvec; base += stride * kUnroll) {
    int4 v[kUnroll][kMaxRanks];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t i = base + u * stride;
#pragma unroll
      for (int r = 0; r < kMaxRanks; ++r) {
        if (r < world && i < nvec) v[u][r] = reinterpret_cast<const int4*>(P.data[r] + off)[i];  // same order on every rank
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t i = base + u * stride;
      if (i >= nvec) continue;
      Acc8<T> acc;
      acc.zero();
#pragma unroll
      for (int r = 0; r < kMaxRanks; ++r) {
        if (r < world) acc.add(v[u][r]);
      }
      out[i] = acc.pack();
    }
  }
}

// Two-shot: reduce-scatter into this rank's slice (kept in the second half of its data region), barrier, all-gather.
// Block b owns vector range [b*chunk, (b+1)*chunk) in EVERY phase and on EVERY rank, so the same-index block barrier
// is sufficient: block b only ever touches data staged / reduced by the peers' block b.
template <typename T>
__global__ void __launch_bounds__(kThreads) allreduce_2shot_kernel(Peers P, const int4* __restrict__ in, int4* __restrict__ out,
                                                                   int64_t nvec, int64_t half_off_vec, int rank, int world) {
  int4* mine = reinterpret_cast<int4*>(P.data[rank]);
  const int64_t chunk = (nvec + gridDim.x - 1) / gridDim.x;
  const int64_t c0 = min(nvec, (int64_t)blockIdx.x * chunk), c1 = min(nvec, c0 + chunk);
  for (int64_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) mine[i] = in[i];
  block_barrier(P, rank, world);
  const int64_t per = (c1 - c0 + world - 1) / world;
  const int64_t lo = min(c1, c0 + rank * per), hi = min(c1, lo + per);
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    Acc8<T> acc;
    acc.zero();
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r) {
      if (r < world) acc.add(reinterpret_cast<const int4*>(P.data[(rank + r) % world])[i]);
    }
    mine[half_off_vec + i] = acc.pack();
  }
  block_barrier(P, rank, world);
  for (int r = 0; r < world; ++r) {
    const int src = (rank + r) % world;
    const int64_t slo = min(c1, c0 + src * per), shi = min(c1, slo + per);
    const int4* sp = reinterpret_cast<const int4*>(P.data[src]) + half_off_vec;
    for (int64_t i = slo + threadIdx.x; i < shi; i += blockDim.x) out[i] = sp[i];
  }
  block_barrier(P, rank, world);
}

// Sum `world` partial slabs that peers have written into this rank's data region (slab s at offset s*slab_vec).
// Producers bump `counter` once per delivered row-tile; the kernel waits until it reaches calls * per_call, where
// `calls` is a per-block counter kept on the device (so the kernel is replayable inside a CUDA graph).
template <typename T>
__global__ void __launch_bounds__(kThreads) reduce_slabs_kernel(const uint8_t* __restrict__ base, int4* __restrict__ out, int64_t nvec,
                                                                int64_t slab_vec, int world, const uint32_t* __restrict__ counter,
                                                                uint32_t* __restrict__ calls, uint32_t per_call) {
  if (threadIdx.x == 0) {
    const uint32_t c = calls[blockIdx.x] + 1;
    calls[blockIdx.x] = c;
    const uint32_t expect = c * per_call;
    while ((int32_t)(ld_acquire_sys(counter) - expect) < 0) {}
  }
  __syncthreads();
  const int4* b = reinterpret_cast<const int4*>(base);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    Acc8<T> acc;
    acc.zero();
    for (int s = 0; s < world; ++s) acc.add(b[s * slab_vec + i]);
    out[i] = acc.pack();
  }
}

__global__ void spin_wait_kernel(const uint32_t* flag, uint32_t target) {
  while ((int32_t)(ld_acquire_sys(flag) - target) < 0) {}
}

Peers make_peers(const int64_t* data_ptrs, const int64_t* pad_ptrs, int world) {
  Peers P;
  for (int i = 0; i < kMaxRanks; ++i) {
    P.data[i] = i < world ? reinterpret_cast<uint8_t*>(data_ptrs[i]) : nullptr;
    P.pad[i] = i < world ? reinterpret_cast<uint32_t*>(pad_ptrs[i]) : nullptr;
  }
  return P;
}

}  // namespace

extern "C" {

// pad words: [0,64) barrier epochs | [64,576) barrier flags | 576,577 RS arrival counters (parity 0/1) |
//            [640,800) reduce call counters parity 0 | [800,960) parity 1
int rb_symm_pad_words() { return 1024; }
int rb_symm_counter_word() { return kMaxBlocks + kMaxBlocks * kMaxRanks; }
int rb_symm_calls_word(int parity) { return 640 + 160 * parity; }

int rb_symm_barrier(const int64_t* data_ptrs, const int64_t* pad_ptrs, int rank, int world, cudaStream_t s) {
  if (world > kMaxRanks) return -1;
  barrier_kernel<<<1, kThreads, 0, s>>>(make_peers(data_ptrs, pad_ptrs, world), rank, world);
  return 0;
}

// dt: 0 fp32, 1 bf16, 2 fp16.  nbytes multiple of 16.  algo: 1 one-shot, 2 two-shot (needs 2x nbytes of data region).
int rb_symm_allreduce(const int64_t* data_ptrs, const int64_t* pad_ptrs, const void* in, void* out, int64_t nbytes, int rank, int world,
                      int dt, int algo, cudaStream_t s) {
  if (world > kMaxRanks || (nbytes & 15)) return -1;
  Peers P = make_peers(data_ptrs, pad_ptrs, world);
  const int64_t nvec = nbytes / 16;
  int blocks = (int)((nvec + kThreads - 1) / kThreads);
  blocks = blocks < 1 ? 1 : (blocks > 36 ? 36 : blocks);  // few fat blocks: this op is latency / link bound, leave SMs to compute
  const int64_t half = (nvec + 63) / 64 * 64;
  if (algo == 3) {  // `in` lives inside this rank's data region; the same offset is read on every peer
    const int64_t off = reinterpret_cast<const uint8_t*>(in) - P.data[rank];
    if (off < 0 || (off & 15)) return -3;
    int nb = (int)((nvec + 2 * kThreads - 1) / (2 * kThreads));
    nb = nb < 1 ? 1 : (nb > kMaxBlocks ? kMaxBlocks : nb);
#define RB_GO3(T) allreduce_symm_kernel<T, 2><<<nb, kThreads, 0, s>>>(P, off, (int4*)out, nvec, rank, world)
    if (dt == 0) RB_GO3(float); else if (dt == 1) RB_GO3(__nv_bfloat16); else if (dt == 2) RB_GO3(__half); else return -2;
#undef RB_GO3
    return 0;
  }
#define RB_GO(T)                                                                                                              \
  if (algo == 1) allreduce_1shot_kernel<T><<<blocks, kThreads, 0, s>>>(P, (const int4*)in, (int4*)out, nvec, rank, world);       \
  else allreduce_2shot_kernel<T><<<blocks, kThreads, 0, s>>>(P, (const int4*)in, (int4*)out, nvec, half, rank, world);
  if (dt == 0) { RB_GO(float) } else if (dt == 1) { RB_GO(__nv_bfloat16) } else if (dt == 2) { RB_GO(__half) } else return -2;
#undef RB_GO
  return 0;
}

int rb_spin_wait(const uint32_t* flag, uint32_t target, cudaStream_t s) {
  spin_wait_kernel<<<1, 1, 0, s>>>(flag, target);
  return 0;
}

int rb_reduce_slabs(const void* base, void* out, int64_t nbytes, int64_t slab_bytes, int world, const uint32_t* counter, uint32_t* calls,
                    uint32_t per_call, int dt, cudaStream_t s) {
  if ((nbytes & 15) || (slab_bytes & 15)) return -1;
  const int64_t nvec = nbytes / 16;
  int blocks = (int)((nvec + kThreads - 1) / kThreads);
  blocks = blocks < 1 ? 1 : (blocks > rb::kNumSMs ? rb::kNumSMs : blocks);
#define RB_GO(T) reduce_slabs_kernel<T><<<blocks, kThreads, 0, s>>>((const uint8_t*)base, (int4*)out, nvec, slab_bytes / 16, world, counter, calls, per_call)
  if (dt == 0) RB_GO(float); else if (dt == 1) RB_GO(__nv_bfloat16); else if (dt == 2) RB_GO(__half); else return -2;
#undef RB_GO
  return 0;
}

}  // extern "C"
