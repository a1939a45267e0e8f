// Shared pieces of the peer-memory collectives: the peer-pointer table, system-scope flag accesses and the
// same-index-block barrier across ranks (see allreduce.cu for the protocol description).
#pragma once
#include "common.cuh"

namespace rbcomm {

constexpr int kMaxRanks = 8;
constexpr int kMaxBlocks = 64;
constexpr int kThreads = 512;

// Layout of the signal pad (uint32 words):
//   [0, kMaxBlocks)                               per-block epoch counters (local use only)
//   [kMaxBlocks, kMaxBlocks + kMaxBlocks*kMaxRanks)  flags[block][src_rank]
struct Peers {
  uint8_t* data[kMaxRanks];
  uint32_t* pad[kMaxRanks];
};

RB_DEVICE void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
RB_DEVICE uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All blocks with the same blockIdx on all ranks rendezvous.  Must be called by every thread of the block.
RB_DEVICE void block_barrier(const Peers& P, int rank, int world) {
  __syncthreads();
  __shared__ uint32_t epoch_s;
  uint32_t* my_pad = P.pad[rank];
  if (threadIdx.x == 0) {
    epoch_s = my_pad[blockIdx.x] + 1;
    my_pad[blockIdx.x] = epoch_s;
  }
  __syncthreads();
  const uint32_t epoch = epoch_s;
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(P.pad[threadIdx.x] + kMaxBlocks + blockIdx.x * kMaxRanks + rank, epoch);
    const uint32_t* flag = my_pad + kMaxBlocks + blockIdx.x * kMaxRanks + threadIdx.x;
    // epochs only grow; a peer may already be one barrier ahead
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {}
  }
  __syncthreads();
}


}  // namespace rbcomm
