// Expert-parallel token exchange as peer stores over NVLink / NVSwitch, driven entirely from the device.
//
// The reference has no expert parallelism (modules/moe/token_dispatcher.py:17-27 never leaves the rank) and syncs the host
// once per MoE layer for `tokens_per_expert.cpu()` (moe/experts.py:186).  Here experts are partitioned over the EP group
// and a MoE layer runs
//
//     ep_plan      every rank posts its per-expert (count, start) pairs into the owners' symmetric memory, one barrier,
//                  then derives ALL segment tables on the device: where each of my expert groups lands in its owner's
//                  receive buffer (rows sorted by local expert, then source rank — exactly the layout the grouped tcgen05
//                  GEMM consumes through its device-side prefix sums), and where every received group goes back to
//   ep_move_rows   dispatch: one kernel stores my token rows straight into the owners' receive buffers (16-byte peer
//                  stores, the destination layout is written directly: no pack / unpack, no all-to-all split lists)
//   grouped GEMMs  on the received rows, offsets from the plan (ops/gemm.py::grouped_linear)
//   ep_move_rows   combine: the same kernel with the return table stores every output row back into the source rank's
//                  buffer at the position its input row came from
//
// with no `.tolist()` / `.item()` anywhere: token counts never reach the host.  Barriers are the epoch-flag block barriers of
// comm_common.cuh (graph-capturable).
#include "comm_common.cuh"

namespace {
using namespace rbcomm;

constexpr int kMaxExperts = 256;  // per rank-local table in shared memory

// Symmetric layout (per rank), all offsets in bytes from the data base:
//   post region: int32 [2 parities][world src][e_local][2]  = (count, start) posted by every source rank
struct PlanParams {
  Peers P;
  int64_t post_off;          // byte offset of the post region in every rank's data region
  const int* counts;         // [E] my assignment count per GLOBAL expert (rows of x_sorted are grouped in this order)
  int E, e_local, rank, world, parity;
  // outputs (local device tensors)
  int* send_tab;             // [E][4]            (src_row, n_rows, dst_rank, dst_row)      dispatch table
  int* ret_tab;              // [e_local*world][4] (src_row, n_rows, dst_rank, dst_row)      combine table
  int* per_expert;           // [e_local + 1]     rows per local expert, then total
  int* overflow;             // [1] set to 1 when the rows I must receive exceed `cap_rows`
  int cap_rows;
};

__global__ void __launch_bounds__(kThreads) ep_plan_kernel(PlanParams p) {
  __shared__ int s_start[kMaxExperts + 1];
  const int tid = threadIdx.x;
  const int E = p.E, el = p.e_local, W = p.world;
  // my start offsets (exclusive prefix over global experts) — E <= 256: one thread does the scan
  if (tid == 0) {
    int acc = 0;
    for (int e = 0; e < E; ++e) { s_start[e] = acc; acc += p.counts[e]; }
    s_start[E] = acc;
  }
  __syncthreads();
  // post (count, start) of expert e to its owner
  for (int e = tid; e < E; e += blockDim.x) {
    const int owner = e / el, le = e % el;
    int* post = reinterpret_cast<int*>(p.P.data[owner] + p.post_off) + (((int64_t)p.parity * W + p.rank) * el + le) * 2;
    post[0] = p.counts[e];
    post[1] = s_start[e];
  }
  __threadfence_system();
  block_barrier(p.P, p.rank, W);
  // ---- receiver side: my own post region now holds everybody's counts for my experts
  const int* mine = reinterpret_cast<const int*>(p.P.data[p.rank] + p.post_off) + (int64_t)p.parity * W * el * 2;
  if (tid == 0) {
    int row = 0;
    for (int le = 0; le < el; ++le) {
      int n_e = 0;
      for (int src = 0; src < W; ++src) {
        const int cnt = mine[(src * el + le) * 2], start = mine[(src * el + le) * 2 + 1];
        int* t = p.ret_tab + (le * W + src) * 4;
        t[0] = row; t[1] = cnt; t[2] = src; t[3] = start;
        row += cnt;
        n_e += cnt;
      }
      p.per_expert[le] = n_e;
    }
    p.per_expert[el] = row;
    if (row > p.cap_rows) *p.overflow = 1;
  }
  // ---- sender side: where do my groups land at their owners?  Read the owner's post matrix (peer loads, tiny).
  for (int e = tid; e < E; e += blockDim.x) {
    const int owner = e / el, le = e % el;
    const int* theirs = reinterpret_cast<const int*>(p.P.data[owner] + p.post_off) + (int64_t)p.parity * W * el * 2;
    int row = 0;
    for (int l2 = 0; l2 < le; ++l2)
      for (int src = 0; src < W; ++src) row += theirs[(src * el + l2) * 2];
    for (int src = 0; src < p.rank; ++src) row += theirs[(src * el + le) * 2];
    int* t = p.send_tab + e * 4;
    t[0] = s_start[e]; t[1] = p.counts[e]; t[2] = owner; t[3] = row;
  }
  __threadfence_system();
  block_barrier(p.P, p.rank, W);  // nobody re-posts (next call, other parity region is used anyway) before all tables are built
}

// Segmented row copy into peer buffers.  table[s] = (src_row, n_rows, dst_rank, dst_row); rows are `row_bytes` long (multiple
// of 16).  dst buffers: byte offset `dst_off` in every rank's data region.  Leading barrier: the destination buffers are free
// (their previous consumer finished); trailing barrier: all rows have landed everywhere.
__global__ void __launch_bounds__(kThreads) ep_move_rows_kernel(Peers P, const int* __restrict__ table, int n_seg,
                                                                const uint8_t* __restrict__ src, int64_t src_pitch, int64_t dst_off,
                                                                int64_t dst_pitch, int row_bytes, int dst_cap_rows, int rank, int world) {
  block_barrier(P, rank, world);
  const int nvec = row_bytes / 16;
  for (int s = 0; s < n_seg; ++s) {
    const int src_row = table[s * 4], n = table[s * 4 + 1], dst_rank = table[s * 4 + 2], dst_row = table[s * 4 + 3];
    uint8_t* dst = P.data[dst_rank] + dst_off;
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
      if (dst_row + r >= dst_cap_rows) break;  // overflow is reported by the plan kernel; never write out of bounds
      const int4* a = reinterpret_cast<const int4*>(src + (int64_t)(src_row + r) * src_pitch);
      int4* b = reinterpret_cast<int4*>(dst + (int64_t)(dst_row + r) * dst_pitch);
      for (int i = threadIdx.x; i < nvec; i += blockDim.x) b[i] = a[i];
    }
  }
  __threadfence_system();
  block_barrier(P, rank, world);
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

int rb_ep_plan(const int64_t* data_ptrs, const int64_t* pad_ptrs, int64_t post_off, const int* counts, int E, int e_local, int rank, int world,
               int parity, int* send_tab, int* ret_tab, int* per_expert, int* overflow, int cap_rows, cudaStream_t s) {
  if (world > kMaxRanks || E > kMaxExperts || e_local * world != E) return -1;
  PlanParams p{make_peers(data_ptrs, pad_ptrs, world), post_off, counts, E, e_local, rank, world, parity, send_tab, ret_tab, per_expert,
               overflow, cap_rows};
  ep_plan_kernel<<<1, kThreads, 0, s>>>(p);
  return 0;
}

int rb_ep_move_rows(const int64_t* data_ptrs, const int64_t* pad_ptrs, const int* table, int n_seg, const void* src, int64_t src_pitch,
                    int64_t dst_off, int64_t dst_pitch, int row_bytes, int dst_cap_rows, int rank, int world, int blocks, cudaStream_t s) {
  if (world > kMaxRanks || (row_bytes & 15) || (src_pitch & 15) || (dst_pitch & 15) || (dst_off & 15)) return -1;
  if (blocks < 1) blocks = 1;
  if (blocks > kMaxBlocks) blocks = kMaxBlocks;
  ep_move_rows_kernel<<<blocks, kThreads, 0, s>>>(make_peers(data_ptrs, pad_ptrs, world), table, n_seg, (const uint8_t*)src, src_pitch, dst_off,
                                                  dst_pitch, row_bytes, dst_cap_rows, rank, world);
  return 0;
}

}  // extern "C"
