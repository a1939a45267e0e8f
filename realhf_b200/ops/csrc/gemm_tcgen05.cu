// Persistent warp-specialised GEMM for sm_100a: TMA -> shared (128B swizzle) -> tcgen05.mma -> TMEM -> epilogue.
//
//   D[M,N] (+)= A x B (+ bias)      bf16/fp16 operands, fp32 accumulation in tensor memory.
//
// Both operands may be K-major or MN-major, which covers the three GEMMs of a linear layer without
// any transposed copy (reference: cuBLAS via torch.matmul, parallelism/model_parallel/modules.py:302,382,532):
//   forward  y  = x  @ W^T : A = x  [M,K]  K-major,   B = W [N,K]            K-major
//   dgrad    dx = dy @ W   : A = dy [M,N'] K-major,   B = W [N',K'] as [K,N] MN-major
//   wgrad    dW = dy^T @ x : A = dy [T,N'] as [K,M]   MN-major, B = x [T,K'] as [K,N] MN-major (fp32 accumulate-into)
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 MMA issuer (one elected lane issues tcgen05.mma) + TMEM
// allocator, warps 2-5 epilogue (each owns the TMEM lane quadrant warp_id%4).  Three pipelines:
// smem full/empty ring (TMA <-> MMA), TMEM full/empty double buffer (MMA <-> epilogue), and a static
// persistent tile loop (one CTA per SM).  The epilogue converts and stores straight from TMEM
// registers with 16-byte stores, with optional bias and accumulate-into-C.
#include "gemm_common.cuh"

namespace {

template <int BN> struct Cfg {
  static constexpr int kStages = BN >= 256 ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// Fused tensor-parallel modes (kMode): 0 plain; 1 GEMM -> reduce-scatter: the epilogue stores every output row straight
// into the inbox slab of the rank that owns it (peer memory over NVLink) and bumps that rank's arrival counter;
// 2 all-gather -> GEMM: A row-blocks are TMA-loaded directly from the owning rank's symmetric buffer.
constexpr int kMaxTP = 8;
struct FusedParams {
  int64_t peer_base[kMaxTP];     // mode 1: byte address of this call's inbox on every rank
  int64_t peer_counter[kMaxTP];  // mode 1: byte address of the arrival counter on every rank
  int64_t slab_elems;            // mode 1: rows_per_rank * N
  int rows_per_rank, my_rank, world;
  const int* group_offsets;      // mode 3 (grouped / MoE): device prefix sums of rows per group, [num_groups + 1]
  int num_groups;
  int group_b_rows;              // mode 3: rows of the stacked B tensor that belong to one group
};
constexpr int kMaxGroups = 256;
struct TmaArray { CUtensorMap m[kMaxTP]; };

RB_DEVICE void red_add_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// L2-friendly rasterisation: tiles are walked in bands of kGroupM m-tiles, m fastest inside a band, so the
// ~148 tiles in flight cover a roughly square region of C and every A / B panel is fetched from HBM once.
constexpr int kGroupM = 16;
RB_DEVICE void tile_coords(int tile, int tiles_m, int tiles_n, int bn, int& m0, int& n0) {
  const int per_group = kGroupM * tiles_n;
  const int group = tile / per_group, within = tile - group * per_group;
  const int gm0 = group * kGroupM;
  const int gsize = min(kGroupM, tiles_m - gm0);
  m0 = (gm0 + within % gsize) * BM;
  n0 = (within / gsize) * bn;
}

// kMC > 1: thread-block cluster of kMC CTAs working on kMC adjacent n-tiles of the SAME m-tile (small-M / decode
// shapes).  Every CTA loads 1/kMC of the shared A tile and TMA-multicasts it to the whole cluster, so A is fetched
// from L2 once per cluster instead of once per CTA; a smem stage is recycled only after all kMC consumers released it
// (tcgen05.commit multicast onto every CTA's empty barrier).
template <int BN, bool kAMN, bool kBMN, typename OutT, int kFmt, int kMC, int kMode = 0>
__global__ void __launch_bounds__(kThreads, 1) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a,
                                                                   const __grid_constant__ CUtensorMap tma_b, Params p,
                                                                   const __grid_constant__ FusedParams fp,
                                                                   const __grid_constant__ TmaArray tma_a_peers) {
  static_assert(kMC == 1 || !kAMN, "A multicast is implemented for K-major A");
  static_assert(kMode == 0 || (kMC == 1 && !kAMN), "fused TP modes: K-major A, no cluster");
  using C = Cfg<BN>;
  // mode 3: tiles of group g are [s_tile_prefix[g], s_tile_prefix[g+1]); inside a group n-tiles are outer, m-tiles inner
  __shared__ int s_tile_prefix[kMode == 3 ? kMaxGroups + 1 : 1];
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = RB_CEIL_DIV(p.M, BM), tiles_n = RB_CEIL_DIV(p.N, BN);
  if constexpr (kMode == 3) {
    if (threadIdx.x == 0) {
      int acc = 0;
      s_tile_prefix[0] = 0;
      for (int g = 0; g < fp.num_groups; ++g) {
        acc += RB_CEIL_DIV(fp.group_offsets[g + 1] - fp.group_offsets[g], BM) * tiles_n;
        s_tile_prefix[g + 1] = acc;
      }
    }
    __syncthreads();
  }
  // with multicast every CTA of a cluster runs the same number of iterations (padded tiles load zeros, store nothing)
  const int num_tiles = kMode == 3 ? s_tile_prefix[fp.num_groups]
                                   : (kMC == 1 ? tiles_m * tiles_n : RB_CEIL_DIV(tiles_m * tiles_n, kMC) * kMC);
  // grouped tile -> (row start, row end, n0, B row offset)
  auto group_tile = [&](int tile, int& m0, int& m_end, int& n0, int& b_row0) {
    int g = 0;
    while (g + 1 < fp.num_groups && s_tile_prefix[g + 1] <= tile) ++g;
    const int r0 = fp.group_offsets[g], r1 = fp.group_offsets[g + 1];
    const int mt = RB_CEIL_DIV(r1 - r0, BM);
    const int local = tile - s_tile_prefix[g];
    m0 = r0 + (local % mt) * BM;
    m_end = r1;
    n0 = (local / mt) * BN;
    b_row0 = g * fp.group_b_rows;
  };
  const int num_kb = RB_CEIL_DIV(p.K, BK);
  const uint32_t cta_rank = kMC == 1 ? 0u : (blockIdx.x % kMC);
  constexpr uint16_t kMcMask = (uint16_t)((1u << kMC) - 1u);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tma_a);
    ptx::prefetch_tensormap(&tma_b);
    for (int i = 0; i < C::kStages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[i]), kMC);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(&tmem_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&tmem_empty[i]), 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_ptr), C::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (kMC > 1) ptx::cluster_sync();  // peers' barriers must be initialised before any remote arrive / multicast
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m0, n0, m_end = 0, b_row0 = 0;
        if constexpr (kMode == 3) group_tile(tile, m0, m_end, n0, b_row0);
        else tile_coords(tile, tiles_m, tiles_n, BN, m0, n0);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
          ptx::mbar_arrive_expect_tx(fb, C::kStageBytes);
          const uint32_t sa = ptx::smem_u32(smem_a + stage * C::kABytes);
          const uint32_t sb = ptx::smem_u32(smem_b + stage * C::kBBytes);
          const int k0 = kb * BK;
          if constexpr (kAMN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) ptx::tma_load_2d(sa + j * (BK * 128), &tma_a, fb, m0 + 64 * j, k0);
          } else if constexpr (kMC > 1) {
            constexpr int kRows = BM / kMC;  // my slice of the A tile, delivered to every CTA of the cluster
            ptx::tma_load_2d_mcast(sa + cta_rank * (kRows * 128), &tma_a, fb, k0, m0 + (int)cta_rank * kRows, kMcMask);
          } else if constexpr (kMode == 2) {
            const int src = m0 / fp.rows_per_rank;  // rows_per_rank % BM == 0 (checked on the host)
            ptx::tma_load_2d(sa, &tma_a_peers.m[src], fb, k0, m0 - src * fp.rows_per_rank);
          } else {
            ptx::tma_load_2d(sa, &tma_a, fb, k0, m0);
          }
          if constexpr (kBMN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) ptx::tma_load_2d(sb + j * (BK * 128), &tma_b, fb, n0 + 64 * j, k0 + b_row0);
          } else {
            ptx::tma_load_2d(sb, &tma_b, fb, k0, n0 + b_row0);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(kFmt, BM, BN, kAMN ? 1 : 0, kBMN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(ptx::smem_u32(&tmem_empty[as]), aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem_a + stage * C::kABytes);
          const uint32_t sb = ptx::smem_u32(smem_b + stage * C::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = kAMN ? ptx::make_smem_desc_sw128(sa + k * 2048, BK * 128, 1024)
                                        : ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bdesc = kBMN ? ptx::make_smem_desc_sw128(sb + k * 2048, BK * 128, 1024)
                                        : ptx::make_smem_desc_sw128(sb + k * 32, 16, 1024);
            ptx::tc_mma_f16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (kMC > 1) ptx::tc_commit_mcast(ptx::smem_u32(&empty_bar[stage]), kMcMask);
          else ptx::tc_commit(ptx::smem_u32(&empty_bar[stage]));  // smem slot reusable once these MMAs retire
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        ptx::tc_commit(ptx::smem_u32(&tmem_full[as]));  // accumulator complete -> epilogue
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================================================== epilogue (4 warps, TMEM lane quadrant = warp % 4)
    const int quad = warp & 3;
    int as = 0;
    uint32_t aphase = 0;
    OutT* Cp = reinterpret_cast<OutT*>(p.C);
    const OutT* bias = reinterpret_cast<const OutT*>(p.bias);
    const bool vec_ok = kMode == 1 ? (p.N % (16 / sizeof(OutT)) == 0)
                                   : ((p.ldc % (16 / sizeof(OutT)) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0));
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m0, n0, m_end = p.M, b_row0 = 0;
      if constexpr (kMode == 3) group_tile(tile, m0, m_end, n0, b_row0);
      else tile_coords(tile, tiles_m, tiles_n, BN, m0, n0);
      ptx::mbar_wait(ptx::smem_u32(&tmem_full[as]), aphase);
      ptx::tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        ptx::tc_ld_32x32(taddr + c * 32, r);
        ptx::tc_wait_ld();
        const int col = n0 + c * 32;
        const int n_valid = min(32, p.N - col);
        if (row < m_end && n_valid > 0) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < n_valid) v[i] += rb::to_f(bias[col + i]);
          }
          OutT* dst = Cp + (int64_t)row * p.ldc + col;
          if constexpr (kMode == 1) {
            const int d = row / fp.rows_per_rank;
            dst = reinterpret_cast<OutT*>(fp.peer_base[d]) + (int64_t)fp.my_rank * fp.slab_elems +
                  (int64_t)(row - d * fp.rows_per_rank) * p.N + col;
          }
          if (p.accumulate) {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < n_valid) v[i] += rb::to_f(dst[i]);
          }
          store_chunk<OutT>(dst, v, n_valid, vec_ok);
        }
      }
      if constexpr (kMode == 1) {
        // publish this tile's rows: one counter bump per destination (warp-aggregated when the 32 rows share a rank)
        __threadfence_system();
        const bool valid = row < p.M && n0 < p.N;
        const int d = valid ? row / fp.rows_per_rank : -1;
        const int d0 = __shfl_sync(0xffffffffu, d, 0);
        const unsigned same = __ballot_sync(0xffffffffu, d == d0);
        if (same == 0xffffffffu) {
          if (lane == 0 && d0 >= 0) red_add_release_sys(reinterpret_cast<uint32_t*>(fp.peer_counter[d0]), 32u);
        } else if (valid) {
          red_add_release_sys(reinterpret_cast<uint32_t*>(fp.peer_counter[d]), 1u);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tmem_empty[as]));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (kMC > 1) ptx::cluster_sync();  // no CTA may exit while peers still multicast into / arrive on its smem
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}


// ---------------------------------------------------------------------------------------------- stream-K (decode shapes)
// M <= 128 (one m-tile): the GEMM streams the weight matrix once and is bound by how evenly the weight bytes are
// spread over the SMs (each SM ingests ~46-57 B/clk through TMA), not by the tensor cores.  Tiling only over N gives
// tiles_n CTAs (16 for a 4096-wide projection with BN=256) and every CTA re-reads the whole activation panel.  Stream-K
// instead cuts the flattened (n-tile, k-block) iteration space into gridDim.x equal contiguous ranges, so every SM
// streams the same number of bytes whatever N and K are.  A range that covers only part of a tile produces a partial
// accumulator: it is written as fp32 to one of the CTA's two workspace slots, the CTA bumps the tile's arrival counter,
// and once all S CTAs that touched the tile have arrived each of them sums and stores 1/S of the tile's rows (a
// serial "owner adds everything" fix-up measured 4-9x slower: 128 threads chasing S x 128 KB through L2).  While the
// epilogue warps wait, the TMA / MMA warps already run the CTA's next segment (TMEM is double-buffered).  A CTA only
// ever waits on tiles at or before its own, so with all CTAs resident (grid <= #SMs, 1 CTA/SM) the scheme cannot
// deadlock.  The last reader re-arms the tile counters, so no per-call memset is needed and CUDA graphs can replay it.  A is loaded with a box of only round_up(M, 8) rows; the remaining rows
// of the 128-row UMMA tile hold stale shared memory and only ever influence accumulator rows that are never stored.
struct StreamKParams {
  float* ws;         // [2 * gridDim.x][BM][BN] fp32 partials: slot 2g for CTA g's first segment, 2g+1 for its last
  uint32_t* flags;   // [2 * tiles_n]: (arrived, done) per tile, zero at entry and exit
  int a_box_rows;
  int bn, stages;
  long long* dbg;    // optional [gridDim.x][8] globaltimer stamps (profiling builds of the benchmark only)
  // fp8 (e4m3) operands: out[m, n] = acc[m, n] * scale_a[m] * scale_b[n] (per-token activation scale, per-output-channel weight
  // scale); nullptr for 16-bit operands
  const float* scale_a;
  const float* scale_b;
};

RB_DEVICE uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
RB_DEVICE long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
RB_DEVICE void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// kFmt: 0 fp16, 1 bf16, 2 fp8 e4m3 (W8A8 generation path: same byte geometry -- a k-block is one 128-byte swizzle row, i.e.
// 128 fp8 elements instead of 64 bf16 -- `tcgen05.mma.kind::f8f6f4`, per-row x per-column scales in the epilogue)
template <typename OutT, int kFmt>
__global__ void __launch_bounds__(kThreads, 1) gemm_streamk_kernel(const __grid_constant__ CUtensorMap tma_a,
                                                                   const __grid_constant__ CUtensorMap tma_b, Params p,
                                                                   StreamKParams sk) {
  // tile width and pipeline depth are run-time values here: the host picks BN (any multiple of 16 up to 256) so that
  // tiles_n x split lands just under the SM count, which matters more for these shapes than compile-time unrolling
  const int BN = sk.bn, n_stages = sk.stages;
  // A stages are packed to the rows that TMA actually writes (round_up(M, 8) x 128 B).  The MMA still addresses 128 rows
  // from each stage base, so it reads on into the following stages / the B region (always inside the allocation, see
  // the host-side layout): those rows only feed accumulator rows >= M, which are never stored.  With M = 16 a stage
  // shrinks from 28 KB to 14 KB and twice as many loads are in flight per SM, which is what bounds these shapes.
  const int kABytes = sk.a_box_rows * 128;
  const int b_bytes = BN * BK * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + n_stages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + n_stages * b_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + n_stages;
  uint64_t* tmem_full = bars + 2 * n_stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  __shared__ int slot_of[160];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int KBE = kFmt == 2 ? 2 * BK : BK;  // elements per k-block (128 bytes either way)
  const int num_kb = RB_CEIL_DIV(p.K, KBE);
  const int64_t units = (int64_t)RB_CEIL_DIV(p.N, BN) * num_kb;
  const int G = gridDim.x, g = blockIdx.x;
  const int u_begin = (int)(units * g / G), u_end = (int)(units * (g + 1) / G);
  const uint32_t stage_tx = (uint32_t)sk.a_box_rows * 128u + (uint32_t)b_bytes;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tma_a);
    ptx::prefetch_tensormap(&tma_b);
    for (int i = 0; i < n_stages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(&tmem_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&tmem_empty[i]), 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_ptr), 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // Programmatic dependent launch: the prologue above (barrier init, TMEM allocation, descriptor prefetch) already
  // overlapped the previous kernel's tail.  The weight operand B does not depend on the predecessor either, so the
  // producer below fills the whole smem ring with weight tiles BEFORE `griddepcontrol.wait` -- HBM keeps streaming
  // weights while the preceding norm / activation / attention kernel drains -- and only the activation loads (A), the
  // output stores and the stream-K workspace wait for the predecessor.
  rb::pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      // first pass over the ring: weights now, activations after the dependency wait
      const int n_pre = min(n_stages, u_end - u_begin);
      for (int i = 0; i < n_pre; ++i) {
        const int u = u_begin + i;
        const int tile = u / num_kb, kb = u - tile * num_kb;
        const uint32_t fb = ptx::smem_u32(&full_bar[i]);
        ptx::mbar_arrive_expect_tx(fb, stage_tx);
        ptx::tma_load_2d(ptx::smem_u32(smem_b + i * b_bytes), &tma_b, fb, kb * KBE, tile * BN);
      }
      rb::pdl_wait();
      for (int i = 0; i < n_pre; ++i) {
        const int u = u_begin + i;
        const int kb = u - (u / num_kb) * num_kb;
        ptx::tma_load_2d(ptx::smem_u32(smem_a + i * kABytes), &tma_a, ptx::smem_u32(&full_bar[i]), kb * KBE, 0);
      }
      int stage = n_pre == n_stages ? 0 : n_pre;
      uint32_t phase = n_pre == n_stages ? 1 : 0;
      for (int u = u_begin + n_pre; u < u_end; ++u) {
        const int tile = u / num_kb, kb = u - tile * num_kb;
        ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
        ptx::mbar_arrive_expect_tx(fb, stage_tx);
        ptx::tma_load_2d(ptx::smem_u32(smem_a + stage * kABytes), &tma_a, fb, kb * KBE, 0);
        ptx::tma_load_2d(ptx::smem_u32(smem_b + stage * b_bytes), &tma_b, fb, kb * KBE, tile * BN);
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = kFmt == 2 ? ptx::make_idesc_f8(0, 0, BM, BN) : ptx::make_idesc_f16(kFmt, BM, BN, 0, 0);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      int u = u_begin;
      while (u < u_end) {
        const int tile = u / num_kb;
        const int seg_end = min(u_end, (tile + 1) * num_kb);
        ptx::mbar_wait(ptx::smem_u32(&tmem_empty[as]), aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int first = 1; u < seg_end; ++u, first = 0) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem_a + stage * kABytes);
          const uint32_t sb = ptx::smem_u32(smem_b + stage * b_bytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {  // 4 instructions of 32 bytes of K each: 16 bf16 or 32 fp8 elements
            const uint64_t da = ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024), db = ptx::make_smem_desc_sw128(sb + k * 32, 16, 1024);
            if constexpr (kFmt == 2) ptx::tc_mma_f8(d_tmem, da, db, idesc, (first && k == 0) ? 0u : 1u);
            else ptx::tc_mma_f16(d_tmem, da, db, idesc, (first && k == 0) ? 0u : 1u);
          }
          ptx::tc_commit(ptx::smem_u32(&empty_bar[stage]));
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
        ptx::tc_commit(ptx::smem_u32(&tmem_full[as]));
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const bool row_ok = row < p.M;
    const int et = threadIdx.x - 64;  // 0..127 over the epilogue warps
    rb::pdl_wait();  // output buffer, bias and the stream-K workspace may still be in use by the predecessor
    int as = 0;
    uint32_t aphase = 0;
    OutT* Cp = reinterpret_cast<OutT*>(p.C);
    const OutT* bias = reinterpret_cast<const OutT*>(p.bias);
    const bool vec_ok = (p.ldc % (16 / sizeof(OutT)) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    auto cta_of = [&](int unit) { return (int)((((int64_t)unit + 1) * G - 1) / units); };
    if (sk.dbg && et == 0) sk.dbg[g * 8 + 0] = gtimer();
    int pending[2], n_pending = 0;  // a range has at most two partial segments: its first and its last
    int u = u_begin;
    while (u < u_end) {
      const int tile = u / num_kb;
      const int tile_u0 = tile * num_kb, tile_u1 = tile_u0 + num_kb;
      const int seg_end = min(u_end, tile_u1);
      const bool whole = u == tile_u0 && seg_end == tile_u1;
      const int n0 = tile * BN;
      ptx::mbar_wait(ptx::smem_u32(&tmem_full[as]), aphase);
      ptx::tc_fence_after();
      if (sk.dbg && et == 0 && u == u_begin) sk.dbg[g * 8 + 1] = gtimer();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * 256;
      float* my_slot = sk.ws + (int64_t)(2 * g + (u == u_begin ? 0 : 1)) * (BM * 256);
      const int n_end = min(p.N, n0 + BN);
#pragma unroll 1
      for (int c = 0; c < (BN + 31) / 32; ++c) {
        uint32_t r[32];
        ptx::tc_ld_32x32(taddr + c * 32, r);  // a 16-column tail reads past the accumulator (inside the allocation); masked below
        ptx::tc_wait_ld();
        const int col = n0 + c * 32;
        const int n_valid = min(32, n_end - col);
        if (!row_ok || n_valid <= 0) continue;
        if (!whole) {  // fp32 partial -> my workspace slot, laid out [col/4][row] in float4 units: a warp stores 512 contiguous bytes
          float4* dst = reinterpret_cast<float4*>(my_slot) + (c * 8) * BM + row;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            dst[i * BM] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                      __uint_as_float(r[4 * i + 3]));
          continue;
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        if constexpr (kFmt == 2) {
          const float sa = sk.scale_a[row];
#pragma unroll
          for (int i = 0; i < 32; ++i) if (i < n_valid) v[i] *= sa * sk.scale_b[col + i];
        }
        if (bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (i < n_valid) v[i] += rb::to_f(bias[col + i]);
        }
        store_chunk<OutT>(Cp + (int64_t)row * p.ldc + col, v, n_valid, vec_ok);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tmem_empty[as]));  // the MMA warp may start the next segment now
      if (++as == 2) { as = 0; aphase ^= 1; }
      if (!whole) {
        epi_bar_sync();  // all four quadrants of my partial are written; the release below is cumulative over them
        if (et == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(sk.flags + 2 * tile), "r"(1u) : "memory");
        pending[n_pending++] = tile;
      }
      u = seg_end;
    }
    if (sk.dbg && et == 0) sk.dbg[g * 8 + 2] = gtimer();
    // ---- cooperative fix-up, deferred until every partial of this CTA is posted (waiting inside the loop would chain
    // tile t's fix-up behind tile t-1's through the CTA that spans both): every CTA that touched a tile reduces 1/S of it
    for (int pi = 0; pi < n_pending; ++pi) {
      const int tile = pending[pi];
      const int tile_u0 = tile * num_kb, tile_u1 = tile_u0 + num_kb;
      const int n0 = tile * BN;
      const int g_first = cta_of(tile_u0), g_last = cta_of(tile_u1 - 1);
      const int S = g_last - g_first + 1, j = g - g_first;
      uint32_t* arrived = sk.flags + 2 * tile;
      if (et == 0) {
        while (ld_acquire_gpu(arrived) < (uint32_t)S) {}
      }
      epi_bar_sync();
      if (sk.dbg && et == 0) sk.dbg[g * 8 + 3 + 2 * pi] = gtimer();
      // member j sums columns [c_lo, c_hi) (float4 units) of all rows.  A warp works on blocks of 8 rows x 4 float4-columns:
      // its loads are four full 128-byte lines per member and its stores eight full 32-byte sectors.  Four blocks x eight
      // members are in flight per thread, because this phase is bound by L2 round trips, not bytes.
      for (int m = et; m < S; m += 128) {
        const int gm = g_first + m;
        slot_of[m] = 2 * gm + (((int)(units * gm / G) / num_kb) == tile ? 0 : 1);
      }
      epi_bar_sync();
      const int kC4 = BN / 4;
      const int c_lo = kC4 * j / S, c_hi = kC4 * (j + 1) / S;
      const int n_rb = RB_CEIL_DIV(p.M, 8), n_cb = RB_CEIL_DIV(c_hi - c_lo, 4);
      const int n_blocks = n_rb * n_cb;
      const bool vec4 = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 7) == 0);
      const float4* ws4 = reinterpret_cast<const float4*>(sk.ws);
      for (int bb = quad; bb < n_blocks; bb += 16) {
        float4 acc[4];
        int off[4];  // float4 offset inside a slot, or -1
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int blk = bb + 4 * q;
          const int rr = (blk % n_rb) * 8 + (lane & 7), c4 = c_lo + (blk / n_rb) * 4 + (lane >> 3);
          off[q] = (blk < n_blocks && rr < p.M && c4 < c_hi && n0 + c4 * 4 < p.N) ? c4 * BM + rr : -1;
        }
        for (int m0 = 0; m0 < S; m0 += 8) {
          float4 t[8][4];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const bool mk = m0 + k < S;
            const float4* slot = ws4 + (int64_t)slot_of[mk ? m0 + k : 0] * (BM * 256 / 4);
#pragma unroll
            for (int q = 0; q < 4; ++q) t[k][q] = (mk && off[q] >= 0) ? __ldcg(slot + off[q]) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { acc[q].x += t[k][q].x; acc[q].y += t[k][q].y; acc[q].z += t[k][q].z; acc[q].w += t[k][q].w; }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (off[q] < 0) continue;
          const int orow = off[q] % BM, col = n0 + (off[q] / BM) * 4;
          float v[4] = {acc[q].x, acc[q].y, acc[q].z, acc[q].w};
          const int nv = min(4, p.N - col);
          if constexpr (kFmt == 2) {
            const float sa = sk.scale_a[orow];
            for (int i = 0; i < nv; ++i) v[i] *= sa * sk.scale_b[col + i];
          }
          if (bias != nullptr) {
            for (int i = 0; i < nv; ++i) v[i] += rb::to_f(bias[col + i]);
          }
          OutT* dst = Cp + (int64_t)orow * p.ldc + col;
          if (nv == 4 && vec4) {
            rb::Pack<OutT, 4> pk;
#pragma unroll
            for (int i = 0; i < 4; ++i) pk.v[i] = rb::from_f<OutT>(v[i]);
            *reinterpret_cast<rb::Pack<OutT, 4>*>(dst) = pk;
          } else {
            for (int i = 0; i < nv; ++i) dst[i] = rb::from_f<OutT>(v[i]);
          }
        }
      }
      epi_bar_sync();  // every thread of this CTA is done reading the group's slots
      if (sk.dbg && et == 0) sk.dbg[g * 8 + 4 + 2 * pi] = gtimer();
      if (et == 0) {
        const uint32_t old = atomicAdd(arrived + 1, 1u);
        if (old == (uint32_t)S - 1u) { arrived[1] = 0u; __threadfence(); arrived[0] = 0u; }  // last reader re-arms the tile
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------- host side
static const FusedParams kNoFused{};
static const TmaArray kNoPeers{};

template <int BN, bool kAMN, bool kBMN, typename OutT, int kFmt, int kMC = 1, int kMode = 0>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, int num_sms, cudaStream_t s,
           const FusedParams& fp = kNoFused, const TmaArray& peers = kNoPeers) {
  auto kern = gemm_tcgen05_kernel<BN, kAMN, kBMN, OutT, kFmt, kMC, kMode>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::kSmemBytes) != cudaSuccess) return -2;
    configured = true;
  }
  const int tiles = RB_CEIL_DIV(p.M, BM) * RB_CEIL_DIV(p.N, BN);
  if constexpr (kMC == 1) {
    const int grid = tiles < num_sms ? tiles : num_sms;
    kern<<<grid, kThreads, Cfg<BN>::kSmemBytes, s>>>(ta, tb, p, fp, peers);
  } else {
    const int padded = RB_CEIL_DIV(tiles, kMC) * kMC;
    const int cap = (num_sms / kMC) * kMC;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(padded < cap ? padded : cap);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg<BN>::kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kMC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, ta, tb, p, fp, peers) != cudaSuccess) return -3;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}


template <typename OutT, int kFmt>
int launch_streamk(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, const StreamKParams& sk, int grid, cudaStream_t s) {
  auto kern = gemm_streamk_kernel<OutT, kFmt>;
  constexpr int kMaxSmem = 227 * 1024 - 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess) return -2;
    configured = true;
  }
  const int smem = sk.stages * (sk.a_box_rows * 128 + sk.bn * BK * 2) + 1024 + 512;
  if (rb::launch_pdl(kern, dim3(grid), dim3(kThreads), (size_t)smem, s, ta, tb, p, sk) != cudaSuccess) return -3;
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

template <int BN, typename OutT, int kFmt>
int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, int sms, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch<BN, false, false, OutT, kFmt>(ta, tb, p, sms, s);
  if constexpr (BN >= 64) {
    if (!a_mn && b_mn) return launch<BN, false, true, OutT, kFmt>(ta, tb, p, sms, s);
    if (a_mn && b_mn) return launch<BN, true, true, OutT, kFmt>(ta, tb, p, sms, s);
    if (a_mn && !b_mn) return launch<BN, true, false, OutT, kFmt>(ta, tb, p, sms, s);
  } else {
    if (a_mn && !b_mn) return launch<BN, true, false, OutT, kFmt>(ta, tb, p, sms, s);
  }
  return -4;
}

template <typename OutT, int kFmt>
int dispatch_bn(int bn, bool a_mn, bool b_mn, int mc, const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, int sms, cudaStream_t s) {
  if (mc == 4) {  // small-M K-major shapes only (checked by the caller)
    if (bn == 32) return launch<32, false, false, OutT, kFmt, 4>(ta, tb, p, sms, s);
    if (bn == 64) return launch<64, false, false, OutT, kFmt, 4>(ta, tb, p, sms, s);
    return -6;
  }
  switch (bn) {
    case 256: return dispatch_major<256, OutT, kFmt>(a_mn, b_mn, ta, tb, p, sms, s);
    case 128: return dispatch_major<128, OutT, kFmt>(a_mn, b_mn, ta, tb, p, sms, s);
    case 64: return dispatch_major<64, OutT, kFmt>(a_mn, b_mn, ta, tb, p, sms, s);
    case 32: return dispatch_major<32, OutT, kFmt>(a_mn, b_mn, ta, tb, p, sms, s);
    default: return -5;
  }
}

}  // namespace

extern "C" int rb_gemm_2cta(const void* A, const void* B, void* C, const void* bias, int M, int N, int K, int64_t lda, int64_t ldb,
                            int64_t ldc, int a_mn, int b_mn, int in_dt, int out_dt, int accumulate, int bn, int num_sms, cudaStream_t s);

extern "C" {

// in_dt: 1 bf16, 2 fp16.  out_dt: 0 fp32, 1 bf16, 2 fp16.
// A: a_mn ? [K, M] : [M, K] with row pitch lda;  B: b_mn ? [K, N] : [N, K] with row pitch ldb.
// bn = 0 picks the tile width from the problem size.
int rb_gemm_tcgen05(const void* A, const void* B, void* C, const void* bias, int M, int N, int K, int64_t lda, int64_t ldb,
                    int64_t ldc, int a_mn, int b_mn, int in_dt, int out_dt, int accumulate, int bn, int num_sms, int mc_req, cudaStream_t s) {
  static const int mc_env = [] { const char* e = getenv("REAL_GEMM_MULTICAST"); return e ? atoi(e) : 0; }();
  static const int cta2_env = [] { const char* e = getenv("REAL_GEMM_2CTA"); return e ? atoi(e) : 1; }();
  // CTA-pair kernel (gemm_2cta.cu) for everything with at least two m-tiles: mc_req == 2 forces it, -1 follows the env default
  if ((mc_req == 2 || (mc_req < 0 && cta2_env != 0)) && M > BM && (bn == 0 || bn == 128 || bn == 256)) {
    const int rc = rb_gemm_2cta(A, B, C, bias, M, N, K, lda, ldb, ldc, a_mn, b_mn, in_dt, out_dt, accumulate, bn, num_sms, s);
    if (rc != -40) return rc;
  }
  if (mc_req == 2) mc_req = 0;
  const int mc_mode = mc_req >= 0 ? mc_req : mc_env;  // TMA-multicast clusters for small M: opt-in (measured slower than plain tiles so far)
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (in_dt != 1 && in_dt != 2) return -10;
  if ((lda % 8) || (ldb % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return -11;
  if (num_sms <= 0) num_sms = rb::kNumSMs;
  if (bn == 0 && M <= BM && !a_mn && !b_mn && mc_mode != 0 && N >= 256) {
    bn = 32;  // decode shapes: many small tiles balance the 148 SMs; A traffic is shared by TMA multicast
  }
  if (bn == 0) {
    const int tm = RB_CEIL_DIV(M, BM);
    const int cands[4] = {256, 128, 64, 32};
    bn = b_mn ? 64 : 32;
    for (int i = 0; i < 4; ++i) {
      if (b_mn && cands[i] < 64) continue;
      if ((int64_t)tm * RB_CEIL_DIV(N, cands[i]) >= num_sms || cands[i] <= (b_mn ? 64 : 32)) { bn = cands[i]; break; }
    }
    while (bn > N && bn > (b_mn ? 64 : 32)) bn >>= 1;
  }
  if (b_mn && bn < 64) return -12;
  // decode-shaped problems (one m-tile, K-major operands): cluster of 4 CTAs shares the A tile by TMA multicast
  int mc = 1;
  if (M <= BM && !a_mn && !b_mn && bn <= 64 && RB_CEIL_DIV(N, bn) >= 8 && mc_mode != 0) mc = 4;
  CUtensorMap ta, tb;
  const int bf = in_dt == 1;
  bool ok = a_mn ? make_tmap(&ta, A, bf, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, BK)
                 : make_tmap(&ta, A, bf, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BK, (uint32_t)(BM / mc));
  ok = ok && (b_mn ? make_tmap(&tb, B, bf, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, BK)
                   : make_tmap(&tb, B, bf, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BK, (uint32_t)bn));
  if (!ok) return -13;
  Params p{C, bias, ldc, M, N, K, accumulate};
#define RB_GO(OutT, FMT) return dispatch_bn<OutT, FMT>(bn, a_mn != 0, b_mn != 0, mc, ta, tb, p, num_sms, s)
  if (in_dt == 1) {
    if (out_dt == 1) RB_GO(__nv_bfloat16, 1);
    if (out_dt == 0) RB_GO(float, 1);
  } else {
    if (out_dt == 2) RB_GO(__half, 0);
    if (out_dt == 0) RB_GO(float, 0);
  }
#undef RB_GO
  return -14;
}

// Small-M GEMM (decode shapes): M <= 128, K-major operands, no accumulate.  `ws` holds 2 * num_sms * 128 * 256 floats,
// `flags` 8192 zero-initialised words (persistent per device; calls sharing them must be stream-ordered).
// bn = 0 / split = 0: pick the tile width (multiple of 16) and the K split from a byte-ingest cost model:
//   t = max(weight bytes / HBM, bytes per CTA / ~70 GB/s per-SM TMA ingest) + fix-up(split > 1),
// over the layouts with tiles_n * split <= #SMs (tile-aligned split: one fix-up round per CTA).
// in_dt: 1 bf16, 2 fp16, 3 fp8 e4m3 (then scale_a [M] / scale_b [N] fp32 are required; lda / ldb count bytes = elements)
static int streamk_impl(const void* A, const void* B, void* C, const void* bias, int M, int N, int K, int64_t lda, int64_t ldb,
                        int64_t ldc, int in_dt, int out_dt, int bn, int split, int num_sms, void* ws, void* flags, void* dbg,
                        const float* scale_a, const float* scale_b, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (M > BM || (in_dt != 1 && in_dt != 2 && in_dt != 3)) return -30;
  const int esz = in_dt == 3 ? 1 : 2, ld_align = 16 / esz;
  if ((lda % ld_align) || (ldb % ld_align) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return -11;
  if (in_dt == 3 && (scale_a == nullptr || scale_b == nullptr)) return -34;
  if (num_sms <= 0) num_sms = rb::kNumSMs;
  if (num_sms > 160) num_sms = 160;
  const int kbe = 128 / esz;  // elements per 128-byte k-block
  const int num_kb = RB_CEIL_DIV(K, kbe);
  const int box_rows = ((M + 7) / 8) * 8;
  if (bn == 0) {
    double best = 1e30;
    for (int cand = 32; cand <= 256; cand += 16) {
      const int tiles = RB_CEIL_DIV(N, cand);
      if (tiles > num_sms && cand < 256) continue;
      int smax = tiles <= num_sms ? num_sms / tiles : 1;
      if (smax > num_kb) smax = num_kb;
      for (int sp = 1; sp <= smax; ++sp) {
        const double waves = (double)RB_CEIL_DIV(tiles * sp, num_sms);
        const double cta_bytes = (double)(box_rows + cand) * 128.0 * RB_CEIL_DIV(num_kb, sp) * waves;
        double t = cta_bytes / 70e3;                              // us
        const double hbm = (double)N * K * esz / 6.5e6;           // us
        if (t < hbm) t = hbm;
        const double kb_floor = 0.33 * RB_CEIL_DIV(num_kb, sp) * waves;  // measured per-k-block pipeline floor
        if (t < kb_floor) t = kb_floor;
        if (sp > 1) t += 5.0 + 4.5 * (double)M * cand * 4.0 / 60e3;  // partial write + barrier + cooperative reduce (measured)
        if (t < best) { best = t; bn = cand; split = sp; }
      }
    }
  }
  if (bn % 16 || bn < 16 || bn > 256) return -5;
  static const bool verbose = getenv("REAL_GEMM_DEBUG") != nullptr;
  if (verbose) fprintf(stderr, "[rb_gemm_smallm] M=%d N=%d K=%d -> bn=%d split=%d\n", M, N, K, bn, split);
  const int tiles_n = RB_CEIL_DIV(N, bn);
  if (tiles_n > 4096) return -31;
  const int64_t units = (int64_t)tiles_n * num_kb;
  int grid;
  if (split > 0 && (int64_t)tiles_n * split <= num_sms && split <= num_kb) {
    grid = tiles_n * split;            // tile-aligned: CTA ranges never straddle a tile
  } else {
    grid = (int)(units < num_sms ? units : num_sms);  // general stream-K: equal contiguous ranges, up to two fix-ups per CTA
  }
  // smem: [stages x A(box_rows x 128 B)] [stages x B(bn x 128 B)] [barriers, 512 B]; the last A stage is read up to 16 KB
  // past its base, which must still fall inside the allocation -> require stages * b_bytes >= 16 KB (true for stages >= 2)
  const int a_bytes = box_rows * 128, b_bytes = bn * BK * 2;
  int stages = (227 * 1024 - 1024 - 1024 - 512) / (a_bytes + b_bytes);
  if (stages > 16) stages = 16;
  if (stages < 2) return -32;
  while (stages * b_bytes < BM * BK * 2 && stages < 16) ++stages;
  if (stages * b_bytes < BM * BK * 2) return -33;
  CUtensorMap ta, tb;
  const int bf = in_dt == 1;
  bool ok = in_dt == 3 ? make_tmap_u8(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, (uint32_t)box_rows) &&
                             make_tmap_u8(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 128, (uint32_t)bn)
                       : make_tmap(&ta, A, bf, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BK, (uint32_t)box_rows) &&
                             make_tmap(&tb, B, bf, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BK, (uint32_t)bn);
  if (!ok) return -13;
  Params p{C, bias, ldc, M, N, K, 0};
  StreamKParams sk{reinterpret_cast<float*>(ws), reinterpret_cast<uint32_t*>(flags), box_rows, bn, stages,
                   reinterpret_cast<long long*>(dbg), scale_a, scale_b};
  if (in_dt == 3) {
    if (out_dt == 1) return launch_streamk<__nv_bfloat16, 2>(ta, tb, p, sk, grid, s);
    if (out_dt == 0) return launch_streamk<float, 2>(ta, tb, p, sk, grid, s);
    return -14;
  }
  if (in_dt == 1) {
    if (out_dt == 1) return launch_streamk<__nv_bfloat16, 1>(ta, tb, p, sk, grid, s);
    if (out_dt == 0) return launch_streamk<float, 1>(ta, tb, p, sk, grid, s);
  } else {
    if (out_dt == 2) return launch_streamk<__half, 0>(ta, tb, p, sk, grid, s);
    if (out_dt == 0) return launch_streamk<float, 0>(ta, tb, p, sk, grid, s);
  }
  return -14;
}

int rb_gemm_streamk(const void* A, const void* B, void* C, const void* bias, int M, int N, int K, int64_t lda, int64_t ldb,
                    int64_t ldc, int in_dt, int out_dt, int bn, int split, int num_sms, void* ws, void* flags, void* dbg,
                    cudaStream_t s) {
  if (in_dt == 3) return -30;
  return streamk_impl(A, B, C, bias, M, N, K, lda, ldb, ldc, in_dt, out_dt, bn, split, num_sms, ws, flags, dbg, nullptr, nullptr, s);
}

// W8A8 decode GEMM: A [M, K] / B [N, K] e4m3 bytes, C[m, n] = (sum_k A B) * scale_a[m] * scale_b[n] (+ bias), out bf16 / fp32.
int rb_gemm_streamk_fp8(const void* A, const void* B, void* C, const void* bias, const float* scale_a, const float* scale_b, int M,
                        int N, int K, int64_t lda, int64_t ldb, int64_t ldc, int out_dt, int bn, int split, int num_sms, void* ws,
                        void* flags, cudaStream_t s) {
  return streamk_impl(A, B, C, bias, M, N, K, lda, ldb, ldc, 3, out_dt, bn, split, num_sms, ws, flags, nullptr, scale_a, scale_b, s);
}

// Grouped GEMM (MoE experts) in ONE launch: rows [off[g], off[g+1]) of A are multiplied by group g's weight.
//   A [M, K] K-major (tokens sorted by group), group_offsets: device int32 [G + 1] (no host copy of the counts needed).
//   b_mn == 0: B is [G * N, K] (weights [G, N, K]):  C[rows_g] = A[rows_g] @ W_g^T
//   b_mn == 1: B is [G * K, N] (weights [G, K, N]): C[rows_g] = A[rows_g] @ W_g      (dgrad of the first form)
// Tile -> (group, m, n) is resolved on the device from the prefix sums; a tile never crosses a group boundary (its tail rows
// are computed and masked in the epilogue).
int rb_gemm_grouped(const void* A, const void* B, void* C, const int* group_offsets, int G, int M, int N, int K, int64_t lda,
                    int64_t ldb, int64_t ldc, int b_mn, int num_sms, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0 || G <= 0) return 0;
  if (G > kMaxGroups) return -50;
  if ((lda % 8) || (ldb % 8)) return -21;
  if (num_sms <= 0) num_sms = rb::kNumSMs;
  const int bn = N >= 256 ? 256 : 128;
  FusedParams fp{};
  fp.group_offsets = group_offsets; fp.num_groups = G; fp.group_b_rows = b_mn ? K : N;
  CUtensorMap ta, tb;
  bool ok = make_tmap(&ta, A, 1, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BK, BM);
  ok = ok && (b_mn ? make_tmap(&tb, B, 1, (uint64_t)G * K, (uint64_t)N, (uint64_t)ldb, 64, BK)
                   : make_tmap(&tb, B, 1, (uint64_t)G * N, (uint64_t)K, (uint64_t)ldb, BK, (uint32_t)bn));
  if (!ok) return -13;
  Params p{C, nullptr, ldc, M, N, K, 0};
  // upper bound of the tile count (the exact number is only known on the device): grid = persistent CTAs
  const int64_t max_tiles = ((int64_t)RB_CEIL_DIV(M, BM) + G) * RB_CEIL_DIV(N, bn);
  const int grid_sms = (int)(max_tiles < num_sms ? max_tiles : num_sms);
#define RB_G(BNV, BMN) return launch<BNV, false, BMN, __nv_bfloat16, 1, 1, 3>(ta, tb, p, grid_sms, s, fp, kNoPeers)
  if (bn == 256) { if (b_mn) RB_G(256, true); else RB_G(256, false); }
  else { if (b_mn) RB_G(128, true); else RB_G(128, false); }
#undef RB_G
  return -24;
}

// Fused tensor-parallel GEMMs (bf16, K-major A).  mode 1: GEMM -> reduce-scatter scatter phase; mode 2: all-gather -> GEMM.
//   mode 1: A [M, K] local, output rows go to rank row / rows_per_rank (inbox slabs, see FusedParams); C is unused.
//   mode 2: A is the concatenation over ranks of [rows_per_rank, K] buffers at `peer_a[r]` (row pitch lda); M = world*rows_per_rank.
int rb_gemm_fused_tp(int mode, const void* A, const int64_t* peer_a, const void* B, void* C, int M, int N, int K, int64_t lda, int64_t ldb,
                     int64_t ldc, int b_mn, const int64_t* peer_base, const int64_t* peer_counter, int rows_per_rank, int my_rank,
                     int world, int num_sms, cudaStream_t s) {
  if (world > kMaxTP || M <= 0 || N <= 0 || K <= 0) return -20;
  if ((lda % 8) || (ldb % 8)) return -21;
  if (num_sms <= 0) num_sms = rb::kNumSMs;
  const int tm = RB_CEIL_DIV(M, BM);
  int bn = ((int64_t)tm * RB_CEIL_DIV(N, 256) >= num_sms || N % 256 == 0) ? 256 : 128;
  if (N < 256) bn = 128;
  FusedParams fp{};
  TmaArray peers{};
  CUtensorMap ta{}, tb;
  fp.rows_per_rank = rows_per_rank; fp.my_rank = my_rank; fp.world = world; fp.slab_elems = (int64_t)rows_per_rank * N;
  bool ok = b_mn ? make_tmap(&tb, B, 1, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, BK)
                 : make_tmap(&tb, B, 1, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BK, (uint32_t)bn);
  if (mode == 1) {
    for (int r = 0; r < world; ++r) { fp.peer_base[r] = peer_base[r]; fp.peer_counter[r] = peer_counter[r]; }
    ok = ok && make_tmap(&ta, A, 1, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BK, BM);
  } else if (mode == 2) {
    if (rows_per_rank % BM != 0 || M != rows_per_rank * world) return -22;
    for (int r = 0; r < world; ++r)
      ok = ok && make_tmap(&peers.m[r], reinterpret_cast<const void*>(peer_a[r]), 1, (uint64_t)rows_per_rank, (uint64_t)K, (uint64_t)lda, BK, BM);
    ta = peers.m[my_rank];
  } else {
    return -23;
  }
  if (!ok) return -13;
  Params p{C, nullptr, ldc, M, N, K, 0};
#define RB_F(BNV, BMN, MODE) return launch<BNV, false, BMN, __nv_bfloat16, 1, 1, MODE>(ta, tb, p, num_sms, s, fp, peers)
  if (mode == 1) {
    if (bn == 256) { if (b_mn) RB_F(256, true, 1); else RB_F(256, false, 1); }
    else { if (b_mn) RB_F(128, true, 1); else RB_F(128, false, 1); }
  } else {
    if (bn == 256) { if (b_mn) RB_F(256, true, 2); else RB_F(256, false, 2); }
    else { if (b_mn) RB_F(128, true, 2); else RB_F(128, false, 2); }
  }
#undef RB_F
  return -24;
}

}  // extern "C"
