// CTA-pair GEMM for sm_100a: tcgen05.mma.cta_group::2, one 256 x BN output tile per pair of SMs.
//
//   D[M,N] (+)= A x B (+ bias)      bf16/fp16 operands, fp32 accumulation in tensor memory.
//
// Why a second kernel: ncu on the single-CTA 128x256 tile kernel shows the tensor pipe ~80% active with the MMA warp
// never starved by TMA -- each 128x256x16 MMA reads 12 KB of shared memory per 128 cycles while TMA writes the next
// stage at the same rate, which saturates the SM's shared-memory port.  With cta_group::2 the two SMs of a TPC execute
// one 256 x BN x 16 MMA together: every CTA stages its own 128 rows of A but only HALF of the B tile (BN/2 columns), and
// the tensor cores of both SMs read both halves.  Shared-memory traffic per SM drops by a third, the stage shrinks from
// 48 KB to 32 KB (6 pipeline stages instead of 4), and B is fetched from L2 once per pair.
//
// Roles per CTA (6 warps, as in gemm_tcgen05.cu): warp 0 TMA producer (both CTAs load; every transfer completes on the
// LEADER's full barrier), warp 1 lane 0 of the leader (cluster rank 0) issues all MMAs and multicasts tcgen05.commit
// onto both CTAs' empty / accumulator-full barriers, warps 2-5 drain the CTA's own 128 accumulator rows from TMEM and
// release the accumulator stage on the leader's barrier (8 arrivals: 4 warps x 2 CTAs).
// Operand majors as in the tile kernel (K-major or MN-major each), so fwd / dgrad / wgrad all run here.
#include "gemm_common.cuh"

namespace {

template <int BN> struct Cfg2 {
  static_assert(BN == 128 || BN == 256, "2-CTA tiles are 256x128 or 256x256");
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / 2) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN == 256 ? 6 : 8;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

RB_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
RB_DEVICE uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
// TMA load whose completion is signalled on an mbarrier that may live in the peer CTA of the pair.
RB_DEVICE void tma_load_2d_cg2(uint32_t smem_dst, const void* tmap, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(tmap), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
RB_DEVICE void tc_mma_f16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
RB_DEVICE void tc_commit_cg2(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
RB_DEVICE void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile("{\n\t.reg .b32 r;\n\tmapa.shared::cluster.u32 r, %0, %1;\n\tmbarrier.arrive.release.cluster.shared::cluster.b64 _, [r];\n\t}" ::"r"(bar),
               "r"(cta)
               : "memory");
}
RB_DEVICE void tmem_alloc_cg2(uint32_t smem_result, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "r"(cols) : "memory");
}
RB_DEVICE void tmem_relinquish_cg2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
RB_DEVICE void tmem_dealloc_cg2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// Tiles are 256 rows tall here; bands of 8 pair-tiles keep the in-flight set roughly square (see tile_coords in gemm_tcgen05.cu).
constexpr int kGroupM2 = 8;
RB_DEVICE void tile_coords2(int tile, int tiles_m, int tiles_n, int bn, int m_rot, int& m0, int& n0) {
  const int per_group = kGroupM2 * tiles_n;
  const int group = tile / per_group, within = tile - group * per_group;
  const int gm0 = group * kGroupM2;
  const int gsize = min(kGroupM2, tiles_m - gm0);
  m0 = ((gm0 + within % gsize + m_rot) % tiles_m) * (2 * BM);
  n0 = (within / gsize) * bn;
}
RB_DEVICE uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

RB_DEVICE float glu_act_fwd(float g, int kind) {
  if (kind == 0) return g / (1.f + __expf(-g));
  const float u = 0.7978845608028654f * (g + 0.044715f * g * g * g);
  return 0.5f * g * (1.f + tanhf(u));
}

// kGlu: gated-linear-unit epilogue (SwiGLU / GeGLU fused into the gate|up projection).  The pair's B tile is made of
// 128 GATE rows (loaded by CTA 0) and the 128 matching UP rows (loaded by CTA 1) of the fused [gate; up] weight -- two TMA boxes
// at different row coordinates, no change to the parameter layout -- so accumulator columns [0,128) and [128,256) of every row
// hold gate_j and up_j for the same 128 output features and the epilogue writes act(gate) * up directly.
template <int BN, bool kAMN, bool kBMN, typename OutT, int kFmt, bool kGlu = false>
__global__ void __launch_bounds__(kThreads, 1) gemm_2cta_kernel(const __grid_constant__ CUtensorMap tma_a,
                                                                const __grid_constant__ CUtensorMap tma_b, Params p) {
  using C = Cfg2<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;                      // used on the leader only
  uint64_t* empty_bar = bars + C::kStages;        // one per CTA, armed by the leader's multicast commit
  uint64_t* tmem_full = bars + 2 * C::kStages;    // one per CTA
  uint64_t* tmem_empty = tmem_full + 2;           // leader only, 8 arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  constexpr int kTileN = kGlu ? BN / 2 : BN;  // output columns per pair tile
  const int tiles_m = RB_CEIL_DIV(p.M, 2 * BM), tiles_n = RB_CEIL_DIV(p.N, kTileN);
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = RB_CEIL_DIV(p.K, BK);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tma_a);
    ptx::prefetch_tensormap(&tma_b);
    for (int i = 0; i < C::kStages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(&tmem_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&tmem_empty[i]), 8);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(ptx::smem_u32(tmem_ptr), C::kTmemCols);
    tmem_relinquish_cg2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();  // the peer's barriers must exist before any remote arrive / multicast commit / cross-CTA TMA completion
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int m0, n0;
        tile_coords2(tile, tiles_m, tiles_n, kTileN, p.m_rot, m0, n0);
        m0 += (int)rank * BM;          // my 128 rows of A (and of the output)
        if constexpr (kGlu) n0 += (int)rank * p.glu_F;  // CTA 0: gate rows of these features, CTA 1: the matching up rows
        else n0 += (int)rank * (BN / 2);                // my half of the B tile
        if (p.ready_flags != nullptr && m0 < p.M) {  // A rows of this tile may still be in flight from a peer
          const uint32_t* f = p.ready_flags + m0 / p.rows_per_flag;
          while ((int32_t)(ld_acquire_gpu_u32(f) - p.ready_epoch) < 0) {}
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb_local = ptx::smem_u32(&full_bar[stage]);
          if (leader) ptx::mbar_arrive_expect_tx(fb_local, 2 * C::kStageBytes);  // both CTAs' bytes land on this barrier
          const uint32_t fb = mapa_u32(fb_local, 0);
          const uint32_t sa = ptx::smem_u32(smem_a + stage * C::kABytes);
          const uint32_t sb = ptx::smem_u32(smem_b + stage * C::kBBytes);
          const int k0 = kb * BK;
          if constexpr (kAMN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d_cg2(sa + j * (BK * 128), &tma_a, fb, m0 + 64 * j, k0);
          } else {
            tma_load_2d_cg2(sa, &tma_a, fb, k0, m0);
          }
          if constexpr (kBMN) {
#pragma unroll
            for (int j = 0; j < BN / 2 / 64; ++j) tma_load_2d_cg2(sb + j * (BK * 128), &tma_b, fb, n0 + 64 * j, k0);
          } else {
            tma_load_2d_cg2(sb, &tma_b, fb, k0, n0);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(kFmt, 2 * BM, BN, kAMN ? 1 : 0, kBMN ? 1 : 0);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
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
            tc_mma_f16_cg2(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit_cg2(ptx::smem_u32(&empty_bar[stage]), 0b11);  // frees the stage in both CTAs once these MMAs retire
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit_cg2(ptx::smem_u32(&tmem_full[as]), 0b11);  // accumulator complete -> both epilogues
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================================================== epilogue (4 warps per CTA, own 128 accumulator rows)
    const int quad = warp & 3;
    int as = 0;
    uint32_t aphase = 0;
    OutT* Cp = reinterpret_cast<OutT*>(p.C);
    const OutT* bias = reinterpret_cast<const OutT*>(p.bias);
    const bool vec_ok = (p.ldc % (16 / sizeof(OutT)) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int m0, n0;
      tile_coords2(tile, tiles_m, tiles_n, kTileN, p.m_rot, m0, n0);
      ptx::mbar_wait(ptx::smem_u32(&tmem_full[as]), aphase);
      ptx::tc_fence_after();
      const int row = m0 + (int)rank * BM + quad * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;
      if constexpr (kGlu) {
        OutT* raw = reinterpret_cast<OutT*>(p.glu_raw);
        const bool raw_vec = raw != nullptr && (p.ld_raw % 8 == 0) && ((reinterpret_cast<uintptr_t>(raw) & 15) == 0) && (p.glu_F % 8 == 0);
#pragma unroll 1
        for (int c = 0; c < BN / 2 / 32; ++c) {
          uint32_t rg[32], ru[32];
          ptx::tc_ld_32x32(taddr + c * 32, rg);
          ptx::tc_ld_32x32(taddr + BN / 2 + c * 32, ru);
          ptx::tc_wait_ld();
          const int col = n0 + c * 32;
          const int n_valid = min(32, p.N - col);
          if (row < p.M && n_valid > 0) {
            float g[32], u[32], v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              g[i] = __uint_as_float(rg[i]);
              u[i] = __uint_as_float(ru[i]);
              // the unfused path rounds gate / up to the activation dtype before the nonlinearity: do the same, so that the saved
              // raw projections (backward) and the activation are consistent
              const float gr = rb::to_f(rb::from_f<OutT>(g[i])), ur = rb::to_f(rb::from_f<OutT>(u[i]));
              v[i] = glu_act_fwd(gr, p.glu_act) * ur;
            }
            store_chunk<OutT>(Cp + (int64_t)row * p.ldc + col, v, n_valid, vec_ok);
            if (raw != nullptr) {
              store_chunk<OutT>(raw + (int64_t)row * p.ld_raw + col, g, n_valid, raw_vec);
              store_chunk<OutT>(raw + (int64_t)row * p.ld_raw + p.glu_F + col, u, n_valid, raw_vec);
            }
          }
        }
      } else {
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        ptx::tc_ld_32x32(taddr + c * 32, r);
        ptx::tc_wait_ld();
        const int col = n0 + c * 32;
        const int n_valid = min(32, p.N - col);
        if (row < p.M && n_valid > 0) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < n_valid) v[i] += rb::to_f(bias[col + i]);
          }
          OutT* dst = Cp + (int64_t)row * p.ldc + col;
          if (p.accumulate) {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < n_valid) v[i] += rb::to_f(dst[i]);
          }
          store_chunk<OutT>(dst, v, n_valid, vec_ok);
        }
      }
      }
      if (p.done_counters != nullptr) {
        __threadfence();
        __syncwarp();
        const int wrow = m0 + (int)rank * BM + quad * 32;
        if (lane == 0 && wrow < p.M) atomicAdd(p.done_counters + wrow / p.rows_per_flag, 1u);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(ptx::smem_u32(&tmem_empty[as]), 0);  // leader's barrier, 8 arrivals per stage
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();  // neither CTA may free TMEM or exit while the pair's MMAs / remote arrivals are in flight
  if (warp == 1) {
    ptx::tc_fence_after();
    tmem_dealloc_cg2(tmem_base, C::kTmemCols);
  }
}

template <int BN, bool kAMN, bool kBMN, typename OutT, int kFmt, bool kGlu = false>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, int num_sms, cudaStream_t s) {
  auto kern = gemm_2cta_kernel<BN, kAMN, kBMN, OutT, kFmt, kGlu>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<BN>::kSmemBytes) != cudaSuccess) return -2;
    configured = true;
  }
  const int tiles = RB_CEIL_DIV(p.M, 2 * BM) * RB_CEIL_DIV(p.N, kGlu ? BN / 2 : BN);
  const int pairs = num_sms / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg2<BN>::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, ta, tb, p) != cudaSuccess) return -3;
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

template <int BN, typename OutT, int kFmt>
int dispatch_major2(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, int sms, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch2<BN, false, false, OutT, kFmt>(ta, tb, p, sms, s);
  if (!a_mn && b_mn) return launch2<BN, false, true, OutT, kFmt>(ta, tb, p, sms, s);
  if (a_mn && b_mn) return launch2<BN, true, true, OutT, kFmt>(ta, tb, p, sms, s);
  return launch2<BN, true, false, OutT, kFmt>(ta, tb, p, sms, s);
}

}  // namespace

extern "C" {

// Same contract as rb_gemm_tcgen05 (gemm_tcgen05.cu); bn must be 128 or 256 (0 picks).  Returns -40 when the problem is
// better served by the single-CTA kernel (one m-tile) so the caller can fall through.
int rb_gemm_2cta_gated(const void* A, const void* B, void* C, const void* bias, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc,
                       int a_mn, int b_mn, int in_dt, int out_dt, int accumulate, int bn, int num_sms, const uint32_t* ready_flags,
                       uint32_t ready_epoch, int rows_per_flag, int m_rot_rows, uint32_t* done_counters, cudaStream_t s);

int rb_gemm_2cta(const void* A, const void* B, void* C, const void* bias, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc,
                 int a_mn, int b_mn, int in_dt, int out_dt, int accumulate, int bn, int num_sms, cudaStream_t s) {
  return rb_gemm_2cta_gated(A, B, C, bias, M, N, K, lda, ldb, ldc, a_mn, b_mn, in_dt, out_dt, accumulate, bn, num_sms, nullptr, 0, 1, 0, nullptr, s);
}

// ready_flags != nullptr: rows [i * rows_per_flag, (i+1) * rows_per_flag) of A may only be read once flag[i] >= ready_epoch
// (rows_per_flag must be a multiple of 128); m_rot_rows: process the m-tiles starting at this row first.
int rb_gemm_2cta_gated(const void* A, const void* B, void* C, const void* bias, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc,
                       int a_mn, int b_mn, int in_dt, int out_dt, int accumulate, int bn, int num_sms, const uint32_t* ready_flags,
                       uint32_t ready_epoch, int rows_per_flag, int m_rot_rows, uint32_t* done_counters, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (in_dt != 1 && in_dt != 2) return -10;
  if ((lda % 8) || (ldb % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return -11;
  const bool gated = ready_flags != nullptr || done_counters != nullptr;
  if (M <= BM && !gated) return -40;
  if (gated && (rows_per_flag <= 0 || rows_per_flag % BM != 0 || a_mn)) return -41;
  if (num_sms <= 0) num_sms = rb::kNumSMs;
  if (bn == 0) {
    // 256x256 pair tiles, unless they would leave most SM pairs idle (then the single-CTA kernel's smaller tiles win)
    const int64_t pair_tiles = (int64_t)RB_CEIL_DIV(M, 2 * BM) * RB_CEIL_DIV(N, 256);
    if (!gated && pair_tiles * 10 < (int64_t)(num_sms / 2) * 7) return -40;
    bn = N > 128 ? 256 : 128;
  }
  if (bn != 128 && bn != 256) return -5;
  CUtensorMap ta, tb;
  const int bf = in_dt == 1;
  bool ok = a_mn ? make_tmap(&ta, A, bf, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, BK)
                 : make_tmap(&ta, A, bf, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BK, (uint32_t)BM);
  ok = ok && (b_mn ? make_tmap(&tb, B, bf, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, BK)
                   : make_tmap(&tb, B, bf, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BK, (uint32_t)(bn / 2)));
  if (!ok) return -13;
  Params p{C, bias, ldc, M, N, K, accumulate, ready_flags, ready_epoch, rows_per_flag > 0 ? rows_per_flag : 1,
           (m_rot_rows / (2 * BM)) % RB_CEIL_DIV(M, 2 * BM), done_counters, 0, 0, nullptr, 0};
#define RB_GO2(OutT, FMT)                                                                     \
  return bn == 256 ? dispatch_major2<256, OutT, FMT>(a_mn != 0, b_mn != 0, ta, tb, p, num_sms, s) \
                   : dispatch_major2<128, OutT, FMT>(a_mn != 0, b_mn != 0, ta, tb, p, num_sms, s)
  if (in_dt == 1) {
    if (out_dt == 1) RB_GO2(__nv_bfloat16, 1);
    if (out_dt == 0) RB_GO2(float, 1);
  } else {
    if (out_dt == 2) RB_GO2(__half, 0);
    if (out_dt == 0) RB_GO2(float, 0);
  }
#undef RB_GO2
  return -14;
}

// act[M, F] = glu(A[M,K] x [gate; up][2F, K]^T) with the activation in the epilogue; raw (optional) [M, 2F] receives gate | up.
// in_dt: 1 bf16, 2 fp16 (outputs use the same dtype).  act_kind: 0 silu, 1 gelu-tanh.
int rb_gemm_2cta_glu(const void* A, const void* B, void* act, void* raw, int M, int F, int K, int64_t lda, int64_t ldb, int64_t ld_act,
                     int64_t ld_raw, int in_dt, int act_kind, int num_sms, cudaStream_t s) {
  if (M <= 0 || F <= 0 || K <= 0) return 0;
  if (in_dt != 1 && in_dt != 2) return -10;
  if ((lda % 8) || (ldb % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return -11;
  if (num_sms <= 0) num_sms = rb::kNumSMs;
  CUtensorMap ta, tb;
  const int bf = in_dt == 1;
  bool ok = make_tmap(&ta, A, bf, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BK, (uint32_t)BM);
  ok = ok && make_tmap(&tb, B, bf, (uint64_t)(2 * F), (uint64_t)K, (uint64_t)ldb, BK, 128);
  if (!ok) return -13;
  Params p{act, nullptr, ld_act, M, F, K, 0, nullptr, 0, 1, 0, nullptr, F, act_kind, raw, ld_raw};
  if (in_dt == 1) return launch2<256, false, false, __nv_bfloat16, 1, true>(ta, tb, p, num_sms, s);
  return launch2<256, false, false, __half, 0, true>(ta, tb, p, num_sms, s);
}

}  // extern "C"
