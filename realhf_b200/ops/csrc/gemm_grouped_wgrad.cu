// Grouped weight-gradient GEMM for MoE experts in ONE launch:
//
//   dW[g] [M, N] (+)= dY_g^T X_g          dY [Tp, M], X [Tp, N] hold the tokens sorted by expert, expert g owns the row range
//                                          [off[g], off[g+1]) and every off[g] is a multiple of 64 (the dispatch pads each
//                                          expert's block with zero rows), so a 64-row K block never straddles two experts.
//
// Reference: grouped_gemm's `gmm(..., trans_a=True)` in the experts' backward (impl/model/modules/moe/experts.py); the first
// version here ran one GEMM per expert after a host read of the token counts.  This kernel is the wgrad form of the tile GEMM
// (`gemm_tcgen05.cu`: both operands MN-major, K = tokens) with the K range taken per tile from the device-side offsets:
// tile -> (expert, m-tile, n-tile), persistent CTAs, TMA producer warp / single-thread tcgen05.mma issuer / four epilogue
// warps, smem ring + double-buffered TMEM accumulators.  An expert without tokens gets zeros (or is left alone when
// accumulating).
//
// STATUS: compiled for sm_100a, not yet run on hardware: opt-in (`REAL_MOE_GROUPED_WGRAD=1`), GPU test gated by
// REAL_TEST_EXPERIMENTAL=1; the padding / offset logic around it is tested on CPU.
#include "gemm_common.cuh"

namespace {

struct WgradParams {
  void* C;          // [G, M, N]
  const int* off;   // [G + 1], multiples of 64
  int G, M, N;
  int accumulate;
};

template <int BN> struct WCfg {
  static constexpr int kStages = BN >= 256 ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;   // [2 chunks of 64 m][64 k rows][128 B]
  static constexpr int kBBytes = BN * BK * 2;   // [BN/64 chunks][64 k rows][128 B]
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BN, typename OutT>
__global__ void __launch_bounds__(kThreads, 1) grouped_wgrad_kernel(const __grid_constant__ CUtensorMap tma_a,
                                                                    const __grid_constant__ CUtensorMap tma_b, WgradParams p) {
  using C = WCfg<BN>;
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
  const int tpg = tiles_m * tiles_n;
  const int num_tiles = p.G * tpg;
  // tile -> (expert, m0, n0, first token row, number of 64-row K blocks); m fastest so neighbouring CTAs share the X panel
  auto coords = [&](int tile, int& g, int& m0, int& n0, int& r0, int& nkb) {
    g = tile / tpg;
    const int local = tile - g * tpg;
    m0 = (local % tiles_m) * BM;
    n0 = (local / tiles_m) * BN;
    r0 = p.off[g];
    nkb = (p.off[g + 1] - r0) / BK;
  };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tma_a);
    ptx::prefetch_tensormap(&tma_b);
    for (int i = 0; i < C::kStages; ++i) {
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
    ptx::tmem_alloc(ptx::smem_u32(tmem_ptr), 2 * BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int g, m0, n0, r0, nkb;
        coords(tile, g, m0, n0, r0, nkb);
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
          ptx::mbar_arrive_expect_tx(fb, C::kStageBytes);
          const uint32_t sa = ptx::smem_u32(smem_a + stage * C::kABytes);
          const uint32_t sb = ptx::smem_u32(smem_b + stage * C::kBBytes);
          const int k0 = r0 + kb * BK;
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) ptx::tma_load_2d(sa + j * (BK * 128), &tma_a, fb, m0 + 64 * j, k0);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) ptx::tma_load_2d(sb + j * (BK * 128), &tma_b, fb, n0 + 64 * j, k0);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(1, BM, BN, 1, 1);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int g, m0, n0, r0, nkb;
        coords(tile, g, m0, n0, r0, nkb);
        ptx::mbar_wait(ptx::smem_u32(&tmem_empty[as]), aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem_a + stage * C::kABytes);
          const uint32_t sb = ptx::smem_u32(smem_b + stage * C::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            ptx::tc_mma_f16(d_tmem, ptx::make_smem_desc_sw128(sa + k * 2048, BK * 128, 1024),
                            ptx::make_smem_desc_sw128(sb + k * 2048, BK * 128, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::tc_commit(ptx::smem_u32(&empty_bar[stage]));
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        ptx::tc_commit(ptx::smem_u32(&tmem_full[as]));  // with nkb == 0 nothing is pending: arrives at once
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    const int quad = warp & 3;
    int as = 0;
    uint32_t aphase = 0;
    OutT* Cp = reinterpret_cast<OutT*>(p.C);
    const bool vec_ok = (p.N % (16 / sizeof(OutT)) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int g, m0, n0, r0, nkb;
      coords(tile, g, m0, n0, r0, nkb);
      ptx::mbar_wait(ptx::smem_u32(&tmem_full[as]), aphase);
      ptx::tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        ptx::tc_ld_32x32(taddr + c * 32, r);  // warp-collective: executed by every lane whatever its row
        ptx::tc_wait_ld();
        const int col = n0 + c * 32;
        const int n_valid = min(32, p.N - col);
        if (row < p.M && n_valid > 0 && !(nkb == 0 && p.accumulate)) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = nkb == 0 ? 0.f : __uint_as_float(r[i]);  // no tokens: the accumulator is stale
          OutT* dst = Cp + ((int64_t)g * p.M + row) * p.N + col;
          if (p.accumulate) {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < n_valid) v[i] += rb::to_f(dst[i]);
          }
          store_chunk<OutT>(dst, v, n_valid, vec_ok);
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
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN, typename OutT>
int launch_wgrad(const CUtensorMap& ta, const CUtensorMap& tb, const WgradParams& p, int num_sms, cudaStream_t s) {
  using C = WCfg<BN>;
  auto kern = grouped_wgrad_kernel<BN, OutT>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) != cudaSuccess) return -30;
    attr_set = true;
  }
  const int64_t tiles = (int64_t)p.G * RB_CEIL_DIV(p.M, BM) * RB_CEIL_DIV(p.N, BN);
  const int grid = (int)(tiles < num_sms ? tiles : num_sms);
  kern<<<grid, kThreads, C::kSmemBytes, s>>>(ta, tb, p);
  return cudaGetLastError() == cudaSuccess ? 0 : -31;
}

}  // namespace

// dy [Tp, M], x [Tp, N] bf16 (row pitches ld*), offsets: device int32 [G + 1] with every entry a multiple of 64 and
// offsets[G] <= Tp; out [G, M, N] contiguous, fp32 (out_dt 0) or bf16 (out_dt 1); accumulate: out += result.
extern "C" int rb_gemm_grouped_wgrad(const void* dy, const void* x, void* out, const int* offsets, int G, int Tp, int M, int N,
                                     int64_t ld_dy, int64_t ld_x, int out_dt, int accumulate, int num_sms, cudaStream_t s) {
  if (G <= 0 || M <= 0 || N <= 0) return 0;
  if (Tp <= 0 || Tp % BK != 0) return -1;
  if ((ld_dy % 8) || (ld_x % 8) || (reinterpret_cast<uintptr_t>(dy) & 15) || (reinterpret_cast<uintptr_t>(x) & 15)) return -11;
  if (M % 8 || N % 8) return -12;
  if (num_sms <= 0) num_sms = rb::kNumSMs;
  CUtensorMap ta, tb;
  // MN-major operands: the tensor maps address [rows = tokens, cols = features] with 64 x 64 boxes (same as the dense wgrad)
  if (!make_tmap(&ta, dy, 1, (uint64_t)Tp, (uint64_t)M, (uint64_t)ld_dy, 64, BK) ||
      !make_tmap(&tb, x, 1, (uint64_t)Tp, (uint64_t)N, (uint64_t)ld_x, 64, BK))
    return -13;
  WgradParams p{out, offsets, G, M, N, accumulate};
  const bool wide = N % 256 == 0;
  if (out_dt == 0) return wide ? launch_wgrad<256, float>(ta, tb, p, num_sms, s) : launch_wgrad<128, float>(ta, tb, p, num_sms, s);
  if (out_dt == 1) return wide ? launch_wgrad<256, __nv_bfloat16>(ta, tb, p, num_sms, s) : launch_wgrad<128, __nv_bfloat16>(ta, tb, p, num_sms, s);
  return -14;
}
