// Packed variable-length (causal) attention forward for sm_100a: TMA -> shared (128B swizzle) -> tcgen05.mma -> TMEM.
//
//   O[t, h, :] = softmax(scale * Q[t, h, :] K[seq(t), kv(h), :]^T  (+ causal mask)) V        bf16 / fp16, fp32 softmax
//
// Reference: flash-attn's `flash_attn_varlen_func` called from `impl/model/modules/attn.py:240-262`.  One CTA owns a
// 128-row query tile of one (sequence, head) and walks the key / value tiles of that sequence (up to the diagonal).
//
// CTA = 6 warps.  warp 0: TMA producer (Q once, then K_j / V_j into two-stage rings; K and V have their own barriers so a K
// stage is recycled as soon as S_j = Q K_j^T retired, before P_j V_j even started).  warp 1: MMA issuer + TMEM allocator.
// warps 2-5: softmax, one thread per query row (thread <-> TMEM lane), no cross-thread reductions.
//
// Tensor memory (512 columns): S double buffer [0,128) [128,256) and a double buffer for the per-tile product
// T_j = P_j V_j at [256,256+D) [384,384+D).  The running output stays in registers: O <- O * alpha_j + T_j with
// alpha_j = exp2((m_{j-1} - m_j) * scale * log2 e); accumulating per tile in TMEM and folding in registers costs one
// tcgen05.ld of T_j but no TMEM read-modify-write of O when the row maximum moves.  Issue order on the tensor pipe is
// S_0, S_1, PV_0, S_2, PV_1, ...: S_{j+1} is in flight while the softmax warps work on S_j, and the fold of T_{j-1} happens
// after P_j has been handed to the MMA warp, so it overlaps PV_j.
//
// P_j is written by the softmax threads as bf16 straight into the canonical K-major SWIZZLE_128B layout (the same bytes a
// TMA load of a [128 x 64] box would produce), fenced into the async proxy and consumed as the A operand of PV_j; V_j is
// the MN-major B operand (keys are the K dimension, head-dim contiguous), exactly the wgrad operand form of the GEMM.
//
// Outputs: O [T, nq, D] and LSE [nq, T] (natural log, scale applied) in the layout flash-attn's backward consumes.
//
// STATUS: compiled and SASS-checked for sm_100a; not yet run on hardware (written after the round's GPU budget was
// spent).  It is therefore opt-in (`REAL_ATTN=tcgen05`) and its GPU test is gated behind `REAL_TEST_EXPERIMENTAL=1`.
#include "gemm_common.cuh"

namespace {

constexpr int kBQ = 128;   // query rows per CTA (= UMMA M)
constexpr int kBKV = 128;  // keys per tile (= UMMA N of S, K of PV)
constexpr int kAttnThreads = 192;

struct AttnParams {
  const int* cu_seqlens;  // [B + 1]
  void* out;              // [T, nq, D]
  float* lse;             // [nq, T]
  int64_t out_ld;         // elements between consecutive tokens of `out`
  int T, nq, nkv;
  float scale;
  int causal;
};

template <int kD> struct ACfg {
  static constexpr int kDBlocks = kD / 64;               // 64-element (128 B) column blocks of Q / K / V
  static constexpr int kQBytes = kBQ * kD * 2;
  static constexpr int kKVBytes = kBKV * kD * 2;          // one K or V stage
  static constexpr int kPBytes = kBQ * kBKV * 2;
  static constexpr int kSmemBytes = kQBytes + 4 * kKVBytes + kPBytes + 1024 /*align*/ + 256 /*barriers*/;
};

enum Bar { Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 7, S_FULL = 9, S_EMPTY = 11, P_FULL = 13, P_EMPTY = 14,
           O_FULL = 15, O_EMPTY = 17, NUM_BARS = 19 };

RB_DEVICE void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <typename T> RB_DEVICE uint32_t pack2(float lo, float hi);
template <> RB_DEVICE uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> RB_DEVICE uint32_t pack2<__half>(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int kD, typename T, int kFmt>
__global__ void __launch_bounds__(kAttnThreads, 1) attn_fwd_kernel(const __grid_constant__ CUtensorMap tma_q,
                                                                   const __grid_constant__ CUtensorMap tma_k,
                                                                   const __grid_constant__ CUtensorMap tma_v, AttnParams p) {
  using C = ACfg<kD>;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int qt = gridDim.x - 1 - blockIdx.x;  // long (late) query tiles first: they walk the most key tiles
  const int tok0 = p.cu_seqlens[seq];
  const int L = p.cu_seqlens[seq + 1] - tok0;
  const int q0 = qt * kBQ;
  if (q0 >= L) return;  // uniform for the CTA, before any barrier / TMEM allocation
  const int hk = head / (p.nq / p.nkv);
  const int kv_len = p.causal ? min(L, q0 + kBQ) : L;
  const int n_kv = RB_CEIL_DIV(kv_len, kBKV);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sQ = ptx::smem_u32(smem);
  const uint32_t sK = sQ + C::kQBytes;
  const uint32_t sV = sK + 2 * C::kKVBytes;
  const uint32_t sP = sV + 2 * C::kKVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kQBytes + 4 * C::kKVBytes + C::kPBytes);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
  auto bar = [&](int i) { return ptx::smem_u32(&bars[i]); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tma_q);
    ptx::prefetch_tensormap(&tma_k);
    ptx::prefetch_tensormap(&tma_v);
    for (int i = 0; i < NUM_BARS; ++i) {
      const bool softmax_side = (i == S_EMPTY || i == S_EMPTY + 1 || i == P_FULL || i == O_EMPTY || i == O_EMPTY + 1);
      ptx::mbar_init(bar(i), softmax_side ? 4 : 1);  // one arrival per softmax warp
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

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bar(Q_FULL), C::kQBytes);
#pragma unroll
      for (int kb = 0; kb < C::kDBlocks; ++kb)
        ptx::tma_load_2d(sQ + kb * (kBQ * 128), &tma_q, bar(Q_FULL), head * kD + kb * 64, tok0 + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t par = (uint32_t)(j >> 1) & 1u;
        const int row = tok0 + j * kBKV;
        ptx::mbar_wait(bar(K_EMPTY + st), par ^ 1);
        ptx::mbar_arrive_expect_tx(bar(K_FULL + st), C::kKVBytes);
#pragma unroll
        for (int kb = 0; kb < C::kDBlocks; ++kb)
          ptx::tma_load_2d(sK + st * C::kKVBytes + kb * (kBKV * 128), &tma_k, bar(K_FULL + st), hk * kD + kb * 64, row);
        ptx::mbar_wait(bar(V_EMPTY + st), par ^ 1);
        ptx::mbar_arrive_expect_tx(bar(V_FULL + st), C::kKVBytes);
        // V is the MN-major B operand of P V: it is staged as two 64-key halves, each [d chunk][64 keys][128 B], i.e. exactly
        // the stage format (and the 8 KB chunk stride) of the GEMM's MN-major operands
#pragma unroll
        for (int half = 0; half < kBKV / 64; ++half)
#pragma unroll
          for (int kb = 0; kb < C::kDBlocks; ++kb)
            ptx::tma_load_2d(sV + st * C::kKVBytes + half * (64 * kD * 2) + kb * (64 * 128), &tma_v, bar(V_FULL + st),
                             hk * kD + kb * 64, row + half * 64);
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_f16(kFmt, kBQ, kBKV, 0, 0);  // Q, K both K-major (head-dim contiguous)
      constexpr uint32_t idesc_o = ptx::make_idesc_f16(kFmt, kBQ, kD, 0, 1);     // P K-major, V MN-major
      auto issue_s = [&](int j) {
        const int st = j & 1;
        const uint32_t par = (uint32_t)(j >> 1) & 1u;
        ptx::mbar_wait(bar(K_FULL + st), par);
        ptx::mbar_wait(bar(S_EMPTY + st), par ^ 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * (kBQ * 128) + (kk & 3) * 32;
          ptx::tc_mma_f16(tmem_base + st * 128, ptx::make_smem_desc_sw128(sQ + off, 16, 1024),
                          ptx::make_smem_desc_sw128(sK + st * C::kKVBytes + off, 16, 1024), idesc_s, kk != 0 ? 1u : 0u);
        }
        ptx::tc_commit(bar(K_EMPTY + st));  // K stage reusable once S_j retired
        ptx::tc_commit(bar(S_FULL + st));
      };
      ptx::mbar_wait(bar(Q_FULL), 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_s(j + 1);
        const int st = j & 1;
        const uint32_t par = (uint32_t)(j >> 1) & 1u;
        ptx::mbar_wait(bar(P_FULL), (uint32_t)j & 1u);
        ptx::mbar_wait(bar(V_FULL + st), par);
        ptx::mbar_wait(bar(O_EMPTY + st), par ^ 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kBKV / 16; ++kk) {
          // A: 16 keys = 32 B inside the 128 B row of key block kk/4; B: 16 key rows = 2048 B down every 64-wide d chunk of
          // key half kk/4
          ptx::tc_mma_f16(tmem_base + 256 + st * 128,
                          ptx::make_smem_desc_sw128(sP + (kk >> 2) * (kBQ * 128) + (kk & 3) * 32, 16, 1024),
                          ptx::make_smem_desc_sw128(sV + st * C::kKVBytes + (kk >> 2) * (64 * kD * 2) + (kk & 3) * 2048, 64 * 128, 1024),
                          idesc_o, kk != 0 ? 1u : 0u);
        }
        ptx::tc_commit(bar(V_EMPTY + st));
        ptx::tc_commit(bar(P_EMPTY));
        ptx::tc_commit(bar(O_FULL + st));
      }
    }
  } else {
    // ===================================================== softmax + output (4 warps; thread = query row = TMEM lane)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int q_pos = q0 + row;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;
    const uint32_t p_row = sP + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
    float O[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) O[i] = 0.f;

    auto fold = [&](int t, float a) {  // O <- O * a + T_t
      const int b = t & 1;
      ptx::mbar_wait(bar(O_FULL + b), (uint32_t)(t >> 1) & 1u);
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < kD / 32; ++c) {
        uint32_t r[32];
        ptx::tc_ld_32x32(lane_addr + 256 + b * 128 + c * 32, r);
        ptx::tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) O[c * 32 + i] = fmaf(O[c * 32 + i], a, __uint_as_float(r[i]));
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(O_EMPTY + b));
    };

    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      const int kv0 = j * kBKV;
      ptx::mbar_wait(bar(S_FULL + b), (uint32_t)(j >> 1) & 1u);
      ptx::tc_fence_after();
      // only the last tile of the sequence and the diagonal tile need per-element masks (uniform over the CTA)
      const bool need_mask = (kv0 + kBKV > L) || (p.causal && kv0 + kBKV - 1 > q0);
      const int lim = (p.causal ? min(L, q_pos + 1) : L) - kv0;  // local key columns [0, lim) are visible to this row
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < kBKV / 32; ++c) {
        uint32_t r[32];
        ptx::tc_ld_32x32(lane_addr + b * 128 + c * 32, r);
        ptx::tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(r[i]);
          if (need_mask && c * 32 + i >= lim) s = -INFINITY;
          mx = fmaxf(mx, s);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float m_eff = (m_new == -INFINITY) ? 0.f : m_new;  // fully masked so far: keep exp2 arguments finite
      const float alpha = exp2f((m_run - m_eff) * sl2);
      const float neg_m = -m_eff * sl2;
      if (j > 0) ptx::mbar_wait(bar(P_EMPTY), (uint32_t)(j - 1) & 1u);  // PV_{j-1} finished reading the P buffer
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < kBKV / 32; ++c) {
        uint32_t r[32];
        ptx::tc_ld_32x32(lane_addr + b * 128 + c * 32, r);
        ptx::tc_wait_ld();
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float e0 = exp2f(fmaf(__uint_as_float(r[i]), sl2, neg_m));
          float e1 = exp2f(fmaf(__uint_as_float(r[i + 1]), sl2, neg_m));
          if (need_mask) {
            if (c * 32 + i >= lim) e0 = 0.f;
            if (c * 32 + i + 1 >= lim) e1 = 0.f;
          }
          sum += e0 + e1;
          w[i >> 1] = pack2<T>(e0, e1);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t piece = (uint32_t)(c * 4 + t);  // 16-byte piece (8 keys) of this row: key block piece/8, slot piece%8
          st_shared_v4(p_row + (piece >> 3) * (kBQ * 128) + (((piece & 7) ^ sw) << 4), w[4 * t], w[4 * t + 1], w[4 * t + 2],
                       w[4 * t + 3]);
        }
      }
      ptx::tc_fence_before();     // our tcgen05.ld of S_j are complete (wait::ld) before the MMA warp may overwrite the buffer
      ptx::fence_proxy_async();   // P_j (generic-proxy stores) visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(bar(S_EMPTY + b));
        ptx::mbar_arrive(bar(P_FULL));
      }
      l_run = fmaf(l_run, alpha, sum);
      m_run = m_new;
      if (j > 0) fold(j - 1, alpha_prev);
      alpha_prev = alpha;
    }
    fold(n_kv - 1, alpha_prev);

    if (q_pos < L) {
      const float inv = 1.f / l_run;
      T* dst = reinterpret_cast<T*>(p.out) + (int64_t)(tok0 + q_pos) * p.out_ld + (int64_t)head * kD;
#pragma unroll
      for (int i = 0; i < kD; i += 8) {
        uint4 v;
        v.x = pack2<T>(O[i] * inv, O[i + 1] * inv);
        v.y = pack2<T>(O[i + 2] * inv, O[i + 3] * inv);
        v.z = pack2<T>(O[i + 4] * inv, O[i + 5] * inv);
        v.w = pack2<T>(O[i + 6] * inv, O[i + 7] * inv);
        *reinterpret_cast<uint4*>(dst + i) = v;
      }
      p.lse[(int64_t)head * p.T + tok0 + q_pos] = m_run * p.scale + __logf(l_run);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

template <int kD, typename T, int kFmt>
int launch_attn_fwd(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const AttnParams& p, int B,
                    int max_seqlen, cudaStream_t s) {
  using C = ACfg<kD>;
  auto kern = attn_fwd_kernel<kD, T, kFmt>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) != cudaSuccess) return -30;
    attr_set = true;
  }
  dim3 grid(RB_CEIL_DIV(max_seqlen, kBQ), p.nq, B);
  kern<<<grid, kAttnThreads, C::kSmemBytes, s>>>(mq, mk, mv, p);
  return cudaGetLastError() == cudaSuccess ? 0 : -31;
}

}  // namespace

// q / k / v: [T, heads, hd] views (unit stride over hd, `heads * hd` contiguous columns, row pitch *_ld elements) -- e.g. the
// three column ranges of the fused QKV projection.  out [T, nq, hd] with row pitch out_ld, lse [nq, T] fp32.
// dt: 1 = bf16, 2 = fp16.  Returns 0 or a negative code (unsupported shape / descriptor failure).
extern "C" int rb_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, const int* cu_seqlens,
                           int64_t q_ld, int64_t k_ld, int64_t v_ld, int64_t out_ld, int T, int B, int nq, int nkv, int hd,
                           int max_seqlen, float scale, int causal, int dt, cudaStream_t s) {
  if (hd != 128 && hd != 64) return -1;
  if (dt != 1 && dt != 2) return -2;
  if (nq % nkv != 0 || T <= 0 || B <= 0 || max_seqlen <= 0) return -3;
  if ((q_ld | k_ld | v_ld | out_ld) % 8 != 0) return -4;
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) return -5;
  CUtensorMap mq, mk, mv;
  const int bf = dt == 1;
  if (!make_tmap(&mq, q, bf, T, (uint64_t)nq * hd, q_ld, 64, kBQ)) return -10;
  if (!make_tmap(&mk, k, bf, T, (uint64_t)nkv * hd, k_ld, 64, kBKV)) return -11;
  if (!make_tmap(&mv, v, bf, T, (uint64_t)nkv * hd, v_ld, 64, 64)) return -12;  // 64-key halves, see the producer
  AttnParams p{cu_seqlens, out, lse, out_ld, T, nq, nkv, scale, causal};
  if (hd == 128) {
    return bf ? launch_attn_fwd<128, __nv_bfloat16, 1>(mq, mk, mv, p, B, max_seqlen, s)
              : launch_attn_fwd<128, __half, 0>(mq, mk, mv, p, B, max_seqlen, s);
  }
  return bf ? launch_attn_fwd<64, __nv_bfloat16, 1>(mq, mk, mv, p, B, max_seqlen, s)
            : launch_attn_fwd<64, __half, 0>(mq, mk, mv, p, B, max_seqlen, s);
}
