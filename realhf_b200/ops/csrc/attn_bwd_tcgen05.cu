// Packed variable-length (causal) attention backward for sm_100a, same building blocks as attn_fwd_tcgen05.cu.
//
// Two passes over the (q block, kv block) pairs, each with ONE resident 128-row tile pair and a streamed ring of 64-row tile
// pairs, so that every accumulator lives in tensor memory for the whole CTA and nothing needs atomics or rescaling:
//
//   pass "dKV" (CTA = 128 keys of one kv head): resident K_j, V_j; streams Q_t, dO_t over all q blocks at / below the
//       diagonal and over the q heads of the GQA group.  Transposed orientation -- tensor-memory lanes are KEYS:
//         S^T = K_j Q_t^T      dP^T = V_j dO_t^T                         (two M=128, N=64, K=D MMAs per step)
//         P^T = exp2(S^T * scale*log2e - lse_q * log2e)                  (softmax statistics vary along columns)
//         dS^T = P^T o (dP^T - delta_q) * scale
//         dV_j += P^T dO_t     dK_j += dS^T Q_t                          (two M=128, N=D, K=64 MMAs per step)
//   pass "dQ" (CTA = 128 query rows of one q head): resident Q_i, dO_i; streams K_t, V_t up to the diagonal:
//         S = Q_i K_t^T        dP = dO_i V_t^T      dS = P o (dP - delta) * scale      dQ_i += dS K_t
//
// The streamed tiles are used twice with two different descriptors over the same bytes: as the K-major B operand of the
// S-type MMAs (rows = N) and as the MN-major B operand of the accumulating MMAs (rows = K), exactly the two operand forms of
// the GEMM kernel (forward / wgrad).  P^T, dS^T (or dS) are written by the softmax threads as bf16 in the K-major
// SWIZZLE_128B layout and consumed as A operands.  S / dP tensor-memory buffers and the P-type shared buffers are double
// buffered: the tensor pipe runs the S-type MMAs of step t+1 while the threads work on step t.
// Recomputing S and dP in both passes costs 7 instead of 5 MMAs per block pair; in exchange the result is deterministic,
// there is no fp32 dQ scratch tensor to zero / convert, and dK / dV of a GQA group are reduced inside one CTA.
//
// delta[h, t] = sum_d dO[t, h, d] * O[t, h, d] comes from a small pre-pass; LSE is the forward's ([nq, T], natural log).
//
// STATUS: compiled and SASS-checked for sm_100a; not yet run on hardware (see attn_fwd_tcgen05.cu).  Opt-in:
// `REAL_ATTN_BWD=tcgen05`; its numerics test needs `REAL_TEST_EXPERIMENTAL=1`.
#include "gemm_common.cuh"

namespace {

constexpr int kR = 128;   // resident rows per CTA (= UMMA M)
constexpr int kX = 64;    // streamed rows per step (= UMMA N of the S-type MMAs, K of the accumulating MMAs)
constexpr int kBwdThreads = 192;

struct BwdParams {
  const int* cu_seqlens;
  const float* lse;    // [nq, T]
  const float* delta;  // [nq, T]
  void* out1;          // dKV pass: dV;  dQ pass: dQ      ([T, heads, D] views, row pitch ld1)
  void* out2;          // dKV pass: dK
  int64_t ld1, ld2;
  int T, nq, nkv;
  float scale;
  int causal;
};

template <int kD, bool kDKV> struct BCfg {
  static constexpr int kDBlocks = kD / 64;
  static constexpr int kRBytes = kR * kD * 2;     // one resident tile
  static constexpr int kXBytes = kX * kD * 2;     // one streamed tile
  static constexpr int kPBytes = kR * kX * 2;     // one P-type buffer [128 x 64] bf16 = one 16 KB swizzle block
  static constexpr int kPBufs = kDKV ? 2 : 1;     // P^T and dS^T, or dS only
  static constexpr int kSmemData = 2 * kRBytes + 4 * kXBytes + 2 * kPBufs * kPBytes;
  static constexpr int kSmemBytes = kSmemData + 1024 /*align*/ + 256 /*barriers*/ + (kDKV ? 2 * 2 * kX * 4 : 0) /*stats*/;
};

enum BBar { R_FULL = 0, X_FULL = 1, X_EMPTY = 3, S_FULL = 5, S_EMPTY = 7, P_FULL = 9, P_EMPTY = 11, ACC_FULL = 13, B_NUM_BARS = 14 };

RB_DEVICE void st_shared_v4b(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <typename T> RB_DEVICE uint32_t pack2b(float lo, float hi);
template <> RB_DEVICE uint32_t pack2b<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> RB_DEVICE uint32_t pack2b<__half>(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
RB_DEVICE void softmax_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// tma_r1 / tma_r2: resident operands (box 64 cols x 128 rows); tma_x1 / tma_x2: streamed operands (box 64 x 64).
//   kDKV: r1 = K, r2 = V (kv-head columns), x1 = Q, x2 = dO (q-head columns).   !kDKV: r1 = Q, r2 = dO, x1 = K, x2 = V.
template <int kD, typename T, int kFmt, bool kDKV>
__global__ void __launch_bounds__(kBwdThreads, 1) attn_bwd_kernel(const __grid_constant__ CUtensorMap tma_r1,
                                                                  const __grid_constant__ CUtensorMap tma_r2,
                                                                  const __grid_constant__ CUtensorMap tma_x1,
                                                                  const __grid_constant__ CUtensorMap tma_x2, BwdParams p) {
  using C = BCfg<kD, kDKV>;
  const int seq = blockIdx.z, head = blockIdx.y;  // kDKV: kv head; else q head
  const int tile = kDKV ? (int)blockIdx.x : (int)(gridDim.x - 1 - blockIdx.x);  // heaviest tiles first in both passes
  const int tok0 = p.cu_seqlens[seq];
  const int L = p.cu_seqlens[seq + 1] - tok0;
  const int r0 = tile * kR;  // first key (dKV) / first query (dQ) of this CTA inside the sequence
  if (r0 >= L) return;
  const int G = p.nq / p.nkv;
  const int nblk_all = RB_CEIL_DIV(L, kX);
  // streamed block range [blk0, blk1): dKV -> query blocks that can see these keys; dQ -> key blocks these queries can see
  const int blk0 = kDKV ? (p.causal ? r0 / kX : 0) : 0;
  const int blk1 = kDKV ? nblk_all : (p.causal ? RB_CEIL_DIV(min(L, r0 + kR), kX) : nblk_all);
  const int nblk = blk1 - blk0;
  const int n_steps = kDKV ? nblk * G : nblk;  // dKV also walks the q heads of the group (outer loop)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sR1 = ptx::smem_u32(smem);
  const uint32_t sR2 = sR1 + C::kRBytes;
  const uint32_t sXr = sR2 + C::kRBytes;            // stage b: X1 at sXr + b * 2 * kXBytes, X2 right after it
  const uint32_t sP = sXr + 4 * C::kXBytes;         // buffer b: sP + b * kPBufs * kPBytes (+ kPBytes for the second one)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kSmemData);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + B_NUM_BARS);
  float* stats = reinterpret_cast<float*>(smem + C::kSmemData + 256);  // dKV only: [2 buffers][lse | delta][64]
  auto bar = [&](int i) { return ptx::smem_u32(&bars[i]); };
  // TMEM columns: S buffers [0,64) [64,128); dP buffers [128,192) [192,256); accumulators at 256 (dV / dQ) and 384 (dK)
  constexpr uint32_t kTmS = 0, kTmDP = 128, kTmAcc1 = 256, kTmAcc2 = 384;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tma_r1);
    ptx::prefetch_tensormap(&tma_r2);
    ptx::prefetch_tensormap(&tma_x1);
    ptx::prefetch_tensormap(&tma_x2);
    for (int i = 0; i < B_NUM_BARS; ++i) {
      const bool thread_side = (i == S_EMPTY || i == S_EMPTY + 1 || i == P_FULL || i == P_FULL + 1);
      ptx::mbar_init(bar(i), thread_side ? 4 : 1);
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

  // step t -> (q head, streamed block)
  auto step_head = [&](int t) { return kDKV ? head * G + t / nblk : head; };
  auto step_blk = [&](int t) { return blk0 + (kDKV ? t % nblk : t); };
  const int r_head = head;                                   // column block of the resident tiles
  const int x_kv_head = kDKV ? head : head / G;              // dQ pass streams the kv head of this q head

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bar(R_FULL), 2 * C::kRBytes);
#pragma unroll
      for (int kb = 0; kb < C::kDBlocks; ++kb) {
        ptx::tma_load_2d(sR1 + kb * (kR * 128), &tma_r1, bar(R_FULL), r_head * kD + kb * 64, tok0 + r0);
        ptx::tma_load_2d(sR2 + kb * (kR * 128), &tma_r2, bar(R_FULL), r_head * kD + kb * 64, tok0 + r0);
      }
      for (int t = 0; t < n_steps; ++t) {
        const int b = t & 1;
        const uint32_t par = (uint32_t)(t >> 1) & 1u;
        const int row = tok0 + step_blk(t) * kX;
        const int col = (kDKV ? step_head(t) : x_kv_head) * kD;
        ptx::mbar_wait(bar(X_EMPTY + b), par ^ 1);
        ptx::mbar_arrive_expect_tx(bar(X_FULL + b), 2 * C::kXBytes);
        const uint32_t x1 = sXr + b * 2 * C::kXBytes, x2 = x1 + C::kXBytes;
#pragma unroll
        for (int kb = 0; kb < C::kDBlocks; ++kb) {
          ptx::tma_load_2d(x1 + kb * (kX * 128), &tma_x1, bar(X_FULL + b), col + kb * 64, row);
          ptx::tma_load_2d(x2 + kb * (kX * 128), &tma_x2, bar(X_FULL + b), col + kb * 64, row);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_f16(kFmt, kR, kX, 0, 0);   // resident x streamed^T, both K-major
      constexpr uint32_t idesc_a = ptx::make_idesc_f16(kFmt, kR, kD, 0, 1);   // P-type (K-major) x streamed (MN-major)
      auto issue_sdp = [&](int t) {
        const int b = t & 1;
        const uint32_t par = (uint32_t)(t >> 1) & 1u;
        ptx::mbar_wait(bar(X_FULL + b), par);
        ptx::mbar_wait(bar(S_EMPTY + b), par ^ 1);
        ptx::tc_fence_after();
        const uint32_t x1 = sXr + b * 2 * C::kXBytes, x2 = x1 + C::kXBytes;
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          const uint32_t ro = (kk >> 2) * (kR * 128) + (kk & 3) * 32, xo = (kk >> 2) * (kX * 128) + (kk & 3) * 32;
          ptx::tc_mma_f16(tmem_base + kTmS + b * 64, ptx::make_smem_desc_sw128(sR1 + ro, 16, 1024),
                          ptx::make_smem_desc_sw128(x1 + xo, 16, 1024), idesc_s, kk != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          const uint32_t ro = (kk >> 2) * (kR * 128) + (kk & 3) * 32, xo = (kk >> 2) * (kX * 128) + (kk & 3) * 32;
          ptx::tc_mma_f16(tmem_base + kTmDP + b * 64, ptx::make_smem_desc_sw128(sR2 + ro, 16, 1024),
                          ptx::make_smem_desc_sw128(x2 + xo, 16, 1024), idesc_s, kk != 0 ? 1u : 0u);
        }
        ptx::tc_commit(bar(S_FULL + b));
      };
      ptx::mbar_wait(bar(R_FULL), 0);
      issue_sdp(0);
      for (int t = 0; t < n_steps; ++t) {
        if (t + 1 < n_steps) issue_sdp(t + 1);
        const int b = t & 1;
        const uint32_t par = (uint32_t)(t >> 1) & 1u;
        ptx::mbar_wait(bar(P_FULL + b), par);
        ptx::tc_fence_after();
        const uint32_t x1 = sXr + b * 2 * C::kXBytes, x2 = x1 + C::kXBytes;
        const uint32_t pb = sP + b * C::kPBufs * C::kPBytes;
#pragma unroll
        for (int kk = 0; kk < kX / 16; ++kk) {
          const uint32_t acc = (t | kk) != 0 ? 1u : 0u;
          // A: 16 streamed rows = 32 B inside the 128 B row of the P-type buffer; B: 16 rows = 2048 B down every 64-wide chunk
          if constexpr (kDKV) {
            ptx::tc_mma_f16(tmem_base + kTmAcc1, ptx::make_smem_desc_sw128(pb + kk * 32, 16, 1024),
                            ptx::make_smem_desc_sw128(x2 + kk * 2048, kX * 128, 1024), idesc_a, acc);  // dV += P^T dO
            ptx::tc_mma_f16(tmem_base + kTmAcc2, ptx::make_smem_desc_sw128(pb + C::kPBytes + kk * 32, 16, 1024),
                            ptx::make_smem_desc_sw128(x1 + kk * 2048, kX * 128, 1024), idesc_a, acc);  // dK += dS^T Q
          } else {
            ptx::tc_mma_f16(tmem_base + kTmAcc1, ptx::make_smem_desc_sw128(pb + kk * 32, 16, 1024),
                            ptx::make_smem_desc_sw128(x1 + kk * 2048, kX * 128, 1024), idesc_a, acc);  // dQ += dS K
          }
        }
        ptx::tc_commit(bar(X_EMPTY + b));
        ptx::tc_commit(bar(P_EMPTY + b));
      }
      ptx::tc_commit(bar(ACC_FULL));
    }
  } else {
    // ===================================================== softmax-gradient threads (thread = resident row = TMEM lane)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int tid = row;  // 0..127 among these four warps (warps 2,3,4,5 own quadrants 2,3,0,1)
    const int r_pos = r0 + row;  // key (dKV) / query (dQ) position of this thread inside the sequence
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;
    const uint32_t p_row_off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    float lse_r = INFINITY, delta_r = 0.f;  // dQ pass: per-row statistics in registers
    if constexpr (!kDKV) {
      if (r_pos < L) {
        lse_r = p.lse[(int64_t)head * p.T + tok0 + r_pos] * 1.4426950408889634f;
        delta_r = p.delta[(int64_t)head * p.T + tok0 + r_pos];
      }
    }

    for (int t = 0; t < n_steps; ++t) {
      const int b = t & 1;
      const uint32_t par = (uint32_t)(t >> 1) & 1u;
      const int x0 = step_blk(t) * kX;  // first streamed position (query for dKV, key for dQ)
      if constexpr (kDKV) {
        // statistics of the 64 streamed queries, shared by all 128 threads: [lse * log2e | delta]; rows past the end of the
        // sequence get lse = +inf, which zeroes their probabilities
        const int h = step_head(t);
        const int c = tid & 63, qpos = x0 + c;
        float v = tid < 64 ? INFINITY : 0.f;
        if (qpos < L) {
          const int64_t idx = (int64_t)h * p.T + tok0 + qpos;
          v = tid < 64 ? p.lse[idx] * 1.4426950408889634f : p.delta[idx];
        }
        stats[b * 128 + tid] = v;
        softmax_bar_sync();
      }
      ptx::mbar_wait(bar(S_FULL + b), par);
      ptx::tc_fence_after();
      ptx::mbar_wait(bar(P_EMPTY + b), par ^ 1);  // the accumulating MMAs of step t-2 are done with these shared buffers
      bool need_mask;
      if constexpr (kDKV) need_mask = p.causal && (r0 + kR - 1 > x0);                    // some key > some query
      else need_mask = (x0 + kX > L) || (p.causal && x0 + kX - 1 > r0);                   // key past the end / above the diagonal
      const uint32_t pb = sP + b * C::kPBufs * C::kPBytes + p_row_off;
#pragma unroll
      for (int c = 0; c < kX / 32; ++c) {
        uint32_t rs[32], rd[32];
        ptx::tc_ld_32x32(lane_addr + kTmS + b * 64 + c * 32, rs);
        ptx::tc_ld_32x32(lane_addr + kTmDP + b * 64 + c * 32, rd);
        ptx::tc_wait_ld();
        uint32_t wp[16], wd[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float pr[2], ds[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = c * 32 + i + e;
            const float lse_c = kDKV ? stats[b * 128 + col] : lse_r;
            const float del_c = kDKV ? stats[b * 128 + 64 + col] : delta_r;
            float pv = exp2f(fmaf(__uint_as_float(rs[i + e]), sl2, -lse_c));
            if (need_mask) {
              const int xpos = x0 + col;
              const bool vis = kDKV ? (r_pos <= xpos) : (xpos < L && (!p.causal || xpos <= r_pos));
              if (!vis) pv = 0.f;
            }
            pr[e] = pv;
            ds[e] = pv * (__uint_as_float(rd[i + e]) - del_c) * p.scale;
          }
          wp[i >> 1] = pack2b<T>(pr[0], pr[1]);
          wd[i >> 1] = pack2b<T>(ds[0], ds[1]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t off = (((uint32_t)(c * 4 + q)) ^ sw) << 4;  // 16-byte piece (8 streamed positions) of this row
          if constexpr (kDKV) {
            st_shared_v4b(pb + off, wp[4 * q], wp[4 * q + 1], wp[4 * q + 2], wp[4 * q + 3]);
            st_shared_v4b(pb + C::kPBytes + off, wd[4 * q], wd[4 * q + 1], wd[4 * q + 2], wd[4 * q + 3]);
          } else {
            st_shared_v4b(pb + off, wd[4 * q], wd[4 * q + 1], wd[4 * q + 2], wd[4 * q + 3]);
          }
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(bar(S_EMPTY + b));
        ptx::mbar_arrive(bar(P_FULL + b));
      }
    }

    // ---- accumulators -> global
    ptx::mbar_wait(bar(ACC_FULL), 0);
    ptx::tc_fence_after();
    // tcgen05.ld is warp-collective: every lane executes the loads, only the stores are predicated on the row being real
    const bool row_ok = r_pos < L;
#pragma unroll
    for (int which = 0; which < (kDKV ? 2 : 1); ++which) {
      T* dst = reinterpret_cast<T*>(which == 0 ? p.out1 : p.out2) + (int64_t)(tok0 + r_pos) * (which == 0 ? p.ld1 : p.ld2) +
               (int64_t)head * kD;  // dKV: kv head; dQ: q head
#pragma unroll
      for (int c = 0; c < kD / 32; ++c) {
        uint32_t r[32];
        ptx::tc_ld_32x32(lane_addr + (which == 0 ? kTmAcc1 : kTmAcc2) + c * 32, r);
        ptx::tc_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 v;
            v.x = pack2b<T>(__uint_as_float(r[i]), __uint_as_float(r[i + 1]));
            v.y = pack2b<T>(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
            v.z = pack2b<T>(__uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
            v.w = pack2b<T>(__uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
            *reinterpret_cast<uint4*>(dst + c * 32 + i) = v;
          }
        }
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

// delta[h, t] = sum_d dO[t, h, d] * O[t, h, d]; one warp per (t, h)
template <typename T>
__global__ void attn_delta_kernel(const T* __restrict__ out, const T* __restrict__ dout, float* __restrict__ delta, int64_t o_ld,
                                  int64_t do_ld, int Tn, int nq, int D) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= Tn * nq) return;
  const int t = w / nq, h = w - t * nq;
  const T* o = out + (int64_t)t * o_ld + (int64_t)h * D;
  const T* g = dout + (int64_t)t * do_ld + (int64_t)h * D;
  float acc = 0.f;
  for (int i = lane * 2; i < D; i += 64) acc += rb::to_f(o[i]) * rb::to_f(g[i]) + rb::to_f(o[i + 1]) * rb::to_f(g[i + 1]);
  acc = rb::warp_sum(acc);
  if (lane == 0) delta[(int64_t)h * Tn + t] = acc;
}

template <int kD, typename T, int kFmt, bool kDKV>
int launch_bwd(const CUtensorMap& r1, const CUtensorMap& r2, const CUtensorMap& x1, const CUtensorMap& x2, const BwdParams& p,
               int heads, int B, int max_seqlen, cudaStream_t s) {
  using C = BCfg<kD, kDKV>;
  auto kern = attn_bwd_kernel<kD, T, kFmt, kDKV>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) != cudaSuccess) return -30;
    attr_set = true;
  }
  dim3 grid(RB_CEIL_DIV(max_seqlen, kR), heads, B);
  kern<<<grid, kBwdThreads, C::kSmemBytes, s>>>(r1, r2, x1, x2, p);
  return cudaGetLastError() == cudaSuccess ? 0 : -31;
}

template <int kD, typename T, int kFmt>
int run_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse, float* delta, void* dq,
            void* dk, void* dv, const int* cu, int64_t q_ld, int64_t k_ld, int64_t v_ld, int64_t o_ld, int64_t do_ld, int64_t dq_ld,
            int64_t dk_ld, int64_t dv_ld, int Tn, int B, int nq, int nkv, int max_seqlen, float scale, int causal, cudaStream_t s) {
  const int bf = kFmt == 1;
  {
    const int warps = Tn * nq, threads = 256;
    attn_delta_kernel<T><<<RB_CEIL_DIV(warps * 32, threads), threads, 0, s>>>(reinterpret_cast<const T*>(out),
                                                                              reinterpret_cast<const T*>(dout), delta, o_ld, do_ld,
                                                                              Tn, nq, kD);
    if (cudaGetLastError() != cudaSuccess) return -20;
  }
  CUtensorMap q128, do128, k128, v128, q64, do64, k64, v64;
  if (!make_tmap(&k128, k, bf, Tn, (uint64_t)nkv * kD, k_ld, 64, kR) || !make_tmap(&v128, v, bf, Tn, (uint64_t)nkv * kD, v_ld, 64, kR) ||
      !make_tmap(&q64, q, bf, Tn, (uint64_t)nq * kD, q_ld, 64, kX) || !make_tmap(&do64, dout, bf, Tn, (uint64_t)nq * kD, do_ld, 64, kX) ||
      !make_tmap(&q128, q, bf, Tn, (uint64_t)nq * kD, q_ld, 64, kR) || !make_tmap(&do128, dout, bf, Tn, (uint64_t)nq * kD, do_ld, 64, kR) ||
      !make_tmap(&k64, k, bf, Tn, (uint64_t)nkv * kD, k_ld, 64, kX) || !make_tmap(&v64, v, bf, Tn, (uint64_t)nkv * kD, v_ld, 64, kX))
    return -10;
  BwdParams pkv{cu, lse, delta, dv, dk, dv_ld, dk_ld, Tn, nq, nkv, scale, causal};
  int rc = launch_bwd<kD, T, kFmt, true>(k128, v128, q64, do64, pkv, nkv, B, max_seqlen, s);
  if (rc != 0) return rc;
  BwdParams pq{cu, lse, delta, dq, nullptr, dq_ld, 0, Tn, nq, nkv, scale, causal};
  return launch_bwd<kD, T, kFmt, false>(q128, do128, k64, v64, pq, nq, B, max_seqlen, s);
}

}  // namespace

// All tensors are [T, heads, hd] views with unit stride over hd and packed heads; `delta` is an [nq, T] fp32 scratch tensor.
// dt: 1 = bf16, 2 = fp16.  Returns 0 or a negative code.
extern "C" int rb_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                           float* delta, void* dq, void* dk, void* dv, const int* cu_seqlens, int64_t q_ld, int64_t k_ld,
                           int64_t v_ld, int64_t o_ld, int64_t do_ld, int64_t dq_ld, int64_t dk_ld, int64_t dv_ld, int T, int B,
                           int nq, int nkv, int hd, int max_seqlen, float scale, int causal, int dt, cudaStream_t s) {
  if (hd != 128 && hd != 64) return -1;
  if (dt != 1 && dt != 2) return -2;
  if (nq % nkv != 0 || T <= 0 || B <= 0 || max_seqlen <= 0) return -3;
  if ((q_ld | k_ld | v_ld | o_ld | do_ld | dq_ld | dk_ld | dv_ld) % 8 != 0) return -4;
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)dout | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 15) return -5;
#define RB_RUN(D, TT, F)                                                                                                          \
  return run_bwd<D, TT, F>(q, k, v, out, dout, lse, delta, dq, dk, dv, cu_seqlens, q_ld, k_ld, v_ld, o_ld, do_ld, dq_ld, dk_ld, \
                           dv_ld, T, B, nq, nkv, max_seqlen, scale, causal, s)
  if (hd == 128) {
    if (dt == 1) RB_RUN(128, __nv_bfloat16, 1);
    RB_RUN(128, __half, 0);
  }
  if (dt == 1) RB_RUN(64, __nv_bfloat16, 1);
  RB_RUN(64, __half, 0);
#undef RB_RUN
}
