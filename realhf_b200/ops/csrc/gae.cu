// GAE advantages / returns for sm_100a.
//
// The reference runs one *thread* per sequence with a serial backward loop on the default stream
// and two host syncs (csrc/cugae/gae.cu:10-58).  GAE is the linear recurrence
//     A_t = a_t + c_t * A_{t+1}
// so here one *warp* owns a sequence and walks it backwards 32 tokens at a time, solving each
// chunk with a 5-step affine suffix scan in registers (composition (c,a)o(c',a') = (cc', a+ca')).
// Launches on the caller's stream, never synchronises with the host.
//
//   gae_1d_misalign      packed varlen: rewards [sum L_i], values [sum (L_i+1)], cu_seqlens [bs+1]
//   ppo_rewards_gae      same, but builds the PPO reward (-kl_ctl*(logp-ref_logp) + clipped score on
//                        the last token) in the same pass and also emits kl_rewards
//   gae_2d               padded [bs,T] with done / truncate flags ("olp" and "nolp" semantics of the
//                        reference's pygae2d_* functions, realhf/impl/model/utils/ppo_functional.py:312-399)
#include "common.cuh"

namespace {

struct Affine { float c, a; };  // x -> a + c*x

// Suffix scan over the 32 lanes: lane l gets the composition f_l o f_{l+1} o ... o f_31.
RB_DEVICE Affine warp_suffix_scan(Affine f) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float c2 = __shfl_down_sync(0xffffffffu, f.c, d);
    float a2 = __shfl_down_sync(0xffffffffu, f.a, d);
    if (lane + d < 32) { f.a = fmaf(f.c, a2, f.a); f.c *= c2; }
  }
  return f;
}

// Generic backward walk of one sequence of length L by one warp.  `coef(t)` returns the affine map
// of step t; `emit(t, A)` stores the result.
template <typename Coef, typename Emit>
RB_DEVICE void warp_backward_scan(int L, Coef coef, Emit emit) {
  const int lane = threadIdx.x & 31;
  float carry = 0.f;
  for (int hi = L; hi > 0; hi -= 32) {
    const int t = hi - 32 + lane;  // lanes map to increasing t; chunk covers [hi-32, hi)
    Affine f{1.f, 0.f};            // identity for t < 0 (only in the first chunk of the sequence)
    if (t >= 0) f = coef(t);
    Affine s = warp_suffix_scan(f);
    const float A = fmaf(s.c, carry, s.a);
    if (t >= 0) emit(t, A);
    carry = __shfl_sync(0xffffffffu, A, 0);
    // lane 0 of a partial chunk (t<0) carries identity o ... which equals the value at t=0; unused after.
  }
}

template <bool kFusedReward>
__global__ void __launch_bounds__(128) gae_1d_kernel(
    const float* __restrict__ rewards,      // [total]      (unused when kFusedReward)
    const float* __restrict__ logp,         // [total]      (kFusedReward)
    const float* __restrict__ ref_logp,     // [total]      (kFusedReward)
    const float* __restrict__ scores,       // [bs]         (kFusedReward)
    const float* __restrict__ values,       // [total + bs]
    const int* __restrict__ cu_seqlens,     // [bs + 1] over rewards
    const bool* __restrict__ bootstrap,     // [bs]  true -> keep V_{L}
    float* __restrict__ adv, float* __restrict__ ret, float* __restrict__ kl_rewards, float* __restrict__ tot_rewards,
    int bs, float gamma, float lam, float kl_ctl, float clip_reward) {
  const int seq = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (seq >= bs) return;
  const int r0 = cu_seqlens[seq], L = cu_seqlens[seq + 1] - r0;
  const float* v = values + r0 + seq;
  const bool boot = bootstrap[seq];
  float score = 0.f;
  if constexpr (kFusedReward) {
    score = fminf(fmaxf(scores[seq], -clip_reward), clip_reward);
    if (boot) score = 0.f;  // sequence was truncated (no EOS): no terminal reward
  }
  const float gl = gamma * lam;
  warp_backward_scan(
      L,
      [&](int t) {
        float r;
        if constexpr (kFusedReward) {
          const float kl = -kl_ctl * (logp[r0 + t] - ref_logp[r0 + t]);
          kl_rewards[r0 + t] = kl;
          r = kl + (t == L - 1 ? score : 0.f);
          tot_rewards[r0 + t] = r;
        } else {
          r = rewards[r0 + t];
        }
        float nv = v[t + 1];
        if (t == L - 1 && !boot) nv = 0.f;
        return Affine{gl, r + gamma * nv - v[t]};
      },
      [&](int t, float A) {
        adv[r0 + t] = A;
        ret[r0 + t] = A + v[t];
      });
}

// mode 0: "olp"  gae = (delta + gamma*lam*(1-done[t+1]) * gae) * (1 - trunc[t+1]), next value masked by done
// mode 1: "nolp" gae = delta + gamma*lam*(1-reset[t+1])*(1-trunc[t+1]) * gae,      next value masked by reset
__global__ void __launch_bounds__(128) gae_2d_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                     const bool* __restrict__ dones, const bool* __restrict__ truncs,
                                                     float* __restrict__ adv, float* __restrict__ ret, int bs, int T,
                                                     float gamma, float lam, int mode) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= bs) return;
  const float* r = rewards + (size_t)row * T;
  const float* v = values + (size_t)row * (T + 1);
  const bool* d = dones + (size_t)row * (T + 1);
  const bool* tr = truncs + (size_t)row * (T + 1);
  float* ao = adv + (size_t)row * T;
  float* ro = ret + (size_t)row * T;
  warp_backward_scan(
      T,
      [&](int t) {
        const float nd = d[t + 1] ? 0.f : 1.f, nt = tr[t + 1] ? 0.f : 1.f;
        const float delta = r[t] + gamma * v[t + 1] * nd - v[t];
        if (mode == 0) return Affine{gamma * lam * nd * nt, delta * nt};
        return Affine{gamma * lam * nd * nt, delta};
      },
      [&](int t, float A) {
        ao[t] = A;
        ro[t] = A + v[t];
      });
}

}  // namespace

extern "C" {

void rb_gae_1d_misalign(const float* rewards, const float* values, const int* cu_seqlens, const bool* bootstrap,
                        float* adv, float* ret, int bs, float gamma, float lam, cudaStream_t stream) {
  if (bs == 0) return;
  const int wpb = 4;
  gae_1d_kernel<false><<<RB_CEIL_DIV(bs, wpb), wpb * 32, 0, stream>>>(
      rewards, nullptr, nullptr, nullptr, values, cu_seqlens, bootstrap, adv, ret, nullptr, nullptr, bs, gamma, lam, 0.f, 0.f);
}

void rb_ppo_rewards_gae(const float* logp, const float* ref_logp, const float* scores, const float* values,
                        const int* cu_seqlens, const bool* no_eos, float* adv, float* ret, float* kl_rewards,
                        float* tot_rewards, int bs, float gamma, float lam, float kl_ctl, float clip_reward,
                        cudaStream_t stream) {
  if (bs == 0) return;
  const int wpb = 4;
  gae_1d_kernel<true><<<RB_CEIL_DIV(bs, wpb), wpb * 32, 0, stream>>>(
      nullptr, logp, ref_logp, scores, values, cu_seqlens, no_eos, adv, ret, kl_rewards, tot_rewards, bs, gamma, lam,
      kl_ctl, clip_reward);
}

void rb_gae_2d(const float* rewards, const float* values, const bool* dones, const bool* truncs, float* adv, float* ret,
               int bs, int T, float gamma, float lam, int mode, cudaStream_t stream) {
  if (bs == 0 || T == 0) return;
  const int wpb = 4;
  gae_2d_kernel<<<RB_CEIL_DIV(bs, wpb), wpb * 32, 0, stream>>>(rewards, values, dones, truncs, adv, ret, bs, T, gamma,
                                                               lam, mode);
}

}  // extern "C"
