// Fused sampling step for generation: temperature, min-length EOS suppression, top-k, top-p, categorical sample,
// log-prob of the sample, and the bit-packed "filtered" mask — one CTA per sequence, one pass over global memory.
//
// Replaces the eager chain of the reference's `genstep` (fp32 cast, topk, sort, softmax, cumsum, scatter,
// Categorical, gather, mask compare: nn/real_llm_generate.py:26-141, utils/logits_warper.py) — about 1 ms per decode
// step at [128, 32000] — with radix selection in shared memory:
//   * the row is staged once in shared memory as fp32 (V * 4 bytes <= 200 KB);
//   * top-k threshold = k-th largest key by a 4 x 8-bit radix descent over per-warp histograms;
//   * top-p threshold = the same descent with probability mass instead of counts (keep x iff the mass of strictly
//     larger logits is < top_p * total);
//   * the sample is drawn by inverse CDF in index order over the kept set with a block scan (counter-based RNG).
// Ties at a threshold are all kept (same rule as `logits < kth` in the PyTorch reference).
#include "common.cuh"

namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;

RB_DEVICE uint32_t f2key(float x) {  // order-preserving float -> uint
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
RB_DEVICE uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

struct SampleParams {
  const void* logits;
  int64_t row_stride;
  int64_t* next_tok;
  float* logprob;
  uint8_t* mask_bits;
  int64_t mask_stride;
  const bool* unfinished;   // [B] or nullptr: finished rows emit pad / logprob 0
  int V, top_k, eos_id, suppress_eos, greedy, pad_id;
  float inv_temp, top_p;
  uint64_t seed;
  uint32_t step;
  // ---- in-graph mode (step_rows != nullptr): the kernel is the tail of the captured decode step.  The step index lives on the
  // device (one counter per row, advanced here), outputs go to column `step` of [B, n_gen] history buffers, and the kernel
  // also feeds the next replay: next token -> input_ids, cache_lens += 1, unfinished &= (token != eos).
  int* step_rows;
  int64_t* input_ids;
  int* cache_lens;
  bool* unfinished_rw;
  int n_gen, min_new_tokens;
  const int64_t* seed_ptr;   // device-resident seed (a replayed graph must not repeat the random stream of the previous call)
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) sample_kernel(SampleParams p) {
  extern __shared__ float srow[];                       // [V] scaled logits
  __shared__ uint32_t hist[256];
  __shared__ float histm[256];
  __shared__ float red[32];
  __shared__ uint32_t sh_u[4];
  __shared__ float sh_f[4];
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* src = reinterpret_cast<const T*>(p.logits) + (int64_t)row * p.row_stride;
  const int V = p.V;
  const bool in_graph = p.step_rows != nullptr;
  if (in_graph) {
    const int st = min(p.step_rows[row], p.n_gen - 1);
    p.step = (uint32_t)st;
    if (p.seed_ptr != nullptr) p.seed = (uint64_t)*p.seed_ptr;
    p.suppress_eos = (p.eos_id >= 0 && st < p.min_new_tokens) ? 1 : 0;
    p.next_tok += st;                                    // row stride of the history buffers is n_gen (set by the host)
    p.logprob += st;
    if (p.mask_bits != nullptr) p.mask_bits += (int64_t)st * (p.mask_stride / p.n_gen);
    __syncthreads();                                     // everybody has read the counter before thread 0 advances it
  }

  // ---- stage + max
  float mx = -INFINITY;
  for (int j = tid; j < V; j += kThreads) {
    float x = rb::to_f(src[j]) * p.inv_temp;
    if (p.suppress_eos && j == p.eos_id) x = -INFINITY;
    srow[j] = x;
    mx = fmaxf(mx, x);
  }
  mx = rb::block_reduce<true>(mx, red);

  int64_t chosen = 0;
  float chosen_lp = 0.f;
  uint32_t keep_key = 0;  // keep elements with key >= keep_key

  if (p.greedy) {
    // argmax (lowest index among ties)
    int best = V;
    for (int j = tid; j < V; j += kThreads) if (srow[j] == mx) best = min(best, j);
    __syncthreads();
    // block min via shuffles + smem
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) sh_u[0] = 0xffffffffu;
    __syncthreads();
    if (lane == 0) atomicMin(&sh_u[0], (uint32_t)best);
    __syncthreads();
    chosen = sh_u[0];
    float s = 0.f;
    for (int j = tid; j < V; j += kThreads) s += __expf(srow[j] - mx);
    s = rb::block_reduce<false>(s, red);
    chosen_lp = -__logf(s);
    keep_key = 0;
  } else {
    // ---- top-k: k-th largest key by radix descent
    uint32_t prefix = 0, maskbits = 0;
    if (p.top_k < V) {
      uint32_t remaining = (uint32_t)p.top_k;
      for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int b = tid; b < 256; b += kThreads) hist[b] = 0;
        __syncthreads();
        for (int j = tid; j < V; j += kThreads) {
          const uint32_t k = f2key(srow[j]);
          if ((k & maskbits) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
          uint32_t cum = 0;
          int b = 255;
          for (; b > 0; --b) {
            if (cum + hist[b] >= remaining) break;
            cum += hist[b];
          }
          sh_u[0] = (uint32_t)b;
          sh_u[1] = remaining - cum;
        }
        __syncthreads();
        prefix |= sh_u[0] << shift;
        maskbits |= 255u << shift;
        remaining = sh_u[1];
        __syncthreads();
      }
      keep_key = prefix;
    }
    // ---- softmax mass of the top-k set
    float tot = 0.f;
    for (int j = tid; j < V; j += kThreads) {
      const float x = srow[j];
      if (f2key(x) >= keep_key) tot += __expf(x - mx);
    }
    tot = rb::block_reduce<false>(tot, red);
    // ---- top-p: weighted radix descent inside the top-k set
    if (p.top_p < 1.f) {
      const float P = p.top_p * tot;
      float above = 0.f;
      prefix = 0; maskbits = 0;
      for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int b = tid; b < 256; b += kThreads) histm[b] = 0.f;
        __syncthreads();
        for (int j = tid; j < V; j += kThreads) {
          const float x = srow[j];
          const uint32_t k = f2key(x);
          if (k >= keep_key && (k & maskbits) == prefix) atomicAdd(&histm[(k >> shift) & 255u], __expf(x - mx));
        }
        __syncthreads();
        if (tid == 0) {
          float cum = above;
          int b = 255;
          for (; b > 0; --b) {
            if (cum + histm[b] >= P) break;
            cum += histm[b];
          }
          sh_u[0] = (uint32_t)b;
          sh_f[0] = cum;
        }
        __syncthreads();
        prefix |= sh_u[0] << shift;
        maskbits |= 255u << shift;
        above = sh_f[0];
        __syncthreads();
      }
      keep_key = max(keep_key, prefix);
    }
    // ---- kept mass, then inverse-CDF sample in index order
    const int per = (V + kThreads - 1) / kThreads;
    const int j0 = tid * per, j1 = min(V, j0 + per);
    float local = 0.f;
    for (int j = j0; j < j1; ++j) {
      const float x = srow[j];
      if (f2key(x) >= keep_key) local += __expf(x - mx);
    }
    // block exclusive scan of `local`
    float incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    __syncthreads();
    if (lane == 31) red[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      float w = lane < kWarps ? red[lane] : 0.f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      red[lane] = w;  // inclusive warp totals
    }
    __syncthreads();
    const float warp_base = warp == 0 ? 0.f : red[warp - 1];
    const float kept = red[kWarps - 1];
    const float excl = warp_base + incl - local;
    const uint32_t r = mix32((uint32_t)p.seed ^ mix32((uint32_t)(p.seed >> 32) + 0x9e3779b9u * (uint32_t)row) ^ mix32(p.step * 0x85ebca6bu + 1u));
    const float u = ((r >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float target = u * kept;
    if (tid == 0) sh_u[2] = 0xffffffffu;
    __syncthreads();
    if (local > 0.f && target >= excl && target < excl + local) {
      float c = excl;
      int pick = -1;
      for (int j = j0; j < j1; ++j) {
        const float x = srow[j];
        if (f2key(x) >= keep_key) {
          c += __expf(x - mx);
          pick = j;
          if (target < c) break;
        }
      }
      atomicMin(&sh_u[2], (uint32_t)pick);
    }
    __syncthreads();
    if (sh_u[2] == 0xffffffffu) {  // rounding at the very end of the CDF: take the last kept element
      int last = -1;
      for (int j = j1 - 1; j >= j0; --j) if (f2key(srow[j]) >= keep_key) { last = j; break; }
      if (tid == 0) sh_u[3] = 0;
      __syncthreads();
      if (last >= 0) atomicMax(&sh_u[3], (uint32_t)last);
      __syncthreads();
      chosen = sh_u[3];
    } else {
      chosen = sh_u[2];
    }
    chosen_lp = srow[chosen] - mx - __logf(kept);
  }

  const bool live = p.unfinished == nullptr || p.unfinished[row];
  const int64_t out_stride = in_graph ? p.n_gen : 1;
  __syncthreads();  // all reads of unfinished[row] precede its update
  if (tid == 0) {
    const int64_t tok = live ? chosen : (int64_t)p.pad_id;
    p.next_tok[row * out_stride] = tok;
    p.logprob[row * out_stride] = live ? chosen_lp : 0.f;
    if (in_graph) {
      p.input_ids[row] = tok;
      p.cache_lens[row] += 1;
      p.step_rows[row] = (int)p.step + 1;
      if (p.eos_id >= 0) p.unfinished_rw[row] = live && tok != (int64_t)p.eos_id;
    }
  }
  if (p.mask_bits != nullptr) {
    uint8_t* mrow = p.mask_bits + (int64_t)row * p.mask_stride;
    const int nbytes = (V + 7) / 8;
    for (int b = tid; b < nbytes; b += kThreads) {
      uint32_t bits = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = b * 8 + i;
        if (j < V && f2key(srow[j]) < keep_key) bits |= 1u << i;
      }
      mrow[b] = (uint8_t)bits;
    }
  }
}

int launch_sample(const SampleParams& p, int B, int dt, cudaStream_t s) {
  const size_t smem = (size_t)p.V * sizeof(float);
#define RB_GO(T)                                                                                                  \
  {                                                                                                               \
    static bool cfgd = false;                                                                                     \
    if (!cfgd) {                                                                                                  \
      if (cudaFuncSetAttribute(sample_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -2; \
      cfgd = true;                                                                                                \
    }                                                                                                             \
    sample_kernel<T><<<B, kThreads, smem, s>>>(p);                                                                \
  }
  if (dt == 0) RB_GO(float) else if (dt == 1) RB_GO(__nv_bfloat16) else if (dt == 2) RB_GO(__half) else return -3;
#undef RB_GO
  return 0;
}


}  // namespace

extern "C" int rb_sample(const void* logits, int64_t row_stride, int64_t* next_tok, float* logprob, uint8_t* mask_bits,
                         int64_t mask_stride, const bool* unfinished, int B, int V, int top_k, float top_p, float inv_temp,
                         int eos_id, int suppress_eos, int greedy, int pad_id, uint64_t seed, uint32_t step, int dt,
                         cudaStream_t s) {
  if (B == 0) return 0;
  const size_t smem = (size_t)V * sizeof(float);
  if (smem > 200 * 1024) return -1;
  SampleParams p{logits, row_stride, next_tok, logprob, mask_bits, mask_stride, unfinished, V, top_k, eos_id, suppress_eos, greedy,
                 pad_id, inv_temp, top_p, seed, step, nullptr, nullptr, nullptr, nullptr, 1, 0, nullptr};
  return launch_sample(p, B, dt, s);
}

// In-graph variant: see SampleParams.  tok_hist / lp_hist are [B, n_gen], mask_hist [B, n_gen, mask_bytes] (or null).
extern "C" int rb_sample_graph(const void* logits, int64_t row_stride, int64_t* tok_hist, float* lp_hist, uint8_t* mask_hist,
                               int64_t mask_bytes, bool* unfinished, int* step_rows, int64_t* input_ids, int* cache_lens, int n_gen,
                               int min_new_tokens, int B, int V, int top_k, float top_p, float inv_temp, int eos_id, int greedy,
                               int pad_id, const int64_t* seed_ptr, int dt, cudaStream_t s) {
  if (B == 0) return 0;
  if ((size_t)V * sizeof(float) > 200 * 1024) return -1;
  SampleParams p{logits, row_stride, tok_hist, lp_hist, mask_hist, mask_bytes * n_gen, unfinished, V, top_k, eos_id, 0, greedy,
                 pad_id, inv_temp, top_p, 0, 0, step_rows, input_ids, cache_lens, unfinished, n_gen, min_new_tokens, seed_ptr};
  return launch_sample(p, B, dt, s);
}

