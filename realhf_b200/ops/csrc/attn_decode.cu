// Single-token decode attention for sm_100a: fused RoPE + in-place KV append + split-KV online softmax.
//
// One CTA per (sequence, kv head, kv split).  The new token's q/k/v come from the fused QKV projection
// output; RoPE is applied in registers, k/v are appended to the dense cache in place, and the CTA
// streams its slice of the cache once with 16-byte loads (this op is pure HBM bandwidth: 2*S*hd*2 bytes
// per (sequence, kv head)).  A 16-lane half-warp owns one cache row at a time (16 lanes x 8 elements =
// head_dim 128; 8 lanes x 8 for head_dim 64), so a warp has two rows in flight per step and the loop is
// unrolled x4 for memory-level parallelism.  GQA groups (REP query heads per kv head) share every load.
// Replaces flash_attn_with_kvcache (reference: modules/attn.py:238-251).
#include "common.cuh"

namespace {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;

struct DecodeParams {
  const void* qkv;      // [B, (nq + 2 nkv) * hd]
  void* k_cache;        // logical [B, S, nkv, hd] with arbitrary (b, s, h) strides, unit stride on hd
  void* v_cache;
  const int* cache_lens;  // [B] tokens already cached = position of the new token
  void* out;            // [B, nq * hd]
  float* part_acc;      // [B, nq, splits, hd] (splits > 1)
  float* part_ml;       // [B, nq, splits, 2]
  const float* cos;     // [max_pos, rot/2] or nullptr
  const float* sin;
  int64_t qkv_stride;
  int64_t kb, ks, kh;   // cache strides in elements
  int64_t vb, vs, vh;
  int B, nq, nkv, S_max, splits, rot_dim, interleaved;
  float scale;
};

template <typename T> RB_DEVICE void load8(const T* p, float* f) {
  rb::Pack<T, 8> v = *reinterpret_cast<const rb::Pack<T, 8>*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = rb::to_f(v.v[i]);
}
template <typename T> RB_DEVICE void load8_stream(const T* p, float* f) {
  int4 r = rb::ld_stream(p);
  const T* v = reinterpret_cast<const T*>(&r);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = rb::to_f(v[i]);
}
template <typename T> RB_DEVICE void store8(T* p, const float* f) {
  rb::Pack<T, 8> v;
#pragma unroll
  for (int i = 0; i < 8; ++i) v.v[i] = rb::from_f<T>(f[i]);
  *reinterpret_cast<rb::Pack<T, 8>*>(p) = v;
}

// Rotate the 8 elements this lane holds of one head.  LPR = lanes per row (hd/8).
template <int HD>
RB_DEVICE void rope8(float* x, int sub, const float* cos, const float* sin, int pos, int rot_dim, int interleaved) {
  constexpr int LPR = HD / 8;
  const int half = rot_dim / 2;
  const int d0 = sub * 8;
  if (interleaved) {
    if (d0 < rot_dim) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int fi = d0 / 2 + i;
        const float c = cos[(int64_t)pos * half + fi], s = sin[(int64_t)pos * half + fi];
        const float a = x[2 * i], b = x[2 * i + 1];
        x[2 * i] = a * c - b * s;
        x[2 * i + 1] = b * c + a * s;
      }
    }
  } else {
    // partner element lives (half) dims away => (half/8) lanes away inside the row's lane group
    const int lane_off = half / 8;
    const bool first = d0 < half;
    const bool active = d0 < rot_dim;
    float partner[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float up = __shfl_down_sync(0xffffffffu, x[i], lane_off, LPR);
      const float dn = __shfl_up_sync(0xffffffffu, x[i], lane_off, LPR);
      partner[i] = first ? up : dn;
    }
    if (active) {
      const int f0 = first ? d0 : d0 - half;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float c = cos[(int64_t)pos * half + f0 + i], s = sin[(int64_t)pos * half + f0 + i];
        x[i] = first ? (x[i] * c - partner[i] * s) : (x[i] * c + partner[i] * s);
      }
    }
  }
}

template <typename T, int HD, int REP>
__global__ void __launch_bounds__(kThreads) decode_attn_kernel(DecodeParams p) {
  constexpr int LPR = HD / 8;        // lanes per cache row
  constexpr int RPW = 32 / LPR;      // rows per warp per step
  const int b = blockIdx.x, hkv = blockIdx.y, split = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane % LPR, rsel = lane / LPR;
  rb::pdl_trigger();  // successors may start launching (and prefetching weights) while this kernel still waits below
  rb::pdl_wait();
  const int pos = p.cache_lens[b];   // position of the new token; attends to [0, pos]
  const int n_ctx = pos + 1;
  const T* qkv = reinterpret_cast<const T*>(p.qkv) + (int64_t)b * p.qkv_stride;

  // ---- new token: q (REP heads), k, v with RoPE
  float q[REP][8], kn[8], vn[8];
#pragma unroll
  for (int r = 0; r < REP; ++r) {
    load8(qkv + (int64_t)(hkv * REP + r) * HD + sub * 8, q[r]);
    if (p.cos) rope8<HD>(q[r], sub, p.cos, p.sin, pos, p.rot_dim, p.interleaved);
#pragma unroll
    for (int i = 0; i < 8; ++i) q[r][i] *= p.scale;
  }
  load8(qkv + (int64_t)(p.nq + hkv) * HD + sub * 8, kn);
  if (p.cos) rope8<HD>(kn, sub, p.cos, p.sin, pos, p.rot_dim, p.interleaved);
  load8(qkv + (int64_t)(p.nq + p.nkv + hkv) * HD + sub * 8, vn);

  T* kc = reinterpret_cast<T*>(p.k_cache) + (int64_t)b * p.kb + (int64_t)hkv * p.kh;
  T* vc = reinterpret_cast<T*>(p.v_cache) + (int64_t)b * p.vb + (int64_t)hkv * p.vh;
  if (split == 0 && warp == 0 && rsel == 0 && pos < p.S_max) {
    store8(kc + (int64_t)pos * p.ks + sub * 8, kn);
    store8(vc + (int64_t)pos * p.vs + sub * 8, vn);
  }

  // ---- this split's range of *cached* positions [lo, hi) (the new token is handled from registers)
  const int per = RB_CEIL_DIV(pos, p.splits);
  const int lo = min(pos, split * per), hi = min(pos, lo + per);

  float m[REP], l[REP], acc[REP][8];
#pragma unroll
  for (int r = 0; r < REP; ++r) {
    m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[r][i] = 0.f;
  }

  // row groups run different trip counts near the range end: shuffles must name only the group's own lanes
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (rsel * LPR));
  auto consume = [&](const float* kf, const float* vf) {
#pragma unroll
    for (int r = 0; r < REP; ++r) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(q[r][i], kf[i], s);
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(gmask, s, o, LPR);
      const float mn = fmaxf(m[r], s);
      const float corr = __expf(m[r] - mn), pr = __expf(s - mn);
      l[r] = l[r] * corr + pr;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[r][i] = fmaf(acc[r][i], corr, pr * vf[i]);
      m[r] = mn;
    }
  };

  constexpr int UNROLL = 4;
  const int step = kWarps * RPW;
  int s0 = lo + warp * RPW + rsel;
  for (; s0 + (UNROLL - 1) * step < hi; s0 += UNROLL * step) {
    float kf[UNROLL][8], vf[UNROLL][8];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      load8_stream(kc + (int64_t)(s0 + u * step) * p.ks + sub * 8, kf[u]);
      load8_stream(vc + (int64_t)(s0 + u * step) * p.vs + sub * 8, vf[u]);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) consume(kf[u], vf[u]);
  }
  for (; s0 < hi; s0 += step) {
    float kf[8], vf[8];
    load8_stream(kc + (int64_t)s0 * p.ks + sub * 8, kf);
    load8_stream(vc + (int64_t)s0 * p.vs + sub * 8, vf);
    consume(kf, vf);
  }
  // the new token itself: exactly one row-group of the last split takes it
  if (split == p.splits - 1 && warp == 0 && rsel == 0) consume(kn, vn);

  __syncwarp();
  // ---- merge the RPW row-groups of each warp, then the warps
  __shared__ float sm_acc[kWarps][REP][HD];
  __shared__ float sm_m[kWarps][REP], sm_l[kWarps][REP];
#pragma unroll
  for (int r = 0; r < REP; ++r) {
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m[r], o), l2 = __shfl_xor_sync(0xffffffffu, l[r], o);
      const float mn = fmaxf(m[r], m2);
      const float c1 = (m[r] == -INFINITY) ? 0.f : __expf(m[r] - mn), c2 = (m2 == -INFINITY) ? 0.f : __expf(m2 - mn);
      l[r] = l[r] * c1 + l2 * c2;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a2 = __shfl_xor_sync(0xffffffffu, acc[r][i], o);
        acc[r][i] = acc[r][i] * c1 + a2 * c2;
      }
      m[r] = mn;
    }
    if (rsel == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sm_acc[warp][r][sub * 8 + i] = acc[r][i];
      if (sub == 0) { sm_m[warp][r] = m[r]; sm_l[warp][r] = l[r]; }
    }
  }
  __syncthreads();
  // threads cooperatively finalize: REP*HD outputs
  for (int idx = threadIdx.x; idx < REP * HD; idx += kThreads) {
    const int r = idx / HD, d = idx % HD;
    float mg = -INFINITY;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) mg = fmaxf(mg, sm_m[w][r]);
    float lg = 0.f, ag = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const float c = (sm_m[w][r] == -INFINITY) ? 0.f : __expf(sm_m[w][r] - mg);
      lg += sm_l[w][r] * c;
      ag += sm_acc[w][r][d] * c;
    }
    const int hq = hkv * REP + r;
    if (p.splits == 1) {
      reinterpret_cast<T*>(p.out)[(int64_t)b * p.nq * HD + (int64_t)hq * HD + d] = rb::from_f<T>(lg > 0.f ? ag / lg : 0.f);
    } else {
      const int64_t o = ((int64_t)b * p.nq + hq) * p.splits + split;
      p.part_acc[o * HD + d] = ag;
      if (d == 0) { p.part_ml[o * 2] = mg; p.part_ml[o * 2 + 1] = lg; }
    }
  }
}

template <typename T, int HD>
__global__ void decode_attn_reduce_kernel(const float* __restrict__ part_acc, const float* __restrict__ part_ml,
                                          T* __restrict__ out, int splits) {
  const int64_t bh = blockIdx.x;  // b * nq + hq
  const int d = threadIdx.x;
  rb::pdl_trigger();  // successors may start launching (and prefetching weights) while this kernel still waits below
  rb::pdl_wait();
  float mg = -INFINITY;
  for (int s = 0; s < splits; ++s) mg = fmaxf(mg, part_ml[(bh * splits + s) * 2]);
  float lg = 0.f, ag = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float ms = part_ml[(bh * splits + s) * 2];
    const float c = (ms == -INFINITY) ? 0.f : __expf(ms - mg);
    lg += part_ml[(bh * splits + s) * 2 + 1] * c;
    ag += part_acc[(bh * splits + s) * HD + d] * c;
  }
  out[bh * HD + d] = rb::from_f<T>(lg > 0.f ? ag / lg : 0.f);
}

template <typename T, int HD>
int launch_decode(const DecodeParams& p, cudaStream_t s) {
  const int rep = p.nq / p.nkv;
  dim3 grid(p.B, p.nkv, p.splits);
#define RB_L(REP) rb::launch_pdl(decode_attn_kernel<T, HD, REP>, grid, dim3(kThreads), 0, s, p)
  switch (rep) {
    case 1: RB_L(1); break;
    case 2: RB_L(2); break;
    case 4: RB_L(4); break;
    case 8: RB_L(8); break;
    default: return -1;
  }
#undef RB_L
  if (p.splits > 1)
    rb::launch_pdl(decode_attn_reduce_kernel<T, HD>, dim3(p.B * p.nq), dim3(HD), 0, s, (const float*)p.part_acc, (const float*)p.part_ml,
                   reinterpret_cast<T*>(p.out), p.splits);
  return 0;
}

}  // namespace

extern "C" int rb_decode_attention(const void* qkv, void* k_cache, void* v_cache, const int* cache_lens, void* out,
                                   float* part_acc, float* part_ml, const float* cos, const float* sin, int64_t qkv_stride,
                                   int64_t kb, int64_t ks, int64_t kh, int64_t vb, int64_t vs, int64_t vh, int B, int nq,
                                   int nkv, int hd, int S_max, int splits, int rot_dim, int interleaved, float scale, int dt,
                                   cudaStream_t s) {
  if (B == 0) return 0;
  DecodeParams p{qkv, k_cache, v_cache, cache_lens, out, part_acc, part_ml, cos, sin, qkv_stride, kb, ks, kh, vb, vs, vh,
                 B, nq, nkv, S_max, splits, rot_dim, interleaved, scale};
  if (dt == 1 && hd == 128) return launch_decode<__nv_bfloat16, 128>(p, s);
  if (dt == 1 && hd == 64) return launch_decode<__nv_bfloat16, 64>(p, s);
  if (dt == 2 && hd == 128) return launch_decode<__half, 128>(p, s);
  if (dt == 2 && hd == 64) return launch_decode<__half, 64>(p, s);
  return -2;
}
