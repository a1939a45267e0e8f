// torch bindings for attention kernels.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cstdlib>
#include <torch/library.h>
#include <torch/types.h>

#include <vector>

using at::Tensor;

extern "C" int rb_decode_attention(const void* qkv, void* k_cache, void* v_cache, const int* cache_lens, void* out,
                                   float* part_acc, float* part_ml, const float* cos, const float* sin, int64_t qkv_stride,
                                   int64_t kb, int64_t ks, int64_t kh, int64_t vb, int64_t vs, int64_t vh, int B, int nq,
                                   int nkv, int hd, int S_max, int splits, int rot_dim, int interleaved, float scale, int dt,
                                   cudaStream_t s);

// qkv [B, (nq+2nkv)*hd]; caches logical [B, S, nkv, hd] (any b/s/h strides, unit hd stride); cache_lens int32 [B].
Tensor decode_attention(const Tensor& qkv, Tensor k_cache, Tensor v_cache, const Tensor& cache_lens, int64_t nq, int64_t nkv,
                        int64_t hd, double scale, const c10::optional<Tensor>& cos, const c10::optional<Tensor>& sin,
                        int64_t rot_dim, bool interleaved) {
  TORCH_CHECK(qkv.is_cuda() && qkv.dim() == 2 && qkv.stride(1) == 1);
  TORCH_CHECK(k_cache.dim() == 4 && v_cache.dim() == 4 && k_cache.stride(3) == 1 && v_cache.stride(3) == 1);
  TORCH_CHECK(k_cache.scalar_type() == qkv.scalar_type() && v_cache.scalar_type() == qkv.scalar_type());
  TORCH_CHECK(cache_lens.scalar_type() == at::kInt && cache_lens.is_contiguous());
  TORCH_CHECK(qkv.size(1) == (nq + 2 * nkv) * hd && nq % nkv == 0);
  const int B = qkv.size(0);
  const int S = k_cache.size(1);
  c10::cuda::CUDAGuard guard(qkv.device());
  auto out = at::empty({B, nq * hd}, qkv.options());
  // Split the KV range of a (sequence, kv head) over several CTAs only when there are fewer than two CTAs per SM.  Measured on
  // B200 at B=16 x 32 heads (512 CTAs, ctx 384): forcing 2-3 splits made the 4-layer decode step 4% SLOWER (565 -> 588 us;
  // the reduce kernel's launch + the partial round trip cost more than the extra loads in flight buy), so the threshold stays
  // low; `REAL_DECODE_CTAS_PER_SM` overrides it for experiments.
  static const int target = [] { const char* e = getenv("REAL_DECODE_CTAS_PER_SM"); return e ? atoi(e) : 2; }();
  int splits = 1;
  const int64_t ctas = (int64_t)B * nkv;
  if (ctas < (int64_t)target * 148 && S >= 512) {
    splits = (int)std::min<int64_t>(16, ((int64_t)target * 148 + ctas - 1) / ctas);
  }
  Tensor pa, pm;
  if (splits > 1) {
    pa = at::empty({B, nq, splits, hd}, qkv.options().dtype(at::kFloat));
    pm = at::empty({B, nq, splits, 2}, qkv.options().dtype(at::kFloat));
  }
  const float *cp = nullptr, *sp = nullptr;
  if (cos.has_value()) {
    TORCH_CHECK(sin.has_value() && cos->scalar_type() == at::kFloat && cos->is_contiguous() && sin->is_contiguous());
    cp = cos->data_ptr<float>();
    sp = sin->data_ptr<float>();
  }
  const int dt = qkv.scalar_type() == at::kBFloat16 ? 1 : (qkv.scalar_type() == at::kHalf ? 2 : -1);
  int rc = rb_decode_attention(qkv.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), cache_lens.data_ptr<int>(), out.data_ptr(),
                               splits > 1 ? pa.data_ptr<float>() : nullptr, splits > 1 ? pm.data_ptr<float>() : nullptr, cp, sp,
                               qkv.stride(0), k_cache.stride(0), k_cache.stride(1), k_cache.stride(2), v_cache.stride(0),
                               v_cache.stride(1), v_cache.stride(2), B, (int)nq, (int)nkv, (int)hd, S, splits, (int)rot_dim,
                               interleaved, (float)scale, dt, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "decode_attention: unsupported configuration (", rc, ")");
  return out;
}

extern "C" int rb_sample(const void* logits, int64_t row_stride, int64_t* next_tok, float* logprob, uint8_t* mask_bits,
                         int64_t mask_stride, const bool* unfinished, int B, int V, int top_k, float top_p, float inv_temp,
                         int eos_id, int suppress_eos, int greedy, int pad_id, uint64_t seed, uint32_t step, int dt,
                         cudaStream_t s);

// logits [B, V] -> (next token int64 [B], logprob fp32 [B], mask bits uint8 [B, ceil(V/8)] or empty)
std::vector<Tensor> sample(const Tensor& logits, const c10::optional<Tensor>& unfinished, int64_t top_k, double top_p, double inv_temp,
                           int64_t eos_id, bool suppress_eos, bool greedy, int64_t pad_id, int64_t seed, int64_t step, bool want_mask) {
  TORCH_CHECK(logits.is_cuda() && logits.dim() == 2 && logits.stride(1) == 1);
  const int B = logits.size(0), V = logits.size(1);
  c10::cuda::CUDAGuard guard(logits.device());
  auto tok = at::empty({B}, logits.options().dtype(at::kLong));
  auto lp = at::empty({B}, logits.options().dtype(at::kFloat));
  Tensor mask = want_mask ? at::empty({B, (V + 7) / 8}, logits.options().dtype(at::kByte)) : at::empty({0}, logits.options().dtype(at::kByte));
  const bool* uf = nullptr;
  if (unfinished.has_value()) {
    TORCH_CHECK(unfinished->scalar_type() == at::kBool && unfinished->is_contiguous() && unfinished->numel() == B);
    uf = unfinished->data_ptr<bool>();
  }
  const int dt = logits.scalar_type() == at::kFloat ? 0 : (logits.scalar_type() == at::kBFloat16 ? 1 : (logits.scalar_type() == at::kHalf ? 2 : -1));
  int rc = rb_sample(logits.data_ptr(), logits.stride(0), tok.data_ptr<int64_t>(), lp.data_ptr<float>(),
                     want_mask ? mask.data_ptr<uint8_t>() : nullptr, want_mask ? mask.stride(0) : 0, uf, B, V, (int)top_k, (float)top_p,
                     (float)inv_temp, (int)eos_id, suppress_eos, greedy, (int)pad_id, (uint64_t)seed, (uint32_t)step, dt,
                     at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "sample: unsupported configuration (", rc, ")");
  return {tok, lp, mask};
}

extern "C" int rb_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, const int* cu_seqlens, int64_t q_ld,
                           int64_t k_ld, int64_t v_ld, int64_t out_ld, int T, int B, int nq, int nkv, int hd, int max_seqlen,
                           float scale, int causal, int dt, cudaStream_t s);

// q [T, nq, hd], k / v [T, nkv, hd]: views with unit stride over hd and heads packed (stride(1) == hd), any row pitch.
// Returns (out [T, nq, hd], lse [nq, T] fp32).
std::vector<Tensor> attn_fwd(const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& cu_seqlens, int64_t max_seqlen,
                             double scale, bool causal) {
  TORCH_CHECK(q.is_cuda() && q.dim() == 3 && k.dim() == 3 && v.dim() == 3);
  const int64_t T = q.size(0), nq = q.size(1), hd = q.size(2), nkv = k.size(1);
  for (const Tensor* t : {&q, &k, &v}) TORCH_CHECK(t->stride(2) == 1 && t->stride(1) == hd && t->size(0) == T && t->size(2) == hd);
  TORCH_CHECK(v.size(1) == nkv && k.scalar_type() == q.scalar_type() && v.scalar_type() == q.scalar_type());
  TORCH_CHECK(cu_seqlens.scalar_type() == at::kInt && cu_seqlens.is_contiguous() && cu_seqlens.is_cuda());
  c10::cuda::CUDAGuard guard(q.device());
  auto out = at::empty({T, nq, hd}, q.options());
  auto lse = at::empty({nq, T}, q.options().dtype(at::kFloat));
  const int dt = q.scalar_type() == at::kBFloat16 ? 1 : (q.scalar_type() == at::kHalf ? 2 : -1);
  int rc = rb_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr<float>(), cu_seqlens.data_ptr<int>(),
                       q.stride(0), k.stride(0), v.stride(0), out.stride(0), (int)T, (int)cu_seqlens.numel() - 1, (int)nq, (int)nkv,
                       (int)hd, (int)max_seqlen, (float)scale, causal ? 1 : 0, dt, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "attn_fwd: unsupported configuration (", rc, ")");
  return {out, lse};
}

extern "C" int rb_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                           float* delta, void* dq, void* dk, void* dv, const int* cu_seqlens, int64_t q_ld, int64_t k_ld, int64_t v_ld,
                           int64_t o_ld, int64_t do_ld, int64_t dq_ld, int64_t dk_ld, int64_t dv_ld, int T, int B, int nq, int nkv,
                           int hd, int max_seqlen, float scale, int causal, int dt, cudaStream_t s);

// Writes dq / dk / dv (views with the same layout rules as q / k / v, e.g. column ranges of one d(qkv) buffer).
void attn_bwd(const Tensor& dout, const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& out, const Tensor& lse, Tensor dq,
              Tensor dk, Tensor dv, const Tensor& cu_seqlens, int64_t max_seqlen, double scale, bool causal) {
  TORCH_CHECK(q.is_cuda() && q.dim() == 3);
  const int64_t T = q.size(0), nq = q.size(1), hd = q.size(2), nkv = k.size(1);
  for (const Tensor* t : std::initializer_list<const Tensor*>{&q, &k, &v, &out, &dout, &dq, &dk, &dv})
    TORCH_CHECK(t->dim() == 3 && t->stride(2) == 1 && t->stride(1) == hd && t->size(0) == T && t->size(2) == hd &&
                t->scalar_type() == q.scalar_type());
  TORCH_CHECK(out.size(1) == nq && dout.size(1) == nq && dq.size(1) == nq && v.size(1) == nkv && dk.size(1) == nkv && dv.size(1) == nkv);
  TORCH_CHECK(lse.scalar_type() == at::kFloat && lse.is_contiguous() && lse.size(0) == nq && lse.size(1) == T);
  TORCH_CHECK(cu_seqlens.scalar_type() == at::kInt && cu_seqlens.is_contiguous() && cu_seqlens.is_cuda());
  c10::cuda::CUDAGuard guard(q.device());
  auto delta = at::empty({nq, T}, q.options().dtype(at::kFloat));
  const int dt = q.scalar_type() == at::kBFloat16 ? 1 : (q.scalar_type() == at::kHalf ? 2 : -1);
  int rc = rb_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr<float>(),
                       delta.data_ptr<float>(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), cu_seqlens.data_ptr<int>(), q.stride(0),
                       k.stride(0), v.stride(0), out.stride(0), dout.stride(0), dq.stride(0), dk.stride(0), dv.stride(0), (int)T,
                       (int)cu_seqlens.numel() - 1, (int)nq, (int)nkv, (int)hd, (int)max_seqlen, (float)scale, causal ? 1 : 0, dt,
                       at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "attn_bwd: unsupported configuration (", rc, ")");
}

extern "C" int rb_sample_graph(const void* logits, int64_t row_stride, int64_t* tok_hist, float* lp_hist, uint8_t* mask_hist,
                               int64_t mask_bytes, bool* unfinished, int* step_rows, int64_t* input_ids, int* cache_lens, int n_gen,
                               int min_new_tokens, int B, int V, int top_k, float top_p, float inv_temp, int eos_id, int greedy,
                               int pad_id, const int64_t* seed_ptr, int dt, cudaStream_t s);

// The sampling step as the tail of a captured decode step: reads the per-row step counter on the device, writes token /
// log-prob / keep-mask into column `step` of the history buffers and prepares the next replay (input_ids, cache_lens,
// unfinished, step).  No host-side state changes between replays.
void sample_graph(const Tensor& logits, Tensor tok_hist, Tensor lp_hist, const c10::optional<Tensor>& mask_hist, Tensor unfinished,
                  Tensor step_rows, Tensor input_ids, Tensor cache_lens, const Tensor& seed, int64_t min_new_tokens, int64_t top_k,
                  double top_p, double inv_temp, int64_t eos_id, bool greedy, int64_t pad_id) {
  TORCH_CHECK(logits.is_cuda() && logits.dim() == 2 && logits.stride(1) == 1);
  const int64_t B = logits.size(0), V = logits.size(1), n_gen = tok_hist.size(1);
  TORCH_CHECK(tok_hist.is_contiguous() && tok_hist.scalar_type() == at::kLong && tok_hist.size(0) == B);
  TORCH_CHECK(lp_hist.is_contiguous() && lp_hist.scalar_type() == at::kFloat && lp_hist.size(0) == B && lp_hist.size(1) == n_gen);
  TORCH_CHECK(unfinished.scalar_type() == at::kBool && unfinished.numel() == B && step_rows.scalar_type() == at::kInt &&
              step_rows.numel() == B && input_ids.scalar_type() == at::kLong && input_ids.numel() == B &&
              cache_lens.scalar_type() == at::kInt && cache_lens.numel() == B && seed.scalar_type() == at::kLong && seed.is_cuda());
  int64_t mask_bytes = 0;
  uint8_t* mh = nullptr;
  if (mask_hist.has_value()) {
    TORCH_CHECK(mask_hist->is_contiguous() && mask_hist->scalar_type() == at::kByte && mask_hist->dim() == 3 && mask_hist->size(0) == B &&
                mask_hist->size(1) == n_gen && mask_hist->size(2) == (V + 7) / 8);
    mask_bytes = mask_hist->size(2);
    mh = mask_hist->data_ptr<uint8_t>();
  }
  c10::cuda::CUDAGuard guard(logits.device());
  const int dt = logits.scalar_type() == at::kFloat ? 0 : (logits.scalar_type() == at::kBFloat16 ? 1 : (logits.scalar_type() == at::kHalf ? 2 : -1));
  int rc = rb_sample_graph(logits.data_ptr(), logits.stride(0), tok_hist.data_ptr<int64_t>(), lp_hist.data_ptr<float>(), mh, mask_bytes,
                           unfinished.data_ptr<bool>(), step_rows.data_ptr<int>(), input_ids.data_ptr<int64_t>(), cache_lens.data_ptr<int>(),
                           (int)n_gen, (int)min_new_tokens, (int)B, (int)V, (int)top_k, (float)top_p, (float)inv_temp, (int)eos_id,
                           greedy ? 1 : 0, (int)pad_id, seed.data_ptr<int64_t>(), dt, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "sample_graph: unsupported configuration (", rc, ")");
}

void register_attn_ops(torch::Library& m) {
  m.def("sample_graph(Tensor logits, Tensor(a!) tok_hist, Tensor(b!) lp_hist, Tensor(c!)? mask_hist, Tensor(d!) unfinished, Tensor(e!) step_rows, Tensor(f!) input_ids, Tensor(g!) cache_lens, Tensor seed, int min_new_tokens, int top_k, float top_p, float inv_temp, int eos_id, bool greedy, int pad_id) -> ()", &sample_graph);
  m.def("attn_bwd(Tensor dout, Tensor q, Tensor k, Tensor v, Tensor out, Tensor lse, Tensor(a!) dq, Tensor(b!) dk, Tensor(c!) dv, Tensor cu_seqlens, int max_seqlen, float scale, bool causal) -> ()", &attn_bwd);
  m.def("attn_fwd(Tensor q, Tensor k, Tensor v, Tensor cu_seqlens, int max_seqlen, float scale, bool causal) -> Tensor[]", &attn_fwd);
  m.def("sample(Tensor logits, Tensor? unfinished, int top_k, float top_p, float inv_temp, int eos_id, bool suppress_eos, bool greedy, int pad_id, int seed, int step, bool want_mask) -> Tensor[]", &sample);
  m.def("decode_attention(Tensor qkv, Tensor(a!) k_cache, Tensor(b!) v_cache, Tensor cache_lens, int nq, int nkv, int hd, float scale, Tensor? cos, Tensor? sin, int rot_dim, bool interleaved) -> Tensor", &decode_attention);
}
