// torch.ops.realhf_b200.* registrations.  Compiled with g++ (no nvcc): every kernel lives in a .cu file
// behind an `extern "C"` launcher taking raw pointers and the *current* torch CUDA stream.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

#include <vector>

using at::Tensor;

#define CHECK_CUDA(x) TORCH_CHECK((x).is_cuda(), #x " must be a CUDA tensor")
#define CHECK_CONTIG(x) TORCH_CHECK((x).is_contiguous(), #x " must be contiguous")
#define CHECK_IN(x) \
  CHECK_CUDA(x);    \
  CHECK_CONTIG(x)

static inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }
static inline int dt_code(const Tensor& t) {
  switch (t.scalar_type()) {
    case at::kFloat: return 0;
    case at::kBFloat16: return 1;
    case at::kHalf: return 2;
    default: TORCH_CHECK(false, "unsupported dtype ", t.scalar_type());
  }
}
template <typename T> static inline T* optp(const c10::optional<Tensor>& t) {
  return t.has_value() ? reinterpret_cast<T*>(t->data_ptr()) : nullptr;
}

extern "C" {
void rb_set_pdl(int on);
int rb_get_pdl();
void rb_gae_1d_misalign(const float*, const float*, const int*, const bool*, float*, float*, int, float, float, cudaStream_t);
void rb_ppo_rewards_gae(const float*, const float*, const float*, const float*, const int*, const bool*, float*, float*,
                        float*, float*, int, float, float, float, float, cudaStream_t);
void rb_gae_2d(const float*, const float*, const bool*, const bool*, float*, float*, int, int, float, float, int, cudaStream_t);
void rb_segment_copy(const void*, void*, const int64_t*, const int64_t*, const int64_t*, int, int64_t, float, int, cudaStream_t);
int rb_adamw(void*, int, const void*, int, void*, void*, int, float*, int64_t, float, float, float, float, float, int,
             const float*, const int*, int, uint32_t, cudaStream_t);
int rb_sumsq(const void*, int, int64_t, float*, cudaStream_t);
int rb_rmsnorm_fwd(const void*, const void*, const void*, void*, void*, float*, int64_t, int, float, float, int, cudaStream_t);
int rb_rmsnorm_bwd_num_partials();
int rb_layernorm_fwd(const void*, const void*, const void*, void*, float*, float*, int64_t, int, float, int, cudaStream_t);
int rb_layernorm_bwd_num_partials();
int rb_layernorm_bwd(const void*, const void*, const void*, const float*, const float*, void*, float*, float*, void*, void*, int64_t, int,
                     int, cudaStream_t);
int rb_rmsnorm_bwd(const void*, const void*, const void*, const float*, void*, float*, void*, int64_t, int, float, int, const void*, cudaStream_t);
int rb_rope_inplace(void*, const float*, const float*, const int*, int64_t, int, int, int64_t, int, int, int, int, cudaStream_t);
int rb_gated_act_fwd(const void*, void*, int64_t, int, int, int, cudaStream_t);
int rb_gated_act_bwd(const void*, const void*, void*, int64_t, int, int, int, cudaStream_t);
int rb_logprob_fwd(const void*, const int64_t*, const uint8_t*, int64_t, float*, float*, float*, float*, float*, int64_t,
                   int, int64_t, float, int, int, cudaStream_t);
int rb_logprob_bwd(void*, const int64_t*, const uint8_t*, int64_t, const float*, const float*, int64_t, int, int64_t, float,
                   int, int, cudaStream_t);
}

// ---------------------------------------------------------------- GAE
std::vector<Tensor> gae_1d_misalign(const Tensor& rewards, const Tensor& values, const Tensor& cu_seqlens,
                                    const Tensor& bootstrap, double gamma, double lam) {
  CHECK_IN(rewards); CHECK_IN(values); CHECK_IN(cu_seqlens); CHECK_IN(bootstrap);
  TORCH_CHECK(rewards.scalar_type() == at::kFloat && values.scalar_type() == at::kFloat, "fp32 only");
  TORCH_CHECK(cu_seqlens.scalar_type() == at::kInt && bootstrap.scalar_type() == at::kBool);
  const int bs = cu_seqlens.numel() - 1;
  TORCH_CHECK(values.numel() == rewards.numel() + bs, "values must hold one extra entry per sequence");
  c10::cuda::CUDAGuard g(rewards.device());
  auto adv = at::empty_like(rewards), ret = at::empty_like(rewards);
  rb_gae_1d_misalign(rewards.data_ptr<float>(), values.data_ptr<float>(), cu_seqlens.data_ptr<int>(),
                     bootstrap.data_ptr<bool>(), adv.data_ptr<float>(), ret.data_ptr<float>(), bs, gamma, lam, cur_stream());
  return {adv, ret};
}

std::vector<Tensor> ppo_rewards_gae(const Tensor& logp, const Tensor& ref_logp, const Tensor& scores, const Tensor& values,
                                    const Tensor& cu_seqlens, const Tensor& no_eos, double gamma, double lam, double kl_ctl,
                                    double clip_reward) {
  CHECK_IN(logp); CHECK_IN(ref_logp); CHECK_IN(scores); CHECK_IN(values); CHECK_IN(cu_seqlens); CHECK_IN(no_eos);
  TORCH_CHECK(logp.scalar_type() == at::kFloat && ref_logp.scalar_type() == at::kFloat && scores.scalar_type() == at::kFloat &&
              values.scalar_type() == at::kFloat, "fp32 only");
  TORCH_CHECK(cu_seqlens.scalar_type() == at::kInt && no_eos.scalar_type() == at::kBool);
  const int bs = cu_seqlens.numel() - 1;
  TORCH_CHECK(values.numel() == logp.numel() + bs && ref_logp.numel() == logp.numel() && scores.numel() == bs);
  c10::cuda::CUDAGuard g(logp.device());
  auto adv = at::empty_like(logp), ret = at::empty_like(logp), kl = at::empty_like(logp), tot = at::empty_like(logp);
  rb_ppo_rewards_gae(logp.data_ptr<float>(), ref_logp.data_ptr<float>(), scores.data_ptr<float>(), values.data_ptr<float>(),
                     cu_seqlens.data_ptr<int>(), no_eos.data_ptr<bool>(), adv.data_ptr<float>(), ret.data_ptr<float>(),
                     kl.data_ptr<float>(), tot.data_ptr<float>(), bs, gamma, lam, kl_ctl, clip_reward, cur_stream());
  return {adv, ret, kl, tot};
}

std::vector<Tensor> gae_2d(const Tensor& rewards, const Tensor& values, const Tensor& dones, const Tensor& truncs,
                           double gamma, double lam, int64_t mode) {
  CHECK_IN(rewards); CHECK_IN(values); CHECK_IN(dones); CHECK_IN(truncs);
  const int bs = rewards.size(0), T = rewards.size(1);
  TORCH_CHECK(values.size(0) == bs && values.size(1) == T + 1 && dones.size(1) == T + 1 && truncs.size(1) == T + 1);
  TORCH_CHECK(rewards.scalar_type() == at::kFloat && values.scalar_type() == at::kFloat);
  TORCH_CHECK(dones.scalar_type() == at::kBool && truncs.scalar_type() == at::kBool);
  c10::cuda::CUDAGuard g(rewards.device());
  auto adv = at::empty_like(rewards), ret = at::empty_like(rewards);
  rb_gae_2d(rewards.data_ptr<float>(), values.data_ptr<float>(), dones.data_ptr<bool>(), truncs.data_ptr<bool>(),
            adv.data_ptr<float>(), ret.data_ptr<float>(), bs, T, gamma, lam, (int)mode, cur_stream());
  return {adv, ret};
}

// ---------------------------------------------------------------- segment copy
// dst_ptr == 0 -> use `dst` tensor; otherwise a raw (possibly peer-mapped) device address.
void segment_copy(const Tensor& src, const Tensor& dst, int64_t dst_ptr, const Tensor& src_off, const Tensor& dst_off,
                  const Tensor& cum, int64_t total_bytes, double eta, bool use_ema) {
  CHECK_CUDA(src); CHECK_IN(src_off); CHECK_IN(dst_off); CHECK_IN(cum);
  TORCH_CHECK(src_off.scalar_type() == at::kLong && dst_off.scalar_type() == at::kLong && cum.scalar_type() == at::kLong);
  const int n = src_off.numel();
  TORCH_CHECK(dst_off.numel() == n && cum.numel() == n + 1);
  c10::cuda::CUDAGuard g(src.device());
  void* d = dst_ptr ? reinterpret_cast<void*>(dst_ptr) : dst.data_ptr();
  rb_segment_copy(src.data_ptr(), d, src_off.data_ptr<int64_t>(), dst_off.data_ptr<int64_t>(), cum.data_ptr<int64_t>(), n,
                  total_bytes, eta, use_ema ? 1 : 0, cur_stream());
}

// ---------------------------------------------------------------- optimizer
void adamw_step(Tensor p, const Tensor& g, Tensor m, Tensor v, const c10::optional<Tensor>& master, double lr, double b1,
                double b2, double eps, double wd, int64_t step, const c10::optional<Tensor>& scale,
                const c10::optional<Tensor>& skip, bool stochastic, int64_t seed) {
  CHECK_IN(p); CHECK_IN(g); CHECK_IN(m); CHECK_IN(v);
  const int64_t n = p.numel();
  TORCH_CHECK(g.numel() == n && m.numel() == n && v.numel() == n);
  TORCH_CHECK(m.scalar_type() == v.scalar_type());
  if (master.has_value()) TORCH_CHECK(master->scalar_type() == at::kFloat && master->numel() == n);
  if (scale.has_value()) TORCH_CHECK(scale->scalar_type() == at::kFloat);
  if (skip.has_value()) TORCH_CHECK(skip->scalar_type() == at::kInt);
  c10::cuda::CUDAGuard guard(p.device());
  int rc = rb_adamw(p.data_ptr(), dt_code(p), g.data_ptr(), dt_code(g), m.data_ptr(), v.data_ptr(), dt_code(m),
                    optp<float>(master), n, lr, b1, b2, eps, wd, (int)step, optp<const float>(scale), optp<const int>(skip),
                    stochastic ? 1 : 0, (uint32_t)seed, cur_stream());
  TORCH_CHECK(rc == 0, "adamw: unsupported dtype combination");
}

void sumsq_accum(const Tensor& g, Tensor out2) {
  CHECK_IN(g); CHECK_IN(out2);
  TORCH_CHECK(out2.scalar_type() == at::kFloat && out2.numel() >= 2);
  c10::cuda::CUDAGuard guard(g.device());
  TORCH_CHECK(rb_sumsq(g.data_ptr(), dt_code(g), g.numel(), out2.data_ptr<float>(), cur_stream()) == 0);
}

// ---------------------------------------------------------------- norm
std::vector<Tensor> rmsnorm_fwd(const Tensor& x, const c10::optional<Tensor>& residual, const Tensor& w, double eps,
                                double w_offset) {
  CHECK_IN(x); CHECK_IN(w);
  const int H = x.size(-1);
  const int64_t rows = x.numel() / H;
  c10::cuda::CUDAGuard guard(x.device());
  auto y = at::empty_like(x);
  auto rstd = at::empty({rows}, x.options().dtype(at::kFloat));
  Tensor res_out;
  const void* rin = nullptr;
  void* rout = nullptr;
  if (residual.has_value()) {
    CHECK_IN(*residual);
    res_out = at::empty_like(x);
    rin = residual->data_ptr();
    rout = res_out.data_ptr();
  }
  int rc = rb_rmsnorm_fwd(x.data_ptr(), rin, w.data_ptr(), y.data_ptr(), rout, rstd.data_ptr<float>(), rows, H, eps, w_offset,
                          dt_code(x), cur_stream());
  TORCH_CHECK(rc == 0, "rmsnorm_fwd: unsupported shape/dtype");
  if (residual.has_value()) return {y, rstd, res_out};
  return {y, rstd};
}

std::vector<Tensor> rmsnorm_bwd(const Tensor& x, const Tensor& w, const Tensor& dy, const Tensor& rstd, double w_offset,
                                const c10::optional<Tensor>& dres) {
  CHECK_IN(x); CHECK_IN(w); CHECK_IN(dy); CHECK_IN(rstd);
  if (dres.has_value()) { CHECK_IN((*dres)); TORCH_CHECK(dres->numel() == x.numel() && dres->scalar_type() == x.scalar_type()); }
  const int H = x.size(-1);
  const int64_t rows = x.numel() / H;
  c10::cuda::CUDAGuard guard(x.device());
  auto dx = at::empty_like(x);
  auto dw = at::empty_like(w);
  auto partial = at::empty({rb_rmsnorm_bwd_num_partials(), H}, x.options().dtype(at::kFloat));
  int rc = rb_rmsnorm_bwd(x.data_ptr(), w.data_ptr(), dy.data_ptr(), rstd.data_ptr<float>(), dx.data_ptr(),
                          partial.data_ptr<float>(), dw.data_ptr(), rows, H, w_offset, dt_code(x),
                          dres.has_value() ? dres->data_ptr() : nullptr, cur_stream());
  TORCH_CHECK(rc == 0, "rmsnorm_bwd: unsupported shape/dtype");
  return {dx, dw};
}

// y, mean, rstd
std::vector<Tensor> layernorm_fwd(const Tensor& x, const Tensor& w, const c10::optional<Tensor>& b, double eps) {
  CHECK_IN(x); CHECK_IN(w);
  const int H = x.size(-1);
  const int64_t rows = x.numel() / H;
  TORCH_CHECK(w.numel() == H && w.scalar_type() == x.scalar_type() && (!b.has_value() || (b->numel() == H && b->scalar_type() == x.scalar_type())));
  c10::cuda::CUDAGuard guard(x.device());
  auto y = at::empty_like(x);
  auto mean = at::empty({rows}, x.options().dtype(at::kFloat));
  auto rstd = at::empty({rows}, x.options().dtype(at::kFloat));
  int rc = rb_layernorm_fwd(x.data_ptr(), w.data_ptr(), b.has_value() ? b->data_ptr() : nullptr, y.data_ptr(), mean.data_ptr<float>(),
                            rstd.data_ptr<float>(), rows, H, eps, dt_code(x), cur_stream());
  TORCH_CHECK(rc == 0, "layernorm_fwd: unsupported shape/dtype");
  return {y, mean, rstd};
}

// dx, dw, db
std::vector<Tensor> layernorm_bwd(const Tensor& x, const Tensor& w, const Tensor& dy, const Tensor& mean, const Tensor& rstd) {
  CHECK_IN(x); CHECK_IN(w); CHECK_IN(dy); CHECK_IN(mean); CHECK_IN(rstd);
  const int H = x.size(-1);
  const int64_t rows = x.numel() / H;
  c10::cuda::CUDAGuard guard(x.device());
  auto dx = at::empty_like(x);
  auto dw = at::empty_like(w);
  auto db = at::empty_like(w);
  auto partial = at::empty({2, rb_layernorm_bwd_num_partials(), H}, x.options().dtype(at::kFloat));
  float* p0 = partial.data_ptr<float>();
  int rc = rb_layernorm_bwd(x.data_ptr(), w.data_ptr(), dy.data_ptr(), mean.data_ptr<float>(), rstd.data_ptr<float>(), dx.data_ptr(), p0,
                            p0 + (int64_t)rb_layernorm_bwd_num_partials() * H, dw.data_ptr(), db.data_ptr(), rows, H, dt_code(x),
                            cur_stream());
  TORCH_CHECK(rc == 0, "layernorm_bwd: unsupported shape/dtype");
  return {dx, dw, db};
}

// ---------------------------------------------------------------- rope / gated act
void rope_inplace(Tensor x, const Tensor& cos, const Tensor& sin, const Tensor& pos, int64_t n_heads, int64_t hd,
                  int64_t rot_dim, bool interleaved, bool inverse) {
  CHECK_CUDA(x); CHECK_IN(cos); CHECK_IN(sin); CHECK_IN(pos);
  TORCH_CHECK(x.dim() == 2 && x.stride(1) == 1, "x must be [T, row] with unit inner stride");
  TORCH_CHECK(cos.scalar_type() == at::kFloat && sin.scalar_type() == at::kFloat && pos.scalar_type() == at::kInt);
  TORCH_CHECK(cos.size(-1) == rot_dim / 2);
  c10::cuda::CUDAGuard guard(x.device());
  int rc = rb_rope_inplace(x.data_ptr(), cos.data_ptr<float>(), sin.data_ptr<float>(), pos.data_ptr<int>(), x.size(0),
                           (int)n_heads, (int)hd, x.stride(0), (int)rot_dim, interleaved, inverse, dt_code(x), cur_stream());
  TORCH_CHECK(rc == 0, "rope: unsupported shape/dtype");
}

Tensor gated_act_fwd(const Tensor& gu, int64_t kind) {
  CHECK_IN(gu);
  const int F = gu.size(-1) / 2;
  const int64_t rows = gu.numel() / (2 * F);
  c10::cuda::CUDAGuard guard(gu.device());
  auto sizes = gu.sizes().vec();
  sizes.back() = F;
  auto out = at::empty(sizes, gu.options());
  TORCH_CHECK(rb_gated_act_fwd(gu.data_ptr(), out.data_ptr(), rows, F, (int)kind, dt_code(gu), cur_stream()) == 0);
  return out;
}

Tensor gated_act_bwd(const Tensor& gu, const Tensor& dout, int64_t kind) {
  CHECK_IN(gu); CHECK_IN(dout);
  const int F = gu.size(-1) / 2;
  const int64_t rows = gu.numel() / (2 * F);
  c10::cuda::CUDAGuard guard(gu.device());
  auto dgu = at::empty_like(gu);
  TORCH_CHECK(rb_gated_act_bwd(gu.data_ptr(), dout.data_ptr(), dgu.data_ptr(), rows, F, (int)kind, dt_code(gu), cur_stream()) == 0);
  return dgu;
}

// ---------------------------------------------------------------- logprob
// mode 0: returns (logp, lse).  mode 1 (vocab-parallel partials): returns (max, sumexp, target_logit).
std::vector<Tensor> logprob_fwd(const Tensor& logits, const Tensor& labels, const c10::optional<Tensor>& mask,
                                double inv_temp, int64_t vocab_start, bool partials) {
  CHECK_CUDA(logits); CHECK_IN(labels);
  TORCH_CHECK(logits.dim() == 2 && logits.stride(1) == 1 && labels.scalar_type() == at::kLong);
  const int64_t rows = logits.size(0);
  const int V = logits.size(1);
  TORCH_CHECK(labels.numel() == rows);
  const uint8_t* mp = nullptr;
  int64_t ms = 0;
  if (mask.has_value()) {
    CHECK_IN(*mask);
    TORCH_CHECK(mask->scalar_type() == at::kByte && mask->size(0) == rows && mask->size(1) * 8 >= V);
    mp = mask->data_ptr<uint8_t>();
    ms = mask->stride(0);
  }
  c10::cuda::CUDAGuard guard(logits.device());
  auto fopt = logits.options().dtype(at::kFloat);
  auto a = at::empty({rows}, fopt), b = at::empty({rows}, fopt);
  Tensor c;
  int rc;
  if (partials) {
    c = at::empty({rows}, fopt);
    rc = rb_logprob_fwd(logits.data_ptr(), labels.data_ptr<int64_t>(), mp, ms, nullptr, nullptr, a.data_ptr<float>(),
                        b.data_ptr<float>(), c.data_ptr<float>(), rows, V, logits.stride(0), inv_temp, (int)vocab_start,
                        dt_code(logits), cur_stream());
  } else {
    rc = rb_logprob_fwd(logits.data_ptr(), labels.data_ptr<int64_t>(), mp, ms, a.data_ptr<float>(), b.data_ptr<float>(),
                        nullptr, nullptr, nullptr, rows, V, logits.stride(0), inv_temp, (int)vocab_start, dt_code(logits),
                        cur_stream());
  }
  TORCH_CHECK(rc == 0);
  if (partials) return {a, b, c};
  return {a, b};
}

void logprob_bwd_(Tensor logits, const Tensor& labels, const c10::optional<Tensor>& mask, const Tensor& lse,
                  const Tensor& dlogp, double inv_temp, int64_t vocab_start) {
  CHECK_CUDA(logits); CHECK_IN(labels); CHECK_IN(lse); CHECK_IN(dlogp);
  TORCH_CHECK(logits.dim() == 2 && logits.stride(1) == 1);
  TORCH_CHECK(lse.scalar_type() == at::kFloat && dlogp.scalar_type() == at::kFloat);
  const uint8_t* mp = nullptr;
  int64_t ms = 0;
  if (mask.has_value()) { mp = mask->data_ptr<uint8_t>(); ms = mask->stride(0); }
  c10::cuda::CUDAGuard guard(logits.device());
  int rc = rb_logprob_bwd(logits.data_ptr(), labels.data_ptr<int64_t>(), mp, ms, lse.data_ptr<float>(), dlogp.data_ptr<float>(),
                          logits.size(0), (int)logits.size(1), logits.stride(0), inv_temp, (int)vocab_start, dt_code(logits),
                          cur_stream());
  TORCH_CHECK(rc == 0);
}

void register_gemm_ops(torch::Library& m);     // bindings_gemm.cpp
void register_attn_ops(torch::Library& m);     // bindings_attn.cpp
void register_comm_ops(torch::Library& m);     // bindings_comm.cpp

static int64_t set_pdl(int64_t on) { const int old = rb_get_pdl(); if (on >= 0) rb_set_pdl((int)on); return old; }

TORCH_LIBRARY(realhf_b200, m) {
  m.def("set_pdl(int on) -> int", &set_pdl);
  m.def("gae_1d_misalign(Tensor rewards, Tensor values, Tensor cu_seqlens, Tensor bootstrap, float gamma, float lam) -> Tensor[]", &gae_1d_misalign);
  m.def("ppo_rewards_gae(Tensor logp, Tensor ref_logp, Tensor scores, Tensor values, Tensor cu_seqlens, Tensor no_eos, float gamma, float lam, float kl_ctl, float clip_reward) -> Tensor[]", &ppo_rewards_gae);
  m.def("gae_2d(Tensor rewards, Tensor values, Tensor dones, Tensor truncs, float gamma, float lam, int mode) -> Tensor[]", &gae_2d);
  m.def("segment_copy(Tensor src, Tensor dst, int dst_ptr, Tensor src_off, Tensor dst_off, Tensor cum, int total_bytes, float eta, bool use_ema) -> ()", &segment_copy);
  m.def("adamw_step(Tensor(a!) p, Tensor g, Tensor(b!) m, Tensor(c!) v, Tensor? master, float lr, float b1, float b2, float eps, float wd, int step, Tensor? scale, Tensor? skip, bool stochastic, int seed) -> ()", &adamw_step);
  m.def("sumsq_accum(Tensor g, Tensor(a!) out2) -> ()", &sumsq_accum);
  m.def("rmsnorm_fwd(Tensor x, Tensor? residual, Tensor w, float eps, float w_offset) -> Tensor[]", &rmsnorm_fwd);
  m.def("rmsnorm_bwd(Tensor x, Tensor w, Tensor dy, Tensor rstd, float w_offset, Tensor? dres) -> Tensor[]", &rmsnorm_bwd);
  m.def("layernorm_fwd(Tensor x, Tensor w, Tensor? b, float eps) -> Tensor[]", &layernorm_fwd);
  m.def("layernorm_bwd(Tensor x, Tensor w, Tensor dy, Tensor mean, Tensor rstd) -> Tensor[]", &layernorm_bwd);
  m.def("rope_inplace(Tensor(a!) x, Tensor cos, Tensor sin, Tensor pos, int n_heads, int hd, int rot_dim, bool interleaved, bool inverse) -> ()", &rope_inplace);
  m.def("gated_act_fwd(Tensor gu, int kind) -> Tensor", &gated_act_fwd);
  m.def("gated_act_bwd(Tensor gu, Tensor dout, int kind) -> Tensor", &gated_act_bwd);
  m.def("logprob_fwd(Tensor logits, Tensor labels, Tensor? mask, float inv_temp, int vocab_start, bool partials) -> Tensor[]", &logprob_fwd);
  m.def("logprob_bwd_(Tensor(a!) logits, Tensor labels, Tensor? mask, Tensor lse, Tensor dlogp, float inv_temp, int vocab_start) -> ()", &logprob_bwd_);
  register_gemm_ops(m);
  register_attn_ops(m);
  register_comm_ops(m);
}
