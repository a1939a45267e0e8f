// torch bindings for the tcgen05 GEMM family.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

using at::Tensor;

extern "C" int rb_gemm_tcgen05(const void* A, const void* B, void* C, const void* bias, int M, int N, int K, int64_t lda,
                               int64_t ldb, int64_t ldc, int a_mn, int b_mn, int in_dt, int out_dt, int accumulate, int bn,
                               int num_sms, int mc_req, cudaStream_t s);

extern "C" int rb_gemm_streamk(const void* A, const void* B, void* C, const void* bias, int M, int N, int K, int64_t lda,
                               int64_t ldb, int64_t ldc, int in_dt, int out_dt, int bn, int split, int num_sms, void* ws, void* flags, void* dbg,
                               cudaStream_t s);

extern "C" int rb_gemm_streamk_fp8(const void* A, const void* B, void* C, const void* bias, const float* scale_a, const float* scale_b,
                                   int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc, int out_dt, int bn, int split, int num_sms,
                                   void* ws, void* flags, cudaStream_t s);
extern "C" int rb_quant_rows_e4m3(const void* x, void* q, float* scale, int M, int K, int64_t ld_x, int64_t ld_q, int dt, cudaStream_t s);
extern "C" int rb_gated_act_quant_e4m3(const void* gu, void* q, float* scale, int M, int F, int64_t ld_x, int64_t ld_q, int act, int dt,
                                       cudaStream_t s);
extern "C" int rb_add_rmsnorm_quant_e4m3(const void* x, const void* res_in, const void* w, void* res_out, void* q, float* scale, int64_t rows,
                                         int H, int64_t ld_q, float eps, float w_offset, int dt, cudaStream_t s);

extern "C" int rb_gemm_grouped(const void* A, const void* B, void* C, const int* group_offsets, int G, int M, int N, int K, int64_t lda,
                               int64_t ldb, int64_t ldc, int b_mn, int num_sms, cudaStream_t s);

static int dtc(at::ScalarType t) {
  switch (t) {
    case at::kFloat: return 0;
    case at::kBFloat16: return 1;
    case at::kHalf: return 2;
    default: TORCH_CHECK(false, "gemm: unsupported dtype ", t);
  }
}

// a: a_mn ? [K, M] : [M, K];  b: b_mn ? [K, N] : [N, K].  Inner stride must be 1, row pitch a multiple of 8.
// out: optional destination [M, N] (any supported dtype); with accumulate=True computes out += a x b.
Tensor gemm(const Tensor& a, const Tensor& b, const c10::optional<Tensor>& out, const c10::optional<Tensor>& bias, bool a_mn,
            bool b_mn, bool accumulate, c10::optional<at::ScalarType> out_dtype, int64_t bn, int64_t num_sms, int64_t mc) {
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && a.dim() == 2 && b.dim() == 2, "gemm: 2-D CUDA tensors expected");
  TORCH_CHECK(a.stride(1) == 1 && b.stride(1) == 1, "gemm: inner stride must be 1");
  TORCH_CHECK(a.scalar_type() == b.scalar_type(), "gemm: operand dtypes differ");
  const int64_t M = a_mn ? a.size(1) : a.size(0), K = a_mn ? a.size(0) : a.size(1);
  const int64_t N = b_mn ? b.size(1) : b.size(0), Kb = b_mn ? b.size(0) : b.size(1);
  TORCH_CHECK(K == Kb, "gemm: reduction dims differ: ", K, " vs ", Kb);
  c10::cuda::CUDAGuard guard(a.device());
  Tensor c;
  if (out.has_value()) {
    c = *out;
    TORCH_CHECK(c.dim() == 2 && c.size(0) == M && c.size(1) == N && c.stride(1) == 1, "gemm: bad out shape");
  } else {
    TORCH_CHECK(!accumulate, "gemm: accumulate needs `out`");
    c = at::empty({M, N}, a.options().dtype(out_dtype.value_or(a.scalar_type())));
  }
  const void* bp = nullptr;
  if (bias.has_value()) {
    TORCH_CHECK(bias->scalar_type() == c.scalar_type() && bias->numel() == N && bias->is_contiguous());
    bp = bias->data_ptr();
  }
  if (M == 0 || N == 0) return c;
  int rc = rb_gemm_tcgen05(a.data_ptr(), b.data_ptr(), c.data_ptr(), bp, (int)M, (int)N, (int)K, a.stride(0), b.stride(0),
                           c.stride(0), a_mn, b_mn, dtc(a.scalar_type()), dtc(c.scalar_type()), accumulate, (int)bn, (int)num_sms, (int)mc,
                           at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "rb_gemm_tcgen05 failed with code ", rc);
  return c;
}

// Decode-shaped GEMM (M <= 128): y = a @ b^T (+ bias), stream-K over all SMs.  ws: fp32 [>= 2*num_sms*128*256], flags: int32 [8192] zeros.
Tensor gemm_streamk(const Tensor& a, const Tensor& b, const c10::optional<Tensor>& out, const c10::optional<Tensor>& bias, const Tensor& ws, const Tensor& flags,
                    c10::optional<at::ScalarType> out_dtype, int64_t bn, int64_t split, int64_t num_sms, const c10::optional<Tensor>& dbg) {
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && a.dim() == 2 && b.dim() == 2 && a.stride(1) == 1 && b.stride(1) == 1);
  TORCH_CHECK(a.scalar_type() == b.scalar_type() && a.size(1) == b.size(1), "gemm_streamk: operand mismatch");
  const int64_t M = a.size(0), K = a.size(1), N = b.size(0);
  TORCH_CHECK(M <= 128, "gemm_streamk: M <= 128");
  TORCH_CHECK(ws.scalar_type() == at::kFloat && ws.numel() >= 2 * num_sms * 128 * 256 && flags.numel() >= 8192 && flags.element_size() == 4);
  c10::cuda::CUDAGuard guard(a.device());
  Tensor c;
  if (out.has_value()) {
    c = *out;
    TORCH_CHECK(c.dim() == 2 && c.size(0) == M && c.size(1) == N && c.stride(1) == 1, "gemm_streamk: bad out shape");
  } else {
    c = at::empty({M, N}, a.options().dtype(out_dtype.value_or(a.scalar_type())));
  }
  const void* bp = nullptr;
  if (bias.has_value()) {
    TORCH_CHECK(bias->scalar_type() == c.scalar_type() && bias->numel() == N && bias->is_contiguous());
    bp = bias->data_ptr();
  }
  if (M == 0 || N == 0) return c;
  int rc = rb_gemm_streamk(a.data_ptr(), b.data_ptr(), c.data_ptr(), bp, (int)M, (int)N, (int)K, a.stride(0), b.stride(0), c.stride(0),
                           dtc(a.scalar_type()), dtc(c.scalar_type()), (int)bn, (int)split, (int)num_sms, ws.data_ptr(), flags.data_ptr(), dbg.has_value() ? dbg->data_ptr() : nullptr,
                           at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "rb_gemm_streamk failed with code ", rc);
  return c;
}

// Row-wise e4m3 quantisation: x [M, K] (bf16 / fp16 / fp32, inner stride 1) -> (q uint8 [M, K], scale fp32 [M]), x ~ q * scale.
// `q_out` / `scale_out`: optional preallocated destinations (the decode loop reuses them inside a CUDA graph).
std::vector<Tensor> quant_rows_e4m3(const Tensor& x, const c10::optional<Tensor>& q_out, const c10::optional<Tensor>& scale_out) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && x.stride(1) == 1, "quant_rows_e4m3: x must be a CUDA matrix with unit inner stride");
  const int64_t M = x.size(0), K = x.size(1);
  c10::cuda::CUDAGuard guard(x.device());
  Tensor q = q_out.has_value() ? *q_out : at::empty({M, K}, x.options().dtype(at::kByte));
  Tensor sc = scale_out.has_value() ? *scale_out : at::empty({M}, x.options().dtype(at::kFloat));
  TORCH_CHECK(q.scalar_type() == at::kByte && q.dim() == 2 && q.size(0) == M && q.size(1) == K && q.stride(1) == 1 &&
              sc.scalar_type() == at::kFloat && sc.numel() >= M && sc.is_contiguous(), "quant_rows_e4m3: bad destinations");
  if (M == 0) return {q, sc};
  int rc = rb_quant_rows_e4m3(x.data_ptr(), q.data_ptr(), sc.data_ptr<float>(), (int)M, (int)K, x.stride(0), q.stride(0), dtc(x.scalar_type()),
                              at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "rb_quant_rows_e4m3 failed with code ", rc, " (K % 8 == 0, K <= 16384, 16-byte aligned rows)");
  return {q, sc};
}

// gu [M, 2F] = [gate | up] -> (e4m3 bytes [M, F] of act(gate) * up, scale [M]); act_kind: 0 silu, 1 gelu(tanh)
std::vector<Tensor> gated_act_quant_e4m3(const Tensor& gu, int64_t act_kind) {
  TORCH_CHECK(gu.is_cuda() && gu.dim() == 2 && gu.stride(1) == 1 && gu.size(1) % 2 == 0, "gated_act_quant_e4m3: gu is [M, 2F]");
  const int64_t M = gu.size(0), F = gu.size(1) / 2;
  c10::cuda::CUDAGuard guard(gu.device());
  Tensor q = at::empty({M, F}, gu.options().dtype(at::kByte));
  Tensor sc = at::empty({M}, gu.options().dtype(at::kFloat));
  if (M == 0) return {q, sc};
  int rc = rb_gated_act_quant_e4m3(gu.data_ptr(), q.data_ptr(), sc.data_ptr<float>(), (int)M, (int)F, gu.stride(0), q.stride(0), (int)act_kind,
                                   dtc(gu.scalar_type()), at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "rb_gated_act_quant_e4m3 failed with code ", rc);
  return {q, sc};
}

// (x [+ residual]) -> [q, scale] of rmsnorm(x + residual) * (w + w_offset), and the new residual stream when `residual` is given
std::vector<Tensor> add_rmsnorm_quant_e4m3(const Tensor& x, const c10::optional<Tensor>& residual, const Tensor& w, double eps, double w_offset) {
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && w.is_contiguous() && w.scalar_type() == x.scalar_type() && w.numel() == x.size(-1));
  const int64_t H = x.size(-1), rows = x.numel() / H;
  c10::cuda::CUDAGuard guard(x.device());
  Tensor q = at::empty({rows, H}, x.options().dtype(at::kByte));
  Tensor sc = at::empty({rows}, x.options().dtype(at::kFloat));
  const void* rin = nullptr;
  Tensor rout;
  if (residual.has_value()) {
    TORCH_CHECK(residual->is_contiguous() && residual->sizes() == x.sizes() && residual->scalar_type() == x.scalar_type());
    rin = residual->data_ptr();
    rout = at::empty_like(x);
  }
  if (rows > 0) {
    int rc = rb_add_rmsnorm_quant_e4m3(x.data_ptr(), rin, w.data_ptr(), rin ? rout.data_ptr() : nullptr, q.data_ptr(), sc.data_ptr<float>(), rows,
                                       (int)H, q.stride(0), (float)eps, (float)w_offset, dtc(x.scalar_type()), at::cuda::getCurrentCUDAStream().stream());
    TORCH_CHECK(rc == 0, "rb_add_rmsnorm_quant_e4m3 failed with code ", rc);
  }
  if (residual.has_value()) return {q, sc, rout};
  return {q, sc};
}

// W8A8 decode GEMM (M <= 128): y[m, n] = (sum_k a8[m, k] * b8[n, k]) * sa[m] * sb[n] (+ bias); a8 / b8 hold e4m3 bytes.
Tensor gemm_streamk_fp8(const Tensor& a, const Tensor& b, const Tensor& sa, const Tensor& sb, const c10::optional<Tensor>& out,
                        const c10::optional<Tensor>& bias, const Tensor& ws, const Tensor& flags, at::ScalarType out_dtype, int64_t bn,
                        int64_t split, int64_t num_sms) {
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && a.dim() == 2 && b.dim() == 2 && a.stride(1) == 1 && b.stride(1) == 1);
  TORCH_CHECK(a.scalar_type() == at::kByte && b.scalar_type() == at::kByte && a.size(1) == b.size(1), "gemm_streamk_fp8: operands are e4m3 bytes");
  const int64_t M = a.size(0), K = a.size(1), N = b.size(0);
  TORCH_CHECK(M <= 128, "gemm_streamk_fp8: M <= 128");
  TORCH_CHECK(sa.scalar_type() == at::kFloat && sb.scalar_type() == at::kFloat && sa.numel() >= M && sb.numel() == N && sa.is_contiguous() &&
              sb.is_contiguous(), "gemm_streamk_fp8: scales are fp32 [M] / [N]");
  TORCH_CHECK(ws.scalar_type() == at::kFloat && ws.numel() >= 2 * num_sms * 128 * 256 && flags.numel() >= 8192 && flags.element_size() == 4);
  c10::cuda::CUDAGuard guard(a.device());
  Tensor c;
  if (out.has_value()) {
    c = *out;
    TORCH_CHECK(c.dim() == 2 && c.size(0) == M && c.size(1) == N && c.stride(1) == 1, "gemm_streamk_fp8: bad out shape");
  } else {
    c = at::empty({M, N}, a.options().dtype(out_dtype));
  }
  const void* bp = nullptr;
  if (bias.has_value()) {
    TORCH_CHECK(bias->scalar_type() == c.scalar_type() && bias->numel() == N && bias->is_contiguous());
    bp = bias->data_ptr();
  }
  if (M == 0 || N == 0) return c;
  int rc = rb_gemm_streamk_fp8(a.data_ptr(), b.data_ptr(), c.data_ptr(), bp, sa.data_ptr<float>(), sb.data_ptr<float>(), (int)M, (int)N, (int)K,
                               a.stride(0), b.stride(0), c.stride(0), dtc(c.scalar_type()), (int)bn, (int)split, (int)num_sms, ws.data_ptr(),
                               flags.data_ptr(), at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "rb_gemm_streamk_fp8 failed with code ", rc);
  return c;
}

// Grouped GEMM over row groups (MoE experts): a [M, K] sorted by group, offsets int32 [G+1] on the device,
// w [G, N, K] (b_mn = false: y = a_g @ w_g^T) or [G, K, N] (b_mn = true: y = a_g @ w_g).  One launch, no host sync.
Tensor gemm_grouped(const Tensor& a, const Tensor& w, const Tensor& offsets, bool b_mn, int64_t num_sms) {
  TORCH_CHECK(a.is_cuda() && a.dim() == 2 && w.dim() == 3 && a.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16);
  TORCH_CHECK(a.stride(1) == 1 && w.is_contiguous() && offsets.scalar_type() == at::kInt && offsets.is_contiguous() && offsets.is_cuda());
  const int64_t G = w.size(0), M = a.size(0), K = a.size(1);
  const int64_t N = b_mn ? w.size(2) : w.size(1);
  TORCH_CHECK((b_mn ? w.size(1) : w.size(2)) == K && offsets.numel() == G + 1, "gemm_grouped: shape mismatch");
  c10::cuda::CUDAGuard guard(a.device());
  Tensor c = at::empty({M, N}, a.options());
  if (M == 0) return c;
  int rc = rb_gemm_grouped(a.data_ptr(), w.data_ptr(), c.data_ptr(), offsets.data_ptr<int>(), (int)G, (int)M, (int)N, (int)K, a.stride(0),
                           w.stride(1), c.stride(0), b_mn, (int)num_sms, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "rb_gemm_grouped failed with code ", rc);
  return c;
}

extern "C" int rb_gemm_grouped_wgrad(const void* dy, const void* x, void* out, const int* offsets, int G, int Tp, int M, int N,
                                     int64_t ld_dy, int64_t ld_x, int out_dt, int accumulate, int num_sms, cudaStream_t s);

// Grouped wgrad: out[g] (+)= dy[rows of g]^T @ x[rows of g]; rows sorted by group with every group's block padded (zero rows) to a
// multiple of 64, `offsets` int32 [G+1] on the device holding the PADDED block starts.  out [G, M, N] fp32 or bf16.
void gemm_grouped_wgrad(const Tensor& dy, const Tensor& x, Tensor out, const Tensor& offsets, bool accumulate, int64_t num_sms) {
  TORCH_CHECK(dy.is_cuda() && dy.dim() == 2 && x.dim() == 2 && out.dim() == 3 && dy.scalar_type() == at::kBFloat16 &&
              x.scalar_type() == at::kBFloat16 && dy.stride(1) == 1 && x.stride(1) == 1 && out.is_contiguous());
  TORCH_CHECK(out.scalar_type() == at::kFloat || out.scalar_type() == at::kBFloat16);
  TORCH_CHECK(offsets.scalar_type() == at::kInt && offsets.is_contiguous() && offsets.is_cuda());
  const int64_t G = out.size(0), M = out.size(1), N = out.size(2), Tp = dy.size(0);
  TORCH_CHECK(dy.size(1) == M && x.size(1) == N && x.size(0) == Tp && offsets.numel() == G + 1, "gemm_grouped_wgrad: shape mismatch");
  c10::cuda::CUDAGuard guard(dy.device());
  int rc = rb_gemm_grouped_wgrad(dy.data_ptr(), x.data_ptr(), out.data_ptr(), offsets.data_ptr<int>(), (int)G, (int)Tp, (int)M, (int)N,
                                 dy.stride(0), x.stride(0), out.scalar_type() == at::kFloat ? 0 : 1, accumulate ? 1 : 0, (int)num_sms,
                                 at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "rb_gemm_grouped_wgrad failed with code ", rc);
}

extern "C" int rb_gemm_2cta_glu(const void* A, const void* B, void* act, void* raw, int M, int F, int K, int64_t lda, int64_t ldb, int64_t ld_act,
                                int64_t ld_raw, int in_dt, int act_kind, int num_sms, cudaStream_t s);

// act = glu(x @ [gate; up]^T): the gated activation runs in the GEMM epilogue (CTA-pair kernel); with `want_raw` the raw
// gate | up projections [M, 2F] are written too (the backward pass needs them).  Returns [act] or [act, raw].
std::vector<Tensor> gemm_glu(const Tensor& x, const Tensor& w, int64_t act_kind, bool want_raw, int64_t num_sms) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && w.dim() == 2 && x.stride(1) == 1 && w.stride(1) == 1 && w.size(0) % 2 == 0 && w.size(1) == x.size(1));
  TORCH_CHECK(x.scalar_type() == w.scalar_type() && (x.scalar_type() == at::kBFloat16 || x.scalar_type() == at::kHalf));
  const int64_t M = x.size(0), K = x.size(1), F = w.size(0) / 2;
  c10::cuda::CUDAGuard g(x.device());
  auto act = at::empty({M, F}, x.options());
  Tensor raw;
  if (want_raw) raw = at::empty({M, 2 * F}, x.options());
  int rc = rb_gemm_2cta_glu(x.data_ptr(), w.data_ptr(), act.data_ptr(), want_raw ? raw.data_ptr() : nullptr, (int)M, (int)F, (int)K, x.stride(0),
                            w.stride(0), F, 2 * F, x.scalar_type() == at::kBFloat16 ? 1 : 2, (int)act_kind, (int)num_sms,
                            at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "gemm_glu failed: ", rc);
  if (want_raw) return {act, raw};
  return {act};
}

void register_gemm_ops(torch::Library& m) {
  m.def("gemm_glu(Tensor x, Tensor w, int act_kind, bool want_raw, int num_sms) -> Tensor[]", &gemm_glu);
  m.def("gemm_grouped_wgrad(Tensor dy, Tensor x, Tensor(a!) out, Tensor offsets, bool accumulate, int num_sms) -> ()", &gemm_grouped_wgrad);
  m.def("gemm_grouped(Tensor a, Tensor w, Tensor offsets, bool b_mn, int num_sms) -> Tensor", &gemm_grouped);
  m.def("gated_act_quant_e4m3(Tensor gu, int act_kind) -> Tensor[]", &gated_act_quant_e4m3);
  m.def("add_rmsnorm_quant_e4m3(Tensor x, Tensor? residual, Tensor w, float eps, float w_offset) -> Tensor[]", &add_rmsnorm_quant_e4m3);
  m.def("quant_rows_e4m3(Tensor x, Tensor? q_out, Tensor? scale_out) -> Tensor[]", &quant_rows_e4m3);
  m.def("gemm_streamk_fp8(Tensor a, Tensor b, Tensor sa, Tensor sb, Tensor? out, Tensor? bias, Tensor ws, Tensor flags, ScalarType out_dtype, int bn, int split, int num_sms) -> Tensor", &gemm_streamk_fp8);
  m.def("gemm_streamk(Tensor a, Tensor b, Tensor? out, Tensor? bias, Tensor ws, Tensor flags, ScalarType? out_dtype, int bn, int split, int num_sms, Tensor? dbg) -> Tensor", &gemm_streamk);
  m.def("gemm(Tensor a, Tensor b, Tensor? out, Tensor? bias, bool a_mn, bool b_mn, bool accumulate, ScalarType? out_dtype, int bn, int num_sms, int mc) -> Tensor", &gemm);
}
