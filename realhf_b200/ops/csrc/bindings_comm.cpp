// torch bindings for the peer-memory collectives.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

#include <cstring>
#include <vector>

using at::Tensor;

extern "C" {
int rb_symm_pad_words();
int rb_symm_counter_word();
int rb_symm_barrier(const int64_t*, const int64_t*, int, int, cudaStream_t);
int rb_symm_allreduce(const int64_t*, const int64_t*, const void*, void*, int64_t, int, int, int, int, cudaStream_t);
int rb_reduce_slabs(const void*, void*, int64_t, int64_t, int, const uint32_t*, uint32_t*, uint32_t, int, cudaStream_t);
int rb_symm_calls_word(int);
int rb_gemm_2cta_gated(const void* A, const void* B, void* C, const void* bias, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc,
                       int a_mn, int b_mn, int in_dt, int out_dt, int accumulate, int bn, int num_sms, const uint32_t* ready_flags,
                       uint32_t ready_epoch, int rows_per_flag, int m_rot_rows, uint32_t* done_counters, cudaStream_t s);
int rb_spin_wait(const uint32_t* flag, uint32_t target, cudaStream_t s);
int rb_gemm_fused_tp(int mode, const void* A, const int64_t* peer_a, const void* B, void* C, int M, int N, int K, int64_t lda, int64_t ldb,
                     int64_t ldc, int b_mn, const int64_t* peer_base, const int64_t* peer_counter, int rows_per_rank, int my_rank,
                     int world, int num_sms, cudaStream_t s);
}

static int dtc(at::ScalarType t) {
  switch (t) {
    case at::kFloat: return 0;
    case at::kBFloat16: return 1;
    case at::kHalf: return 2;
    default: TORCH_CHECK(false, "unsupported dtype ", t);
  }
}

int64_t symm_pad_words() { return rb_symm_pad_words(); }
int64_t symm_counter_word() { return rb_symm_counter_word(); }

void symm_barrier(const Tensor& anchor, std::vector<int64_t> data_ptrs, std::vector<int64_t> pad_ptrs, int64_t rank) {
  c10::cuda::CUDAGuard g(anchor.device());
  TORCH_CHECK(rb_symm_barrier(data_ptrs.data(), pad_ptrs.data(), (int)rank, (int)data_ptrs.size(),
                              at::cuda::getCurrentCUDAStream().stream()) == 0);
}

// out = sum over ranks of in.  algo: 1 one-shot, 2 two-shot.
void symm_allreduce(const Tensor& in, Tensor out, std::vector<int64_t> data_ptrs, std::vector<int64_t> pad_ptrs, int64_t rank, int64_t algo) {
  TORCH_CHECK(in.is_cuda() && in.is_contiguous() && out.is_contiguous() && in.nbytes() == out.nbytes());
  TORCH_CHECK(in.nbytes() % 16 == 0, "all-reduce payload must be a multiple of 16 bytes");
  c10::cuda::CUDAGuard g(in.device());
  int rc = rb_symm_allreduce(data_ptrs.data(), pad_ptrs.data(), in.data_ptr(), out.data_ptr(), in.nbytes(), (int)rank,
                             (int)data_ptrs.size(), dtc(in.scalar_type()), (int)algo, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "symm_allreduce failed: ", rc);
}

// y = a @ w^T (or a @ w if b_mn) on the CTA-pair kernel, reading row-block i of `a` only after flags[i] >= epoch: the
// all-gather -> GEMM path fills `a` chunk by chunk on a copy stream while the tensor cores already work on the local rows.
Tensor gemm_gated(const Tensor& a, const Tensor& w, bool b_mn, const c10::optional<Tensor>& flags, int64_t epoch, int64_t rows_per_flag,
                  int64_t first_row, int64_t num_sms, const c10::optional<Tensor>& done) {
  TORCH_CHECK(a.is_cuda() && a.dim() == 2 && w.dim() == 2 && a.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16);
  TORCH_CHECK(a.stride(1) == 1 && w.stride(1) == 1);
  const int64_t M = a.size(0), K = a.size(1), N = b_mn ? w.size(1) : w.size(0);
  for (const auto* t : {&flags, &done}) {
    if (t->has_value())
      TORCH_CHECK((*t)->scalar_type() == at::kInt && (*t)->is_contiguous() && (*t)->numel() * rows_per_flag >= M, "gemm_gated: bad flag tensor");
  }
  c10::cuda::CUDAGuard g(a.device());
  auto y = at::empty({M, N}, a.options());
  int rc = rb_gemm_2cta_gated(a.data_ptr(), w.data_ptr(), y.data_ptr(), nullptr, (int)M, (int)N, (int)K, a.stride(0), w.stride(0), N, 0,
                              b_mn, 1, 1, 0, 0, (int)num_sms, flags.has_value() ? reinterpret_cast<const uint32_t*>(flags->data_ptr()) : nullptr,
                              (uint32_t)epoch, (int)rows_per_flag, (int)first_row,
                              done.has_value() ? reinterpret_cast<uint32_t*>(done->data_ptr()) : nullptr,
                              at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "gemm_gated failed: ", rc);
  return y;
}

// Blocks the current stream until flags[index] >= target (device-side spin of one thread; the flag is written by another
// stream, a copy engine or a peer GPU).
void spin_wait(const Tensor& flags, int64_t index, int64_t target) {
  TORCH_CHECK(flags.is_cuda() && flags.scalar_type() == at::kInt && index >= 0 && index < flags.numel());
  c10::cuda::CUDAGuard g(flags.device());
  TORCH_CHECK(rb_spin_wait(reinterpret_cast<const uint32_t*>(flags.data_ptr()) + index, (uint32_t)target,
                           at::cuda::getCurrentCUDAStream().stream()) == 0);
}

int64_t symm_calls_word(int64_t parity) { return rb_symm_calls_word((int)parity); }

// GEMM -> reduce-scatter in one op: y[rows_per_rank, N] = sum_ranks (x_rank @ w_rank^T)[my rows].
//   x [T, Kl], w: b_mn ? [Kl, N] : [N, Kl].  The GEMM epilogue stores every output row into the owning rank's inbox
//   (peer memory) while later tiles are still on the tensor cores; the tail kernel sums the `world` slabs of my rows.
Tensor gemm_rs(const Tensor& x, const Tensor& w, bool b_mn, std::vector<int64_t> peer_inbox, std::vector<int64_t> peer_counter,
               int64_t my_calls_ptr, int64_t rank, int64_t num_sms) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && w.dim() == 2 && x.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16);
  TORCH_CHECK(x.stride(1) == 1 && w.stride(1) == 1);
  const int world = (int)peer_inbox.size();
  const int64_t T = x.size(0), K = x.size(1), N = b_mn ? w.size(1) : w.size(0);
  TORCH_CHECK(T % world == 0, "gemm_rs: token count must divide by the TP size");
  const int rows = (int)(T / world);
  c10::cuda::CUDAGuard g(x.device());
  auto stream = at::cuda::getCurrentCUDAStream().stream();
  int bn = rb_gemm_fused_tp(1, x.data_ptr(), nullptr, w.data_ptr(), nullptr, (int)T, (int)N, (int)K, x.stride(0), w.stride(0), N, b_mn,
                            peer_inbox.data(), peer_counter.data(), rows, (int)rank, world, (int)num_sms, stream);
  TORCH_CHECK(bn == 0, "gemm_rs: fused GEMM launch failed (", bn, ")");
  auto y = at::empty({rows, N}, x.options());
  // arrivals per call at this rank: every source delivers each of my rows once per n-tile
  const int tm = (int)((T + 127) / 128);
  int bnv = (((int64_t)tm * ((N + 255) / 256) >= (num_sms > 0 ? num_sms : 148)) || N % 256 == 0) ? 256 : 128;
  if (N < 256) bnv = 128;
  const uint32_t per_call = (uint32_t)((int64_t)rows * ((N + bnv - 1) / bnv) * world);
  int rc = rb_reduce_slabs(reinterpret_cast<const void*>(peer_inbox[rank]), y.data_ptr(), y.nbytes(), (int64_t)rows * N * 2, world,
                           reinterpret_cast<const uint32_t*>(peer_counter[rank]), reinterpret_cast<uint32_t*>(my_calls_ptr), per_call, 1,
                           stream);
  TORCH_CHECK(rc == 0, "gemm_rs: reduce launch failed");
  return y;
}

// all-gather -> GEMM: y[T, N] = concat_r(x_r) @ w^T where x_r [rows, K] lives at peer_a[r] (symmetric staging buffers).
Tensor ag_gemm(std::vector<int64_t> peer_a, int64_t rows, int64_t K, const Tensor& w, bool b_mn, int64_t rank, int64_t num_sms) {
  TORCH_CHECK(w.is_cuda() && w.dim() == 2 && w.scalar_type() == at::kBFloat16 && w.stride(1) == 1);
  const int world = (int)peer_a.size();
  const int64_t N = b_mn ? w.size(1) : w.size(0);
  c10::cuda::CUDAGuard g(w.device());
  auto y = at::empty({rows * world, N}, w.options());
  int rc = rb_gemm_fused_tp(2, nullptr, peer_a.data(), w.data_ptr(), y.data_ptr(), (int)(rows * world), (int)N, (int)K, K, w.stride(0), N, b_mn,
                            nullptr, nullptr, (int)rows, (int)rank, world, (int)num_sms, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "ag_gemm: fused GEMM launch failed (", rc, ")");
  return y;
}

// ---- symmetric allocations: raw cudaMalloc so that the IPC handle maps exactly this buffer, opened by each peer in ITS
// OWN device context (cudaIpcOpenMemHandle with lazy peer access) - torch's tensor-sharing path opens the handle under
// the owner's device index, which does not make the mapping usable by kernels running on the consumer's GPU.
std::vector<Tensor> symm_alloc(int64_t nbytes, int64_t device) {
  c10::cuda::CUDAGuard g((c10::DeviceIndex)device);
  void* p = nullptr;
  TORCH_CHECK(cudaMalloc(&p, nbytes) == cudaSuccess, "cudaMalloc failed for symmetric buffer of ", nbytes, " bytes");
  TORCH_CHECK(cudaMemset(p, 0, nbytes) == cudaSuccess);
  TORCH_CHECK(cudaDeviceSynchronize() == cudaSuccess);
  cudaIpcMemHandle_t h;
  TORCH_CHECK(cudaIpcGetMemHandle(&h, p) == cudaSuccess, "cudaIpcGetMemHandle failed");
  auto handle = at::empty({(int64_t)sizeof(h)}, at::TensorOptions().dtype(at::kByte));
  memcpy(handle.data_ptr(), &h, sizeof(h));
  auto t = at::from_blob(p, {nbytes}, [](void* q) { cudaFree(q); }, at::TensorOptions().dtype(at::kByte).device(at::kCUDA, (c10::DeviceIndex)device));
  return {t, handle};
}

Tensor symm_open(const Tensor& handle, int64_t nbytes, int64_t device) {
  TORCH_CHECK(handle.is_cpu() && handle.numel() == (int64_t)sizeof(cudaIpcMemHandle_t));
  c10::cuda::CUDAGuard g((c10::DeviceIndex)device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle.data_ptr(), sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  TORCH_CHECK(e == cudaSuccess, "cudaIpcOpenMemHandle failed: ", cudaGetErrorString(e));
  return at::from_blob(p, {nbytes}, [](void* q) { cudaIpcCloseMemHandle(q); },
                       at::TensorOptions().dtype(at::kByte).device(at::kCUDA, (c10::DeviceIndex)device));
}

void register_comm_ops(torch::Library& m) {
  m.def("symm_alloc(int nbytes, int device) -> Tensor[]", &symm_alloc);
  m.def("symm_open(Tensor handle, int nbytes, int device) -> Tensor", &symm_open);
  m.def("symm_pad_words() -> int", &symm_pad_words);
  m.def("symm_counter_word() -> int", &symm_counter_word);
  m.def("symm_barrier(Tensor anchor, int[] data_ptrs, int[] pad_ptrs, int rank) -> ()", &symm_barrier);
  m.def("symm_allreduce(Tensor inp, Tensor(a!) out, int[] data_ptrs, int[] pad_ptrs, int rank, int algo) -> ()", &symm_allreduce);
  m.def("symm_calls_word(int parity) -> int", &symm_calls_word);
  m.def("gemm_rs(Tensor x, Tensor w, bool b_mn, int[] peer_inbox, int[] peer_counter, int my_calls_ptr, int rank, int num_sms) -> Tensor", &gemm_rs);
  m.def("gemm_gated(Tensor a, Tensor w, bool b_mn, Tensor? flags, int epoch, int rows_per_flag, int first_row, int num_sms, Tensor? done) -> Tensor", &gemm_gated);
  m.def("spin_wait(Tensor flags, int index, int target) -> ()", &spin_wait);
  m.def("ag_gemm(int[] peer_a, int rows, int K, Tensor w, bool b_mn, int rank, int num_sms) -> Tensor", &ag_gemm);
}
