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
// vmm.cpp
int rb_vmm_multicast_supported(int dev);
int64_t rb_vmm_granularity(int dev, int ndev);
int rb_vmm_alloc(int64_t size, int64_t align, int dev, uint64_t* ptr, uint64_t* handle, int* fd);
int rb_vmm_import(int fd, int64_t size, int64_t align, int dev, uint64_t* ptr, uint64_t* handle);
int rb_vmm_free(uint64_t ptr, int64_t size, uint64_t handle);
int rb_mc_create(int64_t size, int ndev, uint64_t* mc, int* fd);
int rb_mc_import(int fd, uint64_t* mc);
int rb_mc_add_device(uint64_t mc, int dev);
int rb_mc_bind_and_map(uint64_t mc, uint64_t mem_handle, int64_t size, int64_t align, int dev, uint64_t* mc_ptr);
int rb_mc_unbind(uint64_t mc, int dev, int64_t size);
int rb_close_fd(int fd);
// ep.cu
int rb_ep_plan(const int64_t* data_ptrs, const int64_t* pad_ptrs, int64_t post_off, const int* counts, int E, int e_local, int rank, int world,
               int parity, int* send_tab, int* ret_tab, int* per_expert, int* overflow, int cap_rows, cudaStream_t s);
int rb_ep_move_rows(const int64_t* data_ptrs, const int64_t* pad_ptrs, const int* table, int n_seg, const void* src, int64_t src_pitch,
                    int64_t dst_off, int64_t dst_pitch, int row_bytes, int dst_cap_rows, int rank, int world, int blocks, cudaStream_t s);
// nvls.cu
int rb_nvls_allreduce(const int64_t* data_ptrs, const int64_t* pad_ptrs, uint64_t mc, const void* in, void* out, int64_t nbytes,
                      int64_t off_in, int64_t off_out, int rank, int world, int dt, int mode, int max_blocks, cudaStream_t s);
int rb_nvls_reduce_scatter(const int64_t* data_ptrs, const int64_t* pad_ptrs, uint64_t mc, int64_t off, int64_t nbytes, int dt, float scale,
                           float* stats, int rank, int world, cudaStream_t s);
int rb_nvls_adam_allgather(const int64_t* data_ptrs, const int64_t* pad_ptrs, uint64_t mc, int64_t p_off, int p_dt, const void* g, int g_dt,
                           void* m, void* v, int s_dt, float* master, int64_t n, float lr, float b1, float b2, float eps, float wd,
                           int step, const float* scale_ptr, const int* skip_ptr, int stochastic, uint32_t seed, int rank, int world,
                           cudaStream_t s);
int rb_nvls_ar_add_rmsnorm(const int64_t* data_ptrs, const int64_t* pad_ptrs, uint64_t mc, int64_t off, const void* res_in, const void* w,
                           void* y, void* res_out, int rows, int H, float eps, float w_offset, int rank, int world, int dt, cudaStream_t s);
int rb_nvls_allgather(const int64_t* data_ptrs, const int64_t* pad_ptrs, uint64_t mc, int64_t off, int64_t nbytes, int lead_barrier, int rank,
                      int world, cudaStream_t s);
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

// ---- VMM symmetric memory + NVSwitch multicast (vmm.cpp).  Handles and pointers cross the binding as integers; the python
// side (parallel/symm_mem.py) owns the lifetime and ships the file descriptors between processes over unix sockets.
int64_t vmm_multicast_supported(int64_t device) { return rb_vmm_multicast_supported((int)device); }
int64_t vmm_granularity(int64_t device, int64_t ndev) { return rb_vmm_granularity((int)device, (int)ndev); }

// -> [ptr, handle, fd]
std::vector<int64_t> vmm_alloc(int64_t nbytes, int64_t align, int64_t device) {
  c10::cuda::CUDAGuard g((c10::DeviceIndex)device);
  TORCH_CHECK(cudaFree(nullptr) == cudaSuccess);  // make sure the primary context exists
  uint64_t ptr = 0, h = 0;
  int fd = -1;
  int rc = rb_vmm_alloc(nbytes, align, (int)device, &ptr, &h, &fd);
  TORCH_CHECK(rc == 0, "vmm_alloc(", nbytes, ") failed: ", rc);
  TORCH_CHECK(cudaMemset(reinterpret_cast<void*>(ptr), 0, nbytes) == cudaSuccess);
  TORCH_CHECK(cudaDeviceSynchronize() == cudaSuccess);
  return {(int64_t)ptr, (int64_t)h, (int64_t)fd};
}

// -> [ptr, handle]
std::vector<int64_t> vmm_import(int64_t fd, int64_t nbytes, int64_t align, int64_t device) {
  c10::cuda::CUDAGuard g((c10::DeviceIndex)device);
  TORCH_CHECK(cudaFree(nullptr) == cudaSuccess);
  uint64_t ptr = 0, h = 0;
  int rc = rb_vmm_import((int)fd, nbytes, align, (int)device, &ptr, &h);
  TORCH_CHECK(rc == 0, "vmm_import failed: ", rc);
  return {(int64_t)ptr, (int64_t)h};
}

void vmm_free(int64_t ptr, int64_t nbytes, int64_t handle) { rb_vmm_free((uint64_t)ptr, nbytes, (uint64_t)handle); }

// -> [mc_handle, fd]
std::vector<int64_t> mc_create(int64_t nbytes, int64_t ndev) {
  uint64_t mc = 0;
  int fd = -1;
  int rc = rb_mc_create(nbytes, (int)ndev, &mc, &fd);
  TORCH_CHECK(rc == 0, "cuMulticastCreate failed: ", rc);
  return {(int64_t)mc, (int64_t)fd};
}
int64_t mc_import(int64_t fd) {
  uint64_t mc = 0;
  TORCH_CHECK(rb_mc_import((int)fd, &mc) == 0, "multicast handle import failed");
  return (int64_t)mc;
}
void mc_add_device(int64_t mc, int64_t device) { TORCH_CHECK(rb_mc_add_device((uint64_t)mc, (int)device) == 0, "cuMulticastAddDevice failed"); }
int64_t mc_bind_and_map(int64_t mc, int64_t mem_handle, int64_t nbytes, int64_t align, int64_t device) {
  c10::cuda::CUDAGuard g((c10::DeviceIndex)device);
  uint64_t p = 0;
  int rc = rb_mc_bind_and_map((uint64_t)mc, (uint64_t)mem_handle, nbytes, align, (int)device, &p);
  TORCH_CHECK(rc == 0, "multicast bind/map failed: ", rc);
  return (int64_t)p;
}
void mc_unbind(int64_t mc, int64_t device, int64_t nbytes) { rb_mc_unbind((uint64_t)mc, (int)device, nbytes); }
void close_fd(int64_t fd) { rb_close_fd((int)fd); }

// A tensor view of raw device memory owned elsewhere (the symmetric buffer object keeps the mapping alive).  Built from a
// hand-made storage: `at::from_blob` asks the driver which device the pointer belongs to and rejects VMM mappings of a PEER's
// physical memory ("device of data cuda:1") although they are perfectly usable from this device.
Tensor tensor_from_ptr(int64_t ptr, int64_t nbytes, int64_t device) {
  const at::Device dev(at::kCUDA, (c10::DeviceIndex)device);
  c10::DataPtr dp(reinterpret_cast<void*>(ptr), reinterpret_cast<void*>(ptr), [](void*) {}, dev);
  c10::Storage storage(c10::Storage::use_byte_size_t(), (size_t)nbytes, std::move(dp), /*allocator=*/nullptr, /*resizable=*/false);
  auto t = at::empty({0}, at::TensorOptions().dtype(at::kByte).device(dev));
  t.set_(storage, 0, {nbytes}, {1});
  return t;
}

// ---- NVLS collectives (nvls.cu)
void nvls_allreduce(const c10::optional<Tensor>& in, const c10::optional<Tensor>& out, int64_t nbytes, std::vector<int64_t> data_ptrs,
                    std::vector<int64_t> pad_ptrs, int64_t mc_ptr, int64_t off_in, int64_t off_out, int64_t rank, int64_t dt, int64_t mode,
                    int64_t max_blocks, int64_t device) {
  c10::cuda::CUDAGuard g((c10::DeviceIndex)device);
  if (in.has_value()) TORCH_CHECK(in->is_contiguous() && (int64_t)in->nbytes() == nbytes);
  if (out.has_value()) TORCH_CHECK(out->is_contiguous() && (int64_t)out->nbytes() == nbytes);
  int rc = rb_nvls_allreduce(data_ptrs.data(), pad_ptrs.data(), (uint64_t)mc_ptr, in.has_value() ? in->data_ptr() : nullptr,
                             out.has_value() ? out->data_ptr() : nullptr, nbytes, off_in, off_out, (int)rank, (int)data_ptrs.size(), (int)dt,
                             (int)mode, (int)max_blocks, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "nvls_allreduce failed: ", rc);
}

void nvls_reduce_scatter(std::vector<int64_t> data_ptrs, std::vector<int64_t> pad_ptrs, int64_t mc_ptr, int64_t off, int64_t nbytes, int64_t dt,
                         double scale, Tensor stats, int64_t rank) {
  TORCH_CHECK(stats.is_cuda() && stats.scalar_type() == at::kFloat && stats.numel() >= 2);
  c10::cuda::CUDAGuard g(stats.device());
  int rc = rb_nvls_reduce_scatter(data_ptrs.data(), pad_ptrs.data(), (uint64_t)mc_ptr, off, nbytes, (int)dt, (float)scale, stats.data_ptr<float>(),
                                  (int)rank, (int)data_ptrs.size(), at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "nvls_reduce_scatter failed: ", rc);
}

void nvls_adam_allgather(std::vector<int64_t> data_ptrs, std::vector<int64_t> pad_ptrs, int64_t mc_ptr, int64_t p_off, int64_t p_dt, const Tensor& g,
                         Tensor m, Tensor v, const c10::optional<Tensor>& master, int64_t n, double lr, double b1, double b2, double eps, double wd,
                         int64_t step, const c10::optional<Tensor>& scale, const c10::optional<Tensor>& skip, bool stochastic, int64_t seed,
                         int64_t rank) {
  TORCH_CHECK(g.is_cuda() && m.is_cuda() && v.is_cuda() && g.numel() >= n && m.numel() >= n && v.numel() >= n);
  TORCH_CHECK(m.scalar_type() == v.scalar_type());
  c10::cuda::CUDAGuard gd(g.device());
  int rc = rb_nvls_adam_allgather(data_ptrs.data(), pad_ptrs.data(), (uint64_t)mc_ptr, p_off, (int)p_dt, g.data_ptr(), dtc(g.scalar_type()),
                                  m.data_ptr(), v.data_ptr(), dtc(m.scalar_type()), master.has_value() ? master->data_ptr<float>() : nullptr, n,
                                  (float)lr, (float)b1, (float)b2, (float)eps, (float)wd, (int)step,
                                  scale.has_value() ? scale->data_ptr<float>() : nullptr, skip.has_value() ? skip->data_ptr<int>() : nullptr,
                                  stochastic ? 1 : 0, (uint32_t)seed, (int)rank, (int)data_ptrs.size(), at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "nvls_adam_allgather failed: ", rc);
}

// (h, x_new) = (rmsnorm(sum_ranks(partial) + residual), sum_ranks(partial) + residual); `partial` [rows, H] lives at byte offset `off`
// of every rank's data region.  Without `w` only the reduced (+ residual) tensor is produced.
std::vector<Tensor> nvls_ar_add_rmsnorm(std::vector<int64_t> data_ptrs, std::vector<int64_t> pad_ptrs, int64_t mc_ptr, int64_t off, int64_t rows,
                                        int64_t H, const c10::optional<Tensor>& residual, const c10::optional<Tensor>& w, double eps,
                                        double w_offset, int64_t rank, const Tensor& like) {
  c10::cuda::CUDAGuard g(like.device());
  TORCH_CHECK(like.scalar_type() == at::kBFloat16 || like.scalar_type() == at::kHalf);
  if (residual.has_value()) TORCH_CHECK(residual->is_contiguous() && residual->numel() == rows * H && residual->scalar_type() == like.scalar_type());
  if (w.has_value()) TORCH_CHECK(w->is_contiguous() && w->numel() == H && w->scalar_type() == like.scalar_type());
  auto x_new = at::empty({rows, H}, like.options());
  Tensor y;
  if (w.has_value()) y = at::empty({rows, H}, like.options());
  int rc = rb_nvls_ar_add_rmsnorm(data_ptrs.data(), pad_ptrs.data(), (uint64_t)mc_ptr, off, residual.has_value() ? residual->data_ptr() : nullptr,
                                  w.has_value() ? w->data_ptr() : nullptr, w.has_value() ? y.data_ptr() : nullptr, x_new.data_ptr(), (int)rows,
                                  (int)H, (float)eps, (float)w_offset, (int)rank, (int)data_ptrs.size(), dtc(like.scalar_type()),
                                  at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "nvls_ar_add_rmsnorm failed: ", rc);
  if (w.has_value()) return {y, x_new};
  return {x_new};
}

void nvls_allgather(std::vector<int64_t> data_ptrs, std::vector<int64_t> pad_ptrs, int64_t mc_ptr, int64_t off, int64_t nbytes, bool lead_barrier,
                    int64_t rank, int64_t device) {
  c10::cuda::CUDAGuard g((c10::DeviceIndex)device);
  int rc = rb_nvls_allgather(data_ptrs.data(), pad_ptrs.data(), (uint64_t)mc_ptr, off, nbytes, lead_barrier ? 1 : 0, (int)rank,
                             (int)data_ptrs.size(), at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "nvls_allgather failed: ", rc);
}

// ---- expert-parallel token exchange (ep.cu)
// counts [E] int32 (device) -> [send_tab [E,4], ret_tab [e_local*world,4], per_expert [e_local+1]]; `overflow` is set on the device
// when more rows than cap_rows would arrive here.
std::vector<Tensor> ep_plan(const Tensor& counts, std::vector<int64_t> data_ptrs, std::vector<int64_t> pad_ptrs, int64_t post_off, int64_t e_local,
                            int64_t rank, int64_t parity, Tensor overflow, int64_t cap_rows) {
  TORCH_CHECK(counts.is_cuda() && counts.scalar_type() == at::kInt && counts.is_contiguous());
  TORCH_CHECK(overflow.scalar_type() == at::kInt && overflow.is_cuda());
  const int64_t E = counts.numel(), world = (int64_t)data_ptrs.size();
  c10::cuda::CUDAGuard g(counts.device());
  auto opt = counts.options();
  auto send_tab = at::empty({E, 4}, opt), ret_tab = at::empty({e_local * world, 4}, opt), per_expert = at::empty({e_local + 1}, opt);
  int rc = rb_ep_plan(data_ptrs.data(), pad_ptrs.data(), post_off, counts.data_ptr<int>(), (int)E, (int)e_local, (int)rank, (int)world, (int)parity,
                      send_tab.data_ptr<int>(), ret_tab.data_ptr<int>(), per_expert.data_ptr<int>(), overflow.data_ptr<int>(), (int)cap_rows,
                      at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "ep_plan failed: ", rc);
  return {send_tab, ret_tab, per_expert};
}

// Store the row groups of `src` listed in `table` into the peers' buffers at byte offset dst_off (rows of src.size(1) elements).
void ep_move_rows(const Tensor& src, const Tensor& table, std::vector<int64_t> data_ptrs, std::vector<int64_t> pad_ptrs, int64_t dst_off,
                  int64_t dst_cap_rows, int64_t rank, int64_t blocks) {
  TORCH_CHECK(src.is_cuda() && src.dim() == 2 && src.stride(1) == 1 && table.scalar_type() == at::kInt && table.is_contiguous());
  c10::cuda::CUDAGuard g(src.device());
  const int64_t row_bytes = src.size(1) * src.element_size();
  int rc = rb_ep_move_rows(data_ptrs.data(), pad_ptrs.data(), table.data_ptr<int>(), (int)table.size(0), src.data_ptr(),
                           src.stride(0) * src.element_size(), dst_off, row_bytes, (int)row_bytes, (int)dst_cap_rows, (int)rank,
                           (int)data_ptrs.size(), (int)blocks, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "ep_move_rows failed: ", rc);
}

void register_comm_ops(torch::Library& m) {
  m.def("ep_plan(Tensor counts, int[] data_ptrs, int[] pad_ptrs, int post_off, int e_local, int rank, int parity, Tensor(a!) overflow, int cap_rows) -> Tensor[]", &ep_plan);
  m.def("ep_move_rows(Tensor src, Tensor table, int[] data_ptrs, int[] pad_ptrs, int dst_off, int dst_cap_rows, int rank, int blocks) -> ()", &ep_move_rows);
  m.def("vmm_multicast_supported(int device) -> int", &vmm_multicast_supported);
  m.def("vmm_granularity(int device, int ndev) -> int", &vmm_granularity);
  m.def("vmm_alloc(int nbytes, int align, int device) -> int[]", &vmm_alloc);
  m.def("vmm_import(int fd, int nbytes, int align, int device) -> int[]", &vmm_import);
  m.def("vmm_free(int ptr, int nbytes, int handle) -> ()", &vmm_free);
  m.def("mc_create(int nbytes, int ndev) -> int[]", &mc_create);
  m.def("mc_import(int fd) -> int", &mc_import);
  m.def("mc_add_device(int mc, int device) -> ()", &mc_add_device);
  m.def("mc_bind_and_map(int mc, int mem_handle, int nbytes, int align, int device) -> int", &mc_bind_and_map);
  m.def("mc_unbind(int mc, int device, int nbytes) -> ()", &mc_unbind);
  m.def("close_fd(int fd) -> ()", &close_fd);
  m.def("tensor_from_ptr(int ptr, int nbytes, int device) -> Tensor", &tensor_from_ptr);
  m.def("nvls_allreduce(Tensor? inp, Tensor? out, int nbytes, int[] data_ptrs, int[] pad_ptrs, int mc_ptr, int off_in, int off_out, int rank, int dt, int mode, int max_blocks, int device) -> ()", &nvls_allreduce);
  m.def("nvls_reduce_scatter(int[] data_ptrs, int[] pad_ptrs, int mc_ptr, int off, int nbytes, int dt, float scale, Tensor(a!) stats, int rank) -> ()", &nvls_reduce_scatter);
  m.def("nvls_adam_allgather(int[] data_ptrs, int[] pad_ptrs, int mc_ptr, int p_off, int p_dt, Tensor g, Tensor(a!) m, Tensor(b!) v, Tensor? master, int n, float lr, float b1, float b2, float eps, float wd, int step, Tensor? scale, Tensor? skip, bool stochastic, int seed, int rank) -> ()", &nvls_adam_allgather);
  m.def("nvls_ar_add_rmsnorm(int[] data_ptrs, int[] pad_ptrs, int mc_ptr, int off, int rows, int H, Tensor? residual, Tensor? w, float eps, float w_offset, int rank, Tensor like) -> Tensor[]", &nvls_ar_add_rmsnorm);
  m.def("nvls_allgather(int[] data_ptrs, int[] pad_ptrs, int mc_ptr, int off, int nbytes, bool lead_barrier, int rank, int device) -> ()", &nvls_allgather);
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
