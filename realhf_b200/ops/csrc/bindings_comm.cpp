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
int rb_reduce_slabs(const void*, void*, int64_t, int64_t, int, const uint32_t*, uint32_t, int, cudaStream_t);
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

// out[nbytes] = sum_s slabs[s] where slab s starts at base_ptr + s*slab_bytes; waits for *counter >= expect first.
void reduce_slabs(int64_t base_ptr, Tensor out, int64_t slab_bytes, int64_t world, int64_t counter_ptr, int64_t expect) {
  c10::cuda::CUDAGuard g(out.device());
  int rc = rb_reduce_slabs(reinterpret_cast<const void*>(base_ptr), out.data_ptr(), out.nbytes(), slab_bytes, (int)world,
                           reinterpret_cast<const uint32_t*>(counter_ptr), (uint32_t)expect, dtc(out.scalar_type()),
                           at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, "reduce_slabs failed: ", rc);
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
  m.def("reduce_slabs(int base_ptr, Tensor(a!) out, int slab_bytes, int world, int counter_ptr, int expect) -> ()", &reduce_slabs);
}
