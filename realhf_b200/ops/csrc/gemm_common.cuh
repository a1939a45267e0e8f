// Pieces shared by the tcgen05 GEMM translation units (tile kernel, small-M stream-K kernel, 2-CTA kernel).
#pragma once
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int kThreads = 192;

struct Params {
  void* C;
  const void* bias;
  int64_t ldc;
  int M, N, K;
  int accumulate;  // C += result
  // optional gating of A row-blocks on arrival flags (all-gather -> GEMM: a copy stream fills A chunk by chunk and bumps
  // flag[row / rows_per_flag] to `ready_epoch`); m_rot rotates the m-tile order so the rows that are already local go first
  const uint32_t* ready_flags;
  uint32_t ready_epoch;
  int rows_per_flag;
  int m_rot;
  // optional completion counters (GEMM -> reduce-scatter): every epilogue warp bumps done[row / rows_per_flag] once its 32
  // rows x BN columns of a tile are stored, so a copy stream can ship finished row-blocks while later tiles still compute
  uint32_t* done_counters;
  // optional gated-linear-unit epilogue (2-CTA kernel): B is the fused [gate; up] weight with glu_F rows each, N = glu_F output
  // columns act(gate) * up are written to C, and (when glu_raw != nullptr) the raw gate | up projections to glu_raw [M, 2 glu_F]
  int glu_F;
  int glu_act;       // 0 silu, 1 gelu (tanh approximation)
  void* glu_raw;
  int64_t ld_raw;
};

template <typename T> RB_DEVICE void store_chunk(T* dst, const float* v, int n_valid, bool vec_ok);

template <> RB_DEVICE void store_chunk<float>(float* dst, const float* v, int n_valid, bool vec_ok) {
  if (n_valid == 32 && vec_ok) {
#pragma unroll
    for (int i = 0; i < 8; ++i) reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
    for (int i = 0; i < n_valid; ++i) dst[i] = v[i];
  }
}
template <typename T> RB_DEVICE void store_chunk(T* dst, const float* v, int n_valid, bool vec_ok) {
  if (n_valid == 32 && vec_ok) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      rb::Pack<T, 8> p;
#pragma unroll
      for (int k = 0; k < 8; ++k) p.v[k] = rb::from_f<T>(v[8 * i + k]);
      reinterpret_cast<rb::Pack<T, 8>*>(dst)[i] = p;
    }
  } else {
    for (int i = 0; i < n_valid; ++i) dst[i] = rb::from_f<T>(v[i]);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D row-major tensor [rows, cols] (cols contiguous, row pitch `ld` elements), 2-byte elements, 128B swizzle.
bool make_tmap(CUtensorMap* m, const void* ptr, int is_bf16, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
               uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  // The driver entry point needs a current context in *this* thread; autograd's backward threads only ever
  // called cudaSetDevice, which does not bind the primary context for driver-API calls until a runtime call does.
  static thread_local bool ctx_ready = false;
  if (!ctx_ready) { cudaFree(nullptr); ctx_ready = true; }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[rb_gemm] cuTensorMapEncodeTiled failed: %d ptr=%p rows=%llu cols=%llu ld=%llu box=(%u,%u)\n", (int)r, ptr,
            (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_cols, box_rows);
  }
  return r == CUDA_SUCCESS;
}


// 2-D row-major tensor of BYTES [rows, cols] (fp8 operands), 128B swizzle: a box row is 128 elements = one swizzle row.
bool make_tmap_u8(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  static thread_local bool ctx_ready = false;
  if (!ctx_ready) { cudaFree(nullptr); ctx_ready = true; }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "[rb_gemm] cuTensorMapEncodeTiled(u8) failed: %d rows=%llu cols=%llu ld=%llu\n", (int)r,
                                 (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
  return r == CUDA_SUCCESS;
}

}  // namespace
