// NVLS collectives for sm_100a: kernels that talk to the NVSwitch multicast address of a symmetric buffer.
//
//   multimem.ld_reduce  — ONE load returns the sum of the same address over all GPUs, reduced inside the switch
//   multimem.st         — ONE store lands in every GPU's copy
//
// so a reduce-scatter moves 1/world of the buffer INTO each GPU (instead of (world-1)/world), and an all-gather moves
// 1/world OUT of each GPU.  Built on the VMM symmetric memory of vmm.cpp (`parallel/symm_mem.py::VmmSymmetricBuffer`).
//
// Kernels:
//   nvls_allreduce_kernel      one-shot (every rank ld_reduces everything) / two-shot (ld_reduce my slice, multimem.st it)
//                              all-reduce for the tensor-parallel decode path; graph-capturable (epoch barriers)
//   nvls_rs_sumsq_kernel       ZeRO gradient reduce-scatter fused with the 1/dp average, the bf16 cast and the
//                              sum-of-squares / non-finite statistics of the global grad-norm (K17/K18/C8 of SURVEY §2)
//   nvls_adam_ag_kernel        partitioned AdamW on this rank's shard whose parameter store IS the all-gather: the updated
//                              bf16 weights are written once with multimem.st and arrive in every replica (C9)
//   nvls_allgather_kernel      plain shard broadcast (ZeRO-3 materialise, replica refresh)
//
// Reference counterparts: Megatron's DistributedOptimizer reduce-scatter / all-gather through NCCL
// (realhf/impl/model/backend/megatron.py:883-909, :518) and the dormant custom all-reduce
// (csrc/custom_all_reduce/custom_all_reduce.cuh:173-238).
#include <type_traits>

#include "adam_math.cuh"
#include "comm_common.cuh"

namespace {
using namespace rbcomm;
using namespace rbadam;

// ---------------------------------------------------------------------------------------------- multimem primitives
template <typename T> RB_DEVICE int4 mm_ld_reduce(const void* mc);
template <> RB_DEVICE int4 mm_ld_reduce<__nv_bfloat16>(const void* mc) {
  int4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc) : "memory");
  return r;
}
template <> RB_DEVICE int4 mm_ld_reduce<__half>(const void* mc) {
  int4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc) : "memory");
  return r;
}
template <> RB_DEVICE int4 mm_ld_reduce<float>(const void* mc) {
  int4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc) : "memory");
  return r;
}
RB_DEVICE void mm_st(void* mc, const int4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

template <typename T> struct VecOf { static constexpr int N = 16 / sizeof(T); };

// bulk (ZeRO) kernels: 64 blocks x 1024 threads x 4 vectors in flight = 4 MB outstanding per GPU, enough to cover the NVLink
// round trip of multimem loads at full link rate; the epoch barrier only uses the first `world` threads of a block
constexpr int kBwThreads = 1024;

// ---------------------------------------------------------------------------------------------- all-reduce
// Input: `nvec` 16-byte vectors at byte offset off_in of every rank's data region (staged from `in` first when given).
// kTwoShot = false: out[i] = switch-sum(i) for all i; no trailing barrier — callers alternate between two regions.
// kTwoShot = true : this rank reduces slice [rank*per, ...) and multicasts it to offset off_out of EVERY rank; after the
//                   trailing barrier the full result sits in this rank's own data region (and in `out` when given).
template <typename T, bool kTwoShot>
__global__ void __launch_bounds__(kThreads) nvls_allreduce_kernel(Peers P, uint8_t* __restrict__ mc, const int4* __restrict__ in,
                                                                  int4* __restrict__ out, int64_t off_in, int64_t off_out,
                                                                  int64_t nvec, int rank, int world) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // two-shot: block b owns the same vector chunk in every phase and on every rank, so the same-index block barrier covers
  // the staging copy as well (a block only reads what the peers' block of the same index staged)
  const int64_t chunk = (nvec + gridDim.x - 1) / gridDim.x;
  const int64_t c0 = min(nvec, (int64_t)blockIdx.x * chunk), c1 = min(nvec, c0 + chunk);
  if (in != nullptr) {
    int4* mine = reinterpret_cast<int4*>(P.data[rank] + off_in);
    if constexpr (kTwoShot) {
      for (int64_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) mine[i] = in[i];
    } else {
      for (int64_t i = i0; i < nvec; i += stride) mine[i] = in[i];
    }
  }
  block_barrier(P, rank, world);
  const int4* src = reinterpret_cast<const int4*>(mc + off_in);
  if constexpr (!kTwoShot) {
    for (int64_t i = i0; i < nvec; i += stride) out[i] = mm_ld_reduce<T>(src + i);
  } else {
    const int64_t per = (c1 - c0 + world - 1) / world;
    const int64_t lo = min(c1, c0 + rank * per), hi = min(c1, lo + per);
    int4* dst_mc = reinterpret_cast<int4*>(mc + off_out);
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) mm_st(dst_mc + i, mm_ld_reduce<T>(src + i));
    __threadfence_system();
    block_barrier(P, rank, world);
    if (out != nullptr) {
      const int4* res = reinterpret_cast<const int4*>(P.data[rank] + off_out);
      for (int64_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) out[i] = res[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------- ZeRO reduce-scatter
// g_shard (in place, this rank's copy) = scale * switch-sum over ranks of the gradient bucket slice at byte offset `off`;
// stats[0] += sum of squares of the reduced+scaled values, stats[1] += number of non-finite ones.
// The leading barrier makes every rank's backward writes of this bucket visible; a barrier before the gradients are
// overwritten again is the caller's (the all-gather kernel's trailing barrier provides it).
template <typename TG>
__global__ void __launch_bounds__(kBwThreads) nvls_rs_sumsq_kernel(Peers P, uint8_t* __restrict__ mc, int64_t off, int64_t nvec, float scale,
                                                                 float* __restrict__ stats, int rank, int world) {
  __shared__ float red[32];
  block_barrier(P, rank, world);
  constexpr int V = VecOf<TG>::N;
  const int4* src = reinterpret_cast<const int4*>(mc + off);
  int4* dst = reinterpret_cast<int4*>(P.data[rank] + off);
  float acc = 0.f, bad = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < nvec; base += stride * U) {
    int4 r[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (base + u * stride < nvec) r[u] = mm_ld_reduce<TG>(src + base + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + u * stride;
      if (i >= nvec) continue;
      TG* e = reinterpret_cast<TG*>(&r[u]);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float x = rb::to_f(e[k]) * scale;
        acc = fmaf(x, x, acc);
        bad += isfinite(x) ? 0.f : 1.f;
        e[k] = rb::from_f<TG>(x);
      }
      dst[i] = r[u];
    }
  }
  acc = rb::block_reduce<false>(acc, red);
  bad = rb::block_reduce<false>(bad, red);
  if (threadIdx.x == 0) {
    atomicAdd(stats, acc);
    if (bad != 0.f) atomicAdd(stats + 1, bad);
  }
}

// ---------------------------------------------------------------------------------------------- AdamW + all-gather
// AdamW on `n` elements (multiple of 8 for 2-byte params, of 4 for fp32) of this rank's shard.  The parameter is read from
// the local copy (`p`), the update is stored through the multicast mapping (`p_mc`, same offset), i.e. into every replica.
template <typename TP, typename TG, typename TS, bool kMaster, bool kStochastic>
__global__ void __launch_bounds__(kThreads) nvls_adam_ag_kernel(Peers P, const TP* __restrict__ p, TP* __restrict__ p_mc,
                                                                const TG* __restrict__ g, TS* __restrict__ m, TS* __restrict__ v,
                                                                float* __restrict__ master, int64_t n, float lr, float b1, float b2,
                                                                float eps, float wd, float bc1, float bc2,
                                                                const float* __restrict__ scale_ptr, const int* __restrict__ skip_ptr,
                                                                uint32_t seed, int rank, int world) {
  const bool skip = skip_ptr != nullptr && *skip_ptr != 0;
  if (!skip) {
    const float gscale = scale_ptr ? *scale_ptr : 1.f;
    const float inv_bc1 = 1.f / bc1, inv_bc2 = 1.f / bc2;
    constexpr int V = VecOf<TP>::N;  // one 16-byte multicast store per iteration
    const int64_t nvec = n / V;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
      rb::Pack<TP, V> pp = reinterpret_cast<const rb::Pack<TP, V>*>(p)[i];
      rb::Pack<TG, V> gg = reinterpret_cast<const rb::Pack<TG, V>*>(g)[i];
      rb::Pack<TS, V> mm = reinterpret_cast<rb::Pack<TS, V>*>(m)[i];
      rb::Pack<TS, V> vv = reinterpret_cast<rb::Pack<TS, V>*>(v)[i];
      rb::Pack<float, V> ms;
      if constexpr (kMaster) ms = reinterpret_cast<rb::Pack<float, V>*>(master)[i];
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float grad = rb::to_f(gg.v[k]) * gscale;
        float mk = rb::to_f(mm.v[k]), vk = rb::to_f(vv.v[k]);
        const float w = adam_elem(kMaster ? ms.v[k] : rb::to_f(pp.v[k]), grad, mk, vk, lr, b1, b2, eps, wd, inv_bc1, inv_bc2);
        mm.v[k] = rb::from_f<TS>(mk);
        vv.v[k] = rb::from_f<TS>(vk);
        if constexpr (kMaster) ms.v[k] = w;
        if constexpr (kStochastic) pp.v[k] = sr_bf16(w, hash32(seed ^ (uint32_t)(i * V + k)));
        else pp.v[k] = rb::from_f<TP>(w);
      }
      mm_st(reinterpret_cast<rb::Pack<TP, V>*>(p_mc) + i, *reinterpret_cast<const int4*>(&pp));
      reinterpret_cast<rb::Pack<TS, V>*>(m)[i] = mm;
      reinterpret_cast<rb::Pack<TS, V>*>(v)[i] = vv;
      if constexpr (kMaster) reinterpret_cast<rb::Pack<float, V>*>(master)[i] = ms;
    }
  }
  __threadfence_system();
  block_barrier(P, rank, world);  // every replica has received every shard when the kernel ends anywhere
}

// Broadcast `nvec` vectors of this rank's shard (local pointer `src`) to the same offset of every replica.
__global__ void __launch_bounds__(kBwThreads) nvls_allgather_kernel(Peers P, const int4* __restrict__ src, int4* __restrict__ dst_mc,
                                                                  int64_t nvec, int lead_barrier, int rank, int world) {
  if (lead_barrier) block_barrier(P, rank, world);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < nvec; base += stride * U) {
    int4 r[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (base + u * stride < nvec) r[u] = rb::ld_stream(src + base + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (base + u * stride < nvec) mm_st(dst_mc + base + u * stride, r[u]);
  }
  __threadfence_system();
  block_barrier(P, rank, world);
}

// ---------------------------------------------------------------------------------------------- AR + residual + RMSNorm
// The tensor-parallel decode layer boundary in ONE kernel: the row-parallel GEMM (o-proj / down-proj) wrote its partial
// [B, H] output into symmetric memory at byte offset `off`; this kernel waits for every rank's partial (epoch barrier),
// pulls the in-switch sum row by row (`multimem.ld_reduce`), adds the residual stream, writes the new residual and the
// RMS-normalised activations for the next column-parallel GEMM.  Replaces [all-reduce kernel] + [add+RMSNorm kernel]
// (reference: RowParallelLinear all-reduce, modules.py:1010 + eager _LlamaRMSNorm, modules/mlp.py:425-444).
// No trailing barrier: callers alternate between two regions (see FusedTP.symm_out).
template <typename T, int kMaxVec>
__global__ void __launch_bounds__(kThreads) nvls_ar_add_rmsnorm_kernel(Peers P, const uint8_t* __restrict__ mc, int64_t off,
                                                                       const T* __restrict__ res_in, const T* __restrict__ w,
                                                                       T* __restrict__ y, T* __restrict__ res_out, int rows, int H,
                                                                       float eps, float w_offset, int rank, int world) {
  __shared__ float red[32];
  constexpr int V = 8;
  const int nvec = H / V;
  rb::pdl_trigger();
  rb::pdl_wait();  // the producing GEMM of THIS rank is complete; the barrier below covers the peers'
  block_barrier(P, rank, world);
  const rb::Pack<T, V>* wr = reinterpret_cast<const rb::Pack<T, V>*>(w);
  // two rows per iteration: both rows' multimem loads are issued before either is consumed (a block of the decode path owns
  // rows / 64 <= 2 rows, so the whole kernel is ONE NVLink round trip after the barrier)
  constexpr int R = 2;
  for (int row0 = blockIdx.x; row0 < rows; row0 += R * gridDim.x) {
    int4 raw[R][kMaxVec];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r * gridDim.x;
      if (row < rows) {
        const int4* src = reinterpret_cast<const int4*>(mc + off + (int64_t)row * H * sizeof(T));
#pragma unroll
        for (int it = 0; it < kMaxVec; ++it) {
          const int i = threadIdx.x + it * kThreads;
          if (i < nvec) raw[r][it] = mm_ld_reduce<T>(src + i);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r * gridDim.x;
      if (row >= rows) break;
      float vals[kMaxVec][V];
      float ss = 0.f;
#pragma unroll
      for (int it = 0; it < kMaxVec; ++it) {
        const int i = threadIdx.x + it * kThreads;
        if (i < nvec) {
          const T* a = reinterpret_cast<const T*>(&raw[r][it]);
          if (res_in != nullptr) {
            rb::Pack<T, V> bb = reinterpret_cast<const rb::Pack<T, V>*>(res_in + (int64_t)row * H)[i];
            rb::Pack<T, V> o;
#pragma unroll
            for (int k = 0; k < V; ++k) {
              o.v[k] = rb::from_f<T>(rb::to_f(a[k]) + rb::to_f(bb.v[k]));
              vals[it][k] = rb::to_f(o.v[k]);
            }
            reinterpret_cast<rb::Pack<T, V>*>(res_out + (int64_t)row * H)[i] = o;
          } else {
#pragma unroll
            for (int k = 0; k < V; ++k) vals[it][k] = rb::to_f(a[k]);
            if (res_out != nullptr) reinterpret_cast<int4*>(res_out + (int64_t)row * H)[i] = raw[r][it];
          }
#pragma unroll
          for (int k = 0; k < V; ++k) ss = fmaf(vals[it][k], vals[it][k], ss);
        }
      }
      if (y == nullptr) continue;  // plain all-reduce (+ residual) without a norm
      ss = rb::block_reduce<false>(ss, red);
      const float rstd = rsqrtf(ss / (float)H + eps);
#pragma unroll
      for (int it = 0; it < kMaxVec; ++it) {
        const int i = threadIdx.x + it * kThreads;
        if (i < nvec) {
          rb::Pack<T, V> ww = wr[i], o;
#pragma unroll
          for (int k = 0; k < V; ++k) o.v[k] = rb::from_f<T>(vals[it][k] * rstd * (rb::to_f(ww.v[k]) + w_offset));
          reinterpret_cast<rb::Pack<T, V>*>(y + (int64_t)row * H)[i] = o;
        }
      }
    }
  }
}

Peers make_peers(const int64_t* data_ptrs, const int64_t* pad_ptrs, int world) {
  Peers P;
  for (int i = 0; i < kMaxRanks; ++i) {
    P.data[i] = i < world ? reinterpret_cast<uint8_t*>(data_ptrs[i]) : nullptr;
    P.pad[i] = i < world ? reinterpret_cast<uint32_t*>(pad_ptrs[i]) : nullptr;
  }
  return P;
}

int grid_for(int64_t nvec, int per_thread) {
  int64_t b = (nvec + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
  return (int)(b < 1 ? 1 : (b > kMaxBlocks ? kMaxBlocks : b));
}

}  // namespace

extern "C" {

// dt: 0 fp32, 1 bf16, 2 fp16.  mode 1: one-shot into `out`; mode 2: two-shot (result at off_out of the data region, copied to
// `out` when non-null).  `in` may be null when the producer already wrote offset off_in of this rank's data region.
int rb_nvls_allreduce(const int64_t* data_ptrs, const int64_t* pad_ptrs, uint64_t mc, const void* in, void* out, int64_t nbytes,
                      int64_t off_in, int64_t off_out, int rank, int world, int dt, int mode, int max_blocks, cudaStream_t s) {
  if (world > kMaxRanks || (nbytes & 15) || (off_in & 15) || (off_out & 15) || mc == 0) return -1;
  if (mode == 1 && out == nullptr) return -4;
  Peers P = make_peers(data_ptrs, pad_ptrs, world);
  const int64_t nvec = nbytes / 16;
  int nb = grid_for(nvec, mode == 1 ? 2 : 1);
  if (max_blocks > 0 && nb > max_blocks) nb = max_blocks;
#define RB_GO(T)                                                                                                                   \
  if (mode == 1) nvls_allreduce_kernel<T, false><<<nb, kThreads, 0, s>>>(P, (uint8_t*)mc, (const int4*)in, (int4*)out, off_in, off_out, nvec, rank, world); \
  else nvls_allreduce_kernel<T, true><<<nb, kThreads, 0, s>>>(P, (uint8_t*)mc, (const int4*)in, (int4*)out, off_in, off_out, nvec, rank, world);
  if (dt == 0) { RB_GO(float) } else if (dt == 1) { RB_GO(__nv_bfloat16) } else if (dt == 2) { RB_GO(__half) } else return -2;
#undef RB_GO
  return 0;
}

int rb_nvls_reduce_scatter(const int64_t* data_ptrs, const int64_t* pad_ptrs, uint64_t mc, int64_t off, int64_t nbytes, int dt, float scale,
                           float* stats, int rank, int world, cudaStream_t s) {
  if (world > kMaxRanks || (nbytes & 15) || (off & 15) || mc == 0) return -1;
  if (nbytes == 0) return 0;
  Peers P = make_peers(data_ptrs, pad_ptrs, world);
  const int64_t nvec = nbytes / 16;
  if (dt == 0) nvls_rs_sumsq_kernel<float><<<kMaxBlocks, kBwThreads, 0, s>>>(P, (uint8_t*)mc, off, nvec, scale, stats, rank, world);
  else if (dt == 1) nvls_rs_sumsq_kernel<__nv_bfloat16><<<kMaxBlocks, kBwThreads, 0, s>>>(P, (uint8_t*)mc, off, nvec, scale, stats, rank, world);
  else return -2;
  return 0;
}

// p_off: byte offset of the shard inside the PARAMETER symmetric buffer (data_ptrs / mc describe that buffer).
int rb_nvls_adam_allgather(const int64_t* data_ptrs, const int64_t* pad_ptrs, uint64_t mc, int64_t p_off, int p_dt, const void* g, int g_dt,
                           void* m, void* v, int s_dt, float* master, int64_t n, float lr, float b1, float b2, float eps, float wd,
                           int step, const float* scale_ptr, const int* skip_ptr, int stochastic, uint32_t seed, int rank, int world,
                           cudaStream_t s) {
  if (world > kMaxRanks || mc == 0 || (p_off & 15)) return -1;
  const int64_t vec = p_dt == 0 ? 4 : 8;
  if (n % vec) return -3;
  Peers P = make_peers(data_ptrs, pad_ptrs, world);
  const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  void* p = P.data[rank] + p_off;
  void* p_mc = reinterpret_cast<uint8_t*>(mc) + p_off;
  const int key = p_dt * 4 + g_dt * 2 + s_dt;
#define RB_K(TP, TG, TS, MASTER, SR)                                                                                              \
  nvls_adam_ag_kernel<TP, TG, TS, MASTER, SR><<<kMaxBlocks, kThreads, 0, s>>>(P, (const TP*)p, (TP*)p_mc, (const TG*)g, (TS*)m, (TS*)v, master, n, lr, \
                                                                              b1, b2, eps, wd, bc1, bc2, scale_ptr, skip_ptr, seed, rank, world)
#define RB_CASE(K, TP, TG, TS)                                                                         \
  case K:                                                                                              \
    if (master) RB_K(TP, TG, TS, true, false);                                                         \
    else if (std::is_same<TP, __nv_bfloat16>::value && stochastic) RB_K(TP, TG, TS, false, true);      \
    else RB_K(TP, TG, TS, false, false);                                                               \
    break;
  switch (key) {
    RB_CASE(0, float, float, float)
    RB_CASE(2, float, __nv_bfloat16, float)
    RB_CASE(4, __nv_bfloat16, float, float)
    RB_CASE(5, __nv_bfloat16, float, __nv_bfloat16)
    RB_CASE(6, __nv_bfloat16, __nv_bfloat16, float)
    RB_CASE(7, __nv_bfloat16, __nv_bfloat16, __nv_bfloat16)
    default: return -2;
  }
#undef RB_CASE
#undef RB_K
  return 0;
}

// y = rmsnorm(allreduce(partial) + res_in) * (w + w_offset); res_out = allreduce(partial) + res_in.  `partial` is the [rows, H]
// tensor every rank wrote at byte offset `off` of its data region.  y / res_in may be null (plain all-reduce into res_out).
int rb_nvls_ar_add_rmsnorm(const int64_t* data_ptrs, const int64_t* pad_ptrs, uint64_t mc, int64_t off, const void* res_in, const void* w,
                           void* y, void* res_out, int rows, int H, float eps, float w_offset, int rank, int world, int dt, cudaStream_t s) {
  if (world > kMaxRanks || mc == 0 || (off & 15) || H % 8 || H > 8 * kThreads * 4) return -1;
  if (rows == 0) return 0;
  Peers P = make_peers(data_ptrs, pad_ptrs, world);
  const int nb = rows < kMaxBlocks ? rows : kMaxBlocks;
  const int nvec = H / 8;
#define RB_GO(T, MV)                                                                                                               \
  rb::launch_pdl(nvls_ar_add_rmsnorm_kernel<T, MV>, dim3(nb), dim3(kThreads), 0, s, P, (const uint8_t*)mc, off, (const T*)res_in, (const T*)w, \
                 (T*)y, (T*)res_out, rows, H, eps, w_offset, rank, world)
#define RB_T(T) { if (nvec <= kThreads) RB_GO(T, 1); else if (nvec <= 2 * kThreads) RB_GO(T, 2); else RB_GO(T, 4); }
  if (dt == 1) RB_T(__nv_bfloat16) else if (dt == 2) RB_T(__half) else return -2;
#undef RB_T
#undef RB_GO
  return 0;
}

int rb_nvls_allgather(const int64_t* data_ptrs, const int64_t* pad_ptrs, uint64_t mc, int64_t off, int64_t nbytes, int lead_barrier, int rank,
                      int world, cudaStream_t s) {
  if (world > kMaxRanks || (nbytes & 15) || (off & 15) || mc == 0) return -1;
  Peers P = make_peers(data_ptrs, pad_ptrs, world);
  nvls_allgather_kernel<<<kMaxBlocks, kBwThreads, 0, s>>>(P, (const int4*)(P.data[rank] + off), (int4*)(reinterpret_cast<uint8_t*>(mc) + off),
                                                        nbytes / 16, lead_barrier, rank, world);
  return 0;
}

}  // extern "C"
