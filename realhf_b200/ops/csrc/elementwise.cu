// Small bandwidth-bound fused ops: rotary embedding on packed q/k, gated activation (SwiGLU / GeGLU).
//
// RoPE replaces the Triton `apply_rotary` the reference borrows from flash-attn
// (modules/rotary.py:10, utils/functional.py:447); it rotates q and k *in place inside the fused
// QKV projection output* (row stride = (nq+2nkv)*hd) so no q/k copies are materialised.
// The gated activation reads the fused gate|up projection output [T, 2F] once and writes [T, F]
// (reference: eager `silu(gate) * up`, modules/mlp.py:411-422).
#include "common.cuh"

namespace {

// x: [T, row_stride]; heads [0, n_heads) of width hd at the start of each row are rotated.
// cos/sin: [max_pos, hd/2] fp32.  pos: [T] int32.  inverse=1 applies the transpose rotation (backward).
template <typename T, bool kInterleaved>
__global__ void __launch_bounds__(256) rope_kernel(T* __restrict__ x, const float* __restrict__ cs,
                                                   const float* __restrict__ sn, const int* __restrict__ pos,
                                                   int64_t n_tok, int n_heads, int hd, int64_t row_stride, int rot_dim,
                                                   int inverse) {
  // one thread handles 8 rotation pairs
  const int half = rot_dim / 2;
  const int chunks_per_head = half / 8;
  const int64_t total = n_tok * n_heads * chunks_per_head;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % chunks_per_head);
    const int h = (int)((idx / chunks_per_head) % n_heads);
    const int64_t t = idx / ((int64_t)chunks_per_head * n_heads);
    const int p = pos[t];
    const float* cp = cs + (int64_t)p * half + c * 8;
    const float* sp = sn + (int64_t)p * half + c * 8;
    T* base = x + t * row_stride + (int64_t)h * hd;
    float co[8], si[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { co[k] = cp[k]; si[k] = inverse ? -sp[k] : sp[k]; }
    if constexpr (kInterleaved) {
      // pairs (2i, 2i+1): 16 contiguous elements
      rb::Pack<T, 8> a = reinterpret_cast<rb::Pack<T, 8>*>(base + c * 16)[0];
      rb::Pack<T, 8> b = reinterpret_cast<rb::Pack<T, 8>*>(base + c * 16)[1];
      T* e = reinterpret_cast<T*>(&a);  // a|b contiguous in registers is not guaranteed: handle explicitly
      float v[16];
#pragma unroll
      for (int k = 0; k < 8; ++k) { v[k] = rb::to_f(a.v[k]); v[8 + k] = rb::to_f(b.v[k]); }
      (void)e;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float x0 = v[2 * k], x1 = v[2 * k + 1];
        v[2 * k] = x0 * co[k] - x1 * si[k];
        v[2 * k + 1] = x1 * co[k] + x0 * si[k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { a.v[k] = rb::from_f<T>(v[k]); b.v[k] = rb::from_f<T>(v[8 + k]); }
      reinterpret_cast<rb::Pack<T, 8>*>(base + c * 16)[0] = a;
      reinterpret_cast<rb::Pack<T, 8>*>(base + c * 16)[1] = b;
    } else {
      // pairs (i, i+half)
      rb::Pack<T, 8> a = *reinterpret_cast<rb::Pack<T, 8>*>(base + c * 8);
      rb::Pack<T, 8> b = *reinterpret_cast<rb::Pack<T, 8>*>(base + half + c * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float x0 = rb::to_f(a.v[k]), x1 = rb::to_f(b.v[k]);
        a.v[k] = rb::from_f<T>(x0 * co[k] - x1 * si[k]);
        b.v[k] = rb::from_f<T>(x1 * co[k] + x0 * si[k]);
      }
      *reinterpret_cast<rb::Pack<T, 8>*>(base + c * 8) = a;
      *reinterpret_cast<rb::Pack<T, 8>*>(base + half + c * 8) = b;
    }
  }
}

RB_DEVICE float act_fwd(float g, int kind) {
  if (kind == 0) return g / (1.f + __expf(-g));                                   // silu
  const float u = 0.7978845608028654f * (g + 0.044715f * g * g * g);              // gelu (tanh approximation)
  return 0.5f * g * (1.f + tanhf(u));
}
RB_DEVICE float act_bwd(float g, int kind) {
  if (kind == 0) { const float s = 1.f / (1.f + __expf(-g)); return s * (1.f + g * (1.f - s)); }
  const float u = 0.7978845608028654f * (g + 0.044715f * g * g * g);
  const float th = tanhf(u);
  return 0.5f * (1.f + th) + 0.5f * g * (1.f - th * th) * 0.7978845608028654f * (1.f + 3.f * 0.044715f * g * g);
}

template <typename T>
__global__ void __launch_bounds__(256) gated_act_fwd_kernel(const T* __restrict__ gu, T* __restrict__ out, int64_t n_tok,
                                                            int F, int kind) {
  const int vec_per_row = F / 8;
  const int64_t total = n_tok * vec_per_row;
  rb::pdl_trigger();  // successors may start launching (and prefetching weights) while this kernel still waits below
  rb::pdl_wait();
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx / vec_per_row;
    const int c = (int)(idx % vec_per_row);
    rb::Pack<T, 8> g = *reinterpret_cast<const rb::Pack<T, 8>*>(gu + t * 2 * F + c * 8);
    rb::Pack<T, 8> u = *reinterpret_cast<const rb::Pack<T, 8>*>(gu + t * 2 * F + F + c * 8);
    rb::Pack<T, 8> o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = rb::from_f<T>(act_fwd(rb::to_f(g.v[k]), kind) * rb::to_f(u.v[k]));
    *reinterpret_cast<rb::Pack<T, 8>*>(out + t * F + c * 8) = o;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gated_act_bwd_kernel(const T* __restrict__ gu, const T* __restrict__ dout,
                                                            T* __restrict__ dgu, int64_t n_tok, int F, int kind) {
  const int vec_per_row = F / 8;
  const int64_t total = n_tok * vec_per_row;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx / vec_per_row;
    const int c = (int)(idx % vec_per_row);
    rb::Pack<T, 8> g = *reinterpret_cast<const rb::Pack<T, 8>*>(gu + t * 2 * F + c * 8);
    rb::Pack<T, 8> u = *reinterpret_cast<const rb::Pack<T, 8>*>(gu + t * 2 * F + F + c * 8);
    rb::Pack<T, 8> d = *reinterpret_cast<const rb::Pack<T, 8>*>(dout + t * F + c * 8);
    rb::Pack<T, 8> dg, du;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float gf = rb::to_f(g.v[k]), uf = rb::to_f(u.v[k]), df = rb::to_f(d.v[k]);
      dg.v[k] = rb::from_f<T>(df * uf * act_bwd(gf, kind));
      du.v[k] = rb::from_f<T>(df * act_fwd(gf, kind));
    }
    *reinterpret_cast<rb::Pack<T, 8>*>(dgu + t * 2 * F + c * 8) = dg;
    *reinterpret_cast<rb::Pack<T, 8>*>(dgu + t * 2 * F + F + c * 8) = du;
  }
}

inline int grid_for(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = (int64_t)rb::kNumSMs * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

extern "C" {

static int g_rb_pdl = [] { const char* e = getenv("REAL_PDL"); return (e == nullptr || atoi(e) != 0) ? 1 : 0; }();
int rb_get_pdl() { return g_rb_pdl; }
void rb_set_pdl(int on) { g_rb_pdl = on ? 1 : 0; }


int rb_rope_inplace(void* x, const float* cs, const float* sn, const int* pos, int64_t n_tok, int n_heads, int hd,
                    int64_t row_stride, int rot_dim, int interleaved, int inverse, int dt, cudaStream_t s) {
  if (n_tok == 0) return 0;
  if (rot_dim % 16 != 0 || hd % 8 != 0 || row_stride % 8 != 0) return -1;
  const int64_t total = n_tok * n_heads * (rot_dim / 16);
#define RB_L(T)                                                                                                         \
  if (interleaved) rope_kernel<T, true><<<grid_for(total), 256, 0, s>>>((T*)x, cs, sn, pos, n_tok, n_heads, hd, row_stride, rot_dim, inverse); \
  else rope_kernel<T, false><<<grid_for(total), 256, 0, s>>>((T*)x, cs, sn, pos, n_tok, n_heads, hd, row_stride, rot_dim, inverse);
  if (dt == 0) return -1; else if (dt == 1) { RB_L(__nv_bfloat16) } else if (dt == 2) { RB_L(__half) } else return -1;
#undef RB_L
  return 0;
}

int rb_gated_act_fwd(const void* gu, void* out, int64_t n_tok, int F, int kind, int dt, cudaStream_t s) {
  if (n_tok == 0) return 0;
  if (F % 8 != 0) return -1;
  const int64_t total = n_tok * (F / 8);
  if (dt == 1) rb::launch_pdl(gated_act_fwd_kernel<__nv_bfloat16>, dim3(grid_for(total)), dim3(256), 0, s, (const __nv_bfloat16*)gu, (__nv_bfloat16*)out, n_tok, F, kind);
  else if (dt == 2) rb::launch_pdl(gated_act_fwd_kernel<__half>, dim3(grid_for(total)), dim3(256), 0, s, (const __half*)gu, (__half*)out, n_tok, F, kind);
  else return -1;
  return 0;
}

int rb_gated_act_bwd(const void* gu, const void* dout, void* dgu, int64_t n_tok, int F, int kind, int dt, cudaStream_t s) {
  if (n_tok == 0) return 0;
  if (F % 8 != 0) return -1;
  const int64_t total = n_tok * (F / 8);
  if (dt == 1) gated_act_bwd_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, s>>>((const __nv_bfloat16*)gu, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dgu, n_tok, F, kind);
  else if (dt == 2) gated_act_bwd_kernel<__half><<<grid_for(total), 256, 0, s>>>((const __half*)gu, (const __half*)dout, (__half*)dgu, n_tok, F, kind);
  else return -1;
  return 0;
}

}  // extern "C"
