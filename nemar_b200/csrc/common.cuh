// common.cuh — shared device/host helpers for libnemar_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/nemar_b200.h"

#define NEMAR_API extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void nemar_set_error(const char* fmt, ...);
void nemar_note_conv_kernel(const char* fmt, ...);

#define NEMAR_REQUIRE(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      nemar_set_error(__VA_ARGS__);       \
      return -1;                          \
    }                                     \
  } while (0)

#define NEMAR_LAUNCH_CHECK()                                            \
  do {                                                                  \
    cudaError_t e__ = cudaPeekAtLastError();                            \
    if (e__ != cudaSuccess) {                                           \
      nemar_set_error("%s:%d launch failed: %s", __FILE__, __LINE__,    \
                      cudaGetErrorString(e__));                         \
      return (int)e__;                                                  \
    }                                                                   \
  } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int grid_for(int64_t work, int block, int64_t cap = (int64_t)148 * 32) {
  int64_t g = ceil_div64(work, block);
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------------------------------------
// element access
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// Device-side mirror of nemar_tensor with index helpers.
struct TView {
  char* ptr;
  int n, h, w, c, pad, cs, coff, dtype;
  int hp, wp;  // padded extents
  __host__ __device__ __forceinline__ int64_t pix(int nn, int y, int x) const {  // interior coords
    return ((int64_t)(nn * hp + y + pad) * wp + (x + pad)) * cs + coff;
  }
  __host__ __device__ __forceinline__ int64_t pix_p(int nn, int yp, int xp) const {  // padded coords
    return ((int64_t)(nn * hp + yp) * wp + xp) * cs + coff;
  }
};

static inline TView make_view(const nemar_tensor* t) {
  TView v;
  v.ptr = (char*)t->ptr;
  v.n = t->n; v.h = t->h; v.w = t->w; v.c = t->c; v.pad = t->pad; v.cs = t->cs; v.coff = t->coff;
  v.dtype = t->dtype;
  v.hp = t->h + 2 * t->pad;
  v.wp = t->w + 2 * t->pad;
  return v;
}

static inline bool view_ok(const nemar_tensor* t) {
  return t && t->ptr && t->n > 0 && t->h > 0 && t->w > 0 && t->c > 0 && t->pad >= 0 &&
         t->cs >= t->coff + t->c && t->coff >= 0 && (t->dtype == NEMAR_F32 || t->dtype == NEMAR_BF16);
}
static inline bool same_shape(const nemar_tensor* a, const nemar_tensor* b) {
  return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c;
}

template <typename T> __device__ __forceinline__ float ld(const TView& v, int64_t idx) {
  return to_f<T>(((const T*)v.ptr)[idx]);
}
template <typename T> __device__ __forceinline__ void st(const TView& v, int64_t idx, float val) {
  ((T*)v.ptr)[idx] = from_f<T>(val);
}

// runtime-dtype scalar access (generic conv engine: mixed bf16 / fp32 operands)
__device__ __forceinline__ float ld_rt(const void* p, int dtype, int64_t i) {
  return dtype == NEMAR_F32 ? ((const float*)p)[i] : __bfloat162float(((const __nv_bfloat16*)p)[i]);
}
__device__ __forceinline__ void st_rt(void* p, int dtype, int64_t i, float v) {
  if (dtype == NEMAR_F32) ((float*)p)[i] = v;
  else ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
}

// reflect index for ReflectionPad: i in [-p, n+p) -> [0,n)
__host__ __device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// For interior coordinate s of a dimension of size n carrying a reflect halo p, enumerate the padded
// coordinates whose value is a copy of s (adjoint of reflect padding).  Returns count (1..3).
__device__ __forceinline__ int reflect_sources(int s, int n, int p, int out[3]) {
  int k = 0;
  out[k++] = s + p;
  if (p > 0) {
    if (s >= 1 && s <= p) out[k++] = p - s;
    if (s <= n - 2 && s >= n - 1 - p) out[k++] = p + 2 * (n - 1) - s;
  }
  return k;
}

__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act) {
    case NEMAR_ACT_RELU: return x > 0.f ? x : 0.f;
    case NEMAR_ACT_LRELU: return x > 0.f ? x : 0.2f * x;
    case NEMAR_ACT_TANH: return tanhf(x);
    default: return x;
  }
}
// derivative expressed through the pre-activation value x
__device__ __forceinline__ float act_grad_from_x(float x, int act) {
  switch (act) {
    case NEMAR_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case NEMAR_ACT_LRELU: return x > 0.f ? 1.f : 0.2f;
    case NEMAR_ACT_TANH: { float t = tanhf(x); return 1.f - t * t; }
    default: return 1.f;
  }
}
// derivative expressed through the activation OUTPUT y
__device__ __forceinline__ float act_grad_from_y(float y, int act) {
  switch (act) {
    case NEMAR_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case NEMAR_ACT_LRELU: return y > 0.f ? 1.f : 0.2f;
    case NEMAR_ACT_TANH: return 1.f - y * y;
    default: return 1.f;
  }
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in thread 0.  smem must hold >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* smem) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem[threadIdx.x] : 0.f;
  if (wid == 0) v = warp_sum(v);
  return v;
}

#define DISPATCH_DTYPE(dt, T, ...)                         \
  do {                                                     \
    if ((dt) == NEMAR_F32) { using T = float; __VA_ARGS__; } \
    else { using T = __nv_bfloat16; __VA_ARGS__; }          \
  } while (0)
