// norm_lean.cuh — occupancy-first bf16 kernels for the InstanceNorm passes (forward, statistics, backward-reduce,
// backward-apply).
//
// The generic templates (elementwise.cu) and the first fast path (norm_fast.cuh) keep every per-channel constant
// of a thread's 8 channels in registers: 104-158 registers per thread, i.e. 1-2 CTAs of 256 threads per SM, and a
// measured 1.1-2.3 TB/s on the 256-channel 64x64 maps of the ResnetBlocks (cuobjdump --dump-resource-usage;
// per-shape CUDA-event timings in DESIGN.md).  These passes are pure latency hiding: bytes in flight per SM =
// resident threads x loads in flight per thread.  Here the per-channel constants live in shared memory as fused
// multiply-add coefficients, a thread keeps U = 2 pixels (raw 16-byte loads) in flight, and __launch_bounds__ asks
// for 3-4 CTAs per SM.
//
//   xhat = A*x + B                 A = rstd, B = -mean*rstd            (identity when there are no statistics)
//   dx   = A*g' + C + xhat*D       C = -rstd*mean(g'), D = -rstd*mean(g'*xhat),  g' = fold(dy)*act'(xhat)
//
//   requirements: bf16 storage, 128-bit-accessible views, 256 % (C/8) == 0.
#pragma once
#include "common.cuh"
#include "norm_fast.cuh"

namespace nlean {

using nfast::ldraw;
using nfast::unpack8;
using nfast::pack8;
using nfast::near_border;
using nfast::fold_extra;
using nfast::Range;
using nfast::block_range;

constexpr int U = 2;   // pixels in flight per thread

// ACT >= 0: activation known at compile time (no per-element switch); ACT < 0: run-time value
template <int ACT> __device__ __forceinline__ float actf(float x, int act) { return act_fwd(x, ACT < 0 ? act : ACT); }
template <int ACT> __device__ __forceinline__ float actg(float x, int act) { return act_grad_from_x(x, ACT < 0 ? act : ACT); }

// p -> (p / w, p % w) through a float reciprocal plus one correction step (p < 2^23, exact)
__device__ __forceinline__ void divmod(uint32_t p, uint32_t w, float inv_w, int& q, int& rem) {
  uint32_t t = (uint32_t)__float2int_rz((float)p * inv_w);
  if (t * w > p) --t;
  else if ((t + 1) * w <= p) ++t;
  q = (int)t;
  rem = (int)(p - t * w);
}
// element offset of interior pixel (y, x) / padded pixel (yp, xp) inside ONE sample: fits 32 bits
__device__ __forceinline__ uint32_t off_in(const TView& t, int y, int x) {
  return ((uint32_t)(y + t.pad) * (uint32_t)t.wp + (uint32_t)(x + t.pad)) * (uint32_t)t.cs;
}
__device__ __forceinline__ uint32_t off_pad(const TView& t, int yp, int xp) {
  return ((uint32_t)yp * (uint32_t)t.wp + (uint32_t)xp) * (uint32_t)t.cs;
}
// base of sample nn, channel c0
__device__ __forceinline__ const __nv_bfloat16* sample_base(const TView& t, int nn, int c0) {
  return (const __nv_bfloat16*)t.ptr + (int64_t)nn * t.hp * t.wp * t.cs + t.coff + c0;
}

// per-channel (A, B) of sample nn into shared memory
__device__ __forceinline__ void fill_ab(const float* __restrict__ stats, int nn, int c, float inv_hw, float* sA, float* sB) {
  for (int k = threadIdx.x; k < c; k += blockDim.x) {
    float a = 1.f, b = 0.f;
    if (stats) {
      const float m = __ldg(stats + ((size_t)nn * c + k) * 2) * inv_hw;
      const float var = fmaxf(__ldg(stats + ((size_t)nn * c + k) * 2 + 1) * inv_hw - m * m, 0.f);
      a = rsqrtf(var + 1e-5f);
      b = -m * a;
    }
    sA[k] = a; sB[k] = b;
  }
}

// ---------------------------------------------------------------------------------------------
// forward: y = act(A*x + B) (+ residual); the halo of y is written in the same pass
// ---------------------------------------------------------------------------------------------
template <int ACT, int UU>
__global__ void __launch_bounds__(256, UU > 2 ? 3 : 4)
fwd_kernel(TView x, const float* __restrict__ stats, int act, TView res, int has_res, TView y, int pad_mode, float inv_hw) {
  constexpr int U = UU;
  extern __shared__ float sm[];
  const int nn = blockIdx.y, c = y.c, G = c / 8;
  float* sA = sm; float* sB = sm + c;
  fill_ab(stats, nn, c, inv_hw, sA, sB);
  __syncthreads();
  const Range r = block_range((uint32_t)y.hp * y.wp, G);
  const int c0 = r.cg * 8;
  const __nv_bfloat16* xb = sample_base(x, nn, c0);
  const __nv_bfloat16* rb = sample_base(res, nn, c0);
  __nv_bfloat16* yb = const_cast<__nv_bfloat16*>(sample_base(y, nn, c0));
  const uint32_t uwp = (uint32_t)y.wp;
  const float inv_wp = 1.f / (float)y.wp;
  for (uint32_t p0 = r.lo + r.pl; p0 < r.hi; p0 += U * r.ppi) {
    uint4 rx[U], rr[U];
    uint32_t oidx[U];
    int state[U];          // 0: skip, 1: compute, 2: zero halo
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint32_t p = p0 + k * r.ppi;
      state[k] = 0; oidx[k] = 0;
      rx[k] = make_uint4(0, 0, 0, 0); rr[k] = rx[k];
      if (p < r.hi) {
        int yp, xp;
        divmod(p, uwp, inv_wp, yp, xp);
        int ys = yp - y.pad, xs = xp - y.pad;
        const bool halo = ys < 0 || ys >= y.h || xs < 0 || xs >= y.w;
        oidx[k] = off_pad(y, yp, xp);
        if (halo && pad_mode != NEMAR_PAD_REFLECT) {
          state[k] = 2;
        } else {
          state[k] = 1;
          ys = reflect_idx(ys, y.h); xs = reflect_idx(xs, y.w);
          rx[k] = ldraw(xb + off_in(x, ys, xs));
          if (has_res) rr[k] = ldraw(rb + off_in(res, ys, xs));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (state[k] == 0) continue;
      float v[8];
      unpack8(rx[k], v);
      if (state[k] == 1) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 a = *reinterpret_cast<const float4*>(sA + c0 + 4 * h);
          const float4 b = *reinterpret_cast<const float4*>(sB + c0 + 4 * h);
          v[4 * h + 0] = actf<ACT>(fmaf(v[4 * h + 0], a.x, b.x), act);
          v[4 * h + 1] = actf<ACT>(fmaf(v[4 * h + 1], a.y, b.y), act);
          v[4 * h + 2] = actf<ACT>(fmaf(v[4 * h + 2], a.z, b.z), act);
          v[4 * h + 3] = actf<ACT>(fmaf(v[4 * h + 3], a.w, b.w), act);
        }
        if (has_res) {
          float q[8];
          unpack8(rr[k], q);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += q[j];
        }
      }
      *reinterpret_cast<uint4*>(yb + oidx[k]) = pack8(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// plane reductions.  MODE 0: (sum x, sum x^2).  MODE 1: (sum g', sum g'*xhat)
// ---------------------------------------------------------------------------------------------
template <int MODE, int UU, int ACT>
__global__ void __launch_bounds__(256, (MODE == 1 && UU > 2) ? 3 : 4)
reduce_kernel(TView x, const float* __restrict__ stats, int act, TView dy, int pad_mode, float inv_hw, float* __restrict__ out) {
  extern __shared__ float sm[];     // sacc[2c] (unused since the conflict-free combine) | A[c] | B[c] | partials[256/G][c][2]
  const int nn = blockIdx.y, c = x.c, G = c / 8;
  float* sacc = sm; float* sA = sm + 2 * c; float* sB = sm + 3 * c;
  for (int k = threadIdx.x; k < 2 * c; k += blockDim.x) sacc[k] = 0.f;
  if (MODE == 1) fill_ab(stats, nn, c, inv_hw, sA, sB);
  __syncthreads();
  const Range r = block_range((uint32_t)x.h * x.w, G);
  const int c0 = r.cg * 8;
  float a0[8], a1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a0[k] = 0.f; a1[k] = 0.f; }
  const __nv_bfloat16* xb = sample_base(x, nn, c0);
  const __nv_bfloat16* db = sample_base(dy, nn, c0);
  const bool fold = MODE == 1 && dy.pad > 0 && pad_mode == NEMAR_PAD_REFLECT;
  const uint32_t uw = (uint32_t)x.w;
  const float inv_w = 1.f / (float)x.w;
  for (uint32_t p0 = r.lo + r.pl; p0 < r.hi; p0 += UU * r.ppi) {
    uint4 rx[UU], rd[UU];
    int yy[UU], xx[UU];
#pragma unroll
    for (int k = 0; k < UU; ++k) {
      const uint32_t p = p0 + k * r.ppi;
      rx[k] = make_uint4(0, 0, 0, 0); rd[k] = rx[k]; yy[k] = -1; xx[k] = 0;
      if (p < r.hi) {
        divmod(p, uw, inv_w, yy[k], xx[k]);
        rx[k] = ldraw(xb + off_in(x, yy[k], xx[k]));
        if (MODE == 1) rd[k] = ldraw(db + off_in(dy, yy[k], xx[k]));
      }
    }
#pragma unroll
    for (int k = 0; k < UU; ++k) {
      if (yy[k] < 0) continue;
      float v[8];
      unpack8(rx[k], v);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] += v[j]; a1[j] = fmaf(v[j], v[j], a1[j]); }
      } else {
        float g[8];
        unpack8(rd[k], g);
        if (fold && near_border(yy[k], xx[k], dy.h, dy.w, dy.pad)) fold_extra(dy, nn, yy[k], xx[k], c0, g);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 a = *reinterpret_cast<const float4*>(sA + c0 + 4 * h);
          const float4 b = *reinterpret_cast<const float4*>(sB + c0 + 4 * h);
          const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xh = fmaf(v[4 * h + j], aa[j], bb[j]);
            const float gg = g[4 * h + j] * actg<ACT>(xh, act);
            a0[4 * h + j] += gg;
            a1[4 * h + j] = fmaf(gg, xh, a1[4 * h + j]);
          }
        }
      }
    }
  }
  // Block combine WITHOUT shared-memory atomics: the 256 / G threads that own the same channel group park their partial
  // sums in [pixel lane][c][2] order (16 consecutive floats per thread: conflict-free vector stores), then 2c threads
  // add the pixel lanes up.  (Round 1 used atomicAdd on sacc[(c0 + k) * 2]: the 32 lanes of a warp hit two banks —
  // 16-way conflicts on 16 atomics per thread, ~30 us of a 30 us statistics pass; found with scripts/nbench.py.)
  float* spart = sm + 4 * c;                  // [256 / G][c][2]
  {
    float4* d = reinterpret_cast<float4*>(spart + ((size_t)r.pl * c + c0) * 2);
    d[0] = make_float4(a0[0], a1[0], a0[1], a1[1]);
    d[1] = make_float4(a0[2], a1[2], a0[3], a1[3]);
    d[2] = make_float4(a0[4], a1[4], a0[5], a1[5]);
    d[3] = make_float4(a0[6], a1[6], a0[7], a1[7]);
  }
  __syncthreads();
  const int lanes = 256 / G;
  for (int k = threadIdx.x; k < 2 * c; k += blockDim.x) {
    float v = 0.f;
    for (int l = 0; l < lanes; ++l) v += spart[(size_t)l * 2 * c + k];
    atomicAdd(out + (size_t)nn * c * 2 + k, v);
  }
}

// ---------------------------------------------------------------------------------------------
// backward apply: dx = A*g' + C + xhat*D;  dres (+)= fold(dy);  db += column sums of dx
// ---------------------------------------------------------------------------------------------
template <int ACT, int UU>
__global__ void __launch_bounds__(256, UU > 2 ? 2 : 3)
bwd_apply_kernel(TView x, const float* __restrict__ stats, int act, TView dy, int pad_mode, const float* __restrict__ red,
                 TView dx, TView dres, int has_dres, int dres_acc, float inv_hw, float* __restrict__ dbias) {
  constexpr int U = UU;
  extern __shared__ float sm[];     // A[c] | B[c] | C[c] | D[c] | db[c] | partials[256/G][c]
  const int nn = blockIdx.y, c = x.c, G = c / 8;
  float* sA = sm; float* sB = sm + c; float* sC = sm + 2 * c; float* sD = sm + 3 * c; float* sdb = sm + 4 * c;
  fill_ab(stats, nn, c, inv_hw, sA, sB);
  for (int k = threadIdx.x; k < c; k += blockDim.x) {
    float cc = 0.f, dd = 0.f;
    if (stats) {
      const float a = sA[k];      // written by this same thread in fill_ab (same k -> thread mapping)
      cc = -a * __ldg(red + ((size_t)nn * c + k) * 2) * inv_hw;
      dd = -a * __ldg(red + ((size_t)nn * c + k) * 2 + 1) * inv_hw;
    }
    sC[k] = cc; sD[k] = dd; sdb[k] = 0.f;
  }
  __syncthreads();
  const Range r = block_range((uint32_t)x.h * x.w, G);
  const int c0 = r.cg * 8;
  float bsum[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) bsum[k] = 0.f;
  const __nv_bfloat16* xb = sample_base(x, nn, c0);
  const __nv_bfloat16* db = sample_base(dy, nn, c0);
  __nv_bfloat16* ob = const_cast<__nv_bfloat16*>(sample_base(dx, nn, c0));
  __nv_bfloat16* rb = const_cast<__nv_bfloat16*>(sample_base(dres, nn, c0));
  const bool fold = dy.pad > 0 && pad_mode == NEMAR_PAD_REFLECT;
  const bool racc = has_dres && dres_acc;
  const uint32_t uw = (uint32_t)x.w;
  const float inv_w = 1.f / (float)x.w;
  for (uint32_t p0 = r.lo + r.pl; p0 < r.hi; p0 += U * r.ppi) {
    uint4 rx[U], rd4[U], rold[U];
    int yy[U], xx[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint32_t p = p0 + k * r.ppi;
      rx[k] = make_uint4(0, 0, 0, 0); rd4[k] = rx[k]; rold[k] = rx[k]; yy[k] = -1; xx[k] = 0;
      if (p < r.hi) {
        divmod(p, uw, inv_w, yy[k], xx[k]);
        rx[k] = ldraw(xb + off_in(x, yy[k], xx[k]));
        rd4[k] = ldraw(db + off_in(dy, yy[k], xx[k]));
        if (racc) rold[k] = ldraw(rb + off_in(dres, yy[k], xx[k]));
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (yy[k] < 0) continue;
      float g[8], v[8];
      unpack8(rd4[k], g);
      if (fold && near_border(yy[k], xx[k], dy.h, dy.w, dy.pad)) fold_extra(dy, nn, yy[k], xx[k], c0, g);
      if (has_dres) {
        float t[8];
        unpack8(rold[k], t);     // zeros unless accumulating
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] += g[j];
        *reinterpret_cast<uint4*>(rb + off_in(dres, yy[k], xx[k])) = pack8(t);
      }
      unpack8(rx[k], v);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 a = *reinterpret_cast<const float4*>(sA + c0 + 4 * h);
        const float4 b = *reinterpret_cast<const float4*>(sB + c0 + 4 * h);
        const float4 cc = *reinterpret_cast<const float4*>(sC + c0 + 4 * h);
        const float4 dd = *reinterpret_cast<const float4*>(sD + c0 + 4 * h);
        const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
        const float c4[4] = {cc.x, cc.y, cc.z, cc.w}, d4[4] = {dd.x, dd.y, dd.z, dd.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xh = fmaf(v[4 * h + j], aa[j], bb[j]);
          const float gg = g[4 * h + j] * actg<ACT>(xh, act);
          const float o = fmaf(gg, aa[j], fmaf(xh, d4[j], c4[j]));
          v[4 * h + j] = o;
          bsum[4 * h + j] += o;
        }
      }
      *reinterpret_cast<uint4*>(ob + off_in(dx, yy[k], xx[k])) = pack8(v);
    }
  }
  if (dbias) {
    // same conflict-free block combine as the reductions: [pixel lane][c] partials, then c threads add the lanes up
    float* spart = sm + 5 * c;                // [256 / G][c]
    float4* d = reinterpret_cast<float4*>(spart + (size_t)r.pl * c + c0);
    d[0] = make_float4(bsum[0], bsum[1], bsum[2], bsum[3]);
    d[1] = make_float4(bsum[4], bsum[5], bsum[6], bsum[7]);
    __syncthreads();
    const int lanes = 256 / G;
    for (int k = threadIdx.x; k < c; k += blockDim.x) {
      float v = 0.f;
      for (int l = 0; l < lanes; ++l) v += spart[(size_t)l * c + k];
      atomicAdd(dbias + k, v);
    }
  }
}

// blocks per sample: ~8 CTAs per SM over the whole batch, each with at least two iterations' worth of pixels
static inline int reduce_u() {
  static const int u = [] { const char* e = getenv("NEMAR_LEAN_RED_U"); return e ? atoi(e) : 2; }();
  return u;
}
static inline int stream_u() {     // pixels in flight per thread in the forward / backward-apply passes (2 or 4)
  static const int u = [] { const char* e = getenv("NEMAR_LEAN_U"); return e ? atoi(e) : 2; }();
  return u;
}
// the reductions pay a fixed latency chain per CTA (zero the shared accumulators, coefficient table, shared then
// global atomics), so they want FEWER, longer CTAs than the streaming passes (measured: 4/SM beats 8/SM beats 16/SM)
static inline int chunks_for(int64_t npix, int G, int n, bool reduction = false, bool apply = false) {
  static const int per_sm_f = [] { const char* e = getenv("NEMAR_LEAN_CTAS_PER_SM"); return e ? atoi(e) : 8; }();
  static const int per_sm_a = [] { const char* e = getenv("NEMAR_LEAN_APPLY_PER_SM"); return e ? atoi(e) : 6; }();
  const int per_sm_s = apply ? per_sm_a : per_sm_f;
  static const int per_sm_r = [] { const char* e = getenv("NEMAR_LEAN_RED_PER_SM"); return e ? atoi(e) : 4; }();
  const int per_sm = reduction ? per_sm_r : per_sm_s;
  const int64_t ppi = 256 / G;
  int64_t chunks = (npix + ppi * U * 2 - 1) / (ppi * U * 2);
  int64_t cap = ((int64_t)148 * per_sm + n - 1) / n;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  return (int)chunks;
}

// instantiate `...` with the compile-time activation A (A = -1: run-time switch, e.g. tanh)
#define NLEAN_ACT_SWITCH(act, ...)                                                  \
  switch (act) {                                                                    \
    case NEMAR_ACT_NONE: { constexpr int A = NEMAR_ACT_NONE; __VA_ARGS__; } break;   \
    case NEMAR_ACT_RELU: { constexpr int A = NEMAR_ACT_RELU; __VA_ARGS__; } break;   \
    case NEMAR_ACT_LRELU: { constexpr int A = NEMAR_ACT_LRELU; __VA_ARGS__; } break; \
    default: { constexpr int A = -1; __VA_ARGS__; } break;                           \
  }

// bit 0: forward, bit 1: statistics, bit 2: backward reduce, bit 3: backward apply
static inline int enabled_mask() {
  static const int m = [] { const char* e = getenv("NEMAR_NORM_LEAN"); return e ? atoi(e) : 15; }();
  return m;
}

}  // namespace nlean
