// norm_fast.cuh — bf16 fast paths of the three InstanceNorm passes (forward, backward-reduce, backward-apply).
//
// ncu on the generic templates (profiles/): 12-22 % warps active, long-scoreboard stalls, ~2.5 TB/s — the loop
// body interleaves loads and stores through possibly-aliasing pointers, so only one item's loads are in flight per
// thread.  These kernels fix a thread's channel group (8 bf16 = 16 B), keep its per-channel constants in registers,
// and issue the raw 16-byte loads of U consecutive pixels BEFORE any arithmetic or store.
//   requirements: bf16 storage, 128-bit-accessible views, 256 % (C/8) == 0.
#pragma once
#include "common.cuh"

namespace nfast {

constexpr int U = 4;   // pixels in flight per thread

__device__ __forceinline__ uint4 ldraw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// does the interior pixel (y,x) of a buffer with reflect halo `p` have mirror images in the halo?
__device__ __forceinline__ bool near_border(int y, int x, int h, int w, int p) {
  return !(y > p && y < h - 1 - p && x > p && x < w - 1 - p);
}

// gradient of the padded buffer folded onto interior pixel (y,x): main tap given, mirrored taps added when needed
__device__ __forceinline__ void fold_extra(const TView& d, int nn, int y, int x, int c0, float (&g)[8]) {
  int ys[3], xs[3];
  const int ny = reflect_sources(y, d.h, d.pad, ys);
  const int nx = reflect_sources(x, d.w, d.pad, xs);
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) {
      if (a == 0 && b == 0) continue;   // (ys[0], xs[0]) is the main tap, already loaded
      float t[8];
      unpack8(ldraw((const __nv_bfloat16*)d.ptr + d.pix_p(nn, ys[a], xs[b]) + c0), t);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] += t[k];
    }
}

struct Range { uint32_t lo, hi, ppi; int cg, pl; };
__device__ __forceinline__ Range block_range(uint32_t npix, int G) {
  Range r;
  r.ppi = 256u / (uint32_t)G;                        // pixels covered by the block per iteration
  r.cg = (int)(threadIdx.x % G);
  r.pl = (int)(threadIdx.x / G);
  uint32_t per = (npix + gridDim.x - 1) / gridDim.x;
  per = (per + r.ppi - 1) / r.ppi * r.ppi;
  r.lo = blockIdx.x * per;
  r.hi = r.lo + per;
  if (r.hi > npix) r.hi = npix;
  if (r.lo > npix) r.lo = npix;
  return r;
}

// ---------------------------------------------------------------------------------------------
// forward: y = act(norm(x)) (+ residual), halo of y written in the same pass
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fwd_kernel(TView x, const float* __restrict__ stats, int act, TView res, int has_res, TView y, int pad_mode, float inv_hw) {
  const int nn = blockIdx.y, G = y.c / 8;
  const Range r = block_range((uint32_t)y.hp * y.wp, G);
  const int c0 = r.cg * 8;
  float mean[8], rstd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { mean[k] = 0.f; rstd[k] = 1.f; }
  if (stats) {
    const float* s = stats + ((size_t)nn * y.c + c0) * 2;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float m = __ldg(s + 2 * k) * inv_hw;
      const float var = fmaxf(__ldg(s + 2 * k + 1) * inv_hw - m * m, 0.f);
      mean[k] = m; rstd[k] = rsqrtf(var + 1e-5f);
    }
  }
  const __nv_bfloat16* xb = (const __nv_bfloat16*)x.ptr;
  const __nv_bfloat16* rb = (const __nv_bfloat16*)res.ptr;
  __nv_bfloat16* yb = (__nv_bfloat16*)y.ptr;
  const uint32_t uwp = (uint32_t)y.wp;
  for (uint32_t p0 = r.lo + r.pl; p0 < r.hi; p0 += U * r.ppi) {
    uint4 rx[U], rr[U];
    int64_t oidx[U];
    bool ok[U], zero[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint32_t p = p0 + k * r.ppi;
      ok[k] = p < r.hi;
      zero[k] = false;
      rx[k] = make_uint4(0, 0, 0, 0); rr[k] = rx[k]; oidx[k] = 0;
      if (ok[k]) {
        const int yp = (int)(p / uwp), xp = (int)(p - (uint32_t)yp * uwp);
        int ys = yp - y.pad, xs = xp - y.pad;
        const bool halo = ys < 0 || ys >= y.h || xs < 0 || xs >= y.w;
        oidx[k] = y.pix_p(nn, yp, xp) + c0;
        if (halo && pad_mode != NEMAR_PAD_REFLECT) {
          zero[k] = true;
        } else {
          ys = reflect_idx(ys, y.h); xs = reflect_idx(xs, y.w);
          rx[k] = ldraw(xb + x.pix(nn, ys, xs) + c0);
          if (has_res) rr[k] = ldraw(rb + res.pix(nn, ys, xs) + c0);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (!ok[k]) continue;
      float v[8];
      unpack8(rx[k], v);
      if (!zero[k]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_fwd((v[j] - mean[j]) * rstd[j], act);
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
// plane reductions.  MODE 0: (sum x, sum x^2).  MODE 1: (sum g, sum g*xhat), g = fold(dy) * act'(xhat)
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256)
reduce_kernel(TView x, const float* __restrict__ stats, int act, TView dy, int pad_mode, float inv_hw, float* __restrict__ out) {
  extern __shared__ float sacc[];   // [c][2]
  const int nn = blockIdx.y, c = x.c, G = c / 8;
  for (int k = threadIdx.x; k < 2 * c; k += blockDim.x) sacc[k] = 0.f;
  __syncthreads();
  const Range r = block_range((uint32_t)x.h * x.w, G);
  const int c0 = r.cg * 8;
  float mean[8], rstd[8], a0[8], a1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { mean[k] = 0.f; rstd[k] = 1.f; a0[k] = 0.f; a1[k] = 0.f; }
  if (MODE == 1) {
    const float* s = stats + ((size_t)nn * c + c0) * 2;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float m = __ldg(s + 2 * k) * inv_hw;
      const float var = fmaxf(__ldg(s + 2 * k + 1) * inv_hw - m * m, 0.f);
      mean[k] = m; rstd[k] = rsqrtf(var + 1e-5f);
    }
  }
  const __nv_bfloat16* xb = (const __nv_bfloat16*)x.ptr;
  const __nv_bfloat16* db = (const __nv_bfloat16*)dy.ptr;
  const bool fold = MODE == 1 && dy.pad > 0 && pad_mode == NEMAR_PAD_REFLECT;
  const uint32_t uw = (uint32_t)x.w;
  for (uint32_t p0 = r.lo + r.pl; p0 < r.hi; p0 += U * r.ppi) {
    uint4 rx[U], rd[U];
    int yy[U], xx[U];
    bool ok[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint32_t p = p0 + k * r.ppi;
      ok[k] = p < r.hi;
      rx[k] = make_uint4(0, 0, 0, 0); rd[k] = rx[k]; yy[k] = 0; xx[k] = 0;
      if (ok[k]) {
        yy[k] = (int)(p / uw); xx[k] = (int)(p - (uint32_t)yy[k] * uw);
        rx[k] = ldraw(xb + x.pix(nn, yy[k], xx[k]) + c0);
        if (MODE == 1) rd[k] = ldraw(db + dy.pix(nn, yy[k], xx[k]) + c0);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (!ok[k]) continue;
      float v[8];
      unpack8(rx[k], v);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] += v[j]; a1[j] += v[j] * v[j]; }
      } else {
        float g[8];
        unpack8(rd[k], g);
        if (fold && near_border(yy[k], xx[k], dy.h, dy.w, dy.pad)) fold_extra(dy, nn, yy[k], xx[k], c0, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (v[j] - mean[j]) * rstd[j];
          const float gg = g[j] * act_grad_from_x(xh, act);
          a0[j] += gg; a1[j] += gg * xh;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    atomicAdd(&sacc[(c0 + k) * 2], a0[k]);
    atomicAdd(&sacc[(c0 + k) * 2 + 1], a1[k]);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 2 * c; k += blockDim.x) atomicAdd(out + (size_t)nn * c * 2 + k, sacc[k]);
}

// ---------------------------------------------------------------------------------------------
// backward apply: dx = rstd*(g - mean(g) - xhat*mean(g*xhat))  (or g*act'(x) without statistics);
// dres = fold(dy); db += column sums of dx
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bwd_apply_kernel(TView x, const float* __restrict__ stats, int act, TView dy, int pad_mode, const float* __restrict__ red,
                 TView dx, TView dres, int has_dres, int dres_acc, float inv_hw, float* __restrict__ dbias) {
  extern __shared__ float sdb[];   // [c]
  const int nn = blockIdx.y, c = x.c, G = c / 8;
  if (dbias) {
    for (int k = threadIdx.x; k < c; k += blockDim.x) sdb[k] = 0.f;
    __syncthreads();
  }
  const Range r = block_range((uint32_t)x.h * x.w, G);
  const int c0 = r.cg * 8;
  float mean[8], rstd[8], m1[8], m2[8], bsum[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { mean[k] = 0.f; rstd[k] = 1.f; m1[k] = 0.f; m2[k] = 0.f; bsum[k] = 0.f; }
  if (stats) {
    const float* s = stats + ((size_t)nn * c + c0) * 2;
    const float* rd = red + ((size_t)nn * c + c0) * 2;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float m = __ldg(s + 2 * k) * inv_hw;
      const float var = fmaxf(__ldg(s + 2 * k + 1) * inv_hw - m * m, 0.f);
      mean[k] = m; rstd[k] = rsqrtf(var + 1e-5f);
      m1[k] = __ldg(rd + 2 * k) * inv_hw; m2[k] = __ldg(rd + 2 * k + 1) * inv_hw;
    }
  }
  const __nv_bfloat16* xb = (const __nv_bfloat16*)x.ptr;
  const __nv_bfloat16* db = (const __nv_bfloat16*)dy.ptr;
  __nv_bfloat16* ob = (__nv_bfloat16*)dx.ptr;
  __nv_bfloat16* rb = (__nv_bfloat16*)dres.ptr;
  const bool fold = dy.pad > 0 && pad_mode == NEMAR_PAD_REFLECT;
  const uint32_t uw = (uint32_t)x.w;
  for (uint32_t p0 = r.lo + r.pl; p0 < r.hi; p0 += U * r.ppi) {
    uint4 rx[U], rd4[U], rold[U];
    int yy[U], xx[U];
    bool ok[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint32_t p = p0 + k * r.ppi;
      ok[k] = p < r.hi;
      rx[k] = make_uint4(0, 0, 0, 0); rd4[k] = rx[k]; rold[k] = rx[k]; yy[k] = 0; xx[k] = 0;
      if (ok[k]) {
        yy[k] = (int)(p / uw); xx[k] = (int)(p - (uint32_t)yy[k] * uw);
        rx[k] = ldraw(xb + x.pix(nn, yy[k], xx[k]) + c0);
        rd4[k] = ldraw(db + dy.pix(nn, yy[k], xx[k]) + c0);
        if (has_dres && dres_acc) rold[k] = ldraw(rb + dres.pix(nn, yy[k], xx[k]) + c0);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (!ok[k]) continue;
      float g[8], v[8], o[8];
      unpack8(rd4[k], g);
      if (fold && near_border(yy[k], xx[k], dy.h, dy.w, dy.pad)) fold_extra(dy, nn, yy[k], xx[k], c0, g);
      if (has_dres) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = g[j];
        if (dres_acc) {
          float old[8];
          unpack8(rold[k], old);
#pragma unroll
          for (int j = 0; j < 8; ++j) t[j] += old[j];
        }
        *reinterpret_cast<uint4*>(rb + dres.pix(nn, yy[k], xx[k]) + c0) = pack8(t);
      }
      unpack8(rx[k], v);
      if (stats) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (v[j] - mean[j]) * rstd[j];
          const float gg = g[j] * act_grad_from_x(xh, act);
          o[j] = rstd[j] * (gg - m1[j] - xh * m2[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = g[j] * act_grad_from_x(v[j], act);
      }
      *reinterpret_cast<uint4*>(ob + dx.pix(nn, yy[k], xx[k]) + c0) = pack8(o);
#pragma unroll
      for (int j = 0; j < 8; ++j) bsum[j] += o[j];
    }
  }
  if (dbias) {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&sdb[c0 + k], bsum[k]);
    __syncthreads();
    for (int k = threadIdx.x; k < c; k += blockDim.x) atomicAdd(dbias + k, sdb[k]);
  }
}

// host-side eligibility: bf16, vector-accessible, channel groups divide the block
static inline bool eligible(const nemar_tensor* t) {
  return t->dtype == NEMAR_BF16 && t->c % 8 == 0 && t->cs % 8 == 0 && t->coff % 8 == 0 && ((((uintptr_t)t->ptr) & 15) == 0) &&
         (256 % (t->c / 8)) == 0;
}
// blocks per sample: enough CTAs to fill the machine, each with >= U iterations' worth of pixels
static inline int chunks_for(int64_t npix, int G, int n) {
  const int64_t ppi = 256 / G;
  int64_t chunks = (npix + ppi * U * 2 - 1) / (ppi * U * 2);
  int64_t cap = (148 * 6 + n - 1) / n;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  return (int)chunks;
}

}  // namespace nfast
