// elementwise.cu — the HBM-bound glue of the hot path on padded-NHWC views: layout crossings,
// InstanceNorm(+activation +residual +reflect halo) forward/backward, activation backward, MaxPool2d(2),
// bilinear resize, bias gradient.  One pass per tensor, 128-bit channel vectors when the view allows.
//
// Reference call sites: models/networks.py:24,349-352,375,418-445,584-585; models/stn/layers.py:16,99-105,174;
// models/stn/unet_stn.py:96,166,188-195; models/nemar_model.py:187-188.
#include "common.cuh"
#include "vec.cuh"
#include "norm_lean.cuh"
#include <cstdlib>

// =============================================================================================
// helpers
// =============================================================================================
template <typename T, int V>
__device__ __forceinline__ void load_fold(const TView& d, int nn, int y, int x, int c0, int pad_mode,
                                          float (&g)[V]) {
  const T* base = (const T*)d.ptr;
  // fast path: no halo, or a pixel whose mirror images do not exist (more than `pad` away from every border)
  if (d.pad == 0 || pad_mode != NEMAR_PAD_REFLECT ||
      (y > d.pad && y < d.h - 1 - d.pad && x > d.pad && x < d.w - 1 - d.pad)) {
    ldv<T, V>(base + d.pix(nn, y, x) + c0, g);
    return;
  }
  int ys[3], xs[3];
  const int ny = reflect_sources(y, d.h, d.pad, ys);
  const int nx = reflect_sources(x, d.w, d.pad, xs);
#pragma unroll
  for (int k = 0; k < V; ++k) g[k] = 0.f;
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) {
      float t[V];
      ldv<T, V>(base + d.pix_p(nn, ys[a], xs[b]) + c0, t);
#pragma unroll
      for (int k = 0; k < V; ++k) g[k] += t[k];
    }
}

template <int V>
__device__ __forceinline__ void load_mean_rstd(const float* stats, int nn, int c, int c0, float inv_hw,
                                               float (&mean)[V], float (&rstd)[V]) {
  const float* s = stats + ((int64_t)nn * c + c0) * 2;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    float sm = __ldg(s + 2 * k), sq = __ldg(s + 2 * k + 1);
    float m = sm * inv_hw;
    float var = fmaxf(sq * inv_hw - m * m, 0.f);
    mean[k] = m;
    rstd[k] = rsqrtf(var + 1e-5f);
  }
}

// =============================================================================================
// NCHW fp32  <->  padded NHWC
// =============================================================================================
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, TView d, int pad_mode) {
  const int64_t total = (int64_t)d.n * d.hp * d.wp;
  const int64_t hw = (int64_t)d.h * d.w;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int xp = (int)(i % d.wp);
    int64_t r = i / d.wp;
    int yp = (int)(r % d.hp);
    int nn = (int)(r / d.hp);
    int y = yp - d.pad, x = xp - d.pad;
    bool halo = y < 0 || y >= d.h || x < 0 || x >= d.w;
    T* o = (T*)d.ptr + d.pix_p(nn, yp, xp);
    if (halo && pad_mode != NEMAR_PAD_REFLECT) {
      for (int ch = 0; ch < d.c; ++ch) o[ch] = from_f<T>(0.f);
      continue;
    }
    y = reflect_idx(y, d.h);
    x = reflect_idx(x, d.w);
    const float* s = src + (int64_t)nn * d.c * hw + (int64_t)y * d.w + x;
    for (int ch = 0; ch < d.c; ++ch) o[ch] = from_f<T>(__ldg(s + ch * hw));
  }
}

NEMAR_API int nemar_nchw_to_nhwc(const float* src, const nemar_tensor* dst, int pad_mode, void* stream) {
  NEMAR_REQUIRE(src && view_ok(dst), "nchw_to_nhwc: bad args");
  TView d = make_view(dst);
  int64_t total = (int64_t)d.n * d.hp * d.wp;
  DISPATCH_DTYPE(d.dtype, T, (nchw_to_nhwc_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
                                 src, d, pad_mode)));
  NEMAR_LAUNCH_CHECK();
  return 0;
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(TView s, float* __restrict__ dst, int pad_mode, int accumulate) {
  const int64_t hw = (int64_t)s.h * s.w;
  const int64_t total = (int64_t)s.n * hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % s.w);
    int64_t r = i / s.w;
    int y = (int)(r % s.h);
    int nn = (int)(r / s.h);
    float* o = dst + (int64_t)nn * s.c * hw + (int64_t)y * s.w + x;
    for (int ch = 0; ch < s.c; ++ch) {
      float g[1];
      load_fold<T, 1>(s, nn, y, x, ch, pad_mode, g);
      if (accumulate) o[ch * hw] += g[0];
      else o[ch * hw] = g[0];
    }
  }
}

NEMAR_API int nemar_nhwc_to_nchw(const nemar_tensor* src, float* dst, int pad_mode, int accumulate,
                                 void* stream) {
  NEMAR_REQUIRE(dst && view_ok(src), "nhwc_to_nchw: bad args");
  TView s = make_view(src);
  int64_t total = (int64_t)s.n * s.h * s.w;
  DISPATCH_DTYPE(s.dtype, T, (nhwc_to_nchw_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
                                 s, dst, pad_mode, accumulate)));
  NEMAR_LAUNCH_CHECK();
  return 0;
}

template <typename T> __global__ void fill_channels_kernel(TView t, int c0, int nc) {
  const int64_t total = (int64_t)t.n * t.hp * t.wp * nc;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int ch = (int)(i % nc);
    int64_t p = i / nc;
    ((T*)t.ptr)[p * t.cs + c0 + ch] = from_f<T>(0.f);
  }
}

NEMAR_API int nemar_fill_channels(const nemar_tensor* t, int c0, int nc, void* stream) {
  NEMAR_REQUIRE(t && t->ptr && nc > 0 && c0 >= 0 && c0 + nc <= t->cs, "fill_channels: bad args");
  TView v = make_view(t);
  int64_t total = (int64_t)v.n * v.hp * v.wp * nc;
  DISPATCH_DTYPE(v.dtype, T,
                 (fill_channels_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(v, c0, nc)));
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// view copy (channel-slice "cat" writes) and its adjoint
// =============================================================================================
template <typename T, int V> __global__ void copy_view_kernel(TView s, TView d, int pad_mode) {
  const int G = d.c / V;
  const int64_t total = (int64_t)d.n * d.hp * d.wp * G;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int cg = (int)(i % G);
    int64_t r = i / G;
    int xp = (int)(r % d.wp); r /= d.wp;
    int yp = (int)(r % d.hp);
    int nn = (int)(r / d.hp);
    int y = yp - d.pad, x = xp - d.pad;
    bool halo = y < 0 || y >= d.h || x < 0 || x >= d.w;
    float v[V];
    if (halo && pad_mode != NEMAR_PAD_REFLECT) {
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = 0.f;
    } else {
      y = reflect_idx(y, d.h);
      x = reflect_idx(x, d.w);
      ldv<T, V>((const T*)s.ptr + s.pix(nn, y, x) + cg * V, v);
    }
    stv<T, V>((T*)d.ptr + d.pix_p(nn, yp, xp) + cg * V, v);
  }
}

NEMAR_API int nemar_copy_view(const nemar_tensor* src, const nemar_tensor* dst, int pad_mode, void* stream) {
  NEMAR_REQUIRE(view_ok(src) && view_ok(dst) && same_shape(src, dst) && src->dtype == dst->dtype,
                "copy_view: bad args");
  TView s = make_view(src), d = make_view(dst);
  DISPATCH_DTYPE(d.dtype, T, {
    constexpr int V = VecTraits<T>::V;
    if (view_vec_ok<T>(src) && view_vec_ok<T>(dst)) {
      int64_t total = (int64_t)d.n * d.hp * d.wp * (d.c / V);
      copy_view_kernel<T, V><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(s, d, pad_mode);
    } else {
      int64_t total = (int64_t)d.n * d.hp * d.wp * d.c;
      copy_view_kernel<T, 1><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(s, d, pad_mode);
    }
  });
  NEMAR_LAUNCH_CHECK();
  return 0;
}

__global__ void cast_view_kernel(TView s, TView d) {
  const int64_t total = (int64_t)d.n * d.h * d.w * d.c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int ch = (int)(i % d.c);
    int64_t r = i / d.c;
    int x = (int)(r % d.w); r /= d.w;
    int y = (int)(r % d.h);
    int nn = (int)(r / d.h);
    st_rt(d.ptr, d.dtype, d.pix(nn, y, x) + ch, ld_rt(s.ptr, s.dtype, s.pix(nn, y, x) + ch));
  }
}

NEMAR_API int nemar_cast_view(const nemar_tensor* src, const nemar_tensor* dst, void* stream) {
  NEMAR_REQUIRE(view_ok(src) && view_ok(dst) && same_shape(src, dst), "cast_view: bad args");
  TView s = make_view(src), d = make_view(dst);
  int64_t total = (int64_t)d.n * d.h * d.w * d.c;
  cast_view_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(s, d);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// ---- tap <-> channel transforms (k7 head / tail as 1x1 tensor-core convolutions) ----------------
template <int KY, int KK, int CC>
__global__ void gather_taps_kernel(TView s, TView d, int ky_rt, int k_rt, int c_rt, int sgn) {
  const int k = KK > 0 ? KK : k_rt, c = CC > 0 ? CC : c_rt, ky = KY > 0 ? KY : ky_rt;   // ky x k taps
  // one thread -> 8 consecutive destination channels of one pixel (one 16-byte store when d is bf16)
  const uint32_t groups = d.c / 8;
  const int kc = ky * k * c;
  const uint32_t total = (uint32_t)d.n * d.h * d.w * groups;      // < 2^32 for every size on the path
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    uint32_t r = i / groups;
    const int g = (int)(i - r * groups);
    const int x = (int)(r % (uint32_t)d.w); r /= (uint32_t)d.w;
    const int y = (int)(r % (uint32_t)d.h);
    const int nn = (int)(r / (uint32_t)d.h);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int oc = g * 8 + j;
      float val = 0.f;
      if (oc < kc) {
        int tap = oc / c, ch = oc - tap * c;
        int a = tap / k, b = tap - a * k;
        int sy = y + sgn * a, sx = x + sgn * b;
        if (sy >= 0 && sy < s.h && sx >= 0 && sx < s.w) val = ld_rt(s.ptr, s.dtype, s.pix(nn, sy, sx) + ch);
      }
      v[j] = val;
    }
    const int64_t o = d.pix(nn, y, x) + g * 8;
    if (d.dtype == NEMAR_BF16) {
      uint4 pk;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      *reinterpret_cast<uint4*>((__nv_bfloat16*)d.ptr + o) = pk;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) ((float*)d.ptr)[o + j] = v[j];
    }
  }
}

template <int KY, int KK, int CC>
__global__ void sum_taps_kernel(TView s, TView d, int ky_rt, int k_rt, int c_rt, int sgn, const float* __restrict__ bias, int act) {
  const int k = KK > 0 ? KK : k_rt, c = CC > 0 ? CC : c_rt, ky = KY > 0 ? KY : ky_rt;
  const int64_t total = (int64_t)d.n * d.h * d.w;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % d.w);
    int64_t r = i / d.w;
    int y = (int)(r % d.h);
    int nn = (int)(r / d.h);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};     // c <= 4
    for (int a = 0; a < ky; ++a) {
      int sy = y - sgn * a;
      if (sy < 0 || sy >= s.h) continue;
      for (int b = 0; b < k; ++b) {
        int sx = x - sgn * b;
        if (sx < 0 || sx >= s.w) continue;
        const int64_t base = s.pix(nn, sy, sx) + (a * k + b) * c;
        for (int ch = 0; ch < c; ++ch) acc[ch] += ld_rt(s.ptr, s.dtype, base + ch);
      }
    }
    const int64_t o = d.pix(nn, y, x);
    for (int ch = 0; ch < d.c; ++ch) {
      float v = 0.f;
      if (ch < c) v = act_fwd(acc[ch] + (bias ? __ldg(bias + ch) : 0.f), act);
      st_rt(d.ptr, d.dtype, o + ch, v);
    }
  }
}

NEMAR_API int nemar_gather_taps2(const nemar_tensor* src, const nemar_tensor* dst, int ky, int kx, int c, int sgn, void* stream) {
  NEMAR_REQUIRE(view_ok(src) && view_ok(dst) && src->pad == 0 && dst->pad == 0 && src->n == dst->n, "gather_taps: bad views");
  NEMAR_REQUIRE(ky > 0 && kx > 0 && c > 0 && c <= src->c && dst->c % 8 == 0 && dst->c >= ky * kx * c && (sgn == 1 || sgn == -1) &&
                    (dst->dtype != NEMAR_BF16 || ((dst->cs % 8 == 0) && (dst->coff % 8 == 0) && ((((uintptr_t)dst->ptr) & 15) == 0))),
                "gather_taps: bad arguments");
  TView s = make_view(src), d = make_view(dst);
  int64_t total = (int64_t)d.n * d.h * d.w * (d.c / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (ky == 7 && kx == 7 && c == 3) gather_taps_kernel<7, 7, 3><<<grid_for(total, 256), 256, 0, st>>>(s, d, ky, kx, c, sgn);
  else if (ky == 1 && kx == 7 && c == 3) gather_taps_kernel<1, 7, 3><<<grid_for(total, 256), 256, 0, st>>>(s, d, ky, kx, c, sgn);
  else gather_taps_kernel<0, 0, 0><<<grid_for(total, 256), 256, 0, st>>>(s, d, ky, kx, c, sgn);
  NEMAR_LAUNCH_CHECK();
  return 0;
}
NEMAR_API int nemar_gather_taps(const nemar_tensor* src, const nemar_tensor* dst, int k, int c, int sgn, void* stream) {
  return nemar_gather_taps2(src, dst, k, k, c, sgn, stream);
}

NEMAR_API int nemar_sum_taps2(const nemar_tensor* src, const nemar_tensor* dst, int ky, int kx, int c, int sgn, const float* bias,
                              int act, void* stream) {
  NEMAR_REQUIRE(view_ok(src) && view_ok(dst) && src->pad == 0 && dst->pad == 0 && src->n == dst->n, "sum_taps: bad views");
  NEMAR_REQUIRE(ky > 0 && kx > 0 && c > 0 && c <= 4 && c <= dst->c && src->c >= ky * kx * c && (sgn == 1 || sgn == -1), "sum_taps: bad arguments");
  TView s = make_view(src), d = make_view(dst);
  int64_t total = (int64_t)d.n * d.h * d.w;
  cudaStream_t st = (cudaStream_t)stream;
  if (ky == 7 && kx == 7 && c == 3) sum_taps_kernel<7, 7, 3><<<grid_for(total, 128), 128, 0, st>>>(s, d, ky, kx, c, sgn, bias, act);
  else if (ky == 1 && kx == 7 && c == 3) sum_taps_kernel<1, 7, 3><<<grid_for(total, 128), 128, 0, st>>>(s, d, ky, kx, c, sgn, bias, act);
  else sum_taps_kernel<0, 0, 0><<<grid_for(total, 128), 128, 0, st>>>(s, d, ky, kx, c, sgn, bias, act);
  NEMAR_LAUNCH_CHECK();
  return 0;
}
NEMAR_API int nemar_sum_taps(const nemar_tensor* src, const nemar_tensor* dst, int k, int c, int sgn, const float* bias,
                             int act, void* stream) {
  return nemar_sum_taps2(src, dst, k, k, c, sgn, bias, act, stream);
}

template <typename T, int V>
__global__ void copy_view_bwd_kernel(TView ds, TView dd, int pad_mode, int accumulate) {
  // ds: gradient wrt the copy source (written), dd: gradient wrt the (padded) destination (read+fold)
  const int G = ds.c / V;
  const int64_t total = (int64_t)ds.n * ds.h * ds.w * G;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int cg = (int)(i % G);
    int64_t r = i / G;
    int x = (int)(r % ds.w); r /= ds.w;
    int y = (int)(r % ds.h);
    int nn = (int)(r / ds.h);
    float g[V];
    load_fold<T, V>(dd, nn, y, x, cg * V, pad_mode, g);
    T* o = (T*)ds.ptr + ds.pix(nn, y, x) + cg * V;
    if (accumulate) {
      float old[V];
      ldv<T, V>(o, old);
#pragma unroll
      for (int k = 0; k < V; ++k) g[k] += old[k];
    }
    stv<T, V>(o, g);
  }
}

NEMAR_API int nemar_copy_view_bwd(const nemar_tensor* dsrc_out, const nemar_tensor* ddst_in, int pad_mode,
                                  int accumulate, void* stream) {
  NEMAR_REQUIRE(view_ok(dsrc_out) && view_ok(ddst_in) && same_shape(dsrc_out, ddst_in) &&
                    dsrc_out->dtype == ddst_in->dtype,
                "copy_view_bwd: bad args");
  TView s = make_view(dsrc_out), d = make_view(ddst_in);
  DISPATCH_DTYPE(d.dtype, T, {
    constexpr int V = VecTraits<T>::V;
    if (view_vec_ok<T>(dsrc_out) && view_vec_ok<T>(ddst_in)) {
      int64_t total = (int64_t)s.n * s.h * s.w * (s.c / V);
      copy_view_bwd_kernel<T, V><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(s, d, pad_mode,
                                                                                         accumulate);
    } else {
      int64_t total = (int64_t)s.n * s.h * s.w * s.c;
      copy_view_bwd_kernel<T, 1><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(s, d, pad_mode,
                                                                                         accumulate);
    }
  });
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// per-(n,c) plane reductions (InstanceNorm statistics; backward sums; bias gradient)
// =============================================================================================
// MODE 0: (sum x, sum x^2)             MODE 1: (sum g, sum g*xhat), g = fold(dy)*act'(pre)
// grid = (chunks, n); partial sums are combined in shared memory then one global atomic per (c, stat).
template <typename T, int V, int MODE>
__global__ void __launch_bounds__(256)
plane_reduce_kernel(TView x, const float* __restrict__ stats, int act, TView dy, int pad_mode,
                    float inv_hw, float* __restrict__ out) {
  extern __shared__ float sacc[];  // [c][2]
  const int nn = blockIdx.y;
  const int c = x.c, G = c / V;
  for (int k = threadIdx.x; k < 2 * c; k += blockDim.x) sacc[k] = 0.f;
  __syncthreads();
  // per-sample item counts fit 32 bits (<= 1030^2 * 64): 32-bit index arithmetic keeps the loop HBM-bound
  const uint32_t hw = (uint32_t)x.h * (uint32_t)x.w;
  const uint32_t items = hw * (uint32_t)G;
  const uint32_t per_block = (items + gridDim.x - 1) / gridDim.x;
  // align each block's range to a multiple of G so that a thread keeps one channel group when
  // blockDim % G == 0 (register accumulation); otherwise fall back to per-item shared atomics.
  const uint32_t per_block_al = (per_block + G - 1) / G * G;
  const uint32_t lo = blockIdx.x * per_block_al;
  uint32_t hi = lo + per_block_al;
  if (hi > items || hi < lo) hi = items;
  const bool reg_path = (blockDim.x % G) == 0;
  float a0[V], a1[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { a0[k] = 0.f; a1[k] = 0.f; }
  int my_cg = threadIdx.x % G;
  float hmean[V], hrstd[V];
  if (MODE == 1 && reg_path) load_mean_rstd<V>(stats, nn, c, my_cg * V, inv_hw, hmean, hrstd);
  const uint32_t uw = (uint32_t)x.w, uG = (uint32_t)G;
#pragma unroll 2
  for (uint32_t j = lo + threadIdx.x; j < hi; j += blockDim.x) {
    const uint32_t p = j / uG;
    const int cg = reg_path ? my_cg : (int)(j - p * uG);
    const int yy = (int)(p / uw);
    const int xx = (int)(p - (uint32_t)yy * uw);
    float v[V];
    ldv<T, V>((const T*)x.ptr + x.pix(nn, yy, xx) + cg * V, v);
    float s0[V], s1[V];
    if constexpr (MODE == 0) {
#pragma unroll
      for (int k = 0; k < V; ++k) { s0[k] = v[k]; s1[k] = v[k] * v[k]; }
    } else {
      float g[V];
      load_fold<T, V>(dy, nn, yy, xx, cg * V, pad_mode, g);
      float mean[V], rstd[V];
      if (reg_path) {
#pragma unroll
        for (int k = 0; k < V; ++k) { mean[k] = hmean[k]; rstd[k] = hrstd[k]; }
      } else {
        load_mean_rstd<V>(stats, nn, c, cg * V, inv_hw, mean, rstd);
      }
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float xh = (v[k] - mean[k]) * rstd[k];
        float gg = g[k] * act_grad_from_x(xh, act);
        s0[k] = gg;
        s1[k] = gg * xh;
      }
    }
    if (reg_path) {
#pragma unroll
      for (int k = 0; k < V; ++k) { a0[k] += s0[k]; a1[k] += s1[k]; }
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        atomicAdd(&sacc[(cg * V + k) * 2], s0[k]);
        atomicAdd(&sacc[(cg * V + k) * 2 + 1], s1[k]);
      }
    }
  }
  if (reg_path) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      atomicAdd(&sacc[(my_cg * V + k) * 2], a0[k]);
      atomicAdd(&sacc[(my_cg * V + k) * 2 + 1], a1[k]);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 2 * c; k += blockDim.x)
    atomicAdd(out + (int64_t)nn * c * 2 + k, sacc[k]);
}

template <int MODE>
static int launch_plane_reduce(const nemar_tensor* xt, const float* stats, int act, const nemar_tensor* dyt,
                               int pad_mode, float* out, cudaStream_t s) {
  TView x = make_view(xt);
  TView dy = dyt ? make_view(dyt) : x;
  const float inv_hw = 1.f / ((float)x.h * (float)x.w);
  const int64_t hw = (int64_t)x.h * x.w;
  // bf16, 128-bit-accessible views: the cp.async-ring kernels (norm_lean.cuh); anything else: the generic templates
  if (nlean::eligible(xt) && (!dyt || nlean::eligible(dyt))) {
    const int S = nlean::pipe_stages(MODE == 0 ? 1 : 2);
    const size_t smem = sizeof(float) * 2 * x.c + 16384 + nlean::ring_bytes(S, MODE == 1 ? 2 : 1);
    // MODE 1 with a reflect halo walks the padded pixels of dy (see reduce_pipe_kernel)
    const int opad = (MODE == 1 && dy.pad > 0 && pad_mode == NEMAR_PAD_REFLECT) ? dy.pad : 0;
    const int64_t items = (int64_t)(x.h + 2 * opad) * (x.w + 2 * opad);
    dim3 grid(nlean::pipe_chunks(items, x.c / 8, x.n, smem, MODE == 0 ? 4 : 3), x.n);
    if (MODE == 0) {
      nlean::allow_smem((const void*)nlean::reduce_pipe_kernel<MODE, NEMAR_ACT_NONE>, smem);
      nlean::reduce_pipe_kernel<MODE, NEMAR_ACT_NONE><<<grid, 256, smem, s>>>(x, stats, act, dy, pad_mode, inv_hw, out, S);
    } else {
      NLEAN_ACT_SWITCH(act, (nlean::allow_smem((const void*)nlean::reduce_pipe_kernel<MODE, A>, smem),
                             nlean::reduce_pipe_kernel<MODE, A><<<grid, 256, smem, s>>>(x, stats, act, dy, pad_mode, inv_hw, out, S)));
    }
    NEMAR_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_DTYPE(x.dtype, T, {
    constexpr int VV = VecTraits<T>::V;
    bool vec = view_vec_ok<T>(xt) && (!dyt || view_vec_ok<T>(dyt));
    int G = vec ? x.c / VV : x.c;
    int64_t items = hw * G;
    int chunks = (int)((items + 256 * 16 - 1) / (256 * 16));    // more blocks were measured slower (2C global REDs per block)
    int cap = (148 * 8 + x.n - 1) / x.n;
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    size_t smem = sizeof(float) * 2 * x.c;
    if (vec)
      plane_reduce_kernel<T, VV, MODE><<<dim3(chunks, x.n), 256, smem, s>>>(x, stats, act, dy, pad_mode,
                                                                            inv_hw, out);
    else
      plane_reduce_kernel<T, 1, MODE><<<dim3(chunks, x.n), 256, smem, s>>>(x, stats, act, dy, pad_mode,
                                                                           inv_hw, out);
  });
  NEMAR_LAUNCH_CHECK();
  return 0;
}

NEMAR_API int nemar_instnorm_stats(const nemar_tensor* x, float* stats, void* stream) {
  NEMAR_REQUIRE(view_ok(x) && stats && x->c <= 4096, "instnorm_stats: bad args");
  return launch_plane_reduce<0>(x, nullptr, 0, nullptr, 0, stats, (cudaStream_t)stream);
}

NEMAR_API int nemar_norm_act_bwd_reduce(const nemar_tensor* x, const float* stats, int act,
                                        const nemar_tensor* dy, int pad_mode, float* red, void* stream) {
  NEMAR_REQUIRE(view_ok(x) && view_ok(dy) && stats && red && same_shape(x, dy) && x->dtype == dy->dtype &&
                    x->c <= 4096,
                "norm_act_bwd_reduce: bad args");
  return launch_plane_reduce<1>(x, stats, act, dy, pad_mode, red, (cudaStream_t)stream);
}

// bias gradient: db[c] = sum over n,h,w of dy
template <typename T, int V>
__global__ void __launch_bounds__(256) bias_grad_kernel(TView dy, float* __restrict__ db) {
  extern __shared__ float sacc[];
  const int c = dy.c, G = c / V;
  for (int k = threadIdx.x; k < c; k += blockDim.x) sacc[k] = 0.f;
  __syncthreads();
  const int64_t items = (int64_t)dy.n * dy.h * dy.w * G;
  const bool reg_path = (blockDim.x % G) == 0 && ((int64_t)gridDim.x * blockDim.x) % G == 0;
  float a[V];
#pragma unroll
  for (int k = 0; k < V; ++k) a[k] = 0.f;
  const int my_cg = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) % G);
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < items;
       j += (int64_t)gridDim.x * blockDim.x) {
    int cg = (int)(j % G);
    int64_t p = j / G;
    int xx = (int)(p % dy.w); p /= dy.w;
    int yy = (int)(p % dy.h);
    int nn = (int)(p / dy.h);
    float v[V];
    ldv<T, V>((const T*)dy.ptr + dy.pix(nn, yy, xx) + cg * V, v);
    if (reg_path) {
#pragma unroll
      for (int k = 0; k < V; ++k) a[k] += v[k];
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) atomicAdd(&sacc[cg * V + k], v[k]);
    }
  }
  if (reg_path) {
#pragma unroll
    for (int k = 0; k < V; ++k) atomicAdd(&sacc[my_cg * V + k], a[k]);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < c; k += blockDim.x) atomicAdd(db + k, sacc[k]);
}

NEMAR_API int nemar_bias_grad(const nemar_tensor* dy, float* db, void* stream) {
  NEMAR_REQUIRE(view_ok(dy) && db && dy->c <= 8192, "bias_grad: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  TView d = make_view(dy);
  cudaMemsetAsync(db, 0, sizeof(float) * d.c, s);
  DISPATCH_DTYPE(d.dtype, T, {
    constexpr int VV = VecTraits<T>::V;
    bool vec = view_vec_ok<T>(dy);
    int G = vec ? d.c / VV : d.c;
    int64_t items = (int64_t)d.n * d.h * d.w * G;
    int blocks = grid_for(items, 256 * 8, 148 * 4);
    size_t smem = sizeof(float) * d.c;
    if (vec) bias_grad_kernel<T, VV><<<blocks, 256, smem, s>>>(d, db);
    else bias_grad_kernel<T, 1><<<blocks, 256, smem, s>>>(d, db);
  });
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// norm + activation (+ residual) forward, writing the reflect halo of the destination in the same pass
// =============================================================================================
template <typename T, int V>
__global__ void __launch_bounds__(256)
norm_act_fwd_kernel(TView x, const float* __restrict__ stats, int act, TView res, int has_res, TView y,
                    int pad_mode, float inv_hw) {
  // grid = (chunks, n).  Items of one sample = padded destination positions x channel groups; when the block size
  // is a multiple of the group count every thread keeps ONE channel group, so mean/rstd are computed once.
  const int nn = blockIdx.y;
  const int G = y.c / V;
  const uint32_t items = (uint32_t)y.hp * (uint32_t)y.wp * (uint32_t)G;   // per sample: fits 32 bits
  const bool fixed = (blockDim.x % G) == 0;
  const int my_cg = (int)(threadIdx.x % G);
  float hmean[V], hrstd[V];
  if (stats && fixed) load_mean_rstd<V>(stats, nn, y.c, my_cg * V, inv_hw, hmean, hrstd);
  const uint32_t uG = (uint32_t)G, uwp = (uint32_t)y.wp, step = gridDim.x * blockDim.x;
#pragma unroll 2
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += step) {
    const uint32_t r = i / uG;
    const int cg = fixed ? my_cg : (int)(i - r * uG);
    const int yp = (int)(r / uwp);
    const int xp = (int)(r - (uint32_t)yp * uwp);
    int ys = yp - y.pad, xs = xp - y.pad;
    bool halo = ys < 0 || ys >= y.h || xs < 0 || xs >= y.w;
    float v[V];
    if (halo && pad_mode != NEMAR_PAD_REFLECT) {
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = 0.f;
    } else {
      ys = reflect_idx(ys, y.h);
      xs = reflect_idx(xs, y.w);
      ldv<T, V>((const T*)x.ptr + x.pix(nn, ys, xs) + cg * V, v);
      if (stats) {
        float mean[V], rstd[V];
        if (fixed) {
#pragma unroll
          for (int k = 0; k < V; ++k) { mean[k] = hmean[k]; rstd[k] = hrstd[k]; }
        } else {
          load_mean_rstd<V>(stats, nn, y.c, cg * V, inv_hw, mean, rstd);
        }
#pragma unroll
        for (int k = 0; k < V; ++k) v[k] = (v[k] - mean[k]) * rstd[k];
      }
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = act_fwd(v[k], act);
      if (has_res) {
        float rr[V];
        ldv<T, V>((const T*)res.ptr + res.pix(nn, ys, xs) + cg * V, rr);
#pragma unroll
        for (int k = 0; k < V; ++k) v[k] += rr[k];
      }
    }
    stv<T, V>((T*)y.ptr + y.pix_p(nn, yp, xp) + cg * V, v);
  }
}

static int plane_chunks(int64_t items_per_sample, int n) {
  int64_t chunks = (items_per_sample + 256 * 8 - 1) / (256 * 8);
  int64_t cap = (148 * 16 + n - 1) / n;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  return (int)chunks;
}

NEMAR_API int nemar_norm_act_fwd(const nemar_tensor* x, const float* stats, int act,
                                 const nemar_tensor* residual, const nemar_tensor* y, int pad_mode,
                                 void* stream) {
  NEMAR_REQUIRE(view_ok(x) && view_ok(y) && same_shape(x, y) && x->dtype == y->dtype,
                "norm_act_fwd: x/y mismatch");
  NEMAR_REQUIRE(!residual || (view_ok(residual) && same_shape(residual, x) && residual->dtype == x->dtype),
                "norm_act_fwd: residual mismatch");
  TView xv = make_view(x), yv = make_view(y), rv = residual ? make_view(residual) : xv;
  const float inv_hw = 1.f / ((float)xv.h * (float)xv.w);
  cudaStream_t s = (cudaStream_t)stream;
  if (nlean::eligible(x) && nlean::eligible(y) && (!residual || nlean::eligible(residual))) {
    const int S = nlean::pipe_stages(0);
    const size_t smem = sizeof(float) * 2 * yv.c + nlean::ring_bytes(S, residual ? 2 : 1);
    dim3 grid(nlean::pipe_chunks((int64_t)yv.hp * yv.wp, yv.c / 8, yv.n, smem, 4), yv.n);
    NLEAN_ACT_SWITCH(act, (nlean::allow_smem((const void*)nlean::fwd_pipe_kernel<A>, smem),
                           nlean::fwd_pipe_kernel<A><<<grid, 256, smem, s>>>(xv, stats, act, rv, residual != nullptr, yv, pad_mode, inv_hw, S)));
    NEMAR_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_DTYPE(xv.dtype, T, {
    constexpr int VV = VecTraits<T>::V;
    bool vec = view_vec_ok<T>(x) && view_vec_ok<T>(y) && (!residual || view_vec_ok<T>(residual));
    if (vec) {
      int64_t items = (int64_t)yv.hp * yv.wp * (yv.c / VV);
      norm_act_fwd_kernel<T, VV><<<dim3(plane_chunks(items, yv.n), yv.n), 256, 0, s>>>(xv, stats, act, rv, residual != nullptr,
                                                                                          yv, pad_mode, inv_hw);
    } else {
      int64_t items = (int64_t)yv.hp * yv.wp * yv.c;
      norm_act_fwd_kernel<T, 1><<<dim3(plane_chunks(items, yv.n), yv.n), 256, 0, s>>>(xv, stats, act, rv, residual != nullptr,
                                                                                         yv, pad_mode, inv_hw);
    }
  });
  NEMAR_LAUNCH_CHECK();
  return 0;
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
norm_act_bwd_apply_kernel(TView x, const float* __restrict__ stats, int act, TView dy, int pad_mode,
                          const float* __restrict__ red, TView dx, TView dres, int has_dres,
                          int dres_acc, float inv_hw, float* __restrict__ db) {
  // grid = (chunks, n); thread keeps one channel group when blockDim % G == 0 (statistics hoisted, bias-gradient
  // column sums kept in registers and flushed with one shared + one global atomic per channel per block).
  extern __shared__ float sdb[];   // [c] when db != nullptr
  const int nn = blockIdx.y;
  const int G = x.c / V;
  const uint32_t items = (uint32_t)x.h * (uint32_t)x.w * (uint32_t)G;   // per sample: fits 32 bits
  const bool fixed = (blockDim.x % G) == 0;
  const int my_cg = (int)(threadIdx.x % G);
  if (db) {
    for (int k = threadIdx.x; k < x.c; k += blockDim.x) sdb[k] = 0.f;
    __syncthreads();
  }
  float hmean[V], hrstd[V], hm1[V], hm2[V], bsum[V];
#pragma unroll
  for (int k = 0; k < V; ++k) bsum[k] = 0.f;
  if (stats && fixed) {
    load_mean_rstd<V>(stats, nn, x.c, my_cg * V, inv_hw, hmean, hrstd);
    const float* rd = red + ((int64_t)nn * x.c + my_cg * V) * 2;
#pragma unroll
    for (int k = 0; k < V; ++k) { hm1[k] = __ldg(rd + 2 * k) * inv_hw; hm2[k] = __ldg(rd + 2 * k + 1) * inv_hw; }
  }
  const uint32_t uG = (uint32_t)G, uw = (uint32_t)x.w, step = gridDim.x * blockDim.x;
#pragma unroll 2
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += step) {
    const uint32_t r = i / uG;
    const int cg = fixed ? my_cg : (int)(i - r * uG);
    const int yy = (int)(r / uw);
    const int xx = (int)(r - (uint32_t)yy * uw);
    float g[V], v[V], o[V];
    load_fold<T, V>(dy, nn, yy, xx, cg * V, pad_mode, g);
    if (has_dres) {
      T* rp = (T*)dres.ptr + dres.pix(nn, yy, xx) + cg * V;
      float t[V];
#pragma unroll
      for (int k = 0; k < V; ++k) t[k] = g[k];
      if (dres_acc) {
        float old[V];
        ldv<T, V>(rp, old);
#pragma unroll
        for (int k = 0; k < V; ++k) t[k] += old[k];
      }
      stv<T, V>(rp, t);
    }
    ldv<T, V>((const T*)x.ptr + x.pix(nn, yy, xx) + cg * V, v);
    if (stats) {
      float mean[V], rstd[V], m1[V], m2[V];
      if (fixed) {
#pragma unroll
        for (int k = 0; k < V; ++k) { mean[k] = hmean[k]; rstd[k] = hrstd[k]; m1[k] = hm1[k]; m2[k] = hm2[k]; }
      } else {
        load_mean_rstd<V>(stats, nn, x.c, cg * V, inv_hw, mean, rstd);
        const float* rd = red + ((int64_t)nn * x.c + cg * V) * 2;
#pragma unroll
        for (int k = 0; k < V; ++k) { m1[k] = __ldg(rd + 2 * k) * inv_hw; m2[k] = __ldg(rd + 2 * k + 1) * inv_hw; }
      }
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float xh = (v[k] - mean[k]) * rstd[k];
        float gg = g[k] * act_grad_from_x(xh, act);
        o[k] = rstd[k] * (gg - m1[k] - xh * m2[k]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) o[k] = g[k] * act_grad_from_x(v[k], act);
    }
    stv<T, V>((T*)dx.ptr + dx.pix(nn, yy, xx) + cg * V, o);
    if (db) {
      if (fixed) {
#pragma unroll
        for (int k = 0; k < V; ++k) bsum[k] += o[k];
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) atomicAdd(&sdb[cg * V + k], o[k]);
      }
    }
  }
  if (db) {
    if (fixed) {
#pragma unroll
      for (int k = 0; k < V; ++k) atomicAdd(&sdb[my_cg * V + k], bsum[k]);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < x.c; k += blockDim.x) atomicAdd(db + k, sdb[k]);
  }
}

// zero the halo ring of a padded view (the interior is left alone): grid (chunks, n)
template <typename T> __global__ void zero_halo_kernel(TView t) {
  const int nn = blockIdx.y, p = t.pad;
  const int ring = 2 * p * t.wp + 2 * p * t.h;            // halo pixels of one sample
  const int64_t total = (int64_t)ring * t.c;
  T* base = (T*)t.ptr + (int64_t)nn * t.hp * t.wp * t.cs + t.coff;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % t.c);
    int r = (int)(i / t.c), yp, xp;
    if (r < 2 * p * t.wp) {                                // top and bottom rows, full width
      const int row = r / t.wp;
      xp = r - row * t.wp;
      yp = row < p ? row : t.h + row;                      // rows p .. 2p-1 map to the bottom block
    } else {                                               // left and right columns of the interior rows
      r -= 2 * p * t.wp;
      const int row = r / (2 * p), k = r - row * (2 * p);
      yp = p + row;
      xp = k < p ? k : t.w + k;
    }
    base[((int64_t)yp * t.wp + xp) * t.cs + ch] = from_f<T>(0.f);
  }
}

NEMAR_API int nemar_norm_act_bwd_apply(const nemar_tensor* x, const float* stats, int act,
                                       const nemar_tensor* dy, int pad_mode, const float* red,
                                       const nemar_tensor* dx, const nemar_tensor* dres, int flags,
                                       float* db, void* stream) {
  const int dres_accumulate = flags & 1;          // dres += fold(dy) instead of =
  const bool db_accumulate = (flags & 2) != 0;    // db += column sums (caller's running total) instead of =
  const bool zero_dres_halo = (flags & 4) != 0;   // the halo ring of dres is zeroed here (the passes write its interior)
  if (dres && zero_dres_halo && dres->pad > 0 && view_ok(dres)) {
    TView dv = make_view(dres);
    const int64_t ring_items = (int64_t)(2 * dv.pad * dv.wp + 2 * dv.pad * dv.h) * dv.c;
    DISPATCH_DTYPE(dv.dtype, T, (zero_halo_kernel<T><<<dim3(grid_for(ring_items, 256, 8), dv.n), 256, 0, (cudaStream_t)stream>>>(dv)));
    NEMAR_LAUNCH_CHECK();
  }
  NEMAR_REQUIRE(view_ok(x) && view_ok(dy) && view_ok(dx) && same_shape(x, dy) && same_shape(x, dx) &&
                    x->dtype == dy->dtype && x->dtype == dx->dtype,
                "norm_act_bwd_apply: mismatch");
  NEMAR_REQUIRE(!stats || red, "norm_act_bwd_apply: red required with stats");
  NEMAR_REQUIRE(!dres || (view_ok(dres) && same_shape(dres, x) && dres->dtype == x->dtype),
                "norm_act_bwd_apply: dres mismatch");
  TView xv = make_view(x), dyv = make_view(dy), dxv = make_view(dx), dr = dres ? make_view(dres) : xv;
  const float inv_hw = 1.f / ((float)xv.h * (float)xv.w);
  cudaStream_t s = (cudaStream_t)stream;
  if (nlean::eligible(x) && nlean::eligible(dy) && nlean::eligible(dx) && (!dres || nlean::eligible(dres))) {
    if (db && !db_accumulate) cudaMemsetAsync(db, 0, sizeof(float) * xv.c, s);
    const int S = nlean::pipe_stages(3);
    const int ring_nt = ((dres && dres_accumulate) ? 3 : 2) + ((dyv.pad > 0 && pad_mode == NEMAR_PAD_REFLECT) ? 1 : 0);
    const size_t smem = sizeof(float) * 4 * xv.c + 8192 + nlean::ring_bytes(S, ring_nt);
    dim3 grid(nlean::pipe_chunks((int64_t)xv.h * xv.w, xv.c / 8, xv.n, smem, 2), xv.n);
    NLEAN_ACT_SWITCH(act, (nlean::allow_smem((const void*)nlean::bwd_apply_pipe_kernel<A>, smem),
                           nlean::bwd_apply_pipe_kernel<A><<<grid, 256, smem, s>>>(
                               xv, stats, act, dyv, pad_mode, red, dxv, dr, dres != nullptr, dres_accumulate, inv_hw, db, S)));
    NEMAR_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_DTYPE(xv.dtype, T, {
    constexpr int VV = VecTraits<T>::V;
    bool vec = view_vec_ok<T>(x) && view_vec_ok<T>(dy) && view_vec_ok<T>(dx) && (!dres || view_vec_ok<T>(dres));
    size_t smem = db ? sizeof(float) * xv.c : 0;
    if (db && !db_accumulate) cudaMemsetAsync(db, 0, sizeof(float) * xv.c, s);
    if (vec) {
      int64_t items = (int64_t)xv.h * xv.w * (xv.c / VV);
      norm_act_bwd_apply_kernel<T, VV><<<dim3(plane_chunks(items, xv.n), xv.n), 256, smem, s>>>(
          xv, stats, act, dyv, pad_mode, red, dxv, dr, dres != nullptr, dres_accumulate, inv_hw, db);
    } else {
      int64_t items = (int64_t)xv.h * xv.w * xv.c;
      norm_act_bwd_apply_kernel<T, 1><<<dim3(plane_chunks(items, xv.n), xv.n), 256, smem, s>>>(
          xv, stats, act, dyv, pad_mode, red, dxv, dr, dres != nullptr, dres_accumulate, inv_hw, db);
    }
  });
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// dx = dy * act'(y)  (activation fused in a conv epilogue; y is the activation output)
template <typename T, int V> __global__ void act_bwd_kernel(TView y, TView dy, int act, TView dx) {
  const int G = y.c / V;
  const int64_t total = (int64_t)y.n * y.h * y.w * G;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int cg = (int)(i % G);
    int64_t r = i / G;
    int xx = (int)(r % y.w); r /= y.w;
    int yy = (int)(r % y.h);
    int nn = (int)(r / y.h);
    float a[V], g[V];
    ldv<T, V>((const T*)y.ptr + y.pix(nn, yy, xx) + cg * V, a);
    ldv<T, V>((const T*)dy.ptr + dy.pix(nn, yy, xx) + cg * V, g);
#pragma unroll
    for (int k = 0; k < V; ++k) g[k] *= act_grad_from_y(a[k], act);
    stv<T, V>((T*)dx.ptr + dx.pix(nn, yy, xx) + cg * V, g);
  }
}

NEMAR_API int nemar_act_bwd(const nemar_tensor* y, const nemar_tensor* dy, int act, const nemar_tensor* dx,
                            void* stream) {
  NEMAR_REQUIRE(view_ok(y) && view_ok(dy) && view_ok(dx) && same_shape(y, dy) && same_shape(y, dx) &&
                    y->dtype == dy->dtype && y->dtype == dx->dtype,
                "act_bwd: mismatch");
  TView yv = make_view(y), dyv = make_view(dy), dxv = make_view(dx);
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_DTYPE(yv.dtype, T, {
    constexpr int VV = VecTraits<T>::V;
    if (view_vec_ok<T>(y) && view_vec_ok<T>(dy) && view_vec_ok<T>(dx)) {
      int64_t total = (int64_t)yv.n * yv.h * yv.w * (yv.c / VV);
      act_bwd_kernel<T, VV><<<grid_for(total, 256), 256, 0, s>>>(yv, dyv, act, dxv);
    } else {
      int64_t total = (int64_t)yv.n * yv.h * yv.w * yv.c;
      act_bwd_kernel<T, 1><<<grid_for(total, 256), 256, 0, s>>>(yv, dyv, act, dxv);
    }
  });
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// MaxPool2d(2), floor mode
// =============================================================================================
template <typename T, int V> __global__ void maxpool2_fwd_kernel(TView x, TView y) {
  const int G = y.c / V;
  const int64_t total = (int64_t)y.n * y.h * y.w * G;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int cg = (int)(i % G);
    int64_t r = i / G;
    int xo = (int)(r % y.w); r /= y.w;
    int yo = (int)(r % y.h);
    int nn = (int)(r / y.h);
    float m[V];
    ldv<T, V>((const T*)x.ptr + x.pix(nn, 2 * yo, 2 * xo) + cg * V, m);
#pragma unroll
    for (int q = 1; q < 4; ++q) {
      float v[V];
      ldv<T, V>((const T*)x.ptr + x.pix(nn, 2 * yo + (q >> 1), 2 * xo + (q & 1)) + cg * V, v);
#pragma unroll
      for (int k = 0; k < V; ++k) m[k] = v[k] > m[k] ? v[k] : m[k];
    }
    stv<T, V>((T*)y.ptr + y.pix(nn, yo, xo) + cg * V, m);
  }
}

// gradient goes to the first maximum in row-major window order (ATen max_pool2d tie rule)
template <typename T, int V> __global__ void maxpool2_bwd_kernel(TView x, TView dy, TView dx) {
  const int G = x.c / V;
  const int64_t total = (int64_t)x.n * x.h * x.w * G;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int cg = (int)(i % G);
    int64_t r = i / G;
    int xx = (int)(r % x.w); r /= x.w;
    int yy = (int)(r % x.h);
    int nn = (int)(r / x.h);
    int yo = yy >> 1, xo = xx >> 1;
    float o[V];
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = 0.f;
    if (yo < dy.h && xo < dy.w) {
      const int me = ((yy & 1) << 1) | (xx & 1);
      float w[4][V];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        ldv<T, V>((const T*)x.ptr + x.pix(nn, 2 * yo + (q >> 1), 2 * xo + (q & 1)) + cg * V, w[q]);
      float g[V];
      ldv<T, V>((const T*)dy.ptr + dy.pix(nn, yo, xo) + cg * V, g);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        int arg = 0;
        float m = w[0][k];
#pragma unroll
        for (int q = 1; q < 4; ++q)
          if (w[q][k] > m) { m = w[q][k]; arg = q; }
        o[k] = (arg == me) ? g[k] : 0.f;
      }
    }
    stv<T, V>((T*)dx.ptr + dx.pix(nn, yy, xx) + cg * V, o);
  }
}

NEMAR_API int nemar_maxpool2_fwd(const nemar_tensor* x, const nemar_tensor* y, void* stream) {
  NEMAR_REQUIRE(view_ok(x) && view_ok(y) && x->n == y->n && x->c == y->c && y->h == x->h / 2 &&
                    y->w == x->w / 2 && x->dtype == y->dtype,
                "maxpool2_fwd: mismatch");
  TView xv = make_view(x), yv = make_view(y);
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_DTYPE(xv.dtype, T, {
    constexpr int VV = VecTraits<T>::V;
    if (view_vec_ok<T>(x) && view_vec_ok<T>(y)) {
      int64_t total = (int64_t)yv.n * yv.h * yv.w * (yv.c / VV);
      maxpool2_fwd_kernel<T, VV><<<grid_for(total, 256), 256, 0, s>>>(xv, yv);
    } else {
      int64_t total = (int64_t)yv.n * yv.h * yv.w * yv.c;
      maxpool2_fwd_kernel<T, 1><<<grid_for(total, 256), 256, 0, s>>>(xv, yv);
    }
  });
  NEMAR_LAUNCH_CHECK();
  return 0;
}

NEMAR_API int nemar_maxpool2_bwd(const nemar_tensor* x, const nemar_tensor* y, const nemar_tensor* dy,
                                 const nemar_tensor* dx, void* stream) {
  (void)y;
  NEMAR_REQUIRE(view_ok(x) && view_ok(dy) && view_ok(dx) && same_shape(x, dx) && dy->h == x->h / 2 &&
                    dy->w == x->w / 2 && dy->c == x->c && x->dtype == dy->dtype && x->dtype == dx->dtype,
                "maxpool2_bwd: mismatch");
  TView xv = make_view(x), dyv = make_view(dy), dxv = make_view(dx);
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_DTYPE(xv.dtype, T, {
    constexpr int VV = VecTraits<T>::V;
    if (view_vec_ok<T>(x) && view_vec_ok<T>(dy) && view_vec_ok<T>(dx)) {
      int64_t total = (int64_t)xv.n * xv.h * xv.w * (xv.c / VV);
      maxpool2_bwd_kernel<T, VV><<<grid_for(total, 256), 256, 0, s>>>(xv, dyv, dxv);
    } else {
      int64_t total = (int64_t)xv.n * xv.h * xv.w * xv.c;
      maxpool2_bwd_kernel<T, 1><<<grid_for(total, 256), 256, 0, s>>>(xv, dyv, dxv);
    }
  });
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// bilinear resize, align_corners=False (ATen area_pixel_compute_source_index + upsample_bilinear2d)
// =============================================================================================
struct Lerp { int i0, i1; float l0, l1; };
__device__ __forceinline__ Lerp lerp_src(int o, float scale, int in_size) {
  float src = scale * ((float)o + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  Lerp L;
  L.i0 = (int)src;
  if (L.i0 > in_size - 1) L.i0 = in_size - 1;
  L.i1 = L.i0 + ((L.i0 < in_size - 1) ? 1 : 0);
  L.l1 = src - (float)L.i0;
  L.l0 = 1.f - L.l1;
  return L;
}

template <typename T, int V> __global__ void resize_fwd_kernel(TView x, TView y, float sh, float sw) {
  const int G = y.c / V;
  const int64_t total = (int64_t)y.n * y.h * y.w * G;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int cg = (int)(i % G);
    int64_t r = i / G;
    int xo = (int)(r % y.w); r /= y.w;
    int yo = (int)(r % y.h);
    int nn = (int)(r / y.h);
    Lerp ly = lerp_src(yo, sh, x.h), lx = lerp_src(xo, sw, x.w);
    float a[V], b[V], c[V], d[V], o[V];
    const T* base = (const T*)x.ptr;
    ldv<T, V>(base + x.pix(nn, ly.i0, lx.i0) + cg * V, a);
    ldv<T, V>(base + x.pix(nn, ly.i0, lx.i1) + cg * V, b);
    ldv<T, V>(base + x.pix(nn, ly.i1, lx.i0) + cg * V, c);
    ldv<T, V>(base + x.pix(nn, ly.i1, lx.i1) + cg * V, d);
#pragma unroll
    for (int k = 0; k < V; ++k)
      o[k] = ly.l0 * (lx.l0 * a[k] + lx.l1 * b[k]) + ly.l1 * (lx.l0 * c[k] + lx.l1 * d[k]);
    stv<T, V>((T*)y.ptr + y.pix(nn, yo, xo) + cg * V, o);
  }
}

// candidate output range that may read input index i:  src(o) in (i-1, i+1)
__device__ __forceinline__ void cand_range(int i, float scale, int out_size, int& lo, int& hi) {
  float inv = 1.f / scale;
  lo = (int)floorf(((float)i - 1.f + 0.5f) * inv - 0.5f) - 1;
  hi = (int)ceilf(((float)i + 1.f + 0.5f) * inv - 0.5f) + 1;
  if (lo < 0) lo = 0;
  if (hi > out_size - 1) hi = out_size - 1;
}

template <typename T, int V>
__global__ void resize_bwd_kernel(TView dy, TView dx, float sh, float sw, int accumulate) {
  // gather adjoint: every input element sums the output gradients whose taps touch it
  const int G = dx.c / V;
  const int64_t total = (int64_t)dx.n * dx.h * dx.w * G;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int cg = (int)(i % G);
    int64_t r = i / G;
    int xi = (int)(r % dx.w); r /= dx.w;
    int yi = (int)(r % dx.h);
    int nn = (int)(r / dx.h);
    int ylo, yhi, xlo, xhi;
    cand_range(yi, sh, dy.h, ylo, yhi);
    cand_range(xi, sw, dy.w, xlo, xhi);
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    for (int yo = ylo; yo <= yhi; ++yo) {
      Lerp ly = lerp_src(yo, sh, dx.h);
      float wy = (ly.i0 == yi ? ly.l0 : 0.f) + (ly.i1 == yi ? ly.l1 : 0.f);
      if (wy == 0.f) continue;
      for (int xo = xlo; xo <= xhi; ++xo) {
        Lerp lx = lerp_src(xo, sw, dx.w);
        float wx = (lx.i0 == xi ? lx.l0 : 0.f) + (lx.i1 == xi ? lx.l1 : 0.f);
        if (wx == 0.f) continue;
        float g[V];
        ldv<T, V>((const T*)dy.ptr + dy.pix(nn, yo, xo) + cg * V, g);
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += wy * wx * g[k];
      }
    }
    T* o = (T*)dx.ptr + dx.pix(nn, yi, xi) + cg * V;
    if (accumulate) {
      float old[V];
      ldv<T, V>(o, old);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += old[k];
    }
    stv<T, V>(o, acc);
  }
}

NEMAR_API int nemar_bilinear_resize_fwd(const nemar_tensor* x, const nemar_tensor* y, void* stream) {
  NEMAR_REQUIRE(view_ok(x) && view_ok(y) && x->n == y->n && x->c == y->c && x->dtype == y->dtype,
                "bilinear_resize_fwd: mismatch");
  TView xv = make_view(x), yv = make_view(y);
  float sh = (float)xv.h / (float)yv.h, sw = (float)xv.w / (float)yv.w;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_DTYPE(xv.dtype, T, {
    constexpr int VV = VecTraits<T>::V;
    if (view_vec_ok<T>(x) && view_vec_ok<T>(y)) {
      int64_t total = (int64_t)yv.n * yv.h * yv.w * (yv.c / VV);
      resize_fwd_kernel<T, VV><<<grid_for(total, 256), 256, 0, s>>>(xv, yv, sh, sw);
    } else {
      int64_t total = (int64_t)yv.n * yv.h * yv.w * yv.c;
      resize_fwd_kernel<T, 1><<<grid_for(total, 256), 256, 0, s>>>(xv, yv, sh, sw);
    }
  });
  NEMAR_LAUNCH_CHECK();
  return 0;
}

NEMAR_API int nemar_bilinear_resize_bwd(const nemar_tensor* dy, const nemar_tensor* dx, int accumulate,
                                        void* stream) {
  NEMAR_REQUIRE(view_ok(dy) && view_ok(dx) && dy->n == dx->n && dy->c == dx->c && dy->dtype == dx->dtype,
                "bilinear_resize_bwd: mismatch");
  TView dyv = make_view(dy), dxv = make_view(dx);
  float sh = (float)dxv.h / (float)dyv.h, sw = (float)dxv.w / (float)dyv.w;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_DTYPE(dxv.dtype, T, {
    constexpr int VV = VecTraits<T>::V;
    if (view_vec_ok<T>(dy) && view_vec_ok<T>(dx)) {
      int64_t total = (int64_t)dxv.n * dxv.h * dxv.w * (dxv.c / VV);
      resize_bwd_kernel<T, VV><<<grid_for(total, 256), 256, 0, s>>>(dyv, dxv, sh, sw, accumulate);
    } else {
      int64_t total = (int64_t)dxv.n * dxv.h * dxv.w * dxv.c;
      resize_bwd_kernel<T, 1><<<grid_for(total, 256), 256, 0, s>>>(dyv, dxv, sh, sw, accumulate);
    }
  });
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// NCHW fp32 images: express a plane stack [N*C, H, W] as an NHWC view with c == 1
static nemar_tensor plane_view(const float* p, int planes, int h, int w) {
  nemar_tensor t;
  t.ptr = (void*)p; t.n = planes; t.h = h; t.w = w; t.c = 1; t.pad = 0; t.cs = 1; t.coff = 0;
  t.dtype = NEMAR_F32;
  return t;
}

NEMAR_API int nemar_bilinear_resize_nchw_fwd(const float* x, int n, int c, int h, int w, float* y, int ho,
                                             int wo, void* stream) {
  NEMAR_REQUIRE(x && y && n > 0 && c > 0 && h > 0 && w > 0 && ho > 0 && wo > 0, "resize_nchw_fwd: bad args");
  nemar_tensor xv = plane_view(x, n * c, h, w), yv = plane_view(y, n * c, ho, wo);
  return nemar_bilinear_resize_fwd(&xv, &yv, stream);
}

NEMAR_API int nemar_bilinear_resize_nchw_bwd(const float* dy, int n, int c, int h, int w, float* dx, int ho,
                                             int wo, void* stream) {
  NEMAR_REQUIRE(dy && dx && n > 0 && c > 0 && h > 0 && w > 0 && ho > 0 && wo > 0, "resize_nchw_bwd: bad args");
  nemar_tensor dyv = plane_view(dy, n * c, ho, wo), dxv = plane_view(dx, n * c, h, w);
  return nemar_bilinear_resize_bwd(&dyv, &dxv, 0, stream);
}

// =============================================================================================
// Dropout(0.5) — counter-based hash, one 64-bit draw per element (not Philox-compatible by design)
// =============================================================================================
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// `step`: optional device-resident step counter; the mask of a replayed CUDA graph then changes from step to step
template <typename T> __global__ void dropout_kernel(TView x, TView y, uint64_t seed, uint64_t offset,
                                                     const int64_t* __restrict__ step) {
  if (step) offset += (uint64_t)(*step) * 0x9E3779B97F4A7C15ull;
  const int64_t total = (int64_t)x.n * x.h * x.w * x.c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int ch = (int)(i % x.c);
    int64_t r = i / x.c;
    int xx = (int)(r % x.w); r /= x.w;
    int yy = (int)(r % x.h);
    int nn = (int)(r / x.h);
    uint64_t h = splitmix64(seed ^ splitmix64(offset + (uint64_t)i));
    float v = to_f<T>(((const T*)x.ptr)[x.pix(nn, yy, xx) + ch]);
    ((T*)y.ptr)[y.pix(nn, yy, xx) + ch] = from_f<T>((h >> 63) ? 2.f * v : 0.f);
  }
}

static int dropout_impl(const nemar_tensor* x, const nemar_tensor* y, uint64_t seed, uint64_t offset, const int64_t* step,
                        void* stream) {
  NEMAR_REQUIRE(view_ok(x) && view_ok(y) && same_shape(x, y) && x->dtype == y->dtype, "dropout: mismatch");
  TView xv = make_view(x), yv = make_view(y);
  int64_t total = (int64_t)xv.n * xv.h * xv.w * xv.c;
  DISPATCH_DTYPE(xv.dtype, T, (dropout_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
                                  xv, yv, seed, offset, step)));
  NEMAR_LAUNCH_CHECK();
  return 0;
}

NEMAR_API int nemar_dropout(const nemar_tensor* x, const nemar_tensor* y, uint64_t seed, uint64_t offset,
                            void* stream) {
  return dropout_impl(x, y, seed, offset, nullptr, stream);
}

NEMAR_API int nemar_dropout_dev(const nemar_tensor* x, const nemar_tensor* y, uint64_t seed, uint64_t salt,
                                const int64_t* step_dev, void* stream) {
  NEMAR_REQUIRE(step_dev, "dropout_dev: step counter");
  return dropout_impl(x, y, seed, salt, step_dev, stream);
}

__global__ void counter_add_kernel(int64_t* c, int64_t v) { *c += v; }
NEMAR_API int nemar_counter_add(int64_t* counter_dev, int64_t v, void* stream) {
  NEMAR_REQUIRE(counter_dev, "counter_add: pointer");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter_dev, v);
  NEMAR_LAUNCH_CHECK();
  return 0;
}
