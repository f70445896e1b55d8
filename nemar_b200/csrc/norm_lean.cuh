// norm_lean.cuh — bf16 kernels for the InstanceNorm passes (forward, statistics, backward-reduce, backward-apply).
//
// These passes are HBM-bound streaming kernels over 30-130 MB tensors; what they reach is decided by how many bytes
// an SM keeps in flight and by the fixed cost per CTA.  History (profiles/r01*, r02_nbench_*): the generic templates
// in elementwise.cu interleave loads and stores through possibly-aliasing pointers (one item in flight per thread,
// ~1.1-2.5 TB/s); a first fast path kept a thread's per-channel constants in registers (104-158 registers, 1-2 CTAs
// per SM); the "lean" kernels moved the constants to shared memory as fused multiply-add coefficients and kept two
// raw 16-byte loads per tensor in flight per thread (2.6-2.8 TB/s on the 256-channel 64x64 maps, 4.3 on the
// 134 MB ones).  The kernels below replace all of them: every thread fills a private shared-memory ring with
// cp.async, so the loads of S - 1 later iterations stay in flight through the arithmetic and the stores without
// costing registers.
//
//   xhat = A*x + B                 A = rstd, B = -mean*rstd            (identity when there are no statistics)
//   dx   = A*g' + C + xhat*D       C = -rstd*mean(g'), D = -rstd*mean(g'*xhat),  g' = fold(dy)*act'(xhat)
//
//   requirements (eligible()): bf16 storage, 128-bit-accessible views, 256 % (C/8) == 0; anything else takes the
//   generic templates.
#pragma once
#include <mutex>
#include <unordered_set>
#include "common.cuh"

namespace nlean {

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

// gradient of the padded buffer folded onto interior pixel (y,x): main tap given, mirrored taps added when needed.
// `first` (optional): the FIRST mirrored tap in this enumeration order, already fetched (first_mirror() names it).
__device__ __forceinline__ void fold_extra(const TView& d, int nn, int y, int x, int c0, float (&g)[8], const uint4* first = nullptr) {
  int ys[3], xs[3];
  const int ny = reflect_sources(y, d.h, d.pad, ys);
  const int nx = reflect_sources(x, d.w, d.pad, xs);
  int idx = 0;
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) {
      if (a == 0 && b == 0) continue;   // (ys[0], xs[0]) is the main tap, already loaded
      float t[8];
      unpack8((first && idx == 0) ? *first : ldraw((const __nv_bfloat16*)d.ptr + d.pix_p(nn, ys[a], xs[b]) + c0), t);
      ++idx;
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] += t[k];
    }
}
// padded coordinates of the first mirrored tap of interior pixel (y,x) in fold_extra's order; false: there is none
__device__ __forceinline__ bool first_mirror(const TView& d, int y, int x, int& yp, int& xp) {
  int ys[3], xs[3];
  const int ny = reflect_sources(y, d.h, d.pad, ys);
  const int nx = reflect_sources(x, d.w, d.pad, xs);
  if (nx > 1) { yp = ys[0]; xp = xs[1]; return true; }
  if (ny > 1) { yp = ys[1]; xp = xs[0]; return true; }
  return false;
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

static inline bool eligible(const nemar_tensor* t) {
  return t->dtype == NEMAR_BF16 && t->c % 8 == 0 && t->cs % 8 == 0 && t->coff % 8 == 0 && ((((uintptr_t)t->ptr) & 15) == 0) &&
         (256 % (t->c / 8)) == 0;
}

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

// a thread's 8 coefficients of one table into registers (the thread owns ONE channel group for all its pixels)
__device__ __forceinline__ void coef8(const float* t, int c0, float (&r)[8]) {
  const float4 lo = *reinterpret_cast<const float4*>(t + c0), hi = *reinterpret_cast<const float4*>(t + c0 + 4);
  r[0] = lo.x; r[1] = lo.y; r[2] = lo.z; r[3] = lo.w; r[4] = hi.x; r[5] = hi.y; r[6] = hi.z; r[7] = hi.w;
}

// ---------------------------------------------------------------------------------------------
// per-thread cp.async rings
// ---------------------------------------------------------------------------------------------
// Register-staged loads expose one full memory round trip per iteration, and more loads per thread cost registers,
// i.e. residency (U = 4 and a register-prefetch variant were both slower).  Here every thread owns a private ring of S stages x P
// pixels x NT tensors x 16 bytes in SHARED memory and fills it with cp.async (LDGSTS: no destination registers): the
// loads of iteration i + S - 1 are issued before iteration i is consumed, so (S - 1) * P * NT * 16 bytes per thread
// stay in flight through the arithmetic and the stores.  A slot is read only by the thread that filled it
// (cp.async.wait_group orders a thread's own copies), so the ring needs no block barrier.  The coefficient table is
// built AFTER the first S - 1 stages have been issued: its dependent statistics loads hide behind the first data.
__device__ __forceinline__ void cp_async16(uint4* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most n of this thread's most recent groups are pending (n is uniform; the instruction wants an immediate)
__device__ __forceinline__ void cp_async_wait_pending(int n) {
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
  }
}
constexpr int P = 2;            // pixels per thread and stage
constexpr int MAX_STAGES = 8;

// plane reductions.  MODE 0: (sum x, sum x^2).  MODE 1: (sum g', sum g'*xhat), g' = fold(dy)*act'(xhat).
// The fold of a reflect halo is linear, so MODE 1 does not fold at all: it walks the PADDED pixels q of dy and pairs
// dy(q) with x at the pixel q mirrors (66x66 items instead of 64x64; the re-read rows of x are L2 hits).  The
// synchronous mirror loads of a folding kernel made the CTAs that own the border rows the stragglers of the launch.
template <int MODE, int ACT>
__global__ void __launch_bounds__(256, 3)
reduce_pipe_kernel(TView x, const float* __restrict__ stats, int act, TView dy, int pad_mode, float inv_hw,
                   float* __restrict__ out, int S) {
  constexpr int NT = MODE == 1 ? 2 : 1;
  extern __shared__ float sm[];     // A[c] | B[c] | partials[256/G][c][2] (16 KB) | ring[S][P][NT][256] x 16 B
  const int nn = blockIdx.y, c = x.c, G = c / 8;
  float* sA = sm; float* sB = sm + c;
  float* spart = sm + 2 * c;
  uint4* ring = reinterpret_cast<uint4*>(sm + 2 * c + 4096) + threadIdx.x;
  const bool padded = MODE == 1 && dy.pad > 0 && pad_mode == NEMAR_PAD_REFLECT;
  const int opad = padded ? dy.pad : 0;                  // the walked extent: interior + opad on every side
  const uint32_t ew = (uint32_t)(x.w + 2 * opad);
  const Range r = block_range((uint32_t)(x.h + 2 * opad) * ew, G);
  const int c0 = r.cg * 8;
  const __nv_bfloat16* xb = sample_base(x, nn, c0);
  const __nv_bfloat16* db = sample_base(dy, nn, c0);
  const float inv_w = 1.f / (float)ew;
  const uint32_t first = r.lo + r.pl, step = P * r.ppi;
  const uint32_t niter = (r.hi - r.lo + step - 1) / step;          // the same for every thread of the block
  auto issue = [&](uint32_t it, int slot) {
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const uint32_t p = first + it * step + k * r.ppi;
      if (it < niter && p < r.hi) {
        int yy, xx;
        divmod(p, ew, inv_w, yy, xx);
        yy -= opad; xx -= opad;                            // interior coordinates of the walked pixel (halo: outside)
        cp_async16(ring + ((slot * P + k) * NT) * 256, xb + off_in(x, reflect_idx(yy, x.h), reflect_idx(xx, x.w)));
        if (MODE == 1) cp_async16(ring + ((slot * P + k) * NT + 1) * 256, db + off_in(dy, yy, xx));
      }
    }
    cp_async_commit();
  };
  for (int s = 0; s < S - 1; ++s) issue((uint32_t)s, s);
  if (MODE == 1) fill_ab(stats, nn, c, inv_hw, sA, sB);
  __syncthreads();
  float a0[8], a1[8], cA[8], cB[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a0[k] = 0.f; a1[k] = 0.f; cA[k] = 1.f; cB[k] = 0.f; }
  // the coefficients of this thread's channel group live in REGISTERS: re-reading them from shared memory for every
  // pixel (the cp.async asm is a memory clobber, so the compiler must) kept the L1 / shared pipe 60-78 % busy (ncu)
  if (MODE == 1) { coef8(sA, c0, cA); coef8(sB, c0, cB); }
  int slot = 0, islot = S - 1;
  for (uint32_t it = 0; it < niter; ++it) {
    issue(it + (uint32_t)(S - 1), islot);
    if (++islot == S) islot = 0;
    cp_async_wait_pending(S - 1);
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const uint32_t p = first + it * step + k * r.ppi;
      if (p >= r.hi) continue;
      float v[8];
      unpack8(ring[((slot * P + k) * NT) * 256], v);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] += v[j]; a1[j] = fmaf(v[j], v[j], a1[j]); }
      } else {
        float g[8];
        unpack8(ring[((slot * P + k) * NT + 1) * 256], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = fmaf(v[j], cA[j], cB[j]);
          const float gg = g[j] * actg<ACT>(xh, act);
          a0[j] += gg;
          a1[j] = fmaf(gg, xh, a1[j]);
        }
      }
    }
    if (++slot == S) slot = 0;
  }
  // Block combine WITHOUT shared-memory atomics: the 256 / G threads that own the same channel group park their partial
  // sums in [pixel lane][c][2] order (16 consecutive floats per thread: conflict-free vector stores), then 2c threads
  // add the pixel lanes up.  (Round 1 used atomicAdd on a [c][2] table: the 32 lanes of a warp hit two banks — 16-way
  // conflicts on 16 atomics per thread, ~30 us of a 30 us statistics pass; found with scripts/nbench.py.)
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

// forward: y = act(A*x + B) (+ residual); the halo of y is written in the same pass
template <int ACT>
__global__ void __launch_bounds__(256, 3)
fwd_pipe_kernel(TView x, const float* __restrict__ stats, int act, TView res, int has_res, TView y, int pad_mode,
                float inv_hw, int S) {
  extern __shared__ float sm[];     // A[c] | B[c] | ring[S][P][nt][256] x 16 B
  const int nn = blockIdx.y, c = y.c, G = c / 8;
  const int nt = has_res ? 2 : 1;
  float* sA = sm; float* sB = sm + c;
  uint4* ring = reinterpret_cast<uint4*>(sm + 2 * c) + threadIdx.x;
  const Range r = block_range((uint32_t)y.hp * y.wp, G);
  const int c0 = r.cg * 8;
  const __nv_bfloat16* xb = sample_base(x, nn, c0);
  const __nv_bfloat16* rb = sample_base(res, nn, c0);
  __nv_bfloat16* yb = const_cast<__nv_bfloat16*>(sample_base(y, nn, c0));
  const uint32_t uwp = (uint32_t)y.wp;
  const float inv_wp = 1.f / (float)y.wp;
  const bool reflect = pad_mode == NEMAR_PAD_REFLECT;
  const uint32_t first = r.lo + r.pl, step = P * r.ppi;
  const uint32_t niter = (r.hi - r.lo + step - 1) / step;
  // padded pixel p -> source pixel (ys, xs) of x; false: a zero-halo pixel (nothing to load)
  auto source = [&](uint32_t p, int& ys, int& xs) -> bool {
    int yp, xp;
    divmod(p, uwp, inv_wp, yp, xp);
    ys = yp - y.pad; xs = xp - y.pad;
    const bool halo = ys < 0 || ys >= y.h || xs < 0 || xs >= y.w;
    if (halo && !reflect) return false;
    ys = reflect_idx(ys, y.h); xs = reflect_idx(xs, y.w);
    return true;
  };
  auto issue = [&](uint32_t it, int slot) {
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const uint32_t p = first + it * step + k * r.ppi;
      int ys, xs;
      if (it < niter && p < r.hi && source(p, ys, xs)) {
        cp_async16(ring + ((slot * P + k) * nt) * 256, xb + off_in(x, ys, xs));
        if (has_res) cp_async16(ring + ((slot * P + k) * nt + 1) * 256, rb + off_in(res, ys, xs));
      }
    }
    cp_async_commit();
  };
  for (int s = 0; s < S - 1; ++s) issue((uint32_t)s, s);
  fill_ab(stats, nn, c, inv_hw, sA, sB);
  __syncthreads();
  float cA[8], cB[8];
  coef8(sA, c0, cA);
  coef8(sB, c0, cB);
  int slot = 0, islot = S - 1;
  for (uint32_t it = 0; it < niter; ++it) {
    issue(it + (uint32_t)(S - 1), islot);
    if (++islot == S) islot = 0;
    cp_async_wait_pending(S - 1);
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const uint32_t p = first + it * step + k * r.ppi;
      if (p >= r.hi) continue;
      int ys, xs;
      float v[8];
      if (source(p, ys, xs)) {
        unpack8(ring[((slot * P + k) * nt) * 256], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = actf<ACT>(fmaf(v[j], cA[j], cB[j]), act);
        if (has_res) {
          float q[8];
          unpack8(ring[((slot * P + k) * nt + 1) * 256], q);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += q[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      *reinterpret_cast<uint4*>(yb + p * (uint32_t)y.cs) = pack8(v);      // == off_pad(y, yp, xp)
    }
    if (++slot == S) slot = 0;
  }
}

// backward apply: dx = A*g' + C + xhat*D;  dres (+)= fold(dy);  db += column sums of dx
template <int ACT>
__global__ void __launch_bounds__(256, 2)
bwd_apply_pipe_kernel(TView x, const float* __restrict__ stats, int act, TView dy, int pad_mode, const float* __restrict__ red,
                      TView dx, TView dres, int has_dres, int dres_acc, float inv_hw, float* __restrict__ dbias, int S) {
  extern __shared__ float sm[];     // A[c] | B[c] | C[c] | D[c] | partials[256/G][c] (8 KB) | ring[S][P][nt][256] x 16 B
  const int nn = blockIdx.y, c = x.c, G = c / 8;
  const bool racc = has_dres && dres_acc;
  const bool fold = dy.pad > 0 && pad_mode == NEMAR_PAD_REFLECT;
  // ring entries per pixel: x, dy, [old dres], [first mirrored tap of dy: border pixels fetch it through the ring too —
  // loaded synchronously, it made the CTAs that own the border rows the stragglers of the launch]
  const int ne = racc ? 3 : 2;
  const int nt = ne + (fold ? 1 : 0);
  float* sA = sm; float* sB = sm + c; float* sC = sm + 2 * c; float* sD = sm + 3 * c;
  float* spart = sm + 4 * c;
  uint4* ring = reinterpret_cast<uint4*>(sm + 4 * c + 2048) + threadIdx.x;
  const Range r = block_range((uint32_t)x.h * x.w, G);
  const int c0 = r.cg * 8;
  const __nv_bfloat16* xb = sample_base(x, nn, c0);
  const __nv_bfloat16* db = sample_base(dy, nn, c0);
  __nv_bfloat16* ob = const_cast<__nv_bfloat16*>(sample_base(dx, nn, c0));
  __nv_bfloat16* rb = const_cast<__nv_bfloat16*>(sample_base(dres, nn, c0));
  const uint32_t uw = (uint32_t)x.w;
  const float inv_w = 1.f / (float)x.w;
  const uint32_t first = r.lo + r.pl, step = P * r.ppi;
  const uint32_t niter = (r.hi - r.lo + step - 1) / step;
  auto issue = [&](uint32_t it, int slot) {
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const uint32_t p = first + it * step + k * r.ppi;
      if (it < niter && p < r.hi) {
        int yy, xx;
        divmod(p, uw, inv_w, yy, xx);
        uint4* dst = ring + ((slot * P + k) * nt) * 256;
        cp_async16(dst, xb + off_in(x, yy, xx));
        cp_async16(dst + 256, db + off_in(dy, yy, xx));
        if (racc) cp_async16(dst + 512, rb + off_in(dres, yy, xx));
        int yp, xp;
        if (fold && near_border(yy, xx, dy.h, dy.w, dy.pad) && first_mirror(dy, yy, xx, yp, xp))
          cp_async16(dst + ne * 256, db + off_pad(dy, yp, xp));
      }
    }
    cp_async_commit();
  };
  for (int s = 0; s < S - 1; ++s) issue((uint32_t)s, s);
  fill_ab(stats, nn, c, inv_hw, sA, sB);
  for (int k = threadIdx.x; k < c; k += blockDim.x) {
    float cc = 0.f, dd = 0.f;
    if (stats) {
      const float a = sA[k];      // written by this same thread in fill_ab (same k -> thread mapping)
      cc = -a * __ldg(red + ((size_t)nn * c + k) * 2) * inv_hw;
      dd = -a * __ldg(red + ((size_t)nn * c + k) * 2 + 1) * inv_hw;
    }
    sC[k] = cc; sD[k] = dd;
  }
  __syncthreads();
  float bsum[8], cA[8], cB[8], cC[8], cD[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) bsum[k] = 0.f;
  coef8(sA, c0, cA); coef8(sB, c0, cB); coef8(sC, c0, cC); coef8(sD, c0, cD);
  int slot = 0, islot = S - 1;
  for (uint32_t it = 0; it < niter; ++it) {
    issue(it + (uint32_t)(S - 1), islot);
    if (++islot == S) islot = 0;
    cp_async_wait_pending(S - 1);
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const uint32_t p = first + it * step + k * r.ppi;
      if (p >= r.hi) continue;
      int yy, xx;
      divmod(p, uw, inv_w, yy, xx);
      const uint4* src = ring + ((slot * P + k) * nt) * 256;
      float g[8], v[8];
      unpack8(src[256], g);
      if (fold && near_border(yy, xx, dy.h, dy.w, dy.pad)) fold_extra(dy, nn, yy, xx, c0, g, src + ne * 256);
      if (has_dres) {
        float t[8];
        if (racc) {
          unpack8(src[512], t);
#pragma unroll
          for (int j = 0; j < 8; ++j) t[j] += g[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) t[j] = g[j];
        }
        *reinterpret_cast<uint4*>(rb + off_in(dres, yy, xx)) = pack8(t);
      }
      unpack8(src[0], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = fmaf(v[j], cA[j], cB[j]);
        const float gg = g[j] * actg<ACT>(xh, act);
        const float o = fmaf(gg, cA[j], fmaf(xh, cD[j], cC[j]));
        v[j] = o;
        bsum[j] += o;
      }
      *reinterpret_cast<uint4*>(ob + off_in(dx, yy, xx)) = pack8(v);
    }
    if (++slot == S) slot = 0;
  }
  if (dbias) {
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

// ring depth per pass (kind 0: forward, 1: statistics, 2: backward reduce, 3: backward apply).  NEMAR_LEAN_PIPE sets
// all four, NEMAR_LEAN_PIPE_<FWD|STATS|RED|APPLY> one (tuning knobs of scripts/nbench.py).
static inline int pipe_stages(int kind) {
  static const int all = [] { const char* e = getenv("NEMAR_LEAN_PIPE"); return e ? atoi(e) : -1; }();
  static const int per[4] = {
      [] { const char* e = getenv("NEMAR_LEAN_PIPE_FWD"); return e ? atoi(e) : -1; }(),
      [] { const char* e = getenv("NEMAR_LEAN_PIPE_STATS"); return e ? atoi(e) : -1; }(),
      [] { const char* e = getenv("NEMAR_LEAN_PIPE_RED"); return e ? atoi(e) : -1; }(),
      [] { const char* e = getenv("NEMAR_LEAN_PIPE_APPLY"); return e ? atoi(e) : -1; }()};
  static const int dflt[4] = {4, 4, 3, 3};      // measured: profiles/r02_nbench_norm_pipe.txt
  int v = per[kind] >= 0 ? per[kind] : (all >= 0 ? all : dflt[kind]);
  if (v > MAX_STAGES) v = MAX_STAGES;
  return v < 2 ? 2 : v;
}
static inline size_t ring_bytes(int stages, int nt) { return (size_t)stages * P * nt * 256 * 16; }
// one wave of resident CTAs over the whole batch: `smem` dynamic bytes and `max_ctas` (register bound) per CTA decide
// the residency; each CTA then streams its pixel range through the ring in >= 2 iterations
static inline int pipe_chunks(int64_t npix, int G, int n, size_t smem, int max_ctas) {
  static const int waves = [] { const char* e = getenv("NEMAR_LEAN_WAVES"); return e ? atoi(e) : 1; }();
  int occ = (int)(233472 / (smem + 1024));
  if (occ > max_ctas) occ = max_ctas;
  if (occ < 1) occ = 1;
  const int64_t ppi = 256 / G;
  int64_t chunks = (npix + ppi * P * 2 - 1) / (ppi * P * 2);
  int64_t cap = ((int64_t)148 * occ * (waves < 1 ? 1 : waves)) / n;
  if (cap < 1) cap = 1;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  return (int)chunks;
}
// kernels whose ring exceeds the 48 KB default need the opt-in, once per instantiation (forward and backward run on
// different host threads under autograd)
static inline void allow_smem(const void* kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return;
  static std::mutex mu;
  static std::unordered_set<const void*> done;
  std::lock_guard<std::mutex> lk(mu);
  if (done.insert(kernel).second) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

// instantiate `...` with the compile-time activation A (A = -1: run-time switch, e.g. tanh)
#define NLEAN_ACT_SWITCH(act, ...)                                                  \
  switch (act) {                                                                    \
    case NEMAR_ACT_NONE: { constexpr int A = NEMAR_ACT_NONE; __VA_ARGS__; } break;   \
    case NEMAR_ACT_RELU: { constexpr int A = NEMAR_ACT_RELU; __VA_ARGS__; } break;   \
    case NEMAR_ACT_LRELU: { constexpr int A = NEMAR_ACT_LRELU; __VA_ARGS__; } break; \
    default: { constexpr int A = -1; __VA_ARGS__; } break;                           \
  }

}  // namespace nlean
