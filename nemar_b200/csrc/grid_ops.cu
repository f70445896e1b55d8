// grid_ops.cu — STN head: sampling-grid generation, bilinear grid_sample fwd / scatter-add bwd, and the
// (bilateral) smoothness term.  All HBM-bound: one coalesced pass, 128-bit grid/output accesses where the
// layout allows, warp-shuffle + one atomic per block for scalar reductions.
//
// Reference call sites: models/stn/affine_stn.py:105,128-130; models/stn/unet_stn.py:121-129,167,173-174;
// models/stn/stn_losses.py:4-30.  grid_sample arithmetic follows ATen/native/GridSampler.h:26-36,205-207
// (unnormalise ((g+1)*size-1)/2, floor taps, zeros padding, align_corners=False).
#include "common.cuh"
#include <cstdlib>

// ---------------------------------------------------------------------------------------------
// affine grid
// ---------------------------------------------------------------------------------------------
__global__ void affine_grid_fwd_kernel(const float* __restrict__ theta, const float* __restrict__ bx,
                                       const float* __restrict__ by, int n, int h, int w,
                                       float* __restrict__ grid) {
  // one thread -> two horizontally adjacent grid points (float4 store); w may be odd -> scalar tail
  const int wp = (w + 1) >> 1;
  const int64_t total = (int64_t)n * h * wp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int xp = (int)(i % wp);
    int64_t r = i / wp;
    int y = (int)(r % h);
    int nn = (int)(r / h);
    const float* t = theta + nn * 6;
    float t0 = __ldg(t + 0), t1 = __ldg(t + 1), t2 = __ldg(t + 2);
    float t3 = __ldg(t + 3), t4 = __ldg(t + 4), t5 = __ldg(t + 5);
    float yb = __ldg(by + y);
    int x0 = xp * 2;
    float xa = __ldg(bx + x0);
    float gx0 = fmaf(yb, t1, xa * t0) + t2;
    float gy0 = fmaf(yb, t4, xa * t3) + t5;
    float* g = grid + (((int64_t)nn * h + y) * w + x0) * 2;
    if (x0 + 1 < w) {
      float xb = __ldg(bx + x0 + 1);
      float gx1 = fmaf(yb, t1, xb * t0) + t2;
      float gy1 = fmaf(yb, t4, xb * t3) + t5;
      if ((((uintptr_t)g) & 15) == 0) {
        *reinterpret_cast<float4*>(g) = make_float4(gx0, gy0, gx1, gy1);
      } else {
        g[0] = gx0; g[1] = gy0; g[2] = gx1; g[3] = gy1;
      }
    } else {
      g[0] = gx0; g[1] = gy0;
    }
  }
}

__global__ void affine_grid_bwd_kernel(const float* __restrict__ dgrid, const float* __restrict__ bx,
                                       const float* __restrict__ by, int h, int w,
                                       float* __restrict__ dtheta) {
  // grid: (chunks, n).  Each block reduces a chunk of one sample's h*w points into 6 partial sums.
  __shared__ float red[32];
  const int nn = blockIdx.y;
  const int64_t hw = (int64_t)h * w;
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
  const float2* dg = reinterpret_cast<const float2*>(dgrid) + nn * hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw;
       i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % w), y = (int)(i / w);
    float2 d = __ldg(dg + i);
    float xb = __ldg(bx + x), yb = __ldg(by + y);
    a0 = fmaf(d.x, xb, a0); a1 = fmaf(d.x, yb, a1); a2 += d.x;
    a3 = fmaf(d.y, xb, a3); a4 = fmaf(d.y, yb, a4); a5 += d.y;
  }
  float v[6] = {a0, a1, a2, a3, a4, a5};
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    float s = block_sum(v[k], red);
    if (threadIdx.x == 0) atomicAdd(dtheta + nn * 6 + k, s);
  }
}

NEMAR_API int nemar_affine_grid_fwd(const float* theta, const float* bx, const float* by, int n, int h,
                                    int w, float* grid, void* stream) {
  NEMAR_REQUIRE(theta && bx && by && grid && n > 0 && h > 0 && w > 0, "affine_grid_fwd: bad args");
  int64_t total = (int64_t)n * h * ((w + 1) / 2);
  affine_grid_fwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(theta, bx, by, n, h, w,
                                                                                   grid);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

NEMAR_API int nemar_affine_grid_bwd(const float* dgrid, const float* bx, const float* by, int n, int h,
                                    int w, float* dtheta, void* stream) {
  NEMAR_REQUIRE(dgrid && bx && by && dtheta && n > 0 && h > 0 && w > 0, "affine_grid_bwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(dtheta, 0, sizeof(float) * 6 * n, s);
  int chunks = grid_for((int64_t)h * w, 256, 64);
  affine_grid_bwd_kernel<<<dim3(chunks, n), 256, 0, s>>>(dgrid, bx, by, h, w, dtheta);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// flow grid: grid[n,y,x,:] = (xs[x] + off_x, ys[y] + off_y)
// ---------------------------------------------------------------------------------------------
__global__ void flow_grid_fwd_kernel(const float* __restrict__ off, int64_t sn, int64_t sc, int64_t sy,
                                     int64_t sx, const float* __restrict__ xs,
                                     const float* __restrict__ ys, int n, int h, int w,
                                     float* __restrict__ grid) {
  const int64_t total = (int64_t)n * h * w;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % w);
    int64_t r = i / w;
    int y = (int)(r % h);
    int nn = (int)(r / h);
    const float* o = off + nn * sn + y * sy + x * sx;
    float ox, oy;
    if (sc == 1) {  // channels-last offsets: one 8-byte load
      float2 v = __ldg(reinterpret_cast<const float2*>(o));
      ox = v.x; oy = v.y;
    } else {
      ox = __ldg(o); oy = __ldg(o + sc);
    }
    float2 g = make_float2(__ldg(xs + x) + ox, __ldg(ys + y) + oy);
    reinterpret_cast<float2*>(grid)[i] = g;
  }
}

NEMAR_API int nemar_flow_grid_fwd(const float* off, int64_t sn, int64_t sc, int64_t sy, int64_t sx,
                                  const float* xs, const float* ys, int n, int h, int w, float* grid,
                                  void* stream) {
  NEMAR_REQUIRE(off && xs && ys && grid && n > 0 && h > 0 && w > 0, "flow_grid_fwd: bad args");
  if (sc == 1)
    NEMAR_REQUIRE((((uintptr_t)off) & 7) == 0 && sx % 2 == 0 && sy % 2 == 0 && sn % 2 == 0,
                  "flow_grid_fwd: channels-last offsets must be 8-byte aligned");
  int64_t total = (int64_t)n * h * w;
  flow_grid_fwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(off, sn, sc, sy, sx, xs,
                                                                                 ys, n, h, w, grid);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// grid_sample (bilinear, zeros, align_corners=False)
// ---------------------------------------------------------------------------------------------
struct Taps {
  int x0, y0;          // floor coordinates (the "integer sampling indices")
  float nw, ne, sw, se;
  float ix, iy;
};

__device__ __forceinline__ Taps make_taps(float gx, float gy, int w, int h) {
  Taps t;
  // ATen grid_sampler_unnormalize, align_corners=False: ((coord + 1) * size - 1) / 2.  Both ATen builds evaluate the
  // product and the subtraction as ONE fused multiply-add (the vectorised CPU kernel as fma(coord + 1, size / 2, -0.5),
  // the CUDA kernel through nvcc's contraction) — measured: the fused form reproduces F.grid_sample on CPU to 1.2e-7 at
  // 288 x 384, the two-rounding form only to 1.4e-5 (the forms coincide for power-of-two sizes).  Written as an
  // explicit fma so that the INTEGER tap index does not depend on a compiler's contraction choices.
  t.ix = __fmaf_rn(__fadd_rn(gx, 1.f), (float)w * 0.5f, -0.5f);
  t.iy = __fmaf_rn(__fadd_rn(gy, 1.f), (float)h * 0.5f, -0.5f);
  float fx = floorf(t.ix), fy = floorf(t.iy);
  t.x0 = (int)fx;
  t.y0 = (int)fy;
  float x1 = fx + 1.f, y1 = fy + 1.f;
  t.nw = (x1 - t.ix) * (y1 - t.iy);
  t.ne = (t.ix - fx) * (y1 - t.iy);
  t.sw = (x1 - t.ix) * (t.iy - fy);
  t.se = (t.ix - fx) * (t.iy - fy);
  return t;
}

// Each thread owns PX horizontally adjacent output points of one row: the grid is read with 128-bit
// loads and each channel of each image is written with one vector store.
template <int PX>
__global__ void __launch_bounds__(256)
grid_sample_fwd_kernel(const float* __restrict__ img0, const float* __restrict__ img1, int nimg, int n,
                       int c, int h, int w, const float* __restrict__ grid, int ho, int wo,
                       float* __restrict__ out0, float* __restrict__ out1, int32_t* __restrict__ idx) {
  const int wq = wo / PX;  // host guarantees wo % PX == 0
  const int64_t total = (int64_t)n * ho * wq;
  const int64_t ihw = (int64_t)h * w, ohw = (int64_t)ho * wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int xq = (int)(i % wq);
    int64_t r = i / wq;
    int y = (int)(r % ho);
    int nn = (int)(r / ho);
    const int64_t opix = (int64_t)y * wo + (int64_t)xq * PX;
    const float* gp = grid + ((int64_t)nn * ohw + opix) * 2;
    float g[2 * PX];
    if constexpr (PX == 2) {
      float4 v = __ldg(reinterpret_cast<const float4*>(gp));
      g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
    } else if constexpr (PX == 4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(gp));
      float4 u = __ldg(reinterpret_cast<const float4*>(gp) + 1);
      g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
      g[4] = u.x; g[5] = u.y; g[6] = u.z; g[7] = u.w;
    } else {
      g[0] = __ldg(gp); g[1] = __ldg(gp + 1);
    }
    Taps t[PX];
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      t[p] = make_taps(g[2 * p], g[2 * p + 1], w, h);
      if (idx) {
        int64_t o = ((int64_t)nn * ohw + opix + p) * 2;
        idx[o] = t[p].x0;
        idx[o + 1] = t[p].y0;
      }
    }
    for (int im = 0; im < nimg; ++im) {
      const float* img = (im == 0 ? img0 : img1) + (int64_t)nn * c * ihw;
      float* out = (im == 0 ? out0 : out1) + (int64_t)nn * c * ohw + opix;
      for (int ch = 0; ch < c; ++ch) {
        const float* pl = img + ch * ihw;
        float res[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          const int x0 = t[p].x0, y0 = t[p].y0;
          const bool xin0 = (x0 >= 0) & (x0 < w), xin1 = (x0 + 1 >= 0) & (x0 + 1 < w);
          const bool yin0 = (y0 >= 0) & (y0 < h), yin1 = (y0 + 1 >= 0) & (y0 + 1 < h);
          float acc = 0.f;
          if (yin0) {
            const float* row = pl + (int64_t)y0 * w;
            if (xin0) acc += __ldg(row + x0) * t[p].nw;
            if (xin1) acc += __ldg(row + x0 + 1) * t[p].ne;
          }
          if (yin1) {
            const float* row = pl + (int64_t)(y0 + 1) * w;
            if (xin0) acc += __ldg(row + x0) * t[p].sw;
            if (xin1) acc += __ldg(row + x0 + 1) * t[p].se;
          }
          res[p] = acc;
        }
        float* op = out + ch * ohw;
        if constexpr (PX == 4) {
          *reinterpret_cast<float4*>(op) = make_float4(res[0], res[1], res[2], res[3]);
        } else if constexpr (PX == 2) {
          *reinterpret_cast<float2*>(op) = make_float2(res[0], res[1]);
        } else {
          op[0] = res[0];
        }
      }
    }
  }
}

// One output point per lane: for a smooth field the east taps of lane L are the west taps of lane L+1, so each lane
// fetches only its west column (nw, sw) and receives the east column (ne, se) from its right neighbour by shuffle
// (falling back to its own loads when the neighbour samples elsewhere).  Channel and image loops are unrolled at
// compile time so that all west-column loads of a point are in flight together (the kernel is latency-bound, not
// LSU-bound: measured), and the index arithmetic is 32-bit.
template <int C, int NIMG, int P>
__global__ void __launch_bounds__(256)
grid_sample_fwd_shared_kernel(const float* __restrict__ img0, const float* __restrict__ img1, int n, int h, int w,
                              const float* __restrict__ grid, int ho, int wo, float* __restrict__ out0,
                              float* __restrict__ out1, int32_t* __restrict__ idx) {
  const uint32_t ihw = (uint32_t)h * w, ohw = (uint32_t)ho * wo;
  const uint32_t total = (uint32_t)n * ohw;                 // host guarantees < 2^31
  const uint32_t total_r = (total + 31u) & ~31u;
  const uint32_t step = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  // P independent points per thread per iteration (warp-uniform trip count): all their loads are in flight together
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total_r; i0 += P * step) {
    uint32_t ii[P], nn[P], opix[P];
    bool active[P], live[P];
    float2 g[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const uint32_t i = i0 + p * step;
      live[p] = i < total_r;                                  // warp-uniform
      active[p] = i < total;
      ii[p] = active[p] ? i : total - 1;
      g[p] = __ldg(reinterpret_cast<const float2*>(grid) + ii[p]);
    }
    Taps t[P];
    bool borrow[P], xin0[P], xin1[P], yin0[P], yin1[P];
    int o_nw[P];
    float vnw[P][NIMG][C], vsw[P][NIMG][C], vne[P][NIMG][C], vse[P][NIMG][C];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      nn[p] = ii[p] / ohw;
      opix[p] = ii[p] - nn[p] * ohw;
      t[p] = make_taps(g[p].x, g[p].y, w, h);
      if (idx && active[p]) { idx[(size_t)ii[p] * 2] = t[p].x0; idx[(size_t)ii[p] * 2 + 1] = t[p].y0; }
      const int x0 = t[p].x0, y0 = t[p].y0;
      xin0[p] = (x0 >= 0) & (x0 < w); xin1[p] = (x0 + 1 >= 0) & (x0 + 1 < w);
      yin0[p] = (y0 >= 0) & (y0 < h); yin1[p] = (y0 + 1 >= 0) & (y0 + 1 < h);
      o_nw[p] = y0 * w + x0;
      const size_t ibase = (size_t)nn[p] * C * ihw;
#pragma unroll
      for (int im = 0; im < NIMG; ++im) {
        const float* img = (im == 0 ? img0 : img1) + ibase;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
          vnw[p][im][ch] = (xin0[p] && yin0[p]) ? __ldg(img + (size_t)ch * ihw + o_nw[p]) : 0.f;
          vsw[p][im][ch] = (xin0[p] && yin1[p]) ? __ldg(img + (size_t)ch * ihw + o_nw[p] + w) : 0.f;
        }
      }
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
      if (!live[p]) continue;                                 // warp-uniform: safe around the shuffles
      const int nx0 = __shfl_down_sync(0xffffffffu, t[p].x0, 1);
      const int ny0 = __shfl_down_sync(0xffffffffu, t[p].y0, 1);
      const uint32_t nnn = __shfl_down_sync(0xffffffffu, nn[p], 1);
      borrow[p] = lane < 31 && nnn == nn[p] && ny0 == t[p].y0 && nx0 == t[p].x0 + 1;
#pragma unroll
      for (int im = 0; im < NIMG; ++im)
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
          vne[p][im][ch] = __shfl_down_sync(0xffffffffu, vnw[p][im][ch], 1);
          vse[p][im][ch] = __shfl_down_sync(0xffffffffu, vsw[p][im][ch], 1);
        }
      if (!borrow[p]) {
        const size_t ibase = (size_t)nn[p] * C * ihw;
#pragma unroll
        for (int im = 0; im < NIMG; ++im) {
          const float* img = (im == 0 ? img0 : img1) + ibase;
#pragma unroll
          for (int ch = 0; ch < C; ++ch) {
            vne[p][im][ch] = (xin1[p] && yin0[p]) ? __ldg(img + (size_t)ch * ihw + o_nw[p] + 1) : 0.f;
            vse[p][im][ch] = (xin1[p] && yin1[p]) ? __ldg(img + (size_t)ch * ihw + o_nw[p] + w + 1) : 0.f;
          }
        }
      }
      if (active[p]) {
#pragma unroll
        for (int im = 0; im < NIMG; ++im) {
          float* out = (im == 0 ? out0 : out1) + (size_t)nn[p] * C * ohw + opix[p];
#pragma unroll
          for (int ch = 0; ch < C; ++ch) {
            // same accumulation order as ATen's kernel: nw, ne, sw, se (out-of-range taps hold 0 and add exactly 0)
            float acc = vnw[p][im][ch] * t[p].nw;
            acc += vne[p][im][ch] * t[p].ne;
            acc += vsw[p][im][ch] * t[p].sw;
            acc += vse[p][im][ch] * t[p].se;
            out[(size_t)ch * ohw] = acc;
          }
        }
      }
    }
  }
}

NEMAR_API int nemar_grid_sample_fwd(const float* img0, const float* img1, int nimg, int n, int c, int h,
                                    int w, const float* grid, int ho, int wo, float* out0, float* out1,
                                    int32_t* idx_dump, void* stream) {
  NEMAR_REQUIRE(img0 && grid && out0 && (nimg == 1 || (nimg == 2 && img1 && out1)),
                "grid_sample_fwd: bad pointers");
  NEMAR_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0 && ho > 0 && wo > 0, "grid_sample_fwd: bad dims");
  cudaStream_t s = (cudaStream_t)stream;
  bool al16 = ((((uintptr_t)grid) | ((uintptr_t)out0) | ((uintptr_t)(nimg == 2 ? out1 : out0))) & 15) == 0;
  // (a shared-memory tiled variant — 32x16 output tiles, window of the taps' bounding box staged with coalesced row
  //  loads, scatter-add accumulated in a shared window — was measured SLOWER on B200 at 1024^2: fwd 82 vs 63 us, bwd 223 vs
  //  203 us; the neighbour-lane tap sharing below already removes half of the gathers / REDs.  Removed in round 2.)
  static const int variant = [] { const char* e = getenv("NEMAR_GS_VARIANT"); return e ? atoi(e) : 1; }();
  const int64_t tot = (int64_t)n * ho * wo;
  if (variant >= 1 && c == 3 && (((uintptr_t)grid) & 7) == 0 && tot < (1ll << 31) && (int64_t)h * w < (1ll << 30)) {
    const int blocks = grid_for(tot, 256, 148 * 16);   // P = 2 points per thread was measured slower (61 -> 70 us)
    if (nimg == 2)
      grid_sample_fwd_shared_kernel<3, 2, 1><<<blocks, 256, 0, s>>>(img0, img1, n, h, w, grid, ho, wo, out0, out1, idx_dump);
    else
      grid_sample_fwd_shared_kernel<3, 1, 1><<<blocks, 256, 0, s>>>(img0, img1, n, h, w, grid, ho, wo, out0, out1, idx_dump);
  } else if (wo % 4 == 0 && al16) {
    int64_t total = (int64_t)n * ho * (wo / 4);
    grid_sample_fwd_kernel<4><<<grid_for(total, 256), 256, 0, s>>>(img0, img1, nimg, n, c, h, w, grid, ho,
                                                                    wo, out0, out1, idx_dump);
  } else if (wo % 2 == 0 && al16) {
    int64_t total = (int64_t)n * ho * (wo / 2);
    grid_sample_fwd_kernel<2><<<grid_for(total, 256), 256, 0, s>>>(img0, img1, nimg, n, c, h, w, grid, ho,
                                                                    wo, out0, out1, idx_dump);
  } else {
    int64_t total = (int64_t)n * ho * wo;
    grid_sample_fwd_kernel<1><<<grid_for(total, 256), 256, 0, s>>>(img0, img1, nimg, n, c, h, w, grid, ho,
                                                                    wo, out0, out1, idx_dump);
  }
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// Backward.  One thread per output point.  grad_grid follows ATen's CUDA kernel term by term; the
// grad_input scatter uses warp-level merging of the horizontally shared taps: for a smooth field the
// east taps of point x and the west taps of point x+1 hit the same address, so lane x+1 adopts lane
// x's east contribution through a shuffle and only one RED is issued for the pair.
__global__ void __launch_bounds__(256)
grid_sample_bwd_kernel(const float* __restrict__ img0, const float* __restrict__ img1, int nimg, int n,
                       int c, int h, int w, const float* __restrict__ grid, int ho, int wo,
                       const float* __restrict__ dout0, const float* __restrict__ dout1,
                       float* __restrict__ dimg0, float* __restrict__ dimg1, float* __restrict__ dgrid) {
  const int64_t ihw = (int64_t)h * w, ohw = (int64_t)ho * wo;
  const int64_t total = (int64_t)n * ohw;
  const int64_t total_r = (total + 31) / 32 * 32;  // keep whole warps in the loop for the shuffles
  const int lane = threadIdx.x & 31;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_r;
       i += (int64_t)gridDim.x * blockDim.x) {
    const bool active = i < total;
    int64_t ii = active ? i : total - 1;
    int64_t opix = ii % ohw;
    int nn = (int)(ii / ohw);
    float2 g = __ldg(reinterpret_cast<const float2*>(grid) + ii);
    Taps t = make_taps(g.x, g.y, w, h);
    const int x0 = t.x0, y0 = t.y0;
    const bool xin0 = (x0 >= 0) & (x0 < w), xin1 = (x0 + 1 >= 0) & (x0 + 1 < w);
    const bool yin0 = (y0 >= 0) & (y0 < h), yin1 = (y0 + 1 >= 0) & (y0 + 1 < h);
    const float fx = (float)x0, fy = (float)y0;
    const float x1 = fx + 1.f, y1 = fy + 1.f;
    // does my west column coincide with the previous lane's east column (same sample, same rows)?
    int px0 = __shfl_up_sync(0xffffffffu, x0, 1);
    int py0 = __shfl_up_sync(0xffffffffu, y0, 1);
    int pn = __shfl_up_sync(0xffffffffu, nn, 1);
    int pact = __shfl_up_sync(0xffffffffu, (int)active, 1);
    const bool adopt = active && lane > 0 && pact && pn == nn && py0 == y0 && px0 + 1 == x0;
    // will the next lane adopt my east column?
    const unsigned adopt_mask = __ballot_sync(0xffffffffu, adopt);
    const bool donated = (lane < 31) && ((adopt_mask >> (lane + 1)) & 1u);

    float gix = 0.f, giy = 0.f;
    for (int im = 0; im < nimg; ++im) {
      const float* img = (im == 0 ? img0 : img1) + (int64_t)nn * c * ihw;
      const float* dout = (im == 0 ? dout0 : dout1) + (int64_t)nn * c * ohw + opix;
      float* dimg = (im == 0 ? dimg0 : dimg1);
      if (dimg) dimg += (int64_t)nn * c * ihw;
      for (int ch = 0; ch < c; ++ch) {
        const float go = active ? __ldg(dout + ch * ohw) : 0.f;
        const float* pl = img + ch * ihw;
        float vnw = 0.f, vne = 0.f, vsw = 0.f, vse = 0.f;
        if (yin0) {
          const float* row = pl + (int64_t)y0 * w;
          if (xin0) vnw = __ldg(row + x0);
          if (xin1) vne = __ldg(row + x0 + 1);
        }
        if (yin1) {
          const float* row = pl + (int64_t)(y0 + 1) * w;
          if (xin0) vsw = __ldg(row + x0);
          if (xin1) vse = __ldg(row + x0 + 1);
        }
        gix -= vnw * (y1 - t.iy) * go;
        giy -= vnw * (x1 - t.ix) * go;
        gix += vne * (y1 - t.iy) * go;
        giy -= vne * (t.ix - fx) * go;
        gix -= vsw * (t.iy - fy) * go;
        giy += vsw * (x1 - t.ix) * go;
        gix += vse * (t.iy - fy) * go;
        giy += vse * (t.ix - fx) * go;
        if (dimg) {  // warp-uniform branch (dimg is a kernel argument)
          float cnw = t.nw * go, cne = t.ne * go, csw = t.sw * go, cse = t.se * go;
          float pne = __shfl_up_sync(0xffffffffu, cne, 1);
          float pse = __shfl_up_sync(0xffffffffu, cse, 1);
          if (adopt) { cnw += pne; csw += pse; }
          float* dp = dimg + ch * ihw;
          if (active) {
            if (yin0) {
              float* row = dp + (int64_t)y0 * w;
              if (xin0) atomicAdd(row + x0, cnw);
              if (xin1 && !donated) atomicAdd(row + x0 + 1, cne);
            }
            if (yin1) {
              float* row = dp + (int64_t)(y0 + 1) * w;
              if (xin0) atomicAdd(row + x0, csw);
              if (xin1 && !donated) atomicAdd(row + x0 + 1, cse);
            }
          }
        }
      }
    }
    if (active) {
      // ATen: grad multipliers size/2 for align_corners=False
      float2 o = make_float2(((float)w / 2.f) * gix, ((float)h / 2.f) * giy);
      reinterpret_cast<float2*>(dgrid)[ii] = o;
    }
  }
}

// Backward with the same one-point-per-lane layout: west-column values are loaded once and the east column is taken
// from the right neighbour (needed for grad_grid); the scatter-add merges east contributions into the neighbour's
// west REDs.  Channel / image loops are compile-time, indices 32-bit.
template <int C, int NIMG>
__global__ void __launch_bounds__(256)
grid_sample_bwd_shared_kernel(const float* __restrict__ img0, const float* __restrict__ img1, int n, int h, int w,
                              const float* __restrict__ grid, int ho, int wo, const float* __restrict__ dout0,
                              const float* __restrict__ dout1, float* __restrict__ dimg0, float* __restrict__ dimg1,
                              float* __restrict__ dgrid) {
  const uint32_t ihw = (uint32_t)h * w, ohw = (uint32_t)ho * wo;
  const uint32_t total = (uint32_t)n * ohw;
  const uint32_t total_r = (total + 31u) & ~31u;
  const int lane = threadIdx.x & 31;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total_r; i += gridDim.x * blockDim.x) {
    const bool active = i < total;
    const uint32_t ii = active ? i : total - 1;
    const uint32_t nn = ii / ohw;
    const uint32_t opix = ii - nn * ohw;
    const float2 g = __ldg(reinterpret_cast<const float2*>(grid) + ii);
    const Taps t = make_taps(g.x, g.y, w, h);
    const int x0 = t.x0, y0 = t.y0;
    const bool xin0 = (x0 >= 0) & (x0 < w), xin1 = (x0 + 1 >= 0) & (x0 + 1 < w);
    const bool yin0 = (y0 >= 0) & (y0 < h), yin1 = (y0 + 1 >= 0) & (y0 + 1 < h);
    const float fx = (float)x0, fy = (float)y0, x1 = fx + 1.f, y1 = fy + 1.f;
    const size_t ibase = (size_t)nn * C * ihw;
    const int o_nw = y0 * w + x0;
    float vnw[NIMG][C], vsw[NIMG][C], vne[NIMG][C], vse[NIMG][C], go[NIMG][C];
#pragma unroll
    for (int im = 0; im < NIMG; ++im) {
      const float* img = (im == 0 ? img0 : img1) + ibase;
      const float* dout = (im == 0 ? dout0 : dout1) + (size_t)nn * C * ohw + opix;
#pragma unroll
      for (int ch = 0; ch < C; ++ch) {
        go[im][ch] = active ? __ldg(dout + (size_t)ch * ohw) : 0.f;
        vnw[im][ch] = (xin0 && yin0) ? __ldg(img + (size_t)ch * ihw + o_nw) : 0.f;
        vsw[im][ch] = (xin0 && yin1) ? __ldg(img + (size_t)ch * ihw + o_nw + w) : 0.f;
      }
    }
    // right neighbour: source of my east values; left neighbour: adopts my east contributions
    const int nx0 = __shfl_down_sync(0xffffffffu, x0, 1), ny0 = __shfl_down_sync(0xffffffffu, y0, 1);
    const uint32_t nnn = __shfl_down_sync(0xffffffffu, nn, 1);
    const int nact = __shfl_down_sync(0xffffffffu, (int)active, 1);
    const bool borrow = lane < 31 && nnn == nn && ny0 == y0 && nx0 == x0 + 1;     // values (any neighbour, even inactive)
    const bool donated = borrow && active && nact;                                // my east REDs are issued by lane+1
    const int px0 = __shfl_up_sync(0xffffffffu, x0, 1), py0 = __shfl_up_sync(0xffffffffu, y0, 1);
    const uint32_t pnn = __shfl_up_sync(0xffffffffu, nn, 1);
    const int pact = __shfl_up_sync(0xffffffffu, (int)active, 1);
    const bool adopt = active && lane > 0 && pact && pnn == nn && py0 == y0 && px0 + 1 == x0;
#pragma unroll
    for (int im = 0; im < NIMG; ++im)
#pragma unroll
      for (int ch = 0; ch < C; ++ch) {
        vne[im][ch] = __shfl_down_sync(0xffffffffu, vnw[im][ch], 1);
        vse[im][ch] = __shfl_down_sync(0xffffffffu, vsw[im][ch], 1);
      }
    if (!borrow) {
#pragma unroll
      for (int im = 0; im < NIMG; ++im) {
        const float* img = (im == 0 ? img0 : img1) + ibase;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
          vne[im][ch] = (xin1 && yin0) ? __ldg(img + (size_t)ch * ihw + o_nw + 1) : 0.f;
          vse[im][ch] = (xin1 && yin1) ? __ldg(img + (size_t)ch * ihw + o_nw + w + 1) : 0.f;
        }
      }
    }
    float gix = 0.f, giy = 0.f;
#pragma unroll
    for (int im = 0; im < NIMG; ++im) {
      float* dimg = (im == 0 ? dimg0 : dimg1);
#pragma unroll
      for (int ch = 0; ch < C; ++ch) {
        const float gg = go[im][ch];
        gix -= vnw[im][ch] * (y1 - t.iy) * gg;
        giy -= vnw[im][ch] * (x1 - t.ix) * gg;
        gix += vne[im][ch] * (y1 - t.iy) * gg;
        giy -= vne[im][ch] * (t.ix - fx) * gg;
        gix -= vsw[im][ch] * (t.iy - fy) * gg;
        giy += vsw[im][ch] * (x1 - t.ix) * gg;
        gix += vse[im][ch] * (t.iy - fy) * gg;
        giy += vse[im][ch] * (t.ix - fx) * gg;
        if (dimg) {   // kernel argument: warp-uniform
          float cnw = t.nw * gg, cne = t.ne * gg, csw = t.sw * gg, cse = t.se * gg;
          const float pne = __shfl_up_sync(0xffffffffu, cne, 1);
          const float pse = __shfl_up_sync(0xffffffffu, cse, 1);
          if (adopt) { cnw += pne; csw += pse; }
          if (active) {
            float* dp = dimg + ibase + (size_t)ch * ihw + o_nw;
            if (xin0 && yin0) atomicAdd(dp, cnw);
            if (xin0 && yin1) atomicAdd(dp + w, csw);
            if (!donated) {
              if (xin1 && yin0) atomicAdd(dp + 1, cne);
              if (xin1 && yin1) atomicAdd(dp + w + 1, cse);
            }
          }
        }
      }
    }
    if (active)
      reinterpret_cast<float2*>(dgrid)[ii] = make_float2(((float)w / 2.f) * gix, ((float)h / 2.f) * giy);
  }
}

NEMAR_API int nemar_grid_sample_bwd(const float* img0, const float* img1, int nimg, int n, int c, int h,
                                    int w, const float* grid, int ho, int wo, const float* dout0,
                                    const float* dout1, float* dimg0, float* dimg1, float* dgrid,
                                    void* stream) {
  NEMAR_REQUIRE(img0 && grid && dout0 && dgrid && (nimg == 1 || (nimg == 2 && img1 && dout1)),
                "grid_sample_bwd: bad pointers");
  NEMAR_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0 && ho > 0 && wo > 0, "grid_sample_bwd: bad dims");
  // the kernel treats dimg as warp-uniform per image; it handles one "has dimg" flag per image by
  // running images with/without gradient in the same loop (pointer may be NULL per image).
  int64_t total = (int64_t)n * ho * wo;
  static const int variant = [] { const char* e = getenv("NEMAR_GS_VARIANT"); return e ? atoi(e) : 1; }();
  if (variant >= 1 && c == 3 && total < (1ll << 31) && (int64_t)h * w < (1ll << 30)) {
    const int blocks = grid_for(total, 256, 148 * 16);
    if (nimg == 2)
      grid_sample_bwd_shared_kernel<3, 2><<<blocks, 256, 0, (cudaStream_t)stream>>>(img0, img1, n, h, w, grid, ho, wo, dout0,
                                                                                     dout1, dimg0, dimg1, dgrid);
    else
      grid_sample_bwd_shared_kernel<3, 1><<<blocks, 256, 0, (cudaStream_t)stream>>>(img0, img1, n, h, w, grid, ho, wo, dout0,
                                                                                     dout1, dimg0, dimg1, dgrid);
    NEMAR_LAUNCH_CHECK();
    return 0;
  }
  grid_sample_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      img0, img1, nimg, n, c, h, w, grid, ho, wo, dout0, dout1, dimg0, dimg1, dgrid);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// smoothness (+ bilateral weights)
// ---------------------------------------------------------------------------------------------
struct SmoothArgs {
  const float* def;
  int64_t sn, sc, sy, sx;
  const float* img;
  int img_c;
  float alpha;
  int n, h, w;
  float inv1, inv2, inv34;  // 1/count of each directional mean (already multiplied by `scale`)
};

__device__ __forceinline__ float bil_weight(const SmoothArgs& a, int nn, int ya, int xa, int yb, int xb) {
  // mean over image channels of exp(-alpha*|I(a) - I(b)|)
  const int64_t hw = (int64_t)a.h * a.w;
  const float* base = a.img + (int64_t)nn * a.img_c * hw;
  float s = 0.f;
  for (int ch = 0; ch < a.img_c; ++ch) {
    float va = __ldg(base + ch * hw + (int64_t)ya * a.w + xa);
    float vb = __ldg(base + ch * hw + (int64_t)yb * a.w + xb);
    s += expf(-a.alpha * fabsf(va - vb));
  }
  return s / (float)a.img_c;
}

__global__ void __launch_bounds__(256) smoothness_fwd_kernel(SmoothArgs a, float* __restrict__ loss) {
  __shared__ float red[32];
  const int64_t total = (int64_t)a.n * a.h * a.w;
  const bool bil = a.img != nullptr && a.alpha > 0.f;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % a.w);
    int64_t r = i / a.w;
    int y = (int)(r % a.h);
    int nn = (int)(r / a.h);
    const bool hasd = y + 1 < a.h, hasr = x + 1 < a.w;
    float w1 = 1.f, w2 = 1.f, w3 = 1.f, w4 = 1.f;
    if (bil) {
      if (hasd) w1 = bil_weight(a, nn, y + 1, x, y, x);
      if (hasr) w2 = bil_weight(a, nn, y, x + 1, y, x);
      if (hasd && hasr) {
        w3 = bil_weight(a, nn, y, x, y + 1, x + 1);
        w4 = bil_weight(a, nn, y, x + 1, y + 1, x);
      }
    }
    const float* p = a.def + nn * a.sn + y * a.sy + x * a.sx;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const float* q = p + ch * a.sc;
      float v00 = __ldg(q);
      float v10 = hasd ? __ldg(q + a.sy) : 0.f;
      float v01 = hasr ? __ldg(q + a.sx) : 0.f;
      float v11 = (hasd && hasr) ? __ldg(q + a.sy + a.sx) : 0.f;
      if (hasd) acc += a.inv1 * w1 * fabsf(v10 - v00);
      if (hasr) acc += a.inv2 * w2 * fabsf(v01 - v00);
      if (hasd && hasr) {
        acc += a.inv34 * w3 * fabsf(v00 - v11);
        acc += a.inv34 * w4 * fabsf(v01 - v10);
      }
    }
  }
  float s = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, s);
}

__device__ __forceinline__ float sgn(float v) { return (float)((v > 0.f) - (v < 0.f)); }

// Gather form of the adjoint: every field element collects the (up to) eight differences it takes part in.
__global__ void __launch_bounds__(256)
smoothness_bwd_kernel(SmoothArgs a, const float* __restrict__ gscale, float* __restrict__ ddef) {
  const int64_t total = (int64_t)a.n * a.h * a.w;
  const bool bil = a.img != nullptr && a.alpha > 0.f;
  const float gs = __ldg(gscale);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % a.w);
    int64_t r = i / a.w;
    int y = (int)(r % a.h);
    int nn = (int)(r / a.h);
    const bool up = y > 0, dn = y + 1 < a.h, lf = x > 0, rt = x + 1 < a.w;
    // weights of the differences touching (y,x); names: pairs (a)-(b) as in stn_losses.py
    float w1d = 1.f, w1u = 1.f, w2r = 1.f, w2l = 1.f, w3f = 1.f, w3b = 1.f, w4a = 1.f, w4b = 1.f;
    if (bil) {
      if (dn) w1d = bil_weight(a, nn, y + 1, x, y, x);            // diff_1 at (y,x): me is the minus term
      if (up) w1u = bil_weight(a, nn, y, x, y - 1, x);            // diff_1 at (y-1,x): me is the plus term
      if (rt) w2r = bil_weight(a, nn, y, x + 1, y, x);            // diff_2 at (y,x): minus
      if (lf) w2l = bil_weight(a, nn, y, x, y, x - 1);            // diff_2 at (y,x-1): plus
      if (dn && rt) w3f = bil_weight(a, nn, y, x, y + 1, x + 1);  // diff_3 at (y,x): plus
      if (up && lf) w3b = bil_weight(a, nn, y - 1, x - 1, y, x);  // diff_3 at (y-1,x-1): minus
      if (dn && lf) w4a = bil_weight(a, nn, y, x, y + 1, x - 1);  // diff_4 at (y,x-1): me=(y,x) is plus term
      if (up && rt) w4b = bil_weight(a, nn, y - 1, x + 1, y, x);  // diff_4 at (y-1,x): me is the minus term
    }
    const float* p = a.def + nn * a.sn + y * a.sy + x * a.sx;
    float* dp = ddef + nn * a.sn + y * a.sy + x * a.sx;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const float* q = p + ch * a.sc;
      float v = __ldg(q);
      float gacc = 0.f;
      if (dn) gacc -= a.inv1 * w1d * sgn(__ldg(q + a.sy) - v);
      if (up) gacc += a.inv1 * w1u * sgn(v - __ldg(q - a.sy));
      if (rt) gacc -= a.inv2 * w2r * sgn(__ldg(q + a.sx) - v);
      if (lf) gacc += a.inv2 * w2l * sgn(v - __ldg(q - a.sx));
      if (dn && rt) gacc += a.inv34 * w3f * sgn(v - __ldg(q + a.sy + a.sx));
      if (up && lf) gacc -= a.inv34 * w3b * sgn(__ldg(q - a.sy - a.sx) - v);
      if (dn && lf) gacc += a.inv34 * w4a * sgn(v - __ldg(q + a.sy - a.sx));
      if (up && rt) gacc -= a.inv34 * w4b * sgn(__ldg(q - a.sy + a.sx) - v);
      dp[ch * a.sc] += gs * gacc;
    }
  }
}

static int smooth_args(SmoothArgs& a, const float* def, int64_t sn, int64_t sc, int64_t sy, int64_t sx,
                       const float* img, int img_c, float alpha, int n, int h, int w, float scale) {
  NEMAR_REQUIRE(def && n > 0 && h > 1 && w > 1, "smoothness: bad args (need h,w >= 2)");
  NEMAR_REQUIRE(!(img && alpha > 0.f) || img_c > 0, "smoothness: img_c");
  a.def = def; a.sn = sn; a.sc = sc; a.sy = sy; a.sx = sx;
  a.img = (alpha > 0.f) ? img : nullptr;
  a.img_c = img_c; a.alpha = alpha; a.n = n; a.h = h; a.w = w;
  a.inv1 = scale / ((float)n * 2.f * (float)(h - 1) * (float)w);
  a.inv2 = scale / ((float)n * 2.f * (float)h * (float)(w - 1));
  a.inv34 = scale / ((float)n * 2.f * (float)(h - 1) * (float)(w - 1));
  return 0;
}

NEMAR_API int nemar_smoothness_fwd(const float* def, int64_t sn, int64_t sc, int64_t sy, int64_t sx,
                                   const float* img, int img_c, float alpha, int n, int h, int w,
                                   float scale, float* loss, void* stream) {
  SmoothArgs a;
  int rc = smooth_args(a, def, sn, sc, sy, sx, img, img_c, alpha, n, h, w, scale);
  if (rc) return rc;
  NEMAR_REQUIRE(loss, "smoothness_fwd: loss");
  int64_t total = (int64_t)n * h * w;
  smoothness_fwd_kernel<<<grid_for(total, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(a, loss);
  NEMAR_LAUNCH_CHECK();
  return 0;
}

NEMAR_API int nemar_smoothness_bwd(const float* def, int64_t sn, int64_t sc, int64_t sy, int64_t sx,
                                   const float* img, int img_c, float alpha, int n, int h, int w,
                                   float scale, const float* gscale, float* ddef, void* stream) {
  SmoothArgs a;
  int rc = smooth_args(a, def, sn, sc, sy, sx, img, img_c, alpha, n, h, w, scale);
  if (rc) return rc;
  NEMAR_REQUIRE(gscale && ddef, "smoothness_bwd: pointers");
  int64_t total = (int64_t)n * h * w;
  smoothness_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(a, gscale, ddef);
  NEMAR_LAUNCH_CHECK();
  return 0;
}
