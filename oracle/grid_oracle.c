/* grid_oracle.c — plain-C restatement of the STN head arithmetic.  TEST INFRASTRUCTURE, NOT PRODUCT CODE:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may load it.
 *
 * Follows the reference's call sites models/stn/affine_stn.py:128-130 and models/stn/unet_stn.py:121-129,167,
 * 173-174 into PyTorch ATen (third-party, un-vendored; torch 2.11.0 here):
 *   ATen/native/GridSampler.h:26-36   grid_sampler_unnormalize: ((coord + 1) * size - 1) / 2   (align_corners=False);
 *       evaluated by ATen as ONE fused multiply-add fma(coord + 1, size / 2, -0.5) (vectorised CPU kernel; nvcc contraction
 *       in the CUDA kernel): pinned by tests/test_grid_oracle.py at 288 x 384, where the two-rounding form is 1.4e-5 off
 *   ATen/native/GridSampler.h:205-207 within_bounds_2d; bilinear taps nw/ne/sw/se from floor(ix), floor(iy)
 *   ATen/native/AffineGridGenerator.cpp linspace(-1,1,n)*(n-1)/n base grid, grid = base @ theta^T
 * and models/stn/stn_losses.py:4-30 for the smoothness term.  Pinned by tests/test_grid_oracle.py against
 * torch.nn.functional on this machine (the library the reference calls). */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define API __attribute__((visibility("default")))

static inline int inb(int x, int y, int w, int h) { return x >= 0 && x < w && y >= 0 && y < h; }

API void oracle_affine_grid(const float* theta, const float* bx, const float* by, int n, int h, int w, float* grid) {
  for (int i = 0; i < n; ++i)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        const float* t = theta + 6 * i;
        float* g = grid + (((size_t)i * h + y) * w + x) * 2;
        g[0] = t[0] * bx[x] + t[1] * by[y] + t[2];
        g[1] = t[3] * bx[x] + t[4] * by[y] + t[5];
      }
}

API void oracle_flow_grid(const float* off /* NCHW [n,2,h,w] */, const float* xs, const float* ys, int n, int h, int w,
                          float* grid) {
  for (int i = 0; i < n; ++i)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        float* g = grid + (((size_t)i * h + y) * w + x) * 2;
        g[0] = xs[x] + off[(((size_t)i * 2 + 0) * h + y) * w + x];
        g[1] = ys[y] + off[(((size_t)i * 2 + 1) * h + y) * w + x];
      }
}

/* out [n,c,ho,wo]; idx [n,ho,wo,2] = (floor ix, floor iy) */
API void oracle_grid_sample_fwd(const float* img, int n, int c, int h, int w, const float* grid, int ho, int wo,
                                float* out, int32_t* idx) {
  for (int i = 0; i < n; ++i)
    for (int y = 0; y < ho; ++y)
      for (int x = 0; x < wo; ++x) {
        const float* g = grid + (((size_t)i * ho + y) * wo + x) * 2;
        /* one fused multiply-add, as both ATen builds evaluate it (see the header note) */
        float ix = fmaf(g[0] + 1.f, (float)w * 0.5f, -0.5f);
        float iy = fmaf(g[1] + 1.f, (float)h * 0.5f, -0.5f);
        float fx = floorf(ix), fy = floorf(iy);
        int x0 = (int)fx, y0 = (int)fy;
        float x1 = fx + 1.f, y1 = fy + 1.f;
        float nw = (x1 - ix) * (y1 - iy), ne = (ix - fx) * (y1 - iy), sw = (x1 - ix) * (iy - fy), se = (ix - fx) * (iy - fy);
        if (idx) {
          idx[(((size_t)i * ho + y) * wo + x) * 2] = x0;
          idx[(((size_t)i * ho + y) * wo + x) * 2 + 1] = y0;
        }
        for (int ch = 0; ch < c; ++ch) {
          const float* p = img + ((size_t)i * c + ch) * h * w;
          float acc = 0.f;
          if (inb(x0, y0, w, h)) acc += p[(size_t)y0 * w + x0] * nw;
          if (inb(x0 + 1, y0, w, h)) acc += p[(size_t)y0 * w + x0 + 1] * ne;
          if (inb(x0, y0 + 1, w, h)) acc += p[(size_t)(y0 + 1) * w + x0] * sw;
          if (inb(x0 + 1, y0 + 1, w, h)) acc += p[(size_t)(y0 + 1) * w + x0 + 1] * se;
          out[(((size_t)i * c + ch) * ho + y) * wo + x] = acc;
        }
      }
}

/* dimg (may be NULL, else zero-initialised by the caller) and dgrid [n,ho,wo,2] */
API void oracle_grid_sample_bwd(const float* img, int n, int c, int h, int w, const float* grid, int ho, int wo,
                                const float* dout, float* dimg, float* dgrid) {
  for (int i = 0; i < n; ++i)
    for (int y = 0; y < ho; ++y)
      for (int x = 0; x < wo; ++x) {
        const float* g = grid + (((size_t)i * ho + y) * wo + x) * 2;
        /* one fused multiply-add, as both ATen builds evaluate it (see the header note) */
        float ix = fmaf(g[0] + 1.f, (float)w * 0.5f, -0.5f);
        float iy = fmaf(g[1] + 1.f, (float)h * 0.5f, -0.5f);
        float fx = floorf(ix), fy = floorf(iy);
        int x0 = (int)fx, y0 = (int)fy;
        float x1 = fx + 1.f, y1 = fy + 1.f;
        float nw = (x1 - ix) * (y1 - iy), ne = (ix - fx) * (y1 - iy), sw = (x1 - ix) * (iy - fy), se = (ix - fx) * (iy - fy);
        float gix = 0.f, giy = 0.f;
        for (int ch = 0; ch < c; ++ch) {
          const float* p = img + ((size_t)i * c + ch) * h * w;
          float* dp = dimg ? dimg + ((size_t)i * c + ch) * h * w : NULL;
          float go = dout[(((size_t)i * c + ch) * ho + y) * wo + x];
          if (inb(x0, y0, w, h)) {
            float v = p[(size_t)y0 * w + x0];
            if (dp) dp[(size_t)y0 * w + x0] += nw * go;
            gix -= v * (y1 - iy) * go; giy -= v * (x1 - ix) * go;
          }
          if (inb(x0 + 1, y0, w, h)) {
            float v = p[(size_t)y0 * w + x0 + 1];
            if (dp) dp[(size_t)y0 * w + x0 + 1] += ne * go;
            gix += v * (y1 - iy) * go; giy -= v * (ix - fx) * go;
          }
          if (inb(x0, y0 + 1, w, h)) {
            float v = p[(size_t)(y0 + 1) * w + x0];
            if (dp) dp[(size_t)(y0 + 1) * w + x0] += sw * go;
            gix -= v * (iy - fy) * go; giy += v * (x1 - ix) * go;
          }
          if (inb(x0 + 1, y0 + 1, w, h)) {
            float v = p[(size_t)(y0 + 1) * w + x0 + 1];
            if (dp) dp[(size_t)(y0 + 1) * w + x0 + 1] += se * go;
            gix += v * (iy - fy) * go; giy += v * (ix - fx) * go;
          }
        }
        dgrid[(((size_t)i * ho + y) * wo + x) * 2] = ((float)w / 2.f) * gix;
        dgrid[(((size_t)i * ho + y) * wo + x) * 2 + 1] = ((float)h / 2.f) * giy;
      }
}

/* def NCHW [n,2,h,w], img NCHW [n,c,h,w] or NULL */
API double oracle_smoothness(const float* def, const float* img, int c, float alpha, int n, int h, int w) {
  double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
  for (int i = 0; i < n; ++i)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        double w1 = 1, w2 = 1, w3 = 1, w4 = 1;
        if (img && alpha > 0.f) {
          w1 = w2 = w3 = w4 = 0;
          for (int ch = 0; ch < c; ++ch) {
            const float* p = img + ((size_t)i * c + ch) * h * w;
#define P(yy, xx) p[(size_t)(yy) * w + (xx)]
            if (y + 1 < h) w1 += expf(-alpha * fabsf(P(y + 1, x) - P(y, x)));
            if (x + 1 < w) w2 += expf(-alpha * fabsf(P(y, x + 1) - P(y, x)));
            if (y + 1 < h && x + 1 < w) {
              w3 += expf(-alpha * fabsf(P(y, x) - P(y + 1, x + 1)));
              w4 += expf(-alpha * fabsf(P(y, x + 1) - P(y + 1, x)));
            }
#undef P
          }
          w1 /= c; w2 /= c; w3 /= c; w4 /= c;
        }
        for (int k = 0; k < 2; ++k) {
          const float* d = def + ((size_t)i * 2 + k) * h * w;
#define D(yy, xx) d[(size_t)(yy) * w + (xx)]
          if (y + 1 < h) s1 += w1 * fabs((double)D(y + 1, x) - D(y, x));
          if (x + 1 < w) s2 += w2 * fabs((double)D(y, x + 1) - D(y, x));
          if (y + 1 < h && x + 1 < w) {
            s3 += w3 * fabs((double)D(y, x) - D(y + 1, x + 1));
            s4 += w4 * fabs((double)D(y, x + 1) - D(y + 1, x));
          }
#undef D
        }
      }
  double c1 = (double)n * 2 * (h - 1) * w, c2 = (double)n * 2 * h * (w - 1), c34 = (double)n * 2 * (h - 1) * (w - 1);
  return s1 / c1 + s2 / c2 + s3 / c34 + s4 / c34;
}
