// conv_api.cu — C-ABI dispatch for Conv2d / ConvTranspose2d forward, data gradient and weight gradient.
// Maps (operation, transposed) onto the two gather-GEMM modes and picks the engine (tcgen05 or generic).
// Replaces nn.Conv2d / nn.ConvTranspose2d + autograd (models/networks.py:350,357,369-372,376,426,439,576,
// 583,591,597; models/stn/layers.py:85).
//
// Channel padding: the geometry carries the REAL channel counts (the reference's weight shape); activation
// tensors may carry more channels (x->c >= cin, y->c >= cout; the excess is zero) so that 3/6-channel images
// and 1/2/3-channel heads fit the tensor-core tiles.  Packed weights are [y->c][taps][x->c] (forward pack) and
// [x->c][taps][y->c] (backward pack) with zero padding; `bias` must hold y->c floats.
#include "common.cuh"
#include "conv_internal.cuh"

static bool geom_ok(const nemar_conv_geom* g) {
  // rectangular kernels (the 7 x 1 halves of the generator's k7 head / tail) only as "valid" stride-1 convolutions
  return g && g->cin > 0 && g->cout > 0 && g->kh > 0 && g->kw > 0 && g->stride > 0 && g->pad >= 0 &&
         (g->kh == g->kw || (g->pad == 0 && g->stride == 1 && !g->transposed));
}

NEMAR_API int nemar_pack_weights(const float* w, const nemar_conv_geom* g, int dtype, int cin_p, int cout_p,
                                 void* wf, void* wd, void* stream) {
  NEMAR_REQUIRE(w && geom_ok(g) && cin_p >= g->cin && cout_p >= g->cout, "pack_weights: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = 0;
  if (!g->transposed) {
    // w[cout][cin][kh][kw]
    if (wf) rc = generic_pack(w, wf, dtype, g->cout, cout_p, g->cin, cin_p, g->kh, g->kw, /*w_is_oi=*/1, /*flip=*/0, s);
    if (!rc && wd) rc = generic_pack(w, wd, dtype, g->cin, cin_p, g->cout, cout_p, g->kh, g->kw, /*w_is_oi=*/0, /*flip=*/1, s);
  } else {
    // w[cin][cout][kh][kw]; forward pack [cout][taps][cin_p] flipped, backward pack [cin][taps][cout_p]
    if (wf) rc = generic_pack(w, wf, dtype, g->cout, cout_p, g->cin, cin_p, g->kh, g->kw, /*w_is_oi=*/0, /*flip=*/1, s);
    if (!rc && wd) rc = generic_pack(w, wd, dtype, g->cin, cin_p, g->cout, cout_p, g->kh, g->kw, /*w_is_oi=*/1, /*flip=*/0, s);
  }
  return rc;
}

NEMAR_API int nemar_pack_weights_multi(const nemar_pack_job* jobs_dev, const int* blocks_dev, int nblocks, void* stream) {
  NEMAR_REQUIRE(jobs_dev && blocks_dev && nblocks > 0, "pack_weights_multi: bad args");
  return generic_pack_multi(jobs_dev, blocks_dev, nblocks, (cudaStream_t)stream);
}

static int expected_out(int in, const nemar_conv_geom* g, int k) {
  return (in + 2 * g->pad - k) / g->stride + 1;
}

// geometry of the forward pass as a gather problem; false: bad arguments (error set)
static int fprop_geom(const nemar_tensor* x, const nemar_conv_geom* g, const nemar_tensor* y, GatherGeom& gg) {
  gg.kh = g->kh; gg.kw = g->kw; gg.dst_padded = 0;
  if (!g->transposed) {
    NEMAR_REQUIRE(x->pad <= g->pad, "conv2d_fprop: input halo larger than the conv padding");
    NEMAR_REQUIRE(y->h == expected_out(x->h, g, g->kh) && y->w == expected_out(x->w, g, g->kw), "conv2d_fprop: bad output extent");
    gg.sm = g->stride; gg.sd = 1; gg.pe = gg.pe_x = g->pad - x->pad;
  } else {
    NEMAR_REQUIRE(x->pad == 0, "conv2d_fprop(transposed): input halo not supported");
    int lo_h = (x->h - 1) * g->stride - 2 * g->pad + g->kh, lo_w = (x->w - 1) * g->stride - 2 * g->pad + g->kw;
    NEMAR_REQUIRE(y->h >= lo_h && y->h < lo_h + g->stride && y->w >= lo_w && y->w < lo_w + g->stride,
                  "conv2d_fprop(transposed): bad output extent");
    gg.sm = 1; gg.sd = g->stride; gg.pe = gg.pe_x = g->kh - 1 - g->pad;
  }
  return 0;
}

NEMAR_API int64_t nemar_conv2d_fprop_stats_workspace(const nemar_tensor* x, int w_cin_p, const nemar_conv_geom* g,
                                                     const nemar_tensor* y, int use_tc) {
  if (!use_tc || !view_ok(x) || !view_ok(y) || !geom_ok(g) || y->pad != 0 || w_cin_p != x->c) return 0;
  GatherGeom gg;
  if (fprop_geom(x, g, y, gg)) return 0;
  return tc_gather_stats_workspace(x, y, w_cin_p, gg);
}

NEMAR_API int nemar_conv2d_fprop_ws(const nemar_tensor* x, const void* w_packed, int w_cin_p, const float* bias,
                                    const nemar_conv_geom* g, int act, const nemar_tensor* y, float* stats,
                                    float* stats_ws, int64_t stats_ws_bytes, int use_tc, void* stream) {
  NEMAR_REQUIRE(view_ok(x) && view_ok(y) && w_packed && geom_ok(g), "conv2d_fprop: bad args");
  NEMAR_REQUIRE(x->c >= g->cin && y->c >= g->cout && x->n == y->n && w_cin_p == x->c &&
                    (x->dtype == y->dtype || y->dtype == NEMAR_F32),
                "conv2d_fprop: channel/dtype mismatch (weights are packed for x->c input channels and share x's "
                "dtype; y may be fp32)");
  NEMAR_REQUIRE(y->pad == 0, "conv2d_fprop: output must not carry a halo");
  cudaStream_t s = (cudaStream_t)stream;
  GatherGeom gg;
  int rc = fprop_geom(x, g, y, gg);
  if (rc) return rc;
  if (use_tc && tc_gather_supported(x, y, w_cin_p, gg))   // otherwise: CUDA-core engine (still on the GPU)
    return tc_gather_gemm(x, y, w_packed, w_cin_p, bias, act, stats, gg, s, stats_ws, stats_ws_bytes);
  rc = generic_gather_gemm(x, y, w_packed, x->dtype, w_cin_p, bias, act, gg, s);
  if (!rc && stats) {
    NEMAR_REQUIRE(act == NEMAR_ACT_NONE, "conv2d_fprop: stats need the pre-activation output");
    rc = nemar_instnorm_stats(y, stats, stream);
  }
  return rc;
}

NEMAR_API int nemar_conv2d_fprop(const nemar_tensor* x, const void* w_packed, int w_cin_p, const float* bias,
                                 const nemar_conv_geom* g, int act, const nemar_tensor* y, float* stats,
                                 int use_tc, void* stream) {
  return nemar_conv2d_fprop_ws(x, w_packed, w_cin_p, bias, g, act, y, stats, nullptr, 0, use_tc, stream);
}

NEMAR_API int nemar_conv2d_dgrad(const nemar_tensor* dy, const void* w_packed_d, int w_cout_p,
                                 const nemar_conv_geom* g, const nemar_tensor* dx, int use_tc, void* stream) {
  NEMAR_REQUIRE(view_ok(dy) && view_ok(dx) && w_packed_d && geom_ok(g), "conv2d_dgrad: bad args");
  NEMAR_REQUIRE(dy->c >= g->cout && dx->c >= g->cin && dx->n == dy->n && w_cout_p == dy->c && dy->pad == 0 &&
                    (dx->dtype == dy->dtype || dy->dtype == NEMAR_F32),
                "conv2d_dgrad: channel/dtype mismatch (weights are packed for dy->c channels and share dx's dtype)");
  cudaStream_t s = (cudaStream_t)stream;
  GatherGeom gg;
  gg.kh = g->kh; gg.kw = g->kw;
  if (!g->transposed) {
    NEMAR_REQUIRE(dx->pad <= g->pad, "conv2d_dgrad: dx halo larger than the conv padding");
    gg.sm = 1; gg.sd = g->stride; gg.dst_padded = dx->pad > 0;
    gg.pe = g->kh - 1 - (g->pad - dx->pad); gg.pe_x = g->kw - 1 - (g->pad - dx->pad);
  } else {
    NEMAR_REQUIRE(dx->pad == 0, "conv2d_dgrad(transposed): halo not supported");
    gg.sm = g->stride; gg.sd = 1; gg.pe = gg.pe_x = g->pad; gg.dst_padded = 0;
  }
  if (use_tc && tc_gather_supported(dy, dx, w_cout_p, gg))
    return tc_gather_gemm(dy, dx, w_packed_d, w_cout_p, nullptr, NEMAR_ACT_NONE, nullptr, gg, s);
  return generic_gather_gemm(dy, dx, w_packed_d, dx->dtype, w_cout_p, nullptr, NEMAR_ACT_NONE, gg, s);
}

// conv-view of a weight-gradient problem: for ConvTranspose2d the roles of x and dy swap.
struct WgradView {
  const nemar_tensor *xc, *dyc;
  int pe, co_real, ci_real;
};
static WgradView wgrad_view(const nemar_tensor* x, const nemar_tensor* dy, const nemar_conv_geom* g) {
  WgradView v;
  if (!g->transposed) { v.xc = x; v.dyc = dy; v.pe = g->pad - x->pad; v.co_real = g->cout; v.ci_real = g->cin; }
  else { v.xc = dy; v.dyc = x; v.pe = g->pad - dy->pad; v.co_real = g->cin; v.ci_real = g->cout; }
  return v;
}

NEMAR_API int64_t nemar_conv2d_wgrad_workspace(const nemar_tensor* x, const nemar_tensor* dy,
                                               const nemar_conv_geom* g, int use_tc) {
  if (!use_tc || !view_ok(x) || !view_ok(dy) || !geom_ok(g)) return 0;
  WgradView v = wgrad_view(x, dy, g);
  return tc_wgrad_workspace(v.xc, v.dyc, g->kh, g->kw, g->stride, v.pe);
}

NEMAR_API int nemar_conv2d_wgrad(const nemar_tensor* x, const nemar_tensor* dy, const nemar_conv_geom* g,
                                 float* dw, void* workspace, int64_t workspace_bytes, int use_tc, int accumulate,
                                 void* stream) {
  NEMAR_REQUIRE(view_ok(x) && view_ok(dy) && dw && geom_ok(g), "conv2d_wgrad: bad args");
  NEMAR_REQUIRE(x->c >= g->cin && dy->c >= g->cout && x->n == dy->n, "conv2d_wgrad: channel mismatch");
  WgradView v = wgrad_view(x, dy, g);
  NEMAR_REQUIRE(v.dyc->pad == 0 && v.pe >= 0, "conv2d_wgrad: unsupported halo configuration");
  cudaStream_t s = (cudaStream_t)stream;
  if (use_tc && tc_wgrad_supported(v.xc, v.dyc, g->kh, g->kw, g->stride, v.pe))
    return tc_wgrad(v.xc, v.dyc, dw, v.co_real, v.ci_real, g->kh, g->kw, g->stride, v.pe, workspace, workspace_bytes,
                    accumulate, s);
  // generic engine: narrow the views to the real channels (the padding holds zeros and contributes nothing)
  nemar_tensor xr = *v.xc, dr = *v.dyc;
  xr.c = v.ci_real; dr.c = v.co_real;
  return generic_wgrad(&xr, &dr, dw, g->kh, g->kw, g->stride, v.pe, accumulate, s);
}

NEMAR_API int nemar_conv2d_tc_supported(const nemar_conv_geom* g, int dtype, int h_in, int w_in) {
  (void)h_in; (void)w_in;
  if (!geom_ok(g) || dtype != NEMAR_BF16 || !tc_engine_built()) return 0;
  if (!(g->stride == 1 || g->stride == 2) || g->kh * g->kw > 49) return 0;
  return 1;   // any channel count: the host pads activations/weights to a multiple of 16
}

NEMAR_API int nemar_conv2d_set_option(const char* key, int value) { return tc_set_option(key, value); }
