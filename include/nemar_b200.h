/*
 * nemar_b200.h — C ABI of libnemar_b200.so (B200 / sm_100a engine for the NeMAR training hot path).
 *
 * The reference (moabarar/nemar) has no FFI of its own: every op below is an ATen call made from its
 * Python modules.  Each entry point names the reference call site it replaces (paths relative to the
 * reference checkout).  Conventions:
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch's allocator in our host code);
 *     the library never allocates device memory;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream;
 *   - return value: 0 = ok, <0 = invalid argument (see nemar_last_error()), >0 = cudaError_t;
 *   - activations live in "padded NHWC" tensors described by `nemar_tensor` (channels innermost,
 *     optional spatial halo, optional channel-slice view); 3-channel images / offsets that cross the
 *     reference-facing module boundary are plain NCHW fp32;
 *   - dtype: NEMAR_F32 or NEMAR_BF16 storage; arithmetic is always fp32 (accumulators fp32).
 */
#ifndef NEMAR_B200_H
#define NEMAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEMAR_F32 0
#define NEMAR_BF16 1

#define NEMAR_ACT_NONE 0
#define NEMAR_ACT_RELU 1
#define NEMAR_ACT_LRELU 2 /* negative slope 0.2 (networks.py:576,585,593; stn/layers.py:63) */
#define NEMAR_ACT_TANH 3

#define NEMAR_PAD_ZERO 0
#define NEMAR_PAD_REFLECT 1

/* Element (n,y,x,ch) of the view lives at
 *   ptr[ (((n*(h+2*pad) + y+pad) * (w+2*pad)) + x+pad) * cs + coff + ch ]   (units: elements of dtype). */
typedef struct nemar_tensor {
  void* ptr;
  int32_t n, h, w, c; /* logical extent of the view                                  */
  int32_t pad;        /* spatial halo carried by the buffer on every side            */
  int32_t cs;         /* elements per pixel in the underlying buffer (>= coff + c)   */
  int32_t coff;       /* first channel of this view inside the pixel                 */
  int32_t dtype;      /* NEMAR_F32 | NEMAR_BF16                                      */
} nemar_tensor;

/* Convolution geometry (cross-correlation, as nn.Conv2d).  `transposed` selects nn.ConvTranspose2d
 * semantics (networks.py:369-372) where (cin,cout) are the transposed conv's own in/out channels. */
typedef struct nemar_conv_geom {
  int32_t cin, cout, kh, kw, stride, pad;
  int32_t transposed; /* 0: Conv2d, 1: ConvTranspose2d (output_padding given by out tensor extent) */
} nemar_conv_geom;

const char* nemar_last_error(void);
/* name of the kernel the most recent nemar_conv2d_* call on this thread launched ("" before the first); measurement aid */
const char* nemar_last_conv_kernel(void);
int nemar_version(void);
/* 1 when the tcgen05/TMA implicit-GEMM path can take this geometry+dtype, else 0 (generic path). */
int nemar_conv2d_tc_supported(const nemar_conv_geom* g, int dtype, int h_in, int w_in);
/* Engine tuning switch (process-wide; not a semantic option — results stay within the same tolerance).
 *   "pair": 1 = destinations of k*256 channels run on CTA pairs (tcgen05 cta_group::2, 256x256 tiles), 0 = single-CTA
 *   tiles, 2 / 3 = pairs only in the fprop+dgrad / only in the wgrad kernel.  value < 0 only queries.  Returns the previous value, or -1 for an unknown key.  Initial value:
 *   environment NEMAR_TC_PAIR, else 3. */
int nemar_conv2d_set_option(const char* key, int value);

/* ---------------------------------------------------------------------------------------------
 * Layout crossings at the reference-facing module boundary (NCHW fp32 images <-> engine NHWC).
 * Replaces: implicit in the reference (everything NCHW); torch.cat(...,1) of images
 *           (nemar_model.py:181,197,219,233,247; affine_stn.py:79; unet_stn.py:82) is a channel-slice write.
 * --------------------------------------------------------------------------------------------- */
/* dst view (c == C) <- src NCHW fp32 [N,C,H,W]; fills dst halo per pad_mode (reflect: ReflectionPad2d,
 * networks.py:349,375) */
int nemar_nchw_to_nhwc(const float* src, const nemar_tensor* dst, int pad_mode, void* stream);
/* dst NCHW fp32 [N,C,H,W] (=|+=) src view; when src carries a halo its contents are folded back onto
 * the interior per pad_mode (adjoint of the padding) */
int nemar_nhwc_to_nchw(const nemar_tensor* src, float* dst, int pad_mode, int accumulate, void* stream);
/* zero channels [c0, c0+nc) of every pixel (interior+halo) of the buffer described by t */
int nemar_fill_channels(const nemar_tensor* t, int c0, int nc, void* stream);
/* dst view <- src view (same n,h,w,c; any pads/slices); halo of dst filled per pad_mode */
int nemar_copy_view(const nemar_tensor* src, const nemar_tensor* dst, int pad_mode, void* stream);
/* dst view <- src view with a storage-dtype change (fp32 <-> bf16); same n,h,w,c; interiors only */
int nemar_cast_view(const nemar_tensor* src, const nemar_tensor* dst, void* stream);
/* Tap <-> channel transforms that turn the 3-channel k7 head and the 3-channel k7 tail of the generator
 * (networks.py:349-350,375-376) into 1x1 convolutions for the tensor-core engine (src/dst are plain images,
 * pad must be 0; out-of-range source pixels read as zero):
 *   gather_taps: dst[n,y,x,(a*k+b)*c+ch] = src[n, y+sgn*a, x+sgn*b, ch]   (ch < c; channels >= k*k*c are zeroed)
 *   sum_taps:    dst[n,y,x,ch] = act(bias[ch] + sum_{a,b} src[n, y-sgn*a, x-sgn*b, (a*k+b)*c+ch])   (ch < c) */
int nemar_gather_taps(const nemar_tensor* src, const nemar_tensor* dst, int k, int c, int sgn, void* stream);
/* rectangular tap window ky x kx (a < ky rows, b < kx columns; channel index (a*kx+b)*c+ch); (1, 7) is the column half
 * of the generator's k7 head / tail, whose row half runs as a 7 x 1 tensor-core convolution */
int nemar_gather_taps2(const nemar_tensor* src, const nemar_tensor* dst, int ky, int kx, int c, int sgn, void* stream);
int nemar_sum_taps2(const nemar_tensor* src, const nemar_tensor* dst, int ky, int kx, int c, int sgn, const float* bias,
                    int act, void* stream);
int nemar_sum_taps(const nemar_tensor* src, const nemar_tensor* dst, int k, int c, int sgn, const float* bias,
                   int act, void* stream);
/* dst view (=|+=) src view folded (adjoint of nemar_copy_view) */
int nemar_copy_view_bwd(const nemar_tensor* dsrc_out, const nemar_tensor* ddst_in, int pad_mode,
                        int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Conv2d / ConvTranspose2d.  Replaces nn.Conv2d / nn.ConvTranspose2d forward + autograd
 * (networks.py:350,357,369-372,376,426,439,576,583,591,597; stn/layers.py:85).
 * Weights are consumed in a packed layout produced by nemar_pack_weights:
 *   fprop pack  Wf[cout][kh][kw][cin_p]           (cin_p = cin rounded up to `cin_round`)
 *   dgrad pack  Wd[cin][kh][kw][cout_p]  with taps flipped (so dgrad is itself a correlation)
 * For transposed==1 the roles swap (the transposed conv's forward is the dgrad of a Conv2d).
 * --------------------------------------------------------------------------------------------- */
int nemar_pack_weights(const float* w /* reference layout */, const nemar_conv_geom* g, int dtype,
                       int cin_p, int cout_p, void* wf /* may be NULL */, void* wd /* may be NULL */,
                       void* stream);
/* Every pack of an optimizer's layers in ONE launch (the per-layer call above costs 152 launches per training step at
 * C2).  jobs_dev: device array of nemar_pack_job — the element mapping of ONE pack of nemar_pack_weights
 * (out[o][kh][kw][i] chunk-major, zero padded; w indexed [o][i] or [i][o]; taps optionally flipped), or with kind == 2 a
 * zero-padded fp32 vector copy (padded bias).  blocks_dev: nblocks pairs (job index, 2048-element slice index). */
typedef struct nemar_pack_job {
  const float* w;
  void* out;
  int32_t O, op, I, ip, kh, kw, w_is_oi, flip;
  int32_t dtype; /* NEMAR_F32 / NEMAR_BF16 of `out` */
  int32_t kind;  /* 1: weight pack, 2: padded vector copy */
} nemar_pack_job;
int nemar_pack_weights_multi(const nemar_pack_job* jobs_dev, const int* blocks_dev, int nblocks, void* stream);

/* y = act(conv(x) + bias); x halo (if any) is consumed as real data ("valid" conv over the padded
 * buffer when x->pad == g->pad with reflect halo; zero padding needs no halo).  If stats != NULL,
 * per-(n,cout) [sum, sum of squares] of the pre-activation fp32 values are accumulated into
 * stats[n][cout][2] (must be zeroed by the caller) — InstanceNorm statistics in the conv epilogue.
 * use_tc: 1 = tcgen05/TMA kernel (must be supported), 0 = generic CUDA-core kernel. */
int nemar_conv2d_fprop(const nemar_tensor* x, const void* w_packed, int w_cin_p, const float* bias,
                       const nemar_conv_geom* g, int act, const nemar_tensor* y, float* stats,
                       int use_tc, void* stream);
/* Same, with InstanceNorm statistics produced IN THE CONVOLUTION'S EPILOGUE when the geometry allows it: the tiled
 * tensor-core kernel writes per-tile column sums into `stats_ws` (plain stores) and a small kernel adds them up into
 * stats[n][cout][2] (overwritten: no zeroing needed, no atomics, deterministic).  nemar_conv2d_fprop_stats_workspace
 * returns the bytes `stats_ws` must hold, or 0 when this geometry takes the separate statistics pass (then stats must
 * be zeroed by the caller as above and stats_ws is ignored). */
int64_t nemar_conv2d_fprop_stats_workspace(const nemar_tensor* x, int w_cin_p, const nemar_conv_geom* g,
                                           const nemar_tensor* y, int use_tc);
int nemar_conv2d_fprop_ws(const nemar_tensor* x, const void* w_packed, int w_cin_p, const float* bias,
                          const nemar_conv_geom* g, int act, const nemar_tensor* y, float* stats,
                          float* stats_ws, int64_t stats_ws_bytes, int use_tc, void* stream);
/* dx = conv_dgrad(dy); dx halo (if dx->pad>0) receives the raw padded-buffer gradient. */
int nemar_conv2d_dgrad(const nemar_tensor* dy, const void* w_packed_d, int w_cout_p,
                       const nemar_conv_geom* g, const nemar_tensor* dx, int use_tc, void* stream);
/* dw (reference layout fp32; overwritten, or accumulated into when accumulate != 0 — the flat gradient bucket of
 * the optimizer) = wgrad(x, dy); workspace >= nemar_conv2d_wgrad_workspace() */
int64_t nemar_conv2d_wgrad_workspace(const nemar_tensor* x, const nemar_tensor* dy,
                                     const nemar_conv_geom* g, int use_tc);
int nemar_conv2d_wgrad(const nemar_tensor* x, const nemar_tensor* dy, const nemar_conv_geom* g,
                       float* dw, void* workspace, int64_t workspace_bytes, int use_tc, int accumulate, void* stream);
/* db[c] = sum over n,h,w of dy (overwritten) */
int nemar_bias_grad(const nemar_tensor* dy, float* db, void* stream);

/* ---------------------------------------------------------------------------------------------
 * InstanceNorm2d(affine=False, eps=1e-5) + activation + residual + reflect halo, one pass.
 * Replaces nn.InstanceNorm2d / ReLU / LeakyReLU / ReflectionPad2d / skip-add
 * (networks.py:24,351-352,418,426,432,439,445,584-585; stn/layers.py:16,99-105).
 * --------------------------------------------------------------------------------------------- */
/* stats[n][c][2] += (sum, sumsq) over h,w of x (for producers without a fused epilogue) */
int nemar_instnorm_stats(const nemar_tensor* x, float* stats, void* stream);
/* y = act( norm ? (x-mean)*rstd : x ) + residual ; mean/rstd derived from stats (sum,sumsq)/HW */
int nemar_norm_act_fwd(const nemar_tensor* x, const float* stats /* NULL: no norm */, int act,
                       const nemar_tensor* residual /* NULL */, const nemar_tensor* y, int pad_mode,
                       void* stream);
/* phase A: red[n][c][2] += (sum g, sum g*xhat), g = fold(dy)*act'(.) ; red must be zeroed */
int nemar_norm_act_bwd_reduce(const nemar_tensor* x, const float* stats, int act,
                              const nemar_tensor* dy, int pad_mode, float* red, void* stream);
/* phase B: dx = rstd*(g - mean(g) - xhat*mean(g*xhat)) (or g when stats==NULL);
 *          dres (optional) (=|+=) fold(dy);  db (optional, [c] floats) (=|+=) sum over n,h,w of dx — the bias gradient
 *          of the convolution that produced x, fused into this pass.
 *          flags: bit 0 = dres accumulates, bit 1 = db accumulates (e.g. straight into the optimizer's gradient
 *          bucket), bit 2 = zero the halo ring of dres here (the pass writes only its interior) */
int nemar_norm_act_bwd_apply(const nemar_tensor* x, const float* stats, int act,
                             const nemar_tensor* dy, int pad_mode, const float* red,
                             const nemar_tensor* dx, const nemar_tensor* dres, int flags,
                             float* db, void* stream);
/* dx = dy * act'(y) for an activation fused in a conv epilogue (uses the OUTPUT y) */
int nemar_act_bwd(const nemar_tensor* y, const nemar_tensor* dy, int act, const nemar_tensor* dx,
                  void* stream);

/* MaxPool2d(2) (stn/layers.py:174) */
int nemar_maxpool2_fwd(const nemar_tensor* x, const nemar_tensor* y, void* stream);
int nemar_maxpool2_bwd(const nemar_tensor* x, const nemar_tensor* y, const nemar_tensor* dy,
                       const nemar_tensor* dx, void* stream);
/* F.interpolate(mode='bilinear', align_corners=False) (unet_stn.py:96,166,188-195) on NHWC views */
int nemar_bilinear_resize_fwd(const nemar_tensor* x, const nemar_tensor* y, void* stream);
int nemar_bilinear_resize_bwd(const nemar_tensor* dy, const nemar_tensor* dx, int accumulate,
                              void* stream);
/* same op on NCHW fp32 images (nemar_model.py:187-188,204-205,226-227,240-241,254-255) */
int nemar_bilinear_resize_nchw_fwd(const float* x, int n, int c, int h, int w, float* y, int ho,
                                   int wo, void* stream);
int nemar_bilinear_resize_nchw_bwd(const float* dy, int n, int c, int h, int w, float* dx, int ho,
                                   int wo, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Sampling grids and bilinear grid_sample (the STN head).
 * --------------------------------------------------------------------------------------------- */
/* F.affine_grid(theta.view(-1,2,3), size, align_corners=False) (affine_stn.py:105,128).
 * bx[W], by[H]: base coordinates (linspace(-1,1,n)*(n-1)/n, computed by the caller exactly as ATen).
 * grid [N,H,W,2] fp32 (x,y). */
int nemar_affine_grid_fwd(const float* theta, const float* bx, const float* by, int n, int h, int w,
                          float* grid, void* stream);
/* dtheta[N,6] = adjoint (overwritten) */
int nemar_affine_grid_bwd(const float* dgrid, const float* bx, const float* by, int n, int h, int w,
                          float* dtheta, void* stream);
/* grid = identity + offsets, laid out [N,H,W,2] (unet_stn.py:121-129,167): xs[W], ys[H] are the
 * caller's linspace(-1,1,·); offsets element (n,ch,y,x) at off[n*sn + ch*sc + y*sy + x*sx]
 * (so both the reference NCHW tensor and the engine's NHWC conv output are accepted). */
int nemar_flow_grid_fwd(const float* off, int64_t sn, int64_t sc, int64_t sy, int64_t sx,
                        const float* xs, const float* ys, int n, int h, int w, float* grid,
                        void* stream);
/* F.grid_sample(img, grid, 'bilinear', 'zeros', align_corners=False) (affine_stn.py:129-130,
 * unet_stn.py:173-174; ATen/native/GridSampler.h:26-36,205-207).  img/out NCHW fp32, grid [N,H,W,2].
 * One grid read serves `nimg` (1 or 2) images (the reference samples real_A and fake_B with the same grid).
 * idx_dump (optional, int32 [N,Ho,Wo,2]) receives (x0,y0)=floor coordinates for the bit-exactness test. */
int nemar_grid_sample_fwd(const float* img0, const float* img1, int nimg, int n, int c, int h, int w,
                          const float* grid, int ho, int wo, float* out0, float* out1,
                          int32_t* idx_dump, void* stream);
/* dimg{0,1} may be NULL (no input gradient needed, e.g. real_A); when given they must be zeroed by the
 * caller (scatter-add).  dgrid [N,Ho,Wo,2] is overwritten with the sum over images. */
int nemar_grid_sample_bwd(const float* img0, const float* img1, int nimg, int n, int c, int h, int w,
                          const float* grid, int ho, int wo, const float* dout0, const float* dout1,
                          float* dimg0, float* dimg1, float* dgrid, void* stream);
/* smoothness_loss (stn_losses.py:4-30): loss[0] += scale * sum_k mean(weight_k * |delta_k def|).
 * def element (n,ch,y,x) at def[n*sn+ch*sc+y*sy+x*sx]; img NCHW fp32 [N,C,H,W] or NULL / alpha<=0. */
int nemar_smoothness_fwd(const float* def, int64_t sn, int64_t sc, int64_t sy, int64_t sx,
                         const float* img, int img_c, float alpha, int n, int h, int w, float scale,
                         float* loss, void* stream);
/* ddef (same strides as def, must be zeroed by the caller) += gscale[0] * scale * d loss / d def */
int nemar_smoothness_bwd(const float* def, int64_t sn, int64_t sc, int64_t sy, int64_t sx,
                         const float* img, int img_c, float alpha, int n, int h, int w, float scale,
                         const float* gscale, float* ddef, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Losses (nn.L1Loss nemar_model.py:68,179,195; MSELoss-vs-constant networks.py:237-238,273-275;
 * mean|dtheta| affine_stn.py:136-138).  out[0] += scale * mean(...).  Backward: dx (=|+=) gscale[0]*scale*d/dx.
 * --------------------------------------------------------------------------------------------- */
int nemar_l1_fwd(const float* a, const float* b, int64_t numel, float scale, float* out, void* stream);
int nemar_l1_bwd(const float* a, const float* b, int64_t numel, float scale, const float* gscale,
                 float* da, int accumulate, void* stream);
/* pred is an engine tensor view (the discriminator's 1-channel output) */
int nemar_mse_const_fwd(const nemar_tensor* pred, float target, float scale, float* out, void* stream);
int nemar_mse_const_bwd(const nemar_tensor* pred, float target, float scale, const float* gscale,
                        const nemar_tensor* dpred, void* stream);
int nemar_mean_abs_fwd(const float* a, int64_t numel, float scale, float* out, void* stream);
int nemar_mean_abs_bwd(const float* a, int64_t numel, float scale, const float* gscale, float* da,
                       void* stream);

/* nn.Linear (affine_stn.py:69-72): y[N,O] = act(x[N,I] @ W[O,I]^T + b) ; fp32 */
int nemar_linear_fwd(const float* x, const float* w, const float* b, int n, int i, int o, int act,
                     float* y, void* stream);
int nemar_linear_bwd(const float* x, const float* w, const float* y, const float* dy, int n, int i,
                     int o, int act, float* dx, float* dw, float* db, void* stream);

/* torch.optim.Adam step (nemar_model.py:128-137) over a flat fp32 buffer; step_count is 1-based. */
int nemar_adam_step(float* p, const float* g, float* m, float* v, int64_t numel, float lr, float beta1,
                    float beta2, float eps, int step_count, float grad_scale, void* stream);

/* Same update with the 1-based step number read from DEVICE memory (*step_dev is incremented by the call, in stream
 * order) so that the launch can be replayed from a CUDA graph. */
int nemar_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t numel, float lr, float beta1,
                        float beta2, float eps, int32_t* step_dev, float grad_scale, void* stream);

/* Dropout(0.5) of the ResnetBlock (networks.py:427-428): counter-based mask, y = keep ? 2x : 0 */
int nemar_dropout(const nemar_tensor* x, const nemar_tensor* y, uint64_t seed, uint64_t offset,
                  void* stream);

/* Same mask generator, keyed by (seed, salt, *step_dev): the step number is read from DEVICE memory so that a step
 * captured in a CUDA graph draws a new mask on every replay (forward and backward of one step read the same value).
 * nemar_counter_add bumps such a counter in stream order. */
int nemar_dropout_dev(const nemar_tensor* x, const nemar_tensor* y, uint64_t seed, uint64_t salt,
                      const int64_t* step_dev, void* stream);
int nemar_counter_add(int64_t* counter_dev, int64_t v, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEMAR_B200_H */
