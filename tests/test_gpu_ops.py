"""Op-level parity of every CUDA kernel against ATen on CPU (the library the reference calls) and the C oracle.
All calls go through the C ABI (ctypes) via the autograd bindings.  fp32 storage: tight; bf16 storage: the
inputs are pre-rounded to bf16 so only accumulation order and the output rounding differ."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu

from nemar_b200.engine import functional as F  # noqa: E402
from nemar_b200.engine import lib as L  # noqa: E402
from oracle import grid_oracle as G  # noqa: E402
from oracle import nemar_oracle as O  # noqa: E402

DEV = "cuda"
DT = {"fp32": torch.float32, "bf16": torch.bfloat16}


def gen(*shape, seed=0, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def q(t, dt):
    """round to the storage dtype (on CPU, kept as fp32 values)"""
    return t.to(dt).float()


def nhwc(x_nchw, dt, pad=0, mode="reflect"):
    if pad:
        x_nchw = TF.pad(x_nchw, (pad,) * 4, mode=mode) if mode == "reflect" else TF.pad(x_nchw, (pad,) * 4)
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(dt).to(DEV)


def nchw(t_nhwc):
    return t_nhwc.float().cpu().permute(0, 3, 1, 2).contiguous()


def close(a, b, rtol, atol, what=""):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    assert bool((err <= tol).all()), "%s: max err %.3e (|ref| max %.3e), worst tol ratio %.2f" % (
        what, float(err.max()), float(b.abs().max()), float((err / tol).max()))


def rel_rms(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-12))


# ------------------------------------------------------------------------------------------------
# STN head
# ------------------------------------------------------------------------------------------------
def _grids(n, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    ident = O.identity_grid(h, w).permute(0, 2, 3, 1).repeat(n, 1, 1, 1).contiguous()
    return dict(identity=ident, near=ident + torch.randn((n, h, w, 2), generator=g) * (4.0 / w),
                wild=torch.rand((n, h, w, 2), generator=g) * 2.2 - 1.1)


@pytest.mark.parametrize("h,w", [(64, 64), (37, 53), (256, 256), (30, 2), (512, 512), (1024, 1024), (288, 384)])
def test_grid_sample_indices_bit_exact_and_values(h, w):
    n, c = (1 if h * w >= 512 * 512 else 2), 3
    img = gen(n, c, h, w, seed=3)
    for name, grid in _grids(n, h, w, 5).items():
        out, idx = F.grid_sample_indices(grid.to(DEV), img.to(DEV))
        ref_out, ref_idx = G.grid_sample_fwd(img.numpy(), grid.numpy())
        assert np.array_equal(idx.cpu().numpy(), ref_idx), "indices differ from the oracle (%s)" % name
        aten = TF.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
        close(out, aten, 0, 1e-5, "grid_sample fwd " + name)
        close(out, torch.from_numpy(ref_out), 2e-5, 1e-5, "grid_sample fwd vs C oracle " + name)
        # the training path (no index dump) must reproduce the index-dumping launch bit for bit
        plain = F.GridSampleFn.apply(grid.to(DEV), img.to(DEV), None)
        assert torch.equal(plain, out), "grid_sample with / without the index dump differ (%s)" % name


@pytest.mark.parametrize("h,w", [(64, 64), (37, 53), (128, 256)])
def test_grid_sample_two_images_fwd_bwd(h, w):
    n, c = 2, 3
    img0, img1 = gen(n, c, h, w, seed=1), gen(n, c, h, w, seed=2)
    d0, d1 = gen(n, c, h, w, seed=3), gen(n, c, h, w, seed=4)
    for name, grid in _grids(n, h, w, 9).items():
        gr = grid.clone().requires_grad_(True)
        i0r, i1r = img0.clone().requires_grad_(True), img1.clone().requires_grad_(True)
        o0 = TF.grid_sample(i0r, gr, mode="bilinear", padding_mode="zeros", align_corners=False)
        o1 = TF.grid_sample(i1r, gr, mode="bilinear", padding_mode="zeros", align_corners=False)
        (o0 * d0).sum().add((o1 * d1).sum()).backward()
        gg = grid.to(DEV).requires_grad_(True)
        a0, a1 = img0.to(DEV), img1.to(DEV).requires_grad_(True)      # real_A needs no grad, fake_B does
        e0, e1 = F.GridSampleFn.apply(gg, a0, a1)
        close(e0, o0, 0, 1e-5, "out0 " + name)
        close(e1, o1, 0, 1e-5, "out1 " + name)
        ((e0 * d0.to(DEV)).sum() + (e1 * d1.to(DEV)).sum()).backward()
        close(a1.grad, i1r.grad, 1e-4, 2e-5, "dimg1 " + name)
        close(gg.grad, gr.grad, 2e-4, 5e-3, "dgrid " + name)
        assert a0.grad is None


def test_affine_and_flow_grid_fwd_bwd():
    n, h, w = 3, 48, 80
    theta = (torch.tensor([1, 0, 0, 0, 1, 0.0]).repeat(n, 1) + gen(n, 6, seed=1, scale=0.05)).requires_grad_(True)
    ref = TF.affine_grid(theta.view(-1, 2, 3), (n, 3, h, w), align_corners=False)
    dg = gen(n, h, w, 2, seed=2)
    (ref * dg).sum().backward()
    bx = torch.from_numpy(G.affine_base(w)).to(DEV)
    by = torch.from_numpy(G.affine_base(h)).to(DEV)
    th = theta.detach().to(DEV).requires_grad_(True)
    grid = F.AffineGridFn.apply(th, bx, by)
    close(grid, ref, 0, 5e-7, "affine grid")
    (grid * dg.to(DEV)).sum().backward()
    close(th.grad, theta.grad, 1e-4, 1e-3, "dtheta")
    off = gen(n, 2, h, w, seed=3, scale=0.02)
    xs, ys = torch.linspace(-1, 1, w).to(DEV), torch.linspace(-1, 1, h).to(DEV)
    eng = F.FlowGridFn.apply(off.permute(0, 2, 3, 1).contiguous().to(DEV), xs, ys)
    assert np.array_equal(eng.cpu().numpy(), G.flow_grid(off.numpy())), "flow grid must be bit-exact"


@pytest.mark.parametrize("alpha", [0.0, 1.0])
@pytest.mark.parametrize("h,w", [(33, 47), (128, 128)])
def test_smoothness_fwd_bwd(alpha, h, w):
    n = 2
    d = gen(n, 2, h, w, seed=1, scale=0.02).requires_grad_(True)
    img = torch.rand(n, 3, h, w, generator=torch.Generator().manual_seed(2)) * 2 - 1
    ref = O.smoothness_loss(d, img, alpha) * 0.5
    ref.backward()
    de = d.detach().permute(0, 2, 3, 1).contiguous().to(DEV).requires_grad_(True)
    out = F.SmoothnessFn.apply(de, img.to(DEV), alpha, 0.5)
    close(out, ref.reshape(1), 2e-5, 1e-7, "smoothness")
    (out * 3.0).sum().backward()
    close(de.grad, 3.0 * d.grad.permute(0, 2, 3, 1), 1e-4, 1e-9, "smoothness grad")


# ------------------------------------------------------------------------------------------------
# layout, norm, pooling, resize
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_images_to_nhwc_and_back(prec):
    dt = DT[prec]
    a, b = gen(2, 3, 20, 28, seed=1), gen(2, 3, 20, 28, seed=2)
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = TF.pad(torch.cat([ar, br], 1), (3, 3, 3, 3), mode="reflect")
    dg = gen(*ref.shape, seed=3)
    (ref * dg).sum().backward()
    ae, be = a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    t = F.ImagesToNHWC.apply(3, L.PAD_REFLECT, dt, 8, ae, be)
    assert t.shape == (2, 26, 34, 8)
    close(nchw(t)[:, :6], q(ref, dt), 0, 0, "images->nhwc")
    assert float(t[..., 6:].abs().max()) == 0.0
    dge = torch.zeros(2, 26, 34, 8)
    dge[..., :6] = dg.permute(0, 2, 3, 1)
    (t.float() * dge.to(DEV)).sum().backward()
    tol = 1e-6 if prec == "fp32" else 2e-2
    close(ae.grad, ar.grad, tol, tol, "d images a")
    close(be.grad, br.grad, tol, tol, "d images b")
    y = gen(2, 5, 9, 11, seed=4)
    out = F.ToNCHW.apply(nhwc(y, dt), 3)
    close(out, q(y, dt)[:, :3], 0, 0, "to nchw")


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("act", ["none", "relu", "lrelu"])
@pytest.mark.parametrize("c,h,w,out_pad,res", [(64, 16, 16, 1, True), (32, 9, 13, 0, False), (6, 7, 5, 3, False),
                                               (256, 8, 8, 1, True),
                                               # long enough for the cp.async rings to wrap several times per thread
                                               (256, 64, 64, 1, True), (128, 31, 31, 1, False), (32, 96, 80, 0, False)])
def test_norm_act_fwd_bwd(prec, act, c, h, w, out_pad, res):
    dt = DT[prec]
    n = 2
    x = q(gen(n, c, h, w, seed=1) * 1.5 + 0.3, dt)
    r = q(gen(n, c, h, w, seed=2), dt) if res else None
    xr = x.clone().requires_grad_(True)
    rr = r.clone().requires_grad_(True) if res else None
    y = TF.instance_norm(xr, eps=1e-5)
    y = {"none": lambda v: v, "relu": TF.relu, "lrelu": lambda v: TF.leaky_relu(v, 0.2)}[act](y)
    if res:
        y = y + rr
    yp = TF.pad(y, (out_pad,) * 4, mode="reflect") if out_pad else y
    dg = q(gen(*yp.shape, seed=3), dt)
    (yp * dg).sum().backward()
    code = {"none": L.ACT_NONE, "relu": L.ACT_RELU, "lrelu": L.ACT_LRELU}[act]
    xe = nhwc(x, dt).requires_grad_(True)
    re_ = nhwc(r, dt, 1).requires_grad_(True) if res else None     # residual arrives with its own reflect halo
    stats = torch.zeros(n, c, 2, device=DEV)
    L.call("nemar_instnorm_stats", L.view(xe.detach()), L.fptr(stats), L.stream())
    bias = torch.zeros(c, device=DEV, requires_grad=True)     # receives the fused column-sum (conv bias gradient)
    ye = F.NormActFn.apply(xe, stats, re_, code, 1, out_pad, L.PAD_REFLECT, bias)
    ftol = dict(fp32=(1e-4, 1e-4), bf16=(1.6e-2, 1.6e-2))[prec]
    close(nchw(ye), yp, *ftol, "norm_act fwd")
    (ye.float() * nhwc(dg, dt).float()).sum().backward()
    btol = dict(fp32=(2e-3, 2e-4), bf16=(3e-2, 3e-2))[prec]
    close(nchw(xe.grad), xr.grad, *btol, "norm_act dx")
    colsum = nchw(xe.grad).sum((0, 2, 3))
    close(bias.grad, colsum, 1e-3, 1e-3 * float(nchw(xe.grad).abs().sum((0, 2, 3)).max()) + 1e-6, "fused bias grad")
    if prec == "fp32":
        assert rel_rms(nchw(xe.grad), xr.grad) < 2e-5, "norm_act dx rel-rms %.3e" % rel_rms(nchw(xe.grad), xr.grad)
    if res:
        close(nchw(re_.grad)[:, :, 1:-1, 1:-1], rr.grad, *btol, "norm_act dres")
        assert float(re_.grad[:, 0].abs().max()) == 0.0


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_maxpool_resize_concat(prec):
    dt = DT[prec]
    x = q(gen(2, 16, 13, 18, seed=1), dt)
    x[:, :, :4, :4] = 0.0  # ties: gradient must go to the first maximum
    xr = x.clone().requires_grad_(True)
    y = TF.max_pool2d(xr, 2)
    dg = q(gen(*y.shape, seed=2), dt)
    (y * dg).sum().backward()
    xe = nhwc(x, dt).requires_grad_(True)
    ye = F.MaxPool2Fn.apply(xe)
    close(nchw(ye), y, 0, 0, "maxpool fwd")
    (ye.float() * nhwc(dg, dt).float()).sum().backward()
    close(nchw(xe.grad), xr.grad, 0, 0, "maxpool bwd")
    for (oh, ow) in [(26, 36), (9, 7), (13, 18), (27, 35)]:
        xr = x.clone().requires_grad_(True)
        y = TF.interpolate(xr, (oh, ow), mode="bilinear", align_corners=False)
        dg = q(gen(*y.shape, seed=3), dt)
        (y * dg).sum().backward()
        xe = nhwc(x, dt).requires_grad_(True)
        ye = F.ResizeFn.apply(xe, oh, ow)
        tol = dict(fp32=(1e-5, 1e-5), bf16=(1e-2, 1e-2))[prec]
        close(nchw(ye), y, *tol, "resize fwd %dx%d" % (oh, ow))
        (ye.float() * nhwc(dg, dt).float()).sum().backward()
        close(nchw(xe.grad), xr.grad, tol[0] * 2, tol[1] * 4, "resize bwd %dx%d" % (oh, ow))
    a, b = q(gen(2, 8, 5, 6, seed=4), dt), q(gen(2, 24, 5, 6, seed=5), dt)
    ae, be = nhwc(a, dt).requires_grad_(True), nhwc(b, dt).requires_grad_(True)
    ce = F.Concat.apply(ae, be)
    close(nchw(ce), torch.cat([a, b], 1), 0, 0, "concat")
    dg = q(gen(2, 32, 5, 6, seed=6), dt)
    (ce.float() * nhwc(dg, dt).float()).sum().backward()
    close(nchw(ae.grad), dg[:, :8], 0, 0, "concat da")
    close(nchw(be.grad), dg[:, 8:], 0, 0, "concat db")
    img = gen(2, 3, 32, 48, seed=7).requires_grad_(True)
    ref = TF.interpolate(img, (16, 24), mode="bilinear", align_corners=False)
    ref.backward(torch.ones_like(ref))
    ie = img.detach().to(DEV).requires_grad_(True)
    out = F.ResizeNCHWFn.apply(ie, 16, 24)
    close(out, ref, 1e-6, 1e-6, "resize nchw")
    out.backward(torch.ones_like(out))
    close(ie.grad, img.grad, 1e-6, 1e-6, "resize nchw bwd")


# ------------------------------------------------------------------------------------------------
# convolution (generic engine; the tcgen05 engine is cross-checked in test_gpu_conv_tc.py)
# ------------------------------------------------------------------------------------------------
GEOMS = [  # cin, cout, k, stride, pad, transposed, x_pad(reflect halo), h, w
    (3, 64, 7, 1, 3, False, 3, 20, 24),     # netT head on a reflect-padded image
    (64, 128, 3, 2, 1, False, 0, 16, 20),   # netT downsample
    (64, 64, 3, 1, 1, False, 1, 12, 12),    # ResnetBlock conv on a reflect-padded map
    (128, 64, 3, 2, 1, True, 0, 8, 10),     # netT ConvTranspose2d k3 s2 p1 op1
    (64, 3, 7, 1, 3, False, 3, 16, 16),     # netT tail (skinny output)
    (6, 64, 4, 2, 1, False, 0, 32, 32),     # PatchGAN first layer
    (64, 128, 4, 2, 1, False, 0, 16, 16),   # PatchGAN k4 s2
    (128, 256, 4, 1, 1, False, 0, 9, 9),    # PatchGAN k4 s1 (31x31-style odd output)
    (256, 1, 4, 1, 1, False, 0, 7, 7),      # PatchGAN prediction
    (64, 128, 1, 1, 0, False, 0, 2, 2),     # ResUnet bottleneck 1x1 on a 2x2 map
    (96, 32, 3, 1, 1, False, 0, 15, 17),    # ResUnet up_1 (cat 64+32), odd extent
    (32, 2, 3, 1, 1, False, 0, 16, 16),     # offset head
]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("geom", GEOMS, ids=lambda g: "c%d-%d_k%d_s%d_p%d_%s" % (g[0], g[1], g[2], g[3], g[4], "T" if g[5] else "C"))
def test_conv_generic_fwd_bwd(prec, geom):
    cin, cout, k, stride, pad, transposed, x_pad, h, w = geom
    dt = DT[prec]
    n = 2
    x = q(gen(n, cin, h, w, seed=1), dt)
    wshape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    wt = gen(*wshape, seed=2, scale=(cin * k * k) ** -0.5)
    wq = q(wt, dt)
    b = gen(cout, seed=3, scale=0.1)
    xr, wr, br = x.clone().requires_grad_(True), wq.clone().requires_grad_(True), b.clone().requires_grad_(True)
    if transposed:
        y = TF.conv_transpose2d(xr, wr, br, stride=stride, padding=pad, output_padding=1)
    elif x_pad:
        y = TF.conv2d(TF.pad(xr, (x_pad,) * 4, mode="reflect"), wr, br, stride=stride)
    else:
        y = TF.conv2d(xr, wr, br, stride=stride, padding=pad)
    dg = q(gen(*y.shape, seed=4), dt)
    (y * dg).sum().backward()
    # engine: weights stay fp32 masters; the pack rounds them to the storage dtype
    xe = nhwc(x, dt, x_pad).requires_grad_(True)
    we, be = wt.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    cfg = F.ConvCfg(cin, cout, k, stride, pad, transposed, x_pad, L.ACT_NONE, False, False, 1 if transposed else 0, False)
    ye = F.Conv2dFn.apply(xe, we, be, cfg, F.PackedWeights())
    ftol = dict(fp32=(2e-4, 2e-4), bf16=(1.2e-2, 1.2e-2))[prec]
    close(nchw(ye), y, *ftol, "conv fwd")
    (ye.float() * nhwc(dg, dt).float()).sum().backward()
    gx = nchw(xe.grad)
    if x_pad:   # fold the reflect halo of the engine's padded-buffer gradient back (adjoint of the padding)
        xz = torch.zeros(n, cin, h, w, requires_grad=True)
        TF.pad(xz, (x_pad,) * 4, mode="reflect").backward(gx)
        gx = xz.grad
    btol = dict(fp32=(5e-4, 5e-4), bf16=(2e-2, 2e-2))[prec]
    close(gx, xr.grad, *btol, "conv dgrad")
    if prec == "fp32":
        assert rel_rms(nchw(ye), y) < 1e-5, "conv fwd rel-rms %.3e" % rel_rms(nchw(ye), y)
        assert rel_rms(gx, xr.grad) < 1e-5, "conv dgrad rel-rms %.3e" % rel_rms(gx, xr.grad)
        assert rel_rms(be.grad, br.grad) < 1e-5, "conv bias-grad rel-rms %.3e" % rel_rms(be.grad, br.grad)
    assert rel_rms(we.grad, wr.grad) < dict(fp32=1e-5, bf16=1e-2)[prec], "conv wgrad rel-rms %.3e" % rel_rms(we.grad, wr.grad)
    close(be.grad, br.grad, 1e-3, 1e-2 if prec == "bf16" else 1e-3, "conv bias grad")


@pytest.mark.parametrize("act", ["lrelu", "tanh"])
def test_conv_fused_activation_and_fp32_tail(act):
    """bf16 input, fp32 output tail (offset head / PatchGAN prediction / tanh image)."""
    dt = torch.bfloat16
    x = q(gen(2, 32, 12, 14, seed=1), dt)
    wt = gen(3, 32, 3, 3, seed=2, scale=0.08)
    b = gen(3, seed=3, scale=0.1)
    xr, wr, br = x.clone().requires_grad_(True), q(wt, dt).requires_grad_(True), b.clone().requires_grad_(True)
    y = TF.conv2d(xr, wr, br, padding=1)
    y = TF.leaky_relu(y, 0.2) if act == "lrelu" else torch.tanh(y)
    dg = gen(*y.shape, seed=4)
    (y * dg).sum().backward()
    xe = nhwc(x, dt).requires_grad_(True)
    we, be = wt.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    code = L.ACT_LRELU if act == "lrelu" else L.ACT_TANH
    cfg = F.ConvCfg(32, 3, 3, 1, 1, False, 0, code, False, True, 0, False)
    ye = F.Conv2dFn.apply(xe, we, be, cfg, F.PackedWeights())
    assert ye.dtype == torch.float32
    close(nchw(ye), y, 1e-4, 1e-4, "fused act fwd")
    (ye * nhwc(dg, torch.float32)).sum().backward()
    close(nchw(xe.grad), xr.grad, 2e-2, 2e-2, "fused act dgrad")
    assert rel_rms(we.grad, wr.grad) < 2e-3


# ------------------------------------------------------------------------------------------------
# losses, linear, Adam
# ------------------------------------------------------------------------------------------------
def test_losses_linear_adam():
    a, b = gen(2, 3, 17, 19, seed=1).requires_grad_(True), gen(2, 3, 17, 19, seed=2)
    ref = 100.0 * TF.l1_loss(a, b)
    ref.backward()
    ae = a.detach().to(DEV).requires_grad_(True)
    out = F.L1Fn.apply(ae, b.to(DEV), 100.0)
    close(out, ref.reshape(1), 1e-5, 1e-6, "l1")
    out.sum().backward()
    close(ae.grad, a.grad, 1e-6, 1e-9, "l1 grad")
    p = gen(2, 1, 9, 11, seed=3).requires_grad_(True)
    ref = TF.mse_loss(p, torch.ones_like(p))
    ref.backward()
    pe = p.detach().permute(0, 2, 3, 1).contiguous().to(DEV).requires_grad_(True)
    out = F.MSEConstFn.apply(pe, 1.0, 1.0)
    close(out, ref.reshape(1), 1e-5, 1e-6, "mse")
    out.sum().backward()
    close(nchw(pe.grad), p.grad, 1e-5, 1e-8, "mse grad")
    t = gen(4, 6, seed=4, scale=0.1).requires_grad_(True)
    ref = t.abs().mean()
    ref.backward()
    te = t.detach().to(DEV).requires_grad_(True)
    out = F.MeanAbsFn.apply(te, 1.0)
    out.sum().backward()
    close(out, ref.reshape(1), 1e-5, 1e-7, "mean abs")
    close(te.grad, t.grad, 1e-6, 1e-9, "mean abs grad")
    x, w, bb = gen(3, 1024, seed=5).requires_grad_(True), gen(256, 1024, seed=6, scale=0.03).requires_grad_(True), gen(256, seed=7).requires_grad_(True)
    ref = TF.relu(TF.linear(x, w, bb))
    dg = gen(3, 256, seed=8)
    (ref * dg).sum().backward()
    xe, we, be = [v.detach().to(DEV).requires_grad_(True) for v in (x, w, bb)]
    out = F.LinearFn.apply(xe, we, be, L.ACT_RELU)
    close(out, ref, 1e-4, 1e-4, "linear")
    (out * dg.to(DEV)).sum().backward()
    close(xe.grad, x.grad, 1e-4, 1e-4, "linear dx")
    close(we.grad, w.grad, 1e-4, 1e-4, "linear dw")
    close(be.grad, bb.grad, 1e-4, 1e-4, "linear db")
    # Adam vs torch.optim.Adam over 3 steps (numel not a multiple of 4 exercises the tail)
    p0 = gen(1003, seed=9)
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([pt], lr=2e-4, betas=(0.5, 0.999))
    pe, m, v = torch.zeros(1004, device=DEV), torch.zeros(1004, device=DEV), torch.zeros(1004, device=DEV)
    pe[:1003] = p0.to(DEV)
    for s in range(1, 4):
        g = gen(1003, seed=10 + s)
        pt.grad = g.clone()
        opt.step()
        ge = torch.zeros(1004, device=DEV)
        ge[:1003] = g.to(DEV) * 2.0
        F.adam_step(pe, ge, m, v, 2e-4, 0.5, 0.999, 1e-8, s, 0.5)
    close(pe[:1003], pt.detach(), 1e-6, 1e-7, "adam")


def test_tap_transforms_head_and_tail():
    """k7 conv with 3 input (head) / 3 output (tail) channels == tap<->channel transform + 1x1 conv (fp32 check of
    the two transform kernels against F.conv2d, using torch matmul for the 1x1 part)."""
    n, h, w, k, p = 2, 12, 14, 7, 3
    x = gen(n, 3, h, w, seed=1)
    wt = gen(8, 3, k, k, seed=2, scale=0.1)
    xr, wr = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    ref = TF.conv2d(TF.pad(xr, (p,) * 4, mode="reflect"), wr)
    dg = gen(*ref.shape, seed=3)
    (ref * dg).sum().backward()
    xe = nhwc(x, torch.float32, p).requires_grad_(True)                       # [n,h+6,w+6,3]
    u = F.GatherTapsFn.apply(xe, k, 3, 1, h, w, 160, torch.float32)
    assert float(u[..., 147:].abs().max()) == 0.0
    w1 = wt.permute(0, 2, 3, 1).reshape(8, 147).to(DEV)
    y = u[..., :147] @ w1.t()
    close(nchw(y), ref, 1e-4, 1e-4, "head via gather_taps")
    (y * nhwc(dg, torch.float32)).sum().backward()
    xz = torch.zeros(n, 3, h, w, requires_grad=True)
    TF.pad(xz, (p,) * 4, mode="reflect").backward(nchw(xe.grad))
    close(xz.grad, xr.grad, 1e-4, 1e-4, "head dgrad via sum_taps")
    # tail: 8 -> 3 channels
    x = gen(n, 8, h, w, seed=4)
    wt = gen(3, 8, k, k, seed=5, scale=0.1)
    b = gen(3, seed=6, scale=0.1)
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.tanh(TF.conv2d(TF.pad(xr, (p,) * 4, mode="reflect"), wr, br))
    dg = gen(*ref.shape, seed=7)
    (ref * dg).sum().backward()
    xe = nhwc(x, torch.float32, p).requires_grad_(True)                       # [n,h+6,w+6,8]
    wv = wt.permute(2, 3, 0, 1).reshape(147, 8).to(DEV)
    v = torch.zeros(n, h + 2 * p, w + 2 * p, 160, device=DEV)
    v[..., :147] = xe @ wv.t()
    be = b.to(DEV).requires_grad_(True)
    y = F.SumTapsFn.apply(v, be, k, 3, -1, h, w, 4, L.ACT_TANH)
    close(nchw(y)[:, :3], ref, 1e-4, 1e-4, "tail via sum_taps")
    (y[..., :3] * nhwc(dg, torch.float32)).sum().backward()
    xz = torch.zeros(n, 8, h, w, requires_grad=True)
    TF.pad(xz, (p,) * 4, mode="reflect").backward(nchw(xe.grad))
    close(xz.grad, xr.grad, 1e-4, 1e-4, "tail dgrad via gather_taps")
    close(be.grad, br.grad, 1e-4, 1e-4, "tail bias grad")
