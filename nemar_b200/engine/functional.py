"""Autograd bindings: every differentiable op of the hot path as a torch.autograd.Function whose forward
and backward are single calls into libnemar_b200.so.  PyTorch supplies device memory, streams and the
autograd tape only — no ATen compute kernel runs on the named ops.

Engine tensors are contiguous [N, H+2*pad, W+2*pad, C] (channels innermost); the halo size is static
architecture knowledge passed alongside.  Images crossing the reference-facing boundary are NCHW fp32.
"""
import torch
from torch.autograd import Function

from . import lib as L
from .lib import call, fptr, i64, stream, view, vptr

_weights_epoch = [0]


def bump_weights_epoch():
    """Invalidate packed-weight caches (called by the optimizer step and by load_state_dict)."""
    _weights_epoch[0] += 1


def weights_epoch():
    return _weights_epoch[0]


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _deliver(param, grad):
    """Hand a parameter gradient over.  When the parameter already owns a gradient buffer (the optimizer's flat
    bucket), accumulate into it HERE, on the stream this backward runs on, and tell autograd there is nothing left
    to do: no AccumulateGrad kernel is launched later on some other stream (which also keeps the whole step
    capturable in a CUDA graph).  Otherwise return the gradient for autograd to store."""
    if grad is None:
        return None
    if param is not None and not param.requires_grad:
        return None                      # frozen by set_requires_grad(False): nothing may reach its bucket
    if param is not None and param.is_leaf and param.grad is not None and param.grad.shape == grad.shape:
        param.grad.add_(grad)
        return None
    return grad


# ------------------------------------------------------------------------------------------------
# layout crossings
# ------------------------------------------------------------------------------------------------
class ImagesToNHWC(Function):
    """cat(NCHW fp32 images, dim=1) -> engine tensor with optional reflect/zero halo and channel padding."""

    @staticmethod
    def forward(ctx, pad, pad_mode, dtype, cp, *imgs):
        imgs = [_c(i) for i in imgs]
        n, _, h, w = imgs[0].shape
        chans = [int(i.shape[1]) for i in imgs]
        ctot = sum(chans)
        cp = max(cp, ctot)
        out = torch.empty((n, h + 2 * pad, w + 2 * pad, cp), dtype=dtype, device=imgs[0].device)
        if cp > ctot:
            call("nemar_fill_channels", view(out, pad), ctot, cp - ctot, stream())
        coff = 0
        for img, c in zip(imgs, chans):
            assert img.dtype == torch.float32 and img.shape[0] == n and img.shape[2] == h and img.shape[3] == w
            call("nemar_nchw_to_nhwc", fptr(img), view(out, pad, coff, c), pad_mode, stream())
            coff += c
        ctx.meta = (pad, pad_mode, chans, (n, h, w))
        return out

    @staticmethod
    def backward(ctx, g):
        pad, pad_mode, chans, (n, h, w) = ctx.meta
        g = _c(g)
        grads, coff = [], 0
        for k, c in enumerate(chans):
            if ctx.needs_input_grad[4 + k]:
                d = torch.empty((n, c, h, w), dtype=torch.float32, device=g.device)
                call("nemar_nhwc_to_nchw", view(g, pad, coff, c), fptr(d), pad_mode, 0, stream())
                grads.append(d)
            else:
                grads.append(None)
            coff += c
        return (None, None, None, None, *grads)


class ImageGroupsToNHWC(Function):
    """Batch-concatenation of several channel-concatenations: group j = cat(imgs[j*m : (j+1)*m], dim=1) fills the
    samples [j*n, (j+1)*n) of ONE engine tensor, so that a network can take several (A, B_j) pairs in one pass
    (InstanceNorm keeps samples independent, so this equals separate passes).  The same image may appear in
    several groups (autograd sums its gradients)."""

    @staticmethod
    def forward(ctx, pad, pad_mode, dtype, cp, groups, *imgs):
        imgs = [_c(i) for i in imgs]
        m = len(imgs) // groups
        assert m * groups == len(imgs)
        n, _, h, w = imgs[0].shape
        chans = [int(i.shape[1]) for i in imgs[:m]]
        ctot = sum(chans)
        cp = max(cp, ctot)
        out = torch.empty((groups * n, h + 2 * pad, w + 2 * pad, cp), dtype=dtype, device=imgs[0].device)
        if cp > ctot:
            call("nemar_fill_channels", view(out, pad), ctot, cp - ctot, stream())
        for j in range(groups):
            part, coff = out[j * n:(j + 1) * n], 0
            for img, c in zip(imgs[j * m:(j + 1) * m], chans):
                assert img.dtype == torch.float32 and tuple(img.shape) == (n, c, h, w)
                call("nemar_nchw_to_nhwc", fptr(img), view(part, pad, coff, c), pad_mode, stream())
                coff += c
        ctx.meta = (pad, pad_mode, chans, (n, h, w), groups)
        return out

    @staticmethod
    def backward(ctx, g):
        pad, pad_mode, chans, (n, h, w), groups = ctx.meta
        g = _c(g)
        grads = []
        for j in range(groups):
            part, coff = g[j * n:(j + 1) * n], 0
            for k, c in enumerate(chans):
                if ctx.needs_input_grad[5 + j * len(chans) + k]:
                    d = torch.empty((n, c, h, w), dtype=torch.float32, device=g.device)
                    call("nemar_nhwc_to_nchw", view(part, pad, coff, c), fptr(d), pad_mode, 0, stream())
                    grads.append(d)
                else:
                    grads.append(None)
                coff += c
        return (None, None, None, None, None, *grads)


class ToNCHW(Function):
    """engine tensor (first c channels) -> NCHW fp32 image."""

    @staticmethod
    def forward(ctx, x, c):
        x = _c(x)
        n, h, w, cs = x.shape
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
        call("nemar_nhwc_to_nchw", view(x, 0, 0, c), fptr(out), L.PAD_ZERO, 0, stream())
        ctx.meta = (x.shape, x.dtype, c)
        return out

    @staticmethod
    def backward(ctx, g):
        shape, dtype, c = ctx.meta
        g = _c(g)
        dx = torch.empty(shape, dtype=dtype, device=g.device)
        if shape[3] > c:
            call("nemar_fill_channels", view(dx), c, shape[3] - c, stream())
        call("nemar_nchw_to_nhwc", fptr(g), view(dx, 0, 0, c), L.PAD_ZERO, stream())
        return dx, None


class Concat(Function):
    """channel concat of two engine tensors (unet_stn.py:97)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _c(a), _c(b)
        n, h, w, ca = a.shape
        cb = b.shape[3]
        out = torch.empty((n, h, w, ca + cb), dtype=a.dtype, device=a.device)
        call("nemar_copy_view", view(a), view(out, 0, 0, ca), L.PAD_ZERO, stream())
        call("nemar_copy_view", view(b), view(out, 0, ca, cb), L.PAD_ZERO, stream())
        ctx.meta = (ca, cb)
        return out

    @staticmethod
    def backward(ctx, g):
        ca, cb = ctx.meta
        g = _c(g)
        n, h, w, _ = g.shape
        da = db = None
        if ctx.needs_input_grad[0]:
            da = torch.empty((n, h, w, ca), dtype=g.dtype, device=g.device)
            call("nemar_copy_view_bwd", view(da), view(g, 0, 0, ca), L.PAD_ZERO, 0, stream())
        if ctx.needs_input_grad[1]:
            db = torch.empty((n, h, w, cb), dtype=g.dtype, device=g.device)
            call("nemar_copy_view_bwd", view(db), view(g, 0, ca, cb), L.PAD_ZERO, 0, stream())
        return da, db


# ------------------------------------------------------------------------------------------------
# convolution
# ------------------------------------------------------------------------------------------------
class ConvCfg:
    """Static description of one conv layer call (geometry + epilogue + engine choice).  cin/cout are the REAL
    channel counts (the reference's weight shape); cout_p >= cout is the channel count of the output tensor
    (zero-padded to a multiple of 16 on the tensor-core engine); the input tensor brings its own padding."""

    def __init__(self, cin, cout, k, stride=1, pad=0, transposed=False, x_pad=0, act=L.ACT_NONE, stats=False,
                 out_f32=False, out_pad_t=0, use_tc=False, cout_p=None, defer_bias_grad=False):
        self.geom = L.ConvGeom(cin, cout, k, k, stride, pad, int(transposed))
        self.cin, self.cout, self.k, self.stride, self.pad = cin, cout, k, stride, pad
        self.transposed, self.x_pad, self.act, self.stats = transposed, x_pad, act, stats
        self.out_f32, self.out_pad_t, self.use_tc = out_f32, out_pad_t, use_tc
        self.cout_p = cout if cout_p is None else cout_p
        # True: the bias gradient is produced by the NormActFn consuming this conv's output (fused column sum)
        self.defer_bias_grad = defer_bias_grad

    def out_hw(self, h, w):
        if not self.transposed:
            f = lambda v: (v + 2 * self.pad - self.k) // self.stride + 1
        else:
            f = lambda v: (v - 1) * self.stride - 2 * self.pad + self.k + self.out_pad_t
        return f(h), f(w)


class PackedWeights:
    """Per-parameter cache of the forward / backward packs and the padded bias (rebuilt when the weights epoch
    changes or the layer is called with another channel padding)."""

    def __init__(self):
        self.cache = {}

    def get(self, weight, bias, cfg, dtype, cin_p):
        key = (dtype, cin_p, cfg.cout_p)
        ent = self.cache.get(key)
        # (epoch bumped by the engine's own optimizer / loaders; _version catches any other in-place write — a foreign
        #  load_state_dict, an EMA, a manual weight.copy_)
        stamp = (weights_epoch(), weight.data_ptr(), weight._version, bias._version if bias is not None else 0)
        if ent is None or ent[0] != stamp:
            k2 = cfg.k * cfg.k
            wf = torch.empty(cfg.cout_p * k2 * cin_p, dtype=dtype, device=weight.device)
            wd = torch.empty(cin_p * k2 * cfg.cout_p, dtype=dtype, device=weight.device)
            call("nemar_pack_weights", fptr(weight.detach()), cfg.geom, L.dtype_code(wf), cin_p, cfg.cout_p, vptr(wf),
                 vptr(wd), stream())
            bp = None
            if bias is not None:
                if cfg.cout_p == cfg.cout:
                    bp = bias.detach()
                else:
                    bp = torch.zeros(cfg.cout_p, dtype=torch.float32, device=weight.device)
                    bp[:cfg.cout].copy_(bias.detach())
            ent = (stamp, wf, wd, bp)
            self.cache[key] = ent
        return ent[1], ent[2], ent[3]


def _cast(t, dtype):
    out = torch.empty(t.shape, dtype=dtype, device=t.device)
    call("nemar_cast_view", view(t), view(out), stream())
    return out


class Conv2dFn(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, cfg, packed):
        x = _c(x)
        n, hp, wp, cin_p = x.shape
        h, w = hp - 2 * cfg.x_pad, wp - 2 * cfg.x_pad
        ho, wo = cfg.out_hw(h, w)
        wf, wd, bp = packed.get(weight, bias, cfg, x.dtype, cin_p)
        ydt = torch.float32 if cfg.out_f32 else x.dtype
        y = torch.empty((n, ho, wo, cfg.cout_p), dtype=ydt, device=x.device)
        stats = None
        if cfg.stats:
            stats = torch.zeros((n, cfg.cout_p, 2), dtype=torch.float32, device=x.device)
        # while the conv launches are being timed (bench.py's roofline pass) the InstanceNorm statistics pass is issued
        # as its own call, so that the events around nemar_conv2d_fprop bracket the convolution kernel alone; it is the
        # same kernel on the same buffers that nemar_conv2d_fprop launches itself when handed `stats`
        split_stats = stats is not None and L.TIMER.on
        call("nemar_conv2d_fprop", view(x, cfg.x_pad), vptr(wf), cin_p, fptr(bp), cfg.geom, cfg.act, view(y),
             fptr(None if split_stats else stats), int(cfg.use_tc), stream())
        if split_stats:
            call("nemar_instnorm_stats", view(y), fptr(stats), stream())
        ctx.cfg, ctx.wd = cfg, wd
        ctx.has_bias = bias is not None
        ctx.bias_param = bias
        ctx.save_for_backward(x, weight, y if cfg.act != L.ACT_NONE else None)
        if stats is None:
            return y
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, dstats=None):
        cfg = ctx.cfg
        x, weight, y = ctx.saved_tensors
        dy = _c(dy)
        if cfg.act != L.ACT_NONE:
            g = torch.empty_like(dy)
            call("nemar_act_bwd", view(y), view(dy), cfg.act, view(g), stream())
        else:
            g = dy
        # the tensor-core engine wants the output gradient in the storage dtype of the layer
        gk = _cast(g, x.dtype) if (cfg.use_tc and g.dtype != x.dtype) else g
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            call("nemar_conv2d_dgrad", view(gk), vptr(ctx.wd), cfg.cout_p, cfg.geom, view(dx, cfg.x_pad), int(cfg.use_tc),
                 stream())
        if ctx.needs_input_grad[1]:
            direct = weight.is_leaf and weight.grad is not None and weight.grad.is_contiguous()
            dw = weight.grad if direct else torch.empty(weight.shape, dtype=torch.float32, device=x.device)
            xv, gv = view(x, cfg.x_pad), view(gk)
            ws_bytes = L.lib().nemar_conv2d_wgrad_workspace(L.C.byref(xv), L.C.byref(gv), L.C.byref(cfg.geom),
                                                            int(cfg.use_tc))
            ws = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=x.device) if ws_bytes > 0 else None
            call("nemar_conv2d_wgrad", xv, gv, cfg.geom, fptr(dw), vptr(ws), i64(ws_bytes), int(cfg.use_tc), int(direct),
                 stream())
            if direct:
                dw = None
        if ctx.has_bias and ctx.needs_input_grad[2] and not cfg.defer_bias_grad:
            db = torch.empty(cfg.cout_p, dtype=torch.float32, device=x.device)
            call("nemar_bias_grad", view(g), fptr(db), stream())
            db = _deliver(ctx.bias_param, db[:cfg.cout])
        return dx, dw, db, None, None


class TapsWeightFn(Function):
    """The k x k weight of a <=4-channel head / tail seen as the weight of the equivalent 1x1 convolution:
    head: [co][ci][a][b] -> [co][(a,b,ci)][1][1];  tail: [co][ci][a][b] -> [(a,b,co)][ci][1][1].  A copy of a few
    thousand elements (torch glue); the backward delivers the gradient straight into the parameter's bucket."""

    @staticmethod
    def forward(ctx, weight, mode):
        co, ci, k, _ = weight.shape
        ctx.meta = (mode, co, ci, k)
        ctx.param = weight
        if mode == "head":
            return weight.detach().permute(0, 2, 3, 1).reshape(co, k * k * ci, 1, 1).contiguous()
        return weight.detach().permute(2, 3, 0, 1).reshape(k * k * co, ci, 1, 1).contiguous()

    @staticmethod
    def backward(ctx, g):
        mode, co, ci, k = ctx.meta
        if mode == "head":
            gw = g.reshape(co, k, k, ci).permute(0, 3, 1, 2)
        else:
            gw = g.reshape(k, k, co, ci).permute(2, 3, 0, 1)
        return _deliver(ctx.param, gw), None


class GatherTapsFn(Function):
    """dst[n,y,x,(a*k+b)*c+ch] = src[n, y+sgn*a, x+sgn*b, ch] — turns a k x k conv over c (<=4) channels into a
    1x1 conv over k*k*c channels (generator head, networks.py:349-350)."""

    @staticmethod
    def forward(ctx, x, k, c, sgn, oh, ow, cp, out_dtype):
        x = _c(x)
        n = x.shape[0]
        out = torch.empty((n, oh, ow, cp), dtype=out_dtype, device=x.device)
        call("nemar_gather_taps", view(x), view(out), k, c, sgn, stream())
        ctx.meta = (k, c, sgn, x.shape, x.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        k, c, sgn, shape, dtype = ctx.meta
        g = _c(g)
        dx = torch.empty(shape, dtype=dtype, device=g.device)
        call("nemar_sum_taps", view(g), view(dx), k, c, sgn, None, L.ACT_NONE, stream())
        return dx, None, None, None, None, None, None, None


class SumTapsFn(Function):
    """dst[n,y,x,ch] = act(bias[ch] + sum_{a,b} src[n, y-sgn*a, x-sgn*b, (a*k+b)*c+ch]) — the tap sum that completes
    a k x k conv with c (<=4) output channels computed as a 1x1 conv with k*k*c virtual channels (generator tail,
    networks.py:375-377)."""

    @staticmethod
    def forward(ctx, x, bias, k, c, sgn, oh, ow, cp, act):
        x = _c(x)
        n = x.shape[0]
        y = torch.empty((n, oh, ow, cp), dtype=torch.float32, device=x.device)
        call("nemar_sum_taps", view(x), view(y), k, c, sgn, fptr(bias.detach()) if bias is not None else None, act, stream())
        ctx.meta = (k, c, sgn, x.shape, x.dtype, act, bias is not None)
        ctx.bias_param = bias
        ctx.save_for_backward(y if act != L.ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        k, c, sgn, shape, dtype, act, has_bias = ctx.meta
        (y,) = ctx.saved_tensors
        dy = _c(dy)
        if act != L.ACT_NONE:
            g = torch.empty_like(dy)
            call("nemar_act_bwd", view(y), view(dy), act, view(g), stream())
        else:
            g = dy
        db = None
        if has_bias and ctx.needs_input_grad[1]:
            db = torch.empty(c, dtype=torch.float32, device=dy.device)
            call("nemar_bias_grad", view(g, 0, 0, c), fptr(db), stream())
        dx = torch.empty(shape, dtype=dtype, device=dy.device)
        call("nemar_gather_taps", view(g), view(dx), k, c, sgn, stream())
        return dx, _deliver(ctx.bias_param, db), None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# InstanceNorm + activation (+ residual) with halo
# ------------------------------------------------------------------------------------------------
class NormActFn(Function):
    @staticmethod
    def forward(ctx, x, stats, residual, act, res_pad, out_pad, pad_mode, bias=None):
        """`bias` (optional) is the bias Parameter of the conv that produced x: it takes no part in the forward,
        but its gradient (= column sums of dx) is produced by this op's backward pass for free."""
        x = _c(x)
        n, h, w, c = x.shape
        y = torch.empty((n, h + 2 * out_pad, w + 2 * out_pad, c), dtype=x.dtype, device=x.device)
        rv = None
        if residual is not None:
            residual = _c(residual)
            rv = view(residual, res_pad)
        call("nemar_norm_act_fwd", view(x), fptr(stats), act, rv, view(y, out_pad), pad_mode, stream())
        ctx.meta = (act, res_pad, out_pad, pad_mode, residual.shape if residual is not None else None,
                    bias.numel() if bias is not None else 0)
        ctx.bias_param = bias
        ctx.save_for_backward(x, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        act, res_pad, out_pad, pad_mode, res_shape, nbias = ctx.meta
        x, stats = ctx.saved_tensors
        dy = _c(dy)
        n, h, w, c = x.shape
        red = None
        if stats is not None:
            red = torch.zeros((n, c, 2), dtype=torch.float32, device=x.device)
            call("nemar_norm_act_bwd_reduce", view(x), fptr(stats), act, view(dy, out_pad), pad_mode, fptr(red),
                 stream())
        dx = torch.empty_like(x)
        dres, drv = None, None
        if res_shape is not None and ctx.needs_input_grad[2]:
            dres = (torch.zeros if res_pad > 0 else torch.empty)(res_shape, dtype=x.dtype, device=x.device)
            drv = view(dres, res_pad)
        # the bias gradient of the conv that produced x is the column sum of dx: fused into this pass
        db = None
        if nbias and ctx.needs_input_grad[7]:
            db = torch.empty(c, dtype=torch.float32, device=x.device)
        call("nemar_norm_act_bwd_apply", view(x), fptr(stats), act, view(dy, out_pad), pad_mode, fptr(red), view(dx),
             drv, 0, fptr(db), stream())
        return dx, None, dres, None, None, None, None, _deliver(ctx.bias_param, db[:nbias] if db is not None else None)


class MaxPool2Fn(Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        n, h, w, c = x.shape
        y = torch.empty((n, h // 2, w // 2, c), dtype=x.dtype, device=x.device)
        call("nemar_maxpool2_fwd", view(x), view(y), stream())
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(x)
        call("nemar_maxpool2_bwd", view(x), view(y), view(dy), view(dx), stream())
        return dx


class ResizeFn(Function):
    """bilinear, align_corners=False, on engine tensors."""

    @staticmethod
    def forward(ctx, x, oh, ow):
        x = _c(x)
        n, h, w, c = x.shape
        y = torch.empty((n, oh, ow, c), dtype=x.dtype, device=x.device)
        call("nemar_bilinear_resize_fwd", view(x), view(y), stream())
        ctx.meta = x.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        dx = torch.empty(ctx.meta, dtype=dy.dtype, device=dy.device)
        call("nemar_bilinear_resize_bwd", view(dy), view(dx), 0, stream())
        return dx, None, None


class ResizeNCHWFn(Function):
    """bilinear, align_corners=False, on NCHW fp32 images (multi-resolution discriminator inputs)."""

    @staticmethod
    def forward(ctx, x, oh, ow):
        x = _c(x)
        n, c, h, w = x.shape
        y = torch.empty((n, c, oh, ow), dtype=torch.float32, device=x.device)
        call("nemar_bilinear_resize_nchw_fwd", fptr(x), n, c, h, w, fptr(y), oh, ow, stream())
        ctx.meta = (n, c, h, w, oh, ow)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, oh, ow = ctx.meta
        dy = _c(dy)
        dx = torch.empty((n, c, h, w), dtype=torch.float32, device=dy.device)
        call("nemar_bilinear_resize_nchw_bwd", fptr(dy), n, c, h, w, fptr(dx), oh, ow, stream())
        return dx, None, None


class DropoutFn(Function):
    @staticmethod
    def forward(ctx, x, seed, offset):
        x = _c(x)
        y = torch.empty_like(x)
        call("nemar_dropout", view(x), view(y), L.u64(seed), L.u64(offset), stream())
        ctx.meta = (seed, offset)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        dx = torch.empty_like(dy)
        call("nemar_dropout", view(dy), view(dx), L.u64(ctx.meta[0]), L.u64(ctx.meta[1]), stream())
        return dx, None, None


# ------------------------------------------------------------------------------------------------
# STN head
# ------------------------------------------------------------------------------------------------
class AffineGridFn(Function):
    @staticmethod
    def forward(ctx, theta, bx, by):
        theta = _c(theta)
        n = theta.shape[0]
        h, w = by.numel(), bx.numel()
        grid = torch.empty((n, h, w, 2), dtype=torch.float32, device=theta.device)
        call("nemar_affine_grid_fwd", fptr(theta), fptr(bx), fptr(by), n, h, w, fptr(grid), stream())
        ctx.save_for_backward(bx, by)
        ctx.meta = (n, h, w)
        return grid

    @staticmethod
    def backward(ctx, dgrid):
        bx, by = ctx.saved_tensors
        n, h, w = ctx.meta
        dgrid = _c(dgrid)
        dtheta = torch.empty((n, 6), dtype=torch.float32, device=dgrid.device)
        call("nemar_affine_grid_bwd", fptr(dgrid), fptr(bx), fptr(by), n, h, w, fptr(dtheta), stream())
        return dtheta, None, None


class FlowGridFn(Function):
    """grid[N,H,W,2] = identity(xs, ys) + offsets; offsets are the engine's channels-last fp32 conv output."""

    @staticmethod
    def forward(ctx, off, xs, ys):
        off = _c(off)
        n, h, w, cs = off.shape          # cs >= 2: the conv head may pad its output channels (zeros)
        assert cs >= 2 and off.dtype == torch.float32
        grid = torch.empty((n, h, w, 2), dtype=torch.float32, device=off.device)
        call("nemar_flow_grid_fwd", fptr(off), i64(h * w * cs), i64(1), i64(w * cs), i64(cs), fptr(xs), fptr(ys), n, h, w,
             fptr(grid), stream())
        ctx.cs = cs
        return grid

    @staticmethod
    def backward(ctx, dgrid):
        if ctx.cs == 2:
            return dgrid, None, None
        dgrid = _c(dgrid)
        n, h, w, _ = dgrid.shape
        doff = torch.empty((n, h, w, ctx.cs), dtype=torch.float32, device=dgrid.device)
        call("nemar_fill_channels", view(doff), 2, ctx.cs - 2, stream())
        call("nemar_copy_view", view(dgrid), view(doff, 0, 0, 2), L.PAD_ZERO, stream())
        return doff, None, None


class GridSampleFn(Function):
    """bilinear / zeros / align_corners=False on one or two NCHW fp32 images sharing one grid."""

    @staticmethod
    def forward(ctx, grid, img0, img1):
        grid, img0 = _c(grid), _c(img0)
        n, c, h, w = img0.shape
        ho, wo = grid.shape[1], grid.shape[2]
        nimg = 1 if img1 is None else 2
        out0 = torch.empty((n, c, ho, wo), dtype=torch.float32, device=grid.device)
        out1 = None
        if nimg == 2:
            img1 = _c(img1)
            out1 = torch.empty_like(out0)
        call("nemar_grid_sample_fwd", fptr(img0), fptr(img1), nimg, n, c, h, w, fptr(grid), ho, wo, fptr(out0),
             fptr(out1), None, stream())
        ctx.save_for_backward(grid, img0, img1)
        ctx.meta = (n, c, h, w, ho, wo, nimg)
        if nimg == 1:
            return out0
        return out0, out1

    @staticmethod
    def backward(ctx, d0, d1=None):
        grid, img0, img1 = ctx.saved_tensors
        n, c, h, w, ho, wo, nimg = ctx.meta
        dev = grid.device
        d0 = _c(d0) if d0 is not None else torch.zeros((n, c, ho, wo), dtype=torch.float32, device=dev)
        if nimg == 2:
            d1 = _c(d1) if d1 is not None else torch.zeros((n, c, ho, wo), dtype=torch.float32, device=dev)
        dimg0 = torch.zeros_like(img0) if ctx.needs_input_grad[1] else None
        dimg1 = torch.zeros_like(img1) if (nimg == 2 and ctx.needs_input_grad[2]) else None
        dgrid = torch.empty_like(grid)
        call("nemar_grid_sample_bwd", fptr(img0), fptr(img1), nimg, n, c, h, w, fptr(grid), ho, wo, fptr(d0),
             fptr(d1) if nimg == 2 else None, fptr(dimg0), fptr(dimg1), fptr(dgrid), stream())
        return (dgrid if ctx.needs_input_grad[0] else None), dimg0, dimg1


def grid_sample_indices(grid, img):
    """Integer tap indices (x0,y0) the kernel uses for `grid` — the bit-exactness probe."""
    grid, img = _c(grid), _c(img)
    n, c, h, w = img.shape
    ho, wo = grid.shape[1], grid.shape[2]
    out = torch.empty((n, c, ho, wo), dtype=torch.float32, device=grid.device)
    idx = torch.empty((n, ho, wo, 2), dtype=torch.int32, device=grid.device)
    call("nemar_grid_sample_fwd", fptr(img), None, 1, n, c, h, w, fptr(grid), ho, wo, fptr(out), None,
         vptr(idx), stream())
    return out, idx


class SmoothnessFn(Function):
    """scale * smoothness_loss(offsets, img, alpha); offsets channels-last [N,H,W,2] fp32, img NCHW fp32."""

    @staticmethod
    def forward(ctx, off, img, alpha, scale):
        off = _c(off)
        n, h, w, cs = off.shape
        loss = torch.zeros(1, dtype=torch.float32, device=off.device)
        use_img = img is not None and alpha > 0.0
        if use_img:
            img = _c(img)
            assert img.shape[0] == n and img.shape[2] == h and img.shape[3] == w
        call("nemar_smoothness_fwd", fptr(off), i64(h * w * cs), i64(1), i64(w * cs), i64(cs),
             fptr(img) if use_img else None, int(img.shape[1]) if use_img else 0, float(alpha), n, h, w, float(scale),
             fptr(loss), stream())
        ctx.save_for_backward(off, img if use_img else None)
        ctx.meta = (alpha, scale)
        return loss

    @staticmethod
    def backward(ctx, g):
        off, img = ctx.saved_tensors
        alpha, scale = ctx.meta
        n, h, w, cs = off.shape
        g = _c(g.reshape(1).to(torch.float32))
        doff = torch.zeros_like(off)
        call("nemar_smoothness_bwd", fptr(off), i64(h * w * cs), i64(1), i64(w * cs), i64(cs), fptr(img),
             int(img.shape[1]) if img is not None else 0, float(alpha), n, h, w, float(scale), fptr(g), fptr(doff),
             stream())
        return doff, None, None, None


# ------------------------------------------------------------------------------------------------
# losses, linear
# ------------------------------------------------------------------------------------------------
class L1Fn(Function):
    @staticmethod
    def forward(ctx, a, b, scale):
        a, b = _c(a), _c(b)
        out = torch.zeros(1, dtype=torch.float32, device=a.device)
        call("nemar_l1_fwd", fptr(a), fptr(b), i64(a.numel()), float(scale), fptr(out), stream())
        ctx.save_for_backward(a, b)
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = _c(g.reshape(1).to(torch.float32))
        da = torch.empty_like(a)
        call("nemar_l1_bwd", fptr(a), fptr(b), i64(a.numel()), float(ctx.scale), fptr(g), fptr(da), 0, stream())
        return da, None, None


class MeanAbsFn(Function):
    @staticmethod
    def forward(ctx, a, scale):
        a = _c(a)
        out = torch.zeros(1, dtype=torch.float32, device=a.device)
        call("nemar_mean_abs_fwd", fptr(a), i64(a.numel()), float(scale), fptr(out), stream())
        ctx.save_for_backward(a)
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, g):
        (a,) = ctx.saved_tensors
        g = _c(g.reshape(1).to(torch.float32))
        da = torch.empty_like(a)
        call("nemar_mean_abs_bwd", fptr(a), i64(a.numel()), float(ctx.scale), fptr(g), fptr(da), stream())
        return da, None


class MSEConstFn(Function):
    """scale * mean((pred - target)^2) over an engine tensor (LSGAN, networks.py:237-238,273-275)."""

    @staticmethod
    def forward(ctx, pred, target, scale, c=None):
        pred = _c(pred)
        c = pred.shape[3] if c is None else c       # real channels (the head may pad its output with zeros)
        out = torch.zeros(1, dtype=torch.float32, device=pred.device)
        call("nemar_mse_const_fwd", view(pred, 0, 0, c), float(target), float(scale), fptr(out), stream())
        ctx.save_for_backward(pred)
        ctx.meta = (target, scale, c)
        return out

    @staticmethod
    def backward(ctx, g):
        (pred,) = ctx.saved_tensors
        target, scale, c = ctx.meta
        g = _c(g.reshape(1).to(torch.float32))
        dp = torch.empty_like(pred)
        if pred.shape[3] > c:
            call("nemar_fill_channels", view(dp), c, pred.shape[3] - c, stream())
        call("nemar_mse_const_bwd", view(pred, 0, 0, c), float(target), float(scale), fptr(g), view(dp, 0, 0, c), stream())
        return dp, None, None, None


class MSEConstGroupsFn(Function):
    """LSGAN terms of a batch-concatenated prediction: out[j] = scale * mean((pred[j*n:(j+1)*n] - targets[j])^2)."""

    @staticmethod
    def forward(ctx, pred, targets, scale, c=None):
        pred = _c(pred)
        k = len(targets)
        n = pred.shape[0] // k
        assert n * k == pred.shape[0]
        c = pred.shape[3] if c is None else c
        out = torch.zeros(k, dtype=torch.float32, device=pred.device)
        for j, t in enumerate(targets):
            call("nemar_mse_const_fwd", view(pred[j * n:(j + 1) * n], 0, 0, c), float(t), float(scale), fptr(out[j:j + 1]), stream())
        ctx.save_for_backward(pred)
        ctx.meta = (tuple(targets), scale, c, n)
        return out

    @staticmethod
    def backward(ctx, g):
        (pred,) = ctx.saved_tensors
        targets, scale, c, n = ctx.meta
        g = _c(g.reshape(len(targets)).to(torch.float32))
        dp = torch.empty_like(pred)
        if pred.shape[3] > c:
            call("nemar_fill_channels", view(dp), c, pred.shape[3] - c, stream())
        for j, t in enumerate(targets):
            call("nemar_mse_const_bwd", view(pred[j * n:(j + 1) * n], 0, 0, c), float(t), float(scale), fptr(g[j:j + 1]),
                 view(dp[j * n:(j + 1) * n], 0, 0, c), stream())
        return dp, None, None, None


class LinearFn(Function):
    @staticmethod
    def forward(ctx, x, w, b, act):
        x = _c(x)
        n, i = x.shape
        o = w.shape[0]
        y = torch.empty((n, o), dtype=torch.float32, device=x.device)
        call("nemar_linear_fwd", fptr(x), fptr(w.detach()), fptr(b.detach()) if b is not None else None, n, i, o, act,
             fptr(y), stream())
        ctx.save_for_backward(x, w, y)
        ctx.meta = (act, b is not None)
        ctx.params = (w, b)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        act, has_b = ctx.meta
        dy = _c(dy)
        n, i = x.shape
        o = w.shape[0]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w)
        db = torch.empty(o, dtype=torch.float32, device=x.device) if has_b else None
        call("nemar_linear_bwd", fptr(x), fptr(w.detach()), fptr(y), fptr(dy), n, i, o, act, fptr(dx), fptr(dw),
             fptr(db), stream())
        return dx, _deliver(ctx.params[0], dw), _deliver(ctx.params[1], db), None


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
    call("nemar_adam_step", fptr(p), fptr(g), fptr(m), fptr(v), i64(p.numel()), float(lr), float(beta1), float(beta2),
         float(eps), int(step), float(grad_scale), stream())


def adam_step_dev(p, g, m, v, lr, beta1, beta2, eps, step_dev, grad_scale=1.0):
    call("nemar_adam_step_dev", fptr(p), fptr(g), fptr(m), fptr(v), i64(p.numel()), float(lr), float(beta1), float(beta2),
         float(eps), vptr(step_dev), float(grad_scale), stream())
