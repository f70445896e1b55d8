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


class _ZeroArena:
    """One pre-zeroed fp32 buffer per training step for the many small accumulators of the path (InstanceNorm
    statistics and backward sums of ~80 layers, loss scalars): ONE fill per step instead of ~250 torch.zeros launches.
    Active only between begin_step() and end_step(); outside a step (direct op calls, inference) zeros() falls back to
    torch.zeros.  Slices live until the next begin_step — i.e. for the step whose autograd graph holds them."""

    def __init__(self):
        self.buf, self.pos, self.high, self.active = None, 0, 0, False

    def begin(self, device):
        if self.buf is None or self.buf.device != torch.device(device):
            self.buf = torch.zeros(8 << 20, dtype=torch.float32, device=device)      # 32 MB
            self.high = 0
        elif self.high > 0:
            self.buf[:self.high].zero_()
        self.pos, self.active = 0, True

    def end(self):
        self.active = False

    def zeros(self, shape, device):
        n = 1
        for d in shape:
            n *= int(d)
        n4 = (n + 3) // 4 * 4
        if not self.active or self.buf.device != torch.device(device) or self.pos + n4 > self.buf.numel():
            return torch.zeros(shape, dtype=torch.float32, device=device)
        t = self.buf[self.pos:self.pos + n].view(shape)
        self.pos += n4
        self.high = max(self.high, self.pos)
        return t


ARENA = _ZeroArena()

_WGRAD = {"stream": None, "pending": [], "used": False}


def _wgrad_stream(device):
    """The stream weight gradients run on (None: disabled / not a CUDA device)."""
    from .config import CONFIG
    if not CONFIG.wgrad_stream or device.type != "cuda":
        return None
    s = _WGRAD["stream"]
    if s is None or s.device != device:
        s = _WGRAD["stream"] = torch.cuda.Stream(device)
    return s


def join_wgrad_stream():
    """Order the current stream behind every weight gradient issued so far (called before an optimizer consumes the
    gradient bucket) and release the operands that were kept alive for them."""
    if _WGRAD["used"]:
        torch.cuda.current_stream().wait_stream(_WGRAD["stream"])
        _WGRAD["used"] = False
    _WGRAD["pending"].clear()


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
        # k: an int (square kernel) or (kh, kw) — rectangular kernels only as valid stride-1 convolutions
        kh, kw = (k, k) if isinstance(k, int) else (int(k[0]), int(k[1]))
        self.geom = L.ConvGeom(cin, cout, kh, kw, stride, pad, int(transposed))
        self.kh, self.kw = kh, kw
        self.stats_ws = {}           # input shape -> bytes of per-tile statistics partials (0: separate pass)
        self.cin, self.cout, self.k, self.stride, self.pad = cin, cout, kh, stride, pad
        self.transposed, self.x_pad, self.act, self.stats = transposed, x_pad, act, stats
        self.out_f32, self.out_pad_t, self.use_tc = out_f32, out_pad_t, use_tc
        self.cout_p = cout if cout_p is None else cout_p
        # True: the bias gradient is produced by the NormActFn consuming this conv's output (fused column sum)
        self.defer_bias_grad = defer_bias_grad

    def out_hw(self, h, w):
        if not self.transposed:
            f = lambda v, k: (v + 2 * self.pad - k) // self.stride + 1
        else:
            f = lambda v, k: (v - 1) * self.stride - 2 * self.pad + k + self.out_pad_t
        return f(h, self.kh), f(w, self.kw)


class _PackJob(L.C.Structure):
    _fields_ = [("w", L.C.c_void_p), ("out", L.C.c_void_p), ("O", L.C.c_int32), ("op", L.C.c_int32), ("I", L.C.c_int32),
                ("ip", L.C.c_int32), ("kh", L.C.c_int32), ("kw", L.C.c_int32), ("w_is_oi", L.C.c_int32), ("flip", L.C.c_int32),
                ("dtype", L.C.c_int32), ("kind", L.C.c_int32)]


import weakref  # noqa: E402

_PACK_OWNERS = weakref.WeakSet()      # every live PackedWeights (they die with their Conv module)
_PACK_PLANS = {}                      # flat buffer address -> (key, jobs tensor, blocks tensor, nblocks)


class PackedWeights:
    """Per-parameter cache of the forward / backward packs and the padded bias.  The pack BUFFERS are allocated once
    and refreshed in place: lazily by a per-layer call when the entry is stale (first use, load_state_dict, any foreign
    in-place write), or — on the training path — by ONE multi-tensor launch per optimizer step (refresh_packs)."""

    def __init__(self):
        self.cache = {}
        _PACK_OWNERS.add(self)

    @staticmethod
    def _stamp(weight, bias, owner):
        # epoch: bumped by the engine's loaders / init; _version catches any other in-place write through torch;
        # the flat Adam kernel writes through raw pointers and refreshes the packs itself (refresh_packs)
        src = owner if owner is not None else weight
        return (weights_epoch(), src.data_ptr(), src._version, bias._version if bias is not None else 0)

    def get(self, weight, bias, cfg, dtype, cin_p, owner=None):
        """owner: the Parameter the weight tensor was derived from (tap-transformed heads / tails); entries with an
        owner are refreshed lazily only"""
        key = (dtype, cin_p, cfg.cout_p)
        ent = self.cache.get(key)
        stamp = self._stamp(weight, bias, owner)
        if ent is None:
            k2 = cfg.kh * cfg.kw
            ent = {"wf": torch.empty(cfg.cout_p * k2 * cin_p, dtype=dtype, device=weight.device),
                   "wd": torch.empty(cin_p * k2 * cfg.cout_p, dtype=dtype, device=weight.device),
                   "bp": None, "stamp": None, "cfg": cfg, "cin_p": cin_p, "weight": None, "bias": None, "derived": owner is not None}
            if bias is not None and cfg.cout_p != cfg.cout:
                ent["bp"] = torch.zeros(cfg.cout_p, dtype=torch.float32, device=weight.device)
            self.cache[key] = ent
        if ent["stamp"] != stamp:
            call("nemar_pack_weights", fptr(weight.detach()), cfg.geom, L.dtype_code(ent["wf"]), cin_p, cfg.cout_p,
                 vptr(ent["wf"]), vptr(ent["wd"]), stream())
            if ent["bp"] is not None:
                ent["bp"][:cfg.cout].copy_(bias.detach())
            ent["stamp"] = stamp
            ent["weight"] = (owner if owner is not None else weight)
            ent["bias"] = bias
        bp = ent["bp"] if ent["bp"] is not None else (bias.detach() if bias is not None else None)
        return ent["wf"], ent["wd"], bp


def _pack_jobs(ent):
    """the nemar_pack_job records of one cache entry (same mapping as nemar_pack_weights, conv_api.cu)"""
    cfg, cin_p, w = ent["cfg"], ent["cin_p"], ent["weight"]
    code = L.dtype_code(ent["wf"])
    t = bool(cfg.transposed)
    jobs = [_PackJob(w.data_ptr(), ent["wf"].data_ptr(), cfg.cout, cfg.cout_p, cfg.cin, cin_p, cfg.kh, cfg.kw, 0 if t else 1, 1 if t else 0, code, 1),
            _PackJob(w.data_ptr(), ent["wd"].data_ptr(), cfg.cin, cin_p, cfg.cout, cfg.cout_p, cfg.kh, cfg.kw, 1 if t else 0, 0 if t else 1, code, 1)]
    if ent["bp"] is not None:
        jobs.append(_PackJob(ent["bias"].data_ptr(), ent["bp"].data_ptr(), cfg.cout, cfg.cout_p, 1, 1, 1, 1, 1, 0, L.F32, 2))
    return jobs


def refresh_packs(flat_p):
    """After an optimizer step: re-pack, in ONE launch, every cached weight pack whose parameter lives in `flat_p`
    (the optimizer's flat buffer), and mark those entries current.  Entries derived from a parameter through a host-side
    transform (k7 heads / tails) are invalidated instead and re-packed by their next use."""
    lo = flat_p.data_ptr()
    hi = lo + flat_p.numel() * 4
    ents = []
    for owner in list(_PACK_OWNERS):
        for ent in owner.cache.values():
            w = ent["weight"]
            if w is None or not (lo <= w.data_ptr() < hi):
                continue
            if ent["derived"]:
                ent["stamp"] = None
            else:
                ents.append(ent)
    if not ents:
        return
    key = tuple((id(e), e["wf"].data_ptr(), e["weight"].data_ptr()) for e in ents)
    plan = _PACK_PLANS.get(lo)
    if plan is None or plan[0] != key:
        jobs = [j for e in ents for j in _pack_jobs(e)]
        blocks = []
        for ji, j in enumerate(jobs):
            total = j.op * j.kh * j.kw * j.ip
            blocks += [(ji, b) for b in range((total + 2047) // 2048)]
        arr = (_PackJob * len(jobs))(*jobs)
        jt = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(flat_p.device)
        bt = torch.tensor(blocks, dtype=torch.int32).reshape(-1).to(flat_p.device)
        plan = (key, jt, bt, len(blocks))
        _PACK_PLANS[lo] = plan
    call("nemar_pack_weights_multi", vptr(plan[1]), vptr(plan[2]), plan[3], stream())
    for e in ents:
        e["stamp"] = PackedWeights._stamp(e["weight"], e["bias"], None)


def _cast(t, dtype):
    out = torch.empty(t.shape, dtype=dtype, device=t.device)
    call("nemar_cast_view", view(t), view(out), stream())
    return out


class Conv2dFn(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, cfg, packed, owner=None):
        x = _c(x)
        n, hp, wp, cin_p = x.shape
        h, w = hp - 2 * cfg.x_pad, wp - 2 * cfg.x_pad
        ho, wo = cfg.out_hw(h, w)
        wf, wd, bp = packed.get(weight, bias, cfg, x.dtype, cin_p, owner)
        ydt = torch.float32 if cfg.out_f32 else x.dtype
        y = torch.empty((n, ho, wo, cfg.cout_p), dtype=ydt, device=x.device)
        stats = ws = None
        ws_bytes = 0
        if cfg.stats:
            xv, yv = view(x, cfg.x_pad), view(y)
            key = (tuple(x.shape), cfg.x_pad)
            ws_bytes = cfg.stats_ws.get(key)
            if ws_bytes is None:      # can the conv's epilogue produce the statistics (per-tile partials)?
                ws_bytes = int(L.lib().nemar_conv2d_fprop_stats_workspace(L.C.byref(xv), cin_p, L.C.byref(cfg.geom), L.C.byref(yv),
                                                                           int(cfg.use_tc)))
                cfg.stats_ws[key] = ws_bytes
            if ws_bytes > 0 and not L.TIMER.on:
                stats = torch.empty((n, cfg.cout_p, 2), dtype=torch.float32, device=x.device)     # overwritten by the finalize
                ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=x.device)
            else:
                ws_bytes = 0
                stats = ARENA.zeros((n, cfg.cout_p, 2), x.device)
        # while the conv launches are being timed (bench.py's roofline pass) the InstanceNorm statistics come from the
        # separate pass, issued as its own call, so that the events around the conv call bracket the convolution kernel alone
        split_stats = stats is not None and L.TIMER.on
        call("nemar_conv2d_fprop_ws", view(x, cfg.x_pad), vptr(wf), cin_p, fptr(bp), cfg.geom, cfg.act, view(y),
             fptr(None if split_stats else stats), fptr(ws), i64(ws_bytes), int(cfg.use_tc), stream())
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
            xv, gv = view(x, cfg.x_pad), view(gk)
            ws_bytes = L.lib().nemar_conv2d_wgrad_workspace(L.C.byref(xv), L.C.byref(gv), L.C.byref(cfg.geom),
                                                            int(cfg.use_tc))
            side = _wgrad_stream(x.device) if direct else None
            if side is not None:
                # off the critical path: the gradient lands in the optimizer's bucket, which only the Adam launch reads
                # (join_wgrad_stream).  Operands stay referenced until that join: they may outlive this autograd node.
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    ws = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=x.device) if ws_bytes > 0 else None
                    call("nemar_conv2d_wgrad", xv, gv, cfg.geom, fptr(weight.grad), vptr(ws), i64(ws_bytes), int(cfg.use_tc), 1,
                         stream())
                _WGRAD["pending"].append((x, gk, ws))
                _WGRAD["used"] = True
            else:
                dw = weight.grad if direct else torch.empty(weight.shape, dtype=torch.float32, device=x.device)
                ws = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=x.device) if ws_bytes > 0 else None
                call("nemar_conv2d_wgrad", xv, gv, cfg.geom, fptr(dw), vptr(ws), i64(ws_bytes), int(cfg.use_tc), int(direct),
                     stream())
                if direct:
                    dw = None
        if ctx.has_bias and ctx.needs_input_grad[2] and not cfg.defer_bias_grad:
            db = torch.empty(cfg.cout_p, dtype=torch.float32, device=x.device)
            call("nemar_bias_grad", view(g), fptr(db), stream())
            db = _deliver(ctx.bias_param, db[:cfg.cout])
        return dx, dw, db, None, None, None


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
        if mode == "tail":
            return weight.detach().permute(2, 3, 0, 1).reshape(k * k * co, ci, 1, 1).contiguous()
        # column taps folded into channels, row taps kept: a k x 1 convolution
        if mode == "head_x":       # [co][ci][a][b] -> [co][(b,ci)][a][1]
            return weight.detach().permute(0, 3, 1, 2).reshape(co, k * ci, k, 1).contiguous()
        assert mode == "tail_x"    # [co][ci][a][b] -> [(b,co)][ci][a][1]
        return weight.detach().permute(3, 0, 1, 2).reshape(k * co, ci, k, 1).contiguous()

    @staticmethod
    def backward(ctx, g):
        mode, co, ci, k = ctx.meta
        if mode == "head":
            gw = g.reshape(co, k, k, ci).permute(0, 3, 1, 2)
        elif mode == "tail":
            gw = g.reshape(k, k, co, ci).permute(2, 3, 0, 1)
        elif mode == "head_x":
            gw = g.reshape(co, k, ci, k).permute(0, 2, 3, 1)       # [co][b][ci][a] -> [co][ci][a][b]
        else:
            gw = g.reshape(k, co, ci, k).permute(1, 2, 3, 0)       # [b][co][ci][a] -> [co][ci][a][b]
        return _deliver(ctx.param, gw), None


class GatherTapsFn(Function):
    """dst[n,y,x,(a*k+b)*c+ch] = src[n, y+sgn*a, x+sgn*b, ch] — turns a k x k conv over c (<=4) channels into a
    1x1 conv over k*k*c channels (generator head, networks.py:349-350)."""

    @staticmethod
    def forward(ctx, x, k, c, sgn, oh, ow, cp, out_dtype, ky=None):
        """ky: tap rows (default k: a square window); ky = 1 gathers the column taps only"""
        x = _c(x)
        n = x.shape[0]
        ky = k if ky is None else ky
        out = torch.empty((n, oh, ow, cp), dtype=out_dtype, device=x.device)
        call("nemar_gather_taps2", view(x), view(out), ky, k, c, sgn, stream())
        ctx.meta = (ky, k, c, sgn, x.shape, x.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        ky, k, c, sgn, shape, dtype = ctx.meta
        g = _c(g)
        dx = torch.empty(shape, dtype=dtype, device=g.device)
        call("nemar_sum_taps2", view(g), view(dx), ky, k, c, sgn, None, L.ACT_NONE, stream())
        return dx, None, None, None, None, None, None, None, None


class SumTapsFn(Function):
    """dst[n,y,x,ch] = act(bias[ch] + sum_{a,b} src[n, y-sgn*a, x-sgn*b, (a*k+b)*c+ch]) — the tap sum that completes
    a k x k conv with c (<=4) output channels computed as a 1x1 conv with k*k*c virtual channels (generator tail,
    networks.py:375-377)."""

    @staticmethod
    def forward(ctx, x, bias, k, c, sgn, oh, ow, cp, act, ky=None):
        x = _c(x)
        n = x.shape[0]
        ky = k if ky is None else ky
        y = torch.empty((n, oh, ow, cp), dtype=torch.float32, device=x.device)
        call("nemar_sum_taps2", view(x), view(y), ky, k, c, sgn, fptr(bias.detach()) if bias is not None else None, act, stream())
        ctx.meta = (ky, k, c, sgn, x.shape, x.dtype, act, bias is not None)
        ctx.bias_param = bias
        ctx.save_for_backward(y if act != L.ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        ky, k, c, sgn, shape, dtype, act, has_bias = ctx.meta
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
        call("nemar_gather_taps2", view(g), view(dx), ky, k, c, sgn, stream())
        return dx, _deliver(ctx.bias_param, db), None, None, None, None, None, None, None, None


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
            red = ARENA.zeros((n, c, 2), x.device)
            call("nemar_norm_act_bwd_reduce", view(x), fptr(stats), act, view(dy, out_pad), pad_mode, fptr(red),
                 stream())
        dx = torch.empty_like(x)
        dres, drv, flags = None, None, 0
        if res_shape is not None and ctx.needs_input_grad[2]:
            dres = torch.empty(res_shape, dtype=x.dtype, device=x.device)
            drv = view(dres, res_pad)
            if res_pad > 0:
                flags |= 4                      # the pass zeroes the halo ring of dres (it writes only the interior)
        # the bias gradient of the conv that produced x is the column sum of dx: fused into this pass, and accumulated
        # straight into the optimizer's gradient bucket when the parameter owns one
        db, bias, direct = None, ctx.bias_param, False
        if nbias and ctx.needs_input_grad[7] and bias.requires_grad:
            direct = bias.is_leaf and bias.grad is not None and nbias == c and bias.grad.is_contiguous()
            db = bias.grad if direct else torch.empty(c, dtype=torch.float32, device=x.device)
            if direct:
                flags |= 2
        call("nemar_norm_act_bwd_apply", view(x), fptr(stats), act, view(dy, out_pad), pad_mode, fptr(red), view(dx),
             drv, flags, fptr(db), stream())
        dbias = None if (db is None or direct) else _deliver(bias, db[:nbias])
        return dx, None, dres, None, None, None, None, dbias


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
    """Dropout(0.5) (networks.py:427-428).  The mask is a hash of (seed, salt, step): `salt` identifies the call site
    within a step (host constant), the step number is read by the kernel from device memory, so a captured step draws
    a fresh mask on every replay and the backward pass regenerates the forward's mask."""

    @staticmethod
    def forward(ctx, x, seed, salt, step_dev):
        x = _c(x)
        y = torch.empty_like(x)
        call("nemar_dropout_dev", view(x), view(y), L.u64(seed), L.u64(salt), vptr(step_dev), stream())
        ctx.meta = (seed, salt)
        ctx.step_dev = step_dev
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        dx = torch.empty_like(dy)
        call("nemar_dropout_dev", view(dy), view(dx), L.u64(ctx.meta[0]), L.u64(ctx.meta[1]), vptr(ctx.step_dev), stream())
        return dx, None, None, None


def begin_step(device):
    """Start of one optimize_parameters: advance the device step counter (a 1-thread kernel, captured with the step)
    and restart the per-step numbering of the dropout call sites."""
    from .config import CONFIG, step_counter
    c = step_counter(device)
    call("nemar_counter_add", vptr(c), i64(1), stream())
    CONFIG.dropout_calls = 0
    ARENA.begin(device)


def end_step():
    ARENA.end()


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
        loss = ARENA.zeros((1,), off.device)
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
        out = ARENA.zeros((1,), a.device)
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
        out = ARENA.zeros((1,), a.device)
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
        out = ARENA.zeros((1,), pred.device)
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
        out = ARENA.zeros((k,), pred.device)
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
