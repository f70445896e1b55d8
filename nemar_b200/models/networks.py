"""Translation network (ResNet generator), PatchGAN discriminator, GAN objective and weight init, rebuilt
on the B200 engine.  Mirrors the reference's factory surface and parameter tree
(models/networks.py:62-209,215-281,323-446,556-602): same `define_G/define_D/GANLoss/init_weights` names and
arguments, same state_dict keys (`model.1.weight`, `model.10.conv_block.5.bias`, ...), so checkpoints
interchange with the reference.  Everything below nn.Module.forward is libnemar_b200.so.
"""
import torch
import torch.nn as nn

from ..engine import functional as F
from ..engine import lib as L
from ..engine.config import CONFIG


# ------------------------------------------------------------------------------------------------
# parameter holders
# ------------------------------------------------------------------------------------------------
class Holder(nn.Module):
    """Anonymous container used to reproduce the reference's attribute paths."""

    def forward(self, *a, **k):  # pragma: no cover - containers are never called
        raise RuntimeError("container module")


def tc_policy(cin, cout, k, stride, transposed, dtype, cin_p):
    """Should this layer run on the tcgen05 engine?  (bf16 storage, input channels already a multiple of 16)"""
    if CONFIG.conv_engine != "auto" or dtype != torch.bfloat16 or cin_p % 16 != 0:
        return False
    kh, kw = (k, k) if isinstance(k, int) else k
    geom = L.ConvGeom(cin, cout, kh, kw, stride, 0, int(transposed))
    return bool(L.lib().nemar_conv2d_tc_supported(L.C.byref(geom), L.BF16, 0, 0))


def image_channels(nc):
    """Channel count of the engine tensor holding `nc` image channels: padded to 16 (zeros) when the tensor-core
    engine is active so that 3/6-channel images fit a tcgen05 K chunk."""
    if CONFIG.conv_engine == "auto" and CONFIG.dtype == torch.bfloat16:
        return (nc + 15) // 16 * 16
    return nc


class Conv(nn.Module):
    """Owns `weight`/`bias` in the reference layout and runs the conv through the engine."""

    def __init__(self, cin, cout, k, stride=1, pad=0, bias=True, transposed=False, out_pad_t=0):
        super().__init__()
        shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        self.weight = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None
        nn.init.normal_(self.weight, 0.0, 0.02)
        self.meta = (cin, cout, k, stride, pad, transposed, out_pad_t)
        self._packed = F.PackedWeights()
        self._cfgs = {}

    def extra_repr(self):
        cin, cout, k, s, p, t, _ = self.meta
        return "%d->%d k%d s%d p%d%s" % (cin, cout, k, s, p, " transposed" if t else "")

    def run(self, x, x_pad=0, act=L.ACT_NONE, stats=False, out_f32=False, defer_bias_grad=False):
        return self._run(x, self.weight, self.bias, self.meta, "direct", x_pad, act, stats, out_f32, defer_bias_grad)

    def _run(self, x, weight, bias, meta, tag, x_pad, act, stats, out_f32, defer_bias_grad):
        cin, cout, k, stride, pad, transposed, opt = meta
        key = (tag, x_pad, act, stats, out_f32, x.dtype, CONFIG.conv_engine, x.shape[3], defer_bias_grad)
        ent = self._cfgs.get(key)
        if ent is None:
            use_tc = tc_policy(cin, cout, k, stride, transposed, x.dtype, x.shape[3])
            cout_p = (cout + 15) // 16 * 16 if use_tc else cout
            cfg = F.ConvCfg(cin, cout, k, stride, pad, transposed, x_pad, act, stats, out_f32, opt, use_tc, cout_p,
                            defer_bias_grad and bias is not None)
            ent = (cfg, self._packed if tag == "direct" else F.PackedWeights())
            self._cfgs[key] = ent
        # tap-transformed weights are temporaries derived from self.weight: their packs are keyed on the parameter
        return F.Conv2dFn.apply(x, weight, bias, ent[0], ent[1], None if tag == "direct" else self.weight)

    # ---- k x k convolutions with <= 4 input or output channels as 1x1 tensor-core convolutions ------------------
    def taps_supported(self, x_dtype):
        cin, cout, k, stride, pad, transposed, _ = self.meta
        return (CONFIG.conv_engine == "auto" and x_dtype == torch.bfloat16 and not transposed and stride == 1 and k > 1 and
                2 * pad == k - 1 and (cin <= 4 or cout <= 4))

    def run_head_xtaps(self, x_padded, stats=True):
        """cin <= 4, k x k on a halo'd input: the COLUMN taps are gathered into channels (k*cin of them: 21 for the
        generator head instead of the 147 of a full tap gather, 64 B per pixel instead of 320), the ROW taps run as a
        k x 1 tensor-core convolution whose TMA boxes shift by rows.  -> (y, stats) like run(stats=True)"""
        cin, cout, k, _, pad, _, _ = self.meta
        n, hp, wp, _ = x_padded.shape
        kc = k * cin
        u = F.GatherTapsFn.apply(x_padded, k, cin, 1, hp, wp - 2 * pad, (kc + 31) // 32 * 32, x_padded.dtype, 1)
        w1 = F.TapsWeightFn.apply(self.weight, "head_x")                      # [co][(b,ci)][k][1]
        return self._run(u, w1, self.bias, (kc, cout, (k, 1), 1, 0, False, 0), "head_xtaps", 0, L.ACT_NONE, stats, False, True)

    def run_tail_xtaps(self, x_padded, act, cp=4):
        """cout <= 4: a k x 1 convolution produces, per pixel of the row-valid / column-halo'd grid, the k*cout partial
        sums of the column taps; the column-tap sum (+bias, act) completes the k x k convolution.
        -> fp32 engine tensor [N,H,W,cp]"""
        cin, cout, k, _, pad, _, _ = self.meta
        n, hp, wp, _ = x_padded.shape
        kc = k * cout
        wv = F.TapsWeightFn.apply(self.weight, "tail_x")                      # [(b,co)][ci][k][1]
        v = self._run(x_padded, wv, None, (cin, kc, (k, 1), 1, 0, False, 0), "tail_xtaps", 0, L.ACT_NONE, False, False, False)
        return F.SumTapsFn.apply(v, self.bias, k, cout, -1, hp - 2 * pad, wp - 2 * pad, cp, act, 1)

    def run_head_taps(self, x_padded, stats=True):
        """cin <= 4: gather the k*k taps of the (halo'd) input into channels, then a 1x1 conv over k*k*cin channels.
        -> (y, stats) like run(stats=True, defer_bias_grad=True)"""
        cin, cout, k, _, pad, _, _ = self.meta
        n, hp, wp, _ = x_padded.shape
        kc = k * k * cin
        u = F.GatherTapsFn.apply(x_padded, k, cin, 1, hp - 2 * pad, wp - 2 * pad, (kc + 31) // 32 * 32, x_padded.dtype)
        w1 = F.TapsWeightFn.apply(self.weight, "head")                        # [co][(a,b,ci)][1][1]
        return self._run(u, w1, self.bias, (kc, cout, 1, 1, 0, False, 0), "head_taps", 0, L.ACT_NONE, stats, False, True)

    def run_tail_taps(self, x_padded, act, cp=4):
        """cout <= 4: a 1x1 conv producing k*k*cout virtual channels on the halo'd grid, then the tap sum (+bias, act).
        -> fp32 engine tensor [N,H,W,cp]"""
        cin, cout, k, _, pad, _, _ = self.meta
        n, hp, wp, _ = x_padded.shape
        kc = k * k * cout
        wv = F.TapsWeightFn.apply(self.weight, "tail")                        # [(a,b,co)][ci][1][1]
        v = self._run(x_padded, wv, None, (cin, kc, 1, 1, 0, False, 0), "tail_taps", 0, L.ACT_NONE, False, False, False)
        return F.SumTapsFn.apply(v, self.bias, k, cout, -1, hp - 2 * pad, wp - 2 * pad, cp, act)


def norm_act(x, stats, act, residual=None, res_pad=0, out_pad=0, pad_mode=L.PAD_REFLECT, bias=None):
    return F.NormActFn.apply(x, stats, residual, act, res_pad, out_pad, pad_mode, bias)


def conv_in_act(conv, x, x_pad, act, out_pad=0, residual=None, res_pad=0):
    """conv -> InstanceNorm (statistics from the conv epilogue) -> act (+residual) -> optional reflect halo.
    The conv's bias gradient is produced by the norm pass's backward (fused column sum)."""
    y, st = conv.run(x, x_pad=x_pad, stats=True, defer_bias_grad=True)
    return norm_act(y, st, act, residual, res_pad, out_pad, bias=conv.bias)


def conv_act(conv, x, x_pad, act, out_pad=0, out_f32=False):
    """conv -> act with no normalisation; a reflect halo on the result needs the separate pass"""
    if out_pad == 0:
        return conv.run(x, x_pad=x_pad, act=act, out_f32=out_f32)
    y = conv.run(x, x_pad=x_pad, defer_bias_grad=True)
    return norm_act(y, None, act, None, 0, out_pad, bias=conv.bias)


class ResnetBlock(nn.Module):
    """x + IN(conv(pad(relu(IN(conv(pad(x))))))) (reference networks.py:389-446), reflect padding."""

    def __init__(self, dim, use_dropout=False, use_bias=True):
        super().__init__()
        self.conv_block = Holder()
        self.conv_block.add_module("1", Conv(dim, dim, 3, 1, 1, bias=use_bias))
        self.second = "6" if use_dropout else "5"
        self.conv_block.add_module(self.second, Conv(dim, dim, 3, 1, 1, bias=use_bias))
        self.use_dropout = use_dropout

    def run(self, t_padded, out_pad):
        """t_padded carries a reflect halo of 1 (it is both the conv input and the skip)."""
        c1 = getattr(self.conv_block, "1")
        c2 = getattr(self.conv_block, self.second)
        if self.use_dropout and self.training:
            u = conv_in_act(c1, t_padded, 1, L.ACT_RELU, out_pad=0)
            # salt: the n-th dropout site of this step (netT runs twice per step: two masks per block), spaced so that
            # the per-element counters of two sites never overlap
            from ..engine.config import step_counter
            CONFIG.dropout_calls += 1
            u = F.DropoutFn.apply(u, CONFIG.dropout_seed, CONFIG.dropout_calls << 40, step_counter(u.device))
            u = norm_act(u, None, L.ACT_NONE, None, 0, 1)
        else:
            u = conv_in_act(c1, t_padded, 1, L.ACT_RELU, out_pad=1)
        return conv_in_act(c2, u, 1, L.ACT_NONE, out_pad=out_pad, residual=t_padded, res_pad=1)


class ResnetGenerator(nn.Module):
    """Reference networks.py:323-386.  NCHW fp32 in, NCHW fp32 out."""

    def __init__(self, input_nc, output_nc, ngf=64, use_dropout=False, n_blocks=6, use_bias=True):
        super().__init__()
        assert n_blocks >= 0
        self.n_blocks, self.input_nc, self.output_nc = n_blocks, input_nc, output_nc
        m = Holder()
        m.add_module("1", Conv(input_nc, ngf, 7, 1, 3, bias=use_bias))
        m.add_module("4", Conv(ngf, ngf * 2, 3, 2, 1, bias=use_bias))
        m.add_module("7", Conv(ngf * 2, ngf * 4, 3, 2, 1, bias=use_bias))
        for i in range(n_blocks):
            m.add_module(str(10 + i), ResnetBlock(ngf * 4, use_dropout, use_bias))
        b = 10 + n_blocks
        m.add_module(str(b), Conv(ngf * 4, ngf * 2, 3, 2, 1, bias=use_bias, transposed=True, out_pad_t=1))
        m.add_module(str(b + 3), Conv(ngf * 2, ngf, 3, 2, 1, bias=use_bias, transposed=True, out_pad_t=1))
        m.add_module(str(b + 7), Conv(ngf, output_nc, 7, 1, 3, bias=True))
        self.model = m

    def forward(self, x):
        m, nb = self.model, self.n_blocks
        g = lambda i: getattr(m, str(i))
        if g(1).taps_supported(CONFIG.dtype):
            t = F.ImagesToNHWC.apply(3, L.PAD_REFLECT, CONFIG.dtype, self.input_nc, x)
            y, st = g(1).run_head_xtaps(t) if CONFIG.k7_xtaps else g(1).run_head_taps(t)
            t = norm_act(y, st, L.ACT_RELU, bias=g(1).bias)
        else:
            t = F.ImagesToNHWC.apply(3, L.PAD_REFLECT, CONFIG.dtype, image_channels(self.input_nc), x)
            t = conv_in_act(g(1), t, 3, L.ACT_RELU)
        t = conv_in_act(g(4), t, 0, L.ACT_RELU)
        t = conv_in_act(g(7), t, 0, L.ACT_RELU, out_pad=1 if nb > 0 else 0)
        for i in range(nb):
            t = g(10 + i).run(t, out_pad=1 if i + 1 < nb else 0)
        b = 10 + nb
        t = conv_in_act(g(b), t, 0, L.ACT_RELU)
        t = conv_in_act(g(b + 3), t, 0, L.ACT_RELU, out_pad=3)
        if g(b + 7).taps_supported(t.dtype):
            y = g(b + 7).run_tail_xtaps(t, L.ACT_TANH) if CONFIG.k7_xtaps else g(b + 7).run_tail_taps(t, L.ACT_TANH)
        else:
            y = g(b + 7).run(t, x_pad=3, act=L.ACT_TANH, out_f32=True)
        return F.ToNCHW.apply(y, self.output_nc)


class NLayerDiscriminator(nn.Module):
    """70x70 PatchGAN (reference networks.py:556-602).  `forward` keeps the reference signature (NCHW in,
    NCHW out); `forward_engine` takes the images to concatenate and returns the engine-layout prediction."""

    def __init__(self, input_nc, ndf=64, n_layers=3, use_bias=True):
        super().__init__()
        self.input_nc, self.n_layers = input_nc, n_layers
        m = Holder()
        m.add_module("0", Conv(input_nc, ndf, 4, 2, 1, bias=True))
        nf, idx, self.mid = 1, 2, []
        for n in range(1, n_layers):
            nf_prev, nf = nf, min(2 ** n, 8)
            m.add_module(str(idx), Conv(ndf * nf_prev, ndf * nf, 4, 2, 1, bias=use_bias))
            self.mid.append(idx)
            idx += 3
        nf_prev, nf = nf, min(2 ** n_layers, 8)
        m.add_module(str(idx), Conv(ndf * nf_prev, ndf * nf, 4, 1, 1, bias=use_bias))
        self.mid.append(idx)
        idx += 3
        m.add_module(str(idx), Conv(ndf * nf, 1, 4, 1, 1, bias=True))
        self.last = idx
        self.model = m

    def forward_engine(self, *imgs, groups=1):
        """imgs: the images whose channel-concatenation is the input; with groups > 1, `groups` consecutive such
        tuples, evaluated as ONE pass over their batch-concatenation (prediction samples are group-major)."""
        m = self.model
        if groups > 1:
            t = F.ImageGroupsToNHWC.apply(0, L.PAD_ZERO, CONFIG.dtype, image_channels(self.input_nc), groups, *imgs)
        else:
            t = F.ImagesToNHWC.apply(0, L.PAD_ZERO, CONFIG.dtype, image_channels(self.input_nc), *imgs)
        t = getattr(m, "0").run(t, act=L.ACT_LRELU)
        for idx in self.mid:
            t = conv_in_act(getattr(m, str(idx)), t, 0, L.ACT_LRELU)
        return getattr(m, str(self.last)).run(t, out_f32=True)   # [N,h,w,>=1] fp32 (channel 0 is the prediction)

    def forward(self, x):
        return F.ToNCHW.apply(self.forward_engine(x), 1)


# ------------------------------------------------------------------------------------------------
# GAN objective
# ------------------------------------------------------------------------------------------------
class GANLoss(nn.Module):
    """Reference networks.py:215-281.  lsgan is the engine path (MSE against a constant, fused);
    'vanilla'/'wgangp' are accepted for interface parity and evaluated with torch ops (not on the
    benchmarked configurations, which all pass --gan_mode lsgan)."""

    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0):
        super().__init__()
        if gan_mode not in ("lsgan", "vanilla", "wgangp"):
            raise NotImplementedError("gan mode %s not implemented" % gan_mode)
        self.register_buffer("real_label", torch.tensor(target_real_label))
        self.register_buffer("fake_label", torch.tensor(target_fake_label))
        self.gan_mode = gan_mode
        self.real_value, self.fake_value = float(target_real_label), float(target_fake_label)

    def engine(self, prediction, target_is_real):
        """LSGAN term on the discriminator's engine-layout output [N,h,w,C>=1] (channel 0 real, rest padding)."""
        target = self.real_value if target_is_real else self.fake_value
        if self.gan_mode == "lsgan":
            return F.MSEConstFn.apply(prediction, target, 1.0, 1).squeeze(0)
        return self.__call__(F.ToNCHW.apply(prediction, 1), target_is_real)

    def engine_groups(self, prediction, targets_are_real):
        """LSGAN terms of a group-major batch-concatenated prediction -> tensor [len(targets_are_real)]."""
        assert self.gan_mode == "lsgan"
        targets = [self.real_value if t else self.fake_value for t in targets_are_real]
        return F.MSEConstGroupsFn.apply(prediction, targets, 1.0, 1)

    def __call__(self, prediction, target_is_real):
        target = self.real_value if target_is_real else self.fake_value
        if self.gan_mode == "lsgan":
            if prediction.dim() == 4 and prediction.shape[1] == 1 and prediction.shape[3] != 1:
                prediction = prediction.permute(0, 2, 3, 1)   # NCHW [N,1,h,w] -> same memory as [N,h,w,1]
            return F.MSEConstFn.apply(prediction.contiguous(), target, 1.0).squeeze(0)
        if self.gan_mode == "vanilla":
            return torch.nn.functional.binary_cross_entropy_with_logits(prediction, torch.full_like(prediction, target))
        return -prediction.mean() if target_is_real else prediction.mean()


# ------------------------------------------------------------------------------------------------
# init + factories
# ------------------------------------------------------------------------------------------------
def init_weights(net, init_type="normal", init_gain=0.02):
    """Reference networks.py:62-95: conv/linear weights by `init_type`, biases zero."""
    def init_func(m):
        if isinstance(m, (Conv, nn.Linear)) and getattr(m, "weight", None) is not None:
            if init_type == "normal":
                nn.init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == "xavier":
                nn.init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == "kaiming":
                nn.init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                nn.init.orthogonal_(m.weight.data, gain=init_gain)
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if m.bias is not None:
                nn.init.constant_(m.bias.data, 0.0)

    print("initialize network with %s" % init_type)
    net.apply(init_func)
    F.bump_weights_epoch()


def init_net(net, init_type="normal", init_gain=0.02, gpu_ids=()):
    """One replica per process: no DataParallel wrapper (reference networks.py:98-113)."""
    init_weights(net, init_type, init_gain)
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.to(torch.device("cuda", torch.cuda.current_device()))
    return net


def define_G(input_nc, output_nc, ngf, netG, norm="instance", use_dropout=False, init_type="normal", init_gain=0.02,
             gpu_ids=()):
    if norm != "instance":
        raise NotImplementedError("the B200 engine implements --norm instance (the NeMAR default); got [%s]" % norm)
    blocks = {"resnet_9blocks": 9, "resnet_6blocks": 6, "resnet_5blocks": 5, "resnet_4blocks": 4, "resnet_3blocks": 3}
    if netG not in blocks:
        raise NotImplementedError("Generator model name [%s] is not recognized" % netG)
    net = ResnetGenerator(input_nc, output_nc, ngf, use_dropout=use_dropout, n_blocks=blocks[netG])
    return init_net(net, init_type, init_gain, gpu_ids)


def define_D(input_nc, ndf, netD, n_layers_D=3, norm="instance", init_type="normal", init_gain=0.02, gpu_ids=()):
    if norm != "instance":
        raise NotImplementedError("the B200 engine implements --norm instance (the NeMAR default); got [%s]" % norm)
    if netD == "basic":
        net = NLayerDiscriminator(input_nc, ndf, n_layers=3)
    elif netD == "n_layers":
        net = NLayerDiscriminator(input_nc, ndf, n_layers=n_layers_D)
    else:
        raise NotImplementedError("Discriminator model name [%s] is not recognized" % netD)
    return init_net(net, init_type, init_gain, gpu_ids)


def get_scheduler(optimizer, opt):
    """Reference networks.py:32-59 builds torch schedulers that train.py never steps; the engine keeps the
    attribute (a callable returning the lr multiplier) so BaseModel.update_learning_rate works."""
    if opt.lr_policy == "linear":
        def rule(epoch):
            return 1.0 - max(0, epoch + opt.epoch_count - opt.niter) / float(opt.niter_decay + 1)
    elif opt.lr_policy == "step":
        def rule(epoch):
            return 0.1 ** (epoch // opt.lr_decay_iters)
    elif opt.lr_policy == "cosine":
        import math

        def rule(epoch):
            return 0.5 * (1.0 + math.cos(math.pi * epoch / max(opt.niter, 1)))
    else:
        raise NotImplementedError("learning rate policy [%s] is not implemented" % opt.lr_policy)
    return LambdaSchedule(optimizer, rule)


class LambdaSchedule:
    def __init__(self, optimizer, rule):
        self.optimizer, self.rule, self.epoch = optimizer, rule, 0
        self.base_lr = optimizer.param_groups[0]["lr"]

    def step(self, *_):
        self.epoch += 1
        for g in self.optimizer.param_groups:
            g["lr"] = self.base_lr * self.rule(self.epoch)
