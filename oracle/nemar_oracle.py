"""CPU oracle for the NeMAR training hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import this
module; the product (nemar_b200/) never does and has no CPU fallback.

What it is: a functional restatement, in plain fp32 PyTorch CPU ops over *state dicts*, of the algorithm of
`NEMARModel.optimize_parameters` (reference models/nemar_model.py:161-288) and everything it calls.  The
arithmetic of the reference lives in a third-party dependency that is not vendored and not pinned by the
reference (PyTorch ATen: conv2d, instance_norm, grid_sampler_2d, affine_grid_generator, upsample_bilinear2d,
Adam; reference scripts/conda_deps.sh:3 — `conda install pytorch torchvision -c pytorch`); the oracle version
for this project is the installed torch 2.11.0, the same library the reference itself would call here.

Parity pinning: the reference ships no tests or golden vectors ("parity unpinned" by its own tests), so the
oracle is pinned against outputs of the reference itself, generated in the build container by
tests/golden/make_golden.py (imports /root/reference in place, loads the SAME seeded state dicts, runs the
reference's optimize_parameters) and committed under tests/golden/*.npz; tests/test_oracle_golden.py replays
them.  The integer-index / grid arithmetic is additionally restated in plain C (oracle/grid_oracle.c).

Each function cites the reference lines it follows.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F


# =================================================================================================
# parameter enumeration (names and shapes exactly as the reference's state_dict)
# =================================================================================================
def resnet_generator_shapes(input_nc=3, output_nc=3, ngf=64, n_blocks=9, use_dropout=False):
    """reference models/networks.py:349-377 (nn.Sequential indices)"""
    s = OrderedDict()

    def conv(idx, co, ci, k):
        s["model.%s.weight" % idx] = (co, ci, k, k)
        s["model.%s.bias" % idx] = (co,)

    conv(1, ngf, input_nc, 7)
    conv(4, ngf * 2, ngf, 3)
    conv(7, ngf * 4, ngf * 2, 3)
    second = 6 if use_dropout else 5
    for i in range(n_blocks):
        conv("%d.conv_block.1" % (10 + i), ngf * 4, ngf * 4, 3)
        conv("%d.conv_block.%d" % (10 + i, second), ngf * 4, ngf * 4, 3)
    b = 10 + n_blocks
    s["model.%d.weight" % b] = (ngf * 4, ngf * 2, 3, 3)       # ConvTranspose2d: [in, out, k, k]
    s["model.%d.bias" % b] = (ngf * 2,)
    s["model.%d.weight" % (b + 3)] = (ngf * 2, ngf, 3, 3)
    s["model.%d.bias" % (b + 3)] = (ngf,)
    conv(b + 7, output_nc, ngf, 7)
    return s


def discriminator_shapes(input_nc=6, ndf=64, n_layers=3):
    """reference models/networks.py:576-597"""
    s = OrderedDict()
    s["model.0.weight"], s["model.0.bias"] = (ndf, input_nc, 4, 4), (ndf,)
    nf, idx = 1, 2
    for n in range(1, n_layers):
        nf_prev, nf = nf, min(2 ** n, 8)
        s["model.%d.weight" % idx], s["model.%d.bias" % idx] = (ndf * nf, ndf * nf_prev, 4, 4), (ndf * nf,)
        idx += 3
    nf_prev, nf = nf, min(2 ** n_layers, 8)
    s["model.%d.weight" % idx], s["model.%d.bias" % idx] = (ndf * nf, ndf * nf_prev, 4, 4), (ndf * nf,)
    idx += 3
    s["model.%d.weight" % idx], s["model.%d.bias" % idx] = (1, ndf * nf, 4, 4), (1,)
    return s


def _resblock_shapes(s, prefix, dim):
    for j in (1, 5):
        s["%s.conv_block.%d.weight" % (prefix, j)] = (dim, dim, 3, 3)
        s["%s.conv_block.%d.bias" % (prefix, j)] = (dim,)


def affine_stn_shapes(nc_a=3, nc_b=3, height=64, width=64):
    """reference models/stn/affine_stn.py:9-19,50-76 (cfg 'A')"""
    s = OrderedDict()
    prev, nf = nc_a + nc_b, 32
    for i in range(5):
        s["net.convs.%d.conv_0.conv2d.weight" % i] = (nf, prev, 3, 3)
        s["net.convs.%d.conv_0.conv2d.bias" % i] = (nf,)
        prev, nf = nf, min(2 * nf, 256)
    s["net.local.0.weight"] = (nf, prev * (height // 32) * (width // 32))
    s["net.local.0.bias"] = (nf,)
    s["net.local.2.weight"] = (6, nf)
    s["net.local.2.bias"] = (6,)
    return s


UNET_NDF = [32, 64, 64, 64, 64, 64, 64]
UNET_NUF = [64, 64, 64, 64, 64, 64, 32]


def unet_stn_shapes(nc_a=3, nc_b=3):
    """reference models/stn/unet_stn.py:12-25,31-77 (cfg 'A')"""
    s = OrderedDict()
    in_nf = nc_a + nc_b
    for i, out_nf in enumerate(UNET_NDF, start=1):
        p = "offset_map.down_%d.conv_0" % i
        s[p + ".conv2d.weight"], s[p + ".conv2d.bias"] = (out_nf, in_nf, 3, 3), (out_nf,)
        _resblock_shapes(s, p + ".resnet_block.model.0", out_nf)
        in_nf = out_nf
    s["offset_map.c1.conv2d.weight"], s["offset_map.c1.conv2d.bias"] = (2 * in_nf, in_nf, 1, 1), (2 * in_nf,)
    for j in range(3):
        _resblock_shapes(s, "offset_map.t.model.%d" % j, 2 * in_nf)
    s["offset_map.c2.conv2d.weight"], s["offset_map.c2.conv2d.bias"] = (in_nf, 2 * in_nf, 1, 1), (in_nf,)
    num = 7
    for out_nf in UNET_NUF:
        s["offset_map.up_%d.conv2d.weight" % num] = (out_nf, in_nf + UNET_NDF[num - 1], 3, 3)
        s["offset_map.up_%d.conv2d.bias" % num] = (out_nf,)
        in_nf = out_nf
        num -= 1
    _resblock_shapes(s, "offset_map.refine.0.model.0", in_nf)
    s["offset_map.refine.1.conv2d.weight"], s["offset_map.refine.1.conv2d.bias"] = (in_nf, in_nf, 1, 1), (in_nf,)
    s["offset_map.output.conv2d.weight"], s["offset_map.output.conv2d.bias"] = (2, in_nf, 3, 3), (2,)
    return s


def seeded_state(shapes, seed, weight_std=0.02, bias_std=0.01, overrides=None):
    """Deterministic test weights (NOT the reference's init RNG stream): every tensor ~ N(0, std) from a
    per-tensor generator.  `overrides` maps a key substring to a std (e.g. a live deformation head)."""
    sd = OrderedDict()
    for i, (k, shp) in enumerate(shapes.items()):
        g = torch.Generator().manual_seed(seed * 7919 + i)
        std = bias_std if k.endswith(".bias") else weight_std
        for sub, v in (overrides or {}).items():
            if sub in k:
                std = v
        sd[k] = torch.randn(shp, generator=g) * std
    return sd


def synthetic_batch(n, h, w, seed=1, c=3):
    """A,B ~ U(-1,1) from torch.Generator().manual_seed(seed)."""
    g = torch.Generator().manual_seed(seed)
    a = torch.rand((n, c, h, w), generator=g) * 2 - 1
    b = torch.rand((n, c, h, w), generator=g) * 2 - 1
    return a, b


# =================================================================================================
# networks as pure functions of (state dict, input)
# =================================================================================================
def _inorm(x):
    """nn.InstanceNorm2d(affine=False, track_running_stats=False) — networks.py:24; layers.py:16"""
    return F.instance_norm(x, eps=1e-5)


def _resblock(sd, p, x, second=5, dropout_p=0.0):
    """reference networks.py:413-446: x + IN(conv(pad(relu(IN(conv(pad x))))))"""
    y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), sd[p + ".conv_block.1.weight"], sd[p + ".conv_block.1.bias"])
    y = F.relu(_inorm(y))
    if dropout_p > 0:
        y = F.dropout(y, dropout_p, training=True)
    y = F.conv2d(F.pad(y, (1, 1, 1, 1), mode="reflect"), sd[p + ".conv_block.%d.weight" % second],
                 sd[p + ".conv_block.%d.bias" % second])
    return x + _inorm(y)


def resnet_generator(sd, x, n_blocks):
    """reference networks.py:349-386"""
    y = F.conv2d(F.pad(x, (3, 3, 3, 3), mode="reflect"), sd["model.1.weight"], sd["model.1.bias"])
    y = F.relu(_inorm(y))
    y = F.relu(_inorm(F.conv2d(y, sd["model.4.weight"], sd["model.4.bias"], stride=2, padding=1)))
    y = F.relu(_inorm(F.conv2d(y, sd["model.7.weight"], sd["model.7.bias"], stride=2, padding=1)))
    for i in range(n_blocks):
        y = _resblock(sd, "model.%d" % (10 + i), y)
    b = 10 + n_blocks
    y = F.relu(_inorm(F.conv_transpose2d(y, sd["model.%d.weight" % b], sd["model.%d.bias" % b], stride=2, padding=1,
                                         output_padding=1)))
    y = F.relu(_inorm(F.conv_transpose2d(y, sd["model.%d.weight" % (b + 3)], sd["model.%d.bias" % (b + 3)], stride=2,
                                         padding=1, output_padding=1)))
    y = F.conv2d(F.pad(y, (3, 3, 3, 3), mode="reflect"), sd["model.%d.weight" % (b + 7)], sd["model.%d.bias" % (b + 7)])
    return torch.tanh(y)


def nlayer_discriminator(sd, x, n_layers=3):
    """reference networks.py:576-602"""
    y = F.leaky_relu(F.conv2d(x, sd["model.0.weight"], sd["model.0.bias"], stride=2, padding=1), 0.2)
    idx = 2
    for _ in range(1, n_layers):
        y = F.leaky_relu(_inorm(F.conv2d(y, sd["model.%d.weight" % idx], sd["model.%d.bias" % idx], stride=2, padding=1)), 0.2)
        idx += 3
    y = F.leaky_relu(_inorm(F.conv2d(y, sd["model.%d.weight" % idx], sd["model.%d.bias" % idx], stride=1, padding=1)), 0.2)
    idx += 3
    return F.conv2d(y, sd["model.%d.weight" % idx], sd["model.%d.bias" % idx], stride=1, padding=1)


def affine_network(sd, img_a, img_b):
    """reference affine_stn.py:78-83 with DownBlock(Conv3 p1 -> IN -> ReLU -> MaxPool2) x5"""
    x = torch.cat([img_a, img_b], 1)
    for i in range(5):
        p = "net.convs.%d.conv_0.conv2d" % i
        x = F.max_pool2d(F.relu(_inorm(F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1))), 2)
    x = x.view(x.size(0), -1)
    x = F.relu(F.linear(x, sd["net.local.0.weight"], sd["net.local.0.bias"]))
    return F.linear(x, sd["net.local.2.weight"], sd["net.local.2.bias"])


def affine_stn(sd, img_a, img_b, apply_on):
    """reference affine_stn.py:108-138"""
    dtheta = affine_network(sd, img_a, img_b)
    theta = dtheta + torch.tensor([1, 0, 0, 0, 1, 0], dtype=dtheta.dtype).unsqueeze(0).repeat(img_a.size(0), 1)
    warped = []
    for img in apply_on:
        grid = F.affine_grid(theta.view(-1, 2, 3), img.size(), align_corners=False)
        warped.append(F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False))
    return warped, torch.mean(torch.abs(dtheta)), theta


def _stn_conv(sd, p, x, act, padding, resblock=False):
    """reference layers.py:99-106 with use_norm=False: conv -> act -> resblock?"""
    x = F.conv2d(x, sd[p + ".conv2d.weight"], sd[p + ".conv2d.bias"], padding=padding)
    if act == "leaky_relu":
        x = F.leaky_relu(x, 0.2)
    if resblock:
        x = _resblock(sd, p + ".resnet_block.model.0", x)
    return x


def res_unet(sd, img_a, img_b):
    """reference unet_stn.py:80-102"""
    x = torch.cat([img_a, img_b], 1)
    skips = {}
    for i in range(1, 8):
        x = skips[i] = _stn_conv(sd, "offset_map.down_%d.conv_0" % i, x, "leaky_relu", 1, resblock=True)
        x = F.max_pool2d(x, 2)
    x = _stn_conv(sd, "offset_map.c1", x, "leaky_relu", 0)
    for j in range(3):
        x = _resblock(sd, "offset_map.t.model.%d" % j, x)
    x = _stn_conv(sd, "offset_map.c2", x, "leaky_relu", 0)
    for i in range(7, 0, -1):
        s = skips[i]
        x = F.interpolate(x, (s.size(2), s.size(3)), mode="bilinear")
        x = _stn_conv(sd, "offset_map.up_%d" % i, torch.cat([x, s], 1), "leaky_relu", 1)
    x = _resblock(sd, "offset_map.refine.0.model.0", x)
    x = _stn_conv(sd, "offset_map.refine.1", x, "leaky_relu", 0)
    return _stn_conv(sd, "offset_map.output", x, None, 1)


def identity_grid(h, w):
    """reference unet_stn.py:121-129: channel 0 = x = linspace(-1,1,W) along W, channel 1 = y"""
    x = torch.linspace(-1.0, 1.0, w)
    y = torch.linspace(-1.0, 1.0, h)
    return torch.stack([x.view(1, w).expand(h, w), y.view(h, 1).expand(h, w)], 0).unsqueeze(0)


def smoothness_loss(deformation, img=None, alpha=0.0):
    """reference stn_losses.py:4-30"""
    d = deformation
    diffs = [d[:, :, 1:, :] - d[:, :, :-1, :], d[:, :, :, 1:] - d[:, :, :, :-1],
             d[:, :, :-1, :-1] - d[:, :, 1:, 1:], d[:, :, :-1, 1:] - d[:, :, 1:, :-1]]
    if img is not None and alpha > 0.0:
        m = img
        pairs = [m[:, :, 1:, :] - m[:, :, :-1, :], m[:, :, :, 1:] - m[:, :, :, :-1],
                 m[:, :, :-1, :-1] - m[:, :, 1:, 1:], m[:, :, :-1, 1:] - m[:, :, 1:, :-1]]
        weights = [torch.mean(torch.exp(-alpha * torch.abs(p)), dim=1, keepdim=True) for p in pairs]
    else:
        weights = [1.0] * 4
    return sum(torch.mean(wt * torch.abs(df)) for wt, df in zip(weights, diffs))


def unet_regularization(deformation, img, alpha, levels):
    """reference unet_stn.py:179-201"""
    dh, dw = deformation.size(2), deformation.size(3)
    img = img.detach()
    reg, factor = 0.0, 1.0
    for i in range(levels):
        if i != 0:
            size = (dh // (2 ** i), dw // (2 ** i))
            d_r = F.interpolate(deformation, size, mode="bilinear", align_corners=False)
            i_r = F.interpolate(img, size, mode="bilinear", align_corners=False)
        else:
            d_r, i_r = deformation, img
        reg = reg + factor * smoothness_loss(d_r, i_r, alpha=alpha)
        factor /= 2.0
    return reg


def unet_stn(sd, img_a, img_b, apply_on, alpha=0.0, levels=1):
    """reference unet_stn.py:148-177"""
    deformation = res_unet(sd, img_a, img_b)
    h, w = img_a.size(2), img_a.size(3)
    grid = (identity_grid(h, w).repeat(img_a.size(0), 1, 1, 1) + deformation).permute(0, 2, 3, 1)
    warped = [F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False) for img in apply_on]
    return warped, unet_regularization(deformation, warped[0], alpha, levels), grid


def lsgan(pred, target_is_real):
    """reference networks.py:237-238,261-275 (MSELoss against an expanded constant)"""
    return F.mse_loss(pred, torch.full_like(pred, 1.0 if target_is_real else 0.0))


# =================================================================================================
# the training step
# =================================================================================================
class OracleConfig:
    def __init__(self, stn_type="affine", n_blocks=6, height=64, width=64, lambda_gan=1.0, lambda_recon=100.0,
                 lambda_smooth=0.0, alpha=0.0, multires_reg=1, multi_resolution=1, lr=2e-4, beta1=0.5, ngf=64, ndf=64):
        self.__dict__.update(locals())
        del self.__dict__["self"]


def make_states(cfg, seed=0, live_head=True):
    """Seeded T / R / D(+scales) state dicts shared by the oracle, the reference (golden) and the engine."""
    T = seeded_state(resnet_generator_shapes(ngf=cfg.ngf, n_blocks=cfg.n_blocks), seed + 1)
    if cfg.stn_type == "affine":
        ov = {"net.local.2.weight": 2e-3, "net.local.2.bias": 5e-2} if live_head else None
        R = seeded_state(affine_stn_shapes(height=cfg.height, width=cfg.width), seed + 2, overrides=ov)
    else:
        ov = {"offset_map.output.conv2d.weight": 2e-2, "offset_map.output.conv2d.bias": 1e-2} if live_head else None
        R = seeded_state(unet_stn_shapes(), seed + 2, overrides=ov)
    Ds = [seeded_state(discriminator_shapes(ndf=cfg.ndf), seed + 3 + i) for i in range(cfg.multi_resolution)]
    return T, R, Ds


class Adam:
    """torch.optim.Adam update rule restated (reference nemar_model.py:128-137; amsgrad off, no decay)."""

    def __init__(self, params, lr, beta1, beta2=0.999, eps=1e-8):
        self.params = list(params)
        self.lr, self.b1, self.b2, self.eps, self.t = lr, beta1, beta2, eps, 0
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]

    def step(self, grads):
        self.t += 1
        bc1, bc2 = 1 - self.b1 ** self.t, 1 - self.b2 ** self.t
        with torch.no_grad():
            for p, g, m, v in zip(self.params, grads, self.m, self.v):
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
                p.addcdiv_(m, denom, value=-self.lr / bc1)


class OracleStep:
    """optimize_parameters (reference nemar_model.py:266-288) on explicit state dicts."""

    def __init__(self, cfg, T, R, Ds):
        self.cfg = cfg
        req = lambda sd: OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items())
        self.T, self.R, self.Ds = req(T), req(R), [req(d) for d in Ds]
        self.opt_T = Adam(self.T.values(), cfg.lr, cfg.beta1)
        self.opt_R = Adam(self.R.values(), cfg.lr, cfg.beta1)
        self.opt_D = Adam([p for d in self.Ds for p in d.values()], cfg.lr, cfg.beta1)
        self.out = {}

    def _stn(self, a, b, apply_on):
        if self.cfg.stn_type == "affine":
            return affine_stn(self.R, a, b, apply_on)
        return unet_stn(self.R, a, b, apply_on, self.cfg.alpha, self.cfg.multires_reg)

    def forward(self, A, B):
        """reference nemar_model.py:161-173"""
        nb = self.cfg.n_blocks
        fake_B = resnet_generator(self.T, A, nb)
        warped, reg, grid = self._stn(A, B, [A, fake_B])
        fake_TR_B = resnet_generator(self.T, warped[0], nb)
        self.out = dict(fake_B=fake_B, registered_real_A=warped[0], fake_TR_B=fake_TR_B, fake_RT_B=warped[1], reg=reg,
                        grid=grid)
        return self.out

    def _resize(self, x, level):
        if level == 0:
            return x
        return F.interpolate(x, (x.size(2) // 2 ** level, x.size(3) // 2 ** level), mode="bilinear", align_corners=False)

    def _gan(self, A, img, real, Ds):
        return sum(lsgan(nlayer_discriminator(d, torch.cat((self._resize(A, i), self._resize(img, i)), 1)), real)
                   for i, d in enumerate(Ds))

    def step(self, A, B):
        cfg = self.cfg
        o = self.forward(A, B)
        # ---- backward_D (reference :217-264): T,R frozen, fakes detached
        d_real = self._gan(A, B, True, self.Ds)
        d_tr = self._gan(A, o["fake_TR_B"].detach(), False, self.Ds)
        d_rt = self._gan(A, o["fake_RT_B"].detach(), False, self.Ds)
        loss_D = 0.5 * cfg.lambda_gan * (d_real + d_tr + d_rt)
        d_params = [p for d in self.Ds for p in d.values()]
        d_grads = torch.autograd.grad(loss_D, d_params)
        self.opt_D.step(d_grads)
        # ---- backward_T_and_R (reference :175-215): D frozen but UPDATED
        Ds_new = [OrderedDict((k, v.detach()) for k, v in d.items()) for d in self.Ds]
        l1_tr = cfg.lambda_recon * F.l1_loss(o["fake_TR_B"], B)
        l1_rt = cfg.lambda_recon * F.l1_loss(o["fake_RT_B"], B)
        gan_tr = cfg.lambda_gan * self._gan(A, o["fake_TR_B"], True, Ds_new)
        gan_rt = cfg.lambda_gan * self._gan(A, o["fake_RT_B"], True, Ds_new)
        smooth = cfg.lambda_smooth * o["reg"]
        loss = l1_tr + l1_rt + gan_tr + gan_rt + smooth
        tr_params = list(self.R.values()) + list(self.T.values())
        grads = torch.autograd.grad(loss, tr_params, allow_unused=True)
        grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, tr_params)]
        nR = len(self.R)
        self.grads = dict(R=grads[:nR], T=grads[nR:], D=list(d_grads))
        self.opt_R.step(grads[:nR])
        self.opt_T.step(grads[nR:])
        f = lambda t: float(t.detach()) if torch.is_tensor(t) else float(t)
        return OrderedDict(L1_TR=f(l1_tr), GAN_TR=f(gan_tr), L1_RT=f(l1_rt), GAN_RT=f(gan_rt), smoothness=f(smooth),
                           D_fake_TR=f(d_tr), D_fake_RT=f(d_rt), D=f(loss_D))


def run_in_dtype(dtype, fn):
    """Evaluate fn() with `dtype` as torch's default floating type (fp64 'truth' runs of the oracle)."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        return fn()
    finally:
        torch.set_default_dtype(old)


def cast_states(dtype, T, R, Ds):
    c = lambda sd: OrderedDict((k, v.to(dtype)) for k, v in sd.items())
    return c(T), c(R), [c(d) for d in Ds]
