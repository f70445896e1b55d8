"""Dense-deformation STN on the B200 engine (reference models/stn/unet_stn.py:12-201): a 7-down / 7-up
residual U-Net regresses a 2-channel offset field; grid = linspace identity + offsets; bilinear
grid_sample; multi-resolution (bilateral) smoothness regulariser."""
import torch
import torch.nn as nn

from ...engine import functional as F
from ...engine import lib as L
from ...engine.config import CONFIG
from .. import networks as N
from .affine_stn import sample_all
from .layers import Conv, DownBlock, ResnetTransformer

ndf = {"A": [32, 64, 64, 64, 64, 64, 64]}
nuf = {"A": [64, 64, 64, 64, 64, 64, 32]}
use_down_resblocks = {"A": True}
resnet_nblocks = {"A": 3}
refine_output = {"A": True}
down_activation = {"A": "leaky_relu"}
up_activation = {"A": "leaky_relu"}


class ResUnet(nn.Module):
    def __init__(self, nc_a, nc_b, cfg, init_func, init_to_identity):
        super().__init__()
        act = down_activation[cfg]
        self.ndown_blocks, self.nup_blocks = len(ndf[cfg]), len(nuf[cfg])
        assert self.ndown_blocks >= self.nup_blocks
        self.in_nc = in_nf = nc_a + nc_b
        skip_nf = {}
        for i, out_nf in enumerate(ndf[cfg], start=1):
            setattr(self, "down_%d" % i, DownBlock(in_nf, out_nf, 3, 1, 1, activation=act, init_func=init_func,
                                                   bias=True, use_resnet=use_down_resblocks[cfg], use_norm=False))
            skip_nf[i] = out_nf
            in_nf = out_nf
        self.has_t = use_down_resblocks[cfg]
        if self.has_t:
            self.c1 = Conv(in_nf, 2 * in_nf, 1, 1, 0, activation=act, init_func=init_func, bias=True)
            self.t = ResnetTransformer(2 * in_nf, resnet_nblocks[cfg], init_func) if resnet_nblocks[cfg] > 0 else None
            self.c2 = Conv(2 * in_nf, in_nf, 1, 1, 0, activation=act, init_func=init_func, bias=True)
        act = up_activation[cfg]
        conv_num = self.ndown_blocks
        for out_nf in nuf[cfg]:
            # the reference passes `init_fun=` (sic) here, so these convs get the Conv default ('kaiming')
            setattr(self, "up_%d" % conv_num, Conv(in_nf + skip_nf[conv_num], out_nf, 3, 1, 1, bias=True,
                                                   activation=act, init_func="kaiming"))
            in_nf = out_nf
            conv_num -= 1
        self.has_refine = refine_output[cfg]
        if self.has_refine:
            self.refine = N.Holder()
            self.refine.add_module("0", ResnetTransformer(in_nf, 1, init_func))
            self.refine.add_module("1", Conv(in_nf, in_nf, 1, 1, 0, init_func=init_func, activation=act))
        self.output = Conv(in_nf, 2, 3, 1, 1, bias=True, init_func=("zeros" if init_to_identity else init_func),
                           activation=None)

    def forward(self, img_a, img_b):
        """-> offsets as an engine tensor [N,H,W,2] fp32 (channels-last == the sampling-grid layout)."""
        x = F.ImagesToNHWC.apply(0, L.PAD_ZERO, CONFIG.dtype, N.image_channels(self.in_nc), img_a, img_b)
        skips = {}
        for i in range(1, self.ndown_blocks + 1):
            x, skips[i] = getattr(self, "down_%d" % i).run(x)
        if self.has_t:
            x = self.c1.run(x, out_pad=1 if self.t is not None else 0)
            if self.t is not None:
                x = self.t.run(x, out_pad=0)
            x = self.c2.run(x)
        last_up = self.ndown_blocks - self.nup_blocks + 1
        for i in range(self.ndown_blocks, self.ndown_blocks - self.nup_blocks, -1):
            s = skips[i]
            if x.shape[1] != s.shape[1] or x.shape[2] != s.shape[2]:
                x = F.ResizeFn.apply(x, s.shape[1], s.shape[2])
            x = F.Concat.apply(x, s)
            x = getattr(self, "up_%d" % i).run(x, out_pad=1 if (self.has_refine and i == last_up) else 0)
        if self.has_refine:
            x = getattr(self.refine, "0").run(x, out_pad=0)
            x = getattr(self.refine, "1").run(x)
        return self.output.run(x, out_f32=True)


class UnetSTN(nn.Module):
    def __init__(self, in_channels_a, in_channels_b, height, width, cfg, init_func, stn_bilateral_alpha,
                 init_to_identity, multi_resolution_regularization):
        super().__init__()
        self.oh, self.ow = height, width
        self.offset_map = ResUnet(in_channels_a, in_channels_b, cfg, init_func, init_to_identity)
        # the reference's identity grid: linspace(-1,1,n) (the align_corners=True identity, used as is)
        self.register_buffer("grid_xs", torch.linspace(-1.0, 1.0, self.ow), persistent=False)
        self.register_buffer("grid_ys", torch.linspace(-1.0, 1.0, self.oh), persistent=False)
        self.alpha = stn_bilateral_alpha
        self.multi_resolution_regularization = multi_resolution_regularization

    def _offsets(self, img_a, img_b):
        deformation = self.offset_map(img_a, img_b)
        up = deformation
        if deformation.size(1) != self.oh and deformation.size(2) != self.ow:   # reference: `and` (unet_stn.py:165)
            up = F.ResizeFn.apply(deformation, self.oh, self.ow)
        return deformation, up

    def get_grid(self, img_a, img_b, return_offsets_only=False):
        _, up = self._offsets(img_a, img_b)
        if return_offsets_only:
            return up
        return F.FlowGridFn.apply(up, self.grid_xs, self.grid_ys)

    def precompute(self, img_a, img_b):
        """Everything that does not depend on `apply_on` (the ResUnet and the sampling grid): NEMARModel runs it on a
        second stream while the translation network processes real_A."""
        deformation, up = self._offsets(img_a, img_b)
        return deformation, F.FlowGridFn.apply(up, self.grid_xs, self.grid_ys)

    def forward(self, img_a, img_b, apply_on=None, pre=None):
        deformation, grid = pre if pre is not None else self.precompute(img_a, img_b)
        if apply_on is None:
            apply_on = [img_a]
        warped = sample_all(lambda h, w, dev: grid, apply_on)
        reg_term = self._calculate_regularization_term(deformation, warped[0])
        return warped, reg_term

    def _calculate_regularization_term(self, deformation, img):
        dh, dw = deformation.size(1), deformation.size(2)
        img = None if img is None else img.detach()
        reg, factor = None, 1.0
        for i in range(self.multi_resolution_regularization):
            if i != 0:
                d_r = F.ResizeFn.apply(deformation, dh // (2 ** i), dw // (2 ** i))
                i_r = F.ResizeNCHWFn.apply(img, dh // (2 ** i), dw // (2 ** i)) if img is not None else None
            elif img is not None and (img.size(2) != dh or img.size(3) != dw):
                d_r, i_r = deformation, F.ResizeNCHWFn.apply(img, dh, dw)
            else:
                d_r, i_r = deformation, img
            term = F.SmoothnessFn.apply(d_r, i_r, float(self.alpha), factor).squeeze(0)
            reg = term if reg is None else reg + term
            factor /= 2.0
        return reg
