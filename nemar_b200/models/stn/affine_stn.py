"""Affine STN on the B200 engine (reference models/stn/affine_stn.py:9-138): a 5-stage conv/IN/ReLU/pool
encoder, two Linear layers regressing d_theta, theta = I + d_theta, affine_grid + bilinear grid_sample of
every `apply_on` image with ONE grid read, reg = mean|d_theta|."""
import torch
import torch.nn as nn

from ...engine import functional as F
from ...engine import lib as L
from ...engine.config import CONFIG
from .. import networks as N
from .layers import DownBlock

cfg_conv1_nf = {"A": 32}
cfg_mlp_nf = {"A": 256}
cfg_use_norm = {"A": True}
cfg_nconvs = {"A": 5}
cfg_use_resnet = {"A": False}
cfg_activation = {"A": "relu"}


class AffineNetwork(nn.Module):
    def __init__(self, in_channels_a, in_channels_b, height, width, cfg="A", init_func="kaiming"):
        super().__init__()
        self.h, self.w = height, width
        nconvs = cfg_nconvs[cfg]
        self.nconvs = nconvs
        self.convs = N.Holder()
        prev_nf, nf = in_channels_a + in_channels_b, cfg_conv1_nf[cfg]
        self.in_nc = prev_nf
        for i in range(nconvs):
            self.convs.add_module(str(i), DownBlock(prev_nf, nf, 3, 1, 1, bias=True, activation=cfg_activation[cfg],
                                                    init_func=init_func, use_norm=cfg_use_norm[cfg],
                                                    use_resnet=cfg_use_resnet[cfg], skip=False, refine=False, pool=True))
            prev_nf, nf = nf, min(2 * nf, cfg_mlp_nf[cfg])
        self.feat_c = prev_nf
        self.local = N.Holder()
        self.local.add_module("0", nn.Linear(prev_nf * (self.h // 2 ** nconvs) * (self.w // 2 ** nconvs), nf, bias=True))
        self.local.add_module("2", nn.Linear(nf, 6, bias=True))
        last = getattr(self.local, "2")
        last.weight.data.normal_(mean=0.0, std=5e-4)   # start at the identity transformation
        last.bias.data.zero_()

    def forward(self, img_a, img_b):
        x = F.ImagesToNHWC.apply(0, L.PAD_ZERO, CONFIG.dtype, N.image_channels(self.in_nc), img_a, img_b)
        for i in range(self.nconvs):
            x = getattr(self.convs, str(i)).run(x)
        feat = F.ToNCHW.apply(x, self.feat_c)             # (c,h,w) flatten order of the reference
        feat = feat.view(feat.size(0), -1)
        l0, l2 = getattr(self.local, "0"), getattr(self.local, "2")
        hid = F.LinearFn.apply(feat, l0.weight, l0.bias, L.ACT_RELU)
        return F.LinearFn.apply(hid, l2.weight, l2.bias, L.ACT_NONE)


class AffineSTN(nn.Module):
    def __init__(self, nc_a, nc_b, height, width, cfg, init_func):
        super().__init__()
        self.net = AffineNetwork(nc_a, nc_b, height, width, cfg, init_func)
        self.register_buffer("identity_theta", torch.tensor([1, 0, 0, 0, 1, 0], dtype=torch.float), persistent=False)
        self._base = {}

    def _base_coords(self, h, w, device):
        """ATen's affine base grid (align_corners=False): linspace(-1,1,n)*(n-1)/n, built like the reference does."""
        key = (h, w, str(device))
        if key not in self._base:
            bx = (torch.linspace(-1, 1, w) * (w - 1) / w).to(device)
            by = (torch.linspace(-1, 1, h) * (h - 1) / h).to(device)
            self._base[key] = (bx.contiguous(), by.contiguous())
        return self._base[key]

    def _get_theta(self, img_a, img_b):
        dtheta = self.net(img_a, img_b)
        return dtheta, dtheta + self.identity_theta.unsqueeze(0)

    def get_grid(self, img_a, img_b):
        _, theta = self._get_theta(img_a, img_b)
        bx, by = self._base_coords(img_a.size(2), img_a.size(3), img_a.device)
        return F.AffineGridFn.apply(theta, bx, by)

    def precompute(self, img_a, img_b):
        """The affine regressor (independent of `apply_on`): see UnetSTN.precompute."""
        return self._get_theta(img_a, img_b)

    def forward(self, img_a, img_b, apply_on=None, pre=None):
        dtheta, theta = pre if pre is not None else self._get_theta(img_a, img_b)
        if apply_on is None:
            apply_on = [img_a]
        warped = sample_all(lambda h, w, dev: F.AffineGridFn.apply(theta, *self._base_coords(h, w, dev)), apply_on)
        reg_term = F.MeanAbsFn.apply(dtheta, 1.0).squeeze(0)
        return warped, reg_term


def sample_all(grid_for_size, images):
    """Warp every image; images of equal size share one grid and are sampled two per kernel launch."""
    out = [None] * len(images)
    groups = {}
    for i, img in enumerate(images):
        groups.setdefault((img.size(2), img.size(3)), []).append(i)
    for (h, w), idxs in groups.items():
        grid = grid_for_size(h, w, images[idxs[0]].device)
        for j in range(0, len(idxs), 2):
            pair = idxs[j:j + 2]
            if len(pair) == 2:
                o0, o1 = F.GridSampleFn.apply(grid, images[pair[0]], images[pair[1]])
                out[pair[0]], out[pair[1]] = o0, o1
            else:
                out[pair[0]] = F.GridSampleFn.apply(grid, images[pair[0]], None)
    return out
