"""Pin the plain-C grid oracle against ATen (the library the reference calls) on this machine."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import grid_oracle as G
from oracle import nemar_oracle as O


def _grids(n, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    ident = O.identity_grid(h, w).permute(0, 2, 3, 1).repeat(n, 1, 1, 1)
    near = ident + torch.randn((n, h, w, 2), generator=g) * (4.0 / w)          # ~2 px deformation
    wild = torch.rand((n, h, w, 2), generator=g) * 2.2 - 1.1                   # ~9 % out of bounds
    return dict(identity=ident, near=near, wild=wild)


@pytest.mark.parametrize("h,w", [(64, 64), (37, 53), (256, 256), (288, 384)])
def test_grid_sample_fwd_bwd_and_indices(h, w):
    n, c = 2, 3
    img = torch.rand((n, c, h, w), generator=torch.Generator().manual_seed(3)) * 2 - 1
    for name, grid in _grids(n, h, w, 5).items():
        out, idx = G.grid_sample_fwd(img.numpy(), grid.numpy())
        ref = F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
        # 2e-6: the fused-multiply-add unnormalisation ATen uses; a two-rounding restatement is 1.4e-5 off at 288 x 384
        np.testing.assert_allclose(out, ref.numpy(), rtol=0, atol=2e-6, err_msg=name)
        # integer tap indices: bit-exact against fma(g + 1, size / 2, -0.5) (exact product in float64, one rounding)
        g1 = (grid + 1).numpy().astype(np.float64)
        ix = (g1[..., 0] * (w / 2) - 0.5).astype(np.float32)
        iy = (g1[..., 1] * (h / 2) - 0.5).astype(np.float32)
        assert np.array_equal(idx[..., 0], np.floor(ix).astype(np.int32)), name
        assert np.array_equal(idx[..., 1], np.floor(iy).astype(np.int32)), name
        # backward
        imgr, gridr = img.clone().requires_grad_(True), grid.clone().requires_grad_(True)
        dout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(7))
        F.grid_sample(imgr, gridr, mode="bilinear", padding_mode="zeros", align_corners=False).backward(dout)
        dimg, dgrid = G.grid_sample_bwd(img.numpy(), grid.numpy(), dout.numpy())
        np.testing.assert_allclose(dimg, imgr.grad.numpy(), rtol=0, atol=2e-5, err_msg=name)
        np.testing.assert_allclose(dgrid, gridr.grad.numpy(), rtol=2e-4, atol=2e-3, err_msg=name)


def test_affine_and_flow_grid():
    theta = torch.tensor([[1.05, 0.02, -0.03, -0.04, 0.97, 0.05], [1, 0, 0, 0, 1, 0]], dtype=torch.float32)
    ref = F.affine_grid(theta.view(-1, 2, 3), (2, 3, 48, 80), align_corners=False)
    np.testing.assert_allclose(G.affine_grid(theta.numpy(), 48, 80), ref.numpy(), rtol=0, atol=3e-7)
    off = torch.randn(2, 2, 40, 56) * 0.01
    ref = (O.identity_grid(40, 56).repeat(2, 1, 1, 1) + off).permute(0, 2, 3, 1)
    assert np.array_equal(G.flow_grid(off.numpy()), ref.contiguous().numpy())


@pytest.mark.parametrize("alpha", [0.0, 1.0])
def test_smoothness(alpha):
    d = torch.randn(2, 2, 33, 47) * 0.02
    img = torch.rand(2, 3, 33, 47) * 2 - 1
    ref = float(O.smoothness_loss(d, img, alpha))
    assert abs(G.smoothness(d.numpy(), img.numpy(), alpha) - ref) < 1e-6 * max(1.0, abs(ref))
