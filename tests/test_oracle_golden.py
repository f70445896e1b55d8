"""Pin the CPU oracle against the golden vectors produced by the reference itself
(tests/golden/make_golden.py ran the reference's NEMARModel.optimize_parameters on the same seeded state)."""
import os

import numpy as np
import pytest
import torch

from oracle import nemar_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")

CASES = {
    "c1_affine64": dict(cfg=dict(stn_type="affine", n_blocks=6, height=64, width=64), batch=2),
    "c2_unet256": dict(cfg=dict(stn_type="unet", n_blocks=9, height=256, width=256, lambda_smooth=200.0), batch=1),
    "c4_multires256": dict(cfg=dict(stn_type="unet", n_blocks=3, height=256, width=256, lambda_smooth=200.0, alpha=1.0,
                                    multires_reg=2, multi_resolution=2, ngf=16, ndf=16), batch=1),
    "c4_ms3_512": dict(cfg=dict(stn_type="unet", n_blocks=9, height=512, width=512, lambda_smooth=200.0, alpha=1.0,
                                multires_reg=3, multi_resolution=3, ngf=16, ndf=16), batch=1),
    "c5_1024": dict(cfg=dict(stn_type="unet", n_blocks=9, height=1024, width=1024, lambda_smooth=200.0, ngf=8, ndf=8), batch=1),
    "ragged288x384": dict(cfg=dict(stn_type="unet", n_blocks=6, height=288, width=384, lambda_smooth=200.0, ngf=16, ndf=16),
                          batch=2),
}


def run_oracle(name, steps):
    case = CASES[name]
    cfg = O.OracleConfig(**case["cfg"])
    T, R, Ds = O.make_states(cfg, seed=11)
    A, B = O.synthetic_batch(case["batch"], cfg.height, cfg.width, seed=1)
    st = O.OracleStep(cfg, T, R, Ds)
    losses, first = [], None
    for s in range(steps):
        losses.append(list(st.step(A, B).values()))
        if s == 0:
            first = {k: v.detach() for k, v in st.out.items() if k.startswith(("fake", "registered"))}
    return st, np.array(losses), first


@pytest.mark.parametrize("name", ["c1_affine64", "c4_multires256", "c2_unet256", "c4_ms3_512", "ragged288x384", "c5_1024"])
def test_oracle_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    steps = 1 if name == "c2_unet256" else g["losses"].shape[0]   # the 256^2 resnet_9 case is slow on CPU
    st, losses, first = run_oracle(name, steps)
    assert list(g["loss_names"]) == ["L1_TR", "GAN_TR", "L1_RT", "GAN_RT", "smoothness", "D_fake_TR", "D_fake_RT", "D"]
    np.testing.assert_allclose(losses, g["losses"][:steps], rtol=2e-4, atol=1e-5)
    stride = int(g["img_stride"])
    for k in ("fake_B", "registered_real_A", "fake_TR_B", "fake_RT_B"):
        np.testing.assert_allclose(first[k][:, :, ::stride, ::stride].numpy(), g["img_" + k], rtol=0, atol=2e-4)
    if steps == g["losses"].shape[0]:
        for tag, sd in (("T", st.T), ("R", st.R), ("D", st.Ds[0])):
            assert list(g["pkeys_" + tag]) == list(sd.keys())
            # biases that feed an InstanceNorm have an analytically zero gradient: what Adam sees there is
            # rounding noise (in the reference too), so only weights are pinned tightly.
            keep = np.array([k.endswith(".weight") for k in sd.keys()])
            pabs = np.array([float(v.detach().double().abs().sum()) for v in sd.values()])
            np.testing.assert_allclose(pabs[keep], g["pabs_" + tag][keep], rtol=2e-4)


def test_shapes_enumerate_reference_counts():
    """SURVEY.md section 8a5 parameter counts of the reference."""
    n = lambda s: sum(int(np.prod(v)) for v in s.values())
    assert n(O.resnet_generator_shapes(n_blocks=9)) == 11378179 and len(O.resnet_generator_shapes(n_blocks=9)) == 48
    assert n(O.resnet_generator_shapes(n_blocks=6)) == 7837699
    assert n(O.discriminator_shapes()) == 2767809 and len(O.discriminator_shapes()) == 10
    assert n(O.unet_stn_shapes()) == 2059170 and len(O.unet_stn_shapes()) == 80
    assert n(O.affine_stn_shapes(height=64, width=64)) == 1243302 and len(O.affine_stn_shapes()) == 14


def test_identity_grid_quirk():
    """Reference identity grid is the align_corners=True identity sampled with align_corners=False:
    ix_j = j*W/(W-1) - 0.5."""
    w = 256
    g = O.identity_grid(4, w)[0, 0, 0]
    ix = ((g + 1) * w - 1) / 2
    assert abs(float(ix[0]) + 0.5) < 1e-6 and abs(float(ix[-1]) - (w - 0.5)) < 1e-4
