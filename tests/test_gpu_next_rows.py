"""SURVEY 8(f) rows on the GPU: checkpoint compatibility (f2), inference path (f3), device input pipeline (f4)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import nemar_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_written_checkpoint_forward_matches_reference_outputs():
    """Load the .pth files the reference's own save_networks wrote and reproduce the reference's outputs."""
    from nemar_b200.engine import functional as F
    from nemar_b200.engine.config import configure
    from nemar_b200.models import networks
    configure("fp32", "generic")
    d = os.path.join(ROOT, "tests", "golden", "refckpt")
    meta = np.load(os.path.join(d, "refckpt_meta.npz"))
    netT = networks.define_G(3, 3, 8, "resnet_6blocks", "instance", False, "normal", 0.02, (0,))
    netD = networks.define_D(6, 8, "basic", 3, "instance", "normal", 0.02, (0,))
    netT.load_state_dict(torch.load(os.path.join(d, "latest_net_T.pth"), map_location="cpu"))
    netD.load_state_dict(torch.load(os.path.join(d, "latest_net_D.pth"), map_location="cpu"))
    F.bump_weights_epoch()
    A, B = torch.from_numpy(meta["A"]).cuda(), torch.from_numpy(meta["B"]).cuda()
    with torch.no_grad():
        fake = netT(A)
        pred = netD(torch.cat([A, B], 1))
    np.testing.assert_allclose(fake.cpu().numpy(), meta["fake_B"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(pred.cpu().numpy(), meta["pred"], rtol=0, atol=2e-4)


def test_checkpoint_roundtrip_with_multires_discriminators_and_adam_state(tmp_path):
    """What the reference forgets to save (nemar_model.py:79-81,108-113): the extra discriminator scales and the
    optimizer state.  Train 3 steps, save, load into a fresh model, and the 4th step must be identical."""
    model, cfg, states, (A, B) = H.build_case("c4_multires256")
    H.run_engine_steps(model, A, B, 3)
    model.save_dir = str(tmp_path)
    model.save_networks("latest")
    assert os.path.exists(os.path.join(str(tmp_path), "latest_net_D_ms1.pth")) and os.path.exists(os.path.join(str(tmp_path), "latest_optim.pth"))
    want = H.run_engine_steps(model, A, B, 1)[0]
    fresh, _, _, _ = H.build_case("c4_multires256", seed=99)
    fresh.save_dir = str(tmp_path)
    fresh.load_networks("latest")
    assert fresh.optimizer_D.step_count == 3 and int(fresh.optimizer_D.step_dev) == 3
    got = H.run_engine_steps(fresh, A, B, 1)[0]
    np.testing.assert_allclose(got, want, rtol=2e-3, atol=1e-4)


def test_inference_path_and_get_grid_1024():
    """BaseModel.test() (forward only, no autograd) and UnetSTN.get_grid at 1024x1024 against the oracle (f3)."""
    kw, batch, extra = H.CASE_FLAGS["c5_1024"]
    cfg = O.OracleConfig(**kw)
    T, R, Ds = O.make_states(cfg, seed=11)
    model, _, _, (A, B) = H.build_case("c5_1024")
    model.set_input({"A": A, "B": B, "A_paths": "", "B_paths": ""})
    model.test()
    assert not model.fake_TR_B.requires_grad
    grid = model.netR.get_grid(A.cuda(), B.cuda())
    with torch.no_grad():
        fake = O.resnet_generator(T, A, cfg.n_blocks)
        warped, reg, ogrid = O.unet_stn(R, A, B, [A, fake], cfg.alpha, cfg.multires_reg)
    assert float((grid.cpu() - ogrid).abs().max()) < 5e-5
    assert float((model.fake_B.cpu() - fake).abs().max()) < 5e-4
    assert float((model.registered_real_A.cpu() - warped[0]).abs().max()) < 1e-3


def test_device_prefetcher_feeds_the_same_batches():
    """f4: pinned double-buffered staging must deliver exactly the loader's batches, in order, also when the consumer
    is slow (buffers are recycled only after the consuming step's work has finished) and when the last batch is short."""
    from nemar_b200.data.prefetch import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    host = [{"A": torch.rand((2 if i < 6 else 1, 3, 64, 64), generator=g), "B": torch.rand((2 if i < 6 else 1, 3, 64, 64), generator=g),
             "A_paths": str(i)} for i in range(7)]
    dev = torch.device("cuda", 0)
    sums = []
    for data in DevicePrefetcher(host, dev):
        torch.cuda.current_stream().wait_event(data["_ready"])
        a = data["A"]
        torch.cuda._sleep(20_000_000)                       # a slow "step" that still reads the slot afterwards
        sums.append((a.double().sum() + data["B"].double().sum()).clone())
    torch.cuda.synchronize()
    want = [float(h["A"].double().sum() + h["B"].double().sum()) for h in host]
    np.testing.assert_allclose([float(s) for s in sums], want, rtol=1e-12)


def test_training_through_prefetcher_equals_direct_feed():
    out = {}
    for mode in ("direct", "prefetch"):
        model, cfg, states, (A, B) = H.build_case("c1_affine64")
        batch = {"A": A, "B": B, "A_paths": "", "B_paths": ""}
        losses = []
        if mode == "direct":
            losses = H.run_engine_steps(model, A, B, 3)
        else:
            from nemar_b200.data.prefetch import DevicePrefetcher
            for data in DevicePrefetcher([batch] * 3, model.device):
                model.set_input(data)
                model.optimize_parameters()
                losses.append(list(model.get_current_losses().values()))
        out[mode] = np.array(losses)
    # same input bits => same step up to the atomics-ordered reductions (two direct runs differ by as much)
    np.testing.assert_allclose(out["prefetch"][0], out["direct"][0], rtol=5e-4, atol=1e-5)
    np.testing.assert_allclose(out["prefetch"], out["direct"], rtol=0.3, atol=0.3)      # later steps: chaotic


def test_dropout_masks_follow_the_device_step_counter():
    """Dropout(0.5) of the ResnetBlock (reference networks.py:427-428; the reference's DEFAULT, nemar_model.py:101-102):
    counter-based mask keyed by (seed, call-site salt, device step counter)."""
    from nemar_b200.engine import functional as F
    x = torch.ones((2, 16, 16, 32), dtype=torch.bfloat16, device="cuda", requires_grad=True)
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    y1 = F.DropoutFn.apply(x, 0x5EED, 1 << 40, step)
    vals = set(torch.unique(y1.detach().float()).tolist())
    assert vals == {0.0, 2.0} and 0.45 < float((y1 > 0).float().mean()) < 0.55
    y1.backward(torch.ones_like(y1))
    assert torch.equal(x.grad, y1.detach()), "backward must regenerate the forward's mask"
    assert torch.equal(F.DropoutFn.apply(x, 0x5EED, 1 << 40, step), y1)
    assert not torch.equal(F.DropoutFn.apply(x, 0x5EED, 2 << 40, step), y1), "another call site, another mask"
    step += 1
    assert not torch.equal(F.DropoutFn.apply(x, 0x5EED, 1 << 40, step), y1), "another step, another mask"


def test_cuda_graph_captures_a_step_with_dropout():
    """The reference's default flags keep Dropout(0.5) in every ResnetBlock; --cuda_graph 1 must capture that step and
    every replay must draw new masks (with --lr 0 and a fixed input the outputs differ only through the masks)."""
    kw, batch, extra = H.CASE_FLAGS["c1_affine64"]
    cfg = O.OracleConfig(**kw)
    from nemar_b200.models import create_model
    from nemar_b200.options.train_options import TrainOptions
    argv = ["--dataroot", "none", "--name", "t", "--checkpoints_dir", "/tmp/nemar_b200_ckpt", "--gpu_ids", "0", "--gan_mode", "lsgan",
            "--stn_type", "affine", "--img_height", "64", "--img_width", "64", "--batch_size", str(batch), "--dataset_mode",
            "synthetic", "--precision", "bf16", "--conv_engine", "auto", "--cuda_graph", "1", "--lr", "0"] + list(extra)
    model = create_model(TrainOptions().parse(argv, quiet=True))        # no --no_dropout
    A, B = O.synthetic_batch(batch, cfg.height, cfg.width, seed=1)
    outs = []
    for _ in range(7):
        model.set_input({"A": A, "B": B, "A_paths": "", "B_paths": ""})
        model.optimize_parameters()
        outs.append(model.fake_B.detach().float().clone())
    torch.cuda.synchronize()
    assert model._graph_state["graph"] is not None and not model._graph_state["failed"], "the step was not captured"
    d_replays = float((outs[5] - outs[6]).abs().mean())
    d_eager = float((outs[0] - outs[1]).abs().mean())
    assert d_eager > 1e-3, "eager steps must draw different masks"
    assert d_replays > 0.3 * d_eager, "replays repeat the captured mask: |diff| %.3g vs eager %.3g" % (d_replays, d_eager)
