"""Module- and step-level parity of the engine against the oracle and the reference's golden vectors."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import nemar_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ["L1_TR", "GAN_TR", "L1_RT", "GAN_RT", "smoothness", "D_fake_TR", "D_fake_RT", "D"]


def _maxerr(a, b):
    return float((a.detach().float().cpu() - b.detach().float().cpu()).abs().max())


def test_networks_forward_fp32_vs_oracle():
    """netT / netR / netD forward, fp32 storage, generic engine: tight agreement with the oracle."""
    model, cfg, (T, R, Ds), (A, B) = H.build_case("c1_affine64")
    Ad, Bd = A.cuda(), B.cuda()
    with torch.no_grad():
        ref = O.resnet_generator(T, A, cfg.n_blocks)
        assert _maxerr(model.netT(Ad), ref) < 2e-4
        ref_d = O.nlayer_discriminator(Ds[0], torch.cat([A, B], 1))
        assert _maxerr(model.netD(torch.cat([Ad, Bd], 1)), ref_d) < 2e-4
        warped, reg, theta = O.affine_stn(R, A, B, [A, ref])
        ew, ereg = model.netR(Ad, Bd, apply_on=[Ad, ref.cuda()])
        assert _maxerr(ew[0], warped[0]) < 2e-4 and _maxerr(ew[1], warped[1]) < 2e-4
        assert abs(float(ereg) - float(reg)) < 1e-5


def test_unet_stn_forward_fp32_vs_oracle():
    model, cfg, (T, R, Ds), (A, B) = H.build_case("c4_multires256")
    with torch.no_grad():
        fake = O.resnet_generator(T, A, cfg.n_blocks)
        warped, reg, grid = O.unet_stn(R, A, B, [A, fake], cfg.alpha, cfg.multires_reg)
        ew, ereg = model.netR(A.cuda(), B.cuda(), apply_on=[A.cuda(), fake.cuda()])
        egrid = model.netR.get_grid(A.cuda(), B.cuda())
    assert _maxerr(egrid, grid) < 5e-5, "sampling grid"
    assert _maxerr(ew[0], warped[0]) < 1e-3 and _maxerr(ew[1], warped[1]) < 1e-3
    assert abs(float(ereg) - float(reg)) < 2e-4 * max(1.0, abs(float(reg)))


# Trajectories: the first step is pinned tightly.  Later steps are NOT reproducible to better than ~10-20 % even
# between the oracle evaluated in fp32 and in fp64 (Adam's first updates are ~lr*sign(g), and the LSGAN gradient
# through InstanceNorm is a cancellation-dominated quantity: fp32-vs-fp64 oracle D_fake_TR differs by 7 % at step 2
# and 21 % at step 3 on c1 — measured, see DESIGN.md "Parity").  They are therefore checked loosely.
L1_COLS, ADV_COLS = [0, 2, 4], [1, 3, 5, 6, 7]     # reconstruction / smoothness terms vs adversarial terms


def _check_traj(losses, gold, first_rtol, first_atol):
    np.testing.assert_allclose(losses[0], gold[0], rtol=first_rtol, atol=first_atol, err_msg="step-1 losses %s" % NAMES)
    if len(losses) > 1:
        later, g = losses[1:], gold[1:len(losses)]
        np.testing.assert_allclose(later[:, L1_COLS], g[:, L1_COLS], rtol=0.05, atol=0.5, err_msg="later-step L1/smoothness")
        np.testing.assert_allclose(later[:, ADV_COLS], g[:, ADV_COLS], rtol=0.6, atol=0.3, err_msg="later-step adversarial terms")


@pytest.mark.parametrize("name,steps", [("c1_affine64", 3), ("c4_multires256", 2)])
def test_training_step_fp32_vs_reference_golden(name, steps):
    """Full optimize_parameters (fp32, generic engine) against the REFERENCE's own logged losses."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    model, cfg, states, (A, B) = H.build_case(name)
    losses = np.array(H.run_engine_steps(model, A, B, steps))
    _check_traj(losses, g["losses"], 3e-4, 2e-5)


def test_first_step_images_fp32_vs_reference_golden():
    g = np.load(os.path.join(GOLD, "c1_affine64.npz"))
    model, cfg, states, (A, B) = H.build_case("c1_affine64")
    H.run_engine_steps(model, A, B, 1)
    for k in ("fake_B", "registered_real_A", "fake_TR_B", "fake_RT_B"):
        np.testing.assert_allclose(getattr(model, k).detach().cpu().numpy(), g["img_" + k], rtol=0, atol=5e-4, err_msg=k)


@pytest.mark.parametrize("name,steps", [("c1_affine64", 3), ("c4_multires256", 2), ("c2_unet256", 2)])
@pytest.mark.parametrize("conv_engine", ["generic", "auto"])
def test_training_step_bf16_vs_reference_golden(name, steps, conv_engine):
    """bf16 storage (fp32 accumulate): stated tolerance 4 % relative on every logged loss (the
    reference itself moves ~1.2 % under bf16 autocast), absolute 0.02 for near-zero terms."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    model, cfg, states, (A, B) = H.build_case(name, precision="bf16", conv_engine=conv_engine)
    losses = np.array(H.run_engine_steps(model, A, B, steps))
    _check_traj(losses, g["losses"], 4e-2, 2e-2)


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def _bucket_names(model):
    d = [k for net in [model.netD, *model.netD_multiresolution] for k, _ in net.named_parameters()]
    tr = [k for net in (model.netR, model.netT) for k, _ in net.named_parameters()]
    return d, tr


def _weights_only(opt, names, flat):
    """gradient bucket with the bias slices zeroed: a bias feeding an InstanceNorm has an analytically zero gradient —
    what any arithmetic produces there is rounding noise (same exclusion as test_gradients_vs_fp64_oracle)"""
    assert len(names) == len(opt.params)
    out = flat.detach().float().cpu().clone()
    for p, o, nm in zip(opt.params, opt.offsets, names):
        if not nm.endswith(".weight"):
            out[o:o + p.numel()] = 0
    return out


def _step_record(model, A, B):
    losses = np.array(H.run_engine_steps(model, A, B, 1)[0])
    d_names, tr_names = _bucket_names(model)
    return (losses, _weights_only(model.optimizer_D, d_names, model.optimizer_D.flat_g),
            _weights_only(model.optimizer_TR, tr_names, model.optimizer_TR.flat_g))


def _case_inputs(name, k):
    g = torch.Generator().manual_seed(100 + k)
    kw, batch, _ = H.CASE_FLAGS[name]
    A = torch.rand((batch, 3, kw["height"], kw["width"]), generator=g) * 2 - 1
    B = torch.rand((batch, 3, kw["height"], kw["width"]), generator=g) * 2 - 1
    return A, B


def _worst_params(opt, names, a, b, top=4):
    """per-parameter relative difference of two flat gradient buckets -> the `top` worst as text"""
    rows = []
    for p, o, nm in zip(opt.params, opt.offsets, names):
        n = p.numel()
        ga, gb = a[o:o + n], b[o:o + n]
        rows.append((float((ga - gb).norm()), float(gb.norm()), nm, tuple(p.shape)))
    rows.sort(reverse=True)
    return "; ".join("%s%s |diff| %.3g |ref| %.3g" % (nm, sh, d, r) for d, r, nm, sh in rows[:top])


def _assert_same_step(got, ref, noise, floors, what, model=None):
    """got / ref / noise: (losses, D bucket, T+R bucket) records.  `noise` is a second evaluation of `ref`'s step by
    the reference path itself: the reductions use floating-point atomics, and the T+R gradient amplifies rounding-level
    differences ~1e5x, so the bound is 3x the path's own run-to-run spread, floored."""
    f_loss, f_d, f_tr = floors
    e_loss = float(np.max(np.abs(got[0] - ref[0]) / (np.abs(ref[0]) + 1e-3)))
    n_loss = float(np.max(np.abs(noise[0] - ref[0]) / (np.abs(ref[0]) + 1e-3)))
    e_d, n_d = _rel(got[1], ref[1]), _rel(noise[1], ref[1])
    e_tr, n_tr = _rel(got[2], ref[2]), _rel(noise[2], ref[2])
    msg = "%s: losses %.3g (run-to-run %.3g), D bucket %.3g (%.3g), T+R bucket %.3g (%.3g)\n  got %s\n  ref %s" % (
        what, e_loss, n_loss, e_d, n_d, e_tr, n_tr, got[0], ref[0])
    if model is not None:
        d_names, _ = _bucket_names(model)
        msg += "\n  worst D parameters: " + _worst_params(model.optimizer_D, d_names, got[1], ref[1])
    print(msg)
    assert e_loss <= max(3 * n_loss, f_loss) and e_d <= max(3 * n_d, f_d) and e_tr <= max(3 * n_tr, f_tr), msg


# (losses, D weights, T+R weights): floors of the comparisons below = the spread measured between two eager evaluations
# of the same step on B200 (fp32: 5e-6 / 3e-3 / 1.8e-2; bf16: 3e-3 / 4e-2 / 0.75 — bf16 keeps almost nothing of the T/R
# gradient)
FLOORS = {"fp32": (5e-5, 5e-3, 5e-2), "bf16": (1e-2, 0.1, 1.0)}


@pytest.mark.parametrize("name,precision,engine", [("c1_affine64", "fp32", "generic"), ("c4_multires256", "fp32", "generic"),
                                                    ("c4_multires256", "bf16", "auto")])
def test_batched_discriminator_equals_separate_passes(name, precision, engine):
    """--batch_d 1 (one discriminator pass per phase over the batch-concatenated (A, B_k) pairs) against one pass per
    pair, as the reference does (nemar_model.py:181,197,219,233,247): same losses and D / T+R gradients.  One model,
    --lr 0 (every step is then the same function of its input), the flag toggled between steps."""
    model, cfg, states, (A, B) = H.build_case(name, precision=precision, conv_engine=engine, more_flags=["--batch_d", "0", "--lr", "0"])
    floors = FLOORS[precision]
    if name == "c1_affine64":          # well-conditioned state (see test_cuda_graph_replay_equals_eager): tight floors
        from tests.test_gpu_fidelity import _trained_state
        _, T, R, Ds, A, B = _trained_state(name)
        H.load_states(model, T, R, Ds)
        floors = (5e-5, 5e-4, 2e-3)
    ref = _step_record(model, A, B)
    model.opt.batch_d = 1
    got = _step_record(model, A, B)
    model.opt.batch_d = 0
    ref_late = _step_record(model, A, B)
    best = ref if _rel(got[1], ref[1]) <= _rel(got[1], ref_late[1]) else ref_late
    _assert_same_step(got, best, ref_late if best is ref else ref, floors, "batch_d 1 vs 0", model)


@pytest.mark.parametrize("name,precision,engine,state", [("c1_affine64", "fp32", "generic", "trained"),
                                                         ("c4_multires256", "bf16", "auto", "init")])
def test_cuda_graph_replay_equals_eager(name, precision, engine, state):
    """--cuda_graph 1: steps 1-3 run eagerly, step 4 is captured and replayed, later steps replay the graph.  With
    --lr 0 every step is the same function of its input, so a replayed step must reproduce the eager step on the same
    input — including inputs the capture never saw.  Yardstick: the spread between two EAGER evaluations of the same
    input (one before the capture, one after the replays).
    The fp32 case runs at the TRAINED state of tests/test_gpu_fidelity.py (oracle weights after 30 steps, structured
    inputs): there the step is well conditioned (fp32 gradients 1e-6 from fp64 in the oracle) and the comparison is held
    to 5e-4 (D) / 2e-3 (T+R) — at the seeded-init / white-noise state of round 1 a single ReLU-derivative flip moved the D
    gradient by 3e-3 and the test had to be calibrated on that noise.  The bf16 case keeps the init state and its floors."""
    model, cfg, states, _ = H.build_case(name, precision=precision, conv_engine=engine, more_flags=["--cuda_graph", "1", "--lr", "0"])
    if state == "trained":
        from tests.test_gpu_fidelity import _trained_state
        _, T, R, Ds, _, _ = _trained_state(name)
        H.load_states(model, T, R, Ds)
        kw, batch, _ = H.CASE_FLAGS[name]
        X = [H.structured_batch(batch, kw["height"], kw["width"], seed=5 + k) for k in range(3)]
        floors = (5e-5, 5e-4, 2e-3)
    else:
        X = [_case_inputs(name, k) for k in range(3)]
        floors = FLOORS[precision]
    early = [_step_record(model, *x) for x in X]               # eager (warm-up of the graph mode)
    replay = [_step_record(model, *x) for x in X]              # capture + replay, replay, replay
    assert model._graph_state["graph"] is not None and not model._graph_state["failed"], "the step was not captured"
    model.opt.cuda_graph = 0
    late = [_step_record(model, *x) for x in X]                # eager again
    assert _rel(early[1][1], early[0][1]) > 1e-2, "different inputs must give different gradients (test self-check)"
    spread = [max(_rel(e[i], l[i]) for e, l in zip(early, late)) for i in (1, 2)]
    l_spread = max(float(np.max(np.abs(e[0] - l[0]) / (np.abs(l[0]) + 1e-3))) for e, l in zip(early, late))
    f_loss, f_d, f_tr = floors
    rows, ok = [], True
    for k, what in enumerate(("the captured input", "input 1", "an input the capture never saw")):
        g, e, l = replay[k], early[k], late[k]
        e_loss = min(float(np.max(np.abs(g[0] - r[0]) / (np.abs(r[0]) + 1e-3))) for r in (e, l))
        e_d, e_tr = min(_rel(g[1], e[1]), _rel(g[1], l[1])), min(_rel(g[2], e[2]), _rel(g[2], l[2]))
        rows.append("replay on %s: losses %.3g, D weights %.3g, T+R weights %.3g" % (what, e_loss, e_d, e_tr))
        ok = ok and e_loss <= max(3 * l_spread, f_loss) and e_d <= max(3 * spread[0], f_d) and e_tr <= max(3 * spread[1], f_tr)
    msg = "eager-vs-eager spread: losses %.3g, D weights %.3g, T+R weights %.3g\n%s" % (l_spread, spread[0], spread[1], "\n".join(rows))
    print(msg)
    assert ok, msg


def test_cuda_graph_replay_trains():
    """Default lr: the captured Adam launches (device-resident step counter) must keep updating the weights on replay."""
    out = {}
    for flag in ("0", "1"):
        model, cfg, states, (A, B) = H.build_case("c1_affine64", more_flags=["--cuda_graph", flag])
        p0 = model.optimizer_TR.flat_p.detach().clone()
        losses = np.array(H.run_engine_steps(model, A, B, 6))
        out[flag] = (losses, float((model.optimizer_TR.flat_p - p0).abs().mean()), float(model.optimizer_D.exp_avg_sq.sum()))
    l0, moved0, v0 = out["0"]
    l1, moved1, v1 = out["1"]
    assert abs(moved1 - moved0) <= 0.05 * moved0, "mean |delta w| after 6 steps: graph %g vs eager %g" % (moved1, moved0)
    assert abs(v1 - v0) <= 0.2 * v0, "Adam second moment: graph %g vs eager %g" % (v1, v0)
    np.testing.assert_allclose(l1[:, L1_COLS], l0[:, L1_COLS], rtol=0.05, atol=0.5)


def test_checkpoint_roundtrip_reference_keys(tmp_path):
    model, cfg, (T, R, Ds), (A, B) = H.build_case("c1_affine64")
    model.save_dir = str(tmp_path)
    model.save_networks("latest")
    sd = torch.load(os.path.join(str(tmp_path), "latest_net_T.pth"))
    assert list(sd.keys()) == list(T.keys())
    assert all(torch.equal(sd[k], T[k]) for k in T)
    model.load_networks("latest")


def _oracle_grads(cfg, T, R, Ds, A, B, dtype, autocast=False):
    def go():
        Tc, Rc, Dc = O.cast_states(dtype, T, R, Ds)
        st = O.OracleStep(cfg, Tc, Rc, Dc)
        if autocast:
            with torch.autocast("cpu", dtype=torch.bfloat16):
                st.step(A.to(dtype), B.to(dtype))
        else:
            st.step(A.to(dtype), B.to(dtype))
        return {k: [g.double() for g in v] for k, v in st.grads.items()}
    return O.run_in_dtype(dtype, go)


@pytest.mark.parametrize("name,precision,engine", [("c1_affine64", "fp32", "generic"), ("c4_multires256", "fp32", "generic"),
                                                   ("c1_affine64", "bf16", "generic"), ("c1_affine64", "bf16", "auto"),
                                                   ("c4_multires256", "bf16", "auto")])
def test_gradients_vs_fp64_oracle(name, precision, engine):
    """Per-tensor weight gradients of both optimizer phases after one optimize_parameters.  Truth = the oracle in
    fp64.  The yardstick is the error the REFERENCE's own arithmetic makes on the same problem:
      fp32 engine: median error per network <= 8x the fp32 oracle's, every tensor <= 20x (both are rounding noise amplified ~1e5x,
      and the engine's atomics-ordered reductions make its noise vary run to run);
      bf16 engine: error <= 1.25x the error of the oracle under torch.autocast(bfloat16) (floor 0.1) — the LSGAN
      gradient through InstanceNorm is common-mode dominated, so ANY bf16 arithmetic loses most of it (the
      reference under autocast is 85-160 % off on netT/netR here; measured, see DESIGN.md "Parity")."""
    model, cfg, (T, R, Ds), (A, B) = H.build_case(name, precision=precision, conv_engine=engine)
    H.run_engine_steps(model, A, B, 1)
    truth = _oracle_grads(cfg, T, R, Ds, A, B, torch.float64)
    if precision == "fp32":
        yard, factor, floor = _oracle_grads(cfg, T, R, Ds, A, B, torch.float32), 20.0, 5e-3
    else:
        yard, factor, floor = _oracle_grads(cfg, T, R, Ds, A, B, torch.float32, autocast=True), 1.25, 0.1
    bad, summary = [], {}
    for tag, net in (("R", model.netR), ("T", model.netT), ("D", model.netD)):
        errs = []
        for i, (k, p) in enumerate(net.named_parameters()):
            if not k.endswith(".weight"):
                continue     # biases feeding an InstanceNorm have zero true gradient (rounding noise in any arithmetic)
            t = truth[tag][i]
            nrm = float(t.norm()) + 1e-30
            e_eng = float((p.grad.detach().double().cpu() - t).norm()) / nrm
            e_ref = float((yard[tag][i] - t).norm()) / nrm
            errs.append((e_eng, e_ref))
            # the reference arithmetic itself has lost this tensor entirely (error > 100 %): both numbers are noise and only
            # their order of magnitude can be held (seen: engine 2.52 against 1.26 on one STN tensor, from run to run)
            pure_noise = e_ref > 1.0 and e_eng < 3.0 * e_ref
            if e_eng > max(factor * e_ref, floor) and not pure_noise:
                bad.append((e_eng, e_ref, tag, k))
        summary[tag] = (float(np.median([a for a, _ in errs])), float(np.median([b for _, b in errs])))
        med_factor = 8.0 if precision == "fp32" else factor
        assert summary[tag][0] <= med_factor * summary[tag][1] + floor, "median gradient error of net%s: %s" % (tag, summary[tag])
    print("median gradient error vs fp64 truth (engine, reference-arithmetic yardstick):", summary)
    msg = "\n".join("%s.%s engine err %.3e, yardstick err %.3e" % (t, k, a, b) for a, b, t, k in sorted(bad, reverse=True)[:30])
    assert not bad, "gradients less accurate than allowed (worst first):\n" + msg
