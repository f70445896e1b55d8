"""bf16 gradient fidelity and convergence as MEASURED properties.

Round 1 compared bf16 gradients only at the seeded-init / white-noise state, where the T/R gradient is what is left
after the LSGAN common mode cancels in every InstanceNorm: there even the fp32 oracle is 6e-3 off the fp64 oracle and
the reference under torch.autocast(bfloat16) is ~100 % off (tests/probes/grad_fidelity_probe.py, profiles/r02_grad_fidelity_cpu.txt).
That state lasts a handful of steps.  Here the step is evaluated where training actually happens:
  * `trained state` = the fp32 ORACLE's weights after 30 steps on a structured batch (smooth images, B = remapped,
    4-px-shifted A): the oracle's fp32 gradients are 1e-6 from fp64 there, the autocast reference 7-9 % (bucket norm);
  * the bf16 engine's gradients are held to an ABSOLUTE bound against the fp64 oracle (not a relative yardstick);
  * a 150-step run checks that the bf16 engine trains like the fp32 engine (same losses, same registration error).
"""
import numpy as np
import pytest
import torch
from collections import OrderedDict

pytestmark = pytest.mark.gpu

from oracle import nemar_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_gpu_model import _oracle_grads  # noqa: E402

TRAIN_STEPS = 30
_cache = {}


def _trained_state(name="c1_affine64"):
    """(cfg, T, R, Ds, A, B): fp32 oracle weights after TRAIN_STEPS steps on the structured batch (CPU, ~40 s)."""
    if name not in _cache:
        kw, batch, _ = H.CASE_FLAGS[name]
        cfg = O.OracleConfig(**kw)
        T, R, Ds = O.make_states(cfg, seed=11)
        A, B = H.structured_batch(batch, cfg.height, cfg.width)
        st = O.OracleStep(cfg, T, R, Ds)
        for _ in range(TRAIN_STEPS):
            st.step(A, B)
        det = lambda sd: OrderedDict((k, v.detach().clone()) for k, v in sd.items())
        _cache[name] = (cfg, det(st.T), det(st.R), [det(d) for d in st.Ds], A, B)
    return _cache[name]


def _bucket_err(net, truth, grads=None):
    """norm-wise error of all weight gradients of a network taken as one vector, and the per-tensor median"""
    num = den = 0.0
    per = []
    for i, (k, p) in enumerate(net.named_parameters()):
        if not k.endswith(".weight"):
            continue
        t = truth[i]
        g = p.grad.detach().double().cpu() if grads is None else grads[i]
        num += float((g - t).norm()) ** 2
        den += float(t.norm()) ** 2
        per.append(float((g - t).norm()) / (float(t.norm()) + 1e-30))
    return (num / den) ** 0.5, float(np.median(per))


# bucket-norm error bounds vs the fp64 oracle at the trained state (T, R, D).  Measured on B200 (profiles/r02_fidelity_gpu.txt).
# Measured (profiles/r02_fidelity_gpu.txt, r02_fp32_gradient_error_probes.txt): fp32 engine D 3e-6; T / R between 4e-6 / 3e-6
# and 2.4e-3 / 8e-4 from run to run (derivative bits of the few activations within rounding of zero, ~1e-4 each:
# DESIGN.md section 3, tests/probes/relu_flip_probe.py; not the one-pass variance, not the Adam step, not stale packs); bf16 engine
# T 6e-2 / R 2.6e-2 / D 5e-2 — BELOW the reference under torch.autocast(bfloat16) on the same state (9.8e-2 / 9.6e-2 / 8.4e-2).
BOUNDS = {"fp32": (6e-3, 3e-3, 1e-4), "bf16": (0.10, 0.06, 0.08)}


@pytest.mark.parametrize("precision,engine", [("fp32", "generic"), ("bf16", "generic"), ("bf16", "auto")])
def test_gradients_at_trained_state_vs_fp64_oracle(precision, engine):
    cfg, T, R, Ds, A, B = _trained_state()
    # default lr on both sides: the T/R phase differentiates through the discriminator AFTER its Adam update
    model, _, _, _ = H.build_case("c1_affine64", precision=precision, conv_engine=engine)
    H.load_states(model, T, R, Ds)
    H.run_engine_steps(model, A, B, 1)
    truth = _oracle_grads(cfg, T, R, Ds, A, B, torch.float64)
    yard = _oracle_grads(cfg, T, R, Ds, A, B, torch.float32, autocast=True)
    rows, ok = [], True
    for (tag, net), bound in zip((("T", model.netT), ("R", model.netR), ("D", model.netD)), BOUNDS[precision]):
        e_b, e_m = _bucket_err(net, truth[tag])
        y_b, y_m = _bucket_err(net, truth[tag], yard[tag])
        rows.append("net%s: engine %s bucket %.3e median %.3e | reference under bf16 autocast: bucket %.3e median %.3e" % (
            tag, precision, e_b, e_m, y_b, y_m))
        ok = ok and e_b <= bound
    msg = "\n".join(rows)
    print(msg)
    assert ok, "gradient error vs fp64 oracle above %s:\n%s" % (BOUNDS[precision], msg)


def _theta_err(model, A, B, shift_px, w):
    """mean |translation - truth| of the affine STN on the batch (x translation truth = the known shift, normalised)"""
    with torch.no_grad():
        dtheta, theta = model.netR._get_theta(A.cuda(), B.cuda())
    truth = torch.tensor([1, 0, -2.0 * shift_px / w, 0, 1, 0], device=theta.device)
    return float((theta - truth).abs().mean())


def test_bf16_engine_trains_like_fp32_engine():
    """150 steps on a fixed structured batch with a known 4-px shift: the bf16 tcgen05 engine and the fp32 generic
    engine must both train (L1 down by > 30 %) and reach the same reconstruction losses / registration error."""
    steps, out = 150, {}
    kw, batch, _ = H.CASE_FLAGS["c1_affine64"]
    A, B = H.structured_batch(batch, kw["height"], kw["width"])
    for precision, engine in (("fp32", "generic"), ("bf16", "auto")):
        model, cfg, states, _ = H.build_case("c1_affine64", precision=precision, conv_engine=engine)
        e0 = _theta_err(model, A, B, 4, kw["width"])
        losses = np.array(H.run_engine_steps(model, A, B, steps))
        tail = losses[-30:].mean(0)
        out[precision] = dict(L1_TR=tail[0], L1_RT=tail[2], first=losses[0], err0=e0, err=_theta_err(model, A, B, 4, kw["width"]))
    print("convergence after %d steps: %s" % (steps, out))
    f, b = out["fp32"], out["bf16"]
    assert f["L1_RT"] < 0.7 * f["first"][2] and b["L1_RT"] < 0.7 * b["first"][2], "both engines must have trained"
    # yardstick: three runs of the fp32 engine ALONE spread 5.6 ... 7.3 on L1_RT here (the adversarial game is chaotic and
    # the reductions are atomics-ordered), so the bound is 30 % on the 30-step tail means, 20 % on the registration error
    for k, tol in (("L1_TR", 0.30), ("L1_RT", 0.30), ("err", 0.20)):
        assert abs(b[k] - f[k]) <= tol * abs(f[k]), "%s: bf16 %.4g vs fp32 %.4g" % (k, b[k], f[k])
