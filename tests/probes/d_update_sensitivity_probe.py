"""CPU experiment (fp64 oracle only): how sensitive are the T / R gradients of one optimize_parameters to rounding-level
noise in the DISCRIMINATOR's gradient?  The T/R phase differentiates through the discriminator AFTER its Adam update,
and Adam's first update is lr * g / (|g| + eps) ~ lr * sign(g): an element whose gradient is smaller than the noise moves
by +-lr at random.  The fp32 engine is 3e-6 off on netD's gradients but 2e-3 on netT's (8e-4 netR) at the default
learning rate, and 1e-4 / 1e-5 with lr = 0 (tests/probes/fp32_grad_error_probe.py).  Here the fp64 oracle's discriminator
gradient is perturbed before the Adam step with (a) relative and (b) absolute (RMS-scaled, per tensor) Gaussian noise
of that size.

    python tests/probes/d_update_sensitivity_probe.py [--noise 3e-6]
"""
import argparse
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import nemar_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402


def tr_grads(cfg, T, R, Ds, A, B, perturb):
    def go():
        Tc, Rc, Dc = O.cast_states(torch.float64, T, R, Ds)
        st = O.OracleStep(cfg, Tc, Rc, Dc)
        real_step = st.opt_D.step
        st.opt_D.step = lambda grads: real_step([perturb(g) for g in grads])
        st.step(A.double(), B.double())
        return st.grads
    return O.run_in_dtype(torch.float64, go)


def bucket(a, b, names):
    num = den = 0.0
    for k, x, y in zip(names, a, b):
        if k.endswith(".weight"):
            num += float((x - y).norm()) ** 2
            den += float(y.norm()) ** 2
    return (num / den) ** 0.5


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--noise", type=float, default=3e-6)
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    kw, batch, _ = H.CASE_FLAGS["c1_affine64"]
    cfg = O.OracleConfig(**kw)
    T, R, Ds = O.make_states(cfg, seed=11)
    A, B = H.structured_batch(batch, cfg.height, cfg.width)
    st = O.OracleStep(cfg, T, R, Ds)
    for _ in range(a.steps):
        st.step(A, B)
    det = lambda sd: OrderedDict((n, v.detach().clone()) for n, v in sd.items())
    T, R, Ds = det(st.T), det(st.R), [det(d) for d in st.Ds]
    gen = torch.Generator().manual_seed(5)
    base = tr_grads(cfg, T, R, Ds, A, B, lambda g: g)
    d_all = torch.cat([g.flatten() for g in base["D"]]).abs()
    print("netD gradient elements: median |g| %.2e, 1%% quantile %.2e, share below 1e-7: %.2e, below 1e-8 (Adam eps): %.2e" % (
        float(d_all.median()), float(d_all.kthvalue(max(1, d_all.numel() // 100)).values), float((d_all < 1e-7).double().mean()),
        float((d_all < 1e-8).double().mean())))
    rel = lambda g: g * (1 + a.noise * torch.randn(g.shape, generator=gen, dtype=g.dtype))
    ab = lambda g: g + a.noise * float(g.pow(2).mean().sqrt()) * torch.randn(g.shape, generator=gen, dtype=g.dtype)
    for tag, fn in (("relative noise %.0e" % a.noise, rel), ("absolute noise %.0e x RMS per tensor" % a.noise, ab)):
        g = tr_grads(cfg, T, R, Ds, A, B, fn)
        print("%-42s -> netT %.2e  netR %.2e" % (tag, bucket(g["T"], base["T"], list(T.keys())), bucket(g["R"], base["R"], list(R.keys()))))


if __name__ == "__main__":
    main()
