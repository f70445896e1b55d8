"""CPU experiment (oracle only): does rounding-ORDER noise in the InstanceNorm plane sums explain the fp32 engine's
run-to-run spread of the T / R gradient error (4e-6 ... 2.4e-3 against the fp64 oracle, profiles/r02_fp32_gradient_error_probes.txt)?
The engine's plane sums are atomics-ordered, so the last bits of every plane mean differ from run to run.  Here the fp32
oracle's InstanceNorm gets its plane mean and variance perturbed by a relative Gaussian noise of one fp32 ulp-ish size
(default 1e-7), several seeds, and the T / R / D weight gradients are compared with the UNPERTURBED fp32 oracle.

    python tests/probes/mean_noise_probe.py [--noise 1e-7] [--seeds 6]
"""
import argparse
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import nemar_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.probes.grad_fidelity_probe import grads  # noqa: E402
from tests.probes.onepass_var_probe import bucket_err  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--noise", type=float, default=1e-7)
    ap.add_argument("--seeds", type=int, default=6)
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    kw, batch, _ = H.CASE_FLAGS["c1_affine64"]
    cfg = O.OracleConfig(**kw)
    T, R, Ds = O.make_states(cfg, seed=11)
    As, Bs = H.structured_batch(batch, cfg.height, cfg.width)
    st = O.OracleStep(cfg, T, R, Ds)
    for _ in range(a.steps):
        st.step(As, Bs)
    det = lambda sd: OrderedDict((n, v.detach().clone()) for n, v in sd.items())
    T, R, Ds = det(st.T), det(st.R), [det(d) for d in st.Ds]
    names = dict(T=list(T.keys()), R=list(R.keys()), D=[k for d in Ds for k in d.keys()])
    base = grads(cfg, T, R, Ds, As, Bs, torch.float32)
    saved = O._inorm
    try:
        for seed in range(a.seeds):
            gen = torch.Generator().manual_seed(100 + seed)

            def noisy_inorm(x):
                m = x.mean((2, 3), keepdim=True)
                v = x.var((2, 3), unbiased=False, keepdim=True)
                m = m * (1 + a.noise * torch.randn(m.shape, generator=gen, dtype=m.dtype)) + \
                    a.noise * x.abs().mean((2, 3), keepdim=True) * torch.randn(m.shape, generator=gen, dtype=m.dtype)
                v = v * (1 + a.noise * torch.randn(v.shape, generator=gen, dtype=v.dtype))
                return (x - m) * torch.rsqrt(v + 1e-5)
            O._inorm = noisy_inorm
            e = bucket_err(base, grads(cfg, T, R, Ds, As, Bs, torch.float32), names)
            print("seed %d: plane statistics perturbed by %.0e -> netT %.2e  netR %.2e  netD %.2e" % (seed, a.noise, e["T"], e["R"], e["D"]))
            sys.stdout.flush()
    finally:
        O._inorm = saved


if __name__ == "__main__":
    main()
