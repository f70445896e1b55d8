"""CPU experiment (oracle only): does the ONE-PASS plane variance E[x^2] - E[x]^2 in fp32 explain the fp32 engine's
2e-3 gradient error on netT (tests/test_gpu_fidelity.py) when ATen's instance_norm (Welford-style) gives 1e-6?
Runs the fp32 oracle at the trained state twice — ATen instance_norm vs a one-pass fp32 restatement of the engine's
statistics (sum, sum of squares, rsqrt(var + eps)) — and prints both errors against the fp64 oracle.

    python tests/probes/onepass_var_probe.py [--steps 30]
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


def onepass_inorm(x):
    if x.dtype != torch.float32:
        return torch.nn.functional.instance_norm(x, eps=1e-5)
    hw = x.shape[2] * x.shape[3]
    s = x.sum((2, 3), keepdim=True)
    q = (x * x).sum((2, 3), keepdim=True)
    m = s / hw
    var = (q / hw - m * m).clamp_min(0.0)
    return (x - m) * torch.rsqrt(var + 1e-5)


def bucket_err(truth, g, names):
    out = {}
    for net in ("T", "R", "D"):
        num = den = 0.0
        for k, t, a in zip(names[net], truth[net], g[net]):
            if k.endswith(".weight"):
                num += float((a - t).norm()) ** 2
                den += float(t.norm()) ** 2
        out[net] = (num / den) ** 0.5
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="c1_affine64")
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    kw, batch, _ = H.CASE_FLAGS[a.case]
    cfg = O.OracleConfig(**kw)
    T, R, Ds = O.make_states(cfg, seed=11)
    As, Bs = H.structured_batch(batch, cfg.height, cfg.width)
    st = O.OracleStep(cfg, T, R, Ds)
    for _ in range(a.steps):
        st.step(As, Bs)
    det = lambda sd: OrderedDict((n, v.detach().clone()) for n, v in sd.items())
    T, R, Ds = det(st.T), det(st.R), [det(d) for d in st.Ds]
    names = dict(T=list(T.keys()), R=list(R.keys()), D=[k for d in Ds for k in d.keys()])
    truth = grads(cfg, T, R, Ds, As, Bs, torch.float64)
    base = bucket_err(truth, grads(cfg, T, R, Ds, As, Bs, torch.float32), names)
    saved = O._inorm
    O._inorm = onepass_inorm
    try:
        one = bucket_err(truth, grads(cfg, T, R, Ds, As, Bs, torch.float32), names)
    finally:
        O._inorm = saved
    print("fp32 oracle vs fp64, trained-%d state, norm-wise over weight tensors:" % a.steps)
    for net in ("T", "R", "D"):
        print("   net%s  ATen instance_norm %.2e | one-pass E[x^2]-E[x]^2 statistics %.2e" % (net, base[net], one[net]))


if __name__ == "__main__":
    main()
