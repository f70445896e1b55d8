"""CPU experiment (fp64 oracle only): how much of the T / R weight gradient hangs on ONE activation-derivative bit?
The fp32 engine's T / R gradients differ from the fp64 oracle by 4e-6 in some runs and 5e-4 ... 2.4e-3 in others
(profiles/r02_fp32_gradient_error_probes.txt), in discrete steps.  ReLU / LeakyReLU / max-pool derivatives are discontinuous:
an element whose pre-activation lies within the two implementations' rounding difference (~1e-6 after InstanceNorm)
gets derivative 1 in one and 0 (0.2) in the other.  This probe (a) counts such borderline elements at the trained state
and (b) flips the derivative bit of the single ReLU / LeakyReLU input closest to zero, one call site at a time
(generator, STN and discriminator; the discriminator's passes of the T/R phase feed both netT and netR), and reports the
change of the T / R gradient (norm-wise over weight tensors).

    python tests/probes/relu_flip_probe.py [--steps 30]
"""
import argparse
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import nemar_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.probes.onepass_var_probe import bucket_err  # noqa: E402

REAL_RELU = torch.nn.functional.relu
REAL_LRELU = torch.nn.functional.leaky_relu


class Flip:
    """relu with the derivative bit of ONE element (the input closest to zero) of ONE call toggled"""
    call, target, near = 0, -1, []

    @staticmethod
    def relu(x, inplace=False):
        Flip.call += 1
        ax = x.detach().abs()
        Flip.near.append((Flip.call, x.numel(), int((ax < 1e-6).sum()), int((ax < 1e-5).sum()), float(ax.min())))
        if Flip.call != Flip.target:
            return REAL_RELU(x)
        mask = (x.detach() > 0).to(x.dtype)
        i = int(ax.flatten().argmin())
        mask.view(-1)[i] = 1.0 - mask.view(-1)[i]
        return x * mask

    @staticmethod
    def leaky_relu(x, negative_slope=0.01, inplace=False):
        """the discriminator's LeakyReLU(0.2): derivative 1 or 0.2 (call sites share the counter with relu)"""
        Flip.call += 1
        ax = x.detach().abs()
        Flip.near.append((Flip.call, x.numel(), int((ax < 1e-6).sum()), int((ax < 1e-5).sum()), float(ax.min())))
        if Flip.call != Flip.target:
            return REAL_LRELU(x, negative_slope)
        pos = x.detach() > 0
        i = int(ax.flatten().argmin())
        pos.view(-1)[i] = ~pos.view(-1)[i]
        return x * torch.where(pos, torch.ones_like(x), torch.full_like(x, negative_slope))


def tr_grads(cfg, T, R, Ds, A, B, target):
    def go():
        Tc, Rc, Dc = O.cast_states(torch.float64, T, R, Ds)
        st = O.OracleStep(cfg, Tc, Rc, Dc)
        Flip.call, Flip.target, Flip.near = 0, target, []
        torch.nn.functional.relu, torch.nn.functional.leaky_relu = Flip.relu, Flip.leaky_relu
        try:
            st.step(A.double(), B.double())
        finally:
            torch.nn.functional.relu, torch.nn.functional.leaky_relu = REAL_RELU, REAL_LRELU
        return st.grads
    return O.run_in_dtype(torch.float64, go)


def main():
    ap = argparse.ArgumentParser()
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
    names = dict(T=list(T.keys()), R=list(R.keys()), D=[k for d in Ds for k in d.keys()])
    base = tr_grads(cfg, T, R, Ds, A, B, -1)
    near = list(Flip.near)
    tot = sum(n for _, n, _, _, _ in near)
    print("ReLU / LeakyReLU inputs of one step: %d calls, %d elements; |x| < 1e-6: %d, |x| < 1e-5: %d" % (
        len(near), tot, sum(c for _, _, c, _, _ in near), sum(c for _, _, _, c, _ in near)))
    for call, n, c6, c5, mn in near:
        if c5:
            print("   call %2d: %7d elements, %d below 1e-6, %d below 1e-5, closest %.2e" % (call, n, c6, c5, mn))
    # flip the closest-to-zero element of the calls that have the smallest inputs
    for call, n, c6, c5, mn in sorted(near, key=lambda t: t[4])[:10]:
        g = tr_grads(cfg, T, R, Ds, A, B, call)
        e = bucket_err(base, g, names)
        print("derivative bit of the element closest to zero in activation call %2d (|x| = %.1e, %d elements) flipped -> netT %.2e  netR %.2e" % (
            call, mn, n, e["T"], e["R"]))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
