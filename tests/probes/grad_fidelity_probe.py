"""CPU probe (oracle only): how accurate are weight gradients of the training step in fp32 / bf16-autocast arithmetic,
measured against the fp64 oracle, as a function of the STATE the step is evaluated in?

  state "init-noise":   seeded N(0, 0.02) weights, A,B ~ U(-1,1) white noise           (round-1's test state)
  state "init-struct":  same weights, smooth images, B = intensity-remapped shift of A
  state "trained-K":    the fp32 oracle's weights after K steps on the structured batch

Per state it prints, per network, the median over weight tensors of |g - g64| / |g64| for the fp32 oracle and for
the oracle under torch.autocast(bfloat16), and the split of the T/R gradient into its L1 and adversarial parts.

    python tests/probes/grad_fidelity_probe.py [--steps 30] [--case c1_affine64]
"""
import argparse
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import nemar_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402


structured_batch = H.structured_batch


def grads(cfg, T, R, Ds, A, B, dtype, autocast=False):
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


def report(tag, cfg, T, R, Ds, A, B):
    truth = grads(cfg, T, R, Ds, A, B, torch.float64)
    g32 = grads(cfg, T, R, Ds, A, B, torch.float32)
    g16 = grads(cfg, T, R, Ds, A, B, torch.float32, autocast=True)
    names = dict(T=list(T.keys()), R=list(R.keys()), D=[k for d in Ds for k in d.keys()])
    out = {}
    for net in ("T", "R", "D"):
        e32, e16, tot_num32, tot_num16, tot_den = [], [], 0.0, 0.0, 0.0
        for k, t, a, b in zip(names[net], truth[net], g32[net], g16[net]):
            if not k.endswith(".weight"):
                continue
            n = float(t.norm()) + 1e-30
            e32.append(float((a - t).norm()) / n)
            e16.append(float((b - t).norm()) / n)
            tot_num32 += float((a - t).norm()) ** 2
            tot_num16 += float((b - t).norm()) ** 2
            tot_den += float(t.norm()) ** 2
        out[net] = dict(med32=float(np.median(e32)), med16=float(np.median(e16)), max16=float(np.max(e16)),
                        bucket32=(tot_num32 / tot_den) ** 0.5, bucket16=(tot_num16 / tot_den) ** 0.5)
    print("[%s]" % tag)
    for net, r in out.items():
        print("   net%s  fp32: median %.2e bucket %.2e | bf16-autocast: median %.3f max %.3f bucket %.3f" % (
            net, r["med32"], r["bucket32"], r["med16"], r["max16"], r["bucket16"]))
    sys.stdout.flush()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="c1_affine64")
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    kw, batch, _ = H.CASE_FLAGS[a.case]
    cfg = O.OracleConfig(**kw)
    T, R, Ds = O.make_states(cfg, seed=11)
    An, Bn = O.synthetic_batch(batch, cfg.height, cfg.width, seed=1)
    report("init-noise", cfg, T, R, Ds, An, Bn)
    As, Bs = structured_batch(batch, cfg.height, cfg.width)
    report("init-struct", cfg, T, R, Ds, As, Bs)
    st = O.OracleStep(cfg, T, R, Ds)
    for k in range(a.steps):
        losses = st.step(As, Bs)
        if (k + 1) % 10 == 0 or k + 1 == a.steps:
            print("   step %d losses %s" % (k + 1, {n: round(v, 4) for n, v in losses.items()}))
            det = lambda sd: OrderedDict((n, v.detach().clone()) for n, v in sd.items())
            report("trained-%d" % (k + 1), cfg, det(st.T), det(st.R), [det(d) for d in st.Ds], As, Bs)


if __name__ == "__main__":
    main()
