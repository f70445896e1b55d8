"""GPU probe, second question: the fp32 engine's T/R gradient error appears only when the discriminator moves
(tests/probes/fp32_grad_error_probe.py: 2e-3 at lr = 2e-4, 1e-4 at lr = 0).  Is the engine's UPDATED discriminator different
from the oracle's, or does the engine's T/R backward see a different discriminator than its forward?
  (1) per tensor: max |D'_engine - D'_oracle64| / lr;
  (2) the fp64 oracle's T/R gradients recomputed with the ENGINE's updated discriminator substituted for its own:
      if the engine then agrees to ~1e-4, the whole discrepancy is the update (Adam on rounding-level gradients);
      if not, the engine's backward through D is inconsistent with the D it evaluated.

    python tests/probes/fp32_grad_error_probe2.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import nemar_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_gpu_fidelity import _trained_state  # noqa: E402


def oracle_step(cfg, T, R, Ds, A, B, d_new=None):
    """fp64 oracle step; d_new: discriminator weights to put in place of the oracle's own Adam update"""
    def go():
        Tc, Rc, Dc = O.cast_states(torch.float64, T, R, Ds)
        st = O.OracleStep(cfg, Tc, Rc, Dc)
        if d_new is not None:
            def put(grads):
                with torch.no_grad():
                    for p, w in zip([p for d in st.Ds for p in d.values()], d_new):
                        p.copy_(w.double())
            st.opt_D.step = put
        st.step(A.double(), B.double())
        return st
    return O.run_in_dtype(torch.float64, go)


def bucket(net, truth):
    num = den = 0.0
    for i, (k, p) in enumerate(net.named_parameters()):
        if k.endswith(".weight"):
            g = p.grad.detach().double().cpu()
            num += float((g - truth[i]).norm()) ** 2
            den += float(truth[i].norm()) ** 2
    return (num / den) ** 0.5


def main():
    lr = 2e-4
    cfg, T, R, Ds, A, B = _trained_state()
    model, _, _, _ = H.build_case("c1_affine64", precision="fp32", conv_engine="generic")
    H.load_states(model, T, R, Ds)
    H.run_engine_steps(model, A, B, 1)
    st = oracle_step(cfg, T, R, Ds, A, B)
    d_eng = [p.detach().cpu().clone() for p in model.netD.parameters()]
    print("(1) updated discriminator, engine vs fp64 oracle: max |diff| / lr per tensor (|g| = norm of the oracle's gradient)")
    for (k, _), we, wo, g in zip(model.netD.named_parameters(), d_eng, list(st.Ds[0].values()), st.grads["D"]):
        d = (we.double() - wo.detach()).abs()
        print("      %-28s max %.3e  mean %.3e   |g| %.2e  elements off by > 0.5 lr: %d of %d" % (
            k, float(d.max()) / lr, float(d.mean()) / lr, float(g.norm()), int((d > 0.5 * lr).sum()), d.numel()))
    print("(2) engine T/R gradients vs the fp64 oracle with ITS OWN updated D : netT %.3e  netR %.3e" % (
        bucket(model.netT, st.grads["T"]), bucket(model.netR, st.grads["R"])))
    st2 = oracle_step(cfg, T, R, Ds, A, B, d_new=d_eng)
    print("    engine T/R gradients vs the fp64 oracle with the ENGINE's updated D: netT %.3e  netR %.3e" % (
        bucket(model.netT, st2.grads["T"]), bucket(model.netR, st2.grads["R"])))


if __name__ == "__main__":
    main()
