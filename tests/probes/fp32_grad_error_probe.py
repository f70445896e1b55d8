"""GPU probe: where does the fp32 engine's weight-gradient error against the fp64 oracle come from?  At the trained
state (tests/test_gpu_fidelity.py) the fp32 ORACLE is 1e-6 from fp64 but the fp32 engine 2e-3 on netT / 8e-4 on netR
(netD 3e-6).  Prints the error per weight tensor of netT / netR in layer order, once with the default learning rate
(the T/R phase differentiates through the discriminator AFTER its Adam update) and once with lr = 0 on both sides
(the discriminator stays put: no coupling through its update).

    python tests/probes/fp32_grad_error_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import nemar_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_gpu_fidelity import _trained_state  # noqa: E402
from tests.test_gpu_model import _oracle_grads  # noqa: E402


def run(lr):
    cfg, T, R, Ds, A, B = _trained_state()
    cfg = O.OracleConfig(**dict(cfg.__dict__, lr=lr))
    model, _, _, _ = H.build_case("c1_affine64", precision="fp32", conv_engine="generic", more_flags=("--lr", str(lr)))
    H.load_states(model, T, R, Ds)
    H.run_engine_steps(model, A, B, 1)
    truth = _oracle_grads(cfg, T, R, Ds, A, B, torch.float64)
    print("== lr = %g" % lr)
    for tag, net in (("T", model.netT), ("R", model.netR), ("D", model.netD)):
        num = den = 0.0
        rows = []
        for i, (k, p) in enumerate(net.named_parameters()):
            t = truth[tag][i]
            g = p.grad.detach().double().cpu()
            e = float((g - t).norm()) / (float(t.norm()) + 1e-30)
            rows.append("      %-40s |g| %.3e  err %.2e" % (k, float(t.norm()), e))
            if k.endswith(".weight"):
                num += float((g - t).norm()) ** 2
                den += float(t.norm()) ** 2
        print("   net%s bucket %.3e" % (tag, (num / den) ** 0.5))
        if tag != "D":
            print("\n".join(rows))
    sys.stdout.flush()


if __name__ == "__main__":
    run(2e-4)
    run(0.0)
