"""GPU probe, fourth question: the fp32 engine's T/R gradient error against the fp64 oracle was 2.3e-3 / 8e-4 in five
runs and 4e-6 / 3e-6 in one (probe 3's first model).  Is that run-to-run noise of the engine (atomics-ordered sums
deciding ReLU masks on degenerate planes), or does it depend on whether the oracle or the engine ran first in the process?
    python tests/probes/fp32_grad_error_probe4.py engine_first|oracle_first
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from tests import helpers as H  # noqa: E402
from tests.test_gpu_fidelity import _trained_state  # noqa: E402
from tests.probes.fp32_grad_error_probe2 import oracle_step, bucket  # noqa: E402


def main():
    order = sys.argv[1]
    cfg, T, R, Ds, A, B = _trained_state()
    truth = oracle_step(cfg, T, R, Ds, A, B) if order == "oracle_first" else None
    model, _, _, _ = H.build_case("c1_affine64", precision="fp32", conv_engine="generic")
    H.load_states(model, T, R, Ds)
    H.run_engine_steps(model, A, B, 1)
    if truth is None:
        truth = oracle_step(cfg, T, R, Ds, A, B)
    print("PROBE4 %-13s netT %.3e  netR %.3e" % (order, bucket(model.netT, truth.grads["T"]), bucket(model.netR, truth.grads["R"])))


if __name__ == "__main__":
    main()
