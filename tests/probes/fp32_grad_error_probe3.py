"""GPU probe, third question.  Known so far (probes 1 and 2): the fp32 engine's T/R gradients are 2e-3 / 8e-4 off the fp64
oracle only when the discriminator moves (lr != 0), the engine's updated discriminator EQUALS the oracle's, and the
error does not change when the oracle is given the engine's discriminator — so the engine's T/R phase is not consistent
with the discriminator it holds.  Which mechanism?  One engine step per variant, same truth:
  base        : defaults
  repack      : every packed-weight cache entry invalidated right after optimizer_D.step() (lazy per-layer re-pack)
  no_batch_d  : --batch_d 0 (one discriminator pass per image pair instead of one batched pass per phase)
  no_streams  : weight gradients and the STN regressor on the main stream

    python tests/probes/fp32_grad_error_probe3.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import nemar_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_gpu_fidelity import _trained_state  # noqa: E402
from tests.probes.fp32_grad_error_probe2 import oracle_step, bucket  # noqa: E402


def engine_errors(truth, T, R, Ds, A, B, flags=(), repack=False, streams=True):
    from nemar_b200.engine import functional as F
    from nemar_b200.engine.config import CONFIG
    model, _, _, _ = H.build_case("c1_affine64", precision="fp32", conv_engine="generic", more_flags=flags)
    H.load_states(model, T, R, Ds)
    saved = CONFIG.wgrad_stream
    if not streams:
        CONFIG.wgrad_stream = False
    if repack:
        real = model.optimizer_D.step

        def step_and_invalidate():
            real()
            F.bump_weights_epoch()
        model.optimizer_D.step = step_and_invalidate
    try:
        H.run_engine_steps(model, A, B, 1)
    finally:
        CONFIG.wgrad_stream = saved
    return bucket(model.netT, truth.grads["T"]), bucket(model.netR, truth.grads["R"])


def main():
    cfg, T, R, Ds, A, B = _trained_state()
    truth = oracle_step(cfg, T, R, Ds, A, B)
    for tag, kw in (("base", {}), ("repack", dict(repack=True)), ("no_batch_d", dict(flags=("--batch_d", "0"))),
                    ("no_streams", dict(flags=("--stream_overlap", "0"), streams=False))):
        t, r = engine_errors(truth, T, R, Ds, A, B, **kw)
        print("PROBE3 %-11s netT %.3e  netR %.3e" % (tag, t, r))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
