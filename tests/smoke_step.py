"""__graft_entry__.smoke(): one small optimize_parameters of the hot path on cuda:0 (64x64, affine STN,
resnet_6blocks, batch 2 — BASELINE.json configs[0]) in bf16 on the default engine, checked against the oracle."""
import numpy as np
import torch


def run():
    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    from oracle import nemar_oracle as O
    from tests import helpers as H
    model, cfg, (T, R, Ds), (A, B) = H.build_case("c1_affine64", precision="bf16", conv_engine="auto")
    losses = np.array(H.run_engine_steps(model, A, B, 1))[0]
    ref = np.array(list(O.OracleStep(cfg, T, R, Ds).step(A, B).values()))
    err = np.abs(losses - ref) / (np.abs(ref) + 0.5)
    print("smoke losses engine:", np.round(losses, 4))
    print("smoke losses oracle:", np.round(ref, 4))
    assert np.all(np.isfinite(losses)) and float(err.max()) < 0.05, "engine deviates from the oracle: %s" % err
    print("smoke OK (max relative deviation %.4f)" % float(err.max()))


if __name__ == "__main__":
    run()
