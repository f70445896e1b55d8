"""Debug (GPU): is the engine's T/R gradient right at batch 4 (the single-process side of tests/dist_check.py)?"""
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import nemar_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_gpu_model import _oracle_grads  # noqa: E402
from tests.test_gpu_fidelity import _bucket_err  # noqa: E402

kw, _, extra = H.CASE_FLAGS["c1_affine64"]
cfg = O.OracleConfig(**kw)
for batch in (2, 4):
    T, R, Ds = O.make_states(cfg, seed=11)
    A, B = H.structured_batch(batch, cfg.height, cfg.width)
    st = O.OracleStep(cfg, T, R, Ds)
    for _ in range(20):
        st.step(A, B)
    det = lambda sd: OrderedDict((k, v.detach().clone()) for k, v in sd.items())
    T, R, Ds = det(st.T), det(st.R), [det(d) for d in st.Ds]
    from nemar_b200.models import create_model
    opt = H.engine_opt(cfg, batch, extra, "fp32", "generic")
    model = create_model(opt)
    H.load_states(model, T, R, Ds)
    # same Adam state as the oracle's optimizers would matter for D': give the engine the oracle's moments
    H.run_engine_steps(model, A, B, 1)
    # oracle with FRESH Adam state on the same weights (like the engine above)
    truth = _oracle_grads(cfg, T, R, Ds, A, B, torch.float64)
    for tag, net in (("T", model.netT), ("R", model.netR), ("D", model.netD)):
        print("batch %d net%s: engine fp32 vs fp64 oracle: bucket %.3e median %.3e" % ((batch, tag) + _bucket_err(net, truth[tag])))
