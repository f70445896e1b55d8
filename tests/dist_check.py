"""Launched by torchrun on N ranks (tests/test_gpu_dist.py): the N-rank data-parallel engine == a single process on
the same global batch, at a state where that comparison is meaningful.

  1. every rank builds the engine from a DIFFERENT seed: the parameter broadcast in setup_optimizers must make the
     replicas identical;
  2. 20 real training steps on each rank's shard of a structured batch (one all-reduce per gradient bucket): the
     replicas must stay bit-synchronous, and the state leaves the ill-conditioned init (tests/test_gpu_fidelity.py);
  3. one more step on every rank, and the same step by rank 0 alone on the whole global batch from the same weights:
     the averaged gradient buckets must agree.

Ranks use NCCL with one GPU each when the box has >= N GPUs; on a smaller box all ranks share cuda:0 and exchange
through gloo (NCCL refuses two ranks per device) — same sharding / bucket / optimizer code either way."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nemar_b200.engine import functional as F  # noqa: E402
from nemar_b200.engine import parallel  # noqa: E402
from oracle import nemar_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402

WARM_STEPS = 20


def step_and_capture(model, A, B):
    """one optimize_parameters; -> the (exchanged, averaged) gradient buckets the two Adam launches consumed"""
    cap = {}
    saved = {}
    for name, opt in (("TR", model.optimizer_TR), ("D", model.optimizer_D)):
        inner = saved[name] = opt.grad_hook

        def hook(flat, inner=inner, name=name):
            scale = inner(flat)
            cap[name] = (flat * scale).detach().clone()
            return scale
        opt.grad_hook = hook
    model.set_input({"A": A, "B": B, "A_paths": "", "B_paths": ""})
    model.optimize_parameters()
    torch.cuda.synchronize()
    for name, opt in (("TR", model.optimizer_TR), ("D", model.optimizer_D)):
        opt.grad_hook = saved[name]
    return cap


def main():
    world = int(os.environ["WORLD_SIZE"])
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    world, rank, local = parallel.init_process_group_from_env(backend)
    dev = local % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    per = 2
    kw, _, extra = H.CASE_FLAGS["c1_affine64"]
    cfg = O.OracleConfig(**kw)
    A, B = H.structured_batch(per * world, cfg.height, cfg.width)

    def build(batch, seed):
        from nemar_b200.models import create_model
        opt = H.engine_opt(cfg, batch, extra, "fp32", "generic", gpu_ids=str(dev), ckpt="/tmp/nemar_dist_%d" % rank)
        torch.manual_seed(seed)                  # engine init draws from the default RNG: different on every rank
        return create_model(opt)

    dp = build(per, 100 + rank)
    ok = True
    # 1. broadcast at construction
    for name, o in (("TR", dp.optimizer_TR), ("D", dp.optimizer_D)):
        ref = o.flat_p.detach().clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(ref, o.flat_p))
        if not same:
            print("rank %d: %s parameters differ from rank 0 after construction" % (rank, name))
        ok = ok and same
    # 2. train
    a, b = parallel.shard_batch(A, rank, world), parallel.shard_batch(B, rank, world)
    for _ in range(WARM_STEPS):
        step_and_capture(dp, a, b)
    for name, o in (("TR", dp.optimizer_TR), ("D", dp.optimizer_D)):
        ref = o.flat_p.detach().clone()
        dist.broadcast(ref, src=0)
        drift = float((ref - o.flat_p).abs().max())
        if drift != 0.0:
            print("rank %d: %s replicas drifted by %.3e after %d steps" % (rank, name, drift, WARM_STEPS))
        ok = ok and drift == 0.0
    p0 = {name: o.flat_p.detach().clone() for name, o in (("TR", dp.optimizer_TR), ("D", dp.optimizer_D))}
    opt0 = {name: o.state_dict() for name, o in (("TR", dp.optimizer_TR), ("D", dp.optimizer_D))}
    # 3. same step, N ranks vs one process
    g_dp = step_and_capture(dp, a, b)
    calls = dp.allreduce.calls
    # (every rank constructs the reference model: construction broadcasts parameters, a collective; only rank 0 steps it)
    single = build(per * world, 7)
    if rank == 0:
        single.optimizer_TR.grad_hook = lambda flat: 1.0       # no exchange: the whole global batch is local
        single.optimizer_D.grad_hook = lambda flat: 1.0
        single.optimizer_TR.flat_p.copy_(p0["TR"])
        single.optimizer_D.flat_p.copy_(p0["D"])
        # the T/R phase differentiates through the discriminator AFTER its Adam update: same moments, same step number
        single.optimizer_TR.load_state_dict(opt0["TR"])
        single.optimizer_D.load_state_dict(opt0["D"])
        F.bump_weights_epoch()
        g_1 = step_and_capture(single, A, B)
        for k in ("TR", "D"):
            rel = float((g_dp[k] - g_1[k]).norm() / g_1[k].norm())
            print("bucket %s: |dp - single| / |single| = %.3e   (all-reduce calls so far: %d, backend %s)" % (k, rel, calls, backend))
            ok = ok and rel < 2e-3      # same math, different fp32 summation order
        ok = ok and calls == 2 * (WARM_STEPS + 1)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "OK" if float(flag) == 1.0 else "FAILED", "world", world, "backend", backend)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
