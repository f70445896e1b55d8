"""One process per GPU data parallelism (replaces nn.DataParallel, reference models/networks.py:108-111,
models/stn/__init__.py:30-35).  Each optimizer owns one flat gradient bucket; the bucket is summed across
ranks with a single all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests) right before the Adam
update, and the 1/world average is folded into the Adam kernel's grad_scale."""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group_from_env(backend=None):
    world, rank, local = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return world, rank, local


def shard_batch(t, rank, world):
    """rank r owns samples [r*b, (r+1)*b) of the global batch."""
    b = t.shape[0] // world
    return t[rank * b:(rank + 1) * b]


class BucketAllReduce:
    """grad_hook for FlatAdam: sum the bucket over ranks, return the averaging scale."""

    def __init__(self, group=None):
        self.group = group
        self.calls = 0

    def __call__(self, flat_grad):
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            self.calls += 1
            return 1.0 / dist.get_world_size(self.group)
        return 1.0
