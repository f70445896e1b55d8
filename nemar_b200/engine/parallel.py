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
        # the CUDA device is the caller's choice (BaseOptions.parse maps LOCAL_RANK through --gpu_ids; bench.py and the
        # tests call torch.cuda.set_device themselves): NCCL binds to whatever device is current at the first collective
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return world, rank, local


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def broadcast_buffers(tensors, src=0):
    """Make every rank start from rank `src`'s values (flat parameter buffers after construction / a checkpoint load):
    replicas then stay bit-synchronous because every rank applies the same averaged gradient."""
    if world_size() > 1:
        for t in tensors:
            dist.broadcast(t, src=src)


def check_global_batch(batch_size):
    """A global batch must split evenly: an uneven last shard would silently drop samples or hang the all-reduce."""
    w = world_size()
    if batch_size % w != 0:
        raise ValueError("--batch_size %d is not divisible by the %d ranks of this job" % (batch_size, w))
    return batch_size // w


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


def shutdown(graph_mode=False):
    """Leave a multi-rank job.  After a captured step with NCCL all-reduces inside, destroy_process_group() never
    returns (measured at 2 ranks, torch 2.11 / NCCL 2.28): such jobs synchronise, meet at a barrier and exit the
    process directly; everything else tears the group down normally."""
    if not dist.is_initialized():
        return
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dist.barrier()
    if graph_mode and dist.get_backend() == "nccl":
        import sys
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    dist.destroy_process_group()
