"""Training entry point with the reference's loop contract (reference train.py:28-81):
    python train.py --dataroot X --name N --model nemar --stn_type unet --gan_mode lsgan ...
With more than one id in --gpu_ids the script re-launches itself as one process per GPU (torchrun-style
env rendezvous on 127.0.0.1); each rank trains on its shard of every batch and gradients are exchanged by
one NCCL all-reduce per optimizer phase."""
import os
import subprocess
import sys
import time

import torch


def _maybe_spawn():
    """--gpu_ids a,b,c without an existing launcher => spawn one rank per id."""
    if "RANK" in os.environ:
        return False
    ids = None
    for i, a in enumerate(sys.argv):
        if a == "--gpu_ids" and i + 1 < len(sys.argv):
            ids = sys.argv[i + 1]
        elif a.startswith("--gpu_ids="):
            ids = a.split("=", 1)[1]
    n = len([s for s in (ids or "0").split(",") if s.strip() and int(s) >= 0])
    if n <= 1:
        return False
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), os.path.abspath(__file__)] + sys.argv[1:]
    sys.exit(subprocess.call(cmd))


class _Mapped:
    """iterable view of a loader with a function applied to every batch"""

    def __init__(self, loader, fn):
        self.loader, self.fn = loader, fn

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for d in self.loader:
            yield self.fn(d)


def main():
    _maybe_spawn()
    from nemar_b200.data import create_dataset
    from nemar_b200.data.prefetch import DevicePrefetcher
    from nemar_b200.engine import parallel
    from nemar_b200.models import create_model
    from nemar_b200.options.train_options import TrainOptions
    from nemar_b200.util.visualizer import Visualizer

    opt = TrainOptions().parse()
    world, rank, _ = parallel.init_process_group_from_env()
    parallel.check_global_batch(opt.batch_size)
    dataset = create_dataset(opt)
    dataset_size = len(dataset)
    print("The number of training images = %d" % dataset_size)
    model = create_model(opt)
    model.setup(opt)
    visualizer = Visualizer(opt)
    total_iters = 0
    # device input pipeline: each rank stages ITS shard of the next global batch (pinned, side stream) while it computes
    shard = (lambda d: {k: (parallel.shard_batch(v, rank, world) if torch.is_tensor(v) else v) for k, v in d.items()}) \
        if world > 1 else (lambda d: d)
    batches = DevicePrefetcher(_Mapped(dataset, shard), model.device) if opt.gpu_ids else _Mapped(dataset, shard)
    for epoch in range(opt.epoch_count, opt.niter + opt.niter_decay + 1):
        epoch_start_time = time.time()
        iter_data_time = time.time()
        epoch_iter = 0
        for i, data in enumerate(batches):
            iter_start_time = time.time()
            if total_iters % opt.print_freq == 0:
                t_data = iter_start_time - iter_data_time
            visualizer.reset()
            total_iters += opt.batch_size
            epoch_iter += opt.batch_size
            model.set_input(data)
            model.optimize_parameters()
            if total_iters % opt.display_freq == 0:
                model.compute_visuals()
                visualizer.display_current_results(model.get_current_visuals(), epoch, total_iters % opt.update_html_freq == 0)
            if total_iters % opt.print_freq == 0 and rank == 0:
                losses = model.get_current_losses()
                t_comp = (time.time() - iter_start_time) / opt.batch_size
                visualizer.print_current_losses(epoch, epoch_iter, losses, t_comp, t_data)
            if total_iters % opt.save_latest_freq == 0 and rank == 0:
                print("saving the latest model (epoch %d, total_iters %d)" % (epoch, total_iters))
                model.save_networks("iter_%d" % total_iters if opt.save_by_iter else "latest")
            iter_data_time = time.time()
        if epoch % opt.save_epoch_freq == 0 and rank == 0:
            print("saving the model at the end of epoch %d, iters %d" % (epoch, total_iters))
            model.save_networks("latest")
            model.save_networks(epoch)
        print("End of epoch %d / %d \t Time Taken: %d sec" % (epoch, opt.niter + opt.niter_decay, time.time() - epoch_start_time))
    parallel.shutdown(graph_mode=bool(getattr(opt, "cuda_graph", 0)))


if __name__ == "__main__":
    main()
