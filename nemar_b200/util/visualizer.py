"""Console / loss_log.txt side of the reference Visualizer (util/visualizer.py:211-227).  visdom, HTML
galleries and TensorBoard are observability side channels outside the hot path and are not rebuilt."""
import os
import time


class Visualizer:
    def __init__(self, opt):
        self.opt = opt
        self.name = opt.name
        self.log_name = os.path.join(opt.checkpoints_dir, opt.name, "loss_log.txt")
        os.makedirs(os.path.dirname(self.log_name), exist_ok=True)
        with open(self.log_name, "a") as f:
            f.write("================ Training Loss (%s) ================\n" % time.strftime("%c"))

    def reset(self):
        pass

    def display_current_results(self, visuals, epoch, save_result):
        pass

    def plot_current_losses(self, epoch, counter_ratio, losses):
        pass

    def print_current_losses(self, epoch, iters, losses, t_comp, t_data):
        message = "(epoch: %d, iters: %d, time: %.3f, data: %.3f) " % (epoch, iters, t_comp, t_data)
        message += " ".join("%s: %.3f" % kv for kv in losses.items())
        print(message)
        with open(self.log_name, "a") as f:
            f.write(message + "\n")
