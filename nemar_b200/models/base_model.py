"""BaseModel contract (reference models/base_model.py:8-234): device selection, setup/schedulers,
checkpoint save/load with the reference's file names and state_dict keys, loss / visual getters,
set_requires_grad.  Re-implemented for one-process-per-GPU execution (no DataParallel unwrapping)."""
import os
from abc import ABC, abstractmethod
from collections import OrderedDict

import torch

from . import networks
from ..engine import functional as F


class BaseModel(ABC):
    def __init__(self, opt):
        self.opt = opt
        self.gpu_ids = opt.gpu_ids
        self.isTrain = opt.isTrain
        self.device = torch.device("cuda", torch.cuda.current_device()) if self.gpu_ids else torch.device("cpu")
        self.save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        self.loss_names, self.model_names, self.visual_names = [], [], []
        self.optimizers, self.image_paths = [], []
        self.metric = 0

    @staticmethod
    def modify_commandline_options(parser, is_train):
        return parser

    @abstractmethod
    def set_input(self, input):
        ...

    @abstractmethod
    def forward(self):
        ...

    @abstractmethod
    def optimize_parameters(self):
        ...

    def setup(self, opt):
        if self.isTrain:
            self.schedulers = [networks.get_scheduler(o, opt) for o in self.optimizers]
        if not self.isTrain or opt.continue_train:
            self.load_networks("iter_%d" % opt.load_iter if opt.load_iter > 0 else opt.epoch)
        self.print_networks(opt.verbose)

    def _nets(self):
        for name in self.model_names:
            if isinstance(name, str):
                yield name, getattr(self, "net" + name)

    def eval(self):
        for _, net in self._nets():
            net.eval()

    def test(self):
        with torch.no_grad():
            self.forward()
            self.compute_visuals()

    def compute_visuals(self):
        pass

    def get_image_paths(self):
        return self.image_paths

    def update_learning_rate(self):
        for s in self.schedulers:
            s.step(self.metric)
        print("learning rate = %.7f" % self.optimizers[0].param_groups[0]["lr"])

    def get_current_visuals(self):
        out = OrderedDict()
        for name in self.visual_names:
            if isinstance(name, str):
                v = getattr(self, name)
                if isinstance(v, list):
                    for i, x in enumerate(v):
                        out["%s_%d" % (name, i)] = x
                else:
                    out[name] = v
        return out

    def get_current_losses(self):
        out = OrderedDict()
        for name in self.loss_names:
            if isinstance(name, str):
                out[name] = float(getattr(self, "loss_" + name))
        return out

    def save_networks(self, epoch):
        """<epoch>_net_<name>.pth with the reference's keys; tensors are cloned off the flat buffers."""
        os.makedirs(self.save_dir, exist_ok=True)
        for name, net in self._nets():
            sd = OrderedDict((k, v.detach().cpu().clone()) for k, v in net.state_dict().items())
            torch.save(sd, os.path.join(self.save_dir, "%s_net_%s.pth" % (epoch, name)))

    def load_networks(self, epoch):
        for name, net in self._nets():
            path = os.path.join(self.save_dir, "%s_net_%s.pth" % (epoch, name))
            print("loading the model from %s" % path)
            sd = torch.load(path, map_location="cpu")
            if hasattr(sd, "_metadata"):
                del sd._metadata
            for k in [k for k in sd if k.endswith(("running_mean", "running_var", "num_batches_tracked"))]:
                sd.pop(k)     # legacy InstanceNorm buffers (reference base_model.py:166-178)
            net.load_state_dict(sd)
        F.bump_weights_epoch()

    def print_networks(self, verbose):
        print("---------- Networks initialized -------------")
        for name, net in self._nets():
            if verbose:
                print(net)
            print("[Network %s] Total number of parameters : %.3f M" % (name, sum(p.numel() for p in net.parameters()) / 1e6))
        print("-----------------------------------------------")

    def set_requires_grad(self, nets, requires_grad=False):
        if not isinstance(nets, list):
            nets = [nets]
        for net in nets:
            if net is not None:
                for p in net.parameters():
                    p.requires_grad = requires_grad
