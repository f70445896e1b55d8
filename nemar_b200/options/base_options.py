"""Flags shared by training and test (reference options/base_options.py:20-138).  Same names, types and
defaults; model- and dataset-specific flags are injected through the registries in a second parse."""
import argparse
import os

import torch

from .. import data, models

# (flag, kwargs)
COMMON = [
    ("--dataroot", dict(required=True, help="dataset root (ignored by --dataset_mode synthetic)")),
    ("--name", dict(type=str, default="experiment_name")),
    ("--gpu_ids", dict(type=str, default="0", help="e.g. 0 | 0,1,2 | -1 for CPU; >1 ids => one process per GPU")),
    ("--checkpoints_dir", dict(type=str, default="./checkpoints")),
    ("--model", dict(type=str, default="nemar")),
    ("--input_nc", dict(type=int, default=3)),
    ("--output_nc", dict(type=int, default=3)),
    ("--ngf", dict(type=int, default=64)),
    ("--ndf", dict(type=int, default=64)),
    ("--netD", dict(type=str, default="basic")),
    ("--netG", dict(type=str, default="resnet_9blocks")),
    ("--n_layers_D", dict(type=int, default=3)),
    ("--norm", dict(type=str, default="instance")),
    ("--init_type", dict(type=str, default="normal")),
    ("--init_gain", dict(type=float, default=0.02)),
    ("--no_dropout", dict(action="store_true")),
    ("--dataset_mode", dict(type=str, default="unaligned")),
    ("--direction", dict(type=str, default="AtoB")),
    ("--serial_batches", dict(action="store_true")),
    ("--num_threads", dict(default=4, type=int)),
    ("--batch_size", dict(type=int, default=1)),
    ("--load_size", dict(type=int, default=286)),
    ("--img_height", dict(type=int, default=288)),
    ("--img_width", dict(type=int, default=384)),
    ("--crop_size", dict(type=int, default=256)),
    ("--max_dataset_size", dict(type=int, default=float("inf"))),
    ("--preprocess", dict(type=str, default="resize_and_crop")),
    ("--no_flip", dict(action="store_true")),
    ("--display_winsize", dict(type=int, default=256)),
    ("--epoch", dict(type=str, default="latest")),
    ("--load_iter", dict(type=int, default=0)),
    ("--verbose", dict(action="store_true")),
    ("--suffix", dict(default="", type=str)),
]


class BaseOptions:
    isTrain = False

    def __init__(self):
        self.initialized = False

    def initialize(self, parser):
        for flag, kw in COMMON:
            parser.add_argument(flag, **kw)
        self.initialized = True
        return parser

    def gather_options(self, argv=None):
        parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
        parser = self.initialize(parser)
        opt, _ = parser.parse_known_args(argv)
        parser = models.get_option_setter(opt.model)(parser, self.isTrain)
        opt, _ = parser.parse_known_args(argv)
        parser = data.get_option_setter(opt.dataset_mode)(parser, self.isTrain)
        self.parser = parser
        return parser.parse_args(argv)

    def print_options(self, opt):
        lines = ["----------------- Options ---------------"]
        for k, v in sorted(vars(opt).items()):
            default = self.parser.get_default(k)
            note = "\t[default: %s]" % str(default) if v != default else ""
            lines.append("{:>25}: {:<30}{}".format(str(k), str(v), note))
        lines.append("----------------- End -------------------")
        message = "\n".join(lines)
        print(message)
        expr_dir = os.path.join(opt.checkpoints_dir, opt.name)
        os.makedirs(expr_dir, exist_ok=True)
        with open(os.path.join(expr_dir, "%s_opt.txt" % opt.phase), "wt") as f:
            f.write(message + "\n")

    def parse(self, argv=None, quiet=False):
        opt = self.gather_options(argv)
        opt.isTrain = self.isTrain
        if opt.suffix:
            opt.name = opt.name + "_" + opt.suffix.format(**vars(opt))
        if not quiet:
            self.print_options(opt)
        ids = [int(s) for s in opt.gpu_ids.split(",") if s.strip() != ""]
        opt.gpu_ids = [i for i in ids if i >= 0]
        if opt.gpu_ids and torch.cuda.is_available():
            # one process per GPU: under a multi-process launch every rank drives gpu_ids[LOCAL_RANK]
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(opt.gpu_ids[local % len(opt.gpu_ids)])
        self.opt = opt
        return opt
