"""Training flags (reference options/train_options.py:10-54)."""
from .base_options import BaseOptions

TRAIN = [
    ("--display_freq", dict(type=int, default=400)),
    ("--display_ncols", dict(type=int, default=4)),
    ("--display_id", dict(type=int, default=-1)),
    ("--display_server", dict(type=str, default="http://localhost")),
    ("--display_env", dict(type=str, default="main")),
    ("--display_port", dict(type=int, default=8097)),
    ("--update_html_freq", dict(type=int, default=1000)),
    ("--print_freq", dict(type=int, default=100)),
    ("--no_html", dict(action="store_true")),
    ("--save_latest_freq", dict(type=int, default=5000)),
    ("--save_epoch_freq", dict(type=int, default=5)),
    ("--save_by_iter", dict(action="store_true")),
    ("--continue_train", dict(action="store_true")),
    ("--epoch_count", dict(type=int, default=1)),
    ("--phase", dict(type=str, default="train")),
    ("--niter", dict(type=int, default=100)),
    ("--niter_decay", dict(type=int, default=100)),
    ("--beta1", dict(type=float, default=0.5)),
    ("--lr", dict(type=float, default=0.0002)),
    ("--gan_mode", dict(type=str, default="vanilla")),
    ("--pool_size", dict(type=int, default=50)),
    ("--lr_policy", dict(type=str, default="linear")),
    ("--lr_decay_iters", dict(type=int, default=50)),
]


class TrainOptions(BaseOptions):
    isTrain = True

    def initialize(self, parser):
        parser = BaseOptions.initialize(self, parser)
        for flag, kw in TRAIN:
            parser.add_argument(flag, **kw)
        return parser
