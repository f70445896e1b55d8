"""Test-time flags (reference options/test_options.py:10-24)."""
from .base_options import BaseOptions


class TestOptions(BaseOptions):
    isTrain = False

    def initialize(self, parser):
        parser = BaseOptions.initialize(self, parser)
        parser.add_argument("--ntest", type=int, default=float("inf"))
        parser.add_argument("--results_dir", type=str, default="./results/")
        parser.add_argument("--aspect_ratio", type=float, default=1.0)
        parser.add_argument("--phase", type=str, default="test")
        parser.add_argument("--eval", action="store_true")
        parser.add_argument("--num_test", type=int, default=50)
        parser.set_defaults(model="nemar")
        parser.set_defaults(load_size=parser.get_default("crop_size"))
        return parser
