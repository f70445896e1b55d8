"""--dataset_mode synthetic: paired A/B tensors ~ U(-1,1) of shape C x img_height x img_width, generated
deterministically per index (the benchmark / parity input of SURVEY.md section 8d)."""
import torch

from .base_dataset import BaseDataset


class SyntheticDataset(BaseDataset):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument("--synthetic_size", type=int, default=64, help="number of samples per epoch")
        parser.add_argument("--synthetic_seed", type=int, default=1)
        return parser

    def __init__(self, opt):
        BaseDataset.__init__(self, opt)
        self.n, self.seed = opt.synthetic_size, opt.synthetic_seed
        self.shape_a = (opt.input_nc, opt.img_height, opt.img_width)
        self.shape_b = (opt.output_nc, opt.img_height, opt.img_width)

    def __len__(self):
        return self.n

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(self.seed * 1000003 + index)
        a = torch.rand(self.shape_a, generator=g) * 2 - 1
        b = torch.rand(self.shape_b, generator=g) * 2 - 1
        return {"A": a, "B": b, "A_paths": "synthetic_A_%d" % index, "B_paths": "synthetic_B_%d" % index}
