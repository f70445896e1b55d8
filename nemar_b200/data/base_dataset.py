"""Dataset ABC (reference data/base_dataset.py:13-60).  Image-file transforms are out of scope for the
engine (the hot path consumes tensors); user datasets return the {'A','B'} dict themselves."""
from abc import ABC, abstractmethod

import torch.utils.data as data


class BaseDataset(data.Dataset, ABC):
    def __init__(self, opt):
        self.opt = opt
        self.root = opt.dataroot

    @staticmethod
    def modify_commandline_options(parser, is_train):
        return parser

    @abstractmethod
    def __len__(self):
        return 0

    @abstractmethod
    def __getitem__(self, index):
        ...
