"""--dataset_mode registry + loader wrapper (reference data/__init__.py:18-93).  `--dataset_mode foo`
resolves to class FooDataset in nemar_b200/data/foo_dataset.py; items are dicts
{'A': CxHxW float in [-1,1], 'B': ..., 'A_paths': str, 'B_paths': str}."""
import importlib
import os

import torch.utils.data

from .base_dataset import BaseDataset


def find_dataset_using_name(dataset_name):
    module = importlib.import_module("%s.%s_dataset" % (__name__, dataset_name))
    wanted = (dataset_name.replace("_", "") + "dataset").lower()
    for name, cls in vars(module).items():
        if name.lower() == wanted and isinstance(cls, type) and issubclass(cls, BaseDataset):
            return cls
    raise NotImplementedError("%s_dataset.py must define a BaseDataset subclass named like %s" % (dataset_name, wanted))


def get_option_setter(dataset_name):
    return find_dataset_using_name(dataset_name).modify_commandline_options


def create_dataset(opt):
    return CustomDatasetDataLoader(opt).load_data()


class CustomDatasetDataLoader:
    def __init__(self, opt):
        self.opt = opt
        self.dataset = find_dataset_using_name(opt.dataset_mode)(opt)
        print("dataset [%s] was created" % type(self.dataset).__name__)
        # multi-rank jobs: every rank iterates the SAME global batches (train.py keeps its shard), so the shuffle must
        # not depend on how much of the default RNG a rank happened to consume, and a short last batch is dropped
        # (an uneven shard would drop samples silently or stall the other ranks' all-reduce)
        world = int(os.environ.get("WORLD_SIZE", "1"))
        gen = None
        if world > 1:
            gen = torch.Generator()
            gen.manual_seed(int(getattr(opt, "data_seed", 0)))
        self.dataloader = torch.utils.data.DataLoader(self.dataset, batch_size=opt.batch_size,
                                                      shuffle=not opt.serial_batches, num_workers=int(opt.num_threads),
                                                      pin_memory=bool(opt.gpu_ids), drop_last=world > 1, generator=gen)

    def load_data(self):
        return self

    def __len__(self):
        return min(len(self.dataset), self.opt.max_dataset_size)

    def __iter__(self):
        for i, batch in enumerate(self.dataloader):
            if i * self.opt.batch_size >= self.opt.max_dataset_size:
                break
            yield batch
