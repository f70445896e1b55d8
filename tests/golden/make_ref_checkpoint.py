"""Write a checkpoint WITH THE REFERENCE'S OWN CODE (BaseModel.save_networks, models/base_model.py:148-164, run in
place from /root/reference on CPU) for the load-compatibility test.  Build container only:

    python tests/golden/make_ref_checkpoint.py

Output: tests/golden/refckpt/latest_net_{T,D}.pth (reduced widths: --ngf 8 --ndf 8, resnet_6blocks; the affine STN's
file is 5 MB at any width and is left out — its keys are covered by the state-dict tests) and refckpt_meta.npz with
the reference's outputs on a seeded input, so that a loader can be checked end to end.
"""
import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def main():
    NEMARModel, TrainOptions = MG.import_reference()
    torch.manual_seed(1234)
    case = dict(stn_type="affine", height=64, width=64, batch=1,
                extra=["--netG", "resnet_6blocks", "--ngf", "8", "--ndf", "8"])
    model, opt = MG.build_reference(case, NEMARModel, TrainOptions)      # the reference's own init_weights
    out_dir = os.path.join(HERE, "refckpt")
    os.makedirs(out_dir, exist_ok=True)
    model.save_dir = "/tmp/nemar_refckpt"
    os.makedirs(model.save_dir, exist_ok=True)
    model.save_networks("latest")                                        # reference code writes the files
    for name in ("T", "D"):
        shutil.copy(os.path.join(model.save_dir, "latest_net_%s.pth" % name), out_dir)
    g = torch.Generator().manual_seed(3)
    A = torch.rand((1, 3, 64, 64), generator=g) * 2 - 1
    B = torch.rand((1, 3, 64, 64), generator=g) * 2 - 1
    with torch.no_grad():
        fake_B = model.netT(A)
        pred = model.netD(torch.cat([A, B], 1))
    np.savez_compressed(os.path.join(out_dir, "refckpt_meta.npz"), A=A.numpy(), B=B.numpy(), fake_B=fake_B.numpy(),
                        pred=pred.numpy())
    for f in sorted(os.listdir(out_dir)):
        print(f, os.path.getsize(os.path.join(out_dir, f)), "bytes")


if __name__ == "__main__":
    main()
