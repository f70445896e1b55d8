"""Generate golden vectors by running the REFERENCE ITSELF (imported in place from /root/reference; nothing
is copied) on seeded inputs and seeded state dicts.  Run in the build container only:

    python tests/golden/make_golden.py

Outputs tests/golden/<case>.npz: per-step losses of the reference's optimize_parameters, first-step output
images (full for small cases, strided samples for large ones) and per-tensor checksums of the updated
parameters.  tests/test_oracle_golden.py replays them against oracle/nemar_oracle.py.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("NEMAR_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import nemar_oracle as O  # noqa: E402

CASES = {
    # C1 of BASELINE.json: 64x64, affine STN, resnet_6blocks, batch 2
    "c1_affine64": dict(stn_type="affine", n_blocks=6, height=64, width=64, batch=2, steps=3,
                        extra=["--netG", "resnet_6blocks"]),
    # C2 shape at batch 1: 256x256, unet STN (cfg A), resnet_9blocks, live smoothness term
    "c2_unet256": dict(stn_type="unet", n_blocks=9, height=256, width=256, batch=1, steps=2, lambda_smooth=200.0,
                       extra=["--netG", "resnet_9blocks", "--lambda_smooth", "200.0"]),
    # C4 ingredients at 256: bilateral alpha, 2 reg levels, 2 discriminator scales, reduced widths
    "c4_multires256": dict(stn_type="unet", n_blocks=3, height=256, width=256, batch=1, steps=2, lambda_smooth=200.0,
                           alpha=1.0, multires_reg=2, multi_resolution=2, ngf=16, ndf=16,
                           extra=["--netG", "resnet_3blocks", "--lambda_smooth", "200.0", "--stn_bilateral_alpha", "1.0",
                                  "--stn_multires_reg", "2", "--multi_resolution", "2", "--ngf", "16", "--ndf", "16"]),
    # C4 geometry at its real size: 512x512, THREE discriminator scales, bilateral alpha 1.0, three regulariser
    # levels, resnet_9blocks (reduced widths keep the CPU replay in seconds)
    "c4_ms3_512": dict(stn_type="unet", n_blocks=9, height=512, width=512, batch=1, steps=2, lambda_smooth=200.0,
                       alpha=1.0, multires_reg=3, multi_resolution=3, ngf=16, ndf=16,
                       extra=["--netG", "resnet_9blocks", "--lambda_smooth", "200.0", "--stn_bilateral_alpha", "1.0",
                              "--stn_multires_reg", "3", "--multi_resolution", "3", "--ngf", "16", "--ndf", "16"]),
    # C5 shape: 1024x1024 dense deformation field, resnet_9blocks, batch 1 (reduced widths keep the CPU replay short)
    "c5_1024": dict(stn_type="unet", n_blocks=9, height=1024, width=1024, batch=1, steps=1, lambda_smooth=200.0,
                    ngf=8, ndf=8,
                    extra=["--netG", "resnet_9blocks", "--lambda_smooth", "200.0", "--ngf", "8", "--ndf", "8"]),
    # ragged edge case: non-square, not a multiple of 128 (the seven floor-pools of the ResUnet go 288 -> 2 and
    # 384 -> 3 through odd extents 9 and 3; the up path resizes to each skip's size), batch 2
    "ragged288x384": dict(stn_type="unet", n_blocks=6, height=288, width=384, batch=2, steps=2, lambda_smooth=200.0,
                          ngf=16, ndf=16,
                          extra=["--netG", "resnet_6blocks", "--lambda_smooth", "200.0", "--ngf", "16", "--ndf", "16"]),
}


def import_reference():
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    for name in ("dominate", "dominate.tags"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    import models as ref_models  # noqa: F401
    from models.nemar_model import NEMARModel
    from options.train_options import TrainOptions
    return NEMARModel, TrainOptions


def build_reference(case, NEMARModel, TrainOptions):
    argv = ["--dataroot", "none", "--name", "golden", "--checkpoints_dir", "/tmp/nemar_golden", "--gpu_ids", "-1",
            "--gan_mode", "lsgan", "--no_dropout", "--stn_type", case["stn_type"], "--img_height", str(case["height"]),
            "--img_width", str(case["width"]), "--batch_size", str(case["batch"]), "--dataset_mode", "base"] + case["extra"]
    old = sys.argv
    sys.argv = ["make_golden"] + argv
    try:
        import data as ref_data
        ref_data.get_option_setter = lambda name: (lambda parser, is_train: parser)   # reference ships no dataset
        opt = TrainOptions().parse()
    finally:
        sys.argv = old
    return NEMARModel(opt), opt


def main():
    NEMARModel, TrainOptions = import_reference()
    torch.manual_seed(0)
    only = sys.argv[1:]          # optional: generate only the named cases (each case is seeded independently)
    for name, case in CASES.items():
        if only and name not in only:
            continue
        cfg = O.OracleConfig(stn_type=case["stn_type"], n_blocks=case["n_blocks"], height=case["height"],
                             width=case["width"], lambda_smooth=case.get("lambda_smooth", 0.0), alpha=case.get("alpha", 0.0),
                             multires_reg=case.get("multires_reg", 1), multi_resolution=case.get("multi_resolution", 1),
                             ngf=case.get("ngf", 64), ndf=case.get("ndf", 64))
        T, R, Ds = O.make_states(cfg, seed=11)
        model, opt = build_reference(case, NEMARModel, TrainOptions)
        model.netT.load_state_dict(T)
        model.netR.load_state_dict(R)
        model.netD.load_state_dict(Ds[0])
        for d_ref, d in zip(model.netD_multiresolution, Ds[1:]):
            d_ref.load_state_dict(d)
        A, B = O.synthetic_batch(case["batch"], case["height"], case["width"], seed=1)
        out = {}
        losses = []
        for step in range(case["steps"]):
            model.set_input({"A": A, "B": B, "A_paths": "", "B_paths": ""})
            model.optimize_parameters()
            losses.append([v for v in model.get_current_losses().values()])
            if step == 0:
                stride = 1 if case["height"] <= 64 else 8
                for k in ("fake_B", "registered_real_A", "fake_TR_B", "fake_RT_B"):
                    out["img_" + k] = getattr(model, k).detach()[:, :, ::stride, ::stride].numpy().copy()
                out["img_stride"] = np.array(stride)
        out["loss_names"] = np.array(list(model.get_current_losses().keys()))
        out["losses"] = np.array(losses, dtype=np.float64)
        for tag, net in (("T", model.netT), ("R", model.netR), ("D", model.netD)):
            sd = net.state_dict()
            out["psum_" + tag] = np.array([float(v.double().sum()) for v in sd.values()])
            out["pabs_" + tag] = np.array([float(v.double().abs().sum()) for v in sd.values()])
            out["pkeys_" + tag] = np.array(list(sd.keys()))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "losses", np.array2string(out["losses"], precision=5), "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
