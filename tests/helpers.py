"""Shared helpers: build the engine's NEMARModel from flags and load the oracle's seeded state dicts."""
import torch

from oracle import nemar_oracle as O

CASE_FLAGS = {
    "c1_affine64": (dict(stn_type="affine", n_blocks=6, height=64, width=64), 2, ["--netG", "resnet_6blocks"]),
    "c2_unet256": (dict(stn_type="unet", n_blocks=9, height=256, width=256, lambda_smooth=200.0), 1,
                   ["--netG", "resnet_9blocks", "--lambda_smooth", "200.0"]),
    "c4_multires256": (dict(stn_type="unet", n_blocks=3, height=256, width=256, lambda_smooth=200.0, alpha=1.0,
                            multires_reg=2, multi_resolution=2, ngf=16, ndf=16), 1,
                       ["--netG", "resnet_3blocks", "--lambda_smooth", "200.0", "--stn_bilateral_alpha", "1.0",
                        "--stn_multires_reg", "2", "--multi_resolution", "2", "--ngf", "16", "--ndf", "16"]),
    "c4_ms3_512": (dict(stn_type="unet", n_blocks=9, height=512, width=512, lambda_smooth=200.0, alpha=1.0,
                        multires_reg=3, multi_resolution=3, ngf=16, ndf=16), 1,
                   ["--netG", "resnet_9blocks", "--lambda_smooth", "200.0", "--stn_bilateral_alpha", "1.0",
                    "--stn_multires_reg", "3", "--multi_resolution", "3", "--ngf", "16", "--ndf", "16"]),
    "c5_1024": (dict(stn_type="unet", n_blocks=9, height=1024, width=1024, lambda_smooth=200.0, ngf=8, ndf=8), 1,
                ["--netG", "resnet_9blocks", "--lambda_smooth", "200.0", "--ngf", "8", "--ndf", "8"]),
    "ragged288x384": (dict(stn_type="unet", n_blocks=6, height=288, width=384, lambda_smooth=200.0, ngf=16, ndf=16), 2,
                      ["--netG", "resnet_6blocks", "--lambda_smooth", "200.0", "--ngf", "16", "--ndf", "16"]),
}


def engine_opt(cfg, batch, extra, precision="fp32", conv_engine="generic", gpu_ids="0", ckpt="/tmp/nemar_b200_ckpt"):
    from nemar_b200.options.train_options import TrainOptions
    argv = ["--dataroot", "none", "--name", "t", "--checkpoints_dir", ckpt, "--gpu_ids", gpu_ids, "--gan_mode", "lsgan",
            "--no_dropout", "--stn_type", cfg.stn_type, "--img_height", str(cfg.height), "--img_width", str(cfg.width),
            "--batch_size", str(batch), "--dataset_mode", "synthetic", "--precision", precision, "--conv_engine",
            conv_engine] + list(extra)
    return TrainOptions().parse(argv, quiet=True)


def build_case(name, precision="fp32", conv_engine="generic", seed=11, gpu_ids="0", more_flags=()):
    """-> (engine model with the seeded weights loaded, oracle cfg, (T,R,Ds) states, (A,B) batch)"""
    from nemar_b200.models import create_model
    kw, batch, extra = CASE_FLAGS[name]
    cfg = O.OracleConfig(**kw)
    opt = engine_opt(cfg, batch, list(extra) + list(more_flags), precision, conv_engine, gpu_ids)
    model = create_model(opt)
    T, R, Ds = O.make_states(cfg, seed=seed)
    load_states(model, T, R, Ds)
    A, B = O.synthetic_batch(batch, cfg.height, cfg.width, seed=1)
    return model, cfg, (T, R, Ds), (A, B)


def load_states(model, T, R, Ds):
    from nemar_b200.engine import functional as F
    model.netT.load_state_dict(T)
    model.netR.load_state_dict(R)
    model.netD.load_state_dict(Ds[0])
    for net, d in zip(model.netD_multiresolution, Ds[1:]):
        net.load_state_dict(d)
    F.bump_weights_epoch()


def run_engine_steps(model, A, B, steps):
    losses = []
    for _ in range(steps):
        model.set_input({"A": A, "B": B, "A_paths": "", "B_paths": ""})
        model.optimize_parameters()
        losses.append(list(model.get_current_losses().values()))
    torch.cuda.synchronize()
    return losses


def structured_batch(n, h, w, seed=5, shift=4):
    """Smooth random images in [-1, 1]; B = another intensity mapping of A, shifted by `shift` pixels along x (a
    registration problem with a known answer, unlike the white-noise batches of the step goldens)."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand((n, 3, h // 8, w // 8), generator=g) * 2 - 1
    A = torch.nn.functional.interpolate(low, (h, w), mode="bicubic", align_corners=False).clamp(-1, 1)
    A = (A + 0.05 * (torch.rand((n, 3, h, w), generator=g) * 2 - 1)).clamp(-1, 1)
    mapped = torch.tanh(1.5 * A.flip(1)) * 0.9
    B = torch.roll(mapped, shifts=shift, dims=3)
    return A, B
