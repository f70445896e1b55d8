"""--stn_type / --stn_cfg registry (reference models/stn/__init__.py:10-48)."""
import torch

from .affine_stn import AffineSTN
from .unet_stn import UnetSTN

sampling_align_corners = False
sampling_mode = "bilinear"


def modify_commandline_options(parser, is_train=True):
    parser.add_argument("--stn_cfg", type=str, default="A", help="configuration letter of the STN")
    parser.add_argument("--stn_type", type=str, default="affine", help="unet | affine")
    if is_train:
        parser.add_argument("--stn_bilateral_alpha", type=float, default=0.0,
                            help="bilateral coefficient of the smoothness loss (unet STN)")
        parser.add_argument("--stn_no_identity_init", action="store_true",
                            help="do not start the unet STN from the identity transformation")
        parser.add_argument("--stn_multires_reg", type=int, default=1,
                            help="number of resolutions the smoothness regulariser is applied on")
    return parser


def define_stn(opt, stn_type="affine"):
    nc_a = opt.input_nc if opt.direction == "AtoB" else opt.output_nc
    nc_b = opt.output_nc if opt.direction == "AtoB" else opt.input_nc
    stn = None
    if stn_type == "affine":
        stn = AffineSTN(nc_a, nc_b, opt.img_height, opt.img_width, opt.stn_cfg, opt.init_type)
    if stn_type == "unet":
        stn = UnetSTN(nc_a, nc_b, opt.img_height, opt.img_width, opt.stn_cfg, opt.init_type, opt.stn_bilateral_alpha,
                      (not opt.stn_no_identity_init), opt.stn_multires_reg)
    if stn is None:
        raise NotImplementedError("STN type [%s] is not recognized" % stn_type)
    if len(opt.gpu_ids) > 0:     # one replica per process: no DataParallel wrapper
        assert torch.cuda.is_available()
        stn.to(torch.device("cuda", torch.cuda.current_device()))
    return stn
