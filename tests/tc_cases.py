"""tcgen05 engine cross-check cases: every conv geometry of the hot path, run through the C ABI on both engines
(bf16 storage) and compared; a few are also compared with ATen on CPU.  On the tensor-core engine 3/6-channel
inputs and 1/2/3-channel outputs are zero-padded to 16 channels."""
import torch
import torch.nn.functional as TF

from nemar_b200.engine import functional as F
from nemar_b200.engine import lib as L

# cin, cout, k, stride, pad, transposed, x_pad, n, h, w, act, out_f32, note
CASES = [
    (64, 64, 3, 1, 1, False, 0, 2, 16, 16, 0, False, "k3 s1 zero-pad 64ch (ResUnet mid)"),
    (64, 128, 1, 1, 0, False, 0, 16, 2, 2, 0, False, "1x1 bottleneck on 2x2 maps (TN=32 > batch)"),
    (256, 256, 3, 1, 1, False, 1, 2, 64, 64, 0, False, "netT ResnetBlock conv on a reflect-padded map"),
    (128, 128, 3, 1, 1, False, 1, 4, 2, 2, 0, False, "ResUnet bottleneck resblock 2x2 reflect"),
    (64, 128, 3, 2, 1, False, 0, 2, 32, 40, 0, False, "netT k3 s2 (TMA traversal stride)"),
    (128, 256, 3, 2, 1, False, 0, 2, 16, 16, 0, False, "netT k3 s2 #2"),
    (256, 128, 3, 2, 1, True, 0, 2, 16, 16, 0, False, "netT ConvTranspose k3 s2 p1 op1 (parity classes)"),
    (128, 64, 3, 2, 1, True, 0, 2, 8, 12, 0, False, "netT ConvTranspose #2"),
    (64, 128, 4, 2, 1, False, 0, 2, 32, 32, 0, False, "PatchGAN k4 s2"),
    (128, 256, 4, 2, 1, False, 0, 2, 16, 16, 0, False, "PatchGAN k4 s2 #2"),
    (256, 512, 4, 1, 1, False, 0, 2, 16, 16, 0, False, "PatchGAN k4 s1 (15x15 out)"),
    (128, 64, 3, 1, 1, False, 0, 2, 15, 17, 0, False, "ResUnet up conv on cat(64+64), odd extent"),
    (64, 64, 3, 1, 1, False, 1, 3, 8, 8, 0, False, "ResUnet resblock 64ch reflect 8x8, batch 3"),
    # ---- 16 / 32 / 96-channel chunks (SWIZZLE_32B / 64B), padded heads and fp32 tails
    (3, 64, 7, 1, 3, False, 3, 2, 32, 40, 0, False, "netT head 3->64 k7 on a reflect-padded image (cin padded to 16)"),
    (64, 3, 7, 1, 3, False, 3, 2, 32, 32, L.ACT_TANH, True, "netT tail 64->3 k7 + tanh, fp32 out (cout padded to 16)"),
    (6, 64, 4, 2, 1, False, 0, 2, 32, 32, L.ACT_LRELU, False, "PatchGAN first layer 6->64 k4 s2 + LeakyReLU"),
    (6, 32, 3, 1, 1, False, 0, 2, 24, 32, 0, False, "STN first layer 6->32"),
    (32, 32, 3, 1, 1, False, 1, 2, 32, 32, 0, False, "ResUnet 32-ch resblock conv, reflect"),
    (32, 64, 3, 1, 1, False, 0, 2, 16, 24, L.ACT_LRELU, False, "ResUnet down_2 32->64"),
    (96, 32, 3, 1, 1, False, 0, 2, 32, 32, 0, False, "ResUnet up_1 on cat(64+32)"),
    (32, 32, 1, 1, 0, False, 0, 2, 16, 16, L.ACT_LRELU, False, "ResUnet refine 1x1"),
    (32, 2, 3, 1, 1, False, 0, 2, 32, 32, 0, True, "offset head 32->2, fp32 out"),
    (512, 1, 4, 1, 1, False, 0, 2, 9, 9, 0, True, "PatchGAN prediction 512->1, fp32 out"),
    (256, 256, 3, 1, 1, False, 0, 2, 4, 4, 0, False, "affine STN 256->256 on 4x4"),
    # ---- CTA-pair (cta_group::2) geometries: odd tile counts, several 256-channel tiles on either side
    (256, 256, 3, 1, 1, False, 1, 3, 8, 16, 0, False, "256->256 reflect, 3 destination tiles (pair tail)"),
    (256, 512, 3, 1, 1, False, 0, 2, 16, 16, L.ACT_LRELU, False, "256->512: two 256-channel destination tiles, two wgrad pairs"),
    (512, 256, 3, 1, 1, False, 0, 2, 16, 16, 0, False, "512->256: 8 k-chunks per tap, two 256-channel wgrad N tiles"),
    # ---- rectangular 7 x 1 kernels: the row halves of the generator's k7 head / tail (column taps live in the channels)
    (21, 64, (7, 1), 1, 0, False, 0, 2, 38, 32, 0, False, "k7 head as 7x1 over 21 column-tap channels (padded to 32)"),
    (64, 21, (7, 1), 1, 0, False, 0, 2, 38, 38, 0, False, "k7 tail as 7x1 producing 21 column-tap partial sums"),
]

# cases with a 256-multiple channel count on a tensor-core destination (fprop: cout, dgrad: cin) or in the wgrad
PAIR_CASES = [i for i, c in enumerate(CASES) if (c[0] % 256 == 0 or c[1] % 256 == 0) and not c[11]]


def _mk(case, seed=0):
    cin, cout, k, stride, pad, transposed, x_pad, n, h, w = case[:10]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, cin, h, w), generator=g)
    kh, kw = (k, k) if isinstance(k, int) else k
    wshape = (cin, cout, kh, kw) if transposed else (cout, cin, kh, kw)
    wt = torch.randn(wshape, generator=g) * (cin * kh * kw) ** -0.5
    b = torch.randn(cout, generator=g) * 0.1
    return x, wt, b


def _engine(case, x, wt, b, dy, use_tc):
    cin, cout, k, stride, pad, transposed, x_pad, n, h, w, act, out_f32, _ = case
    xp = TF.pad(x, (x_pad,) * 4, mode="reflect") if x_pad else x
    cin_p = (cin + 15) // 16 * 16 if use_tc else cin
    cout_p = (cout + 15) // 16 * 16 if use_tc else cout
    xe = torch.zeros((n, xp.shape[2], xp.shape[3], cin_p), dtype=torch.bfloat16)
    xe[..., :cin] = xp.permute(0, 2, 3, 1).to(torch.bfloat16)
    xe = xe.cuda().requires_grad_(True)
    we, be = wt.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    stats = (act == L.ACT_NONE) and not out_f32
    cfg = F.ConvCfg(cin, cout, k, stride, pad, transposed, x_pad, act, stats, out_f32, 1 if transposed else 0, use_tc, cout_p)
    out = F.Conv2dFn.apply(xe, we, be, cfg, F.PackedWeights())
    y, st = out if stats else (out, None)
    if dy is None:
        dy = torch.randn(y.shape[:3] + (cout,), generator=torch.Generator().manual_seed(5)).to(torch.bfloat16).float()
    dyp = torch.zeros(y.shape, dtype=torch.float32)
    dyp[..., :cout] = dy
    (y.float() * dyp.cuda()).sum().backward()
    torch.cuda.synchronize()
    st = st.cpu()[:, :cout] if st is not None else torch.zeros(1)
    return (y.detach().float().cpu()[..., :cout], st, xe.grad.float().cpu()[..., :cin], we.grad.cpu(), be.grad.cpu(), dy,
            y.detach().float().cpu()[..., cout:], xe.grad.float().cpu()[..., cin:])


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-20))


def run_case(idx, vs_cpu=False):
    """-> dict of relative errors tc vs generic (and optionally vs ATen fp32 on bf16-rounded operands)."""
    case = CASES[idx]
    x, wt, b = _mk(case)
    yg, sg, dxg, dwg, dbg, dy, _, _ = _engine(case, x, wt, b, None, False)
    yt, st, dxt, dwt, dbt, _, ypad, dxpad = _engine(case, x, wt, b, dy, True)
    out = {"case": case[-1], "fwd": rel(yt, yg), "stats": rel(st, sg) if sg.numel() > 1 else 0.0, "dgrad": rel(dxt, dxg),
           "wgrad": rel(dwt, dwg), "bias": rel(dbt, dbg),
           "pad_nonzero": float(ypad.abs().max()) if ypad.numel() else 0.0}
    if case[10] == L.ACT_TANH and ypad.numel():
        out["pad_nonzero"] = 0.0 if float(ypad.abs().max()) == 0.0 else float(ypad.abs().max())
    if vs_cpu:
        cin, cout, k, stride, pad, transposed, x_pad, n, h, w, act, out_f32, _ = case
        xq = x.to(torch.bfloat16).float().requires_grad_(True)
        wq = wt.to(torch.bfloat16).float().requires_grad_(True)
        if transposed:
            y = TF.conv_transpose2d(xq, wq, b, stride=stride, padding=pad, output_padding=1)
        elif x_pad:
            y = TF.conv2d(TF.pad(xq, (x_pad,) * 4, mode="reflect"), wq, b, stride=stride)
        else:
            y = TF.conv2d(xq, wq, b, stride=stride, padding=pad)
        y = {0: lambda v: v, L.ACT_LRELU: lambda v: TF.leaky_relu(v, 0.2), L.ACT_TANH: torch.tanh}[act](y)
        (y * dy.permute(0, 3, 1, 2)).sum().backward()
        out["fwd_cpu"] = rel(yt.permute(0, 3, 1, 2), y.detach())
        out["wgrad_cpu"] = rel(dwt, wq.grad)
    return out


TOL = {"fwd": 6e-3, "stats": 1e-3, "dgrad": 6e-3, "wgrad": 4e-3, "bias": 1e-3, "fwd_cpu": 6e-3, "wgrad_cpu": 4e-3,
       "pad_nonzero": 0.0}


def check(res):
    return [k for k, v in res.items() if k in TOL and not (v <= TOL[k])]

# stride-1 k x k geometries the resident-patch kernel (NEMAR_TC_RP3=1) takes: <= 64 output channels per tile, weight pack
# and two patches within shared memory (fprop and/or dgrad side)
RP3_CASES = [i for i, c in enumerate(CASES) if c[2] != 1 and c[3] == 1 and not c[5] and (c[0] <= 96 or c[1] <= 96)]
