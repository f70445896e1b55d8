"""tcgen05 engine cross-check cases: every conv geometry of the hot path that the tensor-core engine takes,
run through the C ABI on both engines (bf16 storage) and compared; a few are also compared with ATen on CPU."""
import torch
import torch.nn.functional as TF

from nemar_b200.engine import functional as F
from nemar_b200.engine import lib as L

# cin, cout, k, stride, pad, transposed, x_pad, n, h, w, note
CASES = [
    (64, 64, 3, 1, 1, False, 0, 2, 16, 16, "k3 s1 zero-pad 64ch (ResUnet mid)"),
    (64, 128, 1, 1, 0, False, 0, 16, 2, 2, "1x1 bottleneck on 2x2 maps (TN=32 > batch)"),
    (256, 256, 3, 1, 1, False, 1, 2, 64, 64, "netT ResnetBlock conv on a reflect-padded map"),
    (128, 128, 3, 1, 1, False, 1, 4, 2, 2, "ResUnet bottleneck resblock 2x2 reflect"),
    (64, 128, 3, 2, 1, False, 0, 2, 32, 40, "netT k3 s2 (TMA traversal stride)"),
    (128, 256, 3, 2, 1, False, 0, 2, 16, 16, "netT k3 s2 #2"),
    (256, 128, 3, 2, 1, True, 0, 2, 16, 16, "netT ConvTranspose k3 s2 p1 op1 (parity classes)"),
    (128, 64, 3, 2, 1, True, 0, 2, 8, 12, "netT ConvTranspose #2"),
    (64, 128, 4, 2, 1, False, 0, 2, 32, 32, "PatchGAN k4 s2"),
    (128, 256, 4, 2, 1, False, 0, 2, 16, 16, "PatchGAN k4 s2 #2"),
    (256, 512, 4, 1, 1, False, 0, 2, 16, 16, "PatchGAN k4 s1 (15x15 out)"),
    (128, 64, 3, 1, 1, False, 0, 2, 15, 17, "ResUnet up conv on cat(64+64), odd extent"),
    (64, 64, 3, 1, 1, False, 1, 3, 8, 8, "ResUnet resblock 64ch reflect 8x8, batch 3"),
]


def _mk(case, seed=0):
    cin, cout, k, stride, pad, transposed, x_pad, n, h, w, _ = case
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, cin, h, w), generator=g)
    wshape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    wt = torch.randn(wshape, generator=g) * (cin * k * k) ** -0.5
    b = torch.randn(cout, generator=g) * 0.1
    return x, wt, b


def _engine(case, x, wt, b, dy, use_tc):
    cin, cout, k, stride, pad, transposed, x_pad, n, h, w, _ = case
    xp = TF.pad(x, (x_pad,) * 4, mode="reflect") if x_pad else x
    xe = xp.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda().requires_grad_(True)
    we, be = wt.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    cfg = F.ConvCfg(cin, cout, k, stride, pad, transposed, x_pad, L.ACT_NONE, True, False, 1 if transposed else 0, use_tc)
    y, stats = F.Conv2dFn.apply(xe, we, be, cfg, F.PackedWeights())
    if dy is None:
        dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16).cuda()
    (y.float() * dy.float()).sum().backward()
    torch.cuda.synchronize()
    return y.detach().float().cpu(), stats.cpu(), xe.grad.float().cpu(), we.grad.cpu(), be.grad.cpu(), dy


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-20))


def run_case(idx, vs_cpu=False):
    """-> dict of relative errors tc vs generic (and optionally vs ATen fp32 on bf16-rounded operands)."""
    case = CASES[idx]
    x, wt, b = _mk(case)
    yg, sg, dxg, dwg, dbg, dy = _engine(case, x, wt, b, None, False)
    yt, st, dxt, dwt, dbt, _ = _engine(case, x, wt, b, dy, True)
    out = {"case": case[-1], "fwd": rel(yt, yg), "stats": rel(st, sg), "dgrad": rel(dxt, dxg), "wgrad": rel(dwt, dwg),
           "bias": rel(dbt, dbg)}
    if vs_cpu:
        cin, cout, k, stride, pad, transposed, x_pad, n, h, w, _ = case
        xq = x.to(torch.bfloat16).float().requires_grad_(True)
        wq = wt.to(torch.bfloat16).float().requires_grad_(True)
        if transposed:
            y = TF.conv_transpose2d(xq, wq, b, stride=stride, padding=pad, output_padding=1)
        elif x_pad:
            y = TF.conv2d(TF.pad(xq, (x_pad,) * 4, mode="reflect"), wq, b, stride=stride)
        else:
            y = TF.conv2d(xq, wq, b, stride=stride, padding=pad)
        (y * dy.float().cpu().permute(0, 3, 1, 2)).sum().backward()
        out["fwd_cpu"] = rel(yt.permute(0, 3, 1, 2), y.detach())
        out["wgrad_cpu"] = rel(dwt, wq.grad)
    return out


TOL = {"fwd": 6e-3, "stats": 1e-3, "dgrad": 6e-3, "wgrad": 2e-3, "bias": 1e-3, "fwd_cpu": 6e-3, "wgrad_cpu": 2e-3}


def check(res):
    return [k for k, v in res.items() if k in TOL and not (v <= TOL[k])]
