"""Kernel-level A/B tool: times fprop / dgrad / wgrad of single conv geometries through the C ABI (CUDA events, L2
flushed between launches) under several environment variants — seconds of GPU time per comparison instead of a
whole-model bench.  Geometries: the C2 layers of the hot path (batch 16), or --case N for tests/tc_cases.py entries.

  python scripts/kbench.py --variants "" "NEMAR_TC_RP3=1" --layers stn32 stn96 resblock
  python scripts/kbench.py --variants "NEMAR_TC_PAIR=0" "NEMAR_TC_PAIR=1" --layers resblock d512

Each variant runs in its own child process (the engine reads its switches once per process); a child that traps or
hangs is reported and does not take the others down.  Numbers are per launch: microseconds, TFLOP/s (2*MAC) and
algorithmic GB/s (operands + result once)."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name: (cin, cout, k, stride, pad, transposed, x_pad, n, h, w, act, out_f32)   — h, w = input extent without halo
LAYERS = {
    "resblock": (256, 256, 3, 1, 1, False, 1, 16, 64, 64, 0, False),      # netT ResnetBlock conv (72 % of the FLOPs)
    "down2": (128, 256, 3, 2, 1, False, 0, 16, 128, 128, 0, False),
    "down1": (64, 128, 3, 2, 1, False, 0, 16, 256, 256, 0, False),
    "up1": (256, 128, 3, 2, 1, True, 0, 16, 64, 64, 0, False),
    "up2": (128, 64, 3, 2, 1, True, 0, 16, 128, 128, 0, False),
    "stn32": (32, 32, 3, 1, 1, False, 1, 16, 256, 256, 0, False),          # ResUnet full-resolution resblock conv
    "stn6": (6, 32, 3, 1, 1, False, 0, 16, 256, 256, 0, False),
    "stn96": (96, 32, 3, 1, 1, False, 0, 16, 256, 256, 0, False),
    "stn32_64": (32, 64, 3, 1, 1, False, 0, 16, 128, 128, 0, False),
    "stn64": (64, 64, 3, 1, 1, False, 1, 16, 128, 128, 0, False),
    "stn128_64": (128, 64, 3, 1, 1, False, 0, 16, 128, 128, 0, False),
    "offset": (32, 2, 3, 1, 1, False, 0, 16, 256, 256, 0, True),
    "d64": (6, 64, 4, 2, 1, False, 0, 48, 256, 256, 2, False),
    "d128": (64, 128, 4, 2, 1, False, 0, 48, 128, 128, 0, False),
    "d256": (128, 256, 4, 2, 1, False, 0, 48, 64, 64, 0, False),
    "d512": (256, 512, 4, 1, 1, False, 0, 48, 32, 32, 0, False),
    "head1x1": (147, 64, 1, 1, 0, False, 0, 16, 256, 256, 0, False),
    "tail1x1": (64, 147, 1, 1, 0, False, 0, 16, 262, 262, 0, False),
}


def child(spec, reps):
    import torch
    import torch.nn.functional as TF
    from nemar_b200.engine import functional as F
    cin, cout, k, stride, pad, transposed, x_pad, n, h, w, act, out_f32 = spec
    g = torch.Generator().manual_seed(0)
    x = torch.randn((n, cin, h, w), generator=g)
    xp = TF.pad(x, (x_pad,) * 4, mode="reflect") if x_pad else x
    cin_p, cout_p = (cin + 15) // 16 * 16, (cout + 15) // 16 * 16
    xe = torch.zeros((n, xp.shape[2], xp.shape[3], cin_p), dtype=torch.bfloat16)
    xe[..., :cin] = xp.permute(0, 2, 3, 1).to(torch.bfloat16)
    xe = xe.cuda().requires_grad_(True)
    wshape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    we = (torch.randn(wshape, generator=g) * (cin * k * k) ** -0.5).cuda().requires_grad_(True)
    be = torch.zeros(cout).cuda().requires_grad_(True)
    cfg = F.ConvCfg(cin, cout, k, stride, pad, transposed, x_pad, act, False, out_f32, 1 if transposed else 0, True, cout_p)
    packed = F.PackedWeights()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    from nemar_b200.engine import lib as L
    L.TIMER.min_flops = 0.0
    out = {}
    y = None
    for it in range(reps + 2):
        flush.zero_()
        L.TIMER.enable(1 if it >= 2 else 0)
        xe.grad = None
        we.grad = None
        y = F.Conv2dFn.apply(xe, we, be, cfg, packed)
        dy = torch.ones_like(y)
        flush.zero_()
        y.backward(dy)
        for d in L.TIMER.collect().values():       # records are keyed on the kernel instance; regroup by pass
            for op, po in d["ops"].items():
                o = out.setdefault(op, {"ms": 0.0, "n": 0, "flops": po["flops"] / max(po["n"], 1)})
                o["ms"] += po["ms"]
                o["n"] += po["n"]
    L.TIMER.enable(0)
    pix_out = y.shape[0] * y.shape[1] * y.shape[2]
    pix_in = xe.shape[0] * xe.shape[1] * xe.shape[2]
    esz_out = 4 if out_f32 else 2
    byts = {"fprop": pix_in * cin_p * 2 + pix_out * cout_p * esz_out, "dgrad": pix_out * cout_p * 2 + pix_in * cin_p * 2,
            "wgrad": pix_in * cin_p * 2 + pix_out * cout_p * 2}
    res = {}
    for op, o in out.items():
        us = 1e3 * o["ms"] / max(o["n"], 1)
        res[op] = {"us": round(us, 1), "tflops": round(o["flops"] / (us * 1e-6) / 1e12, 1), "gbs": round(byts[op] / (us * 1e-6) / 1e9, 0)}
    print("KBENCH " + json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", nargs="*", default=["resblock", "stn32", "stn96", "stn64", "d512"])
    ap.add_argument("--case", type=int, nargs="*", default=[], help="tests/tc_cases.py entries instead of --layers")
    ap.add_argument("--variants", nargs="*", default=[""], help='environment variants, e.g. "" "NEMAR_TC_RP3=1 NEMAR_TC_RP3_STAGES=4"')
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--timeout", type=int, default=120)
    ap.add_argument("--child", type=str, default=None)
    args = ap.parse_args()
    if args.child is not None:
        child(tuple(json.loads(args.child)), args.reps)
        return
    specs = [(name, LAYERS[name]) for name in args.layers]
    if args.case:
        from tests import tc_cases
        specs = [("case%d" % i, tc_cases.CASES[i][:12]) for i in args.case]
    for name, spec in specs:
        print("%s  %s" % (name, spec))
        for var in args.variants:
            env = dict(os.environ)
            for kv in var.split():
                k, v = kv.split("=", 1)
                env[k] = v
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", json.dumps(list(spec)), "--reps", str(args.reps)],
                                   capture_output=True, text=True, timeout=args.timeout, env=env)
                line = [l for l in p.stdout.splitlines() if l.startswith("KBENCH ")]
                if line:
                    r = json.loads(line[0][7:])
                    print("   %-44s %s" % (var or "(default)", "  ".join("%s %7.1f us %6.1f TF/s %5.0f GB/s" % (op, v["us"], v["tflops"], v["gbs"]) for op, v in sorted(r.items()))))
                else:
                    print("   %-44s FAILED rc=%d %s" % (var or "(default)", p.returncode, (p.stderr or p.stdout)[-300:].replace("\n", " | ")))
            except subprocess.TimeoutExpired:
                print("   %-44s TIMEOUT" % (var or "(default)"))
            sys.stdout.flush()


if __name__ == "__main__":
    main()
