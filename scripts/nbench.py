"""InstanceNorm-pass micro-benchmark (A/B tool like scripts/kbench.py): times the statistics pass, the normalise+act
forward, and the two backward passes through the C ABI on the C2 shapes, L2 flushed between launches.

  python scripts/nbench.py --by_variant --variants "" "NEMAR_LEAN_PIPE=3" "NEMAR_LEAN_PIPE_APPLY=4" --shapes res256 stn32

GB/s = algorithmic bytes (every operand once) / time."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name: (n, h, w, c, out_pad, residual)
SHAPES = {
    "res256": (16, 64, 64, 256, 1, False),      # ResnetBlock first norm (+ReLU, reflect halo)
    "res256r": (16, 64, 64, 256, 1, True),      # ResnetBlock second norm (+residual)
    "up128": (16, 128, 128, 128, 0, False),
    "head64": (16, 256, 256, 64, 0, False),
    "stn32": (16, 256, 256, 32, 1, False),
    "d512": (48, 31, 31, 512, 0, False),
}


def child(spec, reps):
    import torch
    from nemar_b200.engine import lib as L
    from nemar_b200.engine.lib import call, fptr, view, stream
    n, h, w, c, op, has_res = spec
    dev = "cuda"
    x = torch.randn((n, h, w, c), device=dev).to(torch.bfloat16)
    res = torch.randn((n, h + 2, w + 2, c), device=dev).to(torch.bfloat16) if has_res else None
    y = torch.empty((n, h + 2 * op, w + 2 * op, c), dtype=torch.bfloat16, device=dev)
    dy = torch.randn((n, h + 2 * op, w + 2 * op, c), device=dev).to(torch.bfloat16)
    dx = torch.empty_like(x)
    dres = torch.empty_like(res) if has_res else None
    stats = torch.zeros((n, c, 2), device=dev)
    red = torch.zeros((n, c, 2), device=dev)
    db = torch.zeros(c, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    act = L.ACT_NONE if has_res else L.ACT_RELU
    px_in, px_out = n * h * w, n * (h + 2 * op) * (w + 2 * op)
    byts = {"stats": px_in * c * 2, "fwd": (px_in * (2 if has_res else 1) + px_out) * c * 2,
            "bwd_reduce": (px_in + px_out) * c * 2, "bwd_apply": (px_in * (3 if has_res else 2) + px_out) * c * 2}
    ops = {
        "stats": lambda: call("nemar_instnorm_stats", view(x), fptr(stats), stream()),
        "fwd": lambda: call("nemar_norm_act_fwd", view(x), fptr(stats), act, view(res, 1) if has_res else None, view(y, op), L.PAD_REFLECT, stream()),
        "bwd_reduce": lambda: call("nemar_norm_act_bwd_reduce", view(x), fptr(stats), act, view(dy, op), L.PAD_REFLECT, fptr(red), stream()),
        "bwd_apply": lambda: call("nemar_norm_act_bwd_apply", view(x), fptr(stats), act, view(dy, op), L.PAD_REFLECT, fptr(red), view(dx),
                                  view(dres, 1) if has_res else None, 4 if has_res else 0, fptr(db), stream()),
    }
    out = {}
    for name, fn in ops.items():
        ts = []
        for it in range(reps + 2):
            if name in ("stats", "bwd_reduce"):
                (stats if name == "stats" else red).zero_()
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        out[name] = {"us": round(ms * 1e3, 1), "gbs": round(byts[name] / (ms / 1e3) / 1e9)}
    print("NBENCH " + json.dumps(out), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", nargs="*", default=["res256", "res256r", "head64", "stn32"])
    ap.add_argument("--variants", nargs="*", default=[""])
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--child", type=str, default=None)
    ap.add_argument("--by_variant", action="store_true", help="one process per variant (all shapes inside it)")
    args = ap.parse_args()
    if args.child is not None:
        if args.child == "null":
            for spec in json.loads(os.environ["NBENCH_SHAPES"]):
                child(tuple(spec[1:]), args.reps)
        else:
            child(tuple(json.loads(args.child)), args.reps)
        return
    if args.by_variant:
        # one process per variant, every shape inside it (a fresh interpreter costs more than the measurement)
        for var in args.variants:
            env = dict(os.environ)
            for kv in var.split():
                k, v = kv.split("=", 1)
                env[k] = v
            print(var or "(default)")
            env["NBENCH_SHAPES"] = json.dumps([[name] + list(SHAPES[name]) for name in args.shapes])
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "null", "--reps", str(args.reps)],
                               capture_output=True, text=True, timeout=300, env=env)
            lines = [l for l in p.stdout.splitlines() if l.startswith("NBENCH ")]
            if len(lines) != len(args.shapes):
                print("   FAILED rc=%d %s" % (p.returncode, (p.stderr or p.stdout)[-400:].replace("\n", " | ")))
            for name, line in zip(args.shapes, lines):
                r = json.loads(line[7:])
                print("   %-8s %s" % (name, "  ".join("%s %6.1f us %5d GB/s" % (k, v["us"], v["gbs"]) for k, v in r.items())))
            sys.stdout.flush()
        return
    for name in args.shapes:
        print("%s  %s" % (name, SHAPES[name]))
        for var in args.variants:
            env = dict(os.environ)
            for kv in var.split():
                k, v = kv.split("=", 1)
                env[k] = v
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", json.dumps(list(SHAPES[name])), "--reps", str(args.reps)],
                                   capture_output=True, text=True, timeout=120, env=env)
                line = [l for l in p.stdout.splitlines() if l.startswith("NBENCH ")]
                if line:
                    r = json.loads(line[0][7:])
                    print("   %-50s %s" % (var or "(default)", "  ".join("%s %6.1f us %5d GB/s" % (k, v["us"], v["gbs"]) for k, v in r.items())))
                else:
                    print("   %-50s FAILED rc=%d %s" % (var or "(default)", p.returncode, (p.stderr or p.stdout)[-300:].replace("\n", " | ")))
            except subprocess.TimeoutExpired:
                print("   %-50s TIMEOUT" % (var or "(default)"))
            sys.stdout.flush()


if __name__ == "__main__":
    main()
