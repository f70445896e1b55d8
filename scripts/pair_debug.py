"""CTA-pair (cta_group::2) kernels against the single-CTA kernels on one geometry, same process, by region.
usage: python scripts/pair_debug.py [case index, default 2]   (run under `timeout`: a protocol bug traps or hangs)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from nemar_b200.engine import lib as L  # noqa: E402
from tests import tc_cases as t  # noqa: E402


def set_pair(v):
    return L.lib().nemar_conv2d_set_option(C.c_char_p(b"pair"), C.c_int(v))


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-20))


def by_region(name, a, b):
    """a, b: [n, h, w, c] (pair, single)"""
    print("  %s total rel err %.3e" % (name, rel(a, b)))
    if rel(a, b) < 1e-3:
        return
    c = a.shape[3]
    for half in range(0, c, 128):
        print("    channels %d..%d: %.3e" % (half, min(half + 128, c) - 1, rel(a[..., half:half + 128], b[..., half:half + 128])))
    for n in range(a.shape[0]):
        bands = ["%.1e" % rel(a[n, y:y + 4], b[n, y:y + 4]) for y in range(0, a.shape[1], 4)]
        print("    sample %d, 4-row bands: %s" % (n, " ".join(bands)))


def main():
    idx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    case = t.CASES[idx]
    print("case %d: %s" % (idx, case[-1]), flush=True)
    x, wt, b = t._mk(case)
    set_pair(0)
    y0, s0, dx0, dw0, db0, dy, _, _ = t._engine(case, x, wt, b, None, True)
    print("single-CTA run done", flush=True)
    set_pair(int(os.environ.get("PAIR_MODE", "1")))
    y1, s1, dx1, dw1, db1, _, _, _ = t._engine(case, x, wt, b, dy, True)
    print("pair run done", flush=True)
    by_region("fprop y", y1, y0)
    by_region("dgrad dx", dx1, dx0)
    print("  wgrad total rel err %.3e" % rel(dw1, dw0))
    if rel(dw1, dw0) > 1e-3:
        co = dw1.shape[0]
        for h in range(0, co, 128):
            print("    dw rows (out channels) %d..%d: %.3e" % (h, min(h + 128, co) - 1, rel(dw1[h:h + 128], dw0[h:h + 128])))
        ci = dw1.shape[1]
        for h in range(0, ci, 128):
            print("    dw cols (in channels) %d..%d: %.3e" % (h, min(h + 128, ci) - 1, rel(dw1[:, h:h + 128], dw0[:, h:h + 128])))
        for tap in range(dw1.shape[2] * dw1.shape[3]):
            a, bb = dw1.reshape(co, ci, -1)[..., tap], dw0.reshape(co, ci, -1)[..., tap]
            print("    tap %d: %.3e" % (tap, rel(a, bb)))


if __name__ == "__main__":
    main()
