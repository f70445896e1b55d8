"""Is one optimize_parameters step (lr 0) a reproducible function of its input?  Evaluates a fixed input several times,
with other inputs in between, and prints the pairwise relative differences of the D / T+R weight-gradient buckets."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests import helpers as H  # noqa: E402
from tests import test_gpu_model as T  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c1_affine64"
    precision = sys.argv[2] if len(sys.argv) > 2 else "fp32"
    engine = sys.argv[3] if len(sys.argv) > 3 else "generic"
    flags = ["--lr", "0"] + sys.argv[4:]
    model, cfg, states, _ = H.build_case(name, precision=precision, conv_engine=engine, more_flags=flags)
    X = [T._case_inputs(name, k) for k in range(3)]
    seq = [0, 0, 0, 1, 1, 0, 0, 2, 0, 2, 1]
    recs = [T._step_record(model, *X[k]) for k in seq]
    print("case %s %s %s %s; input sequence %s" % (name, precision, engine, flags, seq))
    for k in range(3):
        idx = [i for i, s in enumerate(seq) if s == k]
        print(" input %d evaluated at steps %s" % (k, [i + 1 for i in idx]))
        for what, col in (("D weights", 1), ("T+R weights", 2)):
            rows = []
            for a in idx:
                rows.append(" ".join("%8.1e" % T._rel(recs[a][col], recs[b][col]) for b in idx))
            print("  %s pairwise rel diff:\n    %s" % (what, "\n    ".join(rows)))
        l = np.array([recs[i][0] for i in idx])
        print("  losses max rel spread: %.2e" % float(np.max((l.max(0) - l.min(0)) / (np.abs(l.mean(0)) + 1e-3))))


if __name__ == "__main__":
    main()
