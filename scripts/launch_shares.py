"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: share of the summed kernel time,
ms per step, launches per step.
    python scripts/launch_shares.py launches.csv STEPS "header note" > profiles/<name>.csv
"""
import collections
import csv
import sys


def main():
    path, steps = sys.argv[1], int(sys.argv[2])
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0.0, 0])
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ui]]      # -> ms
        agg[r[ki]][0] += v
        agg[r[ki]][1] += 1
    tot = sum(a[0] for a in agg.values())
    n = sum(a[1] for a in agg.values())
    for line in note.split("\\n"):
        if line:
            print("# " + line)
    print("# %d identical steps captured; %d launches and %.2f ms of summed kernel time per step" % (steps, n // steps, tot / steps))
    print("share_pct,ms_per_step,launches_per_step,avg_us,kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%.2f,%.3f,%.1f,%.1f,%s" % (100 * a[0] / tot, a[0] / steps, a[1] / steps, 1e3 * a[0] / a[1], k.replace(",", ";")))


if __name__ == "__main__":
    main()
