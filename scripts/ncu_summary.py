"""Summarise .ncu-rep files (ncu --set full captures) into one small CSV / JSON: the metrics DESIGN.md and bench.py quote.
    python scripts/ncu_summary.py out.json a.ncu-rep b.ncu-rep ...
"""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "launch__waves_per_multiprocessor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    rows_out = []
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = {"file": rep.split("/")[-1]}
            for i, h in enumerate(hdr):
                if h in ("Kernel Name", "Grid Size", "Block Size"):
                    d[h] = r[i]
                elif h in WANT:
                    d[h] = {"value": r[i], "unit": units[i]}
            rows_out.append(d)
    json.dump(rows_out, open(out_path, "w"), indent=1)
    for d in rows_out:
        g = lambda k: (d.get(k) or {}).get("value", "?") + " " + (d.get(k) or {}).get("unit", "")
        print("%-60s %-14s %s | dram r %s w %s | tensor %s | L2 hit %s | warps %s" % (
            d.get("Kernel Name", "")[:60], d.get("Grid Size", ""), g("gpu__time_duration.sum"), g("dram__bytes_read.sum"), g("dram__bytes_write.sum"),
            g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), g("lts__t_sector_hit_rate.pct"),
            g("sm__warps_active.avg.pct_of_peak_sustained_active")))


if __name__ == "__main__":
    main()
