"""profiles/dominant_kernel_ncu.json from an on-box ncu summary (scripts/ncu_summary.py) and the digest of the library the
capture ran on.  bench.py reports `roofline.traffic` from this file only while the digest matches the built library.
    python scripts/make_dominant_profile.py gpurun_out/<tag>_ncu_full_summary.json gpurun_out/<tag>_lib_digest.txt "<source note>"
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(d, key):
    v = d.get(key)
    return float(v["value"].replace(",", "")) if v else None


def to_mb(d, key):
    v = d.get(key)
    if not v:
        return None
    x = float(v["value"].replace(",", ""))
    return x * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[v["unit"]]


def to_us(d, key):
    v = d.get(key)
    x = float(v["value"].replace(",", ""))
    return x * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[v["unit"]]


def main():
    rows = json.load(open(sys.argv[1]))
    digest = open(sys.argv[2]).read().strip()
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    # the 256 -> 256 k3 ResnetBlock instance: 512 destination tiles of 128 px x 256 ch, two per CTA -> grid (256, 1, 1)
    dom = [r for r in rows if "tc_gather_kernel<256" in r.get("Kernel Name", "") and r.get("Grid Size", "").replace(" ", "") == "(256,1,1)"]
    if not dom:
        raise SystemExit("no tc_gather_kernel<256,...> launch with grid (256,1,1) in %s" % sys.argv[1])
    avg = lambda f, k: round(sum(f(r, k) for r in dom) / len(dom), 3)
    per = {"launches_captured": len(dom),
           "gpu__time_duration_us": avg(to_us, "gpu__time_duration.sum"),
           "dram__bytes_read_MB": avg(to_mb, "dram__bytes_read.sum"),
           "dram__bytes_write_MB": avg(to_mb, "dram__bytes_write.sum"),
           "algorithmic_bytes_MB": 70.3,
           "l1tex__m_xbar2l1tex_read_bytes_MB": avg(to_mb, "l1tex__m_xbar2l1tex_read_bytes.sum"),
           "lts__t_sector_hit_rate_pct": avg(num, "lts__t_sector_hit_rate.pct"),
           "sm__pipe_tensor_cycles_active_pct_of_peak_sustained_active": avg(num, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
           "sm__pipe_tensor_cycles_active_pct_of_peak_sustained_elapsed": avg(num, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
           "launch__registers_per_thread": dom[0].get("launch__registers_per_thread", {}).get("value"),
           "grid": dom[0].get("Grid Size")}
    sys.path.insert(0, ROOT)
    from nemar_b200 import build as B
    # (the capture ran on the library whose whole-source digest is `digest`; run this script before touching csrc/ so that
    #  the conv-engine digest below describes the same sources)
    out = {"source": note, "file": "profiles/dominant_kernel_ncu.json", "lib_digest": digest, "conv_tc_digest": B.conv_tc_digest(),
           "kernel": "tc_gather_kernel<256, 64, false>  (ResnetBlock conv 256->256 k3 on the reflect-padded 66x66 map, batch 16: 512 tiles of "
                     "128 pixels x 256 channels over 256 CTAs, two tiles per CTA)",
           "per_launch": per,
           "reading": "DRAM traffic (%.1f MB) against the algorithmic 70.3 MB: the bf16 output stays in the 126 MB L2.  Operand stream L2 -> SM: "
                      "%.0f MB per launch (the bound: ~6300 B/clk chip-wide).  Tensor pipe active %.0f %% of the kernel's active cycles."
                      % (per["dram__bytes_read_MB"] + per["dram__bytes_write_MB"], per["l1tex__m_xbar2l1tex_read_bytes_MB"],
                         per["sm__pipe_tensor_cycles_active_pct_of_peak_sustained_active"]),
           "other_kernels_same_session": [r for r in rows if r not in dom]}
    json.dump(out, open(os.path.join(ROOT, "profiles", "dominant_kernel_ncu.json"), "w"), indent=1)
    print(json.dumps(per, indent=1))


if __name__ == "__main__":
    main()
