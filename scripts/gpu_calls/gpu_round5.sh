#!/bin/bash
mkdir -p gpurun_out
timeout 300 python - <<'PY'
import sys, json, torch
sys.path.insert(0, ".")
import bench
peaks = bench.load_peaks()
dev = torch.device("cuda", 0)
for size, n in ((1024, 4), (256, 16), (512, 8)):
    r = bench.grid_sample_bench(dev, peaks, size=size, n=n)
    print("grid_sample", size, n, json.dumps({k: r[k] for k in ("fwd", "bwd")}))
PY
B="python bench.py --profile --steps 1 --warmup 0 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:norm_act_bwd_apply|plane_reduce_kernel" -s 95 -c 10 -o gpurun_out/prof_bwd -f $B > gpurun_out/ncu_bwd.log 2>&1
echo "rc=$?"
