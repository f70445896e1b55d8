#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 0" "1 0" "1 1"; do
  set -- $cfg
  NEMAR_TC_STAGGER=$1 NEMAR_FUSED_STATS=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  python - "$cfg" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_x.json").read().strip().splitlines()[-1])
print("stagger/fused",sys.argv[1],"value",d["value"],"ms/step",d["ms_per_step"])
for k,v in d["roofline"]["by_kernel"].items(): print("  ",k, v["ms"], v["tflops"], list(v["top"].items())[:2])
PY
done
