#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=line -p no:cacheprovider 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
timeout 600 python bench.py --steps 5 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 2 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_a.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms/step",d["ms_per_step"],"e2e",d["e2e"]["value"],"launches",d["gpu_launches"])
for k,v in d["roofline"]["by_kernel"].items(): print("  ",k, v["ms"], v["n"], v["tflops"], list(v["top"].items())[:5])
d=json.loads(open("gpurun_out/bench_b.json").read().strip().splitlines()[-1])
print("ALL-CALL TIMING (5 steps): ms/step", d["ms_per_step"])
for k,v in d["roofline"]["by_kernel"].items(): print("  %-28s %8.2f ms/step  n/step %5.1f"%(k, v["ms"]/5, v["n"]/5))
PY
