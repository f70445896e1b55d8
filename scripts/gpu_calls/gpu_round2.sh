#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x 2>&1 | tail -8
for wide in 0 1; do
  NEMAR_TC_WIDE=$wide timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 > gpurun_out/bench_w$wide.json 2> gpurun_out/bench_w$wide.err
  python - $wide <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_w%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
r=d["roofline"]; bk=r["by_kernel"]
print("WIDE",sys.argv[1],"value",d["value"],"ms/step",d["ms_per_step"],"launches",d["gpu_launches"])
for k,v in bk.items(): print("  ",k, v["ms"], v["n"], v["tflops"], list(v["top"].items())[:4])
PY
done
