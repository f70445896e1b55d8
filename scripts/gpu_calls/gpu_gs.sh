#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "grid_sample or smoothness or affine" --tb=short -p no:cacheprovider 2>&1 | tail -3
for v in 0 1; do
NEMAR_GS_VARIANT=$v timeout 300 python - <<'PY'
import sys, json, torch, os
sys.path.insert(0, ".")
import bench
peaks = bench.load_peaks()
dev = torch.device("cuda", 0)
for size, n in ((1024, 4), (256, 16)):
    r = bench.grid_sample_bench(dev, peaks, size=size, n=n)
    print("variant", os.environ["NEMAR_GS_VARIANT"], "grid_sample", size, n, json.dumps({k: r[k] for k in ("fwd", "bwd")}))
PY
done
