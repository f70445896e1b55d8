"""grid_sample micro-benchmark alone (the `grid_sample` object of bench.py's line), for ncu captures."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import bench  # noqa: E402

print(json.dumps(bench.grid_sample_bench(torch.device("cuda", 0), bench.load_peaks(), iters=5)))
