#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 200 python scripts/determinism_probe.py c1_affine64 fp32 generic > $O/f_det_fp32.txt 2>&1; echo rc=$?; grep -v "^initialize\|^model" $O/f_det_fp32.txt | cut -c1-200
timeout -s KILL 200 python scripts/determinism_probe.py c1_affine64 fp32 generic --batch_d 0 > $O/f_det_fp32_b0.txt 2>&1; echo rc=$?; grep -v "^initialize\|^model" $O/f_det_fp32_b0.txt | cut -c1-200 | head -24
timeout -s KILL 600 python -m pytest tests/test_gpu_model.py -q -k "cuda_graph or batched" -p no:cacheprovider -s > $O/f_tests.txt 2>&1; echo "pytest rc=$?"
grep -E "^(batch_d|replay|eager-vs)|^FAILED|passed|failed" $O/f_tests.txt | cut -c1-300 | head -40
