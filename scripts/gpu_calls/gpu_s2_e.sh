#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_model.py -q -k "cuda_graph" -p no:cacheprovider -s > $O/e_tests.txt 2>&1; echo "pytest rc=$?"
grep -E "^(batch_d|replay|first eager)|worst D|^FAILED|passed|failed" $O/e_tests.txt | cut -c1-600 | head -40
