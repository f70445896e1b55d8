#!/bin/bash
# Session-2 call C (1 GPU): extended probe, self-calibrating batch_d / graph tests (full log kept).
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 60 scripts/probe/umma_shift_probe.bin > $O/probe2.txt 2>&1; echo "probe rc=$?"; grep SUMMARY $O/probe2.txt; grep mismatch $O/probe2.txt | grep "base_offset=0" | head -20
timeout -s KILL 400 python -m pytest tests/test_gpu_model.py -q -k "batched or cuda_graph" -p no:cacheprovider > $O/c_tests.txt 2>&1; echo "pytest rc=$?"
grep -E "^(batch_d|replay|E  +batch_d|E  +replay|E +assert|FAILED|PASSED|[0-9]+ (passed|failed))|got \[|ref \[" $O/c_tests.txt | cut -c1-400 | head -60
tail -3 $O/c_tests.txt
