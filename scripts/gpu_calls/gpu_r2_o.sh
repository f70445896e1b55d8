#!/bin/bash
# Round 2: cp.async-ring InstanceNorm passes — parity of the passes, then A/B of ring depths against the register-staged kernels
mkdir -p gpurun_out
O=gpurun_out
echo "== norm tests"; timeout -s KILL 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "norm_act or images" > $O/r2o_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|^FAILED|^ERROR" $O/r2o_tests.txt | cut -c1-300
echo "== nbench"
timeout -s KILL 900 python scripts/nbench.py --by_variant --reps 10 --shapes res256 res256r up128 head64 stn32 --variants "NEMAR_LEAN_PIPE=0" "NEMAR_LEAN_PIPE=3" "NEMAR_LEAN_PIPE=4" "NEMAR_LEAN_PIPE=6" "NEMAR_LEAN_PIPE=8" "NEMAR_LEAN_PIPE=4 NEMAR_LEAN_WAVES=2" > $O/r2o_nbench.txt 2>&1
cat $O/r2o_nbench.txt | cut -c1-260
