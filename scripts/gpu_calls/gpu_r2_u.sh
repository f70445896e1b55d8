#!/bin/bash
# Round 2: norm passes with register-resident coefficients — full GPU suite, nbench, default bench line; fp32 gradient probe 3
mkdir -p gpurun_out
O=gpurun_out
cat nemar_b200/build/stamp > $O/r2u_lib_digest.txt
echo "== tests"; timeout -s KILL 1200 python -m pytest tests/ -m gpu -q -p no:cacheprovider > $O/r2u_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|^FAILED|^ERROR" $O/r2u_tests.txt | cut -c1-300
echo "== nbench"; timeout -s KILL 200 python scripts/nbench.py --by_variant --reps 10 --shapes res256 res256r up128 head64 stn32 --variants "" > $O/r2u_nbench.txt 2>&1; cut -c1-220 $O/r2u_nbench.txt
echo "== probe3"; timeout -s KILL 400 python tests/probes/fp32_grad_error_probe3.py > $O/r2u_probe3.txt 2>&1; grep PROBE3 $O/r2u_probe3.txt
echo "== bench (default flags)"; timeout -s KILL 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>$O/r2u_bench.err | tail -1 > $O/r2u_bench.json; cut -c1-300 $O/r2u_bench.json; tail -2 $O/r2u_bench.err
