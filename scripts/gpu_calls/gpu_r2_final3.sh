#!/bin/bash
# Round 2: full GPU test suite + the default bench line of the final build (traffic from the matching ncu capture)
mkdir -p gpurun_out
O=gpurun_out
echo "== tests"; timeout -s KILL 1500 python -m pytest tests/ -m gpu -q -p no:cacheprovider > $O/r2x_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|^FAILED|^ERROR" $O/r2x_tests.txt | cut -c1-300
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench (default flags)"; timeout -s KILL 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>$O/r2x_bench.err | tail -1 > $O/r2x_bench.json; cut -c1-300 $O/r2x_bench.json; tail -2 $O/r2x_bench.err
cat nemar_b200/build/stamp > $O/r2x_lib_digest.txt
