#!/bin/bash
# Session-2 call D (1 GPU): the whole GPU suite with the new defaults, smoke, default bench line.
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x --durations=12 -p no:cacheprovider > $O/d_tests.txt 2>&1; echo "pytest rc=$?"
grep -E "^(batch_d|replay)|worst D|^FAILED|passed|failed|^E  " $O/d_tests.txt | cut -c1-500 | head -40
grep -A14 "slowest" $O/d_tests.txt | cut -c1-160
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > $O/d_bench.json 2> $O/d_bench.err; echo "bench rc=$?"; cut -c1-1500 $O/d_bench.json
