#!/bin/bash
# Round-end evidence run (1 GPU): tests, smoke, full bench line, reference arm, one-step ncu launch list.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=line -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-400 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
B="python bench.py --profile --steps 1 --warmup 1 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/ncu_list2.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_final.csv
