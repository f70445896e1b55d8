#!/bin/bash
# Round 2, last call: the final build (register-resident norm coefficients) — full GPU suite, smoke, launch list; fp32 probe 4
mkdir -p gpurun_out
O=gpurun_out
cat nemar_b200/build/stamp > $O/r2v_lib_digest.txt
echo "== tests"; timeout -s KILL 900 python -m pytest tests/ -m gpu -q -p no:cacheprovider > $O/r2v_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|^FAILED|^ERROR" $O/r2v_tests.txt | cut -c1-300
echo "== smoke"; timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
B="python bench.py --profile --cuda_graph 0 --steps 1 --warmup 1 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 --torch_gpu_reference 0 --stream_overlap 0"
echo "== ncu launch list"; NEMAR_WGRAD_STREAM=0 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2v_launches.csv $B > $O/r2v_ncu_list.log 2>&1; echo rc=$?
echo "== probe4"
timeout -s KILL 200 python tests/probes/fp32_grad_error_probe4.py engine_first 2>/dev/null | grep PROBE4
timeout -s KILL 200 python tests/probes/fp32_grad_error_probe4.py oracle_first 2>/dev/null | grep PROBE4
timeout -s KILL 200 python tests/probes/fp32_grad_error_probe4.py engine_first 2>/dev/null | grep PROBE4
