#!/bin/bash
# Round 2, final evidence call (after the last kernel change): tests, default bench line, reference arm, launch list,
# `ncu --set full` of the dominant conv instance and the 256-channel InstanceNorm passes (summarised on the box)
mkdir -p gpurun_out
O=gpurun_out
echo "== tests"; timeout -s KILL 1500 python -m pytest tests/ -m gpu -q -p no:cacheprovider > $O/r2y_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|^FAILED|^ERROR" $O/r2y_tests.txt | cut -c1-300
echo "== bench (default flags)"; timeout -s KILL 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>$O/r2y_bench.err | tail -1 > $O/r2y_bench.json; cut -c1-300 $O/r2y_bench.json; tail -2 $O/r2y_bench.err
echo "== bench --impl reference"; timeout -s KILL 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 2 2>$O/r2y_ref.err | tail -1 > $O/r2y_ref.json; cut -c1-200 $O/r2y_ref.json
B="python bench.py --profile --cuda_graph 0 --steps 1 --warmup 1 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 --torch_gpu_reference 0 --stream_overlap 0"
echo "== ncu launch list"; NEMAR_WGRAD_STREAM=0 timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2y_launches.csv $B > $O/r2y_ncu_list.log 2>&1; echo rc=$?
echo "== ncu --set full"
NEMAR_WGRAD_STREAM=0 timeout -s KILL 300 ncu --set full --clock-control none -k regex:tc_gather_kernel -s 5 -c 3 -o /tmp/r2y_gather -f $B > $O/r2y_ncu_gather.log 2>&1; echo rc=$?
NEMAR_WGRAD_STREAM=0 timeout -s KILL 300 ncu --set full --clock-control none -k regex:"bwd_apply_kernel|reduce_kernel|fwd_kernel" -s 170 -c 8 -o /tmp/r2y_norm -f $B > $O/r2y_ncu_norm.log 2>&1; echo rc=$?
python scripts/ncu_summary.py $O/r2y_ncu_full_summary.json /tmp/r2y_gather.ncu-rep /tmp/r2y_norm.ncu-rep | cut -c1-250
cat nemar_b200/build/stamp > $O/r2y_lib_digest.txt
du -sh $O
