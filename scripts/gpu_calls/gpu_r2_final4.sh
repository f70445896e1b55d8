#!/bin/bash
# Round 2, evidence call for the final build (cp.async-ring norm passes, one wgrad-pair CTA per SM): full GPU suite, the
# default bench line, the reference arm, C4 / C5 single-GPU lines, launch list, `ncu --set full` of the dominant conv
# instance and of the norm passes (summarised on the box; .ncu-rep files stay in /tmp)
mkdir -p gpurun_out
O=gpurun_out
cat nemar_b200/build/stamp > $O/r2w_lib_digest.txt
echo "== tests"; timeout -s KILL 1500 python -m pytest tests/ -m gpu -q -p no:cacheprovider > $O/r2w_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|^FAILED|^ERROR" $O/r2w_tests.txt | cut -c1-300
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default flags)"; timeout -s KILL 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>$O/r2w_bench.err | tail -1 > $O/r2w_bench.json; cut -c1-300 $O/r2w_bench.json; tail -2 $O/r2w_bench.err
echo "== bench --impl reference"; timeout -s KILL 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 2 2>$O/r2w_ref.err | tail -1 > $O/r2w_ref.json; cut -c1-200 $O/r2w_ref.json
L="--no_cpu_baseline --grid_sample_bench 0 --torch_gpu_reference 0"
echo "== bench C4"; timeout -s KILL 400 python bench.py --workload C4 --steps 10 --warmup 3 $L 2>/dev/null | tail -1 > $O/r2w_bench_c4.json; cut -c1-200 $O/r2w_bench_c4.json
echo "== bench C5"; timeout -s KILL 400 python bench.py --workload C5 --steps 10 --warmup 3 $L 2>/dev/null | tail -1 > $O/r2w_bench_c5.json; cut -c1-200 $O/r2w_bench_c5.json
B="python bench.py --profile --cuda_graph 0 --steps 1 --warmup 1 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 --torch_gpu_reference 0 --stream_overlap 0"
echo "== ncu launch list"; NEMAR_WGRAD_STREAM=0 timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2w_launches.csv $B > $O/r2w_ncu_list.log 2>&1; echo rc=$?
echo "== ncu --set full"
NEMAR_WGRAD_STREAM=0 timeout -s KILL 300 ncu --set full --clock-control none -k regex:tc_gather_kernel -s 5 -c 3 -o /tmp/r2w_gather -f $B > $O/r2w_ncu_gather.log 2>&1; echo rc=$?
NEMAR_WGRAD_STREAM=0 timeout -s KILL 300 ncu --set full --clock-control none -k regex:"bwd_apply_pipe_kernel|reduce_pipe_kernel|fwd_pipe_kernel|tc_wgrad_pair_kernel" -s 200 -c 10 -o /tmp/r2w_norm -f $B > $O/r2w_ncu_norm.log 2>&1; echo rc=$?
python scripts/ncu_summary.py $O/r2w_ncu_full_summary.json /tmp/r2w_gather.ncu-rep /tmp/r2w_norm.ncu-rep | cut -c1-250
du -sh $O
