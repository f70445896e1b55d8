#!/bin/bash
# Round 2, evidence call: the default bench line, the reference arm, the ncu launch list and `ncu --set full` captures
# (reports are summarised on the box: gpurun copies back at most 64 MiB)
mkdir -p gpurun_out
O=gpurun_out
if [ "$1" != "ncu_only" ]; then
echo "== bench (default flags)"; timeout -s KILL 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>$O/r2z_bench.err | tail -1 > $O/r2z_bench.json; cut -c1-300 $O/r2z_bench.json; tail -2 $O/r2z_bench.err
echo "== bench --impl reference"; timeout -s KILL 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 2 2>$O/r2z_ref.err | tail -1 > $O/r2z_ref.json; cut -c1-300 $O/r2z_ref.json
fi
B="python bench.py --profile --cuda_graph 0 --steps 1 --warmup 1 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 --torch_gpu_reference 0 --stream_overlap 0"
echo "== ncu launch list"; NEMAR_WGRAD_STREAM=0 timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2z_launches.csv $B > $O/r2z_ncu_list.log 2>&1; echo rc=$?
echo "== ncu --set full: gather (256-wide), wgrad pairs, norm passes, grid_sample"
NEMAR_WGRAD_STREAM=0 timeout -s KILL 300 ncu --set full --clock-control none -k regex:tc_gather_kernel -s 30 -c 3 -o /tmp/r2z_gather -f $B > $O/r2z_ncu_gather.log 2>&1; echo rc=$?
NEMAR_WGRAD_STREAM=0 timeout -s KILL 300 ncu --set full --clock-control none -k regex:tc_wgrad_pair -s 4 -c 2 -o /tmp/r2z_wgrad -f $B > $O/r2z_ncu_wgrad.log 2>&1; echo rc=$?
NEMAR_WGRAD_STREAM=0 timeout -s KILL 300 ncu --set full --clock-control none -k regex:"bwd_apply_kernel|reduce_kernel|fwd_kernel|tc_rp3" -s 60 -c 10 -o /tmp/r2z_norm -f $B > $O/r2z_ncu_norm.log 2>&1; echo rc=$?
timeout -s KILL 300 ncu --set full --clock-control none -k regex:grid_sample -s 6 -c 4 -o /tmp/r2z_gs -f python scripts/gs_bench.py > $O/r2z_ncu_gs.log 2>&1; echo rc=$?
python scripts/ncu_summary.py $O/r2z_ncu_full_summary.json /tmp/r2z_gather.ncu-rep /tmp/r2z_wgrad.ncu-rep /tmp/r2z_norm.ncu-rep /tmp/r2z_gs.ncu-rep | cut -c1-250
cat nemar_b200/build/stamp > $O/r2z_lib_digest.txt
du -sh $O
