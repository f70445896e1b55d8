#!/bin/bash
# Round 2: (1) where the fp32 engine's 2e-3 gradient error on netT comes from; (2) `ncu --set full` of the ring norm passes on
# the 256-channel maps (the capture of gpu_r2_final4.sh landed on the STN's tiny maps)
mkdir -p gpurun_out
O=gpurun_out
cat nemar_b200/build/stamp > $O/r2s_lib_digest.txt
echo "== fp32 gradient error probe"; timeout -s KILL 600 python tests/probes/fp32_grad_error_probe.py > $O/r2s_fp32_probe.txt 2>&1; echo rc=$?; grep -E "^==|bucket" $O/r2s_fp32_probe.txt
B="python bench.py --profile --cuda_graph 0 --steps 1 --warmup 1 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 --torch_gpu_reference 0 --stream_overlap 0"
echo "== ncu --set full (norm passes)"
NEMAR_WGRAD_STREAM=0 timeout -s KILL 400 ncu --set full --clock-control none -k regex:"bwd_apply_pipe_kernel|reduce_pipe_kernel|fwd_pipe_kernel" -s 100 -c 70 -o /tmp/r2s_norm -f $B > $O/r2s_ncu_norm.log 2>&1; echo rc=$?
python scripts/ncu_summary.py $O/r2s_ncu_norm_summary.json /tmp/r2s_norm.ncu-rep | grep -v "(1, 16, 1)" | cut -c1-250 | head -60
