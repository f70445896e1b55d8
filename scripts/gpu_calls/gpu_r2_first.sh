#!/bin/bash
# Round 2, first GPU call (1 GPU, ~4 min): everything that was written after round 1's GPU budget was spent.
#   gpurun --timeout 420 -- 'bash scripts/gpu_r2_first.sh'
mkdir -p gpurun_out
O=gpurun_out
echo "== UMMA window probe (K-major all swizzles + MN-major windows with overlapping chunks)"
[ -x scripts/probe/umma_shift_probe.bin ] || nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I nemar_b200/csrc -I include -o scripts/probe/umma_shift_probe.bin scripts/probe/umma_shift_probe.cu
timeout -s KILL 60 scripts/probe/umma_shift_probe.bin > $O/r2_probe.txt 2>&1; echo "rc=$?"; grep SUMMARY $O/r2_probe.txt
echo "== resident-patch kernel: cases"; NEMAR_TC_RP3=1 timeout -s KILL 240 python scripts/tc_check.py 17 16 18 19 0 12 21 13 14 11 22 2>&1 | cut -c1-300 | tee $O/r2_rp3_cases.txt
echo "== resident-patch kernel: timing"; timeout -s KILL 200 python scripts/kbench.py --variants "" "NEMAR_TC_RP3=1" "NEMAR_TC_RP3=1 NEMAR_TC_RP3_STAGES=3" --layers stn32 stn96 stn64 stn6 offset --reps 10 --timeout 60 2>&1 | tee $O/r2_kbench_rp3.txt
echo "== wave-tail split (NEMAR_TC_TAIL=1): correctness on the 256-channel cases, timing on the ResnetBlock conv"
NEMAR_TC_TAIL=1 timeout -s KILL 200 python scripts/tc_check.py 2 9 10 24 25 26 2>&1 | cut -c1-300 | tee $O/r2_tail_cases.txt
timeout -s KILL 200 python scripts/kbench.py --variants "" "NEMAR_TC_TAIL=1" "NEMAR_TC_TPC=1" --layers resblock d512 --reps 10 --timeout 60 2>&1 | tee $O/r2_kbench_tail.txt
echo "== golden-size engine tests"; NEMAR_TEST_UNVALIDATED=1 timeout -s KILL 300 python -m pytest tests/test_gpu_zz_golden_sizes.py -q -p no:cacheprovider > $O/r2_golden_sizes.txt 2>&1; echo rc=$?; tail -15 $O/r2_golden_sizes.txt | cut -c1-300
echo "== bench (default)"; timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 2>$O/r2_bench.err | tail -1 > $O/r2_bench.json; cut -c1-600 $O/r2_bench.json
echo "== ncu --set full: the two dominant kernels of the final build (2 launches each)"
for K in tc_gather_kernel tc_wgrad_pair_kernel; do
  timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 2 -o $O/r2_$K -f \
    python bench.py --profile --cuda_graph 0 --steps 1 --warmup 1 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 > $O/r2_ncu_$K.log 2>&1
  echo "$K rc=$?"
done
