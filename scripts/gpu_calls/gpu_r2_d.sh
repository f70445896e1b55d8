#!/bin/bash
# Round 2, call D: MMA issue-rate probe, the new tests, ncu of the InstanceNorm passes, BN=256 variant
mkdir -p gpurun_out
O=gpurun_out
echo "== UMMA rate probe"; timeout -s KILL 120 scripts/probe/umma_rate_probe.bin > $O/r2d_rate_probe.txt 2>&1; echo rc=$?; cat $O/r2d_rate_probe.txt
echo "== tests"; timeout -s KILL 900 python -m pytest tests/test_gpu_fidelity.py tests/test_gpu_dist.py tests/test_gpu_next_rows.py tests/test_gpu_zz_golden_sizes.py tests/test_gpu_ops.py -q -s -p no:cacheprovider -k "not two_images" > $O/r2d_tests.txt 2>&1; echo rc=$?
grep -E "net[TRD]:|convergence|passed|failed|bucket|DIST_CHECK|^FAILED|^ERROR|Error" $O/r2d_tests.txt | cut -c1-330
echo "== kbench BN=256"; timeout -s KILL 200 python scripts/kbench.py --variants "" "NEMAR_TC_WIDE=1" "NEMAR_TC_PAIR=1" "NEMAR_TC_PAIR=1 NEMAR_TC_PAIR_STAGES=4" "NEMAR_TC_PAIR=1 NEMAR_TC_PAIR_STAGES=6 NEMAR_TC_PAIR_TPC=2" --layers resblock d512 --reps 10 --timeout 60 2>&1 | tee $O/r2d_kbench.txt
echo "== ncu: InstanceNorm passes"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"nlean" -s 24 -c 28 -o $O/r2d_nlean -f \
    python bench.py --profile --cuda_graph 0 --steps 1 --warmup 1 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 --torch_gpu_reference 0 > $O/r2d_ncu_nlean.log 2>&1; echo rc=$?
