#!/bin/bash
# Round 2, call B (1 GPU): bf16 fidelity tests, ncu of the pair / resident-patch kernels, compute-sanitizer
mkdir -p gpurun_out
O=gpurun_out
echo "== fidelity tests"; timeout -s KILL 400 python -m pytest tests/test_gpu_fidelity.py -x -q -s -p no:cacheprovider > $O/r2b_fidelity.txt 2>&1; echo rc=$?; grep -E "net[TRD]:|convergence|passed|failed|Error|error" $O/r2b_fidelity.txt | cut -c1-400
RES='[256,256,3,1,1,false,1,16,64,64,0,false]'
STN='[32,32,3,1,1,false,1,16,256,256,0,false]'
echo "== ncu: pair gather"
NEMAR_TC_PAIR=1 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:tc_gather_pair -s 4 -c 2 -o $O/r2b_pair_gather -f python scripts/kbench.py --child "$RES" --reps 2 > $O/r2b_ncu_pair.log 2>&1; echo rc=$?
echo "== ncu: rp3 stn32"
NEMAR_TC_RP3=1 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:tc_rp3 -s 4 -c 2 -o $O/r2b_rp3_stn32 -f python scripts/kbench.py --child "$STN" --reps 2 > $O/r2b_ncu_rp3.log 2>&1; echo rc=$?
echo "== ncu: default stn32"
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:tc_gather_kernel -s 4 -c 2 -o $O/r2b_gather_stn32 -f python scripts/kbench.py --child "$STN" --reps 2 > $O/r2b_ncu_stn32.log 2>&1; echo rc=$?
echo "== racecheck (c1, bf16 tcgen05 engine)"
timeout -s KILL 420 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 30 python scripts/sanitize_step.py c1_affine64 bf16 auto > $O/r2b_racecheck_bf16.txt 2>&1; echo rc=$?; tail -5 $O/r2b_racecheck_bf16.txt | cut -c1-300
echo "== racecheck (c1, fp32 generic engine)"
timeout -s KILL 300 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 30 python scripts/sanitize_step.py c1_affine64 fp32 generic > $O/r2b_racecheck_fp32.txt 2>&1; echo rc=$?; tail -5 $O/r2b_racecheck_fp32.txt | cut -c1-300
echo "== memcheck (c1, bf16)"
timeout -s KILL 300 compute-sanitizer --tool memcheck --print-limit 30 python scripts/sanitize_step.py c1_affine64 bf16 auto > $O/r2b_memcheck_bf16.txt 2>&1; echo rc=$?; tail -5 $O/r2b_memcheck_bf16.txt | cut -c1-300
