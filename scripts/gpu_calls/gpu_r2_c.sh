#!/bin/bash
# Round 2, call C: new epilogue — correctness on all tc cases, kernel timings, fidelity tests, racecheck of the non-tc kernels
mkdir -p gpurun_out
O=gpurun_out
echo "== tc cases (default)"; timeout -s KILL 500 python scripts/tc_check.py 2>&1 | cut -c1-250 | tee $O/r2c_cases.txt | grep -v " OK " 
echo "   $(grep -c ' OK ' $O/r2c_cases.txt) OK"
echo "== pair + rp3 cases"; timeout -s KILL 300 python -m pytest tests/test_gpu_conv_tc.py -q -x -k "pair or resident" -p no:cacheprovider 2>&1 | tail -3
echo "== kbench"; timeout -s KILL 400 python scripts/kbench.py --variants "" "NEMAR_TC_RP3=1" "NEMAR_TC_PAIR=1" --layers resblock d512 stn32 stn96 stn64 stn6 offset head1x1 tail1x1 down1 up2 --reps 10 --timeout 60 2>&1 | tee $O/r2c_kbench.txt
echo "== fidelity tests"; timeout -s KILL 400 python -m pytest tests/test_gpu_fidelity.py -q -s -p no:cacheprovider > $O/r2c_fidelity.txt 2>&1; echo rc=$?; grep -E "net[TRD]:|convergence|passed|failed|Error|error" $O/r2c_fidelity.txt | cut -c1-400
echo "== dist test"; timeout -s KILL 300 python -m pytest tests/test_gpu_dist.py -q -s -p no:cacheprovider > $O/r2c_dist.txt 2>&1; echo rc=$?; grep -E "bucket|DIST_CHECK|passed|failed|rank" $O/r2c_dist.txt | cut -c1-300
echo "== bench"; timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0 2>$O/r2c_bench.err | tail -1 > $O/r2c_bench.json; cut -c1-400 $O/r2c_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
r=d.get('roofline',{})
for k,v in r.get('by_kernel',{}).items(): print(k, round(v['ms'],2), v['n'], v['tflops'], {a:round(b,2) for a,b in v['top'].items()})
PY
echo "== racecheck, non-tensor-core kernels (c1, bf16 engine)"
timeout -s KILL 420 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 50 --kernel-regex-exclude kns=tc_ python scripts/sanitize_step.py c1_affine64 bf16 auto > $O/r2c_racecheck_nontc.txt 2>&1; echo rc=$?; tail -4 $O/r2c_racecheck_nontc.txt | cut -c1-300
