#!/bin/bash
# Round 2, call E: converged-warp / elect.sync issue of TMA + tcgen05 (no waterfall loops)
mkdir -p gpurun_out
O=gpurun_out
echo "== UMMA rate probe"; timeout -s KILL 120 scripts/probe/umma_rate_probe.bin > $O/r2e_rate_probe.txt 2>&1; echo rc=$?; cat $O/r2e_rate_probe.txt
echo "== tc cases (default)"; timeout -s KILL 500 python scripts/tc_check.py 2>&1 | cut -c1-250 | tee $O/r2e_cases.txt | grep -v " OK "
echo "   $(grep -c ' OK ' $O/r2e_cases.txt) OK"
echo "== pair + rp3 cases"; timeout -s KILL 300 python -m pytest tests/test_gpu_conv_tc.py -q -x -k "pair or resident" -p no:cacheprovider 2>&1 | tail -3
echo "== kbench"; timeout -s KILL 500 python scripts/kbench.py --variants "" "NEMAR_TC_RP3=1" "NEMAR_TC_PAIR=1" "NEMAR_TC_WIDE=1" --layers resblock d512 stn32 stn96 stn64 stn6 offset head1x1 tail1x1 down1 up2 d128 --reps 10 --timeout 60 2>&1 | tee $O/r2e_kbench.txt
echo "== tests"; timeout -s KILL 900 python -m pytest tests/test_gpu_fidelity.py tests/test_gpu_dist.py tests/test_gpu_next_rows.py tests/test_gpu_zz_golden_sizes.py tests/test_gpu_model.py -q -s -p no:cacheprovider > $O/r2e_tests.txt 2>&1; echo rc=$?
grep -E "net[TRD]:|passed|failed|bucket|DIST_CHECK|^FAILED|^ERROR" $O/r2e_tests.txt | cut -c1-330
echo "== bench"; timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline 2>$O/r2e_bench.err | tail -1 > $O/r2e_bench.json; cut -c1-300 $O/r2e_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
r=d.get('roofline',{})
for k,v in r.get('by_kernel',{}).items(): print(k, round(v['ms'],2), v['n'], v['tflops'], {a:round(b,2) for a,b in v['top'].items()})
print({k:d.get(k) for k in ('e2e','torch_gpu_reference','grid_sample')})
PY
tail -5 $O/r2e_bench.err
