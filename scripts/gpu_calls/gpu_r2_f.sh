#!/bin/bash
# Round 2, call F: defaults WIDE + RP3, multi-pack, arena, dropout-in-graph, prefetcher fix
mkdir -p gpurun_out
O=gpurun_out
echo "== conv engine tests (defaults + variants)"; timeout -s KILL 900 python -m pytest tests/test_gpu_conv_tc.py -q -x -p no:cacheprovider 2>&1 | tail -4 | cut -c1-300
echo "== batch-4 debug"; timeout -s KILL 300 python scripts/debug_batch4.py 2>&1 | grep "batch" | tee $O/r2f_batch4.txt
echo "== tests"; timeout -s KILL 1200 python -m pytest tests/test_gpu_fidelity.py tests/test_gpu_dist.py tests/test_gpu_next_rows.py tests/test_gpu_zz_golden_sizes.py tests/test_gpu_model.py tests/test_gpu_ops.py -q -s -p no:cacheprovider > $O/r2f_tests.txt 2>&1; echo rc=$?
grep -E "passed|failed|bucket TR|bucket D|DIST_CHECK|^FAILED|^ERROR" $O/r2f_tests.txt | cut -c1-300
echo "== bench"; timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no_cpu_baseline 2>$O/r2f_bench.err | tail -1 > $O/r2f_bench.json; cut -c1-300 $O/r2f_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
r=d.get('roofline',{})
for k,v in r.get('by_kernel',{}).items(): print(k, round(v['ms'],2), v['n'], v['tflops'], {a:round(b,2) for a,b in v['top'].items()})
print({k:d.get(k) for k in ('e2e','torch_gpu_reference','gpu_launches')})
PY
tail -3 $O/r2f_bench.err
echo "== launch list (ncu, eager step)"
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2f_launches.csv python bench.py --profile --cuda_graph 0 --steps 1 --warmup 1 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 --torch_gpu_reference 0 > $O/r2f_ncu_list.log 2>&1; echo rc=$?
echo "== ncu: InstanceNorm passes"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"reduce_kernel|bwd_apply_kernel|fwd_kernel" -s 40 -c 24 -o $O/r2f_nlean -f \
    python bench.py --profile --cuda_graph 0 --steps 1 --warmup 1 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 --torch_gpu_reference 0 > $O/r2f_ncu_nlean.log 2>&1; echo rc=$?
