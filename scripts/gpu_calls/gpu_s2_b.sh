#!/bin/bash
# Session-2 call B (1 GPU): graph / batch_d tests, per-op breakdown, candidate-default bench, ncu launch list.
mkdir -p gpurun_out
O=gpurun_out
echo "== tests"; timeout -s KILL 600 python -m pytest tests/test_gpu_model.py -q -k "batched or cuda_graph" -p no:cacheprovider 2>&1 | grep -v "^$" | cut -c1-600 | tail -40
B="python bench.py --steps 10 --warmup 3 --no_cpu_baseline --grid_sample_bench 0"
echo "== per-op breakdown (eager, batch_d 1)"; timeout -s KILL 300 $B --kernel_timing 2 --top 14 --batch_d 1 2>$O/b_perop.err | tail -1 > $O/b_perop.json
python - <<PY
import json
r = json.load(open('$O/b_perop.json')); k = r['roofline']['by_kernel']
print('ms/step', r['ms_per_step'], 'launches', r['gpu_launches'])
for a, b in k.items():
    print('  %-28s %6.2f ms/step n=%d' % (a, b['ms'] / r['steps'], b['n'] / r['steps']), {x: round(y / r['steps'], 2) for x, y in list(b['top'].items())[:6]})
PY
echo "== graph + batch_d"; timeout -s KILL 300 $B --kernel_timing 1 --batch_d 1 --cuda_graph 1 2>$O/b_graph.err | tail -1 > $O/b_graph.json; cut -c1-900 $O/b_graph.json; grep -i capture $O/b_graph.err | head
echo "== graph + batch_d + pair3"; NEMAR_TC_PAIR=3 NEMAR_WG_PAIR_STAGES=3 timeout -s KILL 300 $B --kernel_timing 0 --batch_d 1 --cuda_graph 1 2>$O/b_graph_p3.err | tail -1 > $O/b_graph_p3.json; cut -c1-250 $O/b_graph_p3.json
echo "== ncu launch list"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_s2.csv python bench.py --profile --steps 1 --warmup 1 --no_cpu_baseline --grid_sample_bench 0 --kernel_timing 0 --batch_d 1 > $O/ncu_list.log 2>&1
echo "rc=$? lines=$(wc -l < $O/launches_s2.csv)"
