#!/bin/bash
# Run on the GPU box through gpurun: each test file in its own process (a CUDA fault must not poison the rest).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
ls /root/reference baseline/_ref > gpurun_out/ref_presence.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/gpu.txt
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "== $f exit $?"; tail -n 25 gpurun_out/$name.log
done
