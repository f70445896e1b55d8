#!/bin/bash
# Round 2: is the fp32 engine's T/R gradient error the discriminator's update or its backward? (tests/probes/fp32_grad_error_probe2.py)
mkdir -p gpurun_out
timeout -s KILL 500 python tests/probes/fp32_grad_error_probe2.py > gpurun_out/r2t_probe2.txt 2>&1; echo rc=$?
grep -v "Warning\|warn\|initialize\|created\|out\[name\]\|Consider" gpurun_out/r2t_probe2.txt | cut -c1-200
