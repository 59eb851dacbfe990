#!/bin/bash
set -u
out=gpurun_out
for round in 1 2; do
 for pn in 1e6; do
  timeout 900 python tools/gpu_sweep.py $pn gpurun_ab/lib_base.so,gpurun_ab/lib_is64.so,gpurun_ab/lib_rb16.so,gpurun_ab/lib_rb8.so f0x0 2>&1 | sed 's/"wall_ms": [0-9.]*, //; s/"rate".*"recorded"/"recorded"/' | cut -c1-170 | tee -a $out/r2_final_tune.txt
 done
done
