#!/bin/bash
set -u
out=gpurun_out
for round in 1 2; do
 for pn in 1e6 1e5; do
  timeout 900 python tools/gpu_sweep.py $pn gpurun_ab/lib_sp0.so,gpurun_ab/lib_sp1.so f0x0 2>&1 | sed 's/"wall_ms": [0-9.]*, //; s/"rate".*"recorded"/"recorded"/' | cut -c1-170 | tee -a $out/r2_spare.txt
 done
done
