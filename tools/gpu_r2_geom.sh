#!/bin/bash
set -u
out=gpurun_out
for round in 1 2; do
 for pn in 1e6 1e5; do
  timeout 900 python tools/gpu_sweep.py $pn gpurun_ab/lib_g67.so f32x8,f32x7,f32x6 2>&1 | sed 's/"wall_ms": [0-9.]*, //; s/"rate".*"recorded"/"recorded"/' | cut -c1-170 | tee -a $out/r2_geom.txt
 done
done
