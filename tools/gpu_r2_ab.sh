#!/bin/bash
# interleaved A/B of library variants gpurun_ab/lib_<name>.so on configs[1] (default scheduler): tools/gpu_r2_ab.sh name name ...
set -u
out=gpurun_out
mkdir -p $out
libs=$(for v in "$@"; do printf "gpurun_ab/lib_%s.so," $v; done); libs=${libs%,}
for round in 1 2 3; do
  timeout 600 python tools/gpu_sweep.py 1e6 $libs f256x1 2>&1 | sed 's/"wall_ms": [0-9.]*, //; s/"rate".*"recorded"/"recorded"/' | cut -c1-200 | tee -a $out/r2_ab.txt
done
