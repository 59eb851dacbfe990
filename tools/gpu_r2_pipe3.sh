#!/bin/bash
# A/B of pipeline-kernel variants (gpurun_ab/lib_<name>.so) against the round-1 scheduler, photon_n = 1e6 and 1e5
set -u
out=gpurun_out
mkdir -p $out
libs=$(for v in "$@"; do printf "gpurun_ab/lib_%s.so," $v; done); libs=${libs%,}
for round in 1 2; do
  for pn in 1e6 1e5; do
    GRMONTY_B200_OVERLAP=0 timeout 300 python tools/gpu_sweep.py $pn gpurun_ab/lib_$1.so f256x1 2>&1 | sed "s/^/legacy $pn /" | cut -c1-330 | tee -a $out/p3_ab.txt
    timeout 600 python tools/gpu_sweep.py $pn $libs f256x1 2>&1 | sed "s/^/pipe $pn /" | cut -c1-330 | tee -a $out/p3_ab.txt
  done
done
