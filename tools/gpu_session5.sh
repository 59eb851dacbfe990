#!/bin/bash
set -u
out=gpurun_out
mkdir -p $out
GRMONTY_B200_TRACE=1 timeout 900 python tools/gpu_sweep.py 1e6 default f256x1,w256x2,w384x2,w128x2x2,w128x3x2,w128x2x3,w192x2x2,w256x3 2>&1 | tee $out/s5_sweep.txt
for thr in 192,64 160,32 224,32 192,16; do
  echo "== thr $thr"; GRMONTY_B200_WF_THR=$thr timeout 300 python tools/gpu_sweep.py 1e6 default w128x2x2 2>&1 | tee -a $out/s5_thr.txt
done
