#!/bin/bash
# Round-2 third GPU visit: the wavefront kernel -- correctness first (geometry / kernel invariance, parity, edges),
# then the variant sweep against the fused kernel on configs[1].
set -u
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_invariance.py tests/test_gpu_parity.py tests/test_gpu_edge.py -x -q > $out/s3_tests.log 2>&1
echo "tests rc=$?"; tail -15 $out/s3_tests.log
timeout 900 python tools/gpu_sweep.py 1e6 default f256x1,w256x2,w256x3,w384x2,w384x1,w512x1,w128x2 2>&1 | tee $out/s3_sweep.txt
