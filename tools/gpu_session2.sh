#!/bin/bash
# Round-2 second GPU visit: why do 12 warps/SM (168 registers, spills) run slower than 8 (255 registers)?
# ncu --set full of the same full-size generation for both launch geometries + the per-generation profile.
set -u
out=gpurun_out
mkdir -p $out
for v in 256x1 384x1; do
  GRMONTY_B200_PROFILE_MIN_COUNT=700000 timeout 900 ncu --set full --clock-control none --import-source on \
    --profile-from-start off -k regex:transport_kernel -c 1 -f -o $out/s2_transport_$v \
    python tools/gpu_gen_profile.py 200 0 192 $v > $out/s2_ncu_$v.log 2>&1
  echo "ncu $v rc=$?"
done
timeout 300 python tools/gpu_gen_profile.py 200 0 192 256x1 > $out/s2_genprof_256x1.log 2>&1; tail -12 $out/s2_genprof_256x1.log
