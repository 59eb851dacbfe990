#!/bin/bash
# J2 evidence on one B200: configs[4] at its stated size (1024 x 1024, photon_n = 1e8), and ncu --set full of one
# full-size generation on that grid with the L2 access-policy window on and off
set -u
out=gpurun_out
mkdir -p $out
timeout 900 python tools/gpu_configs.py c5 1e8 > $out/r2_configs4_1e8.json 2> $out/r2_configs4_1e8.err
echo "configs4 rc=$?"; cat $out/r2_configs4_1e8.json | cut -c1-900; tail -3 $out/r2_configs4_1e8.err
for w in 1 0; do
  GRMONTY_B200_L2_WINDOW=$w GRMONTY_B200_PROFILE_MIN_COUNT=700000 timeout 900 ncu --set full --clock-control none --import-source on \
    --profile-from-start off -k regex:transport_kernel -c 1 -f -o $out/r2_grid1024_window$w \
    python tools/gpu_gen_profile.py 200 0 1024 f256x1 > $out/r2_grid1024_ncu_window$w.log 2>&1
  echo "ncu window=$w rc=$?"
done
