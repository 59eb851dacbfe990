#!/bin/bash
# Round-2 first GPU visit (one B200): baselines the round-2 kernel work is judged against.
#   1. occupancy sweep of the round-1 kernel (256x1 = 8 warps/SM at 255 regs, 384x1 / 128x3 / 192x2 = 12 warps at
#      168 regs with spills, 512x1 = 16 warps at 128 regs)
#   2. end-of-run bias statistics at photon_n = 1e6 .. 8e6 (seeds of the reference arm's bounded sample)
#   3. 1024^2 grid: bench line at photon_n = 1e6, L2 access-policy window on / off
#   4. compute-sanitizer memcheck + racecheck on a 4000-primary run
set -u
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $out/s1_smi.txt 2>&1
nproc >> $out/s1_smi.txt
echo "== sweep"; timeout 600 python tools/gpu_sweep.py 1e6 default 256x1,384x1,128x3,192x2,512x1,128x2 2>&1 | tee $out/s1_sweep.txt
echo "== end states"; timeout 300 python tools/gpu_end_states.py 2>&1 | tee $out/s1_end_states.txt
echo "== 1024 window on"; timeout 600 python tools/gpu_configs.py c5 1e6 2>&1 | tee $out/s1_c5_window_on.txt
echo "== 1024 window off"; GRMONTY_B200_L2_WINDOW=0 timeout 600 python tools/gpu_configs.py c5 1e6 2>&1 | tee $out/s1_c5_window_off.txt
echo "== sanitizer memcheck"; timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_sanitize_run.py 2000 > $out/s1_memcheck.txt 2>&1; tail -5 $out/s1_memcheck.txt
echo "== sanitizer racecheck"; timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python tools/gpu_sanitize_run.py 1000 > $out/s1_racecheck.txt 2>&1; tail -5 $out/s1_racecheck.txt
