#!/bin/bash
# wavefront kernel diagnosis: per-generation profile of both kernels, phase statistics, one ncu capture
set -u
out=gpurun_out
mkdir -p $out
for v in f256x1 w256x2 w384x1; do
  timeout 300 python tools/gpu_gen_profile.py 200 0 192 $v > $out/s4_genprof_$v.log 2>&1; echo "== $v"; tail -13 $out/s4_genprof_$v.log
done
GRMONTY_B200_TRACE=1 timeout 300 python tools/gpu_sweep.py 1e6 default w256x2 > $out/s4_trace_w256x2.log 2>&1
GRMONTY_B200_TRACE=1 timeout 300 python - > $out/s4_phase_stats.log 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
p = "/tmp/gp_dump_192.txt"
hm = gm.HarmModel(1000000, 4e19); hm.read_file(p); hm.init()
m = hm.model_dict()
for spec in ((256, 2), (384, 1), (384, 2)):
    c = gm.Context(m, threads_per_block=spec[0], slots_per_thread=spec[1])
    c.run(); r = c.result(); print(spec, r["stats"]["transport_ms"], flush=True); c.close()
PY
grep -E "wavefront block|^\(" $out/s4_phase_stats.log
GRMONTY_B200_PROFILE_MIN_COUNT=700000 timeout 900 ncu --set full --clock-control none --import-source on \
    --profile-from-start off -k regex:wavefront_kernel -c 1 -f -o $out/s4_wavefront_w256x2 \
    python tools/gpu_gen_profile.py 200 0 192 w256x2 > $out/s4_ncu.log 2>&1
echo "ncu rc=$?"
