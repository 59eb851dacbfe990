#!/bin/bash
# A/B of library variants built into gpurun_ab/lib_<name>.so: each is copied over the product library and timed with
# the bench command (two rounds, interleaved, so box drift shows).   usage: tools/gpu_ab.sh <name> <name> ...
set -u
lib=cuda-grmonty_b200/libgrmonty_b200.so
cp $lib /tmp/lib_keep.so
mkdir -p gpurun_out
for round in 1 2; do
  for v in "$@"; do
    cp gpurun_ab/lib_$v.so $lib
    timeout 300 python bench.py --steps 3 --warmup 2 --no_cpu_baseline > gpurun_out/ab_${v}_$round.json 2> gpurun_out/ab_${v}_$round.err
    python - <<PY
import json
d = json.load(open("gpurun_out/ab_${v}_$round.json"))
print("$v", $round, "ms/step", round(d["ms_per_step"], 1), "value", round(d["value"] / 1e6, 2), "e2e", round(d["e2e"]["value"] / 1e6, 2), "rec", d["run"]["recorded"])
PY
  done
done
cp /tmp/lib_keep.so $lib
