#!/bin/bash
set -u
out=gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/r2_smoke.log
timeout 600 python bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > $out/r2_bench_ref_as_n8.json 2> $out/r2_bench_ref_as_n8.err; echo "ref(gpus=8) rc=$?"; cut -c1-1400 $out/r2_bench_ref_as_n8.json; tail -3 $out/r2_bench_ref_as_n8.err
