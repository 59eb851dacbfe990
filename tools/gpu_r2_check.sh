#!/bin/bash
# round-2 status check on one B200: the whole GPU test suite, then both bench arms
set -u
out=gpurun_out
mkdir -p $out
rm -f $out/parity_report.json
GRMONTY_B200_PARITY_REPORT=$PWD/$out/parity_report.json timeout 2400 python -m pytest tests -m gpu -q --durations=15 > $out/r2_gputests.log 2>&1
echo "tests rc=$?"; tail -60 $out/r2_gputests.log
timeout 600 python bench.py --steps 3 --warmup 3 > $out/r2_bench.json 2> $out/r2_bench.err
echo "bench rc=$?"; cat $out/r2_bench.json; tail -5 $out/r2_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $out/r2_bench_ref.json 2> $out/r2_bench_ref.err
echo "ref rc=$?"; cut -c1-300 $out/r2_bench_ref.json; tail -5 $out/r2_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline > $out/r2_launches_ncu.log 2>&1
echo "launch list rc=$?"; wc -l $out/r2_launches_bench.csv
GRMONTY_B200_PROFILE_MIN_COUNT=700000 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:transport_kernel -c 1 -f -o $out/r2_transport_32x8 python tools/gpu_gen_profile.py 200 0 192 f0x0 > $out/r2_ncu_32x8.log 2>&1
echo "ncu rc=$?"; tail -2 $out/r2_ncu_32x8.log
