#!/bin/bash
# One GPU-box visit: parity tests, a bench line, the ncu launch list of the same bench command and (optionally)
# one full ncu capture of the dominant transport launch.  Everything lands in gpurun_out/<tag>_*.
#   usage: tools/gpu_round.sh <tag> [tests] [bench] [launches] [ncu]
set -u
tag=${1:-r1}; shift || true
what=${*:-tests bench launches}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $out/${tag}_smi.txt 2>&1
nproc >> $out/${tag}_smi.txt
for w in $what; do
  case $w in
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1
      echo "tests rc=$?"; tail -3 $out/${tag}_tests.log ;;
    bench)
      timeout 900 python bench.py --steps 3 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
      echo "bench rc=$?"; cat $out/${tag}_bench.json ;;
    refarm)
      timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_refarm.json 2> $out/${tag}_refarm.err
      echo "refarm rc=$?"; cat $out/${tag}_refarm.json ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
        --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline \
        > $out/${tag}_launches.log 2>&1
      echo "launches rc=$?" ;;
    ncu)
      GRMONTY_B200_PROFILE_MIN_COUNT=700000 timeout 1200 ncu --set full --clock-control none --import-source on \
        --profile-from-start off -k regex:transport_kernel -c 1 -f -o $out/${tag}_transport \
        python tools/gpu_gen_profile.py 200 0 192 > $out/${tag}_ncu.log 2>&1
      echo "ncu rc=$?" ;;
    genprof)
      timeout 600 python tools/gpu_gen_profile.py 200 0 192 > $out/${tag}_genprof.log 2>&1
      echo "genprof rc=$?"; tail -8 $out/${tag}_genprof.log ;;
  esac
done
