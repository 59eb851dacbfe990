#!/bin/bash
# the 8-GPU visit (gpurun --gpus 8): configs[2] at its stated size, strong scaling at a fixed job, the weak-scaling bench line
set -u
N=${1:-8}
out=gpurun_out
mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
timeout 600 bash -c "$(declare -f run); run $N 29601 tools/gpu_configs2.py 1e8" > $out/r2_configs2_n$N.json 2> $out/r2_configs2_n$N.err
echo "configs2 rc=$?"; cat $out/r2_configs2_n$N.json; tail -3 $out/r2_configs2_n$N.err
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  timeout 600 bash -c "$(declare -f run); run $n $((29610 + n)) bench.py --gpus $n --steps 3 --warmup 3 --scaling strong --photon_n 8e6 --no_cpu_baseline" > $out/r2_strong_n$n.json 2> $out/r2_strong_n$n.err
  echo "strong n=$n rc=$?"; python - <<PY
import json
try:
    d = json.load(open("$out/r2_strong_n$n.json"))
    print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "scaling")}, "e2e", d["e2e"]["value"], "attempts/s", d["work_rates"]["push_attempts_per_s"])
except Exception as e:
    print("no line:", e)
PY
done
timeout 600 bash -c "$(declare -f run); run $N 29630 bench.py --gpus $N --steps 3 --warmup 3" > $out/r2_bench_n$N.json 2> $out/r2_bench_n$N.err
echo "weak N=$N rc=$?"; cut -c1-700 $out/r2_bench_n$N.json; tail -9 $out/r2_bench_n$N.err
