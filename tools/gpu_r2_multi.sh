#!/bin/bash
# round-2 multi-GPU check (run with gpurun --gpus N): the C-ABI allreduce test, the CLI's --gpus N, both bench arms under torchrun
set -u
N=${1:-2}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv,noheader > $out/r2_multi_smi.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -q > $out/r2_multi_test.log 2>&1
echo "multi test rc=$?"; tail -5 $out/r2_multi_test.log
python - > $out/r2_cli_gpus.log 2>&1 <<PY
import os, subprocess, sys, time
sys.path.insert(0, os.getcwd())
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
gm.build_host()
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
for g in (1, $N):
    t0 = time.time()
    o = subprocess.run([gm.CLI, "--harm_dump_path", p, "--spectrum_path", f"/tmp/spec_{g}.txt", "--photon_n", "1000000",
                        "--mass_unit", "4e19", "--gpus", str(g)], capture_output=True, text=True, timeout=600)
    print("gpus", g, "rc", o.returncode, "wall", round(time.time() - t0, 2))
    print(o.stdout[-1500:]); print(o.stderr[-1500:])
PY
echo "cli rc=$?"; tail -30 $out/r2_cli_gpus.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > $out/r2_bench_n$N.json 2> $out/r2_bench_n$N.err
echo "bench N=$N rc=$?"; cat $out/r2_bench_n$N.json; tail -12 $out/r2_bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > $out/r2_bench_ref_n$N.json 2> $out/r2_bench_ref_n$N.err
echo "ref N=$N rc=$?"; cat $out/r2_bench_ref_n$N.json; tail -5 $out/r2_bench_ref_n$N.err
