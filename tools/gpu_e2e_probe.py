#!/usr/bin/env python
"""Where the end-to-end time of one HARMModel.run_simulation goes: create / run / result / destroy timed apart."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_grmonty_b200 as gm
from tools import make_harm_dump
photon_n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
hm = gm.HarmModel(photon_n, 4e19); hm.read_file(p); hm.init()
model = hm.model_dict()
for rep in range(int(os.environ.get('REPS', 3))):
    t0 = time.perf_counter(); c = gm.Context(model, seed=123)
    t1 = time.perf_counter(); c.run()
    t2 = time.perf_counter(); r = c.result()
    t3 = time.perf_counter(); c.close()
    t4 = time.perf_counter()
    st = r["stats"]
    print(f"rep {rep}: create {1e3*(t1-t0):8.1f} ms  run {1e3*(t2-t1):8.1f} ms (kernels {st['kernel_ms']:.1f}, transport "
          f"{st['transport_ms']:.1f}, gens {st['n_generations']}, launches {st['n_kernel_launches']})  result {1e3*(t3-t2):6.1f} ms  "
          f"destroy {1e3*(t4-t3):8.1f} ms")
hm.set_options(seed=123)
for rep in range(int(os.environ.get('REPS', 3))):
    t0 = time.perf_counter(); hm.run_simulation(); t1 = time.perf_counter()
    print(f"HARMModel.run_simulation {1e3*(t1-t0):8.1f} ms  stats {hm.stats()['seconds']*1e3:.1f}")
