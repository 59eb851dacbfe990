#!/bin/bash
set -u
out=gpurun_out
mkdir -p $out
export GRMONTY_B200_WATCHDOG_S=5
GRMONTY_B200_TRACE=1 timeout 300 python - > $out/p2_small.log 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import cuda_grmonty_b200 as gm
d = dict(np.load("tests/golden/functions_48.npz"))
m = {k[6:]: (v.item() if v.ndim == 0 else v) for k, v in d.items() if k.startswith("model_")}
for cap in (0, 1 << 17, 1 << 16):
    try:
        c = gm.Context(m, seed=77, gen0=64, gen_cap=1 << 12, queue_capacity=cap)
        c.run(0, 20000); r = c.result(); c.close()
        print("CAP", cap, r["created"], r["recorded"], r["scattered"], r["stats"]["n_tracked"], flush=True)
    except Exception as e:
        print("CAP", cap, "FAILED", e, flush=True)
PY
grep -v "    generation" $out/p2_small.log | cut -c1-700
