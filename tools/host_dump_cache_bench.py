#!/usr/bin/env python
"""Time HarmModel.read_file on the text dump and on its binary cache.
usage: tools/host_dump_cache_bench.py [n ...]   (grid sizes, default 192 1024)"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cuda_grmonty_b200 as gm
from tools import make_harm_dump

gm.build_host()
for n in [int(a) for a in sys.argv[1:]] or [192, 1024]:
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, f"dump{n}.txt")
        header, table = make_harm_dump.make_dump(n0=n, n1=n)
        make_harm_dump.write_dump(p, header, table)
        m = gm.HarmModel(1000, 4e19)
        t = []
        for _ in range(3):
            t0 = time.perf_counter(); m.read_file(p); t.append(time.perf_counter() - t0)
        text_s = min(t)
        ref = {k: v.copy() for k, v in m.model_dict().items() if isinstance(v, np.ndarray) and v.ndim == 2}
        m.set_dump_cache(True)
        t0 = time.perf_counter(); m.read_file(p); store_s = time.perf_counter() - t0
        t = []
        for _ in range(3):
            t0 = time.perf_counter(); m.read_file(p); t.append(time.perf_counter() - t0)
            assert m.read_from_cache()
        got = m.model_dict()
        assert all(np.array_equal(ref[k], got[k]) for k in ref)
        print(f"{n}x{n}: text {os.path.getsize(p) / 1e6:.1f} MB parse {text_s * 1e3:.1f} ms | parse+store {store_s * 1e3:.1f} ms | "
              f"cache {os.path.getsize(p + '.b200cache') / 1e6:.1f} MB load {min(t) * 1e3:.1f} ms | x{text_s / min(t):.0f}, bit-identical grids")
