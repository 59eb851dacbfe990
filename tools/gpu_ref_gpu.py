#!/usr/bin/env python
"""Same-box comparator: the reference's OWN GPU path (oracle/_ref/grmonty_ref_gpu = its unmodified .cu/.cpp sources
built for sm_100a by `make -C oracle refgpu`) on the bench dump.  usage: tools/gpu_ref_gpu.py [photon_n ...]"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import make_harm_dump
exe = os.path.join(ROOT, "oracle", "_ref", "grmonty_ref_gpu")
cache = os.path.join(ROOT, "oracle", "_ref", "hotcross_table.bin")
p = "/tmp/gp_dump_192.txt"
if not os.path.exists(p):
    make_harm_dump.write_dump(p, *make_harm_dump.make_dump(n0=192, n1=192))
for photon_n in (sys.argv[1:] or ["100000"]):
    t0 = time.time()
    try:
        out = subprocess.run([exe, "--harm_dump_path", p, "--photon_n", photon_n, "--mass_unit", "4e19", "--hotcross_cache", cache],
                             capture_output=True, text=True, timeout=600)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else ""
        try:
            d = json.loads(line)
            d["rate"] = d["created"] / d["run_s"]
            d["impl"] = "reference-gpu (unmodified sources, sm_100a)"
            d["wall_s"] = time.time() - t0
            print(json.dumps(d), flush=True)
        except Exception:
            print("photon_n", photon_n, "rc", out.returncode, "stdout:", out.stdout[-400:], "stderr:", out.stderr[-600:], flush=True)
    except subprocess.TimeoutExpired:
        print("photon_n", photon_n, "timeout", flush=True)
