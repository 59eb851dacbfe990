#!/usr/bin/env python
"""bench.py -- superphotons/s of the B200 transport path (BASELINE.json metric), with the reference CPU build
timed beside it.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1], the README benchmark shape -- synthetic dump019-shaped
HARM dump (192x192, a = 0.9375, tools/make_harm_dump.py), mass_unit = 4e19, photon_n = 1e6 PER GPU (weak
scaling: the job's photon_n is N x 1e6 and rank r tracks the primaries at positions j = r mod N).
A "step" is one complete run_simulation: every primary superphoton of the rank's share is generated,
transported (with all its scattered descendants) and recorded.

    value      whole-job primaries/s with the model resident in HBM (context created once, reset per step)
    e2e        the same through the HARMModel host API with HOST buffers: create (H2D of grid + tables) +
               run + result (D2H of the spectrum) + destroy inside the timed region
    roofline   FP64 pipe: algorithmic flops (650 per vacuum step, 1100 per in-fluid step, 3000 per
               scattering, 628 per extra push attempt) / CUDA-event time of the transport kernels, against the
               DFMA peak measured in the same process (MEASURED_PEAKS.json has no FP64 entry)
    cpu_baseline  the UNMODIFIED reference CPU build (oracle/_ref/grmonty_ref) on all host cores, one
               single-threaded process per core as in BASELINE.md section 3, on a bounded photon_n
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_VAC, FLOP_FLUID, FLOP_SCAT, FLOP_EXTRA = 650.0, 1100.0, 3000.0, 628.0


def dump_path(n0: int, n1: int) -> str:
    """synthetic dump, generated once per box (deterministic)"""
    from tools import make_harm_dump
    p = os.path.join(tempfile.gettempdir(), f"grmonty_b200_dump_{n0}x{n1}.txt")
    if not os.path.exists(p):
        header, table = make_harm_dump.make_dump(n0=n0, n1=n1)
        tmp = p + f".{os.getpid()}.tmp"
        make_harm_dump.write_dump(tmp, header, table)
        os.replace(tmp, p)
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, device: int):
        self.rows, self.proc = [], None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={device}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# End state of the scattering-bias statistics of complete runs on the bench dump (192x192, mass_unit 4e19):
# photon_n -> (max_tau_scatt, scattered / created, recorded / created).  The bounded reference sample starts from
# this state, so that it does the per-primary work of the full run it samples (the reference's bias divides by running
# statistics, harm_model.cpp:1391-1404: a sample that starts from empty statistics scatters ~20 % more per primary
# than the complete run).  photon_n = 1e6: medians of the reference's own 14 complete runs
# (tests/golden/spectrum_192_4e19_1e6*.npz); the larger photon_n: complete runs of the CUDA path
# (profiles/r2_end_states.txt; its statistics track the reference's, tests/test_gpu_spectrum.py), because a complete
# reference run takes 1 - 5 hours per core there.
REF_END_STATE = {1.0e6: (0.00256, 0.626, 0.752), 2.0e6: (0.00269, 0.604, 0.740), 4.0e6: (0.00260, 0.587, 0.732),
                 8.0e6: (0.00281, 0.541, 0.707)}
PRIMARIES_PER_PHOTON_N = 16.118  # created / photon_n on the bench dump: ln(nu_max / nu_min), harm_model.cpp:302


def ref_end_state(photon_n: float):
    """REF_END_STATE interpolated in ln(photon_n) (clamped)"""
    ks = sorted(REF_END_STATE)
    x = min(max(np.log(photon_n), np.log(ks[0])), np.log(ks[-1]))
    return tuple(float(np.interp(x, np.log(ks), [REF_END_STATE[k][i] for k in ks])) for i in range(3))


def run_reference_cpu(dump: str, photon_n: float, mass_unit: float, cores: int, target_s: float, seed0: int = 123,
                      seeded: bool = True) -> dict:
    """A bounded sample of the configs[1] workload on the reference CPU build: `cores` concurrent single-threaded
    processes (the reference's CPU loop has no threading, harm_model.cpp:366-404), each running 1/thin of the primaries
    of a run AT THE JOB'S OWN photon_n, spread over all zones like the run itself (oracle/ref_harness.cpp
    ref_run_thinned: every call is a reference function), with the bias statistics started from the end state of a
    complete run.  rate = sum of primaries / max wall of the transport loop (table init excluded like the reference's
    own timer, harm_model.cpp:409; the hot cross-section table is cached on disk by the harness).
    A process that dies or prints no result line is dropped and reported, not fatal."""
    from oracle import refharness as rh
    total = PRIMARIES_PER_PHOTON_N * photon_n
    thin = max(1.0, total / (6500.0 * max(1.0, target_s)))     # ~6.5 k primaries/s/core
    if os.path.exists(rh.CLI_PATH):
        cmd0 = [rh.CLI_PATH, "--harm_dump_path", dump, "--photon_n", str(int(round(photon_n))), "--mass_unit",
                repr(mass_unit), "--hotcross_cache", rh.HOTCROSS_CACHE, "--thin", repr(thin)]
        if seeded:
            mt, sc, rc = ref_end_state(photon_n)
            cmd0 += ["--bias_stats", f"{mt!r},{int(sc * total)},{int(rc * total)}"]
        procs = [subprocess.Popen(cmd0 + ["--seed", str(seed0 + i)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                  text=True) for i in range(cores)]
        outs, errors = [], []
        for i, p in enumerate(procs):
            so, se = p.communicate()
            line = next((ln for ln in reversed(so.splitlines()) if ln.startswith('{"impl": "reference-cpu"')), None)
            try:
                if p.returncode != 0 or line is None:
                    raise ValueError(f"rc={p.returncode}")
                outs.append(json.loads(line))
            except ValueError as e:
                errors.append(f"process {i} (seed {seed0 + i}): {e}; stderr tail: {se.strip()[-160:]!r}")
        if not outs:
            raise RuntimeError("no reference process finished: " + "; ".join(errors)[:600])
        created = sum(o["created"] for o in outs)
        wall = max(o["run_s"] for o in outs)
        res = {"value": created / wall, "unit": "superphotons/s", "cores": len(outs), "kind": "reference",
               "sample": f"{len(outs)} processes x 1/{thin:.4g} of the primaries of a photon_n={photon_n:g} run "
                         f"(reference CPU build, unmodified sources, all zones, weights of the full run, bias "
                         f"statistics started from a complete run's end state), {created} primaries in "
                         f"{wall:.1f} s of the transport loop",
               "photon_n": photon_n, "thin": thin, "seeded_bias_stats": seeded,
               "recorded": sum(o["recorded"] for o in outs), "scattered": sum(o["scattered"] for o in outs),
               "created": created, "seconds": wall,
               "scattered_per_created": sum(o["scattered"] for o in outs) / max(1, created)}
        if errors:
            res["dropped_processes"] = errors
        return res
    # the reference build is not on this box: time the plain-C port instead (oracle/grmonty_oracle.c)
    import multiprocessing as mp
    port_n = max(200, int(photon_n / thin))
    with mp.Pool(cores) as pool:
        res = pool.map(_oracle_port_run, [(dump, port_n, mass_unit, seed0 + i) for i in range(cores)])
    created = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return {"value": created / wall, "unit": "superphotons/s", "cores": cores, "kind": "port",
            "sample": f"{cores} processes x complete runs at photon_n={port_n} (plain-C port of the reference "
                      f"algorithm; NOT the job's photon_n: the reference build is absent on this box)",
            "photon_n": port_n, "created": created, "seconds": wall}


def run_reference_gpu(dump: str, photon_n: int, mass_unit: float) -> dict | None:
    """The reference's OWN GPU path (oracle/_ref/grmonty_ref_gpu: its unmodified .cu/.cpp sources compiled for
    sm_100a by `make -C oracle refgpu`) on the same dump, same photon_n, same box: SURVEY 8(f) N2.  A performance
    comparator only (float RNG, half-step bug, racy bias statistics: not an oracle)."""
    from oracle import refharness as rh
    exe = os.path.join(os.path.dirname(rh.CLI_PATH), "grmonty_ref_gpu")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe, "--harm_dump_path", dump, "--photon_n", str(int(photon_n)), "--mass_unit",
                              repr(mass_unit), "--hotcross_cache", rh.HOTCROSS_CACHE], capture_output=True, text=True,
                             timeout=600)
        o = json.loads(next(ln for ln in reversed(out.stdout.splitlines()) if ln.startswith('{"impl"')))
    except Exception as e:  # noqa: BLE001 -- a comparator that fails to run is reported, not fatal
        return {"error": str(e)[:200]}
    return {"value": o["created"] / o["run_s"], "unit": "superphotons/s", "kind": "reference GPU build (unmodified "
            "super_photon.cu, sm_100a) on this GPU", "created": o["created"], "recorded": o["recorded"],
            "scattered": o["scattered"], "seconds": o["run_s"], "photon_n": int(photon_n)}


def _oracle_port_run(args):
    dump, photon_n, mass_unit, seed = args
    import cuda_grmonty_b200 as gm
    from oracle import orc
    gm.build_host()
    hm = gm.HarmModel(photon_n, mass_unit)
    hm.read_file(dump)
    hm.init(1)
    M = orc.Model(hm.model_dict(), seed=seed, stats_mode=1, zone_order=1)
    t0 = time.time()
    M.run(budget=0)
    return int(M.m.n_created), time.time() - t0


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the newest committed ncu
    summary of the transport kernel under profiles/ (written by tools/ncu_summary.py); (None, reason) if absent"""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_transport_ncu*.txt")),
                   key=lambda f: (int(re.search(r"r(\d+)_", os.path.basename(f)).group(1)), os.path.getmtime(f)))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for f in reversed(files):
        got = {}
        for ln in open(f):
            m = re.match(r"dram__bytes_(read|write)\.sum\s+([0-9.eE+-]+)\s+(\w+)", ln)
            if m and m.group(3) in scale and m.group(1) not in got:   # the file's FIRST capture is the default kernel's
                got[m.group(1)] = float(m.group(2)) * scale[m.group(3)]
        if len(got) == 2:
            return got["read"] + got["write"], os.path.relpath(f, ROOT)
    return None, "no ncu summary under profiles/"


def flops_of(stats: dict) -> float:
    n_fluid = stats["n_interactions"]
    n_vac = max(0, stats["n_steps"] - n_fluid)
    n_extra = max(0, stats["n_push_attempts"] - stats["n_steps"] - stats["n_scatter_events"])
    return (n_vac * FLOP_VAC + n_fluid * FLOP_FLUID + stats["n_scatter_events"] * FLOP_SCAT + n_extra * FLOP_EXTRA)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--photon_n", type=float, default=1.0e6, help="photon_n per GPU")
    ap.add_argument("--mass_unit", type=float, default=4.0e19)
    ap.add_argument("--n0", type=int, default=192)
    ap.add_argument("--n1", type=int, default=192)
    ap.add_argument("--ref_seconds", type=float, default=0.0,
                    help="CPU seconds per reference process and step (0: 150 / steps, clamped to 6 .. 25)")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --photon_n is the JOB's photon_n, every GPU takes 1/N of it (fixed total work)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.scaling == "strong":
        args.photon_n = args.photon_n / max(world, args.gpus)
    workload = (f"synthetic dump019-shaped HARM dump {args.n0}x{args.n1} (a=0.9375), mass_unit={args.mass_unit:g}, "
                f"photon_n={args.photon_n:g} per GPU (job photon_n={args.photon_n * max(world, args.gpus):g})")
    ref_seconds = args.ref_seconds if args.ref_seconds > 0 else min(25.0, max(6.0, 150.0 / max(1, args.steps)))
    config = {"workload": workload, "photon_n_per_gpu": args.photon_n, "mass_unit": args.mass_unit,
              "grid": [args.n0, args.n1], "sharding": f"photons x{world}, end-of-run allreduce",
              "l2": "per-step working set (photon pool, ~250 B x millions of photons) exceeds L2; no flush needed"}

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        dump = dump_path(args.n0, args.n1)
        cores = os.cpu_count() or 1
        job_photon_n = args.photon_n * max(world, args.gpus)   # the job's photon_n: weak scaling, N x photon_n per GPU
        for _ in range(min(args.warmup, 1)):                   # page cache / table file: one short pass is all a CPU needs
            run_reference_cpu(dump, job_photon_n, args.mass_unit, cores, 1.0)
        t0 = time.time()
        res = [run_reference_cpu(dump, job_photon_n, args.mass_unit, cores, ref_seconds, 123 + 1000 * s)
               for s in range(args.steps)]
        wall = time.time() - t0
        value = sum(r["created"] for r in res) / sum(r["seconds"] for r in res)
        cb = dict(res[-1])
        cb["value"] = value
        dropped = [e for r in res for e in r.get("dropped_processes", [])]
        if dropped:
            cb["dropped_processes"] = dropped
        print(json.dumps({
            "impl": "reference", "metric": "superphotons/sec", "value": value, "unit": "superphotons/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "superphotons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------ B200 arm
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import cuda_grmonty_b200 as gm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the transport path has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    photon_n_total = int(round(args.photon_n * world))

    # untimed set-up: synthetic dump, host model (loader + tables), device context
    if rank == 0:
        dump = dump_path(args.n0, args.n1)
    if dist:
        dist.barrier()
    dump = dump_path(args.n0, args.n1)
    hm = gm.HarmModel(photon_n_total, args.mass_unit)
    t0 = time.time()
    hm.read_file(dump)
    hm.init()
    init_s = time.time() - t0
    model = hm.model_dict()
    ctx = gm.Context(model, seed=123, rank=rank, world=world, device=dev)
    fp64_peak = ctx.fp64_peak()
    # the path's only collective is the product's own: grmonty_b200_allreduce on the device accumulators, over a
    # communicator made through the same ABI (torch.distributed only ships the 128-byte NCCL id and does the timing
    # barrier / max over ranks)
    comm = None
    if dist:
        uid = torch.frombuffer(bytearray(gm.nccl_unique_id() if rank == 0 else bytes(gm.NCCL_ID_BYTES)),
                               dtype=torch.uint8).to(f"cuda:{dev}")
        dist.broadcast(uid, 0)
        comm = gm.nccl_comm_init_rank(uid.cpu().numpy().tobytes(), rank, world, dev)

    def step():
        ctx.reset()
        ctx.run()
        if comm:  # end-of-run reduction of spectrum, counters and max tau, in place on the device
            ctx.allreduce(comm)
        return ctx.result()

    def sync():
        if dist:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    sync()
    sampler = ClockSampler(dev) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    kernel_ms = transport_ms = 0.0
    launches = 0
    flops = 0.0
    res = None
    for _ in range(args.steps):
        res = step()
        kernel_ms += res["stats"]["kernel_ms"]
        transport_ms += res["stats"]["transport_ms"]
        launches += res["stats"]["n_kernel_launches"]
        flops += flops_of(res["stats"])
    ev1.record()
    sync()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    print(f"[bench rank {rank}] wall {1e3 * wall / args.steps:.1f} ms/step, kernels {kernel_ms / args.steps:.1f} ms/step, "
          f"generations {res['stats']['n_generations']}, attempts {res['stats']['n_push_attempts']}", file=sys.stderr)
    # work units of the last step on this rank (every step repeats the same run): tracked photons, accepted geodesic
    # steps, push attempts -- the honest work units next to the headline primaries/s
    work_local = [float(res["stats"][k]) for k in ("n_tracked", "n_steps", "n_push_attempts")]
    local = torch.tensor([wall, kernel_ms, transport_ms, flops, float(launches)] + work_local, dtype=torch.float64,
                         device=f"cuda:{dev}")
    if dist:
        mx = local.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = local.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        wall, kernel_ms, transport_ms = mx[0].item(), mx[1].item(), mx[2].item()
        flops_all, launches_all = sm[3].item(), int(sm[4].item())
        work_all = [sm[5].item(), sm[6].item(), sm[7].item()]
    else:
        flops_all, launches_all = flops, launches
        work_all = work_local
    total = ctx.total_primaries()  # primaries of the whole job (all ranks) per step
    value = total * args.steps / wall

    # ---- e2e: through the HARMModel host API with host buffers (create + H2D + run + D2H + destroy per step)
    # run_simulation = create (H2D) + run + grmonty_b200_allreduce (device, NCCL) + result (D2H) + destroy
    hm.set_options(seed=123, rank=rank, world=world, device=dev, nccl_comm=comm)
    ctx.close()
    hm.run_simulation()  # warm-up
    sync()
    t0 = time.perf_counter()
    t_run = 0.0
    for _ in range(args.steps):
        ta = time.perf_counter()
        hm.run_simulation()
        t_run += time.perf_counter() - ta
    sync()
    e2e_wall = time.perf_counter() - t0
    e2e_stats = hm.stats()
    if dist:
        tw = torch.tensor([e2e_wall], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        e2e_wall = tw.item()
    nz = args.n0 * args.n1
    h2d = 8 * (9 * nz + 221 * 81 + 3 * 201 + 2 * 20001)
    d2h = 8 * (6 * 200 * 13 + 3 + 1 + 10)
    e2e = {"value": total * args.steps / e2e_wall, "unit": "superphotons/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_wall / args.steps,
           "rank0_run_ms": 1e3 * t_run / args.steps,
           "recorded": int(e2e_stats["recorded"]), "created": int(e2e_stats["created"]),
           "api": "HARMModel.run_simulation (create + run + allreduce + result + destroy through the C ABI)"}

    if comm:
        gm.nccl_comm_destroy(comm)
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    achieved = flops_all / world / (transport_ms * 1e-3) / 1e12  # per GPU
    traffic, traffic_src = ncu_traffic()
    out = {
        "metric": "superphotons/sec", "value": value, "unit": "superphotons/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": launches_all,
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE transport launch in the committed ncu
                     # --set full capture (a full-size generation), read from the summary file itself
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": "DFMA micro-benchmark in this process (grmonty_b200_fp64_peak); "
                                    "MEASURED_PEAKS.json has no FP64 entry",
                     "kernel": "transport_kernel<32,8> (csrc/gm_kernels.cuh: the fused per-lane loop, one launch per generation)",
                     "kernel_ms_per_step": transport_ms / args.steps},
        "device_ms_per_step": kernel_ms / args.steps, "init_s": init_s,
        "work_rates": {"tracked_photons_per_s": work_all[0] * args.steps / wall,
                       "geodesic_steps_per_s": work_all[1] * args.steps / wall,
                       "push_attempts_per_s": work_all[2] * args.steps / wall},
        "run": {"primaries_per_step": total, "recorded": res["recorded"], "scattered": res["scattered"],
                "max_tau_scatt": res["max_tau_scatt"], **{k: res["stats"][k] for k in
                ("n_tracked", "n_steps", "n_push_attempts", "n_interactions", "n_scatter_events", "n_generations",
                 "n_live_iterations", "n_slot_iterations")}},
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        out["cpu_baseline"] = run_reference_cpu(dump, args.photon_n, args.mass_unit, cores, ref_seconds)
        ctx_gpu = run_reference_gpu(dump, args.photon_n, args.mass_unit)
        if ctx_gpu:
            out["ref_gpu_baseline"] = ctx_gpu
    json_out.write(json.dumps(out) + "\n")
    json_out.flush()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
