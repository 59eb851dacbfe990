#!/usr/bin/env python
"""BASELINE.json configs[2]: photon_n = 1e8 sharded over the GPUs of one box, end-of-run NCCL all-reduce through the
product ABI.  Launch: python -m torch.distributed.run --nproc-per-node N tools/gpu_configs2.py [photon_n]
Rank 0 prints one JSON line: counters (64-bit: ~1.6e9 primaries, ~2.3e9 tracked photons), bookkeeping identities,
luminosity against the reference ensemble at photon_n = 1e6 (tests/golden), rates."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import cuda_grmonty_b200 as gm
from tools import make_harm_dump

photon_n = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0e8
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p = "/tmp/gp_dump_192.txt"
if rank == 0 and not os.path.exists(p):
    make_harm_dump.write_dump(p + ".tmp", *make_harm_dump.make_dump(n0=192, n1=192))
    os.replace(p + ".tmp", p)
if world > 1:
    dist.barrier()
hm = gm.HarmModel(int(photon_n), 4e19)
hm.read_file(p)
hm.init()
ctx = gm.Context(hm.model_dict(), seed=123, rank=rank, world=world, device=local)
comm = None
if world > 1:
    uid = torch.frombuffer(bytearray(gm.nccl_unique_id() if rank == 0 else bytes(gm.NCCL_ID_BYTES)),
                           dtype=torch.uint8).to(f"cuda:{local}")
    dist.broadcast(uid, 0)
    comm = gm.nccl_comm_init_rank(uid.cpu().numpy().tobytes(), rank, world, local)
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
ctx.run()
mine = ctx.result()
t_run = time.perf_counter() - t0
if comm:
    ctx.allreduce(comm)
res = ctx.result()
torch.cuda.synchronize()
t_all = torch.tensor([time.perf_counter() - t0, t_run], dtype=torch.float64, device=f"cuda:{local}")
work = torch.tensor([float(mine["stats"][k]) for k in ("n_tracked", "n_steps", "n_push_attempts", "n_scatter_events")],
                    dtype=torch.float64, device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    dist.all_reduce(work, op=dist.ReduceOp.SUM)
if rank == 0:
    spec = res["spectrum"]
    parts = [dict(np.load(os.path.join(ROOT, "tests", "golden", f))) for f in
             ("spectrum_192_4e19_1e6.npz", "spectrum_192_4e19_1e6_more.npz", "spectrum_192_4e19_1e6_more2.npz",
              "spectrum_192_4e19_1e6_more3.npz")]
    ref_lum = np.concatenate([q["spec"][..., 1].sum(axis=(1, 2)) for q in parts])
    lum = float(spec[:, :, 1].sum())
    total = ctx.total_primaries()
    out = {"config": f"configs[2]: photon_n={photon_n:g} over {world} GPU(s), 192x192, mass_unit=4e19",
           "n_gpus": world, "created": res["created"], "total_primaries": total, "recorded": res["recorded"],
           "scattered": res["scattered"], "tracked": int(work[0].item()), "steps": int(work[1].item()),
           "push_attempts": int(work[2].item()), "scatter_events": int(work[3].item()),
           "created_exceeds_int32": res["created"] > 2 ** 31 - 1, "tracked_exceeds_int32": work[0].item() > 2 ** 31 - 1,
           "created_equals_total": res["created"] == total,
           "nph_sum_equals_recorded": int(round(spec[:, :, 2].sum())) == res["recorded"],
           "nscatt_sum_equals_scattered": int(round(spec[:, :, 3].sum())) == res["scattered"],
           "finite": bool(np.isfinite(spec).all()), "max_tau_scatt": res["max_tau_scatt"],
           "luminosity": lum, "ref_luminosity_mean_1e6": float(ref_lum.mean()),
           "ref_luminosity_rel_sd_1e6": float(ref_lum.std(ddof=1) / ref_lum.mean()), "n_ref_runs": int(len(ref_lum)),
           "luminosity_rel_diff": lum / float(ref_lum.mean()) - 1,
           "run_s_max_over_ranks": t_all[1].item(), "run_plus_allreduce_s": t_all[0].item(),
           "superphotons_per_s": total / t_all[0].item(),
           "scattered_per_created": res["scattered"] / res["created"], "recorded_per_created": res["recorded"] / res["created"]}
    out["ok"] = bool(out["created_equals_total"] and out["nph_sum_equals_recorded"] and out["nscatt_sum_equals_scattered"]
                     and out["finite"] and abs(out["luminosity_rel_diff"]) < 0.01)
    print(json.dumps(out), flush=True)
ctx.close()
if comm:
    gm.nccl_comm_destroy(comm)
if world > 1:
    dist.destroy_process_group()
