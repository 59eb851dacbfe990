"""Two-GPU test of the path's only collective through the C ABI: grmonty_b200_allreduce with a real ncclComm_t
(created through grmonty_b200_nccl_unique_id / grmonty_b200_nccl_comm_init_rank of the same ABI).  Skipped on boxes with fewer than two GPUs.
Checks: the two ranks' shares add up to the single-GPU run's primaries, after the all-reduce both ranks hold the same
spectrum and counters, and those equal the sum of the per-rank results read before the reduction."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import ctypes as C, glob, os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    import cuda_grmonty_b200 as gm
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)       # only to hand the NCCL unique id around
    # the communicator comes from the product ABI itself (grmonty_b200_nccl_unique_id / _comm_init_rank)
    uid = gm.nccl_unique_id() if rank == 0 else bytes(gm.NCCL_ID_BYTES)
    t = torch.frombuffer(bytearray(uid), dtype=torch.uint8).clone()
    dist.broadcast(t, 0)
    torch.cuda.set_device(rank)
    comm = gm.nccl_comm_init_rank(t.numpy().tobytes(), rank, world, rank)
    g = dict(np.load(os.path.join(%r, "tests", "golden", "functions_48.npz")))
    model = {k[6:]: (v.item() if v.ndim == 0 else v) for k, v in g.items() if k.startswith("model_")}
    ctx = gm.Context(model, seed=123, rank=rank, world=world, device=rank)
    ctx.run()
    mine = ctx.result()
    ctx.allreduce(comm)
    red = ctx.result()
    np.savez(out + f".{rank}.npz", mine_spec=mine["spectrum"], red_spec=red["spectrum"],
             mine_counts=np.array([mine["created"], mine["scattered"], mine["recorded"]], dtype=np.int64),
             red_counts=np.array([red["created"], red["scattered"], red["recorded"]], dtype=np.int64),
             mine_mt=np.array(mine["max_tau_scatt"]), red_mt=np.array(red["max_tau_scatt"]),
             total=np.array(ctx.total_primaries()))
    ctx.close()
    gm.nccl_comm_destroy(comm)
    dist.barrier()
    dist.destroy_process_group()
""") % (ROOT, ROOT)


def test_c_abi_allreduce_over_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = str(tmp_path / "r")
    port = str(29700 + os.getpid() % 200)
    procs = [subprocess.Popen([sys.executable, "-c", WORKER, str(r), "2", port, out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=150) == 0
    a, b = np.load(out + ".0.npz"), np.load(out + ".1.npz")
    assert a["mine_counts"][0] + b["mine_counts"][0] == a["total"]               # shares are complete and disjoint
    assert np.array_equal(a["red_counts"], a["mine_counts"] + b["mine_counts"])   # sum of the three counters
    assert np.array_equal(a["red_counts"], b["red_counts"])
    assert a["red_mt"] == b["red_mt"] == max(a["mine_mt"], b["mine_mt"])          # max of max_tau_scatt
    assert np.array_equal(a["red_spec"], b["red_spec"])                            # both ranks hold the same spectrum
    assert np.allclose(a["red_spec"], a["mine_spec"] + b["mine_spec"], rtol=1e-13, atol=0)
