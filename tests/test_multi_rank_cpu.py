"""World-size-2 coverage of the N > 1 path on CPU (gloo): the photon sharding rule and the end-of-run reduction.

The CUDA kernels cannot run here, so the per-rank compute is the oracle (test infrastructure) driven with the SAME
sharding arguments the C ABI takes (rank, world: positions j with j % world == rank) -- what is under test is the
host-side contract: the ranks' shares are disjoint and complete, a sum-allreduce of the [6][200][13] spectrum and the
three counters plus a max-allreduce of max_tau_scatt (as the bit pattern of a non-negative double, the way bench.py
and grmonty_b200_allreduce do it) reproduces the single-rank result."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    from oracle import orc
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = dict(np.load(os.path.join(%r, "tests", "golden", "functions_48.npz")))
    model = {k[6:]: (v.item() if v.ndim == 0 else v) for k, v in g.items() if k.startswith("model_")}
    M = orc.Model(model, seed=123)
    last = 600
    M.run(0, last, rank, world, 1 << 20, 1 << 20, 0)   # one generation, no suspension: initial statistics throughout
    spec = torch.from_numpy(M.spectrum().copy())
    cnt = torch.tensor([int(M.m.n_created), int(M.m.acc_n_scatt), int(M.m.acc_n_recorded)], dtype=torch.int64)
    mt = torch.from_numpy(np.array([M.m.acc_max_tau_scatt], dtype=np.float64).view(np.int64).copy())
    mine = int(M.m.n_created)
    dist.all_reduce(spec, op=dist.ReduceOp.SUM)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dist.all_reduce(mt, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.savez(out, spec=spec.numpy(), cnt=cnt.numpy(), mt=mt.numpy().view(np.float64), mine=np.array(mine))
    dist.barrier()
    dist.destroy_process_group()
""") % (ROOT, ROOT)


def test_two_ranks_reduce_to_the_single_rank_result(tmp_path, golden_model):
    from oracle import orc
    out = str(tmp_path / "reduced.npz")
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, "-c", WORKER, str(r), "2", port, out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    red = np.load(out)
    M = orc.Model(golden_model, seed=123)
    M.run(0, 600, 0, 1, 1 << 20, 1 << 20, 0)
    assert red["cnt"].tolist() == [int(M.m.n_created), int(M.m.acc_n_scatt), int(M.m.acc_n_recorded)]
    assert red["cnt"][0] == 600 and red["mine"] == 300          # disjoint, complete, balanced shares
    assert red["mt"][0] == M.m.acc_max_tau_scatt                 # max over ranks through the bit pattern
    want = M.spectrum()
    assert np.array_equal(red["spec"][:, :, 2], want[:, :, 2])   # photon counts per bin: exact
    assert np.allclose(red["spec"], want, rtol=1e-12, atol=0)    # sums: up to the order of the additions


def test_shares_partition_any_range():
    """positions first..last-1 split over world ranks: disjoint, complete, sizes differ by at most one"""
    for world in (2, 3, 8):
        for first, last in ((0, 1), (0, 1000), (17, 4099), (5, 5)):
            shares = [[j for j in range(first, last) if j % world == r] for r in range(world)]
            allj = sorted(j for s in shares for j in s)
            assert allj == list(range(first, last))
            sizes = [len(s) for s in shares]
            assert max(sizes) - min(sizes) <= 1
            # the closed form used by grmonty_b200_run_range (gm_api.cu): first index >= lo congruent to rank
            for r in range(world):
                f0 = first + ((r - first % world) % world + world) % world
                count = (last - f0 + world - 1) // world if f0 < last else 0
                assert count == len(shares[r]) and (count == 0 or f0 == shares[r][0])
