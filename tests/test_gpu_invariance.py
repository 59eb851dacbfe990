"""What a run's results may and may not depend on.

* Launch geometry (threads per block, blocks per SM, and which of the transport kernels runs): nothing.  Photons draw
  from counter-based Philox streams keyed by their identity, the scattering-bias statistics are frozen within a
  generation, and a lineage's attempt budget counts its own attempts -- so every integer output is identical and the
  floating-point sums agree up to the order of the atomic additions.
* GPU count (`world`): statistically nothing, bit-wise something.  Rank r of `world` tracks the positions j = r (mod
  world) of the processing sequence with the same Philox key space, so the union over ranks is exactly the world = 1
  photon set (same primaries, same birth states).  The bias statistics, however, are PER RANK (the path's only
  collective is the end-of-run all-reduce, BASELINE.json north_star): each rank runs the schedule of a stand-alone
  run of its share, its scattering decisions see its own running maximum, and the scattered / recorded counts of the
  job therefore differ from the world = 1 run by Monte Carlo noise.  The spectrum is an unbiased estimate for any
  bias, hence for any world.  This file pins that documented behaviour.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def small_run(gm, model, last=20000, gen_cap=1 << 12, **kw):
    c = gm.Context(model, seed=77, gen0=64, gen_cap=gen_cap, **kw)
    c.run(0, last)
    r = c.result()
    c.close()
    return r


INTS = ("created", "recorded", "scattered")
WORK = ("n_tracked", "n_steps", "n_push_attempts", "n_interactions", "n_scatter_events", "n_generations")


@pytest.mark.parametrize("overlap", [2, 1, 3])
def test_results_do_not_depend_on_launch_geometry(golden_model, overlap):
    """overlap 2 (the default): one launch per generation (both kernels, every compiled geometry); 1 / 3: the pipelined
    scheduler (fused kernel, every compiled geometry).  Within a scheduler every integer output is identical."""
    import cuda_grmonty_b200 as gm
    base = None
    for kernel, threads, third in gm.KERNEL_VARIANTS:
        if overlap != gm.OVERLAP_OFF and kernel != gm.KERNEL_FUSED:
            continue
        kw = dict(slots_per_thread=third) if kernel == gm.KERNEL_WAVEFRONT else dict(blocks_per_sm=third)
        r = small_run(gm, golden_model, kernel=kernel, threads_per_block=threads, gen_overlap=overlap, **kw)
        assert r["scattered"] > 1000 and r["stats"]["n_scatter_events"] > 1000   # scattering, children and carry-over ran
        if base is None:
            base = r
            continue
        for k in INTS:
            assert r[k] == base[k], (kernel, threads, third, k)
        for k in WORK:
            assert r["stats"][k] == base["stats"][k], (kernel, threads, third, k)
        assert r["max_tau_scatt"] == base["max_tau_scatt"]
        assert np.array_equal(r["spectrum"][:, :, 2], base["spectrum"][:, :, 2])     # photons per bin
        assert np.array_equal(r["spectrum"][:, :, 3], base["spectrum"][:, :, 3])     # scatterings per bin
        assert np.allclose(r["spectrum"], base["spectrum"], rtol=1e-10, atol=0)      # sums: order of the atomics


@pytest.mark.parametrize("overlap", [2, 1, 3])
def test_queue_capacity_does_not_change_results(golden_model, overlap):
    """pool size (how many generations share a launch of the pipelined scheduler / how a generation is split into
    batches by the round-1 scheduler) is not an input of the physics either"""
    import cuda_grmonty_b200 as gm
    a = small_run(gm, golden_model, gen_overlap=overlap)
    for cap in (1 << 18, 1 << 17):
        b = small_run(gm, golden_model, queue_capacity=cap, gen_overlap=overlap)
        for k in INTS:
            assert a[k] == b[k], (cap, k)
        for k in WORK:
            assert a["stats"][k] == b["stats"][k], (cap, k)
        assert a["max_tau_scatt"] == b["max_tau_scatt"]
        assert np.array_equal(a["spectrum"][:, :, 2], b["spectrum"][:, :, 2])


@pytest.mark.parametrize("overlap", [2, 1])
def test_split_runs_continue_the_generation_clock(golden_model, overlap):
    """run(0, a) + run(a, b) on one context: same primaries as run(0, b); the second call starts from the statistics
    the first left (and its own generation boundaries), so the counts agree statistically"""
    import cuda_grmonty_b200 as gm
    c = gm.Context(golden_model, seed=77, gen0=64, gen_cap=1 << 12, gen_overlap=overlap)
    c.run(0, 8000)
    c.run(8000, 20000)
    r = c.result()
    c.close()
    one = small_run(gm, golden_model, gen_overlap=overlap)
    assert r["created"] == one["created"] == 20000
    assert abs(r["recorded"] / one["recorded"] - 1) < 0.05
    assert r["spectrum"][:, :, 2].sum() == r["recorded"]


def test_world_dependence_is_statistical_only(golden_model):
    import cuda_grmonty_b200 as gm
    last = 60000

    def run(rank, world, seed=77):
        c = gm.Context(golden_model, seed=seed, rank=rank, world=world)
        c.run(0, last)
        r = c.result()
        c.close()
        return r
    c = gm.Context(golden_model, seed=77)
    last = min(last, c.total_primaries())        # the 48 x 48 golden model emits ~32 k primaries at its photon_n
    c.close()
    one = run(0, 1)
    parts = [run(r, 3) for r in range(3)]
    # shares are disjoint and complete: the same primaries are tracked
    assert sum(p["created"] for p in parts) == one["created"] == last
    assert max(p["created"] for p in parts) - min(p["created"] for p in parts) <= 1
    # a rank's result is reproducible ...
    again = run(1, 3)
    for k in INTS:
        assert again[k] == parts[1][k]
    assert np.array_equal(again["spectrum"][:, :, 2], parts[1]["spectrum"][:, :, 2])
    # ... and the job's result agrees with the world = 1 run statistically, not bit-wise (per-rank bias statistics)
    tot = {k: sum(p[k] for p in parts) for k in INTS}
    spec = sum(p["spectrum"] for p in parts)
    lum1, lum3 = one["spectrum"][:, :, 1].sum(), spec[:, :, 1].sum()
    assert abs(lum3 / lum1 - 1) < 0.03                     # ~50 k recorded photons: 1 % noise each
    # The COUNTS are not bias-independent.  Each rank's statistics go through their own start-up ramp, so three ranks of
    # 10 k primaries each see three times the early, over-biased phase of one rank of 32 k: measured here +21 % recorded
    # superphotons (lighter ones: the luminosity above is the same).  The gap closes with the size of a rank's share
    # (profiles/r2_scaling.txt: 8 ranks x 1.6e7 primaries against 1 x 1.6e7); this pins that it is bounded and one-sided.
    assert 1.0 <= tot["recorded"] / one["recorded"] < 1.4
    assert 1.0 <= tot["scattered"] / one["scattered"] < 2.0


@pytest.mark.parametrize("kernel,overlap", [(1, 2), (1, 1), (1, 3), (2, 2)])
def test_checked_build_sees_no_access_outside_the_pool(golden_model, kernel, overlap):
    """compute-sanitizer is closed on the GPU pool, so the test library carries bounds checks of its own: every slot
    number that comes out of a queue entry, a ticket or the allocator is checked against the pool before it is used
    (csrc/gm_transport.cuh chk_slot).  A run with scattering, suspension, carry-over between windows (small pool) and
    the final drain must count no violation -- and give the results of the unchecked product library."""
    import cuda_grmonty_b200 as gm
    c = gm.Context(golden_model, seed=77, gen0=64, gen_cap=1 << 12, kernel=kernel, gen_overlap=overlap,
                   queue_capacity=1 << 17, test_exports=True)
    c.t_bounds_violations()
    c.run(0, 20000)
    r = c.result()
    assert c.t_bounds_violations() == 0
    c.close()
    ref = small_run(gm, golden_model, kernel=kernel, gen_overlap=overlap)
    for k in INTS:
        assert r[k] == ref[k], k
    assert r["stats"]["n_scatter_events"] > 1000 and r["stats"]["n_generations"] > 5


def test_pipelined_scheduler_without_early_starts_reproduces_the_default(golden_model):
    """gen_overlap = 1 starts only the generations at the size cap early; a run that has none (gen_cap above every
    generation) must give the integer results of the default scheduler: same photons, same statistics, same budget
    rule -- only the launches differ (one persistent launch per window instead of one per generation)"""
    import cuda_grmonty_b200 as gm
    a = small_run(gm, golden_model, gen_overlap=2, gen_cap=1 << 20)
    b = small_run(gm, golden_model, gen_overlap=1, gen_cap=1 << 20)
    for k in INTS:
        assert a[k] == b[k], k
    for k in WORK[:-1]:   # (n_generations counts launches of the scheduler, not physics)
        assert a["stats"][k] == b["stats"][k], k
    assert a["max_tau_scatt"] == b["max_tau_scatt"]
    assert np.array_equal(a["spectrum"][:, :, 2], b["spectrum"][:, :, 2])
