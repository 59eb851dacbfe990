"""Function-level parity on the bench grids (192 x 192 and 1024 x 1024) -- not only on the 48 x 48 golden model.

tests/golden/functions_grid_{192,1024}.npz were written by `oracle/make_golden.py functions_grid N` from the UNMODIFIED
reference CPU build loaded with the synthetic dump at that size: get_fluid_params (harm_model.cpp:595-671, with the
x_to_ij clamps of :1406-1434) at random points, cell edges, first / last half cells and the poles; get_fluid_zone and
init_zone on a sample of zones.  The grids themselves are not stored: the tests regenerate the same dump with
tools/make_harm_dump.py and read it through the product's own loader (HarmModel.read_file), so the loader, the
interleaved [n0][n1][8] device layout and the index arithmetic at these strides are all inside the comparison.

CPU (-m "not gpu"): the oracle against the 192^2 vectors.  GPU: the CUDA path against both.
"""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def rowerr(a, b):
    scale = np.max(np.abs(b), axis=1) + 1e-300
    return np.max(np.max(np.abs(a - b), axis=1) / scale)


@pytest.fixture(scope="module")
def models(tmp_path_factory):
    """model dicts of the n x n dumps, built through the product loader (cached per test session)"""
    import cuda_grmonty_b200 as gm
    from tools import make_harm_dump
    gm.build_host()
    cache = {}

    def get(n):
        if n not in cache:
            d = tmp_path_factory.mktemp(f"grid{n}")
            dump = str(d / f"dump{n}.txt")
            make_harm_dump.write_dump(dump, *make_harm_dump.make_dump(n0=n, n1=n))
            hm = gm.HarmModel(100000, 4e19)
            hm.read_file(dump)
            hm.init()
            cache[n] = hm.model_dict()
            os.remove(dump)
        return cache[n]
    return get


def check_model_scalars(model, g):
    for k in ("x_start1", "x_start2", "dx1", "dx2", "x_stop1", "x_stop2"):
        assert model[k] == pytest.approx(float(g[k]), rel=1e-15, abs=1e-300), k
    assert model["bias_norm"] == pytest.approx(float(g["bias_norm"]), rel=1e-12)   # sum over the dump's gdet column


def test_oracle_fluid_lookup_on_the_192_grid(models):
    from oracle import orc
    g = dict(np.load(os.path.join(GOLD, "functions_grid_192.npz")))
    model = models(192)
    check_model_scalars(model, g)
    M = orc.Model(model)
    n_out = 0
    for x, want in zip(g["fluid_x"], g["fluid_params"]):
        got = M.fluid_params(x)
        if want[0] == 0.0:
            n_out += 1
            assert got[0] == 0.0
        else:
            assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 1e-13
            assert abs(got[0] / want[0] - 1) < 1e-13 and abs(got[1] / want[1] - 1) < 1e-13
    assert n_out > 5
    for (i, j), wf, wi in zip(g["zone_ij"], g["zone_fluid"], g["zone_init"]):
        assert np.max(np.abs(M.fluid_zone(int(i), int(j)) - wf)) / np.max(np.abs(wf)) < 1e-13
        got = M.init_zone(int(i), int(j))
        assert np.allclose(got, wi, rtol=1e-12, atol=0)
    assert (g["zone_init"][:, 0] > 0).sum() > 50


@pytest.mark.gpu
@pytest.mark.parametrize("n", [192, 1024])
def test_cuda_fluid_lookup_and_zones_on_the_bench_grids(models, n):
    import cuda_grmonty_b200 as gm
    g = dict(np.load(os.path.join(GOLD, f"functions_grid_{n}.npz")))
    model = models(n)
    check_model_scalars(model, g)
    ctx = gm.Context(model, seed=123, test_exports=True)
    got = ctx.t_fluid_params(g["fluid_x"])
    want = g["fluid_params"]
    inside = want[:, 0] > 0
    assert inside.sum() > 400 and (~inside).sum() > 5
    assert np.all(got[~inside, 0] == 0.0)
    for sl in (slice(0, 3), slice(3, 7), slice(7, 11), slice(11, 15), slice(15, 19)):
        assert rowerr(got[inside][:, sl], want[inside][:, sl]) < 1e-11, sl
    assert np.max(np.abs(got[inside, 0] / want[inside, 0] - 1)) < 1e-11      # n_e on its own scale
    assert np.max(np.abs(got[inside, 1] / want[inside, 1] - 1)) < 1e-11      # theta_e
    # init_zone (photon count expectation and dn_max) on the sampled zones
    nz, dn_max, num = ctx.t_zones()
    zi, zj = g["zone_ij"][:, 0], g["zone_ij"][:, 1]
    wn, wd = g["zone_init"][:, 0], g["zone_init"][:, 1]
    m = wn > 0
    assert np.array_equal(nz[zi, zj] > 0, m)
    assert np.max(np.abs(nz[zi, zj][m] / wn[m] - 1)) < 1e-11
    assert np.max(np.abs(dn_max[zi, zj][m] / wd[m] - 1)) < 1e-11
    # stochastic rounding never moves a zone's count by more than one
    assert np.all(np.abs(num[zi, zj] - wn) <= 1.0)
    ctx.close()
