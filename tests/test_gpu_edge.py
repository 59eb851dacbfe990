"""Edge cases of the run API on the GPU: empty and split ranges, a model that emits nothing, reset, a context that
has not run yet, a device that does not exist.  (Argument validation that needs no GPU is in test_abi_validation.py.)"""
import numpy as np
import pytest

import cuda_grmonty_b200 as gm

pytestmark = pytest.mark.gpu
KW = dict(seed=123, gen0=1 << 10, gen_cap=1 << 12)


def _consistent(r):
    """integer bookkeeping that must hold for any run: every recorded photon sits in exactly one spectrum bin"""
    s = r["spectrum"]
    assert np.isfinite(s).all()
    assert s[:, :, 2].sum() == r["recorded"]          # nph
    assert s[:, :, 3].sum() == r["scattered"]         # nscatt of recorded photons
    assert (s[:, :, 0] >= 0).all() and (s[:, :, 1] >= 0).all()


def test_result_before_any_run_is_empty(golden_model):
    c = gm.Context(golden_model, **KW)
    r = c.result()
    assert r["created"] == r["recorded"] == r["scattered"] == 0 and not r["spectrum"].any()
    assert r["max_tau_scatt"] == pytest.approx(float(golden_model["max_tau_scatt0"]), rel=1e-15)
    c.close()


def test_empty_and_clipped_ranges(golden_model):
    c = gm.Context(golden_model, **KW)
    total = c.total_primaries()
    assert total > 5000
    c.run(0, 0)
    c.run(700, 700)
    c.run(900, 300)                                   # first > last: nothing
    r = c.result()
    assert r["created"] == 0 and r["recorded"] == 0 and not r["spectrum"].any()
    c.run(total - 50, total + 10 ** 9)                # clipped to the run: the last 50 positions
    r = c.result()
    assert r["created"] == 50
    _consistent(r)
    c.close()


def test_split_ranges_accumulate(golden_model):
    c = gm.Context(golden_model, **KW)
    c.run(0, 1500)
    mid = c.result()
    c.run(1500, 4000)
    r = c.result()
    assert mid["created"] == 1500 and r["created"] == 4000
    assert r["recorded"] > mid["recorded"] > 0
    _consistent(mid)
    _consistent(r)
    # the same positions in one call create the same primaries (streams are keyed by the primary index); the
    # scattering-bias history differs between the two schedules, so only the bookkeeping is compared
    d = gm.Context(golden_model, **KW)
    d.run(0, 4000)
    one = d.result()
    assert one["created"] == 4000 and one["recorded"] > 0
    _consistent(one)
    c.close()
    d.close()


def test_reset_gives_the_same_run_again(golden_model):
    c = gm.Context(golden_model, **KW)
    c.run(0, 3000)
    a = c.result()
    c.reset()
    z = c.result()
    assert z["created"] == 0 and not z["spectrum"].any()
    c.run(0, 3000)
    b = c.result()
    assert (a["created"], a["recorded"], a["scattered"]) == (b["created"], b["recorded"], b["scattered"])
    assert np.array_equal(a["spectrum"][:, :, 2], b["spectrum"][:, :, 2])
    assert np.allclose(a["spectrum"], b["spectrum"], rtol=1e-12, atol=0)   # FP64 atomics: order of the sums differs
    c.close()


def test_model_that_emits_nothing(golden_model):
    cold = dict(golden_model)
    cold["u"] = np.asarray(golden_model["u"]) * 1e-9      # theta_e << theta_e_min = 0.3 everywhere
    c = gm.Context(cold, **KW)
    assert c.total_primaries() == 0
    c.run()
    r = c.result()
    assert r["created"] == r["recorded"] == r["scattered"] == 0 and not r["spectrum"].any()
    assert r["stats"]["n_tracked"] == 0
    c.close()


def test_unknown_device_is_an_error_not_a_fallback(golden_model):
    with pytest.raises(gm.GrmontyError, match="out of range"):
        gm.Context(golden_model, device=77)
