"""Argument checking of the C ABI (include/grmonty_b200.h): grmonty_b200_create and grmonty_b200_init_tables reject
a bad config before they touch CUDA, so these run without a GPU.  The reference calls exit() on such errors
(utils.cuh:33-40); here they are return codes with a message."""
import ctypes as C

import numpy as np
import pytest

import cuda_grmonty_b200 as gm

EINVAL = -1


def _create(cfg):
    L = gm.lib()
    h = C.c_void_p()
    rc = L.grmonty_b200_create(C.byref(h), C.byref(cfg))
    msg = L.grmonty_b200_last_error(None).decode()
    if h:
        L.grmonty_b200_destroy(h)
    return rc, msg


def test_create_rejects_null_arguments(golden_model):
    L = gm.lib()
    cfg, keep = gm.make_config(golden_model)
    assert L.grmonty_b200_create(None, C.byref(cfg)) == EINVAL
    h = C.c_void_p()
    assert L.grmonty_b200_create(C.byref(h), None) == EINVAL and not h
    assert "null argument" in L.grmonty_b200_last_error(None).decode()


def test_create_rejects_another_abi(golden_model):
    cfg, keep = gm.make_config(golden_model)
    cfg.abi_version = 1
    rc, msg = _create(cfg)
    assert rc == EINVAL and "ABI mismatch" in msg
    cfg, keep = gm.make_config(golden_model)
    cfg.struct_size -= 8          # a caller compiled against a shorter struct
    rc, msg = _create(cfg)
    assert rc == EINVAL and "ABI mismatch" in msg


@pytest.mark.parametrize("field,value", [("n0", 1), ("n1", 0), ("world", 0), ("rank", -1), ("rank", 1)])
def test_create_rejects_bad_grid_or_sharding(golden_model, field, value):
    cfg, keep = gm.make_config(golden_model)
    setattr(cfg, field, value)
    rc, msg = _create(cfg)
    assert rc == EINVAL and "bad grid or sharding" in msg


@pytest.mark.parametrize("field", ["k_rho", "b_3", "geom_det", "hotcross", "f", "k2", "weight", "nint", "dndlnu_max"])
def test_create_rejects_missing_arrays(golden_model, field):
    cfg, keep = gm.make_config(golden_model)
    setattr(cfg, field, None)
    rc, msg = _create(cfg)
    assert rc == EINVAL and "null input array" in msg


def test_init_tables_rejects_bad_configs(golden_model):
    L = gm.lib()
    out = np.zeros(48 * 48)
    dp = out.ctypes.data_as(C.POINTER(C.c_double))
    assert L.grmonty_b200_init_tables(None, dp, None, None, None, None) == EINVAL
    cfg, keep = gm.make_config(golden_model)
    cfg.abi_version = 7
    assert L.grmonty_b200_init_tables(C.byref(cfg), dp, None, None, None, None) == EINVAL
    assert "ABI mismatch" in L.grmonty_b200_last_error(None).decode()
    cfg, keep = gm.make_config(golden_model)
    cfg.photon_n = 0.0
    assert L.grmonty_b200_init_tables(C.byref(cfg), dp, None, None, None, None) == EINVAL
    cfg, keep = gm.make_config(golden_model)
    cfg.f = None
    assert L.grmonty_b200_init_tables(C.byref(cfg), dp, None, None, None, None) == EINVAL
    assert "F(K) and K2" in L.grmonty_b200_last_error(None).decode()


def test_calls_on_a_null_context_fail_cleanly():
    L = gm.lib()
    assert L.grmonty_b200_run(None) == EINVAL
    assert L.grmonty_b200_run_range(None, 0, 10) == EINVAL
    L.grmonty_b200_destroy(None)          # a no-op, like free(NULL)
    L.grmonty_b200_trim_cache()           # nothing cached: nothing to do
