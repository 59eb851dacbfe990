"""pytest configuration: registers the `gpu` marker and shared fixtures.

`-m "not gpu"` tests run on CPU (oracle vs golden vectors, host logic, C-ABI symbol checks).
`-m gpu` tests are the parity tests proper: CUDA path through the C-ABI vs the oracle / golden vectors.
Nothing here reads /root/reference.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "functions_48.npz")))


@pytest.fixture(scope="session")
def golden_model(golden):
    """model dict (grids, tables, scalars) of the 48x48 golden case, as produced by the reference build"""
    d = {k[len("model_"):]: v for k, v in golden.items() if k.startswith("model_")}
    return {k: (v.item() if v.ndim == 0 else v) for k, v in d.items()}


@pytest.fixture(scope="session")
def orc_model(golden_model):
    from oracle import orc
    return orc.Model(golden_model)
