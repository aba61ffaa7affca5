import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        from sparsefactormixedmodel_b200 import _lib
        return _lib.load().bsfg_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: the product has no CPU path.
    pass


@pytest.fixture(scope="session")
def small_problem():
    """A C1-shaped problem (half-sib A, Z_1 = I, X = 1) with state and lists from the init restatement."""
    from oracle import init_oracle as I
    n, p, k = 60, 40, 6
    setup = I.make_synthetic(n, p, 4, pedigree="halfsib", seed=7)
    pri = I.default_priors(k_init=k)
    rp = I.default_run_parameters()
    st = I.sampler_init(setup, pri, rp, seed=3, lists="reference")
    return dict(setup=setup, state=st, n=n, p=p, k=k, r=n, b=1)


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    denom = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / denom)


def rowwise_relerr(a, b):
    """max over rows of ||a_row - b_row||_inf / ||b_row||_inf (SURVEY.md section 8c tolerance form)."""
    a = np.atleast_2d(np.asarray(a, dtype=np.float64)); b = np.atleast_2d(np.asarray(b, dtype=np.float64))
    num = np.max(np.abs(a - b), axis=1)
    den = np.maximum(np.max(np.abs(b), axis=1), 1e-300)
    return float(np.max(num / den))
