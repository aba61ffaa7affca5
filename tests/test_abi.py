"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/bsfg_b200.h declares; compute calls fail loudly (no CPU fallback) when there is no GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bsfg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bsfg_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sparsefactormixedmodel_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/bsfg_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert lib.bsfg_version() >= 100


def test_reference_export_names_are_mirrored():
    """The eight Rcpp exports of BSFG_functions_c.cpp (lines 26,42,85,137,165,195,240,292) minus the
    init-only GSVD_2_c, plus the north-star spellings."""
    import sparsefactormixedmodel_b200 as pkg
    for name in ("sample_means_c", "sample_Lambda_c", "sample_lambda_c", "sample_factors_scores_c",
                 "sample_px_factors_scores_c", "sample_h2s_discrete_c",
                 "sample_prec_discrete_conditional_c", "sample_F_a_c", "sample_delta_c",
                 "fast_BSFG_sampler"):
        assert callable(getattr(pkg, name))


def test_no_cpu_fallback_without_gpu():
    from sparsefactormixedmodel_b200 import _lib, BSFGError
    import sparsefactormixedmodel_b200 as pkg
    if _lib.load().bsfg_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(BSFGError) as ei:
        pkg.dgemm_tn(np.ones((4, 4)), np.ones((4, 4)))
    assert ei.value.code == _lib.BSFG_ECUDA


def test_product_package_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "sparsefactormixedmodel_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
