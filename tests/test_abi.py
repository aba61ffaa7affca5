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
    """The eight Rcpp exports of BSFG_functions_c.cpp (lines 26,42,85,137,165,195,240,292), the north-star spellings, and
    the device versions of the two R-only steps."""
    import sparsefactormixedmodel_b200 as pkg
    for name in ("GSVD_2_c", "sample_means_c", "sample_Lambda_c", "sample_lambda_c", "sample_factors_scores_c",
                 "sample_px_factors_scores_c", "sample_h2s_discrete_c",
                 "sample_prec_discrete_conditional_c", "sample_F_a_c", "sample_delta_c", "update_k_c",
                 "fast_BSFG_sampler"):
        assert callable(getattr(pkg, name))


def test_r_glue_binds_every_reference_export():
    """r/bsfg_b200.R defines a closure under each Rcpp export name, and every .Call target exists in r/bsfg_glue.c (no R
    in this container: a text-level check that the two files and the header stay in step)."""
    rsrc = open(os.path.join(ROOT, "r", "bsfg_b200.R")).read()
    csrc = open(os.path.join(ROOT, "r", "bsfg_glue.c")).read()
    for name in ("GSVD_2_c", "sample_means_c", "sample_Lambda_c", "sample_lambda_c", "sample_factors_scores_c",
                 "sample_px_factors_scores_c", "sample_h2s_discrete_c", "sample_prec_discrete_conditional_c",
                 "sample_F_a_c", "sample_delta_c", "update_k_c", "bsfg_sweep_loop"):
        assert re.search(r"^%s\s*<-\s*(function|sample_Lambda_c$)" % re.escape(name), rsrc, flags=re.M), name
    targets = set(re.findall(r"\.Call\('(R_bsfg_[A-Za-z0-9_]+)'", rsrc))
    assert targets
    for t in targets:
        assert re.search(r"^SEXP %s\(" % re.escape(t), csrc, flags=re.M), f"{t} is .Call'ed but not defined in bsfg_glue.c"
    # argument counts: every .Call('R_x', a, b, ...) passes as many arguments as `SEXP R_x(SEXP ..., ...)` takes
    def top_level_args(text, start):
        depth, n, i = 0, 1, start
        while True:
            ch = text[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                if depth == 0:
                    return n
                depth -= 1
            elif ch == "," and depth == 0:
                n += 1
            i += 1
    for m in re.finditer(r"\.Call\('(R_bsfg_[A-Za-z0-9_]+)'", rsrc):
        n_call = top_level_args(rsrc, m.end()) - 1                       # minus the symbol name itself
        d = re.search(r"^SEXP %s\(([^)]*)\)" % re.escape(m.group(1)), csrc, flags=re.M)
        n_def = len([a for a in d.group(1).split(",") if a.strip() and a.strip() != "void"])
        assert n_call == n_def, f"{m.group(1)}: .Call passes {n_call} arguments, the C definition takes {n_def}"
    declared = set(_declared_symbols())
    for sym in set(re.findall(r"\b(bsfg_[A-Za-z0-9_]+)\s*\(", csrc)):
        assert sym in declared, f"bsfg_glue.c calls {sym}, which include/bsfg_b200.h does not declare"


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


def test_update_k_add_arm_draw_families_have_their_own_philox_sites():
    """Gamma draws address Philox blocks by element, normal draws by element pair, so two families on one
    site share blocks.  Every site word is unique and the add arm of update_k uses one site per vector family."""
    import re
    csrc = os.path.join(ROOT, "sparsefactormixedmodel_b200", "csrc")
    common = open(os.path.join(csrc, "common.cuh")).read()
    body = common[common.index("enum Site"):]
    body = body[:body.index("};")]
    sites = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"(SITE_\w+)\s*=\s*(\d+)", body))
    assert len(sites) >= 20 and len(set(sites.values())) == len(sites), sites
    stateless = open(os.path.join(csrc, "stateless.inl")).read()
    arm = stateless[stateless.index("one Philox site per draw family"):]
    arm = arm[:arm.index("addk_traits_kernel")]
    used = re.findall(r"make_var\(nullptr, s\.seed, (SITE_\w+),", arm)
    assert used == ["SITE_K_ADD_LP", "SITE_K_ADD_ZL", "SITE_K_ADD_ZFA", "SITE_K_ADD_ZF"], used
