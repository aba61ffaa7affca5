"""Device init (bsfg_init_lists): the lists of fast_BSFG_sampler_init.R:263-325 computed on the GPU
must satisfy the reference's contract (the identities quoted in BSFG_functions_c.cpp:52-54, 98-99) -- U itself is not
unique -- and a sweep run on them must match the oracle run on the same lists."""
import numpy as np
import pytest

from conftest import relerr, rowwise_relerr

pytestmark = pytest.mark.gpu


def _setup(n, p, r2, pedigree, seed, rect=False):
    from oracle import init_oracle as I
    rng = np.random.default_rng(seed)
    setup = I.make_synthetic(n, p, 3, pedigree=pedigree, seed=seed, b=2)
    if rect:                                     # rectangular Z_1: n individuals in r = n/5 lines (Ayroles-shaped)
        r = n // 5
        Z1 = np.zeros((n, r)); Z1[np.arange(n), np.arange(n) // 5] = 1.0
        A = I.random_pedigree_A(r, rng)
        setup["Z_1"] = Z1; setup["A"] = A
    if r2:
        Z2 = np.zeros((n, r2)); Z2[np.arange(n), rng.integers(0, r2, n)] = 1.0
        setup["Z_2"] = Z2
    return setup


@pytest.mark.parametrize("pedigree,rect,r2", [("halfsib", False, 0), ("random", False, 7), ("random", True, 6)])
def test_init_lists_contract(pedigree, rect, r2):
    import sparsefactormixedmodel_b200 as B
    n, p = 60, 20
    setup = _setup(n, p, r2, pedigree, 41, rect)
    X, Z, A = setup["X"], setup["Z_1"], setup["A"]
    rv = B.init_lists(Z, A, X, setup.get("Z_2"))
    b, r = X.shape[1], Z.shape[1]
    Ainv = np.linalg.inv(A)
    assert relerr(rv["Ainv"], Ainv) < 1e-10
    a_, b_ = 0.7, 1.9
    l1, l2, l3 = rv["invert_aI_bZAZ"], rv["invert_aPXA_bDesignDesignT"], rv["invert_aZZt_Ainv"]
    assert np.all(np.diff(l1["s"]) <= 1e-12) and np.all(l1["s"] >= 0)          # descending, like svd()$d
    assert relerr(l1["U"].T @ l1["U"], np.eye(n)) < 1e-12
    got = (l1["U"] / (l1["s"] + a_ / b_)[None, :]) @ l1["U"].T / b_
    assert relerr(got, np.linalg.inv(a_ * np.eye(n) + b_ * Z @ A @ Z.T)) < 1e-10
    D = np.column_stack([X, Z])
    P = np.zeros((b + r, b + r)); P[:b, :b] = 1e-10 * np.eye(b); P[b:, b:] = Ainv
    assert relerr(l2["U"].T @ P @ l2["U"], np.diag(l2["s1"])) < 1e-10
    assert relerr(l2["U"].T @ (D.T @ D) @ l2["U"], np.diag(l2["s2"])) < 1e-10
    assert relerr(l2["s1"] + l2["s2"], np.ones(b + r)) < 1e-13
    assert relerr(l2["Design_U"], D @ l2["U"]) < 1e-12
    got = (l3["U"] / (a_ * l3["s1"] + b_ * l3["s2"])[None, :]) @ l3["U"].T
    assert relerr(got, np.linalg.inv(a_ * Z.T @ Z + b_ * Ainv)) < 1e-10
    if r2:
        l4 = rv["invert_aPXA_bDesignDesignT_rand2"]
        Z2 = setup["Z_2"]
        got = (l4["U"] / (a_ * l4["s1"] + b_ * l4["s2"])[None, :]) @ l4["U"].T
        assert relerr(got, np.linalg.inv(a_ * np.eye(r2) + b_ * Z2.T @ Z2)) < 1e-10
        assert relerr(l4["Design_U"], Z2 @ l4["U"]) < 1e-12


def test_sweep_on_device_made_lists_matches_oracle():
    import sparsefactormixedmodel_b200 as B
    from oracle import checker as O, init_oracle as I
    n, p, k, r2 = 50, 31, 5, 4
    setup = _setup(n, p, r2, "random", 43)
    pri = I.default_priors(k_init=k); rp = I.default_run_parameters()
    st = I.sampler_init(setup, pri, rp, seed=6, lists="stable")
    d, pri, cs = st["data_matrices"], st["priors"], st["current_state"]
    rv = B.init_lists(d["Z_1"], setup["A"], d["X"], d["Z_2"])                 # device lists replace the host ones
    sw = B.BSFGSweep(d, rv, pri, rp, k_init=k, k_max=8)
    try:
        assert sw.mode()[0] == "diag"                                          # the identities hold to 1e-11: fast forms usable
        v = I.draw_variates(np.random.default_rng(7), n, p, k, n, 2 + n, pri, r2=r2)
        sw.set_state(cs)
        ora, _ = O.gibbs_iteration(d, rv, pri, rp, cs, 1, v)
        sw.sweep_injected(v)
        dev = sw.get_state()
        assert rowwise_relerr(dev["Lambda"], ora["Lambda"]) < 1e-9
        assert rowwise_relerr(dev["F"], ora["F"]) < 1e-9
        assert rowwise_relerr(dev["E_a"].T, ora["E_a"].T) < 1e-9
        assert rowwise_relerr(dev["W"].T, ora["W"].T) < 1e-9
        assert relerr(dev["E_a_prec"], ora["E_a_prec"]) < 1e-9
    finally:
        sw.close()
