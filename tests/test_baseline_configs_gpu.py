"""Injected-variate parity at the sizes BASELINE.json names as configs (C2, C3; C1 = the Sim_1 golden
fixture in test_golden_gpu.py; C4/C5 are the bench workloads, covered here through size-independent properties)."""
import numpy as np
import pytest

from conftest import relerr, rowwise_relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _parity(n, p, k, pedigree, seed, iters=1, gram="default"):
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    setup = I.make_synthetic(n, p, max(2, k // 2), pedigree=pedigree, seed=seed)
    pri = I.default_priors(k_init=k); rp = I.default_run_parameters()
    st = I.sampler_init(setup, pri, rp, seed=seed + 1, lists="stable")
    d, rv, pri, cs = st["data_matrices"], st["run_variables"], st["priors"], st["current_state"]
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=k + 8, gram=gram)
    try:
        rng = np.random.default_rng(seed + 2)
        ora = cs
        for it in range(1, iters + 1):
            v = I.draw_variates(rng, n, p, k, n, 1 + n, pri)
            sw.set_state(ora)
            ora, aux = O.gibbs_iteration(d, rv, pri, rp, ora, it, v)
            sw.sweep_injected(v)
            dev = sw.get_state()
            assert rowwise_relerr(dev["Lambda"], ora["Lambda"]) < TOL
            assert rowwise_relerr(dev["F"], ora["F"]) < TOL
            assert rowwise_relerr(dev["F_a"].T, ora["F_a"].T) < TOL
            assert np.array_equal(dev["F_h2"], np.asarray(ora["F_h2"]).ravel())
            assert relerr(dev["B"], ora["B"].reshape(dev["B"].shape)) < TOL
            assert rowwise_relerr(dev["E_a"].T, ora["E_a"].T) < TOL
            for f in ("Lambda_prec", "resid_Y_prec", "E_a_prec", "delta", "tauh"):
                assert relerr(dev[f], ora[f]) < TOL, f
            a = sw.aux()
            assert relerr(a["resid_Y_prec_rate"], aux["resid_Y_prec_rate"]) < TOL
            assert relerr(a["E_a_prec_rate"], aux["E_a_prec_rate"]) < TOL
        assert sw.gram_stats()["fallbacks"] == 0
    finally:
        sw.close()


def test_C2_synthetic_n1000_p1000_k20():
    """BASELINE.json configs[1]: 'synthetic n=1000 individuals, p=1000 traits, k=20 factors, 1 GPU, injected-variate
    equivalence vs Rcpp' (half-sib A with 2 distinct eigenvalues, like Sim_1)."""
    _parity(1000, 1000, 20, "halfsib", 201, iters=2)


def test_C3_ayroles_shaped_n200_p10000_k50():
    """BASELINE.json configs[2]: 'n=200 lines, p=10000 genes, k=50, random pedigree A, 1 GPU'."""
    _parity(200, 10000, 50, "random", 301)


def test_C4_shape_properties_k100():
    """C4's factor count (k = 100, the 4-warp Cholesky with the right-hand-side row in the 13th block) at a size the
    oracle finishes in seconds, both Gram evaluations."""
    _parity(300, 400, 100, "random", 401, gram="expsum")
    _parity(300, 400, 100, "random", 401, gram="exact")


def test_C4_full_size_one_iteration_n5000_p20000_k100():
    """BASELINE.json configs[3] / the `metric` workload at FULL size: one injected-variate Gibbs iteration of the device path
    (both weighted-Gram evaluations: the 256-node exponential sum that is the default, and the exact reference order) against
    the compiled reference (`oracle/_ref`: sample_Lambda_c over 20000 traits with n = 5000-term sums, the dense Z_1 E_a
    products and the p x p `t(E_a) Ainv E_a` of fast_BSFG_sampler.R:245) -- about a minute of host time.  The problem is
    bench.py's own (random 3-generation pedigree, prior-drawn state: tauh spans many decades).  Row-wise 1e-9, exact h2
    indices, no fallback of the exponential sum."""
    import bench
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k = bench.WORKLOADS["C4"]
    d, rv, pri, rp, cs = bench.make_problem(n, p, k, device="cuda:0")
    v = I.draw_variates(np.random.default_rng(4004), n, p, k, n, rv["b"] + n, pri)
    ora, aux = O.gibbs_iteration(d, rv, pri, rp, cs, 1, v)
    for gram in ("expsum", "exact"):
        sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=k, gram=gram)
        try:
            sw.set_state(cs)
            sw.sweep_injected(v)
            dev = sw.get_state()
            assert rowwise_relerr(dev["Lambda"], ora["Lambda"]) < TOL, gram
            assert rowwise_relerr(dev["F"], ora["F"]) < TOL, gram
            assert rowwise_relerr(dev["F_a"].T, ora["F_a"].T) < TOL, gram
            assert np.array_equal(dev["F_h2"], np.asarray(ora["F_h2"]).ravel()), gram
            assert relerr(dev["B"], ora["B"].reshape(dev["B"].shape)) < TOL, gram
            assert rowwise_relerr(dev["E_a"].T, ora["E_a"].T) < TOL, gram
            for f in ("Lambda_prec", "resid_Y_prec", "E_a_prec", "delta", "tauh"):
                assert relerr(dev[f], ora[f]) < TOL, (gram, f)
            a = sw.aux()
            assert relerr(a["resid_Y_prec_rate"], aux["resid_Y_prec_rate"]) < TOL, gram
            assert relerr(a["E_a_prec_rate"], aux["E_a_prec_rate"]) < TOL, gram
            gs = sw.gram_stats()
            assert gs["mode"] == gram and gs["fallbacks"] == 0
        finally:
            sw.close()


def test_C5_factor_count_k200():
    """C5's factor count (k = 200: the 8-warp Cholesky with 26 block rows, 190 KB of tiles per system) at a size the
    reference finishes in seconds (n = 2000, p = 3000), both Gram evaluations."""
    _parity(2000, 3000, 200, "random", 501, gram="expsum")
    _parity(2000, 3000, 200, "random", 501, gram="exact")


def test_factor_count_beyond_shared_memory_k300():
    """k = 300 > 216: the per-trait systems no longer fit shared memory (tiles in the L2 workspace, chol_blocked_big_kernel),
    the F row solves read the 300 x 300 factor from L2 (factor_rows_kernel<16>), k(k+1)/2 = 45150 Gram features.  The reference
    allows k < 2p (BSFG_functions.R:332): p = 200."""
    _parity(260, 200, 300, "random", 601, gram="expsum")
    _parity(260, 200, 300, "random", 601, gram="exact")
