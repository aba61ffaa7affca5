"""Device RNG (Philox4x32-10, csrc/common.cuh) and chain-level checks.  R's rgamma/rnorm streams cannot be replayed 1:1
(SURVEY.md section 7), so the production path is validated the way the survey prescribes: the (shape, rate) parameters
are parity-tested elsewhere, the DRAWS are tested distributionally here (Kolmogorov-Smirnov on back-transformed
variates), and a short chain must recover the covariance structure it was simulated from."""
import numpy as np
import pytest
from scipy import stats

from conftest import relerr

pytestmark = pytest.mark.gpu


def _build(n, p, k, seed, k_true=3):
    from oracle import init_oracle as I
    setup = I.make_synthetic(n, p, k_true, pedigree="halfsib", seed=seed)
    pri = I.default_priors(k_init=k); rp = I.default_run_parameters()
    st = I.sampler_init(setup, pri, rp, seed=seed + 1, lists="stable")
    return setup, st


def test_philox_normals_are_standard_normal():
    """theta = mean + z / sqrt(d) (cpp:75-78): back out z from a Philox draw of sample_means_c; 6000 values, KS vs N(0,1),
    and no correlation between the cosine / sine branches of one Box-Muller block (elements 2q, 2q+1)."""
    import sparsefactormixedmodel_b200 as B
    from oracle import checker as O
    setup, st = _build(100, 60, 4, 3)
    d, rv, cs = st["data_matrices"], st["run_variables"], st["current_state"]
    L = rv["invert_aPXA_bDesignDesignT"]
    Yt = d["Y"] - cs["F"] @ cs["Lambda"].T
    mean = O.sample_means_c(Yt, cs["resid_Y_prec"], cs["E_a_prec"], L, np.zeros((101, 60)))
    a = B.sample_means_c(Yt, cs["resid_Y_prec"], cs["E_a_prec"], L, seed=2024)
    dmat = np.outer(L["s1"], cs["E_a_prec"]) + np.outer(L["s2"], cs["resid_Y_prec"])
    z = (np.linalg.solve(L["U"], a - mean) * np.sqrt(dmat)).ravel(order="F")
    assert stats.kstest(z, "norm").pvalue > 1e-3
    assert abs(np.corrcoef(z[0:-1:2], z[1::2])[0, 1]) < 0.05
    assert abs(stats.skew(z)) < 0.1 and abs(stats.kurtosis(z)) < 0.2


def test_philox_gammas_follow_their_gamma_laws():
    """One iteration with device draws but aux on: precision * rate is a standard Gamma(shape, 1) variate
    (fast_BSFG_sampler.R:235-245): Lambda_prec ~ Gamma((df+1)/2), resid_Y_prec ~ Gamma(a + n/2), E_a_prec ~ Gamma(a + r/2)."""
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k = 60, 400, 6
    setup, st = _build(n, p, k, 5)
    d, rv, pri, rp, cs = (st["data_matrices"], st["run_variables"], st["priors"], st["run_parameters"], st["current_state"])
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=8, seed=77)
    try:
        sw.set_state(cs)
        sw.sweep_injected({})                       # no injected member: every draw comes from Philox, aux is recorded
        out, a = sw.get_state(), sw.aux()
        g = (out["Lambda_prec"] * a["Lambda_prec_rate"]).ravel()
        assert stats.kstest(g, "gamma", args=((pri["Lambda_df"] + 1) / 2,)).pvalue > 1e-3
        g = out["resid_Y_prec"] * a["resid_Y_prec_rate"]
        assert stats.kstest(g, "gamma", args=(pri["resid_Y_prec_shape"] + n / 2,)).pvalue > 1e-3
        g = out["E_a_prec"] * a["E_a_prec_rate"]
        assert stats.kstest(g, "gamma", args=(pri["E_a_prec_shape"] + n / 2,)).pvalue > 1e-3
    finally:
        sw.close()


def test_short_chain_recovers_the_simulated_factor_covariance():
    """Simulation recovery in the spirit of the reference's draw_simulation_diagnostics (plotting_diagnostics.R:22-149):
    data simulated from Lambda_true (generate_simulations_halfSib.R recipe); after burn-in the posterior mean of
    Lambda Lambda' must track Lambda_true Lambda_true' (factor scores have unit variance in the model)."""
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k = 250, 60, 8
    setup, st = _build(n, p, k, 11, k_true=4)
    d, rv, pri, rp, cs = (st["data_matrices"], st["run_variables"], st["priors"], st["run_parameters"], st["current_state"])
    Lt = setup["error_factor_Lambda"]
    true_cov = Lt @ Lt.T
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=12, seed=5)
    try:
        sw.set_state(cs)
        sw.sweep(250)
        acc = np.zeros((p, p)); rp_acc = np.zeros(p)
        for _ in range(30):
            sw.sweep(5)
            s_ = sw.get_state(with_E_a=False)
            acc += s_["Lambda"] @ s_["Lambda"].T / 30
            rp_acc += 1.0 / s_["resid_Y_prec"] / 30
        iu = np.triu_indices(p, 1)
        c = np.corrcoef(acc[iu], true_cov[iu])[0, 1]
        assert c > 0.9, c
        assert relerr(np.diag(acc), np.diag(true_cov)) < 0.5
        assert np.all(rp_acc > 0.02) and np.all(rp_acc < 2.0)        # residual variances near the simulated 0.2 (+ E_a part)
    finally:
        sw.close()
