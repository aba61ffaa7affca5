"""GPU parity against the committed golden fixtures (tests/golden/*.npz: the reference's own setup.mat inputs at
full size -> reference-route init lists -> outputs of the COMPILED REFERENCE, oracle/_ref, see
tests/golden/make_golden.py).  Nothing here runs a checker: stored inputs go through the C ABI, results are
compared with the reference's stored outputs.  Tolerance 1e-9 relative
(BASELINE.json north_star), exact h2 indices."""
import numpy as np
import pytest

from conftest import relerr, rowwise_relerr
from test_oracle import CASES, golden_problem

pytestmark = pytest.mark.gpu
TOL = 1e-9


def test_checker_is_the_compiled_reference():
    """The other `-m gpu` files compare the device with `oracle.checker`; on the GPU box that must be the reference
    library that travelled with the snapshot, not the NumPy restatement."""
    from oracle import checker
    assert checker.KIND == "reference"
    assert checker.sample_Lambda_c.__module__ == "oracle.ref_oracle"


@pytest.mark.parametrize("name", CASES)
def test_sweep_injected_matches_golden(name):
    from sparsefactormixedmodel_b200 import BSFGSweep
    g, d, rv, pri, rp = golden_problem(name)
    k = int(g["meta"]["k"])
    b = d["X"].shape[1]
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=k + 4)
    try:
        # the reference-route lists of the b >= 1, Z_1 = I fixtures are rank deficient:
        # the context must have detected that and chosen the direct forms
        mode, r1, r2 = sw.mode()
        for it in (1, 2):
            sw.set_state(g[f"state{it - 1}"])
            sw.sweep_injected(g[f"variates{it}"])
            dev, want = sw.get_state(), g[f"state{it}"]
            assert rowwise_relerr(dev["Lambda"], want["Lambda"]) < TOL, (mode, r1, r2)
            assert rowwise_relerr(dev["F"], want["F"]) < TOL
            assert rowwise_relerr(dev["F_a"].T, want["F_a"].T) < TOL
            assert np.array_equal(dev["F_h2"], want["F_h2"])
            assert relerr(dev["B"], want["B"].reshape(dev["B"].shape)) < TOL
            assert rowwise_relerr(dev["E_a"].T, want["E_a"].T) < TOL
            for f in ("Lambda_prec", "resid_Y_prec", "E_a_prec", "delta", "tauh"):
                assert relerr(dev[f], want[f]) < TOL, f
            if "Z_2" in d:
                assert rowwise_relerr(dev["W"].T, want["W"].T) < TOL
                assert relerr(dev["W_prec"], want["W_prec"]) < TOL
            assert relerr(dev["Plam"], want["Plam"]) < 10 * TOL
            a = sw.aux()
            for f, val in g[f"aux{it}"].items():
                got = a[f][:len(val)] if f == "delta_rate" else a[f]
                assert relerr(got, val) < TOL, f
                if f == "Y_imputed":
                    obs = ~np.isnan(d["Y"])
                    assert np.array_equal(got[obs], d["Y"][obs])
    finally:
        sw.close()


@pytest.mark.parametrize("name", CASES)
def test_stateless_dropins_match_golden(name):
    """Each Rcpp-named entry point on the stored inputs of iteration 1, chained like fast_BSFG_sampler.R:189-232."""
    import sparsefactormixedmodel_b200 as B
    g, d, rv, pri, rp = golden_problem(name)
    Y, X, Z = g["aux1"].get("Y_imputed", d["Y"]), d["X"], d["Z_1"]      # missing phenotypes: iteration 1's imputed Y
    st, v, want = g["state0"], g["variates1"], g["state1"]
    b, r = X.shape[1], Z.shape[1]
    H = rp["h2_divisions"]
    Yfull = Y
    if "Z_2" in d:
        # second sample_means_c call (R:214): W from the residual of the NEW B, E_a, Lambda and the OLD F
        Wn = B.sample_means_c(Y - X @ want["B"].reshape(b, -1) - Z @ want["E_a"] - st["F"] @ want["Lambda"].T, st["resid_Y_prec"],
                              st["W_prec"], rv["invert_aPXA_bDesignDesignT_rand2"], variates=v["z_W"])
        assert rowwise_relerr(np.asarray(Wn).T, want["W"].T) < TOL
        Y = Y - d["Z_2"] @ st["W"]                     # residuals of the Lambda / (B,E_a) steps carry the OLD W
    Lam = B.sample_Lambda_c(Y - X @ st["B"], st["F"], st["resid_Y_prec"], st["E_a_prec"], st["Plam"],
                            rv["invert_aI_bZAZ"], variates=v["z_Lambda"])
    assert rowwise_relerr(Lam, want["Lambda"]) < TOL
    loc = B.sample_means_c(Y - st["F"] @ want["Lambda"].T, st["resid_Y_prec"], st["E_a_prec"],
                           rv["invert_aPXA_bDesignDesignT"], variates=v["z_means"])
    assert relerr(loc[:b], want["B"].reshape(b, -1)) < TOL
    assert rowwise_relerr(loc[b:].T, want["E_a"].T) < TOL
    h2 = B.sample_h2s_discrete_c(st["F"], H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], variates=v["u_h2"])
    assert np.array_equal(np.asarray(h2).ravel(), want["F_h2"])
    Fa = B.sample_F_a_c(st["F"], Z, want["F_h2"], rv["invert_aZZt_Ainv"], variates=v["z_Fa"])
    assert rowwise_relerr(np.asarray(Fa).T, want["F_a"].T) < TOL
    if "Z_2" in d:
        Y = Yfull - d["Z_2"] @ want["W"]               # the F step sees the NEW W (R:230)
    F = B.sample_factors_scores_c(Y - X @ want["B"].reshape(b, -1) - Z @ want["E_a"], Z, want["Lambda"],
                                  st["resid_Y_prec"], want["F_a"], want["F_h2"], variates=v["z_F"])
    assert rowwise_relerr(F, want["F"]) < TOL
    Fpx = B.sample_px_factors_scores_c(Y - X @ want["B"].reshape(b, -1) - Z @ want["E_a"], Z, want["Lambda"],
                                       st["resid_Y_prec"], want["F_a"], 1.0 / (1.0 - want["F_h2"]),
                                       np.ones_like(want["F_h2"]), variates=v["z_F"])
    assert rowwise_relerr(Fpx, want["F"]) < TOL
    prec = np.asarray(B.sample_prec_discrete_conditional_c(Yfull, H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"],
                                                           st["resid_Y_prec"], variates=g["prec_cond"]["u"])).ravel()
    wp = g["prec_cond"]["Prec"]
    fin = np.isfinite(wp)
    assert np.array_equal(fin, np.isfinite(prec)) and relerr(prec[fin], wp[fin]) < TOL
