"""CPU checks of the oracle itself (no GPU, no library compute calls).

The reference has no tests and cannot be run here (no R/Rcpp/Armadillo), so the oracle is pinned by
  (1) the reference author's twin convention (README.md:37-38): the restatement of the `_c` functions
      (oracle/bsfg_oracle.py, following BSFG_functions_c.cpp) must agree with the independent restatement of
      the R twins (oracle/bsfg_oracle_rtwin.py, following BSFG_functions.R:23-269) under shared variates;
  (2) closed-form identities: every Gaussian conditional recomputed by dense n x n algebra from the model
      statement, without the diagonalisation lists;
  (3) the committed golden fixtures built from the reference's own setup.mat inputs
      (tests/golden/make_golden.py): re-running the oracle on the stored inputs reproduces the stored outputs.
"""
import os

import numpy as np
import pytest

from conftest import relerr, rowwise_relerr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("sim1", "example", "ayroles")


def load_golden(name):
    z = np.load(os.path.join(GOLD, f"{name}.npz"))
    out = {}
    for key in z.files:
        cur = out
        parts = key.split(".")
        for part in parts[:-1]:
            cur = cur.setdefault(part, {})
        val = z[key]
        cur[parts[-1]] = val.item() if val.ndim == 0 else val
    return out


def golden_problem(name):
    from oracle import init_oracle as I
    g = load_golden(name)
    k = int(g["meta"]["k"])
    pri = I.default_priors(k_init=k)
    b = g["data"]["X"].shape[1]
    pri["b_X_prec"] = 1e-10 * np.eye(b)
    rp = I.default_run_parameters()
    rv = dict(g["rv"])
    n, p = g["data"]["Y"].shape
    rv.update(n=n, p=p, r=g["data"]["Z_1"].shape[1], b=b)
    if "Z_2" in g["data"]:
        rv["A_2_inv"] = np.eye(g["data"]["Z_2"].shape[1])            # fast_BSFG_sampler_init.R:264
    return g, g["data"], rv, pri, rp


STATE_FIELDS = ("Lambda", "F", "F_a", "F_h2", "B", "E_a", "Lambda_prec", "resid_Y_prec", "E_a_prec", "delta",
                "tauh", "Plam")


# ------------------------------------------------------------------------------------------------
# (3) golden fixtures
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name):
    from oracle import bsfg_oracle as O
    g, d, rv, pri, rp = golden_problem(name)
    cur = g["state0"]
    for it in (1, 2):
        nxt, aux = O.gibbs_iteration(d, rv, pri, rp, cur, it, g[f"variates{it}"])
        want = g[f"state{it}"]
        for f in STATE_FIELDS + (("W", "W_prec") if "Z_2" in d else ()):
            assert relerr(np.asarray(nxt[f]).ravel(), np.asarray(want[f]).ravel()) < 1e-10, (name, it, f)
        for f, val in g[f"aux{it}"].items():
            assert relerr(aux[f], val) < 1e-10, (name, it, f)
        cur = nxt
    Yc = g["aux1"].get("Y_imputed", d["Y"])
    prec = O.sample_prec_discrete_conditional_c(Yc, rp["h2_divisions"], pri["h2_priors_factors"],
                                                rv["invert_aI_bZAZ"], g["state0"]["resid_Y_prec"], g["prec_cond"]["u"])
    fin = np.isfinite(g["prec_cond"]["Prec"])
    assert np.array_equal(fin, np.isfinite(prec))
    assert relerr(prec[fin], g["prec_cond"]["Prec"][fin]) < 1e-10


@pytest.mark.parametrize("name", CASES)
def test_golden_lists_satisfy_their_contract(name):
    """inv(a P + b D'D) = U diag(1/(a s1 + b s2)) U' (BSFG_functions_c.cpp:52-54) and the aI+bZAZ' analogue
    (cpp:98-99) on the lists the reference's init route produced for the reference's own fixtures."""
    g, d, rv, pri, rp = golden_problem(name)
    X, Z = d["X"], d["Z_1"]
    b = X.shape[1]
    r = Z.shape[1]
    A = np.linalg.inv(rv["Ainv"])
    l1, l2, l3 = rv["invert_aI_bZAZ"], rv["invert_aPXA_bDesignDesignT"], rv["invert_aZZt_Ainv"]
    a_, b_ = 0.7, 1.9
    n = Z.shape[0]
    M = a_ * np.eye(n) + b_ * Z @ A @ Z.T
    got = (l1["U"] / (l1["s"] + a_ / b_)[None, :]) @ l1["U"].T / b_
    assert relerr(got, np.linalg.inv(M)) < 1e-8
    D = np.column_stack([X, Z])
    P = np.zeros((b + r, b + r)); P[:b, :b] = 1e-10 * np.eye(b); P[b:, b:] = rv["Ainv"]
    got = (l2["U"] / (a_ * l2["s1"] + b_ * l2["s2"])[None, :]) @ l2["U"].T
    assert relerr(got, np.linalg.inv(a_ * P + b_ * D.T @ D)) < 1e-6    # reference init is only ~1e-10/cond accurate
    assert relerr(l2["Design_U"], D @ l2["U"]) < 1e-12
    got = (l3["U"] / (a_ * l3["s1"] + b_ * l3["s2"])[None, :]) @ l3["U"].T
    assert relerr(got, np.linalg.inv(a_ * Z.T @ Z + b_ * rv["Ainv"])) < 1e-8


@pytest.mark.parametrize("name", ("sim1", "example"))
@pytest.mark.parametrize("trait_threads", (1, 3))
def test_cpp_oracle_matches_numpy_oracle(name, trait_threads):
    """oracle/c/bsfg_ref.cpp (the compiled CPU baseline of bench.py: reference formulation against SciPy's OpenBLAS) is a
    third restatement; it must agree with the NumPy one on the reference's own fixtures, in the reference's serial
    order and with the per-trait loops under OpenMP."""
    from oracle import bsfg_oracle as O, c_oracle as CO
    g, d, rv, pri, rp = golden_problem(name)
    st, v = g["state0"], g["variates1"]
    want, aux = O.gibbs_iteration(d, rv, pri, rp, st, 1, v, with_update_k=False)
    got, aux_c = CO.gibbs_iteration_c(d, rv, pri, rp, st, v, trait_threads=trait_threads)
    for f in STATE_FIELDS:
        assert relerr(np.asarray(got[f]).ravel(), np.asarray(want[f]).ravel()) < 1e-10, f
    assert relerr(aux_c["resid_Y_prec_rate"], aux["resid_Y_prec_rate"]) < 1e-10
    assert relerr(aux_c["E_a_prec_rate"], aux["E_a_prec_rate"]) < 1e-10


# ------------------------------------------------------------------------------------------------
# (0) THE PIN: NumPy oracle == the reference itself (BSFG_functions_c.cpp compiled unmodified, oracle/_ref)
# ------------------------------------------------------------------------------------------------
def _ref():
    from oracle import ref_oracle as RF
    if not RF.available():
        pytest.skip("oracle/_ref not built and /root/reference not present")
    return RF


def _all_exports_agree(RF, O, Y, X, Z, st, v, rv, pri, H, u_prec, tol=1e-12):
    """Every Rcpp export, chained like one sweep (fast_BSFG_sampler.R:189-232), reference library vs NumPy oracle."""
    b = X.shape[1]
    B = np.asarray(st["B"]).reshape(b, Y.shape[1])
    Yt = Y - X @ B
    a = RF.sample_Lambda_c(Yt, st["F"], st["resid_Y_prec"], st["E_a_prec"], st["Plam"], rv["invert_aI_bZAZ"], v["z_Lambda"])
    o = O.sample_Lambda_c(Yt, st["F"], st["resid_Y_prec"], st["E_a_prec"], st["Plam"], rv["invert_aI_bZAZ"], v["z_Lambda"])
    assert rowwise_relerr(o, a) < 1e-10        # Q_j has cond ~ 1e10 at prior-drawn tauh: two LAPACK routes agree to ~1e-12
    Lam = a
    Yt = Y - st["F"] @ Lam.T
    a = RF.sample_means_c(Yt, st["resid_Y_prec"], st["E_a_prec"], rv["invert_aPXA_bDesignDesignT"], v["z_means"])
    o = O.sample_means_c(Yt, st["resid_Y_prec"], st["E_a_prec"], rv["invert_aPXA_bDesignDesignT"], v["z_means"])
    assert relerr(o, a) < tol
    loc = a
    a = RF.sample_h2s_discrete_c(st["F"], H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], v["u_h2"])
    o = O.sample_h2s_discrete_c(st["F"], H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], v["u_h2"])
    assert np.array_equal(o, a)
    h2 = a
    a = RF.sample_F_a_c(st["F"], Z, h2, rv["invert_aZZt_Ainv"], v["z_Fa"])
    o = O.sample_F_a_c(st["F"], Z, h2, rv["invert_aZZt_Ainv"], v["z_Fa"])
    assert relerr(o, a) < tol
    Fa = a
    Yt = Y - X @ loc[:b] - Z @ loc[b:]
    a = RF.sample_factors_scores_c(Yt, Z, Lam, st["resid_Y_prec"], Fa, h2, v["z_F"])
    o = O.sample_factors_scores_c(Yt, Z, Lam, st["resid_Y_prec"], Fa, h2, v["z_F"])
    assert rowwise_relerr(o, a) < 1e-10
    tau_e = 1.0 / (1.0 - np.minimum(h2, 0.98)) * 1.3
    px = np.linspace(0.5, 2.0, Lam.shape[1])
    a = RF.sample_px_factors_scores_c(Yt, Z, Lam, st["resid_Y_prec"], Fa, tau_e, px, v["z_F"])
    o = O.sample_px_factors_scores_c(Yt, Z, Lam, st["resid_Y_prec"], Fa, tau_e, px, v["z_F"])
    assert rowwise_relerr(o, a) < 1e-10
    a = RF.sample_prec_discrete_conditional_c(Y, H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], st["resid_Y_prec"], u_prec)
    o = O.sample_prec_discrete_conditional_c(Y, H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], st["resid_Y_prec"], u_prec)
    fin = np.isfinite(a)
    assert np.array_equal(fin, np.isfinite(o)) and relerr(o[fin], a[fin]) < tol


@pytest.mark.parametrize("name", CASES)
def test_numpy_oracle_equals_compiled_reference_on_fixtures(name):
    from oracle import bsfg_oracle as O
    RF = _ref()
    g, d, rv, pri, rp, st, v = _twin_inputs(name)
    Y = g["aux1"].get("Y_imputed", d["Y"])
    if "Z_2" in d:
        Y = Y - d["Z_2"] @ st["W"]
    _all_exports_agree(RF, O, Y, d["X"], d["Z_1"], st, v, rv, pri, rp["h2_divisions"], g["prec_cond"]["u"])
    got, want = O.GSVD_2_c(g["gsvd"]["A"], g["gsvd"]["B"]), RF.GSVD_2_c(g["gsvd"]["A"], g["gsvd"]["B"])
    # singular vectors are defined up to sign (and up to rotation inside repeated singular values): compare what
    # the init consumes (fast_BSFG_sampler_init.R:290-294): C^2, S^2 and the two products X C^2 X', X S^2 X'
    for kk in ("C", "S"):
        assert relerr(np.diag(got[kk]) ** 2, np.diag(want[kk]) ** 2) < 1e-9
    for kk in ("C", "S"):
        assert relerr(got["X"] @ got[kk] ** 2 @ got["X"].T, want["X"] @ want[kk] ** 2 @ want["X"].T) < 1e-9


@pytest.mark.parametrize("n,p,k,b,pedigree,rect", [(20, 1, 2, 1, "halfsib", False), (30, 13, 7, 0, "random", False),
                                                    (24, 9, 3, 2, "random", True), (45, 17, 5, 1, "halfsib", True),
                                                    (120, 64, 24, 1, "random", False)])
def test_numpy_oracle_equals_compiled_reference_on_random_problems(n, p, k, b, pedigree, rect):
    """Edge shapes: p = 1, k = 2, b = 0 / 2, odd n (integer n/2 of cpp:274), rectangular Z_1, n distinct eigenvalues."""
    from oracle import bsfg_oracle as O, init_oracle as I
    RF = _ref()
    rng = np.random.default_rng(2000 + n + p)
    setup = I.make_synthetic(n, p, max(2, k // 2), pedigree=pedigree, seed=n + p, b=b)
    if b == 0:
        setup["X"] = np.zeros((n, 0))
    if rect:
        r = n // 2
        Z = np.zeros((n, r)); Z[np.arange(n), np.arange(n) % r] = 1.0
        setup["Z_1"] = Z; setup["A"] = setup["A"][:r, :r]; setup["r"] = r
    pri = I.default_priors(k_init=k); rp = I.default_run_parameters()
    st = I.sampler_init(setup, pri, rp, seed=3, lists="stable")
    d, rv, pri, cs = st["data_matrices"], st["run_variables"], st["priors"], st["current_state"]
    r = d["Z_1"].shape[1]
    v = I.draw_variates(rng, n, p, k, r, b + r, pri)
    _all_exports_agree(RF, O, d["Y"], d["X"], d["Z_1"], cs, v, rv, pri, rp["h2_divisions"], rng.uniform(size=p))


def test_compiled_reference_raises_like_armadillo():
    """chol() of a non-PD Q_j throws in the reference (cpp:124); the flat wrapper reports it instead of aborting."""
    RF = _ref()
    n, p, k = 6, 2, 2
    rng = np.random.default_rng(1)
    lst = {"U": np.eye(n), "s": np.ones(n)}
    with pytest.raises(RuntimeError, match="chol"):
        RF.sample_Lambda_c(rng.standard_normal((n, p)), rng.standard_normal((n, k)), np.ones(p), np.ones(p),
                           -1e6 * np.ones((p, k)), lst, rng.standard_normal((k, p)))


def test_compiled_reference_is_the_unmodified_source():
    """oracle/_ref/SOURCE.sha256 records which file was compiled; where /root/reference is present it must match."""
    import hashlib
    RF = _ref()
    RF.load()
    rec = open(os.path.join(os.path.dirname(GOLD), os.pardir, "oracle", "_ref", "SOURCE.sha256")).read().split()[0]
    if os.path.exists(RF.REFERENCE_SOURCE):
        assert hashlib.sha256(open(RF.REFERENCE_SOURCE, "rb").read()).hexdigest() == rec
    assert len(rec) == 64


# ------------------------------------------------------------------------------------------------
# (1) `_c` restatement == R-twin restatement
# ------------------------------------------------------------------------------------------------
def _twin_inputs(name):
    g, d, rv, pri, rp = golden_problem(name)
    return g, d, rv, pri, rp, g["state0"], g["variates1"]


@pytest.mark.parametrize("name", CASES)
def test_twins_agree_on_reference_fixtures(name):
    from oracle import bsfg_oracle as O, bsfg_oracle_rtwin as R
    g, d, rv, pri, rp, st, v = _twin_inputs(name)
    Y, X, Z = g["aux1"].get("Y_imputed", d["Y"]), d["X"], d["Z_1"]      # missing phenotypes: the imputed Y of iteration 1
    if "Z_2" in d:
        Y = Y - d["Z_2"] @ st["W"]           # the twins are per-function: feed them the W-free residual
    H = rp["h2_divisions"]
    tol = 1e-9
    Yt = Y - X @ st["B"]
    a = O.sample_Lambda_c(Yt, st["F"], st["resid_Y_prec"], st["E_a_prec"], st["Plam"], rv["invert_aI_bZAZ"], v["z_Lambda"])
    b = R.sample_Lambda(Yt, st["F"], st["resid_Y_prec"], st["E_a_prec"], st["Plam"], rv["invert_aI_bZAZ"], v["z_Lambda"])
    assert rowwise_relerr(a, b) < tol
    Lam = a
    Yt = Y - st["F"] @ Lam.T
    a = O.sample_means_c(Yt, st["resid_Y_prec"], st["E_a_prec"], rv["invert_aPXA_bDesignDesignT"], v["z_means"])
    b = R.sample_means(Yt, st["resid_Y_prec"], st["E_a_prec"], rv["invert_aPXA_bDesignDesignT"], v["z_means"])
    assert relerr(a, b) < tol
    a = O.sample_h2s_discrete_c(st["F"], H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], v["u_h2"])
    b = R.sample_h2s_discrete(st["F"], H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], v["u_h2"])
    assert np.array_equal(a, b)
    h2 = a
    a = O.sample_F_a_c(st["F"], Z, h2, rv["invert_aZZt_Ainv"], v["z_Fa"])
    b = R.sample_F_a(st["F"], Z, h2, rv["invert_aZZt_Ainv"], v["z_Fa"])
    assert relerr(a, b) < tol
    Fa = a
    Yt = Y - X @ st["B"] - Z @ st["E_a"]
    a = O.sample_factors_scores_c(Yt, Z, Lam, st["resid_Y_prec"], Fa, h2, v["z_F"])
    b = R.sample_factors_scores(Yt, Z, Lam, st["resid_Y_prec"], Fa, h2, v["z_F"])
    assert rowwise_relerr(a, b) < 1e-8
    # the off-sweep export: twins agree for even n (the integer n/2 of cpp:274 only bites for odd n)
    n = Y.shape[0]
    u = g["prec_cond"]["u"]
    a = O.sample_prec_discrete_conditional_c(Y, H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], st["resid_Y_prec"], u)
    b = R.sample_prec_discrete_conditional(Y, H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], st["resid_Y_prec"], u)
    fin = np.isfinite(a)
    if n % 2 == 0:
        assert np.array_equal(fin, np.isfinite(b)) and relerr(a[fin], b[fin]) < tol
    else:
        # odd n (Example_simulation, n = 187): the `_c` log-density drops half a log(scale) per trait (cpp:274 vs
        # BSFG_functions.R:99), so the twins may pick different grid cells; both stay on the grid rp (1 - h2) / h2
        H2 = np.arange(1, H) / H
        grid = st["resid_Y_prec"][:, None] * (1 - H2)[None, :] / H2[None, :]
        for x in (a, b):
            assert np.all(np.isclose(x[fin, None], grid[fin], rtol=1e-12).any(axis=1))


@pytest.mark.parametrize("n,p,k,b,pedigree,rect", [(20, 1, 2, 1, "halfsib", False), (30, 13, 7, 0, "random", False),
                                                    (24, 9, 3, 2, "random", True), (40, 17, 5, 1, "halfsib", True)])
def test_twins_agree_on_random_problems(n, p, k, b, pedigree, rect):
    """The same twin agreement away from the fixtures: p = 1, k = 2, b = 0 and b = 2, a random pedigree (n distinct
    eigenvalues) and a rectangular incidence matrix Z_1 (r < n), all functions chained like one sweep."""
    from oracle import bsfg_oracle as O, bsfg_oracle_rtwin as R, init_oracle as I
    rng = np.random.default_rng(1000 + n + p)
    setup = I.make_synthetic(n, p, max(2, k // 2), pedigree=pedigree, seed=n + p, b=b)
    if b == 0:
        setup["X"] = np.zeros((n, 0))
    if rect:                                            # two records per line: r = n / 2 random-effect levels
        r = n // 2
        Z = np.zeros((n, r)); Z[np.arange(n), np.arange(n) % r] = 1.0
        setup["Z_1"] = Z; setup["A"] = setup["A"][:r, :r]; setup["r"] = r
    pri = I.default_priors(k_init=k); rp = I.default_run_parameters()
    st = I.sampler_init(setup, pri, rp, seed=3, lists="stable")
    d, rv, pri, cs = st["data_matrices"], st["run_variables"], st["priors"], st["current_state"]
    Y, X, Z = d["Y"], d["X"], d["Z_1"]
    r = Z.shape[1]
    v = I.draw_variates(rng, n, p, k, r, b + r, pri)
    H = rp["h2_divisions"]
    tol = 1e-9
    B = np.asarray(cs["B"]).reshape(b, p)
    Yt = Y - X @ B
    Lam = O.sample_Lambda_c(Yt, cs["F"], cs["resid_Y_prec"], cs["E_a_prec"], cs["Plam"], rv["invert_aI_bZAZ"], v["z_Lambda"])
    assert rowwise_relerr(Lam, R.sample_Lambda(Yt, cs["F"], cs["resid_Y_prec"], cs["E_a_prec"], cs["Plam"], rv["invert_aI_bZAZ"],
                                               v["z_Lambda"])) < tol
    Yt = Y - cs["F"] @ Lam.T
    loc = O.sample_means_c(Yt, cs["resid_Y_prec"], cs["E_a_prec"], rv["invert_aPXA_bDesignDesignT"], v["z_means"])
    assert relerr(loc, R.sample_means(Yt, cs["resid_Y_prec"], cs["E_a_prec"], rv["invert_aPXA_bDesignDesignT"], v["z_means"])) < tol
    h2 = O.sample_h2s_discrete_c(cs["F"], H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], v["u_h2"])
    assert np.array_equal(h2, R.sample_h2s_discrete(cs["F"], H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], v["u_h2"]))
    Fa = O.sample_F_a_c(cs["F"], Z, h2, rv["invert_aZZt_Ainv"], v["z_Fa"])
    assert relerr(Fa, R.sample_F_a(cs["F"], Z, h2, rv["invert_aZZt_Ainv"], v["z_Fa"])) < tol
    Yt = Y - X @ loc[:b] - Z @ loc[b:]
    F = O.sample_factors_scores_c(Yt, Z, Lam, cs["resid_Y_prec"], Fa, h2, v["z_F"])
    assert rowwise_relerr(F, R.sample_factors_scores(Yt, Z, Lam, cs["resid_Y_prec"], Fa, h2, v["z_F"])) < 1e-8


def test_twins_differ_exactly_where_the_reference_does():
    """cpp:274 integer n/2 vs BSFG_functions.R:99 real n/2 (odd n); cpp:326 `> 10000` vs R:229 `== Inf`."""
    from oracle import bsfg_oracle as O, bsfg_oracle_rtwin as R, init_oracle as I
    n, p = 9, 6
    rng = np.random.default_rng(0)
    A = I.halfsib_A(3, 3)
    s, U = np.linalg.eigh(A)
    lst = {"U": U, "s": s}
    Y = rng.standard_normal((n, p))
    rp_ = rng.gamma(2.0, 1.0, p) + 0.5
    pri = np.ones(10) / 10
    lo = O.prec_log_ps(Y, 10, pri, lst, rp_)
    # the only column the integer division touches is h2 = 0: difference = (n/2 - n//2) log(1/rp)
    Yt_u = Y.T @ U
    want0 = np.sum(O.dnorm_std_log(Y.T * np.sqrt(rp_)[:, None]), axis=1) - (n // 2) * np.log(1 / rp_) + np.log(pri[0])
    assert relerr(lo[:, 0], want0) < 1e-13
    real0 = want0 + (n // 2 - n / 2) * np.log(1 / rp_)
    assert np.max(np.abs(real0 - lo[:, 0])) > 1e-3
    del Yt_u
    # F_a: h2 so close to 1 that tau_e > 10000 but finite -> C++ copies F, R samples
    F = rng.standard_normal((n, 2))
    h2 = np.array([1 - 1e-5, 0.5])
    l3 = {"U": U, "s1": np.ones(n) * 0.5, "s2": np.ones(n) * 0.5}
    z = rng.standard_normal((n, 2))
    c = O.sample_F_a_c(F, np.eye(n), h2, l3, z)
    r_ = R.sample_F_a(F, np.eye(n), h2, l3, z)
    assert np.array_equal(c[:, 0], F[:, 0]) and not np.allclose(r_[:, 0], F[:, 0])
    assert relerr(c[:, 1], r_[:, 1]) < 1e-12


# ------------------------------------------------------------------------------------------------
# (1b) `_c` restatement == transcription of the MATLAB originals (BSF-G/*.m), and the R-vs-MATLAB departures
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_matlab_originals_agree_on_reference_fixtures(name):
    from oracle import bsfg_oracle as O, bsfg_oracle_mtwin as M
    g, d, rv, pri, rp, st, v = _twin_inputs(name)
    Y, X, Z = g["aux1"].get("Y_imputed", d["Y"]), d["X"], d["Z_1"]
    if "Z_2" in d:
        Y = Y - d["Z_2"] @ st["W"]
    H = rp["h2_divisions"]
    tol = 1e-9
    Yt = Y - X @ st["B"]
    Lam = O.sample_Lambda_c(Yt, st["F"], st["resid_Y_prec"], st["E_a_prec"], st["Plam"], rv["invert_aI_bZAZ"], v["z_Lambda"])
    assert rowwise_relerr(Lam, M.sample_lambda(Yt, st["F"], st["resid_Y_prec"], st["E_a_prec"], st["Plam"], rv["invert_aI_bZAZ"],
                                               v["z_Lambda"])) < tol
    Yt = Y - st["F"] @ Lam.T
    loc = O.sample_means_c(Yt, st["resid_Y_prec"], st["E_a_prec"], rv["invert_aPXA_bDesignDesignT"], v["z_means"])
    locm = M.sample_means(Yt, st["resid_Y_prec"], st["E_a_prec"], rv["invert_aPXA_bDesignDesignT"], v["z_means"])
    assert locm.shape == loc.T.shape and relerr(locm.T, loc) < tol          # MATLAB returns the transpose (sample_means.m:29)
    h2 = O.sample_h2s_discrete_c(st["F"], H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], v["u_h2"])
    assert np.array_equal(h2, M.sample_h2s_discrete(st["F"], H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], v["u_h2"]))
    Fa = O.sample_F_a_c(st["F"], Z, h2, rv["invert_aZZt_Ainv"], v["z_Fa"])
    assert relerr(Fa, M.sample_F_a(st["F"], Z, h2, rv["invert_aZZt_Ainv"], v["z_Fa"])) < tol
    Yt = Y - X @ st["B"] - Z @ st["E_a"]
    F = O.sample_factors_scores_c(Yt, Z, Lam, st["resid_Y_prec"], Fa, h2, v["z_F"])
    assert rowwise_relerr(F, M.sample_factors_scores(Yt, Z, Lam, st["resid_Y_prec"], Fa, h2, v["z_F"])) < 1e-8


def test_R_port_departs_from_matlab_exactly_where_documented():
    """sample_delta: R's `2:(k-1)` never resamples delta_k, MATLAB's `2:k` does.
    update_k drop arm: R folds delta over every redundant column but the last, MATLAB only over those below the NEW
    k - 1 (row a8).  The add arm and everything else coincide.  The device path follows R."""
    from oracle import bsfg_oracle as O, bsfg_oracle_mtwin as M
    rng = np.random.default_rng(12)
    p, k, n = 11, 6, 9
    LP = rng.gamma(1.5, 1.0, (p, k)); L2 = rng.standard_normal((p, k)) ** 2
    delta = rng.gamma(3.0, 1.0, k) + 0.5; tauh = np.cumprod(delta)
    g = rng.gamma(4.0, 1.0, k)
    dr = O.sample_delta(delta, tauh, LP, 2.1, 0.05, 4.5, 1.0, L2, g[:k - 1])
    dr = np.asarray(dr[0] if isinstance(dr, tuple) else dr).ravel()
    dm, tm, shapes, rates = M.sample_delta(delta, tauh, LP, 2.1, 0.05, 4.5, 1.0, L2, g)
    assert relerr(dm[:k - 1], dr[:k - 1]) < 1e-13                # same draws for delta_1 .. delta_{k-1}
    assert dr[k - 1] == delta[k - 1] and dm[k - 1] != delta[k - 1]
    assert relerr(tm, np.cumprod(dm)) < 1e-15 and len(shapes) == k

    F = rng.standard_normal((n, k)); Fa = rng.standard_normal((n, k)); h2 = rng.integers(1, 40, k) / 50.0
    Lam = rng.standard_normal((p, k))
    v = dict(u_k=0.0, k_g_lp=rng.gamma(1.5, 1.0, p), k_g_delta=rng.gamma(4.5, 1.0), k_z_lambda=rng.standard_normal(p),
             k_u_h2=rng.uniform(), k_z_Fa=rng.standard_normal(n), k_z_F=rng.standard_normal(n))
    args = lambda L: (F, L, Fa, h2, LP, LP * tauh[None, :], delta, tauh, np.eye(n), 3.0, 4.5, 1.0, 1.0, 5e-4, 300, 1e-2, 1.0, v)
    # add arm: identical
    a, b = O.update_k(*args(Lam)), M.update_k(*args(Lam))
    assert a["F"].shape[1] == k + 1
    for f in a:
        assert relerr(np.asarray(b[f]), np.asarray(a[f])) < 1e-13, f
    # drop arm, one redundant column (2, 1-based): identical
    L1 = Lam.copy(); L1[:, 1] *= 1e-6
    a, b = O.update_k(*args(L1)), M.update_k(*args(L1))
    assert a["F"].shape[1] == k - 1
    for f in a:
        assert relerr(np.asarray(b[f]), np.asarray(a[f])) < 1e-13, f
    # drop arm, redundant columns 2 and 4: R folds both, MATLAB only column 2 (4 > new k - 1 = 3)
    L2c = Lam.copy(); L2c[:, [1, 3]] *= 1e-6
    a, b = O.update_k(*args(L2c)), M.update_k(*args(L2c))
    assert a["F"].shape[1] == b["F"].shape[1] == k - 2
    assert relerr(a["delta"], np.array([delta[0], delta[2] * delta[1], delta[4] * delta[3], delta[5]])) < 1e-15
    assert relerr(b["delta"], np.array([delta[0], delta[2] * delta[1], delta[4], delta[5]])) < 1e-15
    for f in ("F", "Lambda", "F_a", "F_h2", "Lambda_prec"):
        assert np.array_equal(np.asarray(a[f]), np.asarray(b[f])), f


# ------------------------------------------------------------------------------------------------
# (2) closed-form identities (dense algebra straight from the model statement)
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dense_case():
    from oracle import init_oracle as I
    n, p, k = 30, 7, 4
    setup = I.make_synthetic(n, p, 2, pedigree="random", seed=3, b=2)
    pri = I.default_priors(k_init=k)
    st = I.sampler_init(setup, pri, I.default_run_parameters(), seed=4, lists="stable")
    return st, np.asarray(setup["A"])


def test_sample_Lambda_is_the_dense_gaussian_conditional(dense_case):
    """Lambda_j | . ~ N(Q^-1 F' S_j^-1 y_j, Q^-1), Q = F' S_j^-1 F + diag(Plam_j), S_j = ZAZ'/ep_j + I/rp_j
    (the marginal model in the comment block of cpp:93-99)."""
    from oracle import bsfg_oracle as O
    st, A = dense_case
    d, rv, cs = st["data_matrices"], st["run_variables"], st["current_state"]
    Y, Z = d["Y"], d["Z_1"]
    n, p = Y.shape
    k = cs["F"].shape[1]
    F = cs["F"]
    zero = O.sample_Lambda_c(Y, F, cs["resid_Y_prec"], cs["E_a_prec"], cs["Plam"], rv["invert_aI_bZAZ"], np.zeros((k, p)))
    cols = [O.sample_Lambda_c(Y, F, cs["resid_Y_prec"], cs["E_a_prec"], cs["Plam"], rv["invert_aI_bZAZ"],
                              np.tile(np.eye(k)[:, [c]], (1, p))) - zero for c in range(k)]
    for j in range(p):
        S = Z @ A @ Z.T / cs["E_a_prec"][j] + np.eye(n) / cs["resid_Y_prec"][j]
        Si = np.linalg.inv(S)
        Q = F.T @ Si @ F + np.diag(cs["Plam"][j])
        mean = np.linalg.solve(Q, F.T @ Si @ Y[:, j])
        assert relerr(zero[j], mean) < 1e-8
        Rinv = np.column_stack([c[j] for c in cols])           # columns of R^-1 (R'R = Q)
        assert relerr(Rinv @ Rinv.T, np.linalg.inv(Q)) < 1e-8


def test_sample_means_is_the_dense_gaussian_conditional(dense_case):
    from oracle import bsfg_oracle as O
    st, A = dense_case
    d, rv, cs, pri = st["data_matrices"], st["run_variables"], st["current_state"], st["priors"]
    Y, X, Z = d["Y"], d["X"], d["Z_1"]
    b, r = X.shape[1], Z.shape[1]
    D = np.column_stack([X, Z])
    P = np.zeros((b + r, b + r)); P[:b, :b] = pri["b_X_prec"]; P[b:, b:] = rv["Ainv"]
    got = O.sample_means_c(Y, cs["resid_Y_prec"], cs["E_a_prec"], rv["invert_aPXA_bDesignDesignT"], np.zeros((b + r, Y.shape[1])))
    for j in range(Y.shape[1]):
        C = P * cs["E_a_prec"][j] + D.T @ D * cs["resid_Y_prec"][j]
        # fixed-effect rows carry prior precision b_X_prec * E_a_prec_j (cpp:52-54 applies `a` to the whole block)
        want = np.linalg.solve(C, D.T @ Y[:, j] * cs["resid_Y_prec"][j])
        assert relerr(got[:, j], want) < 1e-7


def test_sample_F_a_and_factor_scores_dense(dense_case):
    from oracle import bsfg_oracle as O
    st, A = dense_case
    d, rv, cs = st["data_matrices"], st["run_variables"], st["current_state"]
    Y, Z = d["Y"], d["Z_1"]
    n, p = Y.shape
    F, k = cs["F"], cs["F"].shape[1]
    h2 = np.array([0.0, 0.3, 0.62, 0.98])
    got = O.sample_F_a_c(F, Z, h2, rv["invert_aZZt_Ainv"], np.zeros((Z.shape[1], k)))
    assert np.all(got[:, 0] == 0)
    for j in range(1, k):
        te, tu = 1 / (1 - h2[j]), 1 / h2[j]
        want = np.linalg.solve(Z.T @ Z * te + rv["Ainv"] * tu, Z.T @ F[:, j] * te)
        assert relerr(got[:, j], want) < 1e-8
    Lam, rp_ = cs["Lambda"], cs["resid_Y_prec"]
    Fa = got
    gotF = O.sample_factors_scores_c(Y, Z, Lam, rp_, Fa, h2, np.zeros((n, k)))
    te = 1 / (1 - h2)
    Q = Lam.T @ (Lam * rp_[:, None]) + np.diag(te)
    want = np.linalg.solve(Q, (Y @ (Lam * rp_[:, None]) + (Z @ Fa) * te[None, :]).T).T
    assert rowwise_relerr(gotF, want) < 1e-8
    # px variant = same with tau_e supplied; px_factor is ignored (cpp:166-193)
    gotP = O.sample_px_factors_scores_c(Y, Z, Lam, rp_, Fa, te, np.full(k, 7.0), np.zeros((n, k)))
    assert np.array_equal(gotP, gotF)


def test_h2_grid_log_posteriors_are_mvn_densities(dense_case):
    """log_ps[j,i] - log prior_i = log N(F_j; 0, h2 ZAZ' + (1-h2) I)   (cpp:211-226)."""
    from oracle import bsfg_oracle as O
    from scipy.stats import multivariate_normal
    st, A = dense_case
    d, rv, cs, pri = st["data_matrices"], st["run_variables"], st["current_state"], st["priors"]
    Z = d["Z_1"]
    n = Z.shape[0]
    H = 10
    prior = np.linspace(1, 2, H) / 15.0
    lp = O.h2s_log_ps(cs["F"], H, prior, rv["invert_aI_bZAZ"])
    for i in range(H):
        h2 = i / H
        cov = h2 * Z @ A @ Z.T + (1 - h2) * np.eye(n)
        for j in range(cs["F"].shape[1]):
            want = multivariate_normal(mean=np.zeros(n), cov=cov).logpdf(cs["F"][:, j]) + np.log(prior[i])
            assert abs(lp[j, i] - want) < 1e-8 * max(1.0, abs(want))
    # draw: count of cumsum edges below u
    u = np.array([0.0, 0.5, 0.999999, 1.0])
    got = O.sample_h2s_discrete_c(cs["F"], H, prior, rv["invert_aI_bZAZ"], u)
    assert got[0] == 0.0 and np.all((got * H) % 1 == 0) and np.all(got <= 1.0)


def test_sample_delta_follows_the_R_loop():
    """BSFG_functions.R:296 `for(h in 2:(k-1))`: delta_k is never resampled; k = 2 visits h = 2, 1."""
    from oracle import bsfg_oracle as O
    assert O.delta_loop_indices(5) == [2, 3, 4]
    assert O.delta_loop_indices(3) == [2]
    assert O.delta_loop_indices(2) == [2, 1]
    rng = np.random.default_rng(1)
    p, k = 11, 5
    LP = rng.gamma(1.5, 1.0, (p, k)); L2 = rng.standard_normal((p, k)) ** 2
    delta = rng.gamma(3.0, 1.0, k); tauh = np.cumprod(delta)
    g = rng.gamma(5.0, 1.0, k)
    new, shapes, rates = O.sample_delta(delta, tauh, LP, 2.1, 0.05, 4.5, 1.0, L2, g)
    assert new[-1] == delta[-1] and len(shapes) == k - 1
    c = np.sum(LP * L2, axis=0)
    assert abs(rates[0] - (0.05 + 0.5 / delta[0] * np.sum(tauh * c))) < 1e-12 * rates[0]
    assert shapes[0] == 2.1 + 0.5 * p * k and shapes[1] == 4.5 + 0.5 * p * (k - 1)
    d1 = delta.copy(); d1[0] = g[0] / rates[0]
    t1 = np.cumprod(d1)
    assert abs(rates[1] - (1.0 + 0.5 / d1[1] * np.sum(t1[1:] * c[1:]))) < 1e-12 * rates[1]


def test_update_k_arms_follow_R():
    """BSFG_functions.R:311-372: gate `uu < prob && i > 200`; add arm appends in R's draw order; drop arm folds
    delta[red+1] *= delta[red] except for the last column and recomputes tauh/Plam."""
    from oracle import bsfg_oracle as O
    rng = np.random.default_rng(2)
    n, p, k, r = 8, 12, 4, 8
    F = rng.standard_normal((n, k)); Lam = rng.standard_normal((p, k)) + 2.0
    Fa = rng.standard_normal((r, k)); h2 = rng.uniform(size=k)
    LP = rng.gamma(2, 1, (p, k)); delta = rng.gamma(3, 1, k); tauh = np.cumprod(delta); Plam = LP * tauh
    v = dict(u_k=0.0, k_g_lp=rng.gamma(1.5, 1, p), k_g_delta=2.5, k_z_lambda=rng.standard_normal(p), k_u_h2=0.37,
             k_z_Fa=rng.standard_normal(r), k_z_F=rng.standard_normal(n))
    args = (3.0, 4.5, 1.0, 1.0, 0.0005)
    same = O.update_k(F, Lam, Fa, h2, LP, Plam, delta, tauh, np.eye(n), *args, 200, 1e-2, 1.0, v)
    assert same["F"].shape[1] == k                                   # i > 200 gate
    add = O.update_k(F, Lam, Fa, h2, LP, Plam, delta, tauh, np.eye(n), *args, 201, 1e-2, 1.0, v)
    assert add["F"].shape[1] == k + 1
    assert np.allclose(add["Lambda_prec"][:, k], v["k_g_lp"] / 1.5)
    assert np.isclose(add["delta"][k], 2.5) and np.isclose(add["tauh"][k], tauh[-1] * 2.5)
    assert np.allclose(add["Lambda"][:, k], v["k_z_lambda"] / np.sqrt(add["Plam"][:, k]))
    assert np.allclose(add["F_a"][:, k], v["k_z_Fa"] * np.sqrt(0.37))
    assert np.allclose(add["F"][:, k], add["F_a"][:, k] + v["k_z_F"] * np.sqrt(1 - 0.37))
    Lam2 = Lam.copy(); Lam2[:, 1] = 1e-4; Lam2[:, 3] = 1e-5          # columns 2 and 4 (1-based) redundant
    drop = O.update_k(F, Lam2, Fa, h2, LP, Plam, delta, tauh, np.eye(n), *args, 300, 1e-2, 1.0, v)
    assert drop["F"].shape[1] == 2
    dd = delta.copy(); dd[2] *= dd[1]                                  # red = 2 folds into 3; red = 4 = k skipped
    assert np.allclose(drop["delta"], dd[[0, 2]]) and np.allclose(drop["tauh"], np.cumprod(dd[[0, 2]]))
    assert np.allclose(drop["Plam"], LP[:, [0, 2]] * drop["tauh"])
    nope = O.update_k(F, Lam2, Fa, h2, LP, Plam, delta, tauh, np.eye(n), *args, 300, 1e-2, 1.0, dict(v, u_k=0.99))
    assert nope["F"].shape[1] == k                                   # uu >= prob


def test_gsvd_2_c_contract():
    """GSVD_2_c (cpp:27-40): A'A = X C^2 X', B'B = X S^2 X', C^2 + S^2 = I."""
    from oracle import bsfg_oracle as O
    rng = np.random.default_rng(5)
    A = rng.standard_normal((6, 6)); B = rng.standard_normal((6, 6)) + 3 * np.eye(6)
    res = O.GSVD_2_c(A, B)
    X, C, S = res["X"], res["C"], res["S"]
    assert relerr(X @ C @ C @ X.T, A.T @ A) < 1e-10
    assert relerr(X @ S @ S @ X.T, B.T @ B) < 1e-10
    assert relerr(C @ C + S @ S, np.eye(6)) < 1e-12
