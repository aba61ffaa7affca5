"""GPU parity of the fused, device-resident sweep (rotated-data formulation) against the oracle's
reference-formulation Gibbs iteration (fast_BSFG_sampler.R:176-270) under injected variates."""
import numpy as np
import pytest

from conftest import relerr, rowwise_relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _build(n, p, k, lists, pedigree="halfsib", seed=7, b=1):
    from oracle import init_oracle as I
    setup = I.make_synthetic(n, p, max(2, k // 2), pedigree=pedigree, seed=seed, b=b)
    pri = I.default_priors(k_init=k)
    rp = I.default_run_parameters()
    st = I.sampler_init(setup, pri, rp, seed=3, lists=lists)
    return st


def _compare_states(dev, ora, b, tol=TOL, h2_exact=True):
    assert dev["F"].shape == ora["F"].shape
    assert rowwise_relerr(dev["Lambda"], ora["Lambda"]) < tol
    assert rowwise_relerr(dev["F"], ora["F"]) < tol
    assert rowwise_relerr(dev["F_a"].T, ora["F_a"].T) < tol
    if h2_exact:
        assert np.array_equal(dev["F_h2"], np.asarray(ora["F_h2"]).ravel())
    assert relerr(dev["B"], ora["B"].reshape(dev["B"].shape)) < tol
    assert rowwise_relerr(dev["E_a"].T, ora["E_a"].T) < tol
    assert relerr(dev["Lambda_prec"], ora["Lambda_prec"]) < tol
    assert relerr(dev["resid_Y_prec"], ora["resid_Y_prec"]) < tol
    assert relerr(dev["E_a_prec"], ora["E_a_prec"]) < tol
    assert relerr(dev["delta"], ora["delta"]) < tol
    assert relerr(dev["tauh"], ora["tauh"]) < tol
    assert relerr(dev["Plam"], ora["Plam"]) < tol * 10


@pytest.mark.parametrize("gram", ["expsum", "exact"])
@pytest.mark.parametrize("lists,mode,expect", [("stable", "auto", "diag"), ("reference", "auto", "direct"),
                                               ("stable", "direct", "direct")])
def test_sweep_injected_matches_oracle(lists, mode, expect, gram):
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k = 60, 45, 6
    st = _build(n, p, k, lists)
    d, rv, pri, rp, cs = (st["data_matrices"], st["run_variables"], st["priors"], st["run_parameters"], st["current_state"])
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=12, mode=mode, gram=gram)
    try:
        assert sw.gram_stats()["mode"] == gram
        got_mode, r1, r2 = sw.mode()
        assert got_mode == expect, (got_mode, r1, r2)
        rng = np.random.default_rng(21)
        ora = cs
        for it in range(1, 4):
            v = I.draw_variates(rng, n, p, k, n, 1 + n, pri)
            sw.set_state(ora)                      # re-synchronise, then compare one iteration
            ora_next, aux = O.gibbs_iteration(d, rv, pri, rp, ora, it, v)
            sw.sweep_injected(v)
            dev = sw.get_state()
            assert dev["nrun"] == it
            _compare_states(dev, ora_next, 1)
            a = sw.aux()
            assert relerr(a["resid_Y_prec_rate"], aux["resid_Y_prec_rate"]) < TOL
            assert relerr(a["E_a_prec_rate"], aux["E_a_prec_rate"]) < TOL
            assert relerr(a["Lambda_prec_rate"], aux["Lambda_prec_rate"]) < TOL
            nd = len(aux["delta_rate"])
            assert relerr(a["delta_rate"][:nd], aux["delta_rate"]) < TOL
            ora = ora_next
        assert sw.gram_stats()["fallbacks"] == 0
    finally:
        sw.close()


def test_sweep_expsum_falls_back_to_exact_gram_when_range_is_too_wide():
    """(s + E_a_prec/resid_Y_prec) spanning 15 decades cannot be covered by 256 nodes: the device raises the fallback flag,
    the exact Gram GEMM runs for that iteration, the result still matches the oracle, and the event is counted."""
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k = 60, 45, 6
    st = _build(n, p, k, "stable")
    d, rv, pri, rp, cs = (st["data_matrices"], st["run_variables"], st["priors"], st["run_parameters"], st["current_state"])
    cs = dict(cs)
    cs["E_a_prec"] = cs["E_a_prec"].copy(); cs["resid_Y_prec"] = cs["resid_Y_prec"].copy()
    cs["E_a_prec"][0] *= 1e15; cs["resid_Y_prec"][1] *= 1e3
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=12, gram="expsum")
    try:
        v = I.draw_variates(np.random.default_rng(5), n, p, k, n, 1 + n, pri)
        sw.set_state(cs)
        ora, _ = O.gibbs_iteration(d, rv, pri, rp, cs, 1, v)
        sw.sweep_injected(v)
        gs = sw.gram_stats()
        assert gs["fallbacks"] == 1 and gs["c_max"] / gs["c_min"] > 1e16, gs
        dev = sw.get_state()
        assert rowwise_relerr(dev["Lambda"], ora["Lambda"]) < TOL
        assert rowwise_relerr(dev["F"], ora["F"]) < TOL
        # the next iteration has ordinary precisions again -> no further fallback
        v = I.draw_variates(np.random.default_rng(6), n, p, k, n, 1 + n, pri)
        sw.set_state(ora)
        ora2, _ = O.gibbs_iteration(d, rv, pri, rp, ora, 2, v)
        sw.sweep_injected(v)
        assert sw.gram_stats()["fallbacks"] == 1
        assert rowwise_relerr(sw.get_state()["Lambda"], ora2["Lambda"]) < TOL
    finally:
        sw.close()


@pytest.mark.parametrize("lists", ["stable", "reference"])
@pytest.mark.parametrize("gram", ["expsum", "exact"])
def test_sweep_second_random_effect(lists, gram):
    """Z_2 / W (fast_BSFG_sampler.R:211-216, 248-250): the second sample_means_c call with W_prec and the rand2 list,
    every residual that carries Z_2 W, and the W_prec Gamma draw."""
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k, r2 = 55, 41, 5, 9
    rng = np.random.default_rng(77)
    setup = I.make_synthetic(n, p, 3, pedigree="halfsib", seed=19)
    Z2 = np.zeros((n, r2)); Z2[np.arange(n), rng.integers(0, r2, n)] = 1.0
    setup["Z_2"] = Z2
    pri = I.default_priors(k_init=k); rp = I.default_run_parameters()
    st = I.sampler_init(setup, pri, rp, seed=4, lists=lists)
    d, rv, pri, cs = st["data_matrices"], st["run_variables"], st["priors"], st["current_state"]
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=8, gram=gram)
    try:
        ora = cs
        for it in (1, 2):
            v = I.draw_variates(rng, n, p, k, n, 1 + n, pri, r2=r2)
            sw.set_state(ora)
            ora_next, aux = O.gibbs_iteration(d, rv, pri, rp, ora, it, v)
            sw.sweep_injected(v)
            dev = sw.get_state()
            _compare_states(dev, ora_next, 1)
            assert rowwise_relerr(dev["W"].T, ora_next["W"].T) < TOL
            assert relerr(dev["W_prec"], ora_next["W_prec"]) < TOL
            a = sw.aux()
            assert relerr(a["W_prec_rate"], aux["W_prec_rate"]) < TOL
            assert relerr(a["resid_Y_prec_rate"], aux["resid_Y_prec_rate"]) < TOL
            ora = ora_next
        sw.set_state(cs)
        sw.sweep(3)                      # device Philox path with W
        out = sw.get_state()
        assert np.all(np.isfinite(out["W"])) and np.all(out["W_prec"] > 0)
    finally:
        sw.close()


@pytest.mark.parametrize("with_z2", [False, True])
def test_sweep_missing_phenotypes(with_z2):
    """NaN in Y (fast_BSFG_sampler.R:180-184): imputed from the current state + rnorm at the start of every iteration,
    then every rotated quantity is rebuilt.  20% missing, like the Ayroles panel."""
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k, r2 = 45, 38, 4, (6 if with_z2 else 0)
    rng = np.random.default_rng(99)
    setup = I.make_synthetic(n, p, 2, pedigree="halfsib", seed=23)
    Y = setup["Y"].copy(); Y[rng.uniform(size=Y.shape) < 0.2] = np.nan; Y[:, 3] = setup["Y"][:, 3]   # one complete trait
    setup["Y"] = Y
    if with_z2:
        Z2 = np.zeros((n, r2)); Z2[np.arange(n), rng.integers(0, r2, n)] = 1.0
        setup["Z_2"] = Z2
    pri = I.default_priors(k_init=k); rp = I.default_run_parameters()
    st = I.sampler_init(setup, pri, rp, seed=8, lists="stable")
    d, rv, pri, cs = st["data_matrices"], st["run_variables"], st["priors"], st["current_state"]
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=8)
    try:
        sw.set_state(cs)
        ora = cs
        for it in (1, 2, 3):                       # free running: iteration 2+ imputes from the device's own theta
            v = I.draw_variates(rng, n, p, k, n, 1 + n, pri, r2=r2, missing=True)
            ora, aux = O.gibbs_iteration(d, rv, pri, rp, ora, it, v)
            sw.sweep_injected(v)
            a = sw.aux()
            # iteration 1 starts from the same state: the north-star 1e-9; iterations 2, 3 run free (no re-synchronisation),
            # the imputation then sees the device's own theta and the difference compounds by about 10x per iteration
            assert relerr(a["Y_imputed"], aux["Y_imputed"]) < (1e-9 if it == 1 else 1e-8), it
            assert np.array_equal(a["Y_imputed"][~np.isnan(Y)], Y[~np.isnan(Y)])      # observed entries untouched
        dev = sw.get_state()
        _compare_states(dev, ora, 1, tol=1e-7)
        if with_z2:
            assert rowwise_relerr(dev["W"].T, ora["W"].T) < 1e-7
        sw.sweep(2)                                # device Philox path
        assert np.all(np.isfinite(sw.get_state()["Lambda"]))
    finally:
        sw.close()


@pytest.mark.parametrize("b,with_z2", [(2, False), (0, True), (1, False)])
def test_sweep_fixed_lambda(b, with_z2):
    """fast_BSFG_sampler_fixedlambda.R:159-236: Lambda supplied per iteration, everything else resampled; the (B,E_a)
    step removes X B as that driver does."""
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k, r2 = 50, 33, 4, (5 if with_z2 else 0)
    rng = np.random.default_rng(123)
    setup = I.make_synthetic(n, p, 2, pedigree="halfsib", seed=29, b=max(b, 1))
    if b == 0:
        setup["X"] = np.zeros((n, 0))
    if with_z2:
        Z2 = np.zeros((n, r2)); Z2[np.arange(n), rng.integers(0, r2, n)] = 1.0
        setup["Z_2"] = Z2
    pri = I.default_priors(k_init=k); rp = I.default_run_parameters()
    st = I.sampler_init(setup, pri, rp, seed=9, lists="stable")
    d, rv, pri, cs = st["data_matrices"], st["run_variables"], st["priors"], st["current_state"]
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=6)
    try:
        ora = cs
        for it in (1, 2):
            Lam = rng.standard_normal((p, k)) * 0.3                       # a "stored posterior draw"
            v = I.draw_variates(rng, n, p, k, n, b + n, pri, r2=r2)
            sw.set_state(ora)
            ora_next, aux = O.gibbs_iteration_fixedlambda(d, rv, pri, rp, ora, Lam, v)
            sw.sweep_fixed_lambda(Lam, v)
            dev = sw.get_state()
            assert np.array_equal(dev["Lambda"], Lam)
            assert rowwise_relerr(dev["F"], ora_next["F"]) < TOL
            assert rowwise_relerr(dev["F_a"].T, ora_next["F_a"].T) < TOL
            assert np.array_equal(dev["F_h2"], np.asarray(ora_next["F_h2"]).ravel())
            if b > 0:
                assert relerr(dev["B"], ora_next["B"].reshape(dev["B"].shape)) < TOL
            assert rowwise_relerr(dev["E_a"].T, ora_next["E_a"].T) < TOL
            assert relerr(dev["resid_Y_prec"], ora_next["resid_Y_prec"]) < TOL
            assert relerr(dev["E_a_prec"], ora_next["E_a_prec"]) < TOL
            assert relerr(dev["delta"], ora["delta"]) == 0.0 and relerr(dev["Lambda_prec"], ora["Lambda_prec"]) == 0.0
            if with_z2:
                assert rowwise_relerr(dev["W"].T, ora_next["W"].T) < TOL
                assert relerr(dev["W_prec"], ora_next["W_prec"]) < TOL
            ora = ora_next
        sw.sweep_fixed_lambda(Lam)                                        # device Philox draws
        assert np.all(np.isfinite(sw.get_state()["F"]))
    finally:
        sw.close()


@pytest.mark.parametrize("adapt", [False, True])
def test_device_posterior_matches_host_bookkeeping(adapt):
    """bsfg_sweep_collect keeps save_posterior_samples (BSFG_functions.R:375-408) on the device: same chain (same seed),
    samples pulled with get_state at every thinning boundary on the host vs collected on the device must be identical
    for the traces and agree to rounding for the running means (B, E_a go through the mean of theta).
    adapt: the chain starts at nrun = 300 from a state with two crushed loading columns, so update_k drops and adds
    factors between the samples -- the traces keep k_max columns per sample, zero beyond that sample's k (the
    reference grows the rows of Posterior$Lambda / F / F_a instead, BSFG_functions.R:383-392)."""
    from oracle import init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k = 40, 30, 5
    st = _build(n, p, k, "stable", seed=17)
    d, rv, pri, rp, cs = (st["data_matrices"], st["run_variables"], st["priors"], st["run_parameters"], st["current_state"])
    if adapt:
        cs = dict(cs)
        cs["delta"] = np.array([1.5, 1.2, 1.1, 1e12, 1.2]); cs["tauh"] = np.cumprod(cs["delta"])
        cs["Plam"] = cs["Lambda_prec"] * cs["tauh"][None, :]
        cs["nrun"] = 300
        rp = dict(rp, b0=-5.0)                      # prob = 1/exp(b0 + b1 i) > 1: update_k acts every iteration (R:324, 331)
    i0 = int(cs.get("nrun", 0))
    KM = 16 if adapt else 8
    burn, thin, iters = i0 + 3, 2, 11               # samples at i = i0 + 5, 7, 9, 11 (the schedule counts absolute iterations)
    a = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=KM, seed=5)
    bsw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=KM, seed=5)
    try:
        a.set_state(cs); bsw.set_state(cs)
        host = []
        a.sweep(burn - i0 + thin)
        for s in range(4):
            host.append(a.get_state())
            if s < 3:
                a.sweep(thin)
        bsw.posterior_begin(8)
        bsw.collect(iters, burn, thin)
        post = bsw.posterior()
        assert list(post["sp_num"]) == [1, 2, 3, 4] and post["k_max"] == KM
        for s, hs in enumerate(host):
            kk = hs["F"].shape[1]
            for name, rows in (("Lambda", p), ("F", n), ("F_a", n)):
                got = post[name][:, s].reshape(rows, KM, order="F")
                assert np.array_equal(got[:, :kk], hs[name]) and not got[:, kk:].any()
            assert np.array_equal(post["delta"][:kk, s], hs["delta"]) and np.array_equal(post["F_h2"][:kk, s], hs["F_h2"])
            assert np.array_equal(post["resid_Y_prec"][:, s], hs["resid_Y_prec"])
            assert np.array_equal(post["E_a_prec"][:, s], hs["E_a_prec"])
        Bm = np.zeros_like(host[0]["B"]); Em = np.zeros_like(host[0]["E_a"])
        for s, hs in enumerate(host, start=1):
            Bm = (Bm * (s - 1) + hs["B"]) / s; Em = (Em * (s - 1) + hs["E_a"]) / s
        assert relerr(post["B"], Bm) < 1e-12 and relerr(post["E_a"], Em) < 1e-12
        # both chains ended in the same state
        assert np.array_equal(bsw.get_state()["F"], a.get_state()["F"])
        if adapt:
            ks = [hs["F"].shape[1] for hs in host]
            assert len(set(ks)) > 1 or ks[0] != k, f"update_k never changed k ({ks}); test inputs need retuning"
    finally:
        a.close(); bsw.close()


def test_sweep_adaptive_expsum_terms():
    """gram_terms = -1: the node count adapts between bsfg_sweep calls; parity with the oracle is unchanged."""
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k = 60, 45, 6
    st = _build(n, p, k, "stable")
    d, rv, pri, rp, cs = (st["data_matrices"], st["run_variables"], st["priors"], st["run_parameters"], st["current_state"])
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=12, gram="expsum", gram_terms=-1)
    try:
        sw.set_state(cs)
        sw.sweep(2)
        assert sw.gram_stats()["terms"] == 256                   # first call: no range known yet
        sw.sweep(1)
        t = sw.gram_stats()["terms"]
        assert 64 <= t < 256 and t % 16 == 0, t
        rng = np.random.default_rng(3)
        ora = sw.get_state()
        ora["nrun"] = 3
        v = I.draw_variates(rng, n, p, k, n, 1 + n, pri)
        sw.set_state(ora)
        want, _ = O.gibbs_iteration(d, rv, pri, rp, ora, 4, v)
        sw.sweep_injected(v)
        assert sw.gram_stats()["terms"] == t and sw.gram_stats()["fallbacks"] == 0
        assert rowwise_relerr(sw.get_state()["Lambda"], want["Lambda"]) < TOL
    finally:
        sw.close()


@pytest.mark.parametrize("n,p,k,b", [(10, 1, 2, 1), (15, 3, 2, 0), (5, 7, 3, 1), (20, 2, 9, 2)])
def test_sweep_edge_sizes(n, p, k, b):
    """Smallest shapes the reference's R code accepts: a single trait, k = 2 (where R's `2:(k-1)` loop of sample_delta
    counts DOWN, BSFG_functions.R:296), no fixed effects (b = 0), more factors than traits, n not a multiple of anything."""
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    setup = I.make_synthetic(n, max(p, 2), 2, pedigree="halfsib", seed=n + p, b=max(b, 1))
    setup["Y"] = setup["Y"][:, :p]
    if b == 0:
        setup["X"] = np.zeros((n, 0))
    pri = I.default_priors(k_init=k); rp = I.default_run_parameters()
    st = I.sampler_init(setup, pri, rp, seed=2, lists="stable")
    d, rv, pri, cs = st["data_matrices"], st["run_variables"], st["priors"], st["current_state"]
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=k + 2)
    try:
        rng = np.random.default_rng(1)
        ora = cs
        for it in (1, 2):
            v = I.draw_variates(rng, n, p, k, n, b + n, pri)
            sw.set_state(ora)
            ora_next, aux = O.gibbs_iteration(d, rv, pri, rp, ora, it, v)
            sw.sweep_injected(v)
            dev = sw.get_state()
            assert rowwise_relerr(dev["Lambda"], ora_next["Lambda"]) < TOL
            assert rowwise_relerr(dev["F"], ora_next["F"]) < TOL
            assert rowwise_relerr(dev["F_a"].T, ora_next["F_a"].T) < TOL
            assert np.array_equal(dev["F_h2"], np.asarray(ora_next["F_h2"]).ravel())
            if b > 0:
                assert relerr(dev["B"], ora_next["B"].reshape(dev["B"].shape)) < TOL
            assert rowwise_relerr(dev["E_a"].T, ora_next["E_a"].T) < TOL
            for f in ("Lambda_prec", "resid_Y_prec", "E_a_prec", "delta", "tauh"):
                assert relerr(dev[f], ora_next[f]) < TOL, f
            nd = len(aux["delta_rate"])
            assert nd == (3 if k == 2 else k - 1)                      # k = 2: delta_1, then h = 2, 1 -> delta_1 is drawn twice
            assert relerr(sw.aux()["delta_rate"][:nd], aux["delta_rate"]) < TOL
            ora = ora_next
        sw.sweep(3)
        assert np.all(np.isfinite(sw.get_state()["Lambda"]))
    finally:
        sw.close()


def test_sweep_rejects_inconsistent_calls():
    """Errors are loud (negative status -> BSFGError), like Armadillo's exceptions through Rcpp: wrong call order, k out of
    range, k_max beyond the shared-memory limit of the per-trait solver."""
    from oracle import init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep, BSFGError, _lib
    st = _build(20, 8, 3, "stable", seed=3)
    d, rv, pri, rp, cs = (st["data_matrices"], st["run_variables"], st["priors"], st["run_parameters"], st["current_state"])
    sw = BSFGSweep(d, rv, pri, rp, k_init=3, k_max=4)
    try:
        with pytest.raises(BSFGError) as ei:
            sw.sweep(1)                                           # no state yet
        assert ei.value.code == _lib.BSFG_ESTATE
        big = dict(cs)
        for nm in ("F", "F_a", "Lambda", "Lambda_prec", "Plam"):
            big[nm] = np.column_stack([cs[nm]] * 2)               # 6 columns > k_max = 4
        big["delta"] = np.tile(cs["delta"], 2); big["tauh"] = np.tile(cs["tauh"], 2); big["F_h2"] = np.tile(cs["F_h2"], 2)
        with pytest.raises(BSFGError) as ei:
            sw.set_state(big)
        assert ei.value.code == _lib.BSFG_ESHAPE
    finally:
        sw.close()
    with pytest.raises(BSFGError) as ei:
        BSFGSweep(d, rv, pri, rp, k_init=3, k_max=600)             # beyond bsfg_k_limit() = 504
    assert ei.value.code == _lib.BSFG_ESHAPE


def _px_state(cs, pri, rng):
    cs = dict(cs)
    k = cs["F"].shape[1]
    cs["px_factor"] = rng.gamma(pri["px_shape"], 1.0, k) / pri["px_rate"] + 0.05      # fast_BSFG_sampler_init_px.R:221
    cs["F_px"] = cs["F"] * np.sqrt(cs["px_factor"])[None, :]                        # :222
    return cs


def test_sweep_parameter_expanded():
    """fast_BSFG_sampler_px.R:196-311: Lambda_px / F_px / px_factor; Plam divided by px_factor; F and Lambda derived."""
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k = 55, 42, 5
    st = _build(n, p, k, "stable", seed=31)
    d, rv, pri, rp, cs = (st["data_matrices"], st["run_variables"], st["priors"], st["run_parameters"], st["current_state"])
    rng = np.random.default_rng(8)
    ora = _px_state(cs, pri, rng)
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=8)
    try:
        for it in (1, 2, 3):
            v = I.draw_variates(rng, n, p, k, n, 1 + n, pri, px=True)
            sw.set_state(ora)
            ora_next, aux = O.gibbs_iteration_px(d, rv, pri, rp, ora, it, v)
            sw.sweep_injected(v)
            dev = sw.get_state()
            _compare_states(dev, ora_next, 1)
            assert relerr(dev["px_factor"], ora_next["px_factor"]) < TOL
            assert rowwise_relerr(dev["F_px"], ora_next["F_px"]) < TOL
            a = sw.aux()
            assert relerr(a["px_rate"], aux["px_rate"]) < TOL
            assert relerr(a["Lambda_prec_rate"], aux["Lambda_prec_rate"]) < TOL
            ora = ora_next
        sw.set_state(ora)
        sw.sweep(3)                                   # device Philox path of the px loop
        out = sw.get_state()
        assert np.all(np.isfinite(out["F"])) and np.all(out["px_factor"] > 0)
        assert relerr(out["F"] * np.sqrt(out["px_factor"])[None, :], out["F_px"]) < 1e-13
        # device posterior in px mode stores the identified Lambda and F (:317), not Lambda_px / F_px
        sw.posterior_begin(2)
        sw.collect(2, 0, 1)
        post, fin = sw.posterior(), sw.get_state()
        assert relerr(post["Lambda"][:, 1].reshape(p, 8, order="F")[:, :k], fin["Lambda"]) < 1e-13
        assert relerr(post["F"][:, 1].reshape(n, 8, order="F")[:, :k], fin["F"]) < 1e-13
    finally:
        sw.close()


def test_update_k_px_add_and_drop():
    """update_k of BSFG_functions_px.R:301-369: px_factor / F_px follow the added or dropped columns and Plam is recomputed
    WITHOUT the px division in both arms (:328, :348)."""
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k = 40, 30, 5
    st = _build(n, p, k, "stable", seed=13)
    d, rv, pri, rp, cs = (st["data_matrices"], st["run_variables"], st["priors"], st["run_parameters"], st["current_state"])
    rng = np.random.default_rng(4)
    cs = dict(cs)
    # the 4th delta is huge: the last two loading columns are shrunk to ~1e-6 -> redundant -> the drop arm fires first
    # (k: 5 -> 3), after which every column is important and the add arm fires (3 -> 4 -> 5 -> 6)
    cs["delta"] = np.array([1.5, 1.2, 1.1, 1e12, 1.2]); cs["tauh"] = np.cumprod(cs["delta"])
    cs["Plam"] = cs["Lambda_prec"] * cs["tauh"][None, :]
    cs["nrun"] = 300
    ora = _px_state(cs, pri, rng)
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=8)
    try:
        sw.set_state(ora)
        seen = set()
        for it in range(301, 305):
            kk = ora["F"].shape[1]
            v = I.draw_variates(rng, n, p, kk, n, 1 + n, pri, add_arm=True, px=True)
            v["u_k"] = 0.0                                          # always inside the adaptation probability
            ora, _ = O.gibbs_iteration_px(d, rv, pri, rp, ora, it, v)
            sw.sweep_injected(v)
            dev = sw.get_state()
            k2 = ora["F"].shape[1]
            seen.add("add" if k2 > kk else ("drop" if k2 < kk else "same"))
            assert dev["F"].shape[1] == k2
            errs = dict(F_px=rowwise_relerr(dev["F_px"], ora["F_px"]), px_factor=relerr(dev["px_factor"], ora["px_factor"]),
                        Lambda=rowwise_relerr(dev["Lambda"], ora["Lambda"]), Plam=relerr(dev["Plam"], ora["Plam"]),
                        delta=relerr(dev["delta"], ora["delta"]))
            assert max(errs.values()) < TOL, (it, errs)             # state re-synchronised every iteration: the 1e-9 bar
            sw.set_state(ora)
        assert {"add", "drop"} <= seen, seen
    finally:
        sw.close()


def test_sweep_free_running_three_iterations():
    """no re-synchronisation: the chain itself must track the oracle (errors compound, so 1e-7)."""
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k = 50, 64, 5
    st = _build(n, p, k, "stable", seed=11)
    d, rv, pri, rp, cs = (st["data_matrices"], st["run_variables"], st["priors"], st["run_parameters"], st["current_state"])
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=10)
    try:
        sw.set_state(cs)
        rng = np.random.default_rng(22)
        ora = cs
        for it in range(1, 4):
            v = I.draw_variates(rng, n, p, k, n, 1 + n, pri)
            ora, _ = O.gibbs_iteration(d, rv, pri, rp, ora, it, v)
            sw.sweep_injected(v)
        dev = sw.get_state()
        _compare_states(dev, ora, 1, tol=1e-7)
    finally:
        sw.close()


def test_sweep_random_pedigree_rectangular_design():
    """n distinct eigenvalues (random pedigree), b = 2 fixed effects, odd sizes."""
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k = 53, 37, 7
    st = _build(n, p, k, "stable", pedigree="random", seed=5, b=2)
    d, rv, pri, rp, cs = (st["data_matrices"], st["run_variables"], st["priors"], st["run_parameters"], st["current_state"])
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=9)
    try:
        rng = np.random.default_rng(23)
        v = I.draw_variates(rng, n, p, k, n, 2 + n, pri)
        sw.set_state(cs)
        ora, _ = O.gibbs_iteration(d, rv, pri, rp, cs, 1, v)
        sw.sweep_injected(v)
        _compare_states(sw.get_state(), ora, 2)
    finally:
        sw.close()


def test_update_k_add_and_drop():
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k = 40, 30, 5
    st = _build(n, p, k, "stable", seed=13)
    d, rv, pri, rp, cs = (st["data_matrices"], st["run_variables"], st["priors"], st["run_parameters"], st["current_state"])
    cs = dict(cs)
    # moderate shrinkage everywhere -> no redundant column -> the add arm fires when uu < prob and i > 200
    cs["delta"] = np.array([1.5, 1.2, 1.1, 1.3, 1.2]); cs["tauh"] = np.cumprod(cs["delta"])
    cs["Plam"] = cs["Lambda_prec"] * cs["tauh"][None, :]
    cs["nrun"] = 300
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=8)
    try:
        rng = np.random.default_rng(24)
        v = I.draw_variates(rng, n, p, k, n, 1 + n, pri, add_arm=True)
        v["u_k"] = 1e-6
        sw.set_state(cs)
        ora, _ = O.gibbs_iteration(d, rv, pri, rp, cs, 301, v)
        assert ora["F"].shape[1] == k + 1, "oracle did not take the add arm; test inputs need retuning"
        sw.sweep_injected(v)
        dev = sw.get_state()
        assert dev["F"].shape[1] == k + 1
        _compare_states(dev, ora, 1)
        # next iteration runs with k+1 factors
        v2 = I.draw_variates(rng, n, p, k + 1, n, 1 + n, pri)
        v2["u_k"] = 0.99
        ora2, _ = O.gibbs_iteration(d, rv, pri, rp, ora, 302, v2)
        sw.set_state(ora)
        sw.sweep_injected(v2)
        _compare_states(sw.get_state(), ora2, 1)
        # drop arm: crush columns 2 and 4 (0-based 1, 3) with enormous shrinkage -> all |Lambda| < eps
        cs3 = dict(ora2)
        kk = cs3["F"].shape[1]
        delta = np.array([1.5, 1e14, 1e-14 * 1.2, 1e14, 1e-14 * 1.3, 1.1])[:kk]
        cs3["delta"] = delta; cs3["tauh"] = np.cumprod(delta)
        cs3["Plam"] = cs3["Lambda_prec"] * cs3["tauh"][None, :]
        v3 = I.draw_variates(rng, n, p, kk, n, 1 + n, pri)
        v3["u_k"] = 1e-6
        ora3, _ = O.gibbs_iteration(d, rv, pri, rp, cs3, 303, v3)
        assert ora3["F"].shape[1] < kk, "oracle did not take the drop arm; test inputs need retuning"
        sw.set_state(cs3)
        sw.sweep_injected(v3)
        dev3 = sw.get_state()
        assert dev3["F"].shape[1] == ora3["F"].shape[1]
        _compare_states(dev3, ora3, 1, tol=1e-8)
    finally:
        sw.close()


def test_philox_sweep_reproducible_and_finite():
    """production path: device draws; same seed -> identical chain, different seed -> different."""
    from sparsefactormixedmodel_b200 import BSFGSweep
    n, p, k = 60, 45, 6
    st = _build(n, p, k, "stable")
    d, rv, pri, rp, cs = (st["data_matrices"], st["run_variables"], st["priors"], st["run_parameters"], st["current_state"])
    outs = []
    for seed in (5, 5, 6):
        sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=12, seed=seed)
        try:
            sw.set_state(cs)
            sw.sweep(5)
            outs.append(sw.get_state())
            t, launches, _ = sw.timing()
            assert t > 0 and launches > 100
        finally:
            sw.close()
    for name in ("Lambda", "F", "F_a", "resid_Y_prec", "E_a_prec", "delta"):
        assert np.all(np.isfinite(outs[0][name]))
        assert np.array_equal(outs[0][name], outs[1][name])
    assert not np.array_equal(outs[0]["F"], outs[2]["F"])
    assert outs[0]["nrun"] == 5


def test_fast_BSFG_sampler_contract():
    """fast_BSFG_sampler(BSFG_state, n_samples) -> BSFG_state with Posterior filled on the burn/thin schedule."""
    from sparsefactormixedmodel_b200 import fast_BSFG_sampler
    n, p, k = 40, 30, 4
    st = _build(n, p, k, "stable", seed=17)
    st["run_parameters"].update(burn=4, thin=2)
    out = fast_BSFG_sampler(st, 10, seed=3)
    assert out["current_state"]["nrun"] == 10
    assert st["current_state"]["nrun"] == 0                       # input not mutated
    assert out["Posterior"]["Lambda"].shape == (p * k, 3)          # i = 6, 8, 10
    assert out["Posterior"]["E_a"].shape == (n, p)
    out2 = fast_BSFG_sampler(out, 4, seed=3)                       # continuation (README.md:33)
    assert out2["current_state"]["nrun"] == 14
    assert out2["Posterior"]["Lambda"].shape[1] == 5
    # "should be the same as running one continuous chain" (README.md:33): 10 + 4 iterations == 14 in one call
    one = fast_BSFG_sampler(st, 14, seed=3)
    assert np.array_equal(one["Posterior"]["Lambda"], out2["Posterior"]["Lambda"])
    assert np.array_equal(one["current_state"]["F"], out2["current_state"]["F"])
    assert relerr(out2["Posterior"]["E_a"], one["Posterior"]["E_a"]) < 1e-12
    assert relerr(out2["Posterior"]["B"], one["Posterior"]["B"]) < 1e-12


def test_fast_BSFG_sampler_continuation_with_changing_k():
    """The same continuation property while update_k is live (nrun > 200, prob forced above 1): 5 + 4 iterations in two
    calls == 9 in one, with Posterior$Lambda / F / F_a growing rows to the largest k seen (BSFG_functions.R:383-392)."""
    from sparsefactormixedmodel_b200 import fast_BSFG_sampler
    n, p, k = 40, 30, 5
    st = _build(n, p, k, "stable", seed=17)
    st["run_parameters"].update(burn=300, thin=2, b0=-5.0)
    cs = st["current_state"]
    cs["delta"] = np.array([1.5, 1.2, 1.1, 1e12, 1.2]); cs["tauh"] = np.cumprod(cs["delta"])
    cs["Plam"] = cs["Lambda_prec"] * cs["tauh"][None, :]
    cs["nrun"] = 300
    one = fast_BSFG_sampler(st, 9, seed=3)
    a = fast_BSFG_sampler(st, 5, seed=3)
    b = fast_BSFG_sampler(a, 4, seed=3)
    k_end = one["current_state"]["F"].shape[1]
    assert k_end != k and b["current_state"]["F"].shape[1] == k_end
    assert np.array_equal(one["current_state"]["F"], b["current_state"]["F"])
    assert np.array_equal(one["current_state"]["Lambda"], b["current_state"]["Lambda"])
    for name in ("Lambda", "F", "F_a", "delta", "F_h2", "resid_Y_prec", "E_a_prec"):
        assert one["Posterior"][name].shape == b["Posterior"][name].shape, name
        assert np.array_equal(one["Posterior"][name], b["Posterior"][name]), name
    assert one["Posterior"]["Lambda"].shape[1] == 4                      # i = 302, 304, 306, 308
    rows = one["Posterior"]["Lambda"].shape[0]                          # p x the largest k among the SAMPLED iterations
    assert rows % p == 0 and k < rows // p <= k_end
    assert relerr(b["Posterior"]["E_a"], one["Posterior"]["E_a"]) < 1e-12
