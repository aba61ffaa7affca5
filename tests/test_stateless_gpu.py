"""GPU parity tests: every C-ABI drop-in vs the CPU oracle under injected variates.

Tolerance (BASELINE.json north_star): 1e-9 relative in FP64 for continuous draws; discrete h2 indices
exact except where two grid log-likelihoods differ by less than 1e-9 relative."""
import numpy as np
import pytest

from conftest import relerr, rowwise_relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def pkg():
    import sparsefactormixedmodel_b200 as pkg
    assert pkg.load().bsfg_device_count() > 0, "no CUDA device: the product has no CPU path"
    return pkg


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (5, 7, 3), (128, 128, 16), (130, 257, 100), (100, 300, 1000),
                                   (257, 129, 33), (64, 64, 4097), (300, 20, 515),
                                   # stream-K decompositions: several pieces per tile, pieces spanning tile boundaries, complete
                                   # and partial tiles in one launch, the transposed-store path, a single tile over a long K
                                   (100, 5000, 2000), (3000, 100, 4000), (700, 650, 333), (100, 100, 20000), (1275, 2500, 256),
                                   # the K = k product with a Cin: many tiles, few k-tiles (Cin-prefetch kernel), ragged edges
                                   (2601, 4003, 100)])
def test_dgemm_tn(pkg, M, N, K):
    rng = np.random.default_rng(M * 1000 + N * 10 + K)
    At = rng.standard_normal((K, M)); B = rng.standard_normal((K, N)); C0 = rng.standard_normal((M, N))
    got = pkg.dgemm_tn(At, B, alpha=1.5, beta=-0.5, C_in=C0)
    want = 1.5 * At.T @ B - 0.5 * C0
    assert relerr(got, want) < 1e-13
    got = pkg.dgemm_tn(At, B)
    assert relerr(got, At.T @ B) < 1e-13


def test_dgemm_linearity_large(pkg):
    """size-independent property at a Gram-like shape: (A1+A2)'B == A1'B + A2'B, and vs numpy."""
    rng = np.random.default_rng(5)
    K, M, N = 777, 1275, 900
    A1 = rng.standard_normal((K, M)); A2 = rng.standard_normal((K, M)); B = rng.standard_normal((K, N))
    c12 = pkg.dgemm_tn(A1 + A2, B)
    c1 = pkg.dgemm_tn(A1, B); c2 = pkg.dgemm_tn(A2, B)
    assert relerr(c12, c1 + c2) < 1e-12
    assert relerr(c1, A1.T @ B) < 1e-13


@pytest.fixture(params=["expsum", "exact"])
def gram(request, pkg):
    """Both evaluations of the weighted Gram behind sample_Lambda_c; restores the default."""
    pkg.set_gram_default(request.param)
    yield request.param
    pkg.set_gram_default("expsum")


def _prob(small_problem):
    st = small_problem["state"]
    return (st["data_matrices"], st["run_variables"], st["current_state"], st["priors"],
            small_problem["n"], small_problem["p"], small_problem["k"], small_problem["r"], small_problem["b"])


def test_sample_Lambda_c(pkg, small_problem, gram):
    from oracle import checker as O
    d, rv, cs, pri, n, p, k, r, b = _prob(small_problem)
    rng = np.random.default_rng(11)
    z = rng.standard_normal((k, p))
    Yt = d["Y"] - d["X"] @ cs["B"]
    want = O.sample_Lambda_c(Yt, cs["F"], cs["resid_Y_prec"], cs["E_a_prec"], cs["Plam"], rv["invert_aI_bZAZ"], z)
    got = pkg.sample_Lambda_c(Yt, cs["F"], cs["resid_Y_prec"], cs["E_a_prec"], cs["Plam"], rv["invert_aI_bZAZ"],
                              variates=z)
    assert got.shape == (p, k)
    assert rowwise_relerr(got, want) < TOL
    got2 = pkg.sample_lambda_c(Yt, cs["F"], cs["resid_Y_prec"], cs["E_a_prec"], cs["Plam"], rv["invert_aI_bZAZ"],
                               variates=z)
    assert np.array_equal(got, got2)


@pytest.mark.parametrize("n,p,k", [(33, 17, 3), (100, 257, 20), (250, 100, 40), (64, 50, 70), (96, 24, 128), (80, 21, 136),
                                   (120, 19, 200), (72, 9, 216), (90, 12, 217), (110, 10, 300), (64, 6, 448)])
def test_sample_Lambda_c_shapes(pkg, n, p, k, gram):
    """ragged sizes incl. odd n (padded leading dimensions); k across the 4-warp (k <= 128) and 8-warp (k <= 216, the
    shared-memory limit) instantiations of the blocked Cholesky, incl. k not a multiple of 8, and beyond the shared-memory
    limit (tiles in the L2 workspace: k = 217, 300, 448; update_k may grow k towards 2p, BSFG_functions.R:332)."""
    from oracle import checker as O
    rng = np.random.default_rng(n + p + k)
    A = rng.standard_normal((n, n)); A = A @ A.T / n + np.eye(n) * 0.1
    s, U = np.linalg.eigh(A); s = s[::-1].copy(); U = U[:, ::-1].copy()
    Y = rng.standard_normal((n, p)); F = rng.standard_normal((n, k))
    rp = rng.gamma(2, 1, p) + 0.1; ep = rng.gamma(2, 1, p) + 0.1
    Plam = rng.gamma(2, 1, (p, k)) * np.cumprod(rng.gamma(3, 1, k))[None, :]
    z = rng.standard_normal((k, p))
    inv = {"U": U, "s": s}
    want = O.sample_Lambda_c(Y, F, rp, ep, Plam, inv, z)
    got = pkg.sample_Lambda_c(Y, F, rp, ep, Plam, inv, variates=z)
    assert rowwise_relerr(got, want) < TOL


@pytest.mark.parametrize("spread,terms", [(1e4, 0), (1e9, 0), (1e2, 64), (1e2, 512)])
def test_sample_Lambda_c_expsum_range_guard(pkg, spread, terms):
    """The exponential-sum Gram must never lose accuracy silently: with a precision ratio E_a_prec/resid_Y_prec
    spread over 2*log10(spread) decades, or too few nodes, the device guard falls back to the exact GEMM; either way
    the draw matches the oracle to 1e-9."""
    from oracle import checker as O
    rng = np.random.default_rng(3)
    n, p, k = 96, 40, 9
    A = rng.standard_normal((n, n)); A = A @ A.T / n + np.eye(n) * 0.05
    s, U = np.linalg.eigh(A); s = s[::-1].copy(); U = U[:, ::-1].copy()
    Y = rng.standard_normal((n, p)); F = rng.standard_normal((n, k))
    ep = np.exp(rng.uniform(-np.log(spread), np.log(spread), p)); rp = np.exp(rng.uniform(-np.log(spread), np.log(spread), p))
    ep[0], rp[0] = spread, 1 / spread; ep[1], rp[1] = 1 / spread, spread
    Plam = rng.gamma(2, 1, (p, k)) * np.cumprod(rng.gamma(3, 1, k))[None, :]
    z = rng.standard_normal((k, p))
    inv = {"U": U, "s": s}
    want = O.sample_Lambda_c(Y, F, rp, ep, Plam, inv, z)
    pkg.set_gram_default("expsum", terms=terms)
    try:
        got = pkg.sample_Lambda_c(Y, F, rp, ep, Plam, inv, variates=z)
    finally:
        pkg.set_gram_default("expsum", terms=256)
    assert rowwise_relerr(got, want) < TOL


def test_sample_Lambda_c_not_pd(pkg):
    """Armadillo's chol() throws on a non-SPD Qlam; the C ABI reports BSFG_ENOTPD."""
    from sparsefactormixedmodel_b200 import BSFGError, _lib
    rng = np.random.default_rng(0)
    n, p, k = 20, 5, 3
    U = np.eye(n); s = np.ones(n)
    Plam = -1e6 * np.ones((p, k))
    with pytest.raises(BSFGError) as ei:
        pkg.sample_Lambda_c(rng.standard_normal((n, p)), rng.standard_normal((n, k)), np.ones(p), np.ones(p),
                            Plam, {"U": U, "s": s}, variates=np.zeros((k, p)))
    assert ei.value.code == _lib.BSFG_ENOTPD


def test_sample_means_c(pkg, small_problem):
    from oracle import checker as O
    d, rv, cs, pri, n, p, k, r, b = _prob(small_problem)
    rng = np.random.default_rng(12)
    br = b + r
    z = rng.standard_normal((br, p))
    Yt = d["Y"] - cs["F"] @ cs["Lambda"].T
    L = rv["invert_aPXA_bDesignDesignT"]
    want = O.sample_means_c(Yt, cs["resid_Y_prec"], cs["E_a_prec"], L, z)
    got = pkg.sample_means_c(Yt, cs["resid_Y_prec"], cs["E_a_prec"], L, variates=z)
    assert got.shape == (br, p)
    # column j is one trait's [B;E_a] draw
    assert rowwise_relerr(got.T, want.T) < TOL


def test_sample_factors_scores_c(pkg, small_problem):
    from oracle import checker as O
    d, rv, cs, pri, n, p, k, r, b = _prob(small_problem)
    rng = np.random.default_rng(13)
    z = rng.standard_normal((n, k))
    F_h2 = np.array([0.0, 0.3, 0.5, 0.0, 0.9, 0.02])[:k]
    Yt = d["Y"] - d["X"] @ cs["B"] - d["Z_1"] @ cs["E_a"]
    want = O.sample_factors_scores_c(Yt, d["Z_1"], cs["Lambda"], cs["resid_Y_prec"], cs["F_a"], F_h2, z)
    got = pkg.sample_factors_scores_c(Yt, d["Z_1"], cs["Lambda"], cs["resid_Y_prec"], cs["F_a"], F_h2, variates=z)
    assert rowwise_relerr(got, want) < TOL
    tau_e = 1.0 / (1.0 - F_h2)
    want_px = O.sample_px_factors_scores_c(Yt, d["Z_1"], cs["Lambda"], cs["resid_Y_prec"], cs["F_a"], tau_e,
                                           np.ones(k), z)
    got_px = pkg.sample_px_factors_scores_c(Yt, d["Z_1"], cs["Lambda"], cs["resid_Y_prec"], cs["F_a"], tau_e,
                                            np.ones(k), variates=z)
    assert rowwise_relerr(got_px, want_px) < TOL
    assert relerr(got_px, got) < 1e-12


def _check_grid_draw(got_h2, got_lp, want_h2, want_lp, u, H):
    """exact index unless the draw sits where two cumulative probabilities are within 1e-9."""
    assert relerr(got_lp[np.isfinite(want_lp)], want_lp[np.isfinite(want_lp)]) < TOL
    for j in range(len(want_h2)):
        if got_h2[j] != want_h2[j]:
            row = want_lp[j]
            mx = row.max()
            ps = np.exp(row - (mx + np.log(np.sum(np.exp(row - mx)))))
            edge = np.min(np.abs(np.cumsum(ps) - u[j]))
            assert edge < 1e-9, f"row {j}: index {got_h2[j]} vs {want_h2[j]} with edge distance {edge}"


def test_sample_h2s_discrete_c(pkg, small_problem):
    from oracle import checker as O
    d, rv, cs, pri, n, p, k, r, b = _prob(small_problem)
    rng = np.random.default_rng(14)
    H = 50
    for trial in range(3):
        u = rng.uniform(size=k)
        F = cs["F"] * (1.0 + trial)
        want = O.sample_h2s_discrete_c(F, H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], u)
        want_lp = O.h2s_log_ps(F, H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"])
        got, got_lp = pkg.sample_h2s_discrete_c(F, H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"], variates=u,
                                                return_log_ps=True)
        assert got.shape == (k, 1)
        _check_grid_draw(got.ravel(), got_lp, want, want_lp, u, H)
    # u above the last cumsum edge -> count == H -> F_h2 == 1 (reference behaviour, SURVEY 7 "mirror")
    got = pkg.sample_h2s_discrete_c(cs["F"], H, pri["h2_priors_factors"], rv["invert_aI_bZAZ"],
                                    variates=np.full(k, 1.0 + 1e-9))
    assert np.all(got == 1.0)


@pytest.mark.parametrize("n", [60, 61])
def test_sample_prec_discrete_conditional_c(pkg, n):
    """odd n exercises the C++ integer division n/2 (cpp:274)."""
    from oracle import checker as O
    rng = np.random.default_rng(15 + n)
    p, H = 37, 50
    A = rng.standard_normal((n, n)); A = A @ A.T / n + 0.2 * np.eye(n)
    s, U = np.linalg.eigh(A); s = s[::-1].copy(); U = U[:, ::-1].copy()
    inv = {"U": U, "s": s}
    Y = rng.standard_normal((n, p)) * 1.3
    res_prec = rng.gamma(2, 1, p) + 0.2
    pri = np.concatenate([[0.0], np.ones(H - 1)])      # h2_priors[1] = 0 as BSFG_functions.R:81 advises
    u = rng.uniform(size=p)
    want = O.sample_prec_discrete_conditional_c(Y, H, pri, inv, res_prec, u)
    want_lp = O.prec_log_ps(Y, H, pri, inv, res_prec)
    got, got_lp = pkg.sample_prec_discrete_conditional_c(Y, H, pri, inv, res_prec, variates=u, return_log_ps=True)
    assert got.shape == (p, 1)
    assert np.all(np.isneginf(got_lp[:, 0]))
    want_h2 = res_prec / (want + res_prec)
    got_h2 = res_prec / (got.ravel() + res_prec)
    _check_grid_draw(np.round(got_h2 * H) / H, got_lp, np.round(want_h2 * H) / H, want_lp, u, H)
    same = np.round(got_h2 * H) == np.round(want_h2 * H)
    assert relerr(got.ravel()[same], want[same]) < TOL


def test_sample_F_a_c(pkg, small_problem):
    from oracle import checker as O
    d, rv, cs, pri, n, p, k, r, b = _prob(small_problem)
    rng = np.random.default_rng(16)
    z = rng.standard_normal((r, k))
    F_h2 = np.array([0.0, 0.3, 0.5, 0.99995, 0.9, 0.02])[:k]     # branches: tau_e==1, >10000, regular
    want = O.sample_F_a_c(cs["F"], d["Z_1"], F_h2, rv["invert_aZZt_Ainv"], z)
    got = pkg.sample_F_a_c(cs["F"], d["Z_1"], F_h2, rv["invert_aZZt_Ainv"], variates=z)
    assert np.all(got[:, 0] == 0.0)
    assert np.array_equal(got[:, 3], cs["F"][:, 3])
    assert rowwise_relerr(got.T, want.T) < TOL


def test_sample_F_a_c_rectangular_Z(pkg):
    """Ayroles-shaped incidence matrix: r < n lines, several individuals per line."""
    from oracle import checker as O, init_oracle as I
    rng = np.random.default_rng(17)
    n, r, k = 90, 30, 5
    Z1 = np.zeros((n, r)); Z1[np.arange(n), np.arange(n) % r] = 1.0
    A = I.halfsib_A(6, 5)
    _, _, _, inv3 = I.gsvd_lists_stable(np.ones((n, 1)), Z1, A)
    F = rng.standard_normal((n, k)); z = rng.standard_normal((r, k))
    F_h2 = np.array([0.1, 0.0, 0.5, 0.7, 0.98])
    want = O.sample_F_a_c(F, Z1, F_h2, inv3, z)
    got = pkg.sample_F_a_c(F, Z1, F_h2, inv3, variates=z)
    assert rowwise_relerr(got.T, want.T) < TOL


@pytest.mark.parametrize("k", [2, 3, 6, 20])
def test_sample_delta_c(pkg, k):
    from oracle import checker as O
    rng = np.random.default_rng(18 + k)
    p = 45
    delta = rng.gamma(3, 1, k) + 0.5
    tauh = np.cumprod(delta)
    LP = rng.gamma(2, 1, (p, k)); L2 = rng.standard_normal((p, k)) ** 2
    nd = 3 if k == 2 else k - 1
    g = rng.gamma(50, 1, nd)
    want, shapes, rates = O.sample_delta(delta, tauh, LP, 2.1, 1 / 20, 4.5, 1.0, L2, g)
    got, gshapes, grates = pkg.sample_delta_c(delta, tauh, LP, 2.1, 1 / 20, 4.5, 1.0, L2, variates=g,
                                              return_params=True)
    assert len(grates) == nd
    assert relerr(gshapes, shapes) < 1e-14
    assert relerr(grates, rates) < TOL
    assert relerr(got, want) < TOL


def test_device_rng_moments(pkg, small_problem):
    """variates=None draws on the device (Philox): check the draw is the oracle's mean + L'^-1 z with
    z having N(0,1) moments (no 1:1 comparison possible, as with R's own RNG)."""
    d, rv, cs, pri, n, p, k, r, b = _prob(small_problem)
    Yt = d["Y"] - cs["F"] @ cs["Lambda"].T
    L = rv["invert_aPXA_bDesignDesignT"]
    from oracle import checker as O
    mean = O.sample_means_c(Yt, cs["resid_Y_prec"], cs["E_a_prec"], L, np.zeros((b + r, p)))
    a = pkg.sample_means_c(Yt, cs["resid_Y_prec"], cs["E_a_prec"], L, seed=123)
    a2 = pkg.sample_means_c(Yt, cs["resid_Y_prec"], cs["E_a_prec"], L, seed=123)
    c = pkg.sample_means_c(Yt, cs["resid_Y_prec"], cs["E_a_prec"], L, seed=124)
    assert np.array_equal(a, a2) and not np.array_equal(a, c)
    dmat = np.outer(L["s1"], cs["E_a_prec"]) + np.outer(L["s2"], cs["resid_Y_prec"])
    zhat = np.linalg.solve(L["U"], a - mean) * np.sqrt(dmat)
    assert abs(zhat.mean()) < 0.1 and abs(zhat.std() - 1.0) < 0.05


# ---- update_k_c: the R-only update_k behind the same argument list (BSFG_functions.R:311-372) --------------------------------
def _update_k_case(arm, px=False, rect_Z=False):
    from oracle import init_oracle as I
    n, p, k = 40, 30, 5
    rng = np.random.default_rng(77 + len(arm))
    r = 12 if rect_Z else n
    Z = np.zeros((n, r)); Z[np.arange(n), rng.integers(0, r, n)] = 1.0
    if not rect_Z:
        Z = np.eye(n)
    pri = I.default_priors(k_init=k); rp = I.default_run_parameters()
    cs = dict(Lambda=rng.standard_normal((p, k)), Lambda_prec=rng.gamma(1.5, 1.0, (p, k)), F=rng.standard_normal((n, k)),
              F_a=rng.standard_normal((r, k)), F_h2=rng.integers(0, 40, k) / 50.0, nrun=301)
    if arm == "drop":                         # columns 1 and 3 (0-based) and the LAST one are redundant: R skips red == k
        delta = np.array([1.5, 1.2, 1.1, 1.3, 1.2])
        cs["Lambda"][:, [1, 3, 4]] *= 1e-6
    else:
        delta = np.array([1.5, 1.2, 1.1, 1.3, 1.2])
    cs["delta"] = delta; cs["tauh"] = np.cumprod(delta)
    cs["Plam"] = cs["Lambda_prec"] * cs["tauh"][None, :]
    if px:
        cs["px_factor"] = rng.gamma(2.0, 1.0, k) + 0.05
        cs["F_px"] = cs["F"] * np.sqrt(cs["px_factor"])[None, :]
    v = dict(u_k=np.array(0.9 if arm == "none" else 1e-6), k_g_lp=rng.gamma(pri["Lambda_df"] / 2, 1.0, p),
             k_g_delta=np.array(rng.gamma(pri["delta_2_shape"], 1.0)), k_z_lambda=rng.standard_normal(p),
             k_u_h2=np.array(rng.uniform()), k_z_Fa=rng.standard_normal(r), k_z_F=rng.standard_normal(n),
             k_g_px=np.array(rng.gamma(pri["px_shape"], 1.0)))
    return cs, pri, rp, dict(Z_1=Z), v


@pytest.mark.parametrize("arm", ["add", "drop", "none"])
@pytest.mark.parametrize("px", [False, True])
@pytest.mark.parametrize("rect_Z", [False, True])
def test_update_k_c(pkg, arm, px, rect_Z):
    from oracle import checker as O
    cs, pri, rp, dm, v = _update_k_case(arm, px, rect_Z)
    k = cs["F"].shape[1]
    args = (cs["F"], cs["Lambda"], cs["F_a"], cs["F_h2"], cs["Lambda_prec"], cs["Plam"], cs["delta"], cs["tauh"])
    tail = (dm["Z_1"], pri["Lambda_df"], pri["delta_2_shape"], pri["delta_2_rate"])
    sched = (rp["b0"], rp["b1"], cs["nrun"], rp["epsilon"], rp["prop"], v)
    if px:
        want = O.update_k_px(*args, cs["px_factor"], cs["F_px"], *tail, pri["px_shape"], pri["px_rate"], *sched)
    else:
        want = O.update_k(*args, *tail, *sched)
    k2 = want["F"].shape[1]
    assert {"add": k2 == k + 1, "drop": k2 == 2, "none": k2 == k}[arm], "oracle took another arm; test inputs need retuning"
    got = pkg.update_k_c(cs, pri, rp, dm, variates=v, k_max=k + 2)
    assert got["k"] == k2
    for f in ("Lambda", "Lambda_prec", "Plam", "F", "F_a") + (("F_px",) if px else ()):
        assert got[f].shape == want[f].shape, f
        assert relerr(got[f], want[f]) < 1e-12, f
    for f in ("delta", "tauh", "F_h2") + (("px_factor",) if px else ()):
        assert relerr(got[f], np.asarray(want[f]).ravel()) < 1e-12, f
    if arm == "none":                           # nothing is written: bit-identical pass-through
        assert np.array_equal(got["Plam"], cs["Plam"]) and np.array_equal(got["F"], cs["F"])


def test_update_k_c_philox_and_errors(pkg):
    """NULL variates: the draws come from the (seed, nrun) Philox streams -- reproducible, seed dependent, and the iteration
    gate i > 200 (R:331) holds whatever uu is.  A full k_max is an error, as in the sweep."""
    cs, pri, rp, dm, v = _update_k_case("add")
    k = cs["F"].shape[1]
    early = pkg.update_k_c(dict(cs, nrun=150), pri, rp, dm, variates=dict(u_k=np.array(0.0)))
    assert early["k"] == k
    outs = [pkg.update_k_c(cs, pri, rp, dm, variates=dict(u_k=np.array(0.0)), seed=s) for s in (4, 4, 5)]
    assert all(o["k"] == k + 1 for o in outs)
    assert np.array_equal(outs[0]["Lambda"], outs[1]["Lambda"]) and np.array_equal(outs[0]["F"], outs[1]["F"])
    assert not np.array_equal(outs[0]["Lambda"][:, k], outs[2]["Lambda"][:, k])
    assert np.array_equal(outs[0]["Lambda"][:, :k], cs["Lambda"])
    new = outs[0]
    assert np.isfinite(new["Lambda"]).all() and 0.0 <= new["F_h2"][k] < 1.0 and new["delta"][k] > 0
    assert relerr(new["Plam"], new["Lambda_prec"] * np.cumprod(new["delta"])[None, :]) < 1e-14
    with pytest.raises(pkg.BSFGError):
        pkg.update_k_c(cs, pri, rp, dm, variates=v, k_max=k)


# ---- GSVD_2_c (cpp:27-40) ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 7, 64, 301])
def test_GSVD_2_c_contract(pkg, n):
    """U, V are only unique up to signs (and rotations inside repeated singular values), so the check is the contract the
    reference's callers rely on (fast_BSFG_sampler_init.R:289-325) plus the singular values against the oracle's."""
    from oracle import init_oracle as I
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    B = rng.standard_normal((n, n)) + 3.0 * np.eye(n)
    got = pkg.GSVD_2_c(A, B)
    want = I.GSVD_2_c(A, B)
    U, V, X, Cm, Sm = (got[f] for f in ("U", "V", "X", "C", "S"))
    eye = np.eye(n)
    assert np.abs(U.T @ U - eye).max() < 1e-11 and np.abs(V.T @ V - eye).max() < 1e-11
    assert np.abs(Cm @ Cm + Sm @ Sm - eye).max() < 1e-13
    sc = max(1.0, np.abs(A).max(), np.abs(B).max())
    assert np.abs(U @ Cm @ X.T - A).max() < 1e-10 * sc
    assert np.abs(V @ Sm @ X.T - B).max() < 1e-10 * sc
    assert relerr(np.diag(Cm), np.diag(want["C"])) < 1e-10 and relerr(np.diag(Sm), np.diag(want["S"])) < 1e-10
    c = np.diag(Cm)
    assert np.all(np.diff(c) <= 1e-14)          # svd order: descending d, hence descending C


def test_GSVD_2_c_gives_the_init_lists(pkg):
    """The way the reference uses it (init:289-295): U = t(solve(X)), s1 = diag(C)^2, s2 = diag(S)^2 diagonalise both forms."""
    from oracle import init_oracle as I
    rng = np.random.default_rng(5)
    n, r = 30, 12
    Z = np.zeros((n, r)); Z[np.arange(n), np.arange(n) % r] = 1.0
    G = rng.standard_normal((r, r)); Ainv = G @ G.T / r + np.eye(r)
    ZZ = Z.T @ Z
    res = pkg.GSVD_2_c(I.cholcov(ZZ), I.cholcov(Ainv))
    U = np.linalg.inv(res["X"]).T
    s1, s2 = np.diag(res["C"]) ** 2, np.diag(res["S"]) ** 2
    assert np.abs(U.T @ ZZ @ U - np.diag(s1)).max() < 1e-9
    assert np.abs(U.T @ Ainv @ U - np.diag(s2)).max() < 1e-9
    a, b = 0.7, 2.3
    assert relerr(U @ np.diag(1.0 / (a * s1 + b * s2)) @ U.T, np.linalg.inv(a * ZZ + b * Ainv)) < 1e-9


def test_GSVD_2_c_singular_B_is_an_error(pkg):
    with pytest.raises(pkg.BSFGError):
        pkg.GSVD_2_c(np.eye(4), np.zeros((4, 4)))
