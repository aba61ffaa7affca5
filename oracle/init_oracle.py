"""Init restatement + synthetic inputs -- TEST INFRASTRUCTURE ONLY (see oracle/bsfg_oracle.py header).

`sampler_init` follows `/root/reference/R_BSFG/fast_BSFG_sampler_init.R:63-345`
(with `cholcov` = `R_BSFG/BSFG_functions.R:15-21` and `GSVD_2_c` = `BSFG_functions_c.cpp:27-40`).
R's RNG cannot be reproduced, so prior draws use a seeded NumPy Generator; the *structure* (shapes,
rates, `byrow` fills) is the reference's.  `default_priors` / `default_run_parameters` are the
literals of `R_BSFG/model_setup.R:36-63`.

`make_synthetic` builds the inputs of BASELINE.json's configs C2-C5 following the recipe of
`R_BSFG/Runcie_Mukherjee_analyses/Simulations/generate_simulations_halfSib.R:20-37, 72-89`
(sparse Lambda, G/R factor form, Y = L_A' Z G_chol + Z R_chol, X = 1_n, Z_1 = I), with either the
half-sib A of the fixture (`:62-66`) or a random three-generation pedigree (tabular method).
"""
from __future__ import annotations

import numpy as np

from .bsfg_oracle import GSVD_2_c


def default_run_parameters():
    # model_setup.R:36-46
    return dict(b0=1.0, b1=0.0005, epsilon=1e-2, prop=1.00, h2_divisions=50, save_freq=100,
                burn=2000, thin=400, draw_iter=200)


def default_priors(k_init=20, h2_divisions=50):
    # model_setup.R:49-63
    H = h2_divisions
    return dict(k_init=k_init, resid_Y_prec_shape=2.0, resid_Y_prec_rate=1 / 10, E_a_prec_shape=2.0,
                E_a_prec_rate=1 / 10, W_prec_shape=2.0, W_prec_rate=1 / 10, Lambda_df=3.0,
                delta_1_shape=2.1, delta_1_rate=1 / 20, delta_2_shape=4.5, delta_2_rate=1.0,
                h2_priors_factors=np.concatenate([[H - 1.0], np.ones(H - 1)]) / (2.0 * (H - 1)),
                px_shape=0.5, px_rate=0.5)                               # model_setup_px.R:55-56


def cholcov(X):
    """BSFG_functions.R:15-21: U with t(U) %*% U == X via the SVD (keeps all columns)."""
    u, d, _ = np.linalg.svd(X)
    return (u @ np.diag(np.sqrt(d))).T


def block_diag2(a, b):
    out = np.zeros((a.shape[0] + b.shape[0], a.shape[1] + b.shape[1]))
    out[:a.shape[0], :a.shape[1]] = a
    out[a.shape[0]:, a.shape[1]:] = b
    return out


def gsvd_lists_reference(X, Z_1, A, b_X_prec=1e-10):
    """The three diagonalisations exactly as fast_BSFG_sampler_init.R:263-325 computes them."""
    b = X.shape[1]
    Ainv = np.linalg.solve(A, np.eye(A.shape[0]))                       # init:263
    u, d, _ = np.linalg.svd(Z_1 @ A @ Z_1.T)                            # init:278
    inv1 = {"U": u, "s": d}
    Design = np.column_stack([X, Z_1])                                  # init:287
    Design2 = Design.T @ Design
    res = GSVD_2_c(cholcov(block_diag2(b_X_prec * np.eye(b), Ainv)), cholcov(Design2))   # init:289
    U2 = np.linalg.solve(res["X"], np.eye(res["X"].shape[0])).T         # t(solve(X)), init:291
    inv2 = {"U": U2, "s1": np.diag(res["C"]) ** 2, "s2": np.diag(res["S"]) ** 2}
    inv2["Design_U"] = Design @ U2                                      # init:295
    ZZt = Z_1.T @ Z_1                                                   # init:319
    res = GSVD_2_c(cholcov(ZZt), cholcov(Ainv))                         # init:320
    U3 = np.linalg.solve(res["X"], np.eye(res["X"].shape[0])).T
    inv3 = {"U": U3, "s1": np.diag(res["C"]) ** 2, "s2": np.diag(res["S"]) ** 2}
    return Ainv, inv1, inv2, inv3


def _joint_diag(P, D2):
    """U with U' P U = diag(s1), U' D2 U = diag(s2), s1 + s2 = 1 -- the contract of the lists built at
    fast_BSFG_sampler_init.R:289-294, computed through the SPD pencil (D2, P + D2) instead of
    `svd(A inv(B))`, which is singular when [X Z_1] is rank deficient."""
    L = np.linalg.cholesky(P + D2)
    Li = np.linalg.solve(L, np.eye(L.shape[0]))
    M = Li @ D2 @ Li.T
    M = 0.5 * (M + M.T)
    lam, V = np.linalg.eigh(M)
    lam = np.clip(lam, 0.0, 1.0)
    U = Li.T @ V
    return U, 1.0 - lam, lam


def gsvd_lists_stable(X, Z_1, A, b_X_prec=1e-10):
    """Same three lists (same contract: inv(a P + b D'D) = U diag(1/(a s1 + b s2)) U'), computed by
    symmetric eigen-decompositions.  Used for the large synthetic configs, where the lists are
    *inputs* handed identically to the oracle and to the device path."""
    b = X.shape[1]
    Ainv = np.linalg.inv(A)
    Ainv = 0.5 * (Ainv + Ainv.T)
    ZAZ = Z_1 @ A @ Z_1.T
    s, u = np.linalg.eigh(0.5 * (ZAZ + ZAZ.T))
    order = np.argsort(-s)
    inv1 = {"U": np.ascontiguousarray(u[:, order]), "s": np.clip(s[order], 0.0, None)}
    Design = np.column_stack([X, Z_1])
    U2, s1, s2 = _joint_diag(block_diag2(b_X_prec * np.eye(b), Ainv), Design.T @ Design)
    inv2 = {"U": U2, "s1": s1, "s2": s2, "Design_U": Design @ U2}
    U3, s2_3, s1_3 = _joint_diag(Ainv, Z_1.T @ Z_1)      # here s1 <-> Z'Z, s2 <-> Ainv (init:315-320)
    inv3 = {"U": U3, "s1": s1_3, "s2": s2_3}
    return Ainv, inv1, inv2, inv3


def sampler_init(setup, priors, run_parameters, seed=1, lists="reference"):
    """Restates fast_BSFG_sampler_init.R: returns a BSFG_state-shaped dict of dicts.

    setup: dict with Y (n,p), Z_1 (n,r), A (r,r), optional X (n,b) (or (1,n), transposed like init:90),
    optional `simulation` flag (True keeps Y unstandardised, init:76-81)."""
    rng = np.random.default_rng(seed)
    Y = np.array(setup["Y"], dtype=np.float64)
    Z_1 = np.array(setup["Z_1"], dtype=np.float64)
    A = np.array(setup["A"], dtype=np.float64)
    n, p = Y.shape
    r = Z_1.shape[1]
    simulation = bool(setup.get("simulation", "gen_factor_Lambda" in setup))
    Y_missing = np.isnan(Y)
    Mean_Y = np.nanmean(Y, axis=0)
    VY = np.nanvar(Y, axis=0, ddof=1)
    if simulation:
        VY = np.ones(p)
    else:
        Y = (Y - Mean_Y[None, :]) / np.sqrt(VY)[None, :]
    X = setup.get("X")
    if X is None:
        X = np.zeros((n, 0))
    X = np.array(X, dtype=np.float64)
    if X.shape[0] == 1 and n != 1:
        X = X.T                                                         # init:90
    assert X.shape[0] == n
    b = X.shape[1]
    Z_2 = setup.get("Z_2")
    Z_2 = np.zeros((n, 0)) if Z_2 is None else np.array(Z_2, dtype=np.float64)
    if Z_2.shape[0] != n:
        Z_2 = Z_2.T
    assert Z_2.shape[0] == n                                            # init:100
    r2 = Z_2.shape[1]
    priors = dict(priors)
    priors["b_X_prec"] = 1e-10 * np.eye(b)                              # init:114
    k = int(priors["k_init"])
    g = lambda shape, rate, size: rng.gamma(shape, 1.0, size=size) / rate
    resid_Y_prec = g(priors["resid_Y_prec_shape"], priors["resid_Y_prec_rate"], p)
    df = priors["Lambda_df"]
    Lambda_prec = g(df / 2, df / 2, (p, k))
    delta = np.concatenate([g(priors["delta_1_shape"], priors["delta_1_rate"], 1),
                            g(priors["delta_2_shape"], priors["delta_2_rate"], k - 1)])
    tauh = np.cumprod(delta)
    Plam = Lambda_prec * tauh[None, :]
    Lambda = rng.standard_normal((p, k)) * np.sqrt(1.0 / Plam)
    E_a_prec = g(priors["E_a_prec_shape"], priors["E_a_prec_rate"], p)
    E_a = rng.standard_normal((r, p)) * np.sqrt(1.0 / E_a_prec)[None, :]      # byrow fill, init:178
    F_h2 = rng.uniform(size=k)
    F_a = rng.standard_normal((r, k)) * np.sqrt(F_h2)[None, :]
    F = Z_1 @ F_a + rng.standard_normal((n, k)) * np.sqrt(1 - F_h2)[None, :]
    W_prec = g(priors["W_prec_shape"], priors["W_prec_rate"], p)
    W = rng.standard_normal((r2, p)) * np.sqrt(1.0 / W_prec)[None, :]          # byrow fill, init:210
    B = rng.standard_normal((b, p))
    current_state = dict(resid_Y_prec=resid_Y_prec, Lambda_prec=Lambda_prec, delta=delta, tauh=tauh,
                         Plam=Plam, Lambda=Lambda, F_h2=F_h2, E_a_prec=E_a_prec, W_prec=W_prec,
                         F_a=F_a, F=F, E_a=E_a, B=B, W=W, nrun=0)
    fn = gsvd_lists_reference if lists == "reference" else gsvd_lists_stable
    Ainv, inv1, inv2, inv3 = fn(X, Z_1, A)
    inv_rand2 = {}
    if r2 > 0:                                                          # init:300-311, A_2_inv = I (init:264)
        if lists == "reference":
            res = GSVD_2_c(cholcov(np.eye(r2)), cholcov(Z_2.T @ Z_2))
            Uw = np.linalg.solve(res["X"], np.eye(r2)).T
            inv_rand2 = {"U": Uw, "s1": np.diag(res["C"]) ** 2, "s2": np.diag(res["S"]) ** 2}
        else:
            Uw, s1w, s2w = _joint_diag(np.eye(r2), Z_2.T @ Z_2)
            inv_rand2 = {"U": Uw, "s1": s1w, "s2": s2w}
        inv_rand2["Design_U"] = Z_2 @ inv_rand2["U"]
    run_variables = dict(p=p, n=n, r=r, r2=r2, b=b, Mean_Y=Mean_Y, VY=VY, Ainv=Ainv,
                         A_2_inv=np.eye(r2), invert_aI_bZAZ=inv1, invert_aPXA_bDesignDesignT=inv2,
                         invert_aZZt_Ainv=inv3, invert_aPXA_bDesignDesignT_rand2=inv_rand2)
    rp = dict(run_parameters)
    rp["simulation"] = simulation
    return dict(data_matrices=dict(Y=Y, Z_1=Z_1, Z_2=Z_2, X=X, Y_missing=Y_missing),
                run_parameters=rp, run_variables=run_variables, priors=priors,
                current_state=current_state)


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def halfsib_A(n_sires, n_per_sire):
    """A of a paternal half-sib design: 1 on the diagonal, 0.25 within a sire family
    (generate_simulations_halfSib.R:62-66; equals the Sim_1 fixture's A)."""
    fam = np.repeat(np.arange(n_sires), n_per_sire)
    A = 0.25 * (fam[:, None] == fam[None, :]).astype(np.float64)
    np.fill_diagonal(A, 1.0)
    return A


def random_pedigree_A(n, rng, frac_founders=0.2, frac_gen1=0.3):
    """Numerator-relationship matrix of a random 3-generation pedigree (tabular method,
    vectorised per generation).  Gives ~n distinct eigenvalues."""
    n0 = max(2, int(n * frac_founders))
    n1 = max(2, int(n * frac_gen1))
    n2 = n - n0 - n1
    A = np.zeros((n, n))
    A[:n0, :n0] = np.eye(n0)
    lo = n0
    prev = (0, n0)
    for cnt in (n1, n2):
        if cnt <= 0:
            continue
        hi = lo + cnt
        sire = rng.integers(prev[0], prev[1], size=cnt)
        dam = rng.integers(prev[0], prev[1], size=cnt)
        same = sire == dam
        dam[same] = prev[0] + (dam[same] - prev[0] + 1) % (prev[1] - prev[0])
        A[lo:hi, :lo] = 0.5 * (A[sire, :lo] + A[dam, :lo])
        A[:lo, lo:hi] = A[lo:hi, :lo].T
        blk = 0.5 * (A[lo:hi, :lo][:, sire] + A[lo:hi, :lo][:, dam])
        blk = 0.5 * (blk + blk.T)
        np.fill_diagonal(blk, 1.0 + 0.5 * A[sire, dam])
        A[lo:hi, lo:hi] = blk
        prev = (lo, hi)
        lo = hi
    return A


def make_synthetic(n, p, k_true, seed=20131, pedigree="random", b=1):
    """Synthetic setup dict (Y, X, Z_1, A + truth), recipe of generate_simulations_halfSib.R:20-37,72-89."""
    rng = np.random.default_rng(seed)
    if pedigree == "halfsib":
        per = 5
        assert n % per == 0
        A = halfsib_A(n // per, per)
    else:
        A = random_pedigree_A(n, rng)
    factor_h2s = np.linspace(1.0, 0.0, k_true + 1)[:k_true].copy()          # seq(1,0,...)[1:k]
    factor_h2s[0::2] = 0.0                                                   # :51-52
    Lambda = np.zeros((p, k_true))
    lo, hi = max(1, p // 30), max(2, p // 4)
    for h in range(k_true):
        numeff = min(p, int(rng.integers(lo, hi + 1)))
        rows = rng.choice(p, size=numeff, replace=False)
        Lambda[rows, h] = rng.standard_normal(numeff)
    L_A = np.linalg.cholesky(A)
    # factor form of U_act = t(A_chol) Z G_chol, E_act = Z R_chol with G = L diag(h2) L' + 0.2 I (:83-84, :26-28)
    U_act = L_A @ (rng.standard_normal((n, k_true)) @ (Lambda * np.sqrt(factor_h2s)[None, :]).T
                   + np.sqrt(0.2) * rng.standard_normal((n, p)))
    E_act = (rng.standard_normal((n, k_true)) @ (Lambda * np.sqrt(1 - factor_h2s)[None, :]).T
             + np.sqrt(0.2) * rng.standard_normal((n, p)))
    Y = U_act + E_act
    X = np.ones((n, b))
    if b > 1:
        X[:, 1:] = rng.standard_normal((n, b - 1))
    return dict(Y=Y, X=X, Z_1=np.eye(n), A=A, simulation=True, factor_h2s=factor_h2s,
                error_factor_Lambda=Lambda, n=n, p=p, r=n)


def draw_variates(rng, n, p, k, r, br, priors, add_arm=False, r2=0, missing=False, px=False):
    """One iteration's worth of injected variates in the layouts of oracle.bsfg_oracle.gibbs_iteration."""
    df = priors["Lambda_df"]
    v = dict(z_Lambda=rng.standard_normal((k, p)), z_means=rng.standard_normal((br, p)),
             u_h2=rng.uniform(size=k), z_Fa=rng.standard_normal((r, k)), z_F=rng.standard_normal((n, k)),
             g_Lambda_prec=rng.gamma((df + 1) / 2.0, 1.0, size=(p, k)),
             g_resid=rng.gamma(priors["resid_Y_prec_shape"] + n / 2.0, 1.0, size=p),
             g_Ea=rng.gamma(priors["E_a_prec_shape"] + r / 2.0, 1.0, size=p),
             u_k=float(rng.uniform()))
    from .bsfg_oracle import delta_loop_indices
    shapes = [priors["delta_1_shape"] + 0.5 * p * k]
    shapes += [priors["delta_2_shape"] + 0.5 * p * (k - h + 1) for h in delta_loop_indices(k)]
    v["g_delta"] = np.array([rng.gamma(s, 1.0) for s in shapes])
    if missing:
        v["z_Y"] = rng.standard_normal((n, p))
    if px:
        v["g_px"] = rng.gamma(priors["px_shape"] + n / 2.0, 1.0, size=k)
        if add_arm:
            v["k_g_px"] = float(rng.gamma(priors["px_shape"], 1.0))
    if r2 > 0:
        v["z_W"] = rng.standard_normal((r2, p))
        v["g_W"] = rng.gamma(priors["W_prec_shape"] + r2 / 2.0, 1.0, size=p)
    if add_arm:
        v.update(k_g_lp=rng.gamma(df / 2.0, 1.0, size=p),
                 k_g_delta=float(rng.gamma(priors["delta_2_shape"], 1.0)),
                 k_z_lambda=rng.standard_normal(p), k_u_h2=float(rng.uniform()),
                 k_z_Fa=rng.standard_normal(r), k_z_F=rng.standard_normal(n))
    return v
