"""CPU oracle for the BSFG Gibbs sweep -- TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the reference's conditional samplers
(`/root/reference/R_BSFG/BSFG_functions_c.cpp`, the Rcpp `_c` functions) and of the
R-only steps of the sweep (`R_BSFG/BSFG_functions.R:272-372`, `R_BSFG/fast_BSFG_sampler.R:176-270`).
It exists so the CUDA path can be checked; nothing under `sparsefactormixedmodel_b200/`
imports it.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
use it.

PARITY PINNED (round 2): `oracle/_ref/libbsfg_reference.so` is the reference's own `BSFG_functions_c.cpp`,
compiled UNMODIFIED from `/root/reference` against a stand-in `RcppArmadillo.h` (`oracle/ref_shim/`, SciPy's
OpenBLAS/LAPACK underneath; `oracle/ref_oracle.py` binds it).  tests/test_oracle.py asserts that every `_c` function
below equals that library to <= 1e-12 (1e-10 row-wise for the two Cholesky-solve draws) on the reference's three
fixtures at full size and on random edge-shaped problems, and tests/golden/*.npz hold the LIBRARY's outputs.
The R-only steps (`sample_delta`, `update_k`, the loop body) cannot be run (no R here); they stay pinned by
  * a second, independent transcription of the R twins (`oracle/bsfg_oracle_rtwin.py`, following
    `R_BSFG/BSFG_functions.R:23-269`) and a third of the MATLAB originals (`oracle/bsfg_oracle_mtwin.py`);
  * closed-form identities (posterior mean/covariance of each Gaussian draw recomputed by dense
    inversion) on small cases.

Conventions
  * every array is float64; matrices are handled as NumPy 2-D arrays with the same (rows, cols)
    as the R/Armadillo object; "column-major fill" of a variate block means
    `z.reshape(rows, cols, order="F")` of the flat stream, exactly like `reshape(rnorm(k*p),k,p)`
    (`BSFG_functions_c.cpp:111-114`).
  * every random draw is an explicit argument ("injected variates"): N(0,1) blocks `z_*`,
    U(0,1) blocks `u_*`, and standard Gamma(shape,1) blocks `g_*` (a Gamma(shape, rate) draw is
    `g/rate`; the reference's shapes never depend on the data so this is a complete injection).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import solve_triangular

LOG_INV_SQRT_2PI = np.log(1.0 / np.sqrt(2.0 * np.pi))  # BSFG_functions_c.cpp:23


# --------------------------------------------------------------------------------------
# helpers (BSFG_functions_c.cpp:10-24)
# --------------------------------------------------------------------------------------
def sweep_times(x, margin, stats):
    """`sweep_times` (BSFG_functions_c.cpp:10-20): MARGIN 1 scales rows, MARGIN 2 scales columns."""
    stats = np.asarray(stats, dtype=np.float64).ravel()
    if margin == 1:
        return x * stats[:, None]
    return x * stats[None, :]


def dnorm_std_log(x):
    """`dnorm_std_log` (BSFG_functions_c.cpp:22-24)."""
    return LOG_INV_SQRT_2PI - 0.5 * (x * x)


def chol_upper(Q):
    """Armadillo `chol(Q)`: upper-triangular R with R'R = Q; throws when Q is not SPD."""
    try:
        return np.linalg.cholesky(Q).T
    except np.linalg.LinAlgError as e:  # pragma: no cover - mirrors arma's runtime_error
        raise RuntimeError("chol(): decomposition failed") from e


# --------------------------------------------------------------------------------------
# GSVD_2_c (BSFG_functions_c.cpp:27-40)
# --------------------------------------------------------------------------------------
def GSVD_2_c(A, B):
    n = B.shape[1]
    U, d, Vt = np.linalg.svd(A @ np.linalg.inv(B))
    V = Vt.T
    norm_factor = np.sqrt(np.ones(n) + d * d)
    C = np.diag(d / norm_factor)
    S = np.diag(1.0 / norm_factor)
    X = sweep_times(B.T @ V, 2, norm_factor)
    return {"U": U, "V": V, "X": X, "C": C, "S": S}


# --------------------------------------------------------------------------------------
# sample_means_c (BSFG_functions_c.cpp:43-82)
# --------------------------------------------------------------------------------------
def sample_means_c(Y_tilde, resid_Y_prec, E_a_prec, inv_list, z):
    """z: (br, p) N(0,1), column j = trait j (`randn(br,p)`, cpp:68)."""
    U, s1, s2, Design_U = inv_list["U"], inv_list["s1"], inv_list["s2"], inv_list["Design_U"]
    p = Y_tilde.shape[1]
    br = Design_U.shape[1]
    means = sweep_times(Design_U.T @ Y_tilde, 2, resid_Y_prec)          # cpp:65
    out = np.zeros((br, p))
    for j in range(p):                                                   # cpp:75-79
        d = s1 * E_a_prec[j] + s2 * resid_Y_prec[j]
        mlam = means[:, j] / d
        out[:, j] = U @ (mlam + z[:, j] / np.sqrt(d))
    return out


# --------------------------------------------------------------------------------------
# sample_Lambda_c (BSFG_functions_c.cpp:86-135)
# --------------------------------------------------------------------------------------
def sample_Lambda_c(Y_tilde, F, resid_Y_prec, E_a_prec, Plam, inv_list, z):
    """z: (k, p) N(0,1), column j = trait j (`randn(k,p)`, cpp:110). Returns Lambda (p, k)."""
    U, s = inv_list["U"], inv_list["s"]
    p = resid_Y_prec.shape[0]
    k = F.shape[1]
    FtU = F.T @ U                                                        # cpp:107
    UtY = U.T @ Y_tilde                                                  # cpp:108
    Lambda = np.zeros((p, k))
    for j in range(p):                                                   # cpp:119-132
        FUDi = E_a_prec[j] * sweep_times(FtU, 2, 1.0 / (s + E_a_prec[j] / resid_Y_prec[j]))
        means = FUDi @ UtY[:, j]
        Qlam = FUDi @ FtU.T + np.diag(Plam[j, :])
        R = chol_upper(Qlam)                                             # cpp:124
        vlam = solve_triangular(R.T, means, lower=True)                  # cpp:125
        mlam = solve_triangular(R, vlam, lower=False)                    # cpp:126
        ylam = solve_triangular(R, z[:, j], lower=False)                 # cpp:128
        Lambda[j, :] = ylam + mlam                                       # cpp:130
    return Lambda


# --------------------------------------------------------------------------------------
# sample_factors_scores_c / sample_px_factors_scores_c (cpp:138-163 / 166-193)
# --------------------------------------------------------------------------------------
def _factor_scores(Y_tilde, Z_1, Lambda, resid_Y_prec, F_a, tau_e, z):
    Lmsg = sweep_times(Lambda, 1, resid_Y_prec)                          # cpp:146
    S = chol_upper(Lambda.T @ Lmsg + np.diag(tau_e))                     # cpp:149
    rhs = Y_tilde @ Lmsg + sweep_times(Z_1 @ F_a, 2, tau_e)              # cpp:152
    Meta = solve_triangular(S.T, rhs.T, lower=True).T
    F = solve_triangular(S, (Meta + z).T, lower=False).T                 # cpp:160
    return F


def sample_factors_scores_c(Y_tilde, Z_1, Lambda, resid_Y_prec, F_a, F_h2, z):
    """z: (n, k) N(0,1) (`randn(Meta.n_rows, Meta.n_cols)`, cpp:154)."""
    with np.errstate(divide="ignore"):
        tau_e = 1.0 / (1.0 - np.asarray(F_h2, dtype=np.float64).ravel())  # cpp:147
    return _factor_scores(Y_tilde, Z_1, Lambda, resid_Y_prec, F_a, tau_e, z)


def sample_px_factors_scores_c(Y_tilde, Z_1, Lambda_px, resid_Y_prec, F_a_px, tau_e, px_factor, z):
    """`px_factor` is accepted and unused, as in the reference (cpp:172-176)."""
    return _factor_scores(Y_tilde, Z_1, Lambda_px, resid_Y_prec, F_a_px,
                          np.asarray(tau_e, dtype=np.float64).ravel(), z)


# --------------------------------------------------------------------------------------
# discrete-grid draws (cpp:196-238, 241-289)
# --------------------------------------------------------------------------------------
def _draw_from_log_ps(log_ps_row, u, h2_divisions):
    """cpp:228-235: LSE-normalise, `selected = find(u > cumsum(ps))`, value = count / H."""
    mx = np.max(log_ps_row)
    norm_factor = mx + np.log(np.sum(np.exp(log_ps_row - mx)))
    ps = np.exp(log_ps_row - norm_factor)
    count = int(np.sum(u > np.cumsum(ps)))
    return count / float(h2_divisions), ps


def h2s_log_ps(F, h2_divisions, h2_priors, inv_list):
    """The (k, H) table of unnormalised log posteriors of cpp:206-226 (exposed for tie checks)."""
    U, s = inv_list["U"], inv_list["s"]
    k = F.shape[1]
    log_ps = np.zeros((k, h2_divisions))
    std_scores_b = F.T @ U                                               # cpp:209
    with np.errstate(divide="ignore"):
        log_prior = np.log(np.asarray(h2_priors, dtype=np.float64))
    for i in range(h2_divisions):
        h2 = float(i) / h2_divisions
        if h2 > 0:
            std_scores = 1.0 / np.sqrt(h2) * sweep_times(std_scores_b, 2, 1.0 / np.sqrt(s + (1 - h2) / h2))
            det = np.sum(np.log((s + (1 - h2) / h2) * h2) / 2.0) * np.ones(k)
        else:
            std_scores = F.T
            det = np.zeros(k)
        log_ps[:, i] = np.sum(dnorm_std_log(std_scores), axis=1) - det + log_prior[i]
    return log_ps


def sample_h2s_discrete_c(F, h2_divisions, h2_priors, inv_list, u):
    """u: (k,) U(0,1), one per factor in factor order (`randu(1)` inside the j-loop, cpp:232)."""
    log_ps = h2s_log_ps(F, h2_divisions, h2_priors, inv_list)
    k = F.shape[1]
    F_h2 = np.zeros(k)
    for j in range(k):
        F_h2[j], _ = _draw_from_log_ps(log_ps[j, :], u[j], h2_divisions)
    return F_h2


def prec_log_ps(Y, h2_divisions, h2_priors, inv_list, res_prec):
    """(p, H) table of cpp:260-277, including the C++ integer division `n/2` (cpp:274)."""
    U, s = inv_list["U"], inv_list["s"]
    n, p = Y.shape
    log_ps = np.zeros((p, h2_divisions))
    std_scores_b = Y.T @ U                                               # cpp:258
    with np.errstate(divide="ignore"):
        log_prior = np.log(np.asarray(h2_priors, dtype=np.float64))
    for i in range(h2_divisions):
        h2 = float(i) / h2_divisions
        if h2 > 0:
            std_scores = sweep_times(sweep_times(std_scores_b, 2, 1.0 / np.sqrt(s + (1 - h2) / h2)),
                                     1, 1.0 / np.sqrt(h2 / (res_prec * (1 - h2))))
            det_mat = np.log(np.outer(s + (1 - h2) / h2, h2 / (res_prec * (1 - h2)))) / 2.0  # n x p
            det = np.sum(det_mat, axis=0)
        else:
            std_scores = sweep_times(Y.T, 1, np.sqrt(res_prec))
            det = (n // 2) * np.log(1.0 / res_prec)                      # integer division, cpp:274
        log_ps[:, i] = np.sum(dnorm_std_log(std_scores), axis=1) - det + log_prior[i]
    return log_ps


def sample_prec_discrete_conditional_c(Y, h2_divisions, h2_priors, inv_list, res_prec, u):
    """u: (p,) U(0,1), one per trait (cpp:282). Returns Prec (p,) = res_prec (1-h2)/h2 (cpp:286)."""
    log_ps = prec_log_ps(Y, h2_divisions, h2_priors, inv_list, res_prec)
    p = Y.shape[1]
    Trait_h2 = np.zeros(p)
    for j in range(p):
        Trait_h2[j], _ = _draw_from_log_ps(log_ps[j, :], u[j], h2_divisions)
    with np.errstate(divide="ignore"):
        return (res_prec * (1 - Trait_h2)) / Trait_h2


# --------------------------------------------------------------------------------------
# sample_F_a_c (cpp:293-339)
# --------------------------------------------------------------------------------------
def sample_F_a_c(F, Z_1, F_h2, inv_list, z):
    """z: (r, k) N(0,1) = `reshape(rnorm(r*k), r, k)` (cpp:316-319); consumed for all k columns."""
    U, s1, s2 = inv_list["U"], inv_list["s1"], inv_list["s2"]
    k = F.shape[1]
    r = Z_1.shape[1]
    F_h2 = np.asarray(F_h2, dtype=np.float64).ravel()
    with np.errstate(divide="ignore"):
        tau_e = 1.0 / (1.0 - F_h2)
        tau_u = 1.0 / F_h2
    b = U.T @ Z_1.T @ sweep_times(F, 2, tau_e)                           # cpp:311
    F_a = np.zeros((r, k))
    for j in range(k):                                                   # cpp:323-336
        if tau_e[j] == 1:
            F_a[:, j] = 0.0
        elif tau_e[j] > 10000:
            F_a[:, j] = F[:, j]
        else:
            d = s2 * tau_u[j] + s1 * tau_e[j]
            mlam = b[:, j] / d
            F_a[:, j] = U @ (mlam + z[:, j] / np.sqrt(d))
    return F_a


# --------------------------------------------------------------------------------------
# R-only steps: sample_delta (BSFG_functions.R:272-308), update_k (BSFG_functions.R:311-372)
# --------------------------------------------------------------------------------------
def delta_loop_indices(k):
    """1-based h values visited by R's `for(h in 2:(k-1))` (BSFG_functions.R:296).

    R's `2:(k-1)` counts DOWN when k-1 < 2: k=2 gives (2, 1).  k<2 indexes out of range in R."""
    if k < 2:
        raise ValueError("sample_delta needs k >= 2 (R's 2:(k-1) runs off the vector otherwise)")
    hi = k - 1
    return list(range(2, hi + 1)) if hi >= 2 else list(range(2, hi - 1, -1))


def sample_delta(delta, tauh, Lambda_prec, delta_1_shape, delta_1_rate, delta_2_shape, delta_2_rate,
                 Lambda2, g):
    """g: standard-gamma variates, g[0] ~ Gamma(shape_1,1) then one per visited h, in loop order.

    Returns (delta, shapes, rates): shapes/rates are the Gamma parameters of every draw in order,
    which is what the device path is compared on when draws come from device Philox."""
    delta = np.array(delta, dtype=np.float64).ravel().copy()
    tauh = np.array(tauh, dtype=np.float64).ravel().copy()
    k = tauh.shape[0]
    mat = Lambda_prec * Lambda2
    n_genes = mat.shape[0]
    colsum = np.sum(mat, axis=0)
    shapes, rates = [], []
    shape = delta_1_shape + 0.5 * n_genes * k
    rate = delta_1_rate + 0.5 * (1.0 / delta[0]) * np.sum(tauh * colsum)
    shapes.append(shape); rates.append(rate)
    delta[0] = g[0] / rate
    tauh = np.cumprod(delta)
    for t, h in enumerate(delta_loop_indices(k)):
        shape = delta_2_shape + 0.5 * n_genes * (k - h + 1)
        # h < k always holds on this loop range; the `else` arm of BSFG_functions.R:298-302 is
        # only reachable for k == 2 and evaluates to the same number.
        rate = delta_2_rate + 0.5 * (1.0 / delta[h - 1]) * np.sum(tauh[h - 1:] * colsum[h - 1:])
        shapes.append(shape); rates.append(rate)
        delta[h - 1] = g[1 + t] / rate
        tauh = np.cumprod(delta)
    return delta, np.array(shapes), np.array(rates)


def update_k(F, Lambda, F_a, F_h2, Lambda_prec, Plam, delta, tauh, Z_W, Lambda_df, delta_2_shape,
             delta_2_rate, b0, b1, i, epsilon, prop, variates):
    """variates: dict with `u_k` (scalar, drawn every iteration, BSFG_functions.R:325) and, for the
    add-a-factor arm, in R's draw order (`:334-341`): `k_g_lp` (p) ~ Gamma(df/2,1), `k_g_delta` scalar ~
    Gamma(delta_2_shape,1), `k_z_lambda` (p), `k_u_h2` scalar, `k_z_Fa` (r), `k_z_F` (n)."""
    n, k = F.shape
    r = F_a.shape[0]
    p = Lambda.shape[0]
    F_h2 = np.asarray(F_h2, dtype=np.float64).ravel().copy()
    delta = np.asarray(delta, dtype=np.float64).ravel().copy()
    tauh = np.asarray(tauh, dtype=np.float64).ravel().copy()
    prob = 1.0 / np.exp(b0 + b1 * i)
    uu = variates["u_k"]
    lind = np.mean(np.abs(Lambda) < epsilon, axis=0)
    vec = lind >= prop
    num = int(np.sum(vec))
    if uu < prob and i > 200:
        if i > 20 and num == 0 and np.all(lind < 0.995) and k < 2 * p:
            k = k + 1
            Lambda_prec = np.column_stack([Lambda_prec, variates["k_g_lp"] / (Lambda_df / 2.0)])
            delta = np.append(delta, variates["k_g_delta"] / delta_2_rate)
            tauh = np.cumprod(delta)
            Plam = sweep_times(Lambda_prec, 2, tauh)
            Lambda = np.column_stack([Lambda, variates["k_z_lambda"] * np.sqrt(1.0 / Plam[:, k - 1])])
            F_h2 = np.append(F_h2, variates["k_u_h2"])
            F_a = np.column_stack([F_a, variates["k_z_Fa"] * np.sqrt(F_h2[k - 1])])
            F = np.column_stack([F, Z_W @ F_a[:, k - 1] + variates["k_z_F"] * np.sqrt(1 - F_h2[k - 1])])
        elif num > 0:
            nonred = np.where(~vec)[0]
            Lambda = Lambda[:, nonred]
            Lambda_prec = Lambda_prec[:, nonred]
            F = F[:, nonred]
            for red in np.where(vec)[0]:             # 0-based; R skips red == k (1-based)
                if red == k - 1:
                    continue
                delta[red + 1] = delta[red + 1] * delta[red]
            delta = delta[nonred]
            tauh = np.cumprod(delta)
            Plam = sweep_times(Lambda_prec, 2, tauh)
            F_h2 = F_h2[nonred]
            F_a = F_a[:, nonred]
    return {"F": F, "Lambda": Lambda, "F_a": F_a, "F_h2": F_h2, "Lambda_prec": Lambda_prec,
            "Plam": Plam, "delta": delta, "tauh": tauh}


def _c_functions(fns):
    """The `_c` samplers a loop body calls: this module's restatements, or (fns = `oracle.ref_oracle`) the compiled
    reference itself -- same signatures, so the R-side loop restated below can drive either."""
    import sys
    m = sys.modules[__name__] if fns is None else fns
    return (m.sample_Lambda_c, m.sample_means_c, m.sample_h2s_discrete_c, m.sample_F_a_c, m.sample_factors_scores_c,
            m.sample_px_factors_scores_c)


# --------------------------------------------------------------------------------------
# One Gibbs iteration in the reference's formulation (fast_BSFG_sampler.R:176-270)
# --------------------------------------------------------------------------------------
def gibbs_iteration(data, rv, priors, rp, st, i, v, with_update_k=True, fns=None):
    """One pass of the loop body of `fast_BSFG_sampler` with injected variates.

    data: dict Y, X, Z_1 and optionally Z_2 (n x r2; then rv has invert_aPXA_bDesignDesignT_rand2 / A_2_inv, the
          state has W (r2 x p) / W_prec (p) and v has z_W (r2 x p) / g_W (p)); NaN in Y = missing phenotype, imputed
          every iteration from v["z_Y"] (n x p) like fast_BSFG_sampler.R:180-184
    rv  : run_variables (Ainv, invert_aI_bZAZ, invert_aPXA_bDesignDesignT, invert_aZZt_Ainv, n,p,r,b)
    st  : current_state dict (modified copy is returned)
    v   : variates dict: z_Lambda (k,p), z_means (br,p), u_h2 (k), z_Fa (r,k), z_F (n,k),
          g_Lambda_prec (p,k), g_resid (p), g_Ea (p), g_delta (>=k-1), u_k, [add-arm extras]
    Also returns `aux` with the Gamma (shape, rate) parameters of every precision draw.
    fns : None (this module's `_c` restatements) or `oracle.ref_oracle` (the compiled reference)."""
    (sample_Lambda_c, sample_means_c, sample_h2s_discrete_c, sample_F_a_c, sample_factors_scores_c,
     sample_px_factors_scores_c) = _c_functions(fns)
    Y, X, Z_1 = data["Y"], data["X"], data["Z_1"]
    n, p = Y.shape
    r = Z_1.shape[1]
    b = X.shape[1]
    Z_2 = data.get("Z_2")
    r2 = 0 if Z_2 is None else np.asarray(Z_2).shape[1]
    st = {k_: (np.array(v_, dtype=np.float64, copy=True) if isinstance(v_, np.ndarray) else v_)
          for k_, v_ in st.items()}
    B, F = st["B"], st["F"]
    if r2 > 0:
        Z_2 = np.asarray(Z_2, dtype=np.float64)
        W, W_prec = st["W"], st["W_prec"]
        ZW = Z_2 @ W
    else:
        W, W_prec = np.zeros((0, p)), st.get("W_prec", np.zeros(p))
        ZW = 0.0
    aux = {}
    # -- fill in missing phenotypes (R:180-184): conditioning on everything else.  `resids` is
    #    matrix(rnorm(p*n, 0, sqrt(1/resid_Y_prec)), nr=n, nc=p, byrow=T): element (i,j) = z[i*p + j] / sqrt(rp_j);
    #    v["z_Y"] is that N(0,1) stream already arranged as the (n, p) matrix.
    Y_missing = np.isnan(Y)
    if Y_missing.any():
        meanTraits = X @ B + F @ st["Lambda"].T + Z_1 @ st["E_a"] + ZW
        resids = v["z_Y"] * np.sqrt(1.0 / st["resid_Y_prec"])[None, :]
        Y = Y.copy()
        Y[Y_missing] = meanTraits[Y_missing] + resids[Y_missing]
        aux["Y_imputed"] = Y
    # -- Lambda (R:189-191)
    Y_tilde = Y - X @ B - ZW
    Lambda = sample_Lambda_c(Y_tilde, F, st["resid_Y_prec"], st["E_a_prec"], st["Plam"],
                             rv["invert_aI_bZAZ"], v["z_Lambda"])
    # -- B, E_a (R:203-207)
    Y_tilde = Y - F @ Lambda.T - ZW
    loc = sample_means_c(Y_tilde, st["resid_Y_prec"], st["E_a_prec"],
                         rv["invert_aPXA_bDesignDesignT"], v["z_means"])
    B = loc[:b, :]
    E_a = loc[b:b + r, :]
    # -- W (R:211-216): second call of sample_means_c, with W_prec and the rand2 list
    if r2 > 0:
        Y_tilde = Y - X @ B - Z_1 @ E_a - F @ Lambda.T
        W = sample_means_c(Y_tilde, st["resid_Y_prec"], W_prec, rv["invert_aPXA_bDesignDesignT_rand2"], v["z_W"])
        ZW = Z_2 @ W
    # -- F_h2, F_a (R:221, 226)
    F_h2 = sample_h2s_discrete_c(F, rp["h2_divisions"], priors["h2_priors_factors"],
                                 rv["invert_aI_bZAZ"], v["u_h2"])
    F_a = sample_F_a_c(F, Z_1, F_h2, rv["invert_aZZt_Ainv"], v["z_Fa"])
    # -- F (R:230-232)
    Y_tilde = Y - X @ B - Z_1 @ E_a - ZW
    F = sample_factors_scores_c(Y_tilde, Z_1, Lambda, st["resid_Y_prec"], F_a, F_h2, v["z_F"])
    # -- Lambda_prec (R:235-236)
    Lambda2 = Lambda ** 2
    df = priors["Lambda_df"]
    lp_shape = (df + 1) / 2.0
    lp_rate = (df + sweep_times(Lambda2, 2, st["tauh"])) / 2.0
    Lambda_prec = v["g_Lambda_prec"] / lp_rate
    aux["Lambda_prec_shape"], aux["Lambda_prec_rate"] = lp_shape, lp_rate
    # -- resid_Y_prec (R:239-240)
    Y_tilde = Y - X @ B - F @ Lambda.T - Z_1 @ E_a - ZW
    ry_rate = priors["resid_Y_prec_rate"] + 0.5 * np.sum(Y_tilde ** 2, axis=0)
    resid_Y_prec = v["g_resid"] / ry_rate
    aux["resid_Y_prec_shape"], aux["resid_Y_prec_rate"] = priors["resid_Y_prec_shape"] + n / 2.0, ry_rate
    # -- E_a_prec (R:245): diag(t(E_a) %*% Ainv %*% E_a)
    ea_rate = priors["E_a_prec_rate"] + 0.5 * np.einsum("ij,ij->j", E_a, rv["Ainv"] @ E_a)
    E_a_prec = v["g_Ea"] / ea_rate
    aux["E_a_prec_shape"], aux["E_a_prec_rate"] = priors["E_a_prec_shape"] + r / 2.0, ea_rate
    # -- W_prec (R:248-250): diag(t(W) %*% A_2_inv %*% W)
    if r2 > 0:
        A2 = rv.get("A_2_inv")
        A2 = np.eye(r2) if A2 is None else np.asarray(A2, dtype=np.float64)
        w_rate = priors["W_prec_rate"] + 0.5 * np.einsum("ij,ij->j", W, A2 @ W)
        W_prec = v["g_W"] / w_rate
        aux["W_prec_shape"], aux["W_prec_rate"] = priors["W_prec_shape"] + r2 / 2.0, w_rate
    # -- delta, tauh, Plam (R:253-257)
    delta, d_shapes, d_rates = sample_delta(st["delta"], st["tauh"], Lambda_prec,
                                            priors["delta_1_shape"], priors["delta_1_rate"],
                                            priors["delta_2_shape"], priors["delta_2_rate"],
                                            Lambda2, v["g_delta"])
    aux["delta_shape"], aux["delta_rate"] = d_shapes, d_rates
    tauh = np.cumprod(delta)
    Plam = sweep_times(Lambda_prec, 2, tauh)
    out = dict(st)
    out.update(Lambda=Lambda, B=B, E_a=E_a, F_h2=F_h2, F_a=F_a, F=F, Lambda_prec=Lambda_prec,
               resid_Y_prec=resid_Y_prec, E_a_prec=E_a_prec, delta=delta, tauh=tauh, Plam=Plam)
    if r2 > 0:
        out.update(W=W, W_prec=W_prec)
    # -- update_k (R:260-269)
    if with_update_k:
        res = update_k(F, Lambda, F_a, F_h2, Lambda_prec, Plam, delta, tauh, Z_1, df,
                       priors["delta_2_shape"], priors["delta_2_rate"], rp["b0"], rp["b1"], i,
                       rp["epsilon"], rp["prop"], v)
        out.update(res)
    out["nrun"] = i
    return out, aux


# --------------------------------------------------------------------------------------
# One iteration of the fixed-Lambda sampler (fast_BSFG_sampler_fixedlambda.R:159-236)
# --------------------------------------------------------------------------------------
def gibbs_iteration_fixedlambda(data, rv, priors, rp, st, Lambda, v, fns=None):
    """Lambda (p, k) is a stored posterior draw (`matrix(BSFG_state$Lambda[,j], nr=p)`, :163-164); the loop resamples
    B/E_a (after ALSO removing X B, :180), W, F_h2, F_a, F and the three precisions; Lambda, Lambda_prec, delta, k stay."""
    (sample_Lambda_c, sample_means_c, sample_h2s_discrete_c, sample_F_a_c, sample_factors_scores_c,
     sample_px_factors_scores_c) = _c_functions(fns)
    Y, X, Z_1 = data["Y"], data["X"], data["Z_1"]
    n, p = Y.shape
    r = Z_1.shape[1]
    b = X.shape[1]
    Z_2 = data.get("Z_2")
    r2 = 0 if Z_2 is None else np.asarray(Z_2).shape[1]
    st = {k_: (np.array(v_, dtype=np.float64, copy=True) if isinstance(v_, np.ndarray) else v_) for k_, v_ in st.items()}
    B, F = st["B"], st["F"]
    Lambda = np.asarray(Lambda, dtype=np.float64)
    W, W_prec = (st["W"], st["W_prec"]) if r2 > 0 else (np.zeros((0, p)), st.get("W_prec", np.zeros(p)))
    ZW = (np.asarray(Z_2, dtype=np.float64) @ W) if r2 > 0 else 0.0
    aux = {}
    Y_missing = np.isnan(Y)
    if Y_missing.any():                                                       # :170-174
        meanTraits = X @ B + F @ Lambda.T + Z_1 @ st["E_a"] + ZW
        Y = Y.copy()
        Y[Y_missing] = (meanTraits + v["z_Y"] * np.sqrt(1.0 / st["resid_Y_prec"])[None, :])[Y_missing]
        aux["Y_imputed"] = Y
    Y_tilde = Y - F @ Lambda.T - ZW - X @ B                                   # :180 (X B removed too)
    loc = sample_means_c(Y_tilde, st["resid_Y_prec"], st["E_a_prec"], rv["invert_aPXA_bDesignDesignT"], v["z_means"])
    B = loc[:b, :]
    E_a = loc[b:b + r, :]
    if r2 > 0:                                                                # :192-198
        W = sample_means_c(Y - X @ B - Z_1 @ E_a - F @ Lambda.T, st["resid_Y_prec"], W_prec,
                           rv["invert_aPXA_bDesignDesignT_rand2"], v["z_W"])
        ZW = np.asarray(Z_2, dtype=np.float64) @ W
    F_h2 = sample_h2s_discrete_c(F, rp["h2_divisions"], priors["h2_priors_factors"], rv["invert_aI_bZAZ"], v["u_h2"])
    F_a = sample_F_a_c(F, Z_1, F_h2, rv["invert_aZZt_Ainv"], v["z_Fa"])
    F = sample_factors_scores_c(Y - X @ B - Z_1 @ E_a - ZW, Z_1, Lambda, st["resid_Y_prec"], F_a, F_h2, v["z_F"])
    Y_tilde = Y - X @ B - F @ Lambda.T - Z_1 @ E_a - ZW
    ry_rate = priors["resid_Y_prec_rate"] + 0.5 * np.sum(Y_tilde ** 2, axis=0)
    ea_rate = priors["E_a_prec_rate"] + 0.5 * np.einsum("ij,ij->j", E_a, rv["Ainv"] @ E_a)
    out = dict(st)
    out.update(Lambda=Lambda, B=B, E_a=E_a, F_h2=F_h2, F_a=F_a, F=F, resid_Y_prec=v["g_resid"] / ry_rate,
               E_a_prec=v["g_Ea"] / ea_rate)
    aux["resid_Y_prec_rate"], aux["E_a_prec_rate"] = ry_rate, ea_rate
    if r2 > 0:
        A2 = rv.get("A_2_inv")
        A2 = np.eye(r2) if A2 is None else np.asarray(A2, dtype=np.float64)
        w_rate = priors["W_prec_rate"] + 0.5 * np.einsum("ij,ij->j", W, A2 @ W)
        out.update(W=W, W_prec=v["g_W"] / w_rate)
        aux["W_prec_rate"] = w_rate
    return out, aux


# --------------------------------------------------------------------------------------
# Parameter-expanded sampler (fast_BSFG_sampler_px.R:196-311, update_k of BSFG_functions_px.R:301-369)
# --------------------------------------------------------------------------------------
def update_k_px(F, Lambda, F_a, F_h2, Lambda_prec, Plam, delta, tauh, px_factor, F_px, Z_W, Lambda_df, delta_2_shape,
                delta_2_rate, px_shape, px_rate, b0, b1, i, epsilon, prop, variates):
    """As `update_k` plus px_factor / F_px; draw order of the add arm: k_g_lp, k_g_delta, k_z_lambda, k_u_h2, k_z_Fa,
    k_z_F, k_g_px (BSFG_functions_px.R:325-334).  Plam is recomputed WITHOUT the px division in both arms (:328, :348)."""
    res = update_k(F, Lambda, F_a, F_h2, Lambda_prec, Plam, delta, tauh, Z_W, Lambda_df, delta_2_shape, delta_2_rate,
                   b0, b1, i, epsilon, prop, variates)
    px_factor = np.asarray(px_factor, dtype=np.float64).ravel().copy()
    k_old, k_new = F.shape[1], res["F"].shape[1]
    if k_new == k_old + 1:
        px_factor = np.append(px_factor, variates["k_g_px"] / px_rate)
        F_px = np.column_stack([F_px, res["F"][:, k_new - 1] * np.sqrt(px_factor[k_new - 1])])
    elif k_new < k_old:
        lind = np.mean(np.abs(Lambda) < epsilon, axis=0)
        nonred = np.where(~(lind >= prop))[0]
        px_factor = px_factor[nonred]
        F_px = F_px[:, nonred]
    res.update(px_factor=px_factor, F_px=F_px)
    return res


def gibbs_iteration_px(data, rv, priors, rp, st, i, v, with_update_k=True, fns=None):
    """One pass of the loop body of fast_BSFG_sampler_px.R (:196-311).  State additionally holds px_factor (k) and F_px
    (n, k); variates additionally g_px (k) ~ Gamma(px_shape + n/2, 1) and, for the add arm, k_g_px ~ Gamma(px_shape, 1)."""
    (sample_Lambda_c, sample_means_c, sample_h2s_discrete_c, sample_F_a_c, sample_factors_scores_c,
     sample_px_factors_scores_c) = _c_functions(fns)
    Y, X, Z_1 = data["Y"], data["X"], data["Z_1"]
    n, p = Y.shape
    r = Z_1.shape[1]
    b = X.shape[1]
    st = {k_: (np.array(v_, dtype=np.float64, copy=True) if isinstance(v_, np.ndarray) else v_) for k_, v_ in st.items()}
    B, F_px, px = st["B"], st["F_px"], st["px_factor"]
    aux = {}
    Lambda_px = sample_Lambda_c(Y - X @ B, F_px, st["resid_Y_prec"], st["E_a_prec"], st["Plam"], rv["invert_aI_bZAZ"], v["z_Lambda"])
    loc = sample_means_c(Y - F_px @ Lambda_px.T, st["resid_Y_prec"], st["E_a_prec"], rv["invert_aPXA_bDesignDesignT"], v["z_means"])
    B = loc[:b, :]
    E_a = loc[b:b + r, :]
    F_temp = sweep_times(F_px, 2, 1.0 / np.sqrt(px))                                            # :232
    F_h2 = sample_h2s_discrete_c(F_temp, rp["h2_divisions"], priors["h2_priors_factors"], rv["invert_aI_bZAZ"], v["u_h2"])
    F_a = sample_F_a_c(F_temp, Z_1, F_h2, rv["invert_aZZt_Ainv"], v["z_Fa"])
    F_a_px = sweep_times(F_a, 2, np.sqrt(px))                                                    # :249
    tau_e = 1.0 / (px * (1.0 - F_h2))                                                            # :250
    F_px = sample_px_factors_scores_c(Y - X @ B - Z_1 @ E_a, Z_1, Lambda_px, st["resid_Y_prec"], F_a_px, tau_e, px, v["z_F"])
    F_px_resid = F_px - Z_1 @ F_a_px                                                             # :260
    px_rate_post = priors["px_rate"] + 0.5 * np.sum(F_px_resid ** 2, axis=0) / (1.0 - F_h2)
    px = 1.0 / (v["g_px"] / px_rate_post)                                                        # :263
    aux["px_shape"], aux["px_rate"] = priors["px_shape"] + n / 2.0, px_rate_post
    Lambda = sweep_times(Lambda_px, 2, np.sqrt(px))                                              # :267
    Lambda2 = Lambda ** 2
    df = priors["Lambda_df"]
    lp_rate = (df + sweep_times(Lambda2, 2, st["tauh"])) / 2.0
    Lambda_prec = v["g_Lambda_prec"] / lp_rate
    aux["Lambda_prec_rate"] = lp_rate
    Y_tilde = Y - X @ B - F_px @ Lambda_px.T - Z_1 @ E_a
    ry_rate = priors["resid_Y_prec_rate"] + 0.5 * np.sum(Y_tilde ** 2, axis=0)
    ea_rate = priors["E_a_prec_rate"] + 0.5 * np.einsum("ij,ij->j", E_a, rv["Ainv"] @ E_a)
    aux["resid_Y_prec_rate"], aux["E_a_prec_rate"] = ry_rate, ea_rate
    delta, d_shapes, d_rates = sample_delta(st["delta"], st["tauh"], Lambda_prec, priors["delta_1_shape"], priors["delta_1_rate"],
                                            priors["delta_2_shape"], priors["delta_2_rate"], Lambda2, v["g_delta"])
    aux["delta_rate"] = d_rates
    tauh = np.cumprod(delta)
    Plam = sweep_times(sweep_times(Lambda_prec, 2, tauh), 2, 1.0 / px)                           # :289-290
    F = sweep_times(F_px, 2, 1.0 / np.sqrt(px))                                                  # :294
    out = dict(st)
    out.update(Lambda=Lambda, B=B, E_a=E_a, F_h2=F_h2, F_a=F_a, F=F, F_px=F_px, px_factor=px, Lambda_prec=Lambda_prec,
               resid_Y_prec=v["g_resid"] / ry_rate, E_a_prec=v["g_Ea"] / ea_rate, delta=delta, tauh=tauh, Plam=Plam)
    if with_update_k:
        res = update_k_px(F, Lambda, F_a, F_h2, Lambda_prec, Plam, delta, tauh, px, F_px, Z_1, df, priors["delta_2_shape"],
                          priors["delta_2_rate"], priors["px_shape"], priors["px_rate"], rp["b0"], rp["b1"], i, rp["epsilon"],
                          rp["prop"], v)
        out.update(res)
    out["nrun"] = i
    return out, aux
