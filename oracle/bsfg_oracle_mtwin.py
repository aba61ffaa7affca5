"""Third transcription -- of the reference's MATLAB originals -- TEST INFRASTRUCTURE ONLY.

Follows `/root/reference/BSF-G/sample_lambda.m`, `sample_means.m`, `sample_h2s_discrete.m`, `sample_F_a.m`,
`sample_factors_scores.m`, `sample_delta.m` and `update_k.m` (the code the R/Rcpp port was written from), keeping
MATLAB's own formulation: lower Cholesky factors with `\\` and `/` solves, `normpdf`, `bsxfun` scalings, the
transposed return of `sample_means`.  `tests/test_oracle.py` checks it against `oracle/bsfg_oracle.py` (the `_c`
restatement) under shared variates, and checks that the two differ exactly where the R port departs from MATLAB:

  * `sample_delta`: MATLAB resamples delta_2..delta_k (`for h = 2:k`, sample_delta.m:13); R stops at k-1
    (`2:(k-1)`, BSFG_functions.R:296) -- the device path follows R;
  * `update_k`, drop arm: MATLAB folds `delta(red+1) *= delta(red)` over `setdiff(1:k-1, nonred)` AFTER k has been
    decremented (update_k.m:32,36); R loops over `which(vec == 1)` and skips `red == k` with the old k
    (BSFG_functions.R:347-352) -- the device path follows R;
  * `sample_F_a`: MATLAB tests `tau_e == Inf` (sample_F_a.m:26) like the R twin, C++ `tau_e > 10000`;
  * `sample_means` returns location_sample' (p x n; sample_means.m:29), the R/C++ versions n x p.
This transcription itself cannot be run against MATLAB (no MATLAB / Octave here); it is checked against the `_c`
restatement, which tests/test_oracle.py pins to the compiled reference (oracle/_ref).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import solve_triangular
from scipy.stats import norm


def _mldivide_lower(L, b):      # L \ b, L lower triangular
    return solve_triangular(L, b, lower=True)


def _mldivide_upper(Lt, b):     # L' \ b
    return solve_triangular(Lt, b, lower=False)


def sample_lambda(Wtil, F, resid_W_prec, E_a_prec, Plam, inv_list, Zlams):
    # sample_lambda.m:10-28; Zlams = randn(k,p)
    U, s = inv_list["U"], np.asarray(inv_list["s"]).ravel()
    p, k = len(resid_W_prec), F.shape[1]
    FtU = F.T @ U
    UtY = U.T @ Wtil
    Lambda = np.zeros((p, k))
    for j in range(p):
        FUDi = E_a_prec[j] * (FtU * (1.0 / (s + E_a_prec[j] / resid_W_prec[j]))[None, :])     # bsxfun(@times, FtU, 1./(s' + ..))
        means = FUDi @ UtY[:, j]
        Qlam = FUDi @ FtU.T + np.diag(Plam[j, :])
        Llam = np.linalg.cholesky(Qlam)                                                   # chol(Qlam,'lower')
        vlam = _mldivide_lower(Llam, means)
        mlam = _mldivide_upper(Llam.T, vlam)
        ylam = _mldivide_upper(Llam.T, Zlams[:, j])
        Lambda[j, :] = ylam + mlam
    return Lambda


def sample_means(Y_tilde, resid_Y_prec, E_a_prec, inv_list, Zlams):
    # sample_means.m:12-29; Zlams = randn(n,p) with n = size(Design_U,2); returns location_sample' (p x n)
    U, s1, s2, Design_U = inv_list["U"], np.asarray(inv_list["s1"]).ravel(), np.asarray(inv_list["s2"]).ravel(), inv_list["Design_U"]
    p = len(resid_Y_prec)
    n = Design_U.shape[1]
    means = (Design_U.T @ Y_tilde) * np.asarray(resid_Y_prec)[None, :]
    location_sample = np.zeros((n, p))
    for j in range(p):
        d = s1 * E_a_prec[j] + s2 * resid_Y_prec[j]
        mlam = means[:, j] * (1.0 / d)
        location_sample[:, j] = U @ (mlam + Zlams[:, j] * (1.0 / np.sqrt(d)))
    return location_sample.T


def sample_h2s_discrete(F, h2_divisions, h2_priors, inv_list, u):
    # sample_h2s_discrete.m:8-33; u = one rand per factor, in loop order
    U, s = inv_list["U"], np.asarray(inv_list["s"]).ravel()
    k = F.shape[1]
    F_h2 = np.zeros(k)
    log_ps = np.zeros((k, h2_divisions))
    std_scores_b = F.T @ U
    for i in range(1, h2_divisions + 1):
        h2 = (i - 1) / h2_divisions
        std_scores = F
        det = 0.0
        if h2 > 0:
            std_scores = (1.0 / np.sqrt(h2)) * (std_scores_b * (1.0 / np.sqrt(s + (1 - h2) / h2))[None, :]).T
            det = np.sum(np.log((s + (1 - h2) / h2) * h2) / 2.0)
        with np.errstate(divide="ignore"):
            log_ps[:, i - 1] = np.sum(np.log(norm.pdf(std_scores)), axis=0) - det + np.log(h2_priors[i - 1])
    for j in range(k):
        m = np.max(log_ps[j, :])
        norm_factor = m + np.log(np.sum(np.exp(log_ps[j, :] - m)))
        ps_j = np.exp(log_ps[j, :] - norm_factor)
        F_h2[j] = np.sum(u[j] > np.cumsum(ps_j)) / h2_divisions
    return F_h2


def sample_F_a(F, Z_W, F_h2, inv_list, z):
    # sample_F_a.m:12-35; z = randn(n,k) with n = size(Z_W,2)
    U, s1, s2 = inv_list["U"], np.asarray(inv_list["s1"]).ravel(), np.asarray(inv_list["s2"]).ravel()
    k = F.shape[1]
    n = Z_W.shape[1]
    F_h2 = np.asarray(F_h2, dtype=np.float64).ravel()
    with np.errstate(divide="ignore"):
        tau_e = 1.0 / (1.0 - F_h2)
        tau_u = 1.0 / F_h2
    b = U.T @ Z_W.T @ (F * tau_e[None, :])
    F_a = np.zeros((n, k))
    for j in range(k):
        if tau_e[j] == 1:
            F_a[:, j] = 0.0
        elif np.isinf(tau_e[j]):
            F_a[:, j] = F[:, j]
        else:
            d = s2 * tau_u[j] + s1 * tau_e[j]
            mlam = b[:, j] * (1.0 / d)
            F_a[:, j] = U @ (mlam + z[:, j] * (1.0 / np.sqrt(d)))
    return F_a


def sample_factors_scores(W_tilde, Z_W, Lambda, resid_W_prec, F_a, F_h2, z):
    # sample_factors_scores.m:4-8; z = randn(size(Meta)) = n x k.   X / S' = (S \ X')', X / S = (S' \ X')'
    Lmsg = Lambda * np.asarray(resid_W_prec)[:, None]
    tau_e = 1.0 / (1.0 - np.asarray(F_h2, dtype=np.float64).ravel())
    S = np.linalg.cholesky(Lambda.T @ Lmsg + np.diag(tau_e))
    rhs = W_tilde @ Lmsg + (Z_W @ F_a) * tau_e[None, :]
    Meta = _mldivide_upper(S.T, _mldivide_lower(S, rhs.T)).T
    return Meta + _mldivide_upper(S.T, z.T).T


def sample_delta(delta, tauh, Lambda_prec, delta_1_shape, delta_1_rate, delta_2_shape, delta_2_rate, Lambda2, g):
    # sample_delta.m:6-17; g = standard Gamma(shape,1) variates in draw order: delta(1), then h = 2..k (k draws in all)
    delta = np.asarray(delta, dtype=np.float64).ravel().copy()
    tauh = np.asarray(tauh, dtype=np.float64).ravel().copy()
    k = len(tauh)
    mat = Lambda_prec * Lambda2
    n_genes = mat.shape[0]
    cs = np.sum(mat, axis=0)
    rate = delta_1_rate + 0.5 * (1.0 / delta[0]) * np.sum(tauh * cs)
    delta[0] = g[0] / rate                                         # gamrnd(shape, 1/rate)
    tauh = np.cumprod(delta)
    shapes, rates = [delta_1_shape + 0.5 * n_genes * k], [rate]
    for h in range(2, k + 1):
        shape = delta_2_shape + 0.5 * n_genes * (k - h + 1)
        rate = delta_2_rate + 0.5 * (1.0 / delta[h - 1]) * np.sum(tauh[h - 1:] * cs[h - 1:])
        delta[h - 1] = g[h - 1] / rate
        tauh = np.cumprod(delta)
        shapes.append(shape); rates.append(rate)
    return delta, tauh, np.array(shapes), np.array(rates)


def update_k(F, Lambda, F_a, F_h2, Lambda_prec, Plam, delta, tauh, Z_W, Lambda_df, delta_2_shape, delta_2_rate, b0, b1, i,
             epsilon, prop, variates):
    # update_k.m:8-48; variates as oracle.bsfg_oracle.update_k (u_k, and for the add arm k_g_lp, k_g_delta, k_z_lambda,
    # k_u_h2, k_z_Fa, k_z_F in MATLAB's draw order, which is also R's)
    n, k = F.shape
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
            Lambda_prec = np.column_stack([Lambda_prec, variates["k_g_lp"] * (2.0 / Lambda_df)])       # gamrnd(df/2, 2/df)
            delta = np.append(delta, variates["k_g_delta"] * (1.0 / delta_2_rate))
            tauh = np.cumprod(delta)
            Plam = Lambda_prec * tauh[None, :]
            Lambda = np.column_stack([Lambda, variates["k_z_lambda"] * np.sqrt(1.0 / Plam[:, k - 1])])
            F_h2 = np.append(F_h2, variates["k_u_h2"])
            F_a = np.column_stack([F_a, variates["k_z_Fa"] * np.sqrt(F_h2[k - 1])])
            F = np.column_stack([F, Z_W @ F_a[:, k - 1] + variates["k_z_F"] * np.sqrt(1 - F_h2[k - 1])])
        elif num > 0:
            nonred = np.array([h for h in range(1, k + 1) if not vec[h - 1]], dtype=int)     # 1-based, setdiff(1:k, find(vec))
            k = max(k - num, 1)
            Lambda = Lambda[:, nonred - 1]
            Lambda_prec = Lambda_prec[:, nonred - 1]
            F = F[:, nonred - 1]
            for red in [h for h in range(1, k) if h not in set(nonred.tolist())]:            # setdiff(1:k-1, nonred), NEW k
                delta[red] = delta[red] * delta[red - 1]                                      # delta(red+1) *= delta(red)
            delta = delta[nonred - 1]
            tauh = np.cumprod(delta)
            Plam = Lambda_prec * tauh[None, :]
            F_h2 = F_h2[nonred - 1]
            F_a = F_a[:, nonred - 1]
    return {"F": F, "Lambda": Lambda, "F_a": F_a, "F_h2": F_h2, "Lambda_prec": Lambda_prec, "Plam": Plam, "delta": delta,
            "tauh": tauh}
