"""Second, independent transcription -- of the reference's *R twins* -- TEST INFRASTRUCTURE ONLY.

Follows `/root/reference/R_BSFG/BSFG_functions.R:23-269` (sample_Lambda, sample_prec_discrete_conditional,
sample_h2s_discrete, sample_means, sample_F_a, sample_factors_scores), NOT the C++ file.  The
reference's README (README.md:37-38) states the R functions and their `_c` versions "should be
identical (up to RNG differences)"; `tests/test_oracle.py` asserts exactly that between this module
and `oracle/bsfg_oracle.py` under shared variates.  Known, intended divergences between the twins are
kept and exercised by the tests:
  * `sample_prec_discrete_conditional`: R uses real `n/2` (BSFG_functions.R:99), C++ integer `n/2`
    (BSFG_functions_c.cpp:274) -> twins differ for odd n.
  * `sample_F_a`: R tests `tau_e == Inf` (BSFG_functions.R:229), C++ `tau_e > 10000` (cpp:326).
This transcription deliberately uses different primitives (dense inverses / vectorised dnorm) so a
shared bug is unlikely.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cho_factor, cho_solve, solve_triangular
from scipy.stats import norm


def sample_Lambda(Y_tilde, F, resid_Y_prec, E_a_prec, Plam, inv_list, z):
    # BSFG_functions.R:23-61
    U, s = inv_list["U"], inv_list["s"]
    p = len(resid_Y_prec)
    k = F.shape[1]
    FtU = F.T @ U
    UtY = U.T @ Y_tilde
    Lambda = np.zeros((p, k))
    for j in range(p):
        FUDi = E_a_prec[j] * (FtU / (s + E_a_prec[j] / resid_Y_prec[j])[None, :])   # R:46 sweep '/'
        means = FUDi @ UtY[:, j]
        Qlam = FUDi @ FtU.T + np.diag(Plam[j, :])
        Llam = np.linalg.cholesky(Qlam)                      # t(chol(Qlam)), R:51
        vlam = solve_triangular(Llam, means, lower=True)     # forwardsolve, R:52
        mlam = solve_triangular(Llam.T, vlam, lower=False)   # backsolve, R:53
        ylam = solve_triangular(Llam.T, z[:, j], lower=False)
        Lambda[j, :] = ylam + mlam
    return Lambda


def sample_prec_discrete_conditional(Y, h2_divisions, h2_priors, inv_list, res_prec, u):
    # BSFG_functions.R:63-112
    U, s = inv_list["U"], inv_list["s"]
    n, p = Y.shape
    log_ps = np.full((p, h2_divisions), np.nan)
    std_scores_b = Y.T @ U
    for i in range(1, h2_divisions + 1):
        h2 = (i - 1) / h2_divisions
        if h2 > 0:
            std_scores = std_scores_b / np.sqrt(s + (1 - h2) / h2)[None, :]
            std_scores = std_scores / np.sqrt(h2 / (res_prec * (1 - h2)))[:, None]
            det = np.sum(np.log(np.outer(s + (1 - h2) / h2, h2 / (res_prec * (1 - h2)))) / 2, axis=0)
        else:
            std_scores = Y.T / np.sqrt(1 / res_prec)[:, None]
            det = n / 2 * np.log(1 / res_prec)               # real division, R:99
        with np.errstate(divide="ignore"):
            log_ps[:, i - 1] = np.sum(norm.logpdf(std_scores), axis=1) - det + np.log(h2_priors[i - 1])
    Trait_h2 = np.zeros(p)
    for j in range(p):
        mx = np.max(log_ps[j, :])
        norm_factor = mx + np.log(np.sum(np.exp(log_ps[j, :] - mx)))
        ps_j = np.exp(log_ps[j, :] - norm_factor)
        Trait_h2[j] = np.sum(u[j] > np.cumsum(ps_j)) / h2_divisions
    with np.errstate(divide="ignore"):
        return (res_prec * (1 - Trait_h2)) / Trait_h2


def sample_h2s_discrete(F, h2_divisions, h2_priors, inv_list, u):
    # BSFG_functions.R:115-159
    U, s = inv_list["U"], inv_list["s"]
    k = F.shape[1]
    log_ps = np.zeros((k, h2_divisions))
    std_scores_b = F.T @ U
    for i in range(1, h2_divisions + 1):
        h2 = (i - 1) / h2_divisions
        if h2 > 0:
            std_scores = 1 / np.sqrt(h2) * (std_scores_b / np.sqrt(s + (1 - h2) / h2)[None, :])
            det = np.sum(np.log((s + (1 - h2) / h2) * h2) / 2)
        else:
            std_scores = F.T
            det = 0.0
        with np.errstate(divide="ignore"):
            log_ps[:, i - 1] = np.sum(norm.logpdf(std_scores), axis=1) - det + np.log(h2_priors[i - 1])
    F_h2 = np.zeros(k)
    for j in range(k):
        mx = np.max(log_ps[j, :])
        norm_factor = mx + np.log(np.sum(np.exp(log_ps[j, :] - mx)))
        ps_j = np.exp(log_ps[j, :] - norm_factor)
        F_h2[j] = np.sum(u[j] > np.cumsum(ps_j)) / h2_divisions
    return F_h2


def sample_means(Y_tilde, resid_Y_prec, E_a_prec, inv_list, z):
    # BSFG_functions.R:162-202 -- vectorised over traits instead of the j-loop
    U, s1, s2, Design_U = inv_list["U"], inv_list["s1"], inv_list["s2"], inv_list["Design_U"]
    means = (Design_U.T @ Y_tilde) * resid_Y_prec[None, :]
    d = np.outer(s1, E_a_prec) + np.outer(s2, resid_Y_prec)
    return U @ (means / d + z / np.sqrt(d))


def sample_F_a(F, Z_1, F_h2, inv_list, z):
    # BSFG_functions.R:204-242
    U, s1, s2 = inv_list["U"], inv_list["s1"], inv_list["s2"]
    k = F.shape[1]
    r = Z_1.shape[1]
    F_h2 = np.asarray(F_h2, dtype=np.float64).ravel()
    with np.errstate(divide="ignore"):
        tau_e = 1 / (1 - F_h2)
        tau_u = 1 / F_h2
    b = U.T @ (Z_1.T @ (F * tau_e[None, :]))
    F_a = np.zeros((r, k))
    for j in range(k):
        if tau_e[j] == 1:
            F_a[:, j] = 0
        elif np.isinf(tau_e[j]):                             # R:229
            F_a[:, j] = F[:, j]
        else:
            d = s2 * tau_u[j] + s1 * tau_e[j]
            F_a[:, j] = U @ (b[:, j] / d + z[:, j] / np.sqrt(d))
    return F_a


def sample_factors_scores(Y_tilde, Z_1, Lambda, resid_Y_prec, F_a, F_h2, z):
    # BSFG_functions.R:244-269: F' = S^-1 (S^-T rhs' + z'), S = chol(L' diag(rp) L + diag(tau_e)).
    # Written here through the normal equations: F' = Q^-1 rhs' + S^-1 z'.
    Lmsg = Lambda * resid_Y_prec[:, None]
    tau_e = 1 / (1 - np.asarray(F_h2, dtype=np.float64).ravel())
    Q = Lambda.T @ Lmsg + np.diag(tau_e)
    rhs = Y_tilde @ Lmsg + (Z_1 @ F_a) * tau_e[None, :]
    c = cho_factor(Q, lower=False)
    mean_part = cho_solve(c, rhs.T).T
    S = np.triu(c[0])
    noise_part = solve_triangular(S, z.T, lower=False).T
    return mean_part + noise_part
