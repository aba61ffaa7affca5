"""Host-side mirror of the reference's Rcpp interface (R_BSFG/BSFG_functions_c.cpp).

Same function names, argument order and list arguments as the `// [[Rcpp::export()]]` functions, so
call sites such as `fast_BSFG_sampler.R:191` read the same.  Each function forwards to the C ABI of
libbsfg_b200 (CUDA, sm_100a); nothing is computed in Python.  Two keyword-only additions:

  variates=None  inject the N(0,1)/U(0,1)/Gamma(shape,1) draws (layout = the reference's own fill
                 order, e.g. `reshape(rnorm(k*p),k,p)`); None draws on the device (Philox, `seed`).
  seed=0         Philox key for device draws.

Vector results are returned as (k, 1) / (p, 1) column matrices, as RcppArmadillo returns `vec`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_f, check, ptr


def _vec(a, n=None):
    v = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    if n is not None and v.shape[0] != n:
        raise ValueError(f"expected a vector of length {n}, got {v.shape[0]}")
    return v


def sample_Lambda_c(Y_tilde, F, resid_Y_prec, E_a_prec, Plam, invert_aI_bZAZ, *, variates=None, seed=0):
    """BSFG_functions_c.cpp:86-135.  Returns Lambda (p, k)."""
    lib = _lib.load()
    Y = as_f(Y_tilde); F = as_f(F)
    n, p = Y.shape
    k = F.shape[1]
    if F.shape[0] != n:
        raise ValueError("F must have n rows")
    rp = _vec(resid_Y_prec, p); ep = _vec(E_a_prec, p)
    Pl = as_f(Plam, (p, k))
    U = as_f(invert_aI_bZAZ["U"], (n, n)); s = _vec(invert_aI_bZAZ["s"], n)
    z = as_f(variates, (k, p)) if variates is not None else None
    out = np.zeros((p, k), order="F")
    check(lib.bsfg_sample_Lambda_c(ptr(Y), n, p, ptr(F), k, ptr(rp), ptr(ep), ptr(Pl), ptr(U), ptr(s),
                                   ptr(z), C.c_ulonglong(seed), ptr(out)))
    return out


sample_lambda_c = sample_Lambda_c   # spelling used by BASELINE.json


def sample_means_c(Y_tilde, resid_Y_prec, E_a_prec, invert_aPXA_bDesignDesignT, *, variates=None, seed=0):
    """BSFG_functions_c.cpp:43-82.  Returns location_sample (br, p)."""
    lib = _lib.load()
    Y = as_f(Y_tilde)
    n, p = Y.shape
    L = invert_aPXA_bDesignDesignT
    D = as_f(L["Design_U"])
    br = D.shape[1]
    if D.shape[0] != n:
        raise ValueError("Design_U must have n rows")
    U = as_f(L["U"], (br, br)); s1 = _vec(L["s1"], br); s2 = _vec(L["s2"], br)
    rp = _vec(resid_Y_prec, p); ep = _vec(E_a_prec, p)
    z = as_f(variates, (br, p)) if variates is not None else None
    out = np.zeros((br, p), order="F")
    check(lib.bsfg_sample_means_c(ptr(Y), n, p, ptr(rp), ptr(ep), ptr(U), ptr(s1), ptr(s2), ptr(D), br,
                                  ptr(z), C.c_ulonglong(seed), ptr(out)))
    return out


def sample_factors_scores_c(Y_tilde, Z_1, Lambda, resid_Y_prec, F_a, F_h2, *, variates=None, seed=0):
    """BSFG_functions_c.cpp:138-163.  Returns F (n, k)."""
    lib = _lib.load()
    Y = as_f(Y_tilde); Z1 = as_f(Z_1); Lam = as_f(Lambda); Fa = as_f(F_a)
    n, p = Y.shape
    r = Z1.shape[1]
    k = Lam.shape[1]
    if Lam.shape[0] != p or Fa.shape != (r, k) or Z1.shape[0] != n:
        raise ValueError("shape mismatch")
    rp = _vec(resid_Y_prec, p); h2 = _vec(F_h2, k)
    z = as_f(variates, (n, k)) if variates is not None else None
    out = np.zeros((n, k), order="F")
    check(lib.bsfg_sample_factors_scores_c(ptr(Y), n, p, ptr(Z1), r, ptr(Lam), k, ptr(rp), ptr(Fa), ptr(h2),
                                           ptr(z), C.c_ulonglong(seed), ptr(out)))
    return out


def sample_px_factors_scores_c(Y_tilde, Z_1, Lambda_px, resid_Y_prec, F_a_px, tau_e, px_factor, *,
                               variates=None, seed=0):
    """BSFG_functions_c.cpp:166-193 (`px_factor` accepted and unused, like the reference)."""
    lib = _lib.load()
    Y = as_f(Y_tilde); Z1 = as_f(Z_1); Lam = as_f(Lambda_px); Fa = as_f(F_a_px)
    n, p = Y.shape
    r = Z1.shape[1]
    k = Lam.shape[1]
    rp = _vec(resid_Y_prec, p); te = _vec(tau_e, k); px = _vec(px_factor, k)
    z = as_f(variates, (n, k)) if variates is not None else None
    out = np.zeros((n, k), order="F")
    check(lib.bsfg_sample_px_factors_scores_c(ptr(Y), n, p, ptr(Z1), r, ptr(Lam), k, ptr(rp), ptr(Fa), ptr(te),
                                              ptr(px), ptr(z), C.c_ulonglong(seed), ptr(out)))
    return out


def sample_h2s_discrete_c(F, h2_divisions, h2_priors, invert_aI_bZAZ, *, variates=None, seed=0,
                          return_log_ps=False):
    """BSFG_functions_c.cpp:196-238.  Returns F_h2 (k, 1) [and the (k, H) log-posterior table]."""
    lib = _lib.load()
    F = as_f(F)
    n, k = F.shape
    H = int(h2_divisions)
    pri = _vec(h2_priors, H)
    U = as_f(invert_aI_bZAZ["U"], (n, n)); s = _vec(invert_aI_bZAZ["s"], n)
    u = _vec(variates, k) if variates is not None else None
    out = np.zeros(k)
    lp = np.zeros((k, H), order="F") if return_log_ps else None
    check(lib.bsfg_sample_h2s_discrete_c(ptr(F), n, k, H, ptr(pri), ptr(U), ptr(s), ptr(u),
                                         C.c_ulonglong(seed), ptr(out), ptr(lp)))
    out = out.reshape(k, 1)
    return (out, lp) if return_log_ps else out


def sample_prec_discrete_conditional_c(Y, h2_divisions, h2_priors, invert_aI_bZAZ, res_prec, *,
                                       variates=None, seed=0, return_log_ps=False):
    """BSFG_functions_c.cpp:241-289.  Returns Prec (p, 1)."""
    lib = _lib.load()
    Y = as_f(Y)
    n, p = Y.shape
    H = int(h2_divisions)
    pri = _vec(h2_priors, H)
    U = as_f(invert_aI_bZAZ["U"], (n, n)); s = _vec(invert_aI_bZAZ["s"], n)
    rp = _vec(res_prec, p)
    u = _vec(variates, p) if variates is not None else None
    out = np.zeros(p)
    lp = np.zeros((p, H), order="F") if return_log_ps else None
    check(lib.bsfg_sample_prec_discrete_conditional_c(ptr(Y), n, p, H, ptr(pri), ptr(U), ptr(s), ptr(rp), ptr(u),
                                                      C.c_ulonglong(seed), ptr(out), ptr(lp)))
    out = out.reshape(p, 1)
    return (out, lp) if return_log_ps else out


def sample_F_a_c(F, Z_1, F_h2, invert_aZZt_Ainv, *, variates=None, seed=0):
    """BSFG_functions_c.cpp:293-339.  Returns F_a (r, k)."""
    lib = _lib.load()
    F = as_f(F); Z1 = as_f(Z_1)
    n, k = F.shape
    r = Z1.shape[1]
    L = invert_aZZt_Ainv
    U = as_f(L["U"], (r, r)); s1 = _vec(L["s1"], r); s2 = _vec(L["s2"], r)
    h2 = _vec(F_h2, k)
    z = as_f(variates, (r, k)) if variates is not None else None
    out = np.zeros((r, k), order="F")
    check(lib.bsfg_sample_F_a_c(ptr(F), n, k, ptr(Z1), r, ptr(h2), ptr(U), ptr(s1), ptr(s2), ptr(z),
                                C.c_ulonglong(seed), ptr(out)))
    return out


def sample_delta_c(delta, tauh, Lambda_prec, delta_1_shape, delta_1_rate, delta_2_shape, delta_2_rate,
                   Lambda2, *, variates=None, seed=0, return_params=False):
    """Device version of the R-only `sample_delta` (BSFG_functions.R:272-308); same argument order.
    variates: standard Gamma(shape,1) draws in draw order."""
    lib = _lib.load()
    LP = as_f(Lambda_prec); L2 = as_f(Lambda2)
    p, k = LP.shape
    d = _vec(delta, k); t = _vec(tauh, k)
    g = _vec(variates) if variates is not None else None
    out = np.zeros(k)
    nd = C.c_int(0)
    shapes = np.zeros(k + 2); rates = np.zeros(k + 2)
    check(lib.bsfg_sample_delta_c(ptr(d), ptr(t), k, ptr(LP), p, C.c_double(delta_1_shape), C.c_double(delta_1_rate),
                                  C.c_double(delta_2_shape), C.c_double(delta_2_rate), ptr(L2), ptr(g),
                                  C.c_ulonglong(seed), ptr(out), C.byref(nd), ptr(shapes), ptr(rates)))
    if return_params:
        return out, shapes[:nd.value], rates[:nd.value]
    return out


def update_k_c(current_state, priors, run_parameters, data_matrices, *, variates=None, seed=0, k_max=None):
    """Device version of the R-only `update_k(current_state, priors, run_parameters, data_matrices)`
    (BSFG_functions.R:311-372); with `px_factor` and `F_px` in current_state, the variant of BSFG_functions_px.R:301-369.
    Returns a copy of current_state with the factor-indexed members grown / shrunk and `k` set; `nrun` is the iteration
    count i of R:323.  variates: dict with u_k and, for the add-a-factor arm, k_g_lp, k_g_delta, k_z_lambda, k_u_h2, k_z_Fa,
    k_z_F (and k_g_px), or None for the device Philox streams of (seed, nrun)."""
    lib = _lib.load()
    cs = dict(current_state)
    Lam = as_f(cs["Lambda"]); p, k = Lam.shape
    F = as_f(cs["F"]); n = F.shape[0]
    Fa = as_f(cs["F_a"]); r = Fa.shape[0]
    kmax = int(k_max if k_max is not None else k + 1)
    if kmax < k:
        raise ValueError("k_max < k")

    def grow(a, rows):
        out = np.zeros((rows, kmax), order="F")
        out[:, :k] = as_f(a, (rows, k))
        return out

    def growv(a):
        out = np.zeros(kmax)
        out[:k] = _vec(a, k)
        return out

    bufs = dict(Lambda=grow(Lam, p), Lambda_prec=grow(cs["Lambda_prec"], p), Plam=grow(cs["Plam"], p), F=grow(F, n), F_a=grow(Fa, r))
    vecs = dict(delta=growv(cs["delta"]), tauh=growv(cs["tauh"]), F_h2=growv(cs["F_h2"]))
    px = cs.get("px_factor") is not None
    if px:
        vecs["px_factor"] = growv(cs["px_factor"])
        bufs["F_px"] = grow(cs["F_px"], n)
    Z_1 = data_matrices["Z_1"]
    Z = None if (Z_1 is None or (Z_1.shape[0] == Z_1.shape[1] and np.array_equal(Z_1, np.eye(Z_1.shape[0])))) else as_f(Z_1, (n, r))
    pr, hold_p = _lib.make_priors(priors, run_parameters)
    v = _lib.bsfg_variates()
    hold = []
    if variates is not None:
        for name in ("u_k", "k_g_lp", "k_g_delta", "k_z_lambda", "k_u_h2", "k_z_Fa", "k_z_F", "k_g_px"):
            if variates.get(name) is not None:
                a = np.ascontiguousarray(np.asarray(variates[name], dtype=np.float64).ravel())
                hold.append(a)
                setattr(v, name, ptr(a))
    k_out = C.c_int(0)
    check(lib.bsfg_update_k_c(n, p, r, k, kmax, C.c_longlong(int(cs["nrun"])), C.byref(pr), ptr(Z), ptr(bufs["Lambda"]),
                              ptr(bufs["Lambda_prec"]), ptr(bufs["Plam"]), ptr(vecs["delta"]), ptr(vecs["tauh"]),
                              ptr(vecs["F_h2"]), ptr(bufs["F"]), ptr(bufs["F_a"]), ptr(vecs.get("px_factor")),
                              ptr(bufs.get("F_px")), C.byref(v) if variates is not None else None, C.c_ulonglong(seed),
                              C.byref(k_out)))
    kn = k_out.value
    for name, a in bufs.items():
        cs[name] = np.asfortranarray(a[:, :kn])
    for name, a in vecs.items():
        cs[name] = a[:kn].copy()
    cs["k"] = kn
    return cs


def GSVD_2_c(A, B):
    """`GSVD_2_c(A, B)` (BSFG_functions_c.cpp:27-40): list U, V, X, C, S with A = U C X', B = V S X' (C, S diagonal
    matrices like the reference returns them)."""
    lib = _lib.load()
    A = as_f(A); B = as_f(B)
    n = B.shape[1]
    if A.shape != (n, n) or B.shape != (n, n):
        raise ValueError("GSVD_2_c: A and B must be square of the same size (as every call site of the reference has them)")
    U = np.zeros((n, n), order="F"); V = np.zeros((n, n), order="F"); X = np.zeros((n, n), order="F")
    c = np.zeros(n); s = np.zeros(n)
    check(lib.bsfg_GSVD_2_c(ptr(A), ptr(B), n, ptr(U), ptr(V), ptr(X), ptr(c), ptr(s)))
    return {"U": U, "V": V, "X": X, "C": np.diag(c), "S": np.diag(s)}


def dgemm_tn(At, B, alpha=1.0, beta=0.0, C_in=None):
    """C = alpha * At' @ B + beta * C_in through the library's DMMA/TMA GEMM (test/bench hook)."""
    lib = _lib.load()
    At = as_f(At); B = as_f(B)
    K, M = At.shape
    N = B.shape[1]
    out = as_f(C_in, (M, N)).copy(order="F") if C_in is not None else np.zeros((M, N), order="F")
    check(lib.bsfg_dgemm_tn_host(M, N, K, C.c_double(alpha), ptr(At), K, ptr(B), K, C.c_double(beta), ptr(out), M))
    return out
