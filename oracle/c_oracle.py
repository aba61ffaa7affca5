"""ctypes binding of the C++ CPU oracle (oracle/c/bsfg_ref.cpp) -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

`gibbs_iteration_c` has the contract of `oracle.bsfg_oracle.gibbs_iteration` (without update_k, Z_2 and missing data):
one pass of the loop body of fast_BSFG_sampler (R_BSFG/fast_BSFG_sampler.R:176-270) in the reference's formulation,
compiled, calling SciPy's bundled OpenBLAS where RcppArmadillo would call BLAS/LAPACK.  tests/test_oracle.py checks it
against the NumPy restatement; bench.py times it as the CPU baseline (`kind: "port"`).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "c", "libbsfg_oracle_c.so")
_dp = C.POINTER(C.c_double)
_lib = None


def build():
    """make -C oracle/c (called by __graft_entry__.build())."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "c")])
    return _LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.bsfg_ref_iteration.restype = C.c_int
        _lib.bsfg_ref_blas_threads.restype = C.c_int
    return _lib


def blas_threads(n=0):
    """Set (n > 0) and return the OpenBLAS thread count the C++ oracle runs with."""
    return int(load().bsfg_ref_blas_threads(int(n)))


def _f(a, shape=None):
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        a = np.asfortranarray(a.reshape(shape, order="F"))
    return a


def _p(a):
    return a.ctypes.data_as(_dp)


def gibbs_iteration_c(data, rv, priors, rp, st, v, trait_threads=1):
    lib = load()
    Y = _f(data["Y"]); X = _f(data["X"]); Z1 = _f(data["Z_1"])
    n, p = Y.shape
    r, b = Z1.shape[1], X.shape[1]
    k = np.asarray(st["F"]).shape[1]
    H = int(rp["h2_divisions"])
    l1, l2, l3 = rv["invert_aI_bZAZ"], rv["invert_aPXA_bDesignDesignT"], rv["invert_aZZt_Ainv"]
    c = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())
    U1, s = _f(l1["U"]), c(l1["s"])
    U2, s12, s22, DU = _f(l2["U"]), c(l2["s1"]), c(l2["s2"]), _f(l2["Design_U"])
    U3, s13, s23 = _f(l3["U"]), c(l3["s1"]), c(l3["s2"])
    Ainv = _f(rv["Ainv"]); h2p = c(priors["h2_priors_factors"])
    pri = np.array([priors["resid_Y_prec_shape"], priors["resid_Y_prec_rate"], priors["E_a_prec_shape"], priors["E_a_prec_rate"],
                    priors["Lambda_df"], priors["delta_1_shape"], priors["delta_1_rate"], priors["delta_2_shape"],
                    priors["delta_2_rate"]], dtype=np.float64)
    Lam = np.zeros((p, k), order="F"); LP = np.zeros((p, k), order="F"); Plam = _f(st["Plam"]).copy(order="F")
    F = _f(st["F"]).copy(order="F"); Fa = np.zeros((r, k), order="F"); h2 = np.zeros(k)
    B = _f(np.asarray(st["B"]).reshape(b, p)).copy(order="F") if b > 0 else np.zeros((1, p), order="F")
    Ea = np.zeros((r, p), order="F")
    rpv, epv = c(st["resid_Y_prec"]).copy(), c(st["E_a_prec"]).copy()
    delta, tauh = c(st["delta"]).copy(), c(st["tauh"]).copy()
    rates = np.zeros(2 * p)
    zL, zM = _f(v["z_Lambda"], (k, p)), _f(v["z_means"], (b + r, p))
    uh, zFa, zF = c(v["u_h2"]), _f(v["z_Fa"], (r, k)), _f(v["z_F"], (n, k))
    gLP, gR, gE, gD = _f(v["g_Lambda_prec"], (p, k)), c(v["g_resid"]), c(v["g_Ea"]), c(v["g_delta"])
    rc = lib.bsfg_ref_iteration(n, p, k, r, b, H, _p(Y), _p(X), _p(Z1), _p(U1), _p(s), _p(U2), _p(s12), _p(s22), _p(DU), _p(U3),
                                _p(s13), _p(s23), _p(Ainv), _p(h2p), _p(pri), _p(Lam), _p(LP), _p(Plam), _p(F), _p(Fa), _p(h2),
                                _p(B), _p(Ea), _p(rpv), _p(epv), _p(delta), _p(tauh), _p(zL), _p(zM), _p(uh), _p(zFa), _p(zF),
                                _p(gLP), _p(gR), _p(gE), _p(gD), int(trait_threads), _p(rates))
    if rc != 0:
        raise RuntimeError("chol(): decomposition failed" if rc == -2 else f"bsfg_ref_iteration returned {rc}")
    out = dict(st)
    out.update(Lambda=Lam, Lambda_prec=LP, Plam=Plam, F=F, F_a=Fa, F_h2=h2, B=B[:b], E_a=Ea, resid_Y_prec=rpv, E_a_prec=epv,
               delta=delta, tauh=tauh)
    return out, dict(resid_Y_prec_rate=rates[:p].copy(), E_a_prec_rate=rates[p:].copy())
