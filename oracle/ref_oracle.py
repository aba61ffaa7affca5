"""ctypes binding of `oracle/_ref/libbsfg_reference.so` -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

That library IS the reference: `/root/reference/R_BSFG/BSFG_functions_c.cpp`, compiled unmodified from where it lies
(recipe `oracle/ref_shim/Makefile`) against the stand-in `oracle/ref_shim/RcppArmadillo.h` and SciPy's bundled
OpenBLAS/LAPACK.  Each function below has the call signature of its namesake in `oracle/bsfg_oracle.py` (the NumPy
restatement), variates injected in the reference's own fill order, so the two are interchangeable as checkers:

  * `tests/test_oracle.py` pins the NumPy oracle to this library (all eight exports, fixtures + random problems);
  * `-m gpu` parity tests compare the CUDA path with this library directly where it is present;
  * `tests/golden/ref_*.npz` are outputs of this library committed as fixtures (`tests/golden/make_ref_golden.py`),
    because `/root/reference` does not exist on the GPU box (the built `.so` does travel there, it is git-ignored only);
  * `bench.py`'s CPU legs time it (`cpu_baseline.kind = "reference"`).

`gibbs_iteration(...)` = the R loop body (`fast_BSFG_sampler.R:176-270`, restated in `bsfg_oracle.gibbs_iteration`,
R code cannot run here) with every `_c` call dispatched to the compiled reference instead of the restatement.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "ref_shim")
_LIB = os.path.join(_HERE, "_ref", "libbsfg_reference.so")
REFERENCE_SOURCE = "/root/reference/R_BSFG/BSFG_functions_c.cpp"
_dp = C.POINTER(C.c_double)
_lib = None


def can_build():
    return os.path.exists(REFERENCE_SOURCE)


def build():
    """make -C oracle/ref_shim (needs /root/reference; called by __graft_entry__.build() when it is present)."""
    subprocess.check_call(["make", "-s", "-C", _SHIM])
    return _LIB


def available():
    """True when the compiled reference can be used: already built (it travels to the GPU box) or buildable here."""
    return os.path.exists(_LIB) or can_build()


def load():
    global _lib
    if _lib is None:
        if can_build():
            build()                       # make is a no-op when up to date
        if not os.path.exists(_LIB):
            raise RuntimeError("oracle/_ref/libbsfg_reference.so is not built and /root/reference is not present")
        _lib = C.CDLL(_LIB)
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_blas_threads.restype = C.c_int
    return _lib


def blas_threads(n=0):
    return int(load().ref_blas_threads(int(n)))


def seed_fallback_rng(seed):
    load().ref_seed_fallback_rng(C.c_ulonglong(int(seed)))


def _f(a, shape=None):
    a = np.asarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape, order="F")
    return np.asfortranarray(a)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _z(z, shape):
    """Injected variate block in the reference's column-major fill order, or None (internal generator)."""
    return None if z is None else _f(z, shape)


def _check(rc):
    if rc != 0:
        raise RuntimeError(load().ref_last_error().decode())


def _L(v):
    return C.c_long(int(v))


def GSVD_2_c(A, B):
    A, B = _f(A), _f(B)
    n = B.shape[1]
    outs = [np.zeros((n, n), order="F") for _ in range(5)]
    _check(load().ref_GSVD_2_c(_p(A), _p(B), _L(n), *[_p(o) for o in outs]))
    return dict(zip(("U", "V", "X", "C", "S"), outs))


def sample_means_c(Y_tilde, resid_Y_prec, E_a_prec, inv_list, z):
    Y = _f(Y_tilde)
    n, p = Y.shape
    DU = _f(inv_list["Design_U"])
    br = DU.shape[1]
    U, s1, s2 = _f(inv_list["U"]), _f(inv_list["s1"]).ravel(), _f(inv_list["s2"]).ravel()
    rp, ep = _f(resid_Y_prec).ravel(), _f(E_a_prec).ravel()
    zz = _z(z, (br, p))
    out = np.zeros((br, p), order="F")
    _check(load().ref_sample_means_c(_p(Y), _L(n), _L(p), _p(rp), _p(ep), _p(U), _p(s1), _p(s2), _p(DU), _L(br), _p(zz), _p(out)))
    return out


def sample_Lambda_c(Y_tilde, F, resid_Y_prec, E_a_prec, Plam, inv_list, z):
    Y, F = _f(Y_tilde), _f(F)
    n, p = Y.shape
    k = F.shape[1]
    U, s = _f(inv_list["U"]), _f(inv_list["s"]).ravel()
    rp, ep, Pl = _f(resid_Y_prec).ravel(), _f(E_a_prec).ravel(), _f(Plam, (p, k))
    zz = _z(z, (k, p))
    out = np.zeros((p, k), order="F")
    _check(load().ref_sample_Lambda_c(_p(Y), _L(n), _L(p), _p(F), _L(k), _p(rp), _p(ep), _p(Pl), _p(U), _p(s), _p(zz), _p(out)))
    return out


def sample_factors_scores_c(Y_tilde, Z_1, Lambda, resid_Y_prec, F_a, F_h2, z):
    Y, Z1, Lam, Fa = _f(Y_tilde), _f(Z_1), _f(Lambda), _f(F_a)
    n, p = Y.shape
    r, k = Z1.shape[1], Lam.shape[1]
    rp, h2 = _f(resid_Y_prec).ravel(), _f(F_h2).ravel()
    zz = _z(z, (n, k))
    out = np.zeros((n, k), order="F")
    _check(load().ref_sample_factors_scores_c(_p(Y), _L(n), _L(p), _p(Z1), _L(r), _p(Lam), _L(k), _p(rp), _p(Fa), _p(h2), _p(zz), _p(out)))
    return out


def sample_px_factors_scores_c(Y_tilde, Z_1, Lambda_px, resid_Y_prec, F_a_px, tau_e, px_factor, z):
    Y, Z1, Lam, Fa = _f(Y_tilde), _f(Z_1), _f(Lambda_px), _f(F_a_px)
    n, p = Y.shape
    r, k = Z1.shape[1], Lam.shape[1]
    rp, te, px = _f(resid_Y_prec).ravel(), _f(tau_e).ravel(), _f(px_factor).ravel()
    zz = _z(z, (n, k))
    out = np.zeros((n, k), order="F")
    _check(load().ref_sample_px_factors_scores_c(_p(Y), _L(n), _L(p), _p(Z1), _L(r), _p(Lam), _L(k), _p(rp), _p(Fa), _p(te), _p(px),
                                                 _p(zz), _p(out)))
    return out


def sample_h2s_discrete_c(F, h2_divisions, h2_priors, inv_list, u):
    F = _f(F)
    n, k = F.shape
    U, s = _f(inv_list["U"]), _f(inv_list["s"]).ravel()
    pri = _f(h2_priors).ravel()
    uu = None if u is None else _f(u).ravel()
    out = np.zeros(k)
    _check(load().ref_sample_h2s_discrete_c(_p(F), _L(n), _L(k), C.c_int(int(h2_divisions)), _p(pri), _p(U), _p(s), _p(uu), _p(out)))
    return out


def sample_prec_discrete_conditional_c(Y, h2_divisions, h2_priors, inv_list, res_prec, u):
    Y = _f(Y)
    n, p = Y.shape
    U, s = _f(inv_list["U"]), _f(inv_list["s"]).ravel()
    pri, rp = _f(h2_priors).ravel(), _f(res_prec).ravel()
    uu = None if u is None else _f(u).ravel()
    out = np.zeros(p)
    _check(load().ref_sample_prec_discrete_conditional_c(_p(Y), _L(n), _L(p), C.c_int(int(h2_divisions)), _p(pri), _p(U), _p(s), _p(rp),
                                                         _p(uu), _p(out)))
    return out


def sample_F_a_c(F, Z_1, F_h2, inv_list, z):
    F, Z1 = _f(F), _f(Z_1)
    n, k = F.shape
    r = Z1.shape[1]
    U, s1, s2 = _f(inv_list["U"]), _f(inv_list["s1"]).ravel(), _f(inv_list["s2"]).ravel()
    h2 = _f(F_h2).ravel()
    zz = _z(z, (r, k))
    out = np.zeros((r, k), order="F")
    _check(load().ref_sample_F_a_c(_p(F), _L(n), _L(k), _p(Z1), _L(r), _p(h2), _p(U), _p(s1), _p(s2), _p(zz), _p(out)))
    return out


C_FUNCTIONS = ("sample_means_c", "sample_Lambda_c", "sample_factors_scores_c", "sample_px_factors_scores_c",
               "sample_h2s_discrete_c", "sample_prec_discrete_conditional_c", "sample_F_a_c", "GSVD_2_c")


def gibbs_iteration(data, rv, priors, rp, st, i, v, with_update_k=True):
    """`bsfg_oracle.gibbs_iteration` (the R loop body) with the `_c` samplers served by the compiled reference."""
    import sys
    from . import bsfg_oracle as O
    return O.gibbs_iteration(data, rv, priors, rp, st, i, v, with_update_k=with_update_k, fns=sys.modules[__name__])


def gibbs_iteration_px(data, rv, priors, rp, st, i, v, with_update_k=True):
    import sys
    from . import bsfg_oracle as O
    return O.gibbs_iteration_px(data, rv, priors, rp, st, i, v, with_update_k=with_update_k, fns=sys.modules[__name__])
