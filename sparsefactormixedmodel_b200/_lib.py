"""ctypes binding of libbsfg_b200.so (the C ABI declared in include/bsfg_b200.h).

The library is the product; this file only loads it and declares argument types.  There is no
Python/NumPy/torch fallback: if the shared object is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbsfg_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class BSFGError(RuntimeError):
    """Raised for every non-zero bsfg_status (the analogue of Rcpp turning Armadillo exceptions
    into R `stop()` conditions)."""

    def __init__(self, code, msg):
        super().__init__(f"libbsfg_b200 error {code}: {msg}")
        self.code = code


BSFG_OK, BSFG_ECUDA, BSFG_ENOTPD, BSFG_ESHAPE, BSFG_ENCCL, BSFG_EARG, BSFG_ESTATE = 0, -1, -2, -3, -4, -5, -6
BSFG_N_PHASES = 12


class bsfg_config(C.Structure):
    _fields_ = [("n", C.c_int), ("p_total", C.c_int), ("p_offset", C.c_int), ("p_local", C.c_int),
                ("k", C.c_int), ("k_max", C.c_int), ("r", C.c_int), ("b", C.c_int),
                ("h2_divisions", C.c_int), ("rank", C.c_int), ("world", C.c_int),
                ("seed", C.c_ulonglong), ("z1_is_identity", C.c_int), ("mode", C.c_int),
                ("gram_mode", C.c_int), ("gram_terms", C.c_int), ("gram_step", C.c_double), ("r2", C.c_int)]


class bsfg_priors(C.Structure):
    _fields_ = [("resid_Y_prec_shape", C.c_double), ("resid_Y_prec_rate", C.c_double),
                ("E_a_prec_shape", C.c_double), ("E_a_prec_rate", C.c_double),
                ("Lambda_df", C.c_double),
                ("delta_1_shape", C.c_double), ("delta_1_rate", C.c_double),
                ("delta_2_shape", C.c_double), ("delta_2_rate", C.c_double),
                ("b_X_prec", C.c_double), ("h2_priors_factors", c_double_p),
                ("b0", C.c_double), ("b1", C.c_double), ("epsilon", C.c_double), ("prop", C.c_double),
                ("W_prec_shape", C.c_double), ("W_prec_rate", C.c_double),
                ("px_shape", C.c_double), ("px_rate", C.c_double)]


class bsfg_constants(C.Structure):
    _fields_ = [(nm, c_double_p) for nm in
                ("Y", "X", "Z_1", "U1", "s", "U2", "s1_2", "s2_2", "Design_U", "U3", "s1_3", "s2_3", "Ainv",
                 "Z_2", "U4", "s1_4", "s2_4", "A_2_inv")]


class bsfg_state(C.Structure):
    _fields_ = [(nm, c_double_p) for nm in
                ("Lambda", "Lambda_prec", "Plam", "F", "F_a", "F_h2", "B", "E_a", "resid_Y_prec",
                 "E_a_prec", "delta", "tauh")] + [("k", C.c_int), ("nrun", C.c_int)] + \
               [("W", c_double_p), ("W_prec", c_double_p), ("px_factor", c_double_p), ("F_px", c_double_p),
                ("bcast_replicated", C.c_int)]


class bsfg_variates(C.Structure):
    _fields_ = [(nm, c_double_p) for nm in
                ("z_Lambda", "z_means", "u_h2", "z_Fa", "z_F", "g_Lambda_prec", "g_resid", "g_Ea",
                 "g_delta", "u_k", "k_g_lp", "k_g_delta", "k_z_lambda", "k_u_h2", "k_z_Fa", "k_z_F", "z_W", "g_W", "z_Y", "g_px", "k_g_px")]


class bsfg_posterior(C.Structure):
    _fields_ = [(nm, c_double_p) for nm in
                ("Lambda", "F", "F_a", "delta", "F_h2", "resid_Y_prec", "E_a_prec", "W_prec", "B", "E_a", "W")] + \
               [("sp_num", c_int_p), ("k", c_int_p)]


class bsfg_init_out(C.Structure):
    _fields_ = [(nm, c_double_p) for nm in
                ("Ainv", "U1", "s", "U2", "s1_2", "s2_2", "Design_U", "U3", "s1_3", "s2_3", "U4", "s1_4", "s2_4", "Design_U4")]


class bsfg_aux(C.Structure):
    _fields_ = [(nm, c_double_p) for nm in
                ("Lambda_prec_rate", "resid_Y_prec_rate", "E_a_prec_rate", "delta_rate", "h2_log_ps", "W_prec_rate", "Y_imputed", "px_rate")]


# every symbol include/bsfg_b200.h declares (tests/test_abi.py checks this list against the header)
EXPORTS = [
    "bsfg_version", "bsfg_last_error", "bsfg_device_count",
    "bsfg_sample_Lambda_c", "bsfg_sample_lambda_c", "bsfg_sample_means_c",
    "bsfg_sample_factors_scores_c", "bsfg_sample_px_factors_scores_c", "bsfg_sample_h2s_discrete_c",
    "bsfg_sample_prec_discrete_conditional_c", "bsfg_sample_F_a_c", "bsfg_sample_delta_c",
    "bsfg_update_k_c", "bsfg_GSVD_2_c",
    "bsfg_dgemm_tn_host", "bsfg_set_gram_default", "bsfg_get_gram_stats",
    "bsfg_create", "bsfg_destroy", "bsfg_set_priors", "bsfg_upload_constants", "bsfg_set_state",
    "bsfg_get_state", "bsfg_get_aux", "bsfg_sweep", "bsfg_sweep_injected", "bsfg_sweep_fixed_lambda", "bsfg_sweep_px", "bsfg_sweep_px_injected",
    "bsfg_init_lists", "bsfg_posterior_begin", "bsfg_sweep_collect", "bsfg_posterior_count", "bsfg_get_posterior", "bsfg_get_mode",
    "bsfg_get_timing", "bsfg_nccl_unique_id", "bsfg_comm_init", "bsfg_launch_count", "bsfg_k_limit",
]

_lib = None


def load():
    """Load libbsfg_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback for the BSFG sweep.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.bsfg_last_error.restype = C.c_char_p
    lib.bsfg_version.restype = C.c_int
    lib.bsfg_device_count.restype = C.c_int
    lib.bsfg_launch_count.restype = C.c_longlong
    lib.bsfg_k_limit.restype = C.c_int
    for name in EXPORTS:
        getattr(lib, name)      # AttributeError if a declared symbol is not exported
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise BSFGError(rc, load().bsfg_last_error().decode("utf-8", "replace"))


def as_f(a, shape=None):
    """Column-major float64 copy/view of `a` (what R/Armadillo hand the reference functions)."""
    arr = np.asfortranarray(np.asarray(a, dtype=np.float64))
    if shape is not None and tuple(arr.shape) != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {arr.shape}")
    return arr


def ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(c_double_p)


def make_priors(priors, run_parameters):
    """bsfg_priors from the reference's `priors` / `run_parameters` lists (model_setup.R:36-63).  Returns
    (struct, keepalive): the struct points into keepalive's memory."""
    h2p = np.ascontiguousarray(np.asarray(priors["h2_priors_factors"], dtype=np.float64).ravel())
    bx = priors.get("b_X_prec", 1e-10)
    bx = float(np.asarray(bx).ravel()[0]) if np.size(bx) else 1e-10
    pr = bsfg_priors(priors["resid_Y_prec_shape"], priors["resid_Y_prec_rate"], priors["E_a_prec_shape"],
                     priors["E_a_prec_rate"], priors["Lambda_df"], priors["delta_1_shape"],
                     priors["delta_1_rate"], priors["delta_2_shape"], priors["delta_2_rate"], bx,
                     ptr(h2p), run_parameters["b0"], run_parameters["b1"],
                     run_parameters["epsilon"], run_parameters["prop"],
                     priors.get("W_prec_shape", 2.0), priors.get("W_prec_rate", 0.1),
                     priors.get("px_shape", 0.5), priors.get("px_rate", 0.5))
    return pr, h2p
