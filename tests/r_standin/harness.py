"""Python side of the R stand-in (TEST INFRASTRUCTURE ONLY): builds r/bsfg_glue.c -- unmodified -- against
tests/r_standin/Rinternals.h, either linked to the real libbsfg_b200.so (GPU tests) or to the recording mock
(CPU test), and offers `.Call`-like access: call("R_bsfg_sample_Lambda_c", Y, F, ...) with NumPy arrays, dicts
(named R lists), ints, floats, bools and None (R NULL) as arguments."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BUILD = os.path.join(HERE, "_build")
GLUE = os.path.join(ROOT, "r", "bsfg_glue.c")
PKG = os.path.join(ROOT, "sparsefactormixedmodel_b200")

ENTRY_POINTS = {   # name -> number of .Call arguments (r/bsfg_b200.R)
    "R_bsfg_sample_Lambda_c": 8, "R_bsfg_sample_means_c": 6, "R_bsfg_sample_factors_scores_c": 8,
    "R_bsfg_sample_px_factors_scores_c": 9, "R_bsfg_sample_h2s_discrete_c": 6,
    "R_bsfg_sample_prec_discrete_conditional_c": 7, "R_bsfg_sample_F_a_c": 6, "R_bsfg_sample_delta_c": 10,
    "R_bsfg_GSVD_2_c": 2, "R_bsfg_update_k_c": 14, "R_bsfg_create": 14, "R_bsfg_set_state": 2, "R_bsfg_sweep": 2,
    "R_bsfg_get_state": 2,
}


def _stale(target, sources):
    return not os.path.exists(target) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in sources)


def build(mock):
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "glue_mock.so" if mock else "glue_b200.so")
    src = [GLUE, os.path.join(HERE, "rstandin.c")]
    dep = src + [os.path.join(HERE, "Rinternals.h"), os.path.join(ROOT, "include", "bsfg_b200.h")]
    if mock:
        src += [os.path.join(HERE, "mock_bsfg.c"), os.path.join(HERE, "mock_stubs.c")]
        link = ["-Wl,-Bsymbolic"]      # bind the glue's bsfg_* references to the mock even when the real library is already loaded
    else:
        link = ["-L" + PKG, "-lbsfg_b200", "-Wl,-rpath," + PKG]
    if _stale(out, dep + src):
        cmd = ["gcc", "-std=gnu11", "-O1", "-w", "-shared", "-fPIC", "-I" + HERE, "-I" + os.path.join(ROOT, "include")] + src + link + ["-o", out]
        subprocess.check_call(cmd)
    return out


class Glue:
    def __init__(self, mock=False):
        if not mock:
            from sparsefactormixedmodel_b200 import _lib
            _lib.load()                        # the product library first (global symbols), as bsfg_b200.R's dyn.load does
        self.lib = C.CDLL(build(mock), mode=C.RTLD_LOCAL if mock else C.RTLD_GLOBAL)   # the mock's bsfg_* must not leak
        L = self.lib
        P = C.c_void_p
        L.rs_nil.restype = P
        L.rs_real_view.restype = P; L.rs_real_view.argtypes = [P, C.c_ssize_t, C.c_int, C.c_int]
        L.rs_int_vector.restype = P; L.rs_int_vector.argtypes = [P, C.c_ssize_t]
        L.rs_lgl_scalar.restype = P; L.rs_lgl_scalar.argtypes = [C.c_int]
        L.rs_real_scalar.restype = P; L.rs_real_scalar.argtypes = [C.c_double]
        L.rs_list.restype = P; L.rs_list.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(P)]
        L.rs_type.restype = C.c_int; L.rs_type.argtypes = [P]
        L.rs_dims.argtypes = [P, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.rs_call.restype = P; L.rs_call.argtypes = [P, C.c_int, C.POINTER(P)]
        L.rs_last_error.restype = C.c_char_p
        L.rs_run_finalizer.argtypes = [P]
        L.REAL.restype = C.POINTER(C.c_double); L.REAL.argtypes = [P]
        L.INTEGER.restype = C.POINTER(C.c_int); L.INTEGER.argtypes = [P]
        L.VECTOR_ELT.restype = P; L.VECTOR_ELT.argtypes = [P, C.c_ssize_t]
        self._keep = []

    # ---- Python -> SEXP ----
    def sexp(self, x):
        L = self.lib
        if x is None:
            return L.rs_nil()
        if isinstance(x, Raw):
            return x.ptr
        if isinstance(x, bool):
            return L.rs_lgl_scalar(int(x))
        if isinstance(x, (int, np.integer)):
            v = np.array([x], dtype=np.int32); self._keep.append(v)
            return L.rs_int_vector(v.ctypes.data, 1)
        if isinstance(x, (float, np.floating)):
            return L.rs_real_scalar(float(x))
        if isinstance(x, dict):
            names = (C.c_char_p * len(x))(*[k.encode() for k in x])
            elems = (C.c_void_p * len(x))(*[self.sexp(v) for v in x.values()])
            self._keep += [names, elems]
            return L.rs_list(len(x), names, elems)
        a = np.asarray(x)
        if a.dtype.kind in "iu" or a.dtype == np.int32:
            v = np.ascontiguousarray(a.ravel(), dtype=np.int32); self._keep.append(v)
            return L.rs_int_vector(v.ctypes.data, v.size)
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 2:
            if not a.flags.f_contiguous:
                a = np.asfortranarray(a)
            self._keep.append(a)
            return L.rs_real_view(a.ctypes.data, a.size, a.shape[0], a.shape[1])
        a = np.ascontiguousarray(a.ravel()); self._keep.append(a)
        return L.rs_real_view(a.ctypes.data, a.size, -1, -1)

    # ---- SEXP -> Python ----
    def value(self, p):
        L = self.lib
        t = L.rs_type(p)
        if t == 0:
            return None
        ln, nr, nc = C.c_longlong(), C.c_int(), C.c_int()
        L.rs_dims(p, C.byref(ln), C.byref(nr), C.byref(nc))
        if t == 14:
            a = np.ctypeslib.as_array(L.REAL(p), shape=(ln.value,)).copy() if ln.value else np.zeros(0)
            return a.reshape((nr.value, nc.value), order="F") if nr.value >= 0 else a
        if t in (13, 10):
            return np.ctypeslib.as_array(L.INTEGER(p), shape=(ln.value,)).copy()
        if t == 19:
            return [self.value(L.VECTOR_ELT(p, i)) for i in range(ln.value)]
        if t == 22:
            return Raw(p)
        raise TypeError(f"SEXP type {t} not modelled")

    def call(self, name, *args):
        """.Call(name, ...): returns the converted result; an Rf_error() raises RuntimeError with R's message."""
        assert len(args) == ENTRY_POINTS[name], (name, len(args))
        fn = C.cast(getattr(self.lib, name), C.c_void_p)
        argv = (C.c_void_p * max(1, len(args)))(*[self.sexp(a) for a in args])
        out = self.lib.rs_call(fn, len(args), argv)
        if not out:
            raise RuntimeError(self.lib.rs_last_error().decode())
        return self.value(out)

    def finalize(self, raw):
        self.lib.rs_run_finalizer(raw.ptr)

    def reset(self):
        self.lib.rs_reset()
        self._keep.clear()


class Raw:
    """An opaque SEXP (external pointer) handed back to later calls."""
    def __init__(self, ptr):
        self.ptr = ptr
