"""r/bsfg_glue.c compiled UNMODIFIED against the R API stand-in (tests/r_standin/) and linked to a recording mock of the
library: the .Call surface exists with the argument counts r/bsfg_b200.R uses, errors travel as R conditions, and
R_bsfg_set_state hands the library a fully defined bsfg_state (px_factor / F_px were stack garbage).
The GPU counterpart (tests/test_r_glue_gpu.py) links the same glue to the real library."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "r_standin"))
import harness as H  # noqa: E402


@pytest.fixture(scope="module")
def glue():
    g = H.Glue(mock=True)
    yield g
    g.reset()


def test_glue_exports_every_entry_point_the_R_wrapper_calls(glue):
    rsrc = open(os.path.join(H.ROOT, "r", "bsfg_b200.R")).read()
    called = set(re.findall(r"\.Call\('(R_bsfg_\w+)'", rsrc))
    assert called == set(H.ENTRY_POINTS), called ^ set(H.ENTRY_POINTS)
    for name in called:
        assert hasattr(glue.lib, name), name
    # argument counts of the R side: every .Call('name', a, b, ...) has ENTRY_POINTS[name] arguments after the name
    for m in re.finditer(r"\.Call\('(R_bsfg_\w+)'((?:[^()]|\((?:[^()]|\([^()]*\))*\))*)\)", rsrc):
        name, rest = m.group(1), m.group(2)
        depth, n = 0, 0
        for ch in rest:
            depth += ch in "([" ; depth -= ch in ")]"
            n += (ch == "," and depth == 0)
        assert n == H.ENTRY_POINTS[name], (name, n)


def _state(k=3, n=5, p=4, r=5, b=1):
    z = lambda *s: np.asfortranarray(np.arange(float(np.prod(s))).reshape(s) + 1.0)
    return dict(Lambda=z(p, k), Lambda_prec=z(p, k), Plam=z(p, k), F=z(n, k), F_a=z(r, k), F_h2=z(k, 1), B=z(b, p), E_a=z(r, p),
                resid_Y_prec=z(p), E_a_prec=z(p), delta=z(k), tauh=z(k), W=np.zeros((0, p)), W_prec=z(p), nrun=11)


def test_set_state_hands_over_a_fully_defined_struct(glue):
    from sparsefactormixedmodel_b200 import _lib
    ctx = glue.call("R_bsfg_create", np.array([5, 4, 3, 4, 5, 1, 10, 0], dtype=np.int32), np.ones(16), np.full(10, 0.1), np.ones((5, 4)),
                    np.ones((5, 1)), np.eye(5), dict(U=np.eye(5), s=np.ones(5)),
                    dict(U=np.eye(6), s1=np.ones(6), s2=np.ones(6), Design_U=np.ones((5, 6))),
                    dict(U=np.eye(5), s1=np.ones(5), s2=np.ones(5)), None, True, 1.0, None, None)
    st = _state()
    glue.lib.mock_last_state.restype = C.POINTER(_lib.bsfg_state)
    for _ in range(3):
        glue.lib.rs_poison_stack()                       # whatever R_bsfg_set_state does not write is 0xAB..AB
        assert glue.call("R_bsfg_set_state", ctx, st) is None
        got = glue.lib.mock_last_state().contents
        assert got.k == 3 and got.nrun == 11
        assert not got.px_factor and not got.F_px, "px members must be NULL for a non-px chain"
        raw = bytes(C.string_at(C.addressof(got), C.sizeof(_lib.bsfg_state)))
        assert b"\xab" * 8 not in raw, "uninitialised bytes reached bsfg_set_state"
        assert got.Lambda and got.E_a and got.tauh
    assert glue.lib.mock_sizeof_state() == C.sizeof(_lib.bsfg_state)     # ctypes mirror == C header
    kn = glue.call("R_bsfg_get_state", ctx, None)
    assert list(kn) == [3, 7]
    glue.finalize(ctx)                                    # gc(): the finalizer destroys the context exactly once
    glue.finalize(ctx)


def test_errors_become_R_conditions(glue):
    with pytest.raises(RuntimeError, match="list member 'U' missing"):
        glue.call("R_bsfg_sample_Lambda_c", np.ones((3, 2)), np.ones((3, 2)), np.ones(2), np.ones(2), np.ones((2, 2)),
                  dict(s=np.ones(3)), None, 1.0)
    ctx = glue.call("R_bsfg_create", np.array([5, 4, 3, 4, 5, 1, 10, 0], dtype=np.int32), np.ones(16), np.full(10, 0.1), np.ones((5, 4)),
                    np.ones((5, 1)), np.eye(5), dict(U=np.eye(5), s=np.ones(5)),
                    dict(U=np.eye(6), s1=np.ones(6), s2=np.ones(6), Design_U=np.ones((5, 6))),
                    dict(U=np.eye(5), s1=np.ones(5), s2=np.ones(5)), None, True, 1.0, None, None)
    with pytest.raises(RuntimeError, match=r"libbsfg_b200 \(-5\)"):
        glue.call("R_bsfg_sweep", ctx, -1)                # the mock answers BSFG_EARG; chk() turns it into Rf_error
