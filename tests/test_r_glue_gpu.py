"""r/bsfg_glue.c -- compiled unmodified against the R API stand-in (tests/r_standin/) and linked to the REAL
libbsfg_b200.so -- executed on the GPU: every .Call entry point r/bsfg_b200.R uses is driven with the Sim_1 golden
inputs exactly as the R closures would (R numeric matrices, named lists, NULL / numeric `variates`, a numeric `seed`)
and must agree BIT FOR BIT with the ctypes path the other parity tests go through (same C ABI, same inputs), which in
turn is checked against the compiled reference.  Interface mirrored: the closures sourceCpp generates from
BSFG_functions_c.cpp:26,42,85,137,165,195,240,292."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "r_standin"))
import harness as H  # noqa: E402
from test_oracle import golden_problem  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def glue():
    g = H.Glue(mock=False)
    yield g
    g.reset()


def same(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return a.size == b.size and np.array_equal(a.ravel(order="F"), b.ravel(order="F"), equal_nan=True)


def _pri_vector(pri, rp, n):
    v = [pri["resid_Y_prec_shape"], pri["resid_Y_prec_rate"], pri["E_a_prec_shape"], pri["E_a_prec_rate"], pri["Lambda_df"],
         pri["delta_1_shape"], pri["delta_1_rate"], pri["delta_2_shape"], pri["delta_2_rate"],
         float(np.asarray(pri.get("b_X_prec", 1e-10)).ravel()[0]), rp["b0"], rp["b1"], rp["epsilon"], rp["prop"],
         pri.get("W_prec_shape", 2.0), pri.get("W_prec_rate", 0.1), pri.get("px_shape", 0.5), pri.get("px_rate", 0.5)]
    return np.array(v[:n], dtype=np.float64)


@pytest.mark.parametrize("seeded", [False, True])
def test_stateless_closures_bit_for_bit(glue, seeded):
    """The eight Rcpp-named closures + sample_delta_c, injected variates (seeded = False) and device Philox (True)."""
    import sparsefactormixedmodel_b200 as B
    g, d, rv, pri, rp = golden_problem("sim1")
    Y, X, Z = d["Y"], d["X"], d["Z_1"]
    st, v, nx = g["state0"], g["variates1"], g["state1"]
    b = X.shape[1]; H_ = rp["h2_divisions"]; seed = 4242
    var = (lambda name: None) if seeded else (lambda name: v[name])
    kw = lambda name: dict(seed=seed) if seeded else dict(variates=v[name])
    l1, l2, l3 = rv["invert_aI_bZAZ"], rv["invert_aPXA_bDesignDesignT"], rv["invert_aZZt_Ainv"]
    Yt = Y - X @ st["B"].reshape(b, -1)
    a = glue.call("R_bsfg_sample_Lambda_c", Yt, st["F"], st["resid_Y_prec"], st["E_a_prec"], st["Plam"], l1, var("z_Lambda"), float(seed))
    assert same(a, B.sample_Lambda_c(Yt, st["F"], st["resid_Y_prec"], st["E_a_prec"], st["Plam"], l1, **kw("z_Lambda")))
    Ym = Y - st["F"] @ nx["Lambda"].T
    a = glue.call("R_bsfg_sample_means_c", Ym, st["resid_Y_prec"], st["E_a_prec"], l2, var("z_means"), float(seed))
    assert same(a, B.sample_means_c(Ym, st["resid_Y_prec"], st["E_a_prec"], l2, **kw("z_means")))
    a = glue.call("R_bsfg_sample_h2s_discrete_c", st["F"], int(H_), pri["h2_priors_factors"], l1, var("u_h2"), float(seed))
    assert a.shape == (st["F"].shape[1], 1)                      # a k x 1 matrix like RcppArmadillo's wrap(vec)
    assert same(a, B.sample_h2s_discrete_c(st["F"], H_, pri["h2_priors_factors"], l1, **kw("u_h2")))
    a = glue.call("R_bsfg_sample_F_a_c", st["F"], Z, nx["F_h2"], l3, var("z_Fa"), float(seed))
    assert same(a, B.sample_F_a_c(st["F"], Z, nx["F_h2"], l3, **kw("z_Fa")))
    Yf = Y - X @ nx["B"].reshape(b, -1) - Z @ nx["E_a"]
    a = glue.call("R_bsfg_sample_factors_scores_c", Yf, Z, nx["Lambda"], st["resid_Y_prec"], nx["F_a"], nx["F_h2"], var("z_F"), float(seed))
    assert same(a, B.sample_factors_scores_c(Yf, Z, nx["Lambda"], st["resid_Y_prec"], nx["F_a"], nx["F_h2"], **kw("z_F")))
    tau_e, px = 1.0 / (1.0 - nx["F_h2"]), np.ones_like(nx["F_h2"])
    a = glue.call("R_bsfg_sample_px_factors_scores_c", Yf, Z, nx["Lambda"], st["resid_Y_prec"], nx["F_a"], tau_e, px, var("z_F"), float(seed))
    assert same(a, B.sample_px_factors_scores_c(Yf, Z, nx["Lambda"], st["resid_Y_prec"], nx["F_a"], tau_e, px, **kw("z_F")))
    u = None if seeded else g["prec_cond"]["u"]
    a = glue.call("R_bsfg_sample_prec_discrete_conditional_c", Y, int(H_), pri["h2_priors_factors"], l1, st["resid_Y_prec"], u, float(seed))
    assert a.shape == (Y.shape[1], 1)
    assert same(a, B.sample_prec_discrete_conditional_c(Y, H_, pri["h2_priors_factors"], l1, st["resid_Y_prec"],
                                                        **(dict(seed=seed) if seeded else dict(variates=u))))
    L2 = nx["Lambda"] ** 2
    gd = None if seeded else v["g_delta"]
    a = glue.call("R_bsfg_sample_delta_c", st["delta"], st["tauh"], nx["Lambda_prec"], pri["delta_1_shape"], pri["delta_1_rate"],
                  pri["delta_2_shape"], pri["delta_2_rate"], L2, gd, float(seed))
    assert same(a, B.sample_delta_c(st["delta"], st["tauh"], nx["Lambda_prec"], pri["delta_1_shape"], pri["delta_1_rate"],
                                    pri["delta_2_shape"], pri["delta_2_rate"], L2, **(dict(seed=seed) if seeded else dict(variates=gd))))


def test_GSVD_2_c_list(glue):
    import sparsefactormixedmodel_b200 as B
    rng = np.random.default_rng(5)
    n = 12
    A = np.linalg.cholesky(np.cov(rng.standard_normal((n, 3 * n)))).T
    Bm = np.linalg.cholesky(np.cov(rng.standard_normal((n, 3 * n)))).T
    got = glue.call("R_bsfg_GSVD_2_c", A, Bm)
    want = B.GSVD_2_c(A, Bm)
    for i, nm in enumerate(("U", "V", "X", "C", "S")):
        assert same(got[i], want[nm]), nm


def test_update_k_closure_drop_and_add(glue):
    """update_k_c as r/bsfg_b200.R calls it: members padded to k + 1 columns, edited in place, new k returned."""
    import sparsefactormixedmodel_b200 as B
    g, d, rv, pri, rp = golden_problem("sim1")
    st = dict(g["state1"]); st["nrun"] = 400
    k = st["Lambda"].shape[1]; n = st["F"].shape[0]; p = st["Lambda"].shape[0]; r = st["F_a"].shape[0]
    pad = lambda m: np.asfortranarray(np.hstack([np.asarray(m, dtype=np.float64), np.zeros((np.shape(m)[0], 1))]))
    padv = lambda x: np.concatenate([np.asarray(x, dtype=np.float64).ravel(), [0.0]])
    changed = set()
    for variant in ("drop", "add"):
        cs = dict(st)
        if variant == "drop":
            cs["Lambda"] = st["Lambda"].copy(); cs["Lambda"][:, 1] = 1e-9        # a redundant column (all |Lambda| < epsilon)
        else:
            cs["Lambda"] = st["Lambda"] + 1.0                                  # nothing below epsilon: the add arm
        for seed in range(1, 40):                                              # uu < prob decides; find a seed that acts
            want = B.update_k_c(cs, pri, rp, d, seed=seed)
            if want["k"] != k:
                break
        assert (want["k"] < k) if variant == "drop" else (want["k"] == k + 1), (variant, want["k"])   # every redundant column goes
        bufs = {nm: pad(cs[nm]) for nm in ("Lambda", "Lambda_prec", "Plam", "F", "F_a")}
        vecs = {nm: padv(cs[nm]) for nm in ("delta", "tauh", "F_h2")}
        kn = glue.call("R_bsfg_update_k_c", np.array([n, p, r, k, k + 1, cs["nrun"]], dtype=np.int32), _pri_vector(pri, rp, 18),
                       None, bufs["Lambda"], bufs["Lambda_prec"], bufs["Plam"], vecs["delta"], vecs["tauh"], vecs["F_h2"],
                       bufs["F"], bufs["F_a"], None, None, float(seed))
        kn = int(kn[0])
        assert kn == want["k"]
        for nm in bufs:
            assert same(bufs[nm][:, :kn], want[nm]), (variant, nm)
        for nm in vecs:
            assert same(vecs[nm][:kn], want[nm]), (variant, nm)
        changed.add(variant)
    assert changed == {"drop", "add"}


def test_device_sweep_through_the_glue(glue):
    """bsfg_sweep_loop's calls: R_bsfg_create -> R_bsfg_set_state -> R_bsfg_sweep -> R_bsfg_get_state (two-call protocol),
    bit for bit against BSFGSweep with the same Philox seed."""
    from sparsefactormixedmodel_b200 import BSFGSweep
    g, d, rv, pri, rp = golden_problem("sim1")
    cs = dict(g["state0"]); cs["nrun"] = 0
    n, p = d["Y"].shape
    k = cs["F"].shape[1]; r = d["Z_1"].shape[1]; b = d["X"].shape[1]
    seed, k_max = 977, k + 4
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=k_max, seed=seed)
    try:
        sw.set_state(cs)
        sw.sweep(3)
        want = sw.get_state()
    finally:
        sw.close()
    dims = np.array([n, p, k, k_max, r, b, rp["h2_divisions"], 0], dtype=np.int32)
    ident = bool(np.array_equal(d["Z_1"], np.eye(n)))
    ctx = glue.call("R_bsfg_create", dims, _pri_vector(pri, rp, 16), pri["h2_priors_factors"], d["Y"], d["X"], d["Z_1"],
                    rv["invert_aI_bZAZ"], rv["invert_aPXA_bDesignDesignT"], rv["invert_aZZt_Ainv"], rv.get("Ainv"), ident,
                    float(seed), None, None)
    rstate = {nm: np.asfortranarray(np.asarray(cs[nm], dtype=np.float64)) for nm in
              ("Lambda", "Lambda_prec", "Plam", "F", "F_a", "B", "E_a", "resid_Y_prec", "E_a_prec", "delta", "tauh")}
    rstate["F_h2"] = np.asarray(cs["F_h2"], dtype=np.float64).reshape(-1, 1)
    rstate["B"] = rstate["B"].reshape(b, p)
    rstate["W"] = np.zeros((0, p)); rstate["W_prec"] = np.zeros(p); rstate["nrun"] = 0
    for _ in range(2):
        glue.lib.rs_poison_stack()
        glue.call("R_bsfg_set_state", ctx, rstate)
    glue.call("R_bsfg_sweep", ctx, 3)
    kn = glue.call("R_bsfg_get_state", ctx, None)
    assert list(kn) == [want["F"].shape[1], 3]
    kk = int(kn[0])
    out = dict(Lambda=np.zeros((p, kk), order="F"), Lambda_prec=np.zeros((p, kk), order="F"), Plam=np.zeros((p, kk), order="F"),
               F=np.zeros((n, kk), order="F"), F_a=np.zeros((r, kk), order="F"), F_h2=np.zeros((kk, 1)), B=np.zeros((b, p), order="F"),
               E_a=np.zeros((r, p), order="F"), resid_Y_prec=np.zeros(p), E_a_prec=np.zeros(p), delta=np.zeros(kk), tauh=np.zeros(kk),
               W=np.zeros((0, p)), W_prec=np.zeros(p), nrun=0)
    glue.call("R_bsfg_get_state", ctx, out)                      # fills the preallocated members in place, like R's pull()
    for nm in ("Lambda", "Lambda_prec", "Plam", "F", "F_a", "F_h2", "B", "E_a", "resid_Y_prec", "E_a_prec", "delta", "tauh"):
        assert same(out[nm], want[nm]), nm
    glue.finalize(ctx)
