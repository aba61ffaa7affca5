"""Generates tests/golden/*.npz -- run ONCE in the build container (it reads /root/reference, which does
not exist on the GPU box; nothing in tests/, smoke() or bench.py reads /root/reference at run time).

    python tests/golden/make_golden.py

The reference ships no golden vectors, so these fixtures are OUTPUTS OF THE REFERENCE ITSELF
RUN HERE: the reference's own input fixtures, at their FULL size (`R_BSFG/Sim_1/setup.mat` n=250 p=100,
`R_BSFG/Example_simulation/setup.mat` n=187 p=100, `.../Ayroles_et_al_Competitive_fitness/setup.mat` n=200 p=415
r=40 b=2 with Z_2 and its real NaN pattern), pushed through the init restatement (fast_BSFG_sampler_init.R:63-345,
`lists="reference"` = the GSVD route the R code takes) and two Gibbs iterations in which every `_c` sampler call is
served by `oracle/_ref/libbsfg_reference.so` -- `R_BSFG/BSFG_functions_c.cpp` compiled unmodified
(`oracle/ref_shim/Makefile`, `oracle/ref_oracle.py`) -- under seeded, stored variates injected in the reference's
fill order.  The R-only glue between those calls (`fast_BSFG_sampler.R:176-270`, `sample_delta`, `update_k`; R is
not installed) is the restatement in `oracle/bsfg_oracle.py`.  `ref.*` entries are the eight exports called one by
one (GSVD_2_c and sample_prec_discrete_conditional_c are not on the active sweep).  Each file holds every input of
the `_c` boundary (data, the three diagonalisation lists, priors literals, state, variates) and the outputs, so

  * `-m "not gpu"` tests re-run the NumPy oracle from the stored inputs and must reproduce the reference's outputs;
  * `-m gpu` tests push the same stored inputs through the C ABI and compare with the reference's outputs.
"""
import os
import sys

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import bsfg_oracle as O, init_oracle as I, ref_oracle as R  # noqa: E402

REF = "/root/reference/R_BSFG"
CASES = {
    # name: (path, individuals kept, traits kept, k)
    "sim1": (f"{REF}/Sim_1/setup.mat", 250, 100, 20),          # k_init = 20: model_setup.R:58
    "example": (f"{REF}/Example_simulation/setup.mat", 187, 100, 10),
    "ayroles": (f"{REF}/Runcie_Mukherjee_analyses/Ayroles_et_al_Competitive_fitness/setup.mat", 200, 415, 12),
}


def load_setup(name, path, n_keep, p_keep):
    m = sio.loadmat(path)
    Y = np.asarray(m["Y"], dtype=np.float64)
    n = Y.shape[0]
    Z_1 = np.asarray(m["Z_1"], dtype=np.float64)
    if Z_1.shape[0] != n:
        Z_1 = Z_1.T
    X = np.asarray(m["X"], dtype=np.float64)
    if X.shape[0] != n:
        X = X.T
    A = np.asarray(m["A"], dtype=np.float64)
    if name == "ayroles":
        keep = np.arange(p_keep)          # the real missing-data pattern (NaN) is kept: 20% of the panel
        Z_2 = np.asarray(m["Z_2"], dtype=np.float64)
        if Z_2.shape[0] != n:
            Z_2 = Z_2.T
        return dict(Y=Y[:, keep], X=X, Z_1=Z_1, Z_2=Z_2, A=A, simulation=False)
    # simulations: Z_1 = I, individuals are ordered by family -> the first n_keep are whole families
    assert np.array_equal(Z_1, np.eye(n))
    return dict(Y=Y[:n_keep, :p_keep], X=X[:n_keep], Z_1=np.eye(n_keep), A=A[:n_keep, :n_keep], simulation=True)


def flatten(prefix, d, out):
    for k, v in d.items():
        if isinstance(v, dict):
            flatten(f"{prefix}{k}.", v, out)
        elif v is None:
            continue
        else:
            out[f"{prefix}{k}"] = np.asarray(v)


def main():
    I.GSVD_2_c = R.GSVD_2_c          # the init's two GSVD_2_c calls (fast_BSFG_sampler_init.R:289, 320) -> compiled reference
    for name, (path, n_keep, p_keep, k) in CASES.items():
        setup = load_setup(name, path, n_keep, p_keep)
        pri = I.default_priors(k_init=k)
        rp = I.default_run_parameters()
        st = I.sampler_init(setup, pri, rp, seed=101, lists="reference")
        d, rv, pri, cs = st["data_matrices"], st["run_variables"], st["priors"], st["current_state"]
        n, p = d["Y"].shape
        r, b = d["Z_1"].shape[1], d["X"].shape[1]
        rng = np.random.default_rng(202)
        out = {}
        r2 = d["Z_2"].shape[1]
        skip = () if r2 > 0 else ("W", "W_prec")
        flatten("data.", {kk: d[kk] for kk in ("Y", "X", "Z_1") + (("Z_2",) if r2 > 0 else ())}, out)
        flatten("rv.", {kk: rv[kk] for kk in ("Ainv", "invert_aI_bZAZ", "invert_aPXA_bDesignDesignT", "invert_aZZt_Ainv")
                        + (("invert_aPXA_bDesignDesignT_rand2",) if r2 > 0 else ())}, out)
        flatten("state0.", {kk: vv for kk, vv in cs.items() if kk not in skip}, out)
        cur = cs
        for it in (1, 2):
            v = I.draw_variates(rng, n, p, k, r, b + r, pri, r2=r2, missing=bool(np.isnan(d["Y"]).any()))
            nxt, aux = R.gibbs_iteration(d, rv, pri, rp, cur, it, v)      # `_c` calls -> compiled reference
            flatten(f"variates{it}.", v, out)
            flatten(f"state{it}.", {kk: vv for kk, vv in nxt.items() if kk not in skip}, out)
            flatten(f"aux{it}.", {kk: vv for kk, vv in aux.items() if kk.endswith("_rate") or kk == "Y_imputed"}, out)
            cur = nxt
        # the export that is off the active sweep (fast_BSFG_sampler.R:198-199): stored separately
        u = rng.uniform(size=p)
        Yc = out.get("aux1.Y_imputed", d["Y"])
        prec = R.sample_prec_discrete_conditional_c(Yc, rp["h2_divisions"], pri["h2_priors_factors"],
                                                    rv["invert_aI_bZAZ"], cs["resid_Y_prec"], u)
        out["prec_cond.u"] = u
        out["prec_cond.Prec"] = prec
        # GSVD_2_c (cpp:27-40) on the pair the init hands it for the [B;E_a] list (fast_BSFG_sampler_init.R:287-289)
        Dm = np.column_stack([d["X"], d["Z_1"]])
        gA, gB = I.cholcov(I.block_diag2(1e-10 * np.eye(b), rv["Ainv"])), I.cholcov(Dm.T @ Dm)
        out["gsvd.A"], out["gsvd.B"] = gA, gB
        for kk, vv in R.GSVD_2_c(gA, gB).items():
            out[f"gsvd.{kk}"] = vv
        out["meta.k"] = np.int64(k)
        out["meta.generated_by"] = np.array("oracle/_ref/libbsfg_reference.so (BSFG_functions_c.cpp compiled unmodified)")
        dst = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(dst, **out)
        print(name, f"n={n} p={p} r={r} b={b} k={k}", f"{os.path.getsize(dst) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
