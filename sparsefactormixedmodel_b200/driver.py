"""Host-side mirror of the reference's R driver contract.

`fast_BSFG_sampler(BSFG_state, n_samples)` keeps the contract of `R_BSFG/fast_BSFG_sampler.R:1`:
it takes the `BSFG_state` list (here a dict of dicts with the same field names, see
fast_BSFG_sampler_init.R:352-362), runs `n_samples` Gibbs iterations starting at
`current_state$nrun + 1`, records posterior samples on the reference's burn/thin schedule
(`fast_BSFG_sampler.R:272-279` -> `save_posterior_samples`, BSFG_functions.R:375-408) and returns the
updated `BSFG_state`.  The loop body (`:176-270`) runs on the GPU through `BSFGSweep`, a thin wrapper
of the stateful C ABI (include/bsfg_b200.h); the state stays resident in HBM between iterations and is
only read back at thinning boundaries.  Plotting / .RData side effects of the R driver are not part
of the contract reproduced here.

Trait sharding (one process per GPU): every rank constructs `BSFGSweep(..., rank=, world=)` over its
slice of trait columns (`shard_bounds`), and the ranks share a NCCL communicator created from a
unique id that the caller broadcasts (bench.py uses torch.distributed for that plumbing).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_f, check, ptr

_STATE_FIELDS = ("Lambda", "Lambda_prec", "Plam", "F", "F_a", "F_h2", "B", "E_a", "resid_Y_prec",
                 "E_a_prec", "delta", "tauh")
_VARIATE_FIELDS = ("z_Lambda", "z_means", "u_h2", "z_Fa", "z_F", "g_Lambda_prec", "g_resid", "g_Ea",
                   "g_delta", "u_k", "k_g_lp", "k_g_delta", "k_z_lambda", "k_u_h2", "k_z_Fa", "k_z_F", "z_W", "g_W", "z_Y", "g_px", "k_g_px")


_GRAM_MODES = {"default": 0, "exact": 1, "expsum": 2}


def set_gram_default(gram="expsum", terms=0, step=0.0):
    """Process-wide default for the weighted-Gram evaluation of the Lambda draw (include/bsfg_b200.h
    `bsfg_set_gram_default`): "exact" = DMMA GEMM over the n individuals, "expsum" = exponential-sum factorisation
    (library default).  Applies to `sample_Lambda_c` and to sweeps created with gram="default"."""
    check(_lib.load().bsfg_set_gram_default(_GRAM_MODES[gram], int(terms), C.c_double(step)))


def init_lists(Z_1, A, X, Z_2=None, b_X_prec=1e-10):
    """The precomputed members of `run_variables` (fast_BSFG_sampler_init.R:263-325) on the GPU: returns a dict with
    Ainv, invert_aI_bZAZ, invert_aPXA_bDesignDesignT, invert_aZZt_Ainv (and invert_aPXA_bDesignDesignT_rand2 / A_2_inv when
    Z_2 has columns) that can be passed as `run_variables` to BSFGSweep."""
    lib = _lib.load()
    A = as_f(A); X = as_f(X)
    if Z_1 is None:                          # identity incidence matrix (one record per individual): r == n
        n = r = A.shape[0]
        ident = True
    else:
        Z_1 = np.asarray(Z_1, dtype=np.float64)
        n, r = Z_1.shape
        ident = _is_identity(Z_1)
    b = X.shape[1]
    Z_2 = np.zeros((n, 0)) if Z_2 is None else np.asarray(Z_2, dtype=np.float64)
    r2 = Z_2.shape[1]
    br = b + r
    f = lambda *shape: np.zeros(shape, order="F")
    o = dict(Ainv=f(r, r), U1=f(n, n), s=f(n), U2=f(br, br), s1_2=f(br), s2_2=f(br), Design_U=f(n, br), U3=f(r, r), s1_3=f(r),
             s2_3=f(r), U4=f(r2, r2), s1_4=f(r2), s2_4=f(r2), Design_U4=f(n, r2))
    out = _lib.bsfg_init_out(*[ptr(o[nm]) if o[nm].size else None for nm in
                               ("Ainv", "U1", "s", "U2", "s1_2", "s2_2", "Design_U", "U3", "s1_3", "s2_3", "U4", "s1_4", "s2_4",
                                "Design_U4")])
    Z1f = None if ident else as_f(Z_1); Z2f = as_f(Z_2) if r2 > 0 else None
    check(lib.bsfg_init_lists(n, r, b, r2, ptr(Z1f), ptr(A), ptr(X) if b > 0 else None, ptr(Z2f), C.c_double(b_X_prec), C.byref(out)))
    rv = dict(n=n, r=r, b=b, r2=r2, Ainv=o["Ainv"], invert_aI_bZAZ=dict(U=o["U1"], s=o["s"]),
              invert_aPXA_bDesignDesignT=dict(U=o["U2"], s1=o["s1_2"], s2=o["s2_2"], Design_U=o["Design_U"]),
              invert_aZZt_Ainv=dict(U=o["U3"], s1=o["s1_3"], s2=o["s2_3"]), A_2_inv=np.eye(r2),
              invert_aPXA_bDesignDesignT_rand2=(dict(U=o["U4"], s1=o["s1_4"], s2=o["s2_4"], Design_U=o["Design_U4"]) if r2 > 0 else {}))
    return rv


def shard_bounds(p, rank, world):
    """Contiguous, balanced split of the p trait columns: rank g owns [lo, hi)."""
    base, rem = divmod(p, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def _is_identity(Z):
    Z = np.asarray(Z)
    return Z.shape[0] == Z.shape[1] and np.array_equal(Z, np.eye(Z.shape[0]))


class BSFGSweep:
    """Device-resident Gibbs sweep for one chain (one rank's trait shard)."""

    def __init__(self, data_matrices, run_variables, priors, run_parameters, *, k_max=None, k_init=None,
                 seed=0, rank=0, world=1, mode="auto", nccl_id=None, gram="default", gram_terms=0, gram_step=0.0):
        self.lib = _lib.load()
        Y = np.asarray(data_matrices["Y"], dtype=np.float64)
        self.has_missing = bool(np.isnan(Y).any())      # re-imputed on the device every iteration (fast_BSFG_sampler.R:180-184)
        X = np.asarray(data_matrices["X"], dtype=np.float64)
        Z_1 = np.asarray(data_matrices["Z_1"], dtype=np.float64)
        Z_2 = data_matrices.get("Z_2")
        Z_2 = np.zeros((Y.shape[0], 0)) if Z_2 is None else np.asarray(Z_2, dtype=np.float64)
        self.r2 = r2 = Z_2.shape[1]
        n, p = Y.shape
        r = Z_1.shape[1]
        b = X.shape[1]
        self.n, self.p_total, self.r, self.b, self.br = n, p, r, b, b + r
        self.rank, self.world = rank, world
        self.lo, self.hi = shard_bounds(p, rank, world)
        self.pl = self.hi - self.lo
        self.H = int(run_parameters["h2_divisions"])
        k0 = int(k_init if k_init is not None else priors["k_init"])
        self.k_max = int(k_max if k_max is not None else max(2 * k0, k0 + 8))
        ident = _is_identity(Z_1)
        cfg = _lib.bsfg_config(n=n, p_total=p, p_offset=self.lo, p_local=self.pl, k=k0, k_max=self.k_max,
                               r=r, b=b, h2_divisions=self.H, rank=rank, world=world, seed=seed,
                               z1_is_identity=int(ident), mode={"auto": 0, "direct": 1, "diag": 2}[mode],
                               gram_mode=_GRAM_MODES[gram], gram_terms=int(gram_terms), gram_step=float(gram_step), r2=r2)
        self._ctx = C.c_void_p()
        check(self.lib.bsfg_create(C.byref(cfg), C.byref(self._ctx)))
        # priors / run parameters
        pr, self._h2p = _lib.make_priors(priors, run_parameters)
        if self._h2p.shape[0] != self.H:
            raise ValueError("h2_priors_factors must have h2_divisions entries")
        check(self.lib.bsfg_set_priors(self._ctx, C.byref(pr)))
        # constants (this rank's trait columns of Y)
        l1 = run_variables["invert_aI_bZAZ"]; l2 = run_variables["invert_aPXA_bDesignDesignT"]
        l3 = run_variables["invert_aZZt_Ainv"]
        keep = [as_f(Y[:, self.lo:self.hi]), as_f(X), None if ident else as_f(Z_1),
                as_f(l1["U"], (n, n)), np.ascontiguousarray(np.asarray(l1["s"], dtype=np.float64).ravel()),
                as_f(l2["U"], (self.br, self.br)), np.ascontiguousarray(np.asarray(l2["s1"], dtype=np.float64).ravel()),
                np.ascontiguousarray(np.asarray(l2["s2"], dtype=np.float64).ravel()), as_f(l2["Design_U"], (n, self.br)),
                as_f(l3["U"], (r, r)), np.ascontiguousarray(np.asarray(l3["s1"], dtype=np.float64).ravel()),
                np.ascontiguousarray(np.asarray(l3["s2"], dtype=np.float64).ravel()),
                as_f(run_variables["Ainv"], (r, r)) if run_variables.get("Ainv") is not None else None]
        if r2 > 0:      # second random effect: Z_2 and invert_aPXA_bDesignDesignT_rand2 (fast_BSFG_sampler_init.R:300-311)
            l4 = run_variables["invert_aPXA_bDesignDesignT_rand2"]
            A2 = run_variables.get("A_2_inv")
            keep += [as_f(Z_2, (n, r2)), as_f(l4["U"], (r2, r2)),
                     np.ascontiguousarray(np.asarray(l4["s1"], dtype=np.float64).ravel()),
                     np.ascontiguousarray(np.asarray(l4["s2"], dtype=np.float64).ravel()),
                     None if A2 is None or np.array_equal(np.asarray(A2), np.eye(r2)) else as_f(A2, (r2, r2))]
        else:
            keep += [None] * 5
        consts = _lib.bsfg_constants(*[ptr(a) for a in keep])
        check(self.lib.bsfg_upload_constants(self._ctx, C.byref(consts)))
        if world > 1:
            if nccl_id is None:
                raise ValueError("world > 1 needs the 128-byte NCCL unique id (see nccl_unique_id())")
            buf = C.create_string_buffer(bytes(nccl_id), 128)
            check(self.lib.bsfg_comm_init(self._ctx, buf))

    # -- multi-GPU plumbing ----------------------------------------------------------------
    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        check(_lib.load().bsfg_nccl_unique_id(buf))
        return buf.raw

    # -- state -----------------------------------------------------------------------------
    def set_state(self, current_state, local=False, bcast_replicated=False):
        """Upload `current_state` (fast_BSFG_sampler_init.R:238-254 field names); trait-indexed members
        are given for ALL traits and sliced to this rank's shard here, or -- `local=True` -- already hold this
        rank's shard (what get_state returns), which avoids a host copy per call.
        `bcast_replicated=True` (world > 1, a collective call): the replicated members (F / F_px, F_a, F_h2, delta, tauh,
        px_factor) are uploaded by rank 0 only and reach the other GPUs over NCCL; on ranks > 0 they may be missing from
        `current_state` (then `k` must be given as current_state["k"])."""
        cs = current_state
        k = int(np.asarray(cs["F"]).shape[1]) if cs.get("F") is not None else int(cs["k"])
        giver = not (bcast_replicated and self.world > 1) or self.rank == 0
        sl = slice(None) if local else slice(self.lo, self.hi)
        arrs = {}

        def put(name, a):
            arrs[name] = a
            return ptr(a)

        st = _lib.bsfg_state()
        st.Lambda = put("Lambda", as_f(np.asarray(cs["Lambda"])[sl, :])) if cs.get("Lambda") is not None else None
        st.Lambda_prec = put("Lambda_prec", as_f(np.asarray(cs["Lambda_prec"])[sl, :])) if cs.get("Lambda_prec") is not None else None
        st.Plam = put("Plam", as_f(np.asarray(cs["Plam"])[sl, :]))
        st.F = put("F", as_f(cs["F"])) if giver else None
        st.F_a = put("F_a", as_f(cs["F_a"])) if giver and cs.get("F_a") is not None else None
        st.F_h2 = put("F_h2", np.ascontiguousarray(np.asarray(cs["F_h2"], dtype=np.float64).ravel())) if giver and cs.get("F_h2") is not None else None
        st.B = put("B", as_f(np.asarray(cs["B"]).reshape(self.b, -1)[:, sl])) if self.b > 0 else None
        st.E_a = put("E_a", as_f(np.asarray(cs["E_a"])[:, sl])) if cs.get("E_a") is not None else None
        st.resid_Y_prec = put("rp", np.ascontiguousarray(np.asarray(cs["resid_Y_prec"], dtype=np.float64).ravel()[sl]))
        st.E_a_prec = put("ep", np.ascontiguousarray(np.asarray(cs["E_a_prec"], dtype=np.float64).ravel()[sl]))
        if giver:
            st.delta = put("delta", np.ascontiguousarray(np.asarray(cs["delta"], dtype=np.float64).ravel()))
            st.tauh = put("tauh", np.ascontiguousarray(np.asarray(cs["tauh"], dtype=np.float64).ravel()))
        st.k = k
        st.nrun = int(cs.get("nrun", 0))
        st.bcast_replicated = 1 if (bcast_replicated and self.world > 1) else 0
        self.px_mode = cs.get("px_factor") is not None and cs.get("F_px") is not None
        if self.px_mode and giver:      # parameter-expanded state (fast_BSFG_sampler_init_px.R:221-222)
            st.px_factor = put("px_factor", np.ascontiguousarray(np.asarray(cs["px_factor"], dtype=np.float64).ravel()))
            st.F_px = put("F_px", as_f(cs["F_px"]))
        if self.r2 > 0:
            st.W = put("W", as_f(np.asarray(cs["W"])[:, sl]))
            st.W_prec = put("W_prec", np.ascontiguousarray(np.asarray(cs["W_prec"], dtype=np.float64).ravel()[sl]))
        check(self.lib.bsfg_set_state(self._ctx, C.byref(st)))

    def get_state(self, with_E_a=True, out=None, with_replicated=True):
        """Download this rank's shard of `current_state` (same field names; trait-indexed members cover
        columns [lo, hi) only).  `out`: dict of preallocated F-ordered arrays from `alloc_state` to fill in place
        (e.g. pinned host memory); valid while k is unchanged.  `with_replicated=False` leaves F, F_px, F_a, F_h2, delta,
        tauh and px_factor where they are (bit-identical on every rank: one rank reading them back is enough)."""
        probe = _lib.bsfg_state()
        check(self.lib.bsfg_get_state(self._ctx, C.byref(probe)))     # all-NULL call: just k and nrun
        k = probe.k
        if out is None or out["F"].shape[1] != k:
            out = self.alloc_state(k, with_E_a)
        st = _lib.bsfg_state()
        if getattr(self, "px_mode", False) and "px_factor" not in out:
            out["px_factor"] = np.zeros(k); out["F_px"] = np.zeros((self.n, k), order="F")
        for name in _STATE_FIELDS + ("W", "W_prec", "px_factor", "F_px"):
            if not with_replicated and name in ("F", "F_px", "F_a", "F_h2", "delta", "tauh", "px_factor"):
                continue
            if name in out and out[name] is not None and out[name].size:
                setattr(st, name, ptr(out[name]))
        check(self.lib.bsfg_get_state(self._ctx, C.byref(st)))
        out["nrun"] = st.nrun
        return out

    def k(self):
        """Current number of factors (update_k changes it at run time)."""
        probe = _lib.bsfg_state()
        check(self.lib.bsfg_get_state(self._ctx, C.byref(probe)))     # all-NULL call: just k and nrun
        return int(probe.k)

    def alloc_state(self, k, with_E_a=True, empty=None):
        """Host buffers for get_state; `empty(shape)` may return pinned F-ordered float64 arrays."""
        mk = empty or (lambda shape: np.zeros(shape, order="F"))
        out = dict(Lambda=mk((self.pl, k)), Lambda_prec=mk((self.pl, k)), Plam=mk((self.pl, k)), F=mk((self.n, k)),
                   F_a=mk((self.r, k)), F_h2=mk((k,)), B=mk((self.b, self.pl)), resid_Y_prec=mk((self.pl,)),
                   E_a_prec=mk((self.pl,)), delta=mk((k,)), tauh=mk((k,)))
        if with_E_a:
            out["E_a"] = mk((self.r, self.pl))
        out["W"] = mk((self.r2, self.pl)); out["W_prec"] = mk((self.pl,))
        return out

    # -- sweeping --------------------------------------------------------------------------
    def sweep(self, n_iter):
        """`n_iter` Gibbs iterations with device Philox draws (the parameter-expanded loop when the state was set with
        px_factor / F_px, fast_BSFG_sampler_px.R)."""
        fn = self.lib.bsfg_sweep_px if getattr(self, "px_mode", False) else self.lib.bsfg_sweep
        check(fn(self._ctx, int(n_iter)))

    def sweep_injected(self, variates):
        """One iteration with injected variates (dict in the layout of oracle gibbs_iteration; trait-indexed
        blocks are given for ALL traits and sliced here)."""
        sl = slice(self.lo, self.hi)
        hold = []
        v = _lib.bsfg_variates()

        def conv(name, a):
            a = np.asarray(a, dtype=np.float64)
            if name in ("z_Lambda", "z_means", "z_W", "z_Y"):
                a = a[:, sl]
            elif name in ("g_Lambda_prec",):
                a = a[sl, :]
            elif name in ("g_resid", "g_Ea", "g_W", "k_g_lp", "k_z_lambda"):
                a = a.ravel()[sl]
            a = np.asfortranarray(a) if a.ndim == 2 else np.ascontiguousarray(a.ravel())
            hold.append(a)
            return ptr(a)

        for name in _VARIATE_FIELDS:
            if variates.get(name) is not None:
                setattr(v, name, conv(name, variates[name]))
        fn = self.lib.bsfg_sweep_px_injected if getattr(self, "px_mode", False) else self.lib.bsfg_sweep_injected
        check(fn(self._ctx, C.byref(v)))

    def sweep_fixed_lambda(self, Lambda, variates=None):
        """One iteration of the fixed-Lambda sampler (fast_BSFG_sampler_fixedlambda.R:159-236) with `Lambda` (p x k, all
        traits; sliced to this rank's shard here) a stored posterior draw; `variates` as in sweep_injected, or None."""
        L = as_f(np.asarray(Lambda, dtype=np.float64)[self.lo:self.hi, :])
        hold, vptr = [], None
        if variates is not None:
            v = _lib.bsfg_variates()
            sl = slice(self.lo, self.hi)
            for name in ("z_means", "u_h2", "z_Fa", "z_F", "g_resid", "g_Ea", "z_W", "g_W", "z_Y"):
                a = variates.get(name)
                if a is None:
                    continue
                a = np.asarray(a, dtype=np.float64)
                if name in ("z_means", "z_W", "z_Y"):
                    a = a[:, sl]
                elif name in ("g_resid", "g_Ea", "g_W"):
                    a = a.ravel()[sl]
                a = np.asfortranarray(a) if a.ndim == 2 else np.ascontiguousarray(a.ravel())
                hold.append(a)
                setattr(v, name, ptr(a))
            vptr = C.byref(v)
        check(self.lib.bsfg_sweep_fixed_lambda(self._ctx, ptr(L), int(L.shape[1]), vptr))

    # -- device-side posterior bookkeeping (save_posterior_samples, BSFG_functions.R:375-408) -------------
    def posterior_begin(self, capacity):
        check(self.lib.bsfg_posterior_begin(self._ctx, int(capacity)))

    def collect(self, n_iter, burn, thin):
        """n_iter iterations; samples on the reference's burn/thin schedule stay on the device."""
        check(self.lib.bsfg_sweep_collect(self._ctx, int(n_iter), int(burn), int(thin)))

    def posterior(self):
        """Collected samples in the reference's Posterior layout (this rank's traits): traces are (rows*k_max) x S with
        each column the column-major vec of the sample, zero padded to k_max factor columns; B, E_a, W running means."""
        S = C.c_int(0); km = C.c_int(0)
        check(self.lib.bsfg_posterior_count(self._ctx, C.byref(S), C.byref(km)))
        S, km = S.value, km.value
        out = dict(Lambda=np.zeros((self.pl * km, S), order="F"), F=np.zeros((self.n * km, S), order="F"),
                   F_a=np.zeros((self.r * km, S), order="F"), delta=np.zeros((km, S), order="F"), F_h2=np.zeros((km, S), order="F"),
                   resid_Y_prec=np.zeros((self.pl, S), order="F"), E_a_prec=np.zeros((self.pl, S), order="F"),
                   W_prec=np.zeros((self.pl, S), order="F"), B=np.zeros((self.b, self.pl), order="F"),
                   E_a=np.zeros((self.r, self.pl), order="F"), W=np.zeros((self.r2, self.pl), order="F"))
        sp = np.zeros(max(S, 1), dtype=np.int32)
        ks = np.zeros(max(S, 1), dtype=np.int32)
        po = _lib.bsfg_posterior()
        for name in ("Lambda", "F", "F_a", "delta", "F_h2", "resid_Y_prec", "E_a_prec", "W_prec", "B", "E_a", "W"):
            if out[name].size:
                setattr(po, name, ptr(out[name]))
        po.sp_num = sp.ctypes.data_as(_lib.c_int_p)
        po.k = ks.ctypes.data_as(_lib.c_int_p)
        check(self.lib.bsfg_get_posterior(self._ctx, C.byref(po)))
        out["sp_num"] = sp[:S].copy()
        out["k"] = ks[:S].copy()
        out["k_max"] = km
        return out

    def aux(self):
        probe = _lib.bsfg_state()
        check(self.lib.bsfg_get_state(self._ctx, C.byref(probe)))
        k = probe.k
        out = dict(Lambda_prec_rate=np.zeros((self.pl, k), order="F"), resid_Y_prec_rate=np.zeros(self.pl),
                   E_a_prec_rate=np.zeros(self.pl), delta_rate=np.zeros(k + 2), h2_log_ps=np.zeros((k, self.H), order="F"),
                   W_prec_rate=np.zeros(self.pl),
                   Y_imputed=np.zeros((self.n, self.pl) if self.has_missing else (0, 0), order="F"),
                   px_rate=np.zeros(k))
        a = _lib.bsfg_aux(*[ptr(out[nm]) if out[nm].size else None
                            for nm in ("Lambda_prec_rate", "resid_Y_prec_rate", "E_a_prec_rate", "delta_rate", "h2_log_ps",
                                       "W_prec_rate", "Y_imputed", "px_rate")])
        check(self.lib.bsfg_get_aux(self._ctx, C.byref(a)))
        return out

    def mode(self):
        m = C.c_int(0); r1 = C.c_double(0); r2 = C.c_double(0)
        check(self.lib.bsfg_get_mode(self._ctx, C.byref(m), C.byref(r1), C.byref(r2)))
        return {1: "direct", 2: "diag"}.get(m.value, "?"), r1.value, r2.value

    def gram_stats(self):
        m = C.c_int(0); t = C.c_int(0); h = C.c_double(0); fb = C.c_longlong(0); lo = C.c_double(0); hi = C.c_double(0)
        tn = C.c_double(0); tt = C.c_double(0)
        check(self.lib.bsfg_get_gram_stats(self._ctx, C.byref(m), C.byref(t), C.byref(h), C.byref(fb), C.byref(lo), C.byref(hi),
                                           C.byref(tn), C.byref(tt)))
        return dict(mode={1: "exact", 2: "expsum"}.get(m.value, "?"), terms=t.value, step=h.value, fallbacks=fb.value,
                    c_min=lo.value, c_max=hi.value, ms_nodes=tn.value, ms_traits=tt.value)

    def timing(self):
        t = C.c_double(0); n = C.c_longlong(0)
        ph = (C.c_double * _lib.BSFG_N_PHASES)()
        check(self.lib.bsfg_get_timing(self._ctx, C.byref(t), C.byref(n), ph))
        names = ("lambda_prep", "lambda", "lambda_chol", "means", "h2", "F_a", "F_gemm", "F_comm", "F_solve",
                 "precisions", "delta", "update_k")
        return t.value, n.value, dict(zip(names, list(ph)))

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.bsfg_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------
# fast_BSFG_sampler (R_BSFG/fast_BSFG_sampler.R) -- same entry point, state dict in, state dict out
# ------------------------------------------------------------------------------------------
def _save_posterior_samples(sp_num, Posterior, cs, traces_only=False):
    """BSFG_functions.R:375-408: full traces for Lambda, F, F_a, delta, F_h2 and the precisions (rows grow
    when update_k adds factors), running means for B and E_a."""
    def put(name, vec):
        vec = np.asarray(vec, dtype=np.float64).ravel(order="F")
        M = Posterior.get(name)
        if M is None or M.size == 0:
            M = np.zeros((vec.shape[0], sp_num))
        if M.shape[1] < sp_num:
            M = np.column_stack([M, np.zeros((M.shape[0], sp_num - M.shape[1]))])
        if vec.shape[0] > M.shape[0]:
            M = np.vstack([M, np.zeros((vec.shape[0] - M.shape[0], M.shape[1]))])
        M[:vec.shape[0], sp_num - 1] = vec
        Posterior[name] = M
    for name in ("Lambda", "F", "F_a", "delta", "F_h2", "resid_Y_prec", "E_a_prec", "W_prec"):
        put(name, cs[name])
    for name in (() if traces_only else ("B", "E_a", "W")):
        prev = Posterior.get(name)
        cur = np.asarray(cs[name], dtype=np.float64)
        if prev is None or np.shape(prev) != cur.shape:
            prev = np.zeros_like(cur)
        Posterior[name] = (prev * (sp_num - 1) + cur) / sp_num
    return Posterior


def fast_BSFG_sampler(BSFG_state, n_samples, *, seed=None, sweep=None):
    """Run `n_samples` iterations of the BSFG Gibbs sampler on the GPU (fast_BSFG_sampler.R:1).

    BSFG_state: dict with data_matrices, run_parameters, run_variables, priors, current_state and
    (optional) Posterior, as produced by the init (fast_BSFG_sampler_init.R:352-362).
    `sweep`: an existing BSFGSweep to reuse (keeps the rotated data resident across calls); otherwise
    one is created and closed here.  Returns the updated BSFG_state (new dict; inputs are not mutated).
    """
    rp = BSFG_state["run_parameters"]
    cs = dict(BSFG_state["current_state"])
    Posterior = dict(BSFG_state.get("Posterior") or {})
    burn, thin = int(rp.get("burn", 0)), int(rp.get("thin", 1))
    start_i = int(cs.get("nrun", 0))
    own = sweep is None
    if own:
        if seed is None:
            seed = int(BSFG_state.get("RNG", {}).get("seed", 0)) if isinstance(BSFG_state.get("RNG"), dict) else 0
        sweep = BSFGSweep(BSFG_state["data_matrices"], BSFG_state["run_variables"], BSFG_state["priors"], rp,
                          k_init=np.asarray(cs["F"]).shape[1], seed=seed)
    try:
        sweep.set_state(cs)
        end = start_i + int(n_samples)
        # samples that fall inside this call (fast_BSFG_sampler.R:272): kept on the device, read back once at the end
        n_keep = sum(1 for i in range(start_i + 1, end + 1) if i > burn and (i - burn) % thin == 0) if thin > 0 else 0
        if n_keep > 0:
            sweep.posterior_begin(n_keep)
            sweep.collect(int(n_samples), burn, thin)
            post = sweep.posterior()
            km = post["k_max"]
            for s in range(len(post["sp_num"])):
                kk = int(post["k"][s])
                sample = dict(Lambda=post["Lambda"][:, s].reshape(sweep.pl, km, order="F")[:, :kk],
                              F=post["F"][:, s].reshape(sweep.n, km, order="F")[:, :kk],
                              F_a=post["F_a"][:, s].reshape(sweep.r, km, order="F")[:, :kk],
                              delta=post["delta"][:kk, s], F_h2=post["F_h2"][:kk, s],
                              resid_Y_prec=post["resid_Y_prec"][:, s], E_a_prec=post["E_a_prec"][:, s], W_prec=post["W_prec"][:, s])
                Posterior = _save_posterior_samples(int(post["sp_num"][s]), Posterior, sample, traces_only=True)
            # running means of B, E_a, W: combine the device means of this call's samples with what was there before
            first, last = int(post["sp_num"][0]), int(post["sp_num"][-1])
            for name in ("B", "E_a", "W"):
                cur = post[name]
                prev = Posterior.get(name)
                if prev is None or np.shape(prev) != cur.shape or first == 1:
                    Posterior[name] = cur
                else:
                    # the device mean was started at sample `first` with the reference's recursion seeded by zero, i.e. it
                    # equals (sum of this call's samples) / last; add the earlier samples' share
                    Posterior[name] = prev * (first - 1) / last + cur
        else:
            sweep.sweep(int(n_samples))
        new_cs = sweep.get_state(with_E_a=True)
    finally:
        if own:
            sweep.close()
    out = dict(BSFG_state)
    out["current_state"] = new_cs
    out["Posterior"] = Posterior
    return out
