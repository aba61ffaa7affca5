#!/usr/bin/env python
"""bench.py -- Gibbs iterations/second of the BSFG sweep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C4]

A "step" is one Gibbs iteration (the loop body of fast_BSFG_sampler, R_BSFG/fast_BSFG_sampler.R:176-270)
on synthetic data of the named shape.  Default workload: C4 = n 5000 individuals, p 20000 traits, k 100
factors, X = 1_n, Z_1 = I_n, random three-generation pedigree; traits are sharded
over the N ranks (total work fixed -> "strong" scaling).  One JSON line is printed by rank 0.

  value        iterations/s with the chain state resident in HBM (K iterations inside ONE bsfg_sweep call,
               timed with CUDA events on the library's stream; max over ranks)
  e2e          the same metric through the public host API per iteration: pinned host state -> set_state
               (H2D) -> one iteration -> get_state of the posterior-sample fields (D2H)
  roofline     the weighted-Gram FP64 GEMM (dgemm_tn_kernel, DMMA + TMA): algorithmic k(k+1) n p_local
               FLOP / its CUDA-event duration, vs the cuBLAS DGEMM rate measured on this pool
  cpu_baseline the oracle (restatement of the reference formulation) timed on this box's host cores on a
               bounded trait sample (rank 0, N = 1 only)

Synthetic inputs and the three diagonalisations are generated here (torch on the GPU when present, for
speed; not part of any timed region).  The oracle is only imported by the CPU-baseline legs.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, p, k)            BASELINE.json configs
    "C2": (1000, 1000, 20),
    "C3": (200, 10000, 50),
    "C4": (5000, 20000, 100),
    "C5": (20000, 50000, 200),
    "tiny": (96, 160, 8),
}
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
NCU_SUMMARY_FILE = os.path.join(ROOT, "profiles", "ncu_gram_gemm.json")


# --------------------------------------------------------------------------------------------
# synthetic problem (recipe: generate_simulations_halfSib.R:20-37,72-89)
# --------------------------------------------------------------------------------------------
def random_pedigree_A(n, rng, frac_founders=0.2, frac_gen1=0.3):
    n0 = max(2, int(n * frac_founders)); n1 = max(2, int(n * frac_gen1)); n2 = n - n0 - n1
    A = np.zeros((n, n))
    A[:n0, :n0] = np.eye(n0)
    lo, prev = n0, (0, n0)
    for cnt in (n1, n2):
        if cnt <= 0:
            continue
        hi = lo + cnt
        sire = rng.integers(prev[0], prev[1], size=cnt)
        dam = rng.integers(prev[0], prev[1], size=cnt)
        same = sire == dam
        dam[same] = prev[0] + (dam[same] - prev[0] + 1) % (prev[1] - prev[0])
        A[lo:hi, :lo] = 0.5 * (A[sire, :lo] + A[dam, :lo])
        A[:lo, lo:hi] = A[lo:hi, :lo].T
        blk = 0.5 * (A[lo:hi, :lo][:, sire] + A[lo:hi, :lo][:, dam])
        blk = 0.5 * (blk + blk.T)
        np.fill_diagonal(blk, 1.0 + 0.5 * A[sire, dam])
        A[lo:hi, lo:hi] = blk
        prev, lo = (lo, hi), hi
    return A


def make_problem(n, p, k, seed=20131, device=None, lists_route="device", start="prior"):
    """Returns (data_matrices, run_variables, priors, run_parameters, current_state) as host NumPy.

    lists_route  "device": the product's own init (bsfg_init_lists) when a GPU is present;
                 "host": the same eigen route through torch.linalg, never importing the product (the reference arm).
    start        "prior": initial state drawn from the priors (fast_BSFG_sampler_init.R:127-216);
                 "truth": k_true = k simulated factors and the chain started at the simulated Lambda / F / F_a with flat
                 shrinkage, so that no column is redundant and update_k keeps k near its start value (the C5 live leg)."""
    import torch
    dev = torch.device(device if device is not None else ("cuda:0" if torch.cuda.is_available() else "cpu"))
    f64 = torch.float64
    rng = np.random.default_rng(seed)
    gen = torch.Generator(device=dev); gen.manual_seed(seed)
    A_np = random_pedigree_A(n, rng)
    A = torch.from_numpy(A_np).to(dev)
    k_true = k if start == "truth" else max(2, k // 2)
    factor_h2s = np.linspace(1.0, 0.0, k_true + 1)[:k_true].copy()
    factor_h2s[0::2] = 0.0
    Lam_true = np.zeros((p, k_true))
    lo, hi = max(1, p // 30), max(2, p // 4)
    for h in range(k_true):
        numeff = int(rng.integers(lo, hi + 1))
        rows = rng.choice(p, size=numeff, replace=False)
        Lam_true[rows, h] = rng.standard_normal(numeff)
    Lt_ = torch.from_numpy(Lam_true).to(dev)
    h2 = torch.from_numpy(factor_h2s).to(dev)
    L_A = torch.linalg.cholesky(A)
    rn = lambda *s: torch.randn(*s, dtype=f64, device=dev, generator=gen)
    Fa_true = L_A @ rn(n, k_true) * h2.sqrt()[None, :]
    Fe_true = rn(n, k_true) * (1 - h2).sqrt()[None, :]
    Y = Fa_true @ Lt_.T
    Y += L_A @ (np.sqrt(0.2) * rn(n, p))
    Y += Fe_true @ Lt_.T
    Y += np.sqrt(0.2) * rn(n, p)
    Fa_true_h, F_true_h = Fa_true.cpu().numpy(), (Fa_true + Fe_true).cpu().numpy()
    del Fa_true, Fe_true
    X = torch.ones(n, 1, dtype=f64, device=dev)
    b = 1
    # --- the three diagonalisations (contract of fast_BSFG_sampler_init.R:263-325): the library's own device init
    #     (bsfg_init_lists, cuSOLVER) when a GPU is present, else the same eigen route through torch on the CPU
    A_host = np.asfortranarray(A.cpu().numpy())
    X_host = np.asfortranarray(X.cpu().numpy())
    lists = None
    if dev.type == "cuda" and lists_route == "device":
        del A, L_A
        torch.cuda.synchronize(); torch.cuda.empty_cache()
        import sparsefactormixedmodel_b200 as pkg
        with torch.cuda.device(dev):
            lists = pkg.init_lists(None, A_host, X_host, None, b_X_prec=1e-10)
    else:
        s, U = torch.linalg.eigh(A)
        s = s.flip(0).clamp_min(0.0).contiguous(); U = U.flip(1).contiguous()
        Ainv = torch.linalg.inv(A); Ainv = 0.5 * (Ainv + Ainv.T)
        br = b + n
        P = torch.zeros(br, br, dtype=f64, device=dev); P[0, 0] = 1e-10; P[b:, b:] = Ainv
        D = torch.cat([X, torch.eye(n, dtype=f64, device=dev)], dim=1)
        D2 = D.T @ D
        Lc = torch.linalg.cholesky(P + D2)
        Li = torch.linalg.solve_triangular(Lc, torch.eye(br, dtype=f64, device=dev), upper=False)
        M = Li @ D2 @ Li.T; M = 0.5 * (M + M.T)
        lam, V = torch.linalg.eigh(M)
        lam = lam.clamp(0.0, 1.0)
        U2 = Li.T @ V
        t2 = s / (1.0 + s)                      # Z_1 = I: U3 = U diag(t), U3'U3 = t^2 = s1, U3'Ainv U3 = t^2/s = s2
        cpu_ = lambda t: np.asfortranarray(t.cpu().numpy())
        lists = dict(Ainv=cpu_(Ainv), invert_aI_bZAZ=dict(U=cpu_(U), s=s.cpu().numpy()),
                     invert_aPXA_bDesignDesignT=dict(U=cpu_(U2), s1=(1.0 - lam).cpu().numpy(), s2=lam.cpu().numpy(),
                                                     Design_U=cpu_(D @ U2)),
                     invert_aZZt_Ainv=dict(U=cpu_(U * t2.sqrt()[None, :]), s1=t2.cpu().numpy(), s2=(1.0 - t2).cpu().numpy()))
    H = 50
    run_parameters = dict(b0=1.0, b1=0.0005, epsilon=1e-2, prop=1.00, h2_divisions=H, save_freq=100, burn=2000,
                          thin=400, draw_iter=200, simulation=True)
    priors = dict(k_init=k, resid_Y_prec_shape=2.0, resid_Y_prec_rate=0.1, E_a_prec_shape=2.0, E_a_prec_rate=0.1,
                  W_prec_shape=2.0, W_prec_rate=0.1, Lambda_df=3.0, delta_1_shape=2.1, delta_1_rate=1 / 20,
                  delta_2_shape=4.5, delta_2_rate=1.0, b_X_prec=1e-10 * np.eye(b),
                  h2_priors_factors=np.concatenate([[H - 1.0], np.ones(H - 1)]) / (2.0 * (H - 1)))
    # --- initial state from the priors (fast_BSFG_sampler_init.R:127-216)
    g = lambda shape, rate, size: rng.gamma(shape, 1.0, size=size) / rate
    resid_Y_prec = g(2.0, 0.1, p)
    Lambda_prec = g(1.5, 1.5, (p, k))
    delta = np.concatenate([g(2.1, 1 / 20, 1), g(4.5, 1.0, k - 1)])
    tauh = np.cumprod(delta)
    Plam = Lambda_prec * tauh[None, :]
    Lambda = rng.standard_normal((p, k)) * np.sqrt(1.0 / Plam)
    E_a_prec = g(2.0, 0.1, p)
    F_h2 = rng.uniform(size=k)
    F_a = rng.standard_normal((n, k)) * np.sqrt(F_h2)[None, :]
    F = F_a + rng.standard_normal((n, k)) * np.sqrt(1 - F_h2)[None, :]
    B = rng.standard_normal((b, p))
    if start == "truth":
        delta = np.ones(k); delta[0] = 2.0
        tauh = np.cumprod(delta)
        Plam = Lambda_prec * tauh[None, :]
        Lambda = Lam_true.copy()
        F_h2 = np.clip(factor_h2s, 0.0, 0.98)
        F, F_a = F_true_h, Fa_true_h
        B = np.zeros((b, p))
    cpu = lambda t: np.asfortranarray(t.cpu().numpy())
    data = dict(Y=cpu(Y), X=X_host, Z_1=np.eye(n), Z_2=np.zeros((n, 0)))
    rv = dict(p=p, n=n, r=n, r2=0, b=b, Ainv=lists["Ainv"], invert_aI_bZAZ=lists["invert_aI_bZAZ"],
              invert_aPXA_bDesignDesignT=lists["invert_aPXA_bDesignDesignT"], invert_aZZt_Ainv=lists["invert_aZZt_Ainv"])
    cs = dict(resid_Y_prec=resid_Y_prec, Lambda_prec=Lambda_prec, delta=delta, tauh=tauh, Plam=Plam, Lambda=Lambda,
              F_h2=F_h2, E_a_prec=E_a_prec, F_a=F_a, F=F, B=B, E_a=None, nrun=0)
    if dev.type == "cuda":
        torch.cuda.synchronize(); torch.cuda.empty_cache()
    return data, rv, priors, run_parameters, cs


HBM_PEAK_GBS_FALLBACK = 6553.9


def iteration_roofline(n, pl, k, b, phases, gstats, tensor_tf, ms_iter):
    """Whole-iteration speed of light: per phase, floor = max(FLOP / FP64 tensor rate, bytes / HBM copy rate) with the FLOP and
    byte counts of what the device formulation executes (tools/phase_floors.py), next to the CUDA-event
    phase times of the last timed iteration.  The three longest phases are listed with their own fractions."""
    from tools.phase_floors import floors
    hbm = HBM_PEAK_GBS_FALLBACK
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rows, tot_f, tot_m = [], 0.0, 0.0
    for name, flop, byts, what in floors(n, pl, k, b, R=gstats["terms"] if gstats["mode"] == "expsum" else 0, exact=gstats["mode"] != "expsum"):
        t_t, t_b = flop / (tensor_tf * 1e12) * 1e3, byts / (hbm * 1e9) * 1e3
        m = phases.get(name, 0.0) + (phases.get("F_comm", 0.0) if name == "F_solve" else 0.0)
        f = max(t_t, t_b)
        tot_f += f; tot_m += m
        rows.append({"phase": name, "ms": m, "floor_ms": f, "frac": (f / m) if m > 0 else None, "bound": "tensor" if t_t >= t_b else "hbm",
                     "gflop": flop * 1e-9, "gbytes": byts * 1e-9})
    top = sorted(rows, key=lambda r_: -r_["ms"])[:3]
    return {"floor_ms": tot_f, "phases_ms": tot_m, "ms_per_iteration": ms_iter, "frac": tot_f / ms_iter if ms_iter > 0 else None,
            "hbm_gbs": hbm, "tensor_tflops": tensor_tf, "top3_by_time": top, "all": rows}


def algorithmic_flops(n, p, k, b=1):
    """SURVEY.md section 8d W_flop without the fitted-value term (pre-rotation removes it)."""
    return (k * (k + 1) * n * p + 2 * k * n * p + 2 * n * p * k + 2 * n * p * k + 2 * p * k * k
            + p * (k ** 3 / 3 + 3 * k * k) + 2 * k * k * n)


# --------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark(self):
        """Samples taken before this call (warm-up) are dropped, except the last one if the timed region turns out shorter
        than one sampling period."""
        self.first = max(0, len(self.lines) - 1)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines[self.first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU legs: the reference itself on the host cores (the ONLY place bench.py touches oracle/)
# --------------------------------------------------------------------------------------------
def _reference_checker():
    """(module, kind, description): the compiled reference (oracle/_ref: BSFG_functions_c.cpp built unmodified against the
    stand-in RcppArmadillo.h + SciPy OpenBLAS; the R loop body around it is oracle/bsfg_oracle.gibbs_iteration, R itself
    cannot run here), else the C++ restatement oracle/c."""
    from oracle import ref_oracle as R
    if R.available():
        R.load()
        return R, "reference", ("oracle/_ref/libbsfg_reference.so = the reference's BSFG_functions_c.cpp compiled unmodified (stand-in "
                                "RcppArmadillo.h over SciPy OpenBLAS/LAPACK); the R loop body fast_BSFG_sampler.R:176-270 around the _c calls "
                                "is restated in NumPy (no R in the image) with the reference's own operations incl. dense Z_1 E_a and the "
                                "p x p t(E_a) Ainv E_a product")
    from oracle import c_oracle as CO
    return CO, "port", "oracle/c/bsfg_ref.cpp (C++ restatement of the Rcpp path in the reference formulation, SciPy OpenBLAS)"


def _slice_traits(data, cs, sl):
    d = dict(Y=np.asfortranarray(data["Y"][:, sl]), X=data["X"], Z_1=data["Z_1"])
    st = {kk: vv for kk, vv in cs.items() if vv is not None}
    for nm in ("Lambda", "Lambda_prec", "Plam"):
        st[nm] = np.asarray(st[nm])[sl, :]
    for nm in ("resid_Y_prec", "E_a_prec"):
        st[nm] = np.asarray(st[nm])[sl]
    st["B"] = np.asarray(st["B"])[:, sl]
    return d, st


def reference_iterations(data, rv, priors, rp, cs, n, p, k, traits=None, iters=1, seed=1, warm=True):
    """Times `iters` Gibbs iterations of the reference on the host cores, on all p traits (traits=None) or on the first
    `traits` of them.  Returns (seconds per iteration [list], kind, description, cores, blas threads)."""
    from oracle import init_oracle as I
    M, kind, what = _reference_checker()
    cores = os.cpu_count() or 1
    M.blas_threads(cores)
    rng = np.random.default_rng(seed)
    if warm:                      # BLAS thread pool + page-in on a tiny trait block (not timed)
        dw, sw_ = _slice_traits(data, cs, slice(0, min(p, 64)))
        vw = I.draw_variates(rng, n, dw["Y"].shape[1], k, n, rv["b"] + n, priors)
        _one_reference_iteration(M, kind, dw, rv, priors, rp, sw_, vw)
    ps = p if traits is None else min(traits, p)
    d, st = (dict(Y=data["Y"], X=data["X"], Z_1=data["Z_1"]), {kk: vv for kk, vv in cs.items() if vv is not None}) if ps == p \
        else _slice_traits(data, cs, slice(0, ps))
    times = []
    for _ in range(iters):
        v = I.draw_variates(rng, n, ps, k, n, rv["b"] + n, priors)
        t0 = time.perf_counter()
        st = _one_reference_iteration(M, kind, d, rv, priors, rp, st, v)
        times.append(time.perf_counter() - t0)
    return times, kind, what, int(cores), int(M.blas_threads()), ps


def _one_reference_iteration(M, kind, d, rv, priors, rp, st, v):
    if kind == "reference":
        nxt, _ = M.gibbs_iteration(d, rv, priors, rp, st, 1, v, with_update_k=False)   # i = 1 <= 200: update_k gated off either way
    else:
        nxt, _ = M.gibbs_iteration_c(d, rv, priors, rp, st, v, trait_threads=1)
    return nxt


def cpu_baseline(data, rv, priors, rp, cs, n, p, k, sample_traits):
    """The bounded CPU leg printed beside the GPU number: ONE iteration of the reference on the first `sample_traits` traits
    at full n and k, scaled by p / sample (per-trait work dominates and is linear in p; the p^2 r term of
    fast_BSFG_sampler.R:245 is under-counted by the sample, in the reference's favour; the p-independent F_h2 / F_a / F
    steps are over-counted by the scaling -- < 1 % of the iteration).  `--impl reference` measures the whole workload."""
    times, kind, what, cores, threads, ps = reference_iterations(data, rv, priors, rp, cs, n, p, k, traits=sample_traits)
    t = times[0]
    return {"value": 1.0 / (t * (p / ps)), "unit": "iter/s", "cores": cores, "kind": kind, "extrapolated": ps != p,
            "sample": f"{what}; {threads} BLAS threads; 1 iteration on {ps} of {p} traits at full n={n}, k={k}: {t:.2f} s, "
                      f"scaled x{p / ps:.1f}"}


# --------------------------------------------------------------------------------------------
def dist_env():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def bcast_obj(obj, src=0):
    import torch.distributed as dist
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def bcast_problem(prob, rank, local):
    """Rank 0's problem (nested dicts / tuples of NumPy arrays and scalars) to every rank: the structure goes as a small
    pickled object, every array >= 1 MB as a CUDA tensor through NCCL (a pickled 20 GB C5 problem would take minutes)."""
    import torch
    import torch.distributed as dist
    big = []

    def strip(o):
        if isinstance(o, dict):
            return {k: strip(v) for k, v in o.items()}
        if isinstance(o, (tuple, list)):
            return type(o)(strip(v) for v in o)
        if isinstance(o, np.ndarray) and o.nbytes >= (1 << 20):
            big.append(o)
            return ("__big__", len(big) - 1, o.shape, str(o.dtype), bool(o.flags.f_contiguous and o.ndim == 2))
        return o

    skel = strip(prob) if rank == 0 else None
    skel = bcast_obj(skel)
    metas = []

    def collect(o):
        if isinstance(o, dict):
            for v in o.values():
                collect(v)
        elif isinstance(o, tuple) and len(o) == 5 and o[0] == "__big__":
            metas.append(o)
        elif isinstance(o, (tuple, list)):
            for v in o:
                collect(v)

    collect(skel)
    metas.sort(key=lambda m: m[1])
    arrays = []
    for _, idx, shape, dtype, fortran in metas:
        if rank == 0:
            a = big[idx]
            flat = np.ascontiguousarray(a.T if fortran else a).reshape(-1)
            t = torch.from_numpy(flat).to(f"cuda:{local}")
        else:
            t = torch.empty(int(np.prod(shape)), dtype=getattr(torch, dtype), device=f"cuda:{local}")
        dist.broadcast(t, src=0)
        if rank == 0:
            arrays.append(big[idx])
        else:
            h = t.cpu().numpy()
            arrays.append(h.reshape(shape[::-1]).T if fortran else h.reshape(shape))
        del t
    torch.cuda.empty_cache()

    def fill(o):
        if isinstance(o, dict):
            return {k: fill(v) for k, v in o.items()}
        if isinstance(o, tuple) and len(o) == 5 and o[0] == "__big__":
            return arrays[o[1]]
        if isinstance(o, (tuple, list)):
            return type(o)(fill(v) for v in o)
        return o

    return fill(skel)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, world, local = dist_env()
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the BSFG sweep has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    import sparsefactormixedmodel_b200 as pkg
    from sparsefactormixedmodel_b200 import BSFGSweep
    lib = pkg.load()
    n, p, k = WORKLOADS[args.workload]

    t_setup = time.time()
    # C5 (BASELINE.json configs[4]) runs with update_k live (start at iteration 200: the i > 200 gate of BSFG_functions.R:331
    # is open from the first iteration).  Two legs: --c5-leg held  = k fixed at its start value (iterations <= 200);
    #                                               --c5-leg live  = update_k live on a chain started at the simulated truth
    #                                                                with k_true = k, so k stays near its start value
    held = args.workload != "C5" or args.c5_leg == "held"
    if rank == 0:
        prob = make_problem(n, p, k, device=f"cuda:{local}", start="prior" if held else "truth")
    else:
        prob = None
    if world > 1:
        prob = bcast_problem(prob, rank, local)          # setup only, outside every timed region
    data, rv, priors, rp, cs = prob
    nccl_id = None
    if world > 1:
        nccl_id = bcast_obj(BSFGSweep.nccl_unique_id() if rank == 0 else None)
    start_iter = args.start_iter if args.start_iter >= 0 else (0 if held else 200)
    k_max = k if start_iter + args.warmup + args.steps + 8 <= 200 else min(k + 16, pkg.k_limit() if hasattr(pkg, "k_limit") else 216)
    cs = dict(cs); cs["nrun"] = start_iter
    live_k = start_iter + args.warmup + args.steps > 200
    sw = BSFGSweep(data, rv, priors, rp, k_init=k, k_max=k_max, seed=20131, rank=rank, world=world, nccl_id=nccl_id,
                   gram=args.gram)
    sw.set_state(cs)
    mode, r1, r2 = sw.mode()
    t_setup = time.time() - t_setup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput -------------------------------------------------------
    clocks = ClockSampler(local); clocks.start()      # nvidia-smi needs ~0.2 s to deliver its first line: start it before the warm-up
    sw.sweep(args.warmup)
    barrier()
    clocks.mark()                                     # samples from here on belong to the timed region
    launches0 = lib.bsfg_launch_count()
    wall0 = time.perf_counter()
    k_per_iter = None
    if not live_k:
        sw.sweep(args.steps)                  # K iterations, one call, CUDA events inside on the sweep stream
        barrier()
        dev_ms, _, phases = sw.timing()
    else:                                     # k moves: one call per iteration so k can be reported per timed iteration
        dev_ms, k_per_iter, phases = 0.0, [], None
        for _ in range(args.steps):
            sw.sweep(1)
            ms1, _, phases = sw.timing()
            dev_ms += ms1
            k_per_iter.append(int(sw.k()))
        barrier()
    wall = time.perf_counter() - wall0
    launches = lib.bsfg_launch_count() - launches0
    clk = clocks.stop()
    dev_ms = max_over_ranks(dev_ms)
    value = args.steps / (dev_ms * 1e-3)
    gstats = sw.gram_stats()                 # Gram evaluation in use + its two GEMM timings (last timed iteration)

    # ---- end to end through the host API, per iteration -----------------------------------
    def pinned(a):
        a = np.asfortranarray(np.asarray(a, dtype=np.float64))
        t = torch.empty(a.shape[::-1], dtype=torch.float64).pin_memory()   # F-order view of a C-order pinned block
        v = t.numpy().T
        v[...] = a
        return v
    def pinned_empty(shape):
        shape = tuple(shape)
        t = torch.empty(shape[::-1], dtype=torch.float64).pin_memory()
        return t.numpy().T if len(shape) > 1 else t.numpy()
    k_now = sw.get_state(with_E_a=False)["F"].shape[1]                # update_k may have changed k (C5)
    k_trace = [k, k_now]
    host = sw.alloc_state(k_now, with_E_a=False, empty=pinned_empty)  # pinned result buffers, reused every step
    sw.get_state(with_E_a=False, out=host)
    # e2e starts from the state the timed sweep left: every rank keeps ITS shard of the trait-indexed members in pinned
    # host memory (set_state(local=True)), exactly what a chain driver holding per-rank state would pass back
    cur = sw.get_state(with_E_a=False)
    full = dict(cur)
    for nm in ("Lambda", "Lambda_prec", "Plam", "F", "F_a", "B"):
        full[nm] = pinned(full[nm])
    e2e_steps = args.steps
    kk_ = k_now
    h2d = 8 * (3 * sw.pl * kk_ + 2 * n * kk_ + sw.b * sw.pl + 2 * sw.pl + 3 * kk_)     # Lambda, Lambda_prec, Plam, F, F_a, B, precisions, delta/tauh/F_h2
    d2h = 8 * (3 * sw.pl * kk_ + n * kk_ + n * kk_ + kk_ + sw.b * sw.pl + 2 * sw.pl + 2 * kk_)
    # N > 1: the replicated members (F, F_a, F_h2, delta, tauh -- bit-identical on every rank) cross PCIe ONCE per direction:
    # rank 0 uploads them and the library hands them to the other GPUs over NVLink (bsfg_state::bcast_replicated), and only
    # rank 0 reads them back; every rank moves its own trait shard both ways.  h2d / d2h bytes below are rank 0's (the largest).
    rep = world == 1 or rank == 0
    sw.set_state(full, local=True, bcast_replicated=True); sw.sweep(1); sw.get_state(with_E_a=False, out=host, with_replicated=rep)   # warm the path
    barrier()
    t0 = time.perf_counter()
    e2e_parts = [0.0, 0.0, 0.0]
    for _ in range(e2e_steps):
        ta = time.perf_counter()
        sw.set_state(full, local=True, bcast_replicated=True)
        tb = time.perf_counter()
        sw.sweep(1)
        tc = time.perf_counter()
        out = sw.get_state(with_E_a=False, out=host, with_replicated=rep)
        td = time.perf_counter()
        e2e_parts[0] += tb - ta; e2e_parts[1] += tc - tb; e2e_parts[2] += td - tc
        if not live_k and rep:
            full.update({nm: out[nm] for nm in ("F", "F_a", "F_h2", "delta", "tauh")})   # replicated members carry over
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_val = e2e_steps / e2e_s

    if live_k:
        args.no_exact_leg = True          # k moves under update_k: the fixed-k roofline legs do not apply
    npairs = k_now * (k_now + 1) // 2
    pl_local = sw.pl
    if gstats["mode"] == "expsum":
        # dominant tensor kernel: the trait product Q = (G'A)' E, M = k(k+1)/2, N = p_local, K = R nodes
        kern_flop = 2 * npairs * gstats["terms"] * sw.pl
        kern_ms = gstats["ms_traits"]
        kern_name = (f"dgemm_tn_kernel<128,128,6> (exponential-sum weighted Gram, trait product Q = (G'A)'E: "
                     f"M=k(k+1)/2={npairs}, N=p_local={sw.pl}, K=R={gstats['terms']})")
    else:
        kern_flop = k_now * (k_now + 1) * n * sw.pl
        kern_ms = phases["lambda"]
        kern_name = "dgemm_tn_kernel<128,128,6> (exact weighted Gram Q = G'W, M=k(k+1)/2, N=p_local, K=n)"
    achieved = kern_flop / (kern_ms * 1e-3) * 1e-12 if kern_ms > 0 else None
    # DRAM bytes of that launch: the ncu --set full figure when this very launch shape was captured (profiles/), else null;
    # the model (operands read once + result written once) is reported beside it for every shape
    traffic = None
    if os.path.exists(NCU_SUMMARY_FILE):
        try:
            cap = json.load(open(NCU_SUMMARY_FILE))
            if cap.get("shape_" + gstats["mode"]) == [int(npairs), int(sw.pl), int(gstats["terms"] if gstats["mode"] == "expsum" else n)]:
                traffic = cap.get("dram_bytes_per_launch_" + gstats["mode"])
        except Exception:
            traffic = None
    kdim = gstats["terms"] if gstats["mode"] == "expsum" else n
    traffic_model = 8.0 * (npairs * sw.pl + kdim * npairs + kdim * sw.pl)

    # ---- the exact (reference-order) Gram evaluation, timed beside the default: the full-size DMMA GEMM -------
    exact_leg = None
    if gstats["mode"] != "exact" and not args.no_exact_leg:
        sw.close()
        sw = None
        ex = BSFGSweep(data, rv, priors, rp, k_init=k, k_max=k, seed=20131, rank=rank, world=world, nccl_id=None if world == 1 else
                       bcast_obj(BSFGSweep.nccl_unique_id() if rank == 0 else None), gram="exact")
        ex.set_state(cs)
        ex.sweep(3)
        barrier()
        ex.sweep(3)
        barrier()
        ex_ms, _, ex_ph = ex.timing()
        ex_ms = max_over_ranks(ex_ms)
        ex_flop = k * (k + 1) * n * ex.pl
        ex_ach = ex_flop / (ex_ph["lambda"] * 1e-3) * 1e-12
        exact_leg = {"value": 3 / (ex_ms * 1e-3), "unit": "iter/s", "ms_per_step": ex_ms / 3,
                     "roofline": {"bound": "tensor", "achieved": ex_ach, "unit": "TFLOP/s",
                                  "kernel": "dgemm_tn_kernel<128,128,6> (exact weighted Gram Q = G'W, M=k(k+1)/2, N=p_local, K=n)",
                                  "flop_per_launch": ex_flop, "ms_per_launch": ex_ph["lambda"]}}
        ex.close()

    if rank != 0:
        if sw is not None:
            sw.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peak_tf, peak_src = 35.4, "fallback constant (profiles/r01_fp64_peaks.md)"
    if os.path.exists(FP64_PEAK_FILE):
        pk = json.load(open(FP64_PEAK_FILE))
        peak_tf = float(pk.get("dgemm_8192_sustained_tflops", peak_tf))
        peak_src = "cuBLAS DGEMM 8192^3 measured on this pool (profiles/r01_fp64_peak.json; MEASURED_PEAKS.json has no FP64 entry)"
    if exact_leg:
        exact_leg["roofline"]["peak"] = peak_tf
        exact_leg["roofline"]["frac"] = exact_leg["roofline"]["achieved"] / peak_tf
    line = {
        "metric": "Gibbs iter/sec at n=5000,p=20000,k=100 (1/2/4/8 B200) vs Rcpp host" if args.workload == "C4"
                  else f"Gibbs iter/sec at n={n},p={p},k={k}",
        "value": value, "unit": "iter/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT.format(w=args.workload, n=n, p=p, k=k),
                   "parallelism": f"trait-shard x{world}", "l2": "inputs larger than L2 (rotated data 2.4 GB/rank-set, "
                   "Gram operands 1 GB) -- no explicit flush", "formulation": mode,
                   "diag_identity_residuals": [r1, r2],
                   "update_k": (f"LIVE: iterations {start_iter + 1}.. (i > 200), chain started at the simulated truth (k_true = k), k_max {k_max}"
                                if live_k else "gated off by i<=200 (uu still drawn)"),
                   "k_per_timed_iteration": k_per_iter if live_k else k,
                   "gram": {kk: gstats[kk] for kk in ("mode", "terms", "step", "fallbacks", "c_min", "c_max")},
                   "rng": "device Philox4x32-10", "setup_s": round(t_setup, 1), "wall_s_timed_region": round(wall, 3)},
        "clocks": clk,
        "e2e": {"value": e2e_val, "unit": "iter/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "what": f"{e2e_steps} x [set_state(pinned host current_state) -> bsfg_sweep(1) -> get_state(posterior-sample fields, pinned)]; "
                        "E_a (r x p) is resampled before use and only read back on request"
                        + ("; replicated members (F, F_a, F_h2, delta, tauh) cross PCIe once per direction on rank 0 (NCCL broadcast "
                           "to the other GPUs on the way in, read back by rank 0 only); every rank moves its own trait shard; "
                           "bytes are rank 0's" if world > 1 else ""),
                "ms_per_step_parts": {"set_state_h2d": 1e3 * e2e_parts[0] / e2e_steps, "sweep": 1e3 * e2e_parts[1] / e2e_steps,
                                      "get_state_d2h": 1e3 * e2e_parts[2] / e2e_steps}},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic, "traffic_model": traffic_model,
                     "kernel": kern_name,
                     "flop_per_launch": kern_flop, "ms_per_launch": kern_ms, "peak_source": peak_src,
                     "dmma_issue_peak_tflops": 37.1},
        "phases_ms_last_iter": phases,
        "phases_note": "CUDA-event intervals of the last timed iteration, which the library runs on one stream so that every interval "
                       "measures the kernels it brackets; the other iterations overlap the F-only draws and the delta chain on a side stream",
        "iteration_roofline": iteration_roofline(n, pl_local, k_now, 1, phases, gstats, peak_tf, dev_ms / args.steps),
        "sweep_algorithmic_tflops": algorithmic_flops(n, p, k) / world / (dev_ms / args.steps * 1e-3) * 1e-12,
        "sweep_algorithmic_tflops_note": "SURVEY.md 8d W_flop of the exact formulation (k(k+1)np weighted Gram) per GPU / time; in expsum "
                                         "mode fewer FLOPs are executed, so this can exceed the FP64 peak",
    }
    if exact_leg:
        line["exact_gram"] = exact_leg
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(data, rv, priors, rp, cs, n, p, k, args.cpu_sample_traits)
    print(json.dumps(line), flush=True)
    if sw is not None:
        sw.close()
    if world > 1:
        dist.destroy_process_group()


WORKLOAD_TEXT = "{w}: n={n} individuals, p={p} traits, k={k} factors, b=1, Z_1=I, r=n, H=50, random 3-generation pedigree"


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref, the unmodified
    BSFG_functions_c.cpp) on the host cores, on the WHOLE workload -- every trait, the dense Z_1 E_a products and the
    p x p temporary of fast_BSFG_sampler.R:245 included.  One whole iteration at C4 is about a minute of 16 cores, so the arm
    times `--ref-steps` (default 1) iterations, however many --steps asks for, and says so in `steps_timed`; nothing is
    extrapolated.  The problem is built without the product (torch.linalg on the CPU).  Rank 0 only."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    n, p, k = WORKLOADS[args.workload]
    t0 = time.perf_counter()
    data, rv, priors, rp, cs = make_problem(n, p, k, device="cpu", lists_route="host")
    t_setup = time.perf_counter() - t0
    steps = max(1, args.ref_steps)
    times, kind, what, cores, threads, _ = reference_iterations(data, rv, priors, rp, cs, n, p, k, traits=None, iters=steps)
    t_it = float(np.mean(times))
    value = 1.0 / t_it
    line = {"impl": "reference",
            "metric": "Gibbs iter/sec at n=5000,p=20000,k=100 (1/2/4/8 B200) vs Rcpp host" if args.workload == "C4"
                      else f"Gibbs iter/sec at n={n},p={p},k={k}",
            "value": value, "unit": "iter/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "steps_timed": steps, "timed_region_s": float(np.sum(times)), "s_per_iteration": [round(t, 3) for t in times],
            "warmup_note": "one untimed iteration on a 64-trait block (BLAS thread pool, page-in); whole-workload warm-up iterations "
                           "would cost a minute each",
            "ms_per_step": 1e3 * t_it, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT.format(w=args.workload, n=n, p=p, k=k), "setup_s": round(t_setup, 1)},
            "cpu_baseline": {"value": value, "unit": "iter/s", "cores": cores, "kind": kind, "extrapolated": False,
                             "sample": f"{what}; {threads} BLAS threads; {steps} whole iteration(s) on all {p} traits"},
            "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C4", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-traits", type=int, default=1000, help="trait sample of the bounded CPU leg beside the GPU number")
    ap.add_argument("--ref-steps", type=int, default=1, help="--impl reference: whole-workload iterations actually timed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gram", default="default", choices=["default", "expsum", "exact"],
                    help="weighted-Gram evaluation of the Lambda draw (DESIGN.md section 3a)")
    ap.add_argument("--no-exact-leg", action="store_true", help="skip the extra exact-Gram timing beside the default")
    ap.add_argument("--c5-leg", default="live", choices=["live", "held"],
                    help="C5 only: update_k live on a chain started at the simulated truth (k stays near 200), or k held (i <= 200)")
    ap.add_argument("--start-iter", type=int, default=-1, help="current_state$nrun at entry (default 0; 200 for C5 so update_k is live)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
