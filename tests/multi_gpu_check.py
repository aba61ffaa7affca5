"""Trait-sharded sweep check, one process per GPU.  Launched by
tests/test_multi_gpu.py (and by hand under gpurun) as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tests/multi_gpu_check.py

Checks, on every rank:
  1. three injected-variate iterations: this rank's trait shard of the device state matches the oracle's
     full-problem Gibbs iteration to 1e-9 (the oracle is the checker; it runs on every rank, sizes are small);
  2. free-running Philox sweep: the replicated members (F, F_a, F_h2, delta, tauh) are BIT-identical on all
     ranks, and the N-rank chain equals the 1-rank chain (run on rank 0) to 1e-10 -- the Philox element index
     of every trait-indexed draw is the global trait index, so sharding must not change the chain beyond the
     summation order of the allreduce.
  3. the same injected-variate check for the Z_2 + missing-data sweep and for the parameter-expanded sweep.
Prints "multi_gpu_check ok world=N" from rank 0 and exits 0; any mismatch raises.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import relerr, rowwise_relerr
    from oracle import checker as O, init_oracle as I
    from sparsefactormixedmodel_b200 import BSFGSweep

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    def bcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    n, p, k = 96, 131, 8            # p not divisible by the world size: ragged shards
    setup = I.make_synthetic(n, p, 4, pedigree="random", seed=17)
    pri = I.default_priors(k_init=k); rp = I.default_run_parameters()
    st = I.sampler_init(setup, pri, rp, seed=5, lists="stable")
    d, rv, cs = st["data_matrices"], st["run_variables"], st["current_state"]

    nccl_id = bcast(BSFGSweep.nccl_unique_id() if rank == 0 else None)
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=12, seed=99, rank=rank, world=world, nccl_id=nccl_id)
    lo, hi = sw.lo, sw.hi
    tol = 1e-9
    rng = np.random.default_rng(31)
    ora = cs
    for it in range(1, 4):
        v = I.draw_variates(rng, n, p, k, n, 1 + n, pri)
        sw.set_state(ora)
        ora, _ = O.gibbs_iteration(d, rv, pri, rp, ora, it, v)
        sw.sweep_injected(v)
        dev = sw.get_state()
        assert rowwise_relerr(dev["Lambda"], ora["Lambda"][lo:hi]) < tol
        assert rowwise_relerr(dev["F"], ora["F"]) < tol
        assert rowwise_relerr(dev["F_a"].T, ora["F_a"].T) < tol
        assert np.array_equal(dev["F_h2"], np.asarray(ora["F_h2"]).ravel())
        assert relerr(dev["B"], ora["B"].reshape(1, p)[:, lo:hi]) < tol
        assert rowwise_relerr(dev["E_a"].T, ora["E_a"][:, lo:hi].T) < tol
        assert relerr(dev["Lambda_prec"], ora["Lambda_prec"][lo:hi]) < tol
        assert relerr(dev["resid_Y_prec"], ora["resid_Y_prec"][lo:hi]) < tol
        assert relerr(dev["E_a_prec"], ora["E_a_prec"][lo:hi]) < tol
        assert relerr(dev["delta"], ora["delta"]) < tol
        assert relerr(dev["Plam"], ora["Plam"][lo:hi]) < tol * 10

    # free-running Philox chain, N ranks vs 1 rank
    sw.set_state(cs)
    sw.sweep(6)
    got = sw.get_state(with_E_a=False)
    for name in ("F", "F_a", "F_h2", "delta", "tauh"):
        t = torch.from_numpy(np.ascontiguousarray(got[name])).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(t, ref), f"replicated member {name} differs between rank {rank} and rank 0"
    gathered = [None] * world
    dist.all_gather_object(gathered, {nm: got[nm] for nm in ("Lambda", "resid_Y_prec", "E_a_prec", "Lambda_prec")})
    # the same chain with the replicated members supplied by rank 0 only (bsfg_state::bcast_replicated): the other ranks hand
    # over just their trait shards; the result must be bit-identical to the run above
    shard = {nm: cs[nm] for nm in ("Lambda", "Lambda_prec", "Plam", "B", "E_a", "resid_Y_prec", "E_a_prec")}
    shard["nrun"] = cs.get("nrun", 0); shard["k"] = k
    if rank == 0:
        shard.update({nm: cs[nm] for nm in ("F", "F_a", "F_h2", "delta", "tauh")})
    sw.set_state(shard, bcast_replicated=True)
    sw.sweep(6)
    again = sw.get_state(with_E_a=False, with_replicated=(rank == 0))
    for nm in ("Lambda", "resid_Y_prec", "E_a_prec", "Lambda_prec"):
        assert np.array_equal(again[nm], got[nm]), f"bcast_replicated changed {nm} on rank {rank}"
    if rank == 0:
        for nm in ("F", "F_a", "F_h2", "delta", "tauh"):
            assert np.array_equal(again[nm], got[nm]), f"bcast_replicated changed {nm}"
    sw.close()
    if rank == 0:
        one = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=12, seed=99)
        one.set_state(cs)
        one.sweep(6)
        ref = one.get_state(with_E_a=False)
        one.close()
        for nm in ("Lambda", "resid_Y_prec", "E_a_prec", "Lambda_prec"):
            full = np.concatenate([g[nm] for g in gathered], axis=0)
            assert relerr(full, ref[nm]) < 1e-10, (nm, relerr(full, ref[nm]))
        for nm in ("F", "F_a", "delta"):
            assert relerr(got[nm], ref[nm]) < 1e-10, (nm, relerr(got[nm], ref[nm]))
        assert np.array_equal(got["F_h2"], ref["F_h2"])
        print(f"multi_gpu_check ok world={world}", flush=True)
    # ---- second scenario: second random effect (Z_2 / W) + missing phenotypes under trait sharding ------------------
    n, p, k, r2 = 70, 53, 5, 6
    rs = np.random.default_rng(101)
    setup = I.make_synthetic(n, p, 3, pedigree="halfsib", seed=23)
    Y = setup["Y"].copy(); Y[rs.uniform(size=Y.shape) < 0.15] = np.nan
    setup["Y"] = Y
    Z2 = np.zeros((n, r2)); Z2[np.arange(n), rs.integers(0, r2, n)] = 1.0
    setup["Z_2"] = Z2
    pri = I.default_priors(k_init=k)
    st = I.sampler_init(setup, pri, rp, seed=9, lists="stable")
    d, rv, pri, cs = st["data_matrices"], st["run_variables"], st["priors"], st["current_state"]
    nccl_id = bcast(BSFGSweep.nccl_unique_id() if rank == 0 else None)
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=8, seed=7, rank=rank, world=world, nccl_id=nccl_id)
    lo, hi = sw.lo, sw.hi
    sw.set_state(cs)
    ora = cs
    for it in range(1, 4):                       # free running: iterations 2, 3 impute from the device's own theta
        v = I.draw_variates(rs, n, p, k, n, 1 + n, pri, r2=r2, missing=True)
        ora, aux = O.gibbs_iteration(d, rv, pri, rp, ora, it, v)
        sw.sweep_injected(v)
        assert relerr(sw.aux()["Y_imputed"], aux["Y_imputed"][:, lo:hi]) < 1e-8
    dev = sw.get_state()
    assert rowwise_relerr(dev["Lambda"], ora["Lambda"][lo:hi]) < 1e-7
    assert rowwise_relerr(dev["F"], ora["F"]) < 1e-7
    assert rowwise_relerr(dev["W"].T, ora["W"][:, lo:hi].T) < 1e-7
    assert relerr(dev["W_prec"], ora["W_prec"][lo:hi]) < 1e-7
    assert relerr(dev["resid_Y_prec"], ora["resid_Y_prec"][lo:hi]) < 1e-7
    sw.close()
    if rank == 0:
        print(f"multi_gpu_check ok (Z_2 + missing data) world={world}", flush=True)
    # ---- third scenario: parameter-expanded sweep (fast_BSFG_sampler_px.R) under trait sharding --------------------------
    n, p, k = 70, 45, 5
    rs = np.random.default_rng(202)
    setup = I.make_synthetic(n, p, 3, pedigree="halfsib", seed=29)
    pri = I.default_priors(k_init=k)
    st = I.sampler_init(setup, pri, rp, seed=11, lists="stable")
    d, rv, pri, cs = st["data_matrices"], st["run_variables"], st["priors"], st["current_state"]
    ora = dict(cs)
    ora["px_factor"] = rs.gamma(pri["px_shape"], 1.0, k) / pri["px_rate"] + 0.05
    ora["F_px"] = ora["F"] * np.sqrt(ora["px_factor"])[None, :]
    nccl_id = bcast(BSFGSweep.nccl_unique_id() if rank == 0 else None)
    sw = BSFGSweep(d, rv, pri, rp, k_init=k, k_max=8, seed=7, rank=rank, world=world, nccl_id=nccl_id)
    lo, hi = sw.lo, sw.hi
    for it in range(1, 4):
        v = I.draw_variates(rs, n, p, k, n, 1 + n, pri, px=True)
        sw.set_state(ora)
        ora, _ = O.gibbs_iteration_px(d, rv, pri, rp, ora, it, v)
        sw.sweep_injected(v)
        dev = sw.get_state()
        assert rowwise_relerr(dev["Lambda"], ora["Lambda"][lo:hi]) < tol
        assert rowwise_relerr(dev["F_px"], ora["F_px"]) < tol and rowwise_relerr(dev["F"], ora["F"]) < tol
        assert relerr(dev["px_factor"], ora["px_factor"]) < tol
        assert relerr(dev["Plam"], ora["Plam"][lo:hi]) < tol * 10
        assert relerr(dev["resid_Y_prec"], ora["resid_Y_prec"][lo:hi]) < tol
        assert relerr(dev["delta"], ora["delta"]) < tol
    sw.close()
    if rank == 0:
        print(f"multi_gpu_check ok (parameter-expanded sweep) world={world}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
