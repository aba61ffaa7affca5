"""Trait sharding, host side, on CPU: world_size-2 `gloo` processes.

The device sweep shards trait columns over ranks and exchanges exactly two buffers per iteration
(csrc/sweep.inl `allreduce` call sites):
   red1 = [Lambda' diag(rp) Lambda | U'Y Lmsg | theta Lmsg]       -> the F draw (cpp:149,152)
   red2 = [colSums(Lambda_prec * Lambda^2) | #{|Lambda| < eps}]    -> sample_delta / update_k (R:288, 326)
This test runs that decomposition with NumPy on each rank's shard (`shard_bounds`, the same slicing
`BSFGSweep.set_state` / `sweep_injected` apply), all-reduces over gloo, and checks that every rank reconstructs
the oracle's full-problem F draw, delta draw and update_k decision.  It also covers the host plumbing that
needs >1 process: ragged `shard_bounds`, broadcasting the 128-byte NCCL id blob, max-over-ranks timing.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        from scipy.linalg import solve_triangular
        from oracle import bsfg_oracle as O, init_oracle as I
        from sparsefactormixedmodel_b200.driver import shard_bounds
        dist.init_process_group("gloo", rank=rank, world_size=world)

        n, p, k = 25, 19, 4                                   # ragged: 10 + 9 traits
        setup = I.make_synthetic(n, p, 2, pedigree="halfsib", seed=9)
        pri = I.default_priors(k_init=k); rp = I.default_run_parameters()
        st = I.sampler_init(setup, pri, rp, seed=10, lists="stable")
        d, rv, cs = st["data_matrices"], st["run_variables"], st["current_state"]
        v = I.draw_variates(np.random.default_rng(11), n, p, k, n, 1 + n, pri)
        full, _ = O.gibbs_iteration(d, rv, pri, rp, cs, 1, v)      # every rank holds the full answer to check against

        lo, hi = shard_bounds(p, rank, world)
        bounds = [None] * world
        dist.all_gather_object(bounds, (lo, hi))
        assert bounds[0][0] == 0 and bounds[-1][1] == p and all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
        assert max(h - l for l, h in bounds) - min(h - l for l, h in bounds) <= 1

        # the NCCL id blob travels as a Python object from rank 0 (what bench.py / multi_gpu_check.py do)
        blob = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(blob, src=0)
        assert blob[0] == bytes(range(128))

        # ---- per-shard sufficient statistics for the F draw, in rotated coordinates -------------------
        sl = slice(lo, hi)
        Y, X, Z = d["Y"], d["X"], d["Z_1"]
        U1 = rv["invert_aI_bZAZ"]["U"]
        l2 = rv["invert_aPXA_bDesignDesignT"]
        b = X.shape[1]
        Lam, rp_ = full["Lambda"][sl], cs["resid_Y_prec"][sl]
        Lmsg = Lam * rp_[:, None]
        loc = np.vstack([full["B"].reshape(b, p), full["E_a"]])[:, sl]
        theta = np.linalg.solve(l2["U"], loc)                  # location_sample = U theta (cpp:78)
        part = np.concatenate([(Lam.T @ Lmsg).ravel(), ((U1.T @ Y[:, sl]) @ Lmsg).ravel(), (theta @ Lmsg).ravel()])
        t = torch.from_numpy(part.copy())
        dist.all_reduce(t)
        tot = t.numpy()
        LtL = tot[:k * k].reshape(k, k)
        P1 = tot[k * k:k * k + n * k].reshape(n, k)
        P2 = tot[k * k + n * k:].reshape(-1, k)
        YL = U1 @ P1 - l2["Design_U"] @ P2                       # = (Y - X B - Z_1 E_a) Lmsg summed over ALL traits
        tau_e = 1.0 / (1.0 - full["F_h2"])
        S = np.linalg.cholesky(LtL + np.diag(tau_e)).T
        Meta = solve_triangular(S.T, (YL + (Z @ full["F_a"]) * tau_e[None, :]).T, lower=True).T
        F = solve_triangular(S, (Meta + v["z_F"]).T, lower=False).T
        err = np.max(np.abs(F - full["F"])) / np.max(np.abs(full["F"]))
        assert err < 1e-9, err

        # ---- delta / update_k statistics ----------------------------------------------------------
        LP = full["Lambda_prec"][sl]
        part2 = np.concatenate([np.sum(LP * Lam ** 2, axis=0), np.sum(np.abs(Lam) < rp["epsilon"], axis=0).astype(float)])
        t2 = torch.from_numpy(part2.copy())
        dist.all_reduce(t2)
        colsum, cnt = t2.numpy()[:k], t2.numpy()[k:]
        assert np.allclose(colsum, np.sum(full["Lambda_prec"] * full["Lambda"] ** 2, axis=0), rtol=1e-13)
        assert np.array_equal(cnt / p, np.mean(np.abs(full["Lambda"]) < rp["epsilon"], axis=0))

        # ---- max-over-ranks timing reduction used by bench.py ------------------------------------------
        tm = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        assert tm.item() == float(world)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "".join(traceback.format_exception(type(e), e, e.__traceback__))))


def test_sharded_statistics_reconstruct_the_full_draws_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=240) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


@pytest.mark.parametrize("p,world", [(20000, 8), (7, 3), (5, 8), (131, 2)])
def test_shard_bounds_partition(p, world):
    from sparsefactormixedmodel_b200.driver import shard_bounds
    b = [shard_bounds(p, r, world) for r in range(world)]
    assert b[0][0] == 0 and b[-1][1] == p
    assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
    sizes = [h - l for l, h in b]
    assert max(sizes) - min(sizes) <= 1
