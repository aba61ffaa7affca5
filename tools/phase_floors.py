"""Per-phase speed-of-light of one Gibbs iteration next to the measured CUDA-event phase times of a bench.py line.

    python tools/phase_floors.py profiles/r01b_bench_c4_n1.json > profiles/r01b_phase_floors.md

Floors use the MEASURED ceilings of this pool: FP64 tensor rate = cuBLAS DGEMM (profiles/r01_fp64_peak.json) and the HBM copy
rate of MEASURED_PEAKS.json (6553.9 GB/s).  FLOP and byte counts are those of the device formulation
(DESIGN.md section 3): per phase, floor = max(FLOP / tensor rate, bytes / HBM rate); phases that are neither (the Cholesky's
dependency chains, the FP64 Box-Muller) are listed with what bounds them instead.
"""
import json
import re
import sys

TENSOR = 35.45e12          # FLOP/s, cuBLAS DGEMM 8192^3 on this pool
HBM = 6553.9e9             # B/s


def floors(n, p, k, b, R=256, exact=False):
    br, npair = n + b, k * (k + 1) // 2
    D = 8.0
    rows = []
    add = lambda name, flop, byts, what: rows.append((name, flop, byts, what))
    add("lambda_prep", 0, D * (n * npair + 2 * n * p + n * R + R * p),
        "Gram features (write n x k(k+1)/2), weights (read U'Y, write w.U'y), exp-sum nodes A and E")
    if exact:
        add("lambda", 2.0 * npair * n * p, D * npair * p, "exact weighted Gram Q = G'W (K = n); writes Q")
    else:
        add("lambda", 2.0 * R * npair * n + 2.0 * npair * R * p, D * npair * p,
            "node product (G'A) and trait product Q = (G'A)'E; writes Q")
    add("lambda_chol", 2.0 * k * n * p + p * (k ** 3 / 3.0 + 3.0 * k * k), D * (npair * p + n * p),
        "m = (U'F)'(w.U'y), then p Cholesky factorisations + 3 solves of order k (latency bound: dependency chains of 8x8 tiles)")
    add("means", 2.0 * br * p * k, D * 5 * br * p,
        "Design_U'Y - (Design_U'F)Lambda' (read + write), theta draw (read, write theta and theta', 1e8 FP64 Box-Muller normals)")
    add("h2", 0, D * k * n * 2, "grid log-likelihoods from L2-resident U'F")
    add("F_a", 2 * 2.0 * n * n * k, D * 2 * n * n, "U3'F and U3 v (two n x n x k products)")
    add("F_gemm", 2 * 2.0 * n * p * k + 2.0 * p * k * k, D * (2 * n * p), "Lambda'Lmsg, Y Lmsg, theta Lmsg")
    add("F_solve", 2.0 * n * n * k + 2.0 * n * k * k, D * n * n, "Design_U (theta Lmsg) (Y Lmsg needs no rotation), k x k Cholesky, n row solves")
    add("precisions", 2 * 2.0 * n * n * k + 2 * 2.0 * n * p * k + 2.0 * k * k * p + 2.0 * k * k * n, D * (3 * n * p + 2 * n * n),
        "U'F and Design_U'F for the new F, F'Y, F'Design_U theta, F'F Lambda', per-trait reductions + Gamma draws")
    add("delta", 0, D * 2 * p * k, "column sums of Lambda_prec Lambda^2, sequential delta recursion (k dependent Gamma draws)")
    return rows


def main():
    path = sys.argv[1]
    line = json.loads(open(path).read().strip().splitlines()[-1])
    wl = line["config"]["workload"]
    n, p, k = (int(re.search(rf"\b{c}=(\d+)", wl).group(1)) for c in ("n", "p", "k"))
    ngpu = int(line.get("n_gpus", 1))
    ph = line["phases_ms_last_iter"]
    print(f"# Per-phase floors vs measured ({wl.split(':')[0]}: n={n}, p={p}, k={k}; {ngpu} GPU) -- source `{path}`\n")
    print("Floors: FP64 tensor 35.45 TFLOP/s (measured cuBLAS DGEMM), HBM 6553.9 GB/s (measured copy rate); trait-indexed work / #GPUs.\n")
    print("| phase | GFLOP | GB | floor ms | measured ms | floor / measured | what |")
    print("|---|---|---|---|---|---|---|")
    tot_f = tot_m = 0.0
    for name, flop, byts, what in floors(n, p // ngpu, k, 1):
        f = max(flop / TENSOR, byts / HBM) * 1e3
        m = ph.get(name, float("nan"))
        if name == "F_solve":
            m += ph.get("F_comm", 0.0)
        tot_f += f; tot_m += m
        frac = f"{f / m:.0%}" if m >= 0.02 else "(side stream)"      # bench lines taken before the one-stream phase iteration
        print(f"| {name} | {flop / 1e9:.1f} | {byts / 1e9:.2f} | {f:.2f} | {m:.2f} | {frac} | {what} |")
    print(f"| **sum** | | | **{tot_f:.2f}** | **{tot_m:.2f}** | **{tot_f / line['ms_per_step']:.0%}** of the iteration | iteration: {line['ms_per_step']:.2f} ms |")
    print("\nPhases are CUDA-event intervals of the last timed iteration, which the library runs on ONE stream so that every interval measures")
    print("the kernels it brackets; the other iterations overlap h2, F_a, the k x k factor, the Design_U'F refresh and the delta chain on a")
    print("side stream, which is why the iteration (ms_per_step) is shorter than the sum of the phases.  The fraction of record is floor / iteration.")


if __name__ == "__main__":
    main()
