"""Numerical prototype (NumPy, CPU) of an INT8-slice emulation of the sweep's FP64 GEMMs -- the 'Ozaki scheme I' that
B200's tcgen05 INT8 tensor pipe (4.5 POPS nominal, ~100x its DMMA rate per operation) would run.  Not product code and
not used by it: it answers, on BSFG-shaped operands, how many 7-bit slices the contraction needs to stay inside the
parity tolerance (1e-9) and inside plain FP64 GEMM rounding, so that DESIGN.md section 9 can size the kernel
(number of int8 GEMMs = s (s + 1) / 2 for s slices when cross terms with t + u > s + 1 are dropped).

    C = A' B,  A: K x M, B: K x N (the library's TN form).
    Row scaling: every column of A (a row of A') and every column of B is scaled by a power of two to (-1, 1), then cut
    into s signed 7-bit digits q_1..q_s (|q| <= 127) by truncation; digit products accumulate EXACTLY in int32 for
    K < 2^31 / 127^2 = 133k; the s (s + 1) / 2 partial products are combined in FP64 with their power-of-two weights.

Run:  python tools/ozaki_prototype.py            (prints a table; ~20 s)
"""
import numpy as np

BITS = 7


def slices(X, s):
    """X: K x M, scaled per column.  Returns (digits [s, K, M] int8-valued int32, exponents [M])."""
    amax = np.abs(X).max(axis=0)
    e = np.where(amax > 0, np.ceil(np.log2(np.where(amax > 0, amax, 1.0))) + 1, 0.0)    # |x| 2^-e < 0.5
    r = X * np.exp2(-e)[None, :]
    out = np.empty((s,) + X.shape, dtype=np.int32)
    for t in range(s):
        r = r * (1 << BITS)
        q = np.trunc(r)
        out[t] = q.astype(np.int32)
        r = r - q                               # exact: r has fewer significant bits each round
    assert np.abs(out).max() <= 127
    return out, e


def gemm_int8_slices(A, B, s):
    qa, ea = slices(A, s)
    qb, eb = slices(B, s)
    M, N = A.shape[1], B.shape[1]
    C = np.zeros((M, N))
    n_gemm = 0
    for lvl in range(2 * s, 1, -1):             # smallest weights first
        acc = np.zeros((M, N), dtype=np.int64)  # pairs of one level share a weight: summed as integers (TMEM int32 in the kernel)
        for t in range(1, s + 1):
            u = lvl - t
            if 1 <= u <= s and t + u <= s + 1:
                acc += qa[t - 1].T.astype(np.int64) @ qb[u - 1].astype(np.int64)
                n_gemm += 1
        C += acc.astype(np.float64) * 2.0 ** (-BITS * lvl)
    return C * np.exp2(ea)[:, None] * np.exp2(eb)[None, :], n_gemm


def report(name, A, B):
    ref = (A.astype(np.longdouble).T @ B.astype(np.longdouble)).astype(np.float64)
    bound = np.abs(A).T @ np.abs(B)             # componentwise scale of the FP64 rounding bound
    f64 = A.T @ B
    e64 = np.abs(f64 - ref)
    print(f"{name}: K={A.shape[0]} M={A.shape[1]} N={B.shape[1]}   plain FP64 GEMM: err/|A|'|B| = {np.max(e64 / bound):.1e}, "
          f"err/max|C| = {e64.max() / np.abs(ref).max():.1e}")
    for s in (4, 5, 6, 7, 8):
        C, ng = gemm_int8_slices(A, B, s)
        err = np.abs(C - ref)
        print(f"   s={s} ({ng:2d} int8 GEMMs): err/|A|'|B| = {np.max(err / bound):.1e}   err/max|C| = {err.max() / np.abs(ref).max():.1e}"
              f"   worst elementwise rel = {np.max(err / np.maximum(np.abs(ref), 1e-300)):.1e}")


def main():
    rng = np.random.default_rng(1)
    n, p, k = 600, 300, 12
    # rotated factor scores and data, eigenvalues of a pedigree-like relationship matrix, per-trait precisions
    UtF = rng.standard_normal((n, k)) * np.sqrt(rng.uniform(0.3, 2.5, n))[:, None]
    UtY = rng.standard_normal((n, p)) * rng.uniform(0.5, 3.0, p)[None, :]
    s = np.sort(rng.gamma(2.0, 0.6, n))[::-1]
    rp, ep = rng.gamma(2.0, 5.0, p), rng.gamma(2.0, 5.0, p)
    W = 1.0 / (s[:, None] / ep[None, :] + 1.0 / rp[None, :])          # cpp:120: ep rp / (ep + s rp) up to the arrangement
    report("m_j = (U'F)'(w . U'y)   [a1, K = n]", UtF, W * UtY)
    iu = np.triu_indices(k)
    G = UtF[:, iu[0]] * UtF[:, iu[1]]
    report("Q_j = G'w               [a1 exact Gram, K = n]", G, W)
    # exponential-sum form: nodes x traits weights with a huge dynamic range down each column
    t = -8.0 + 0.28 * np.arange(64)
    c = rng.uniform(0.01, 400.0, p)
    E = 0.28 * np.exp(t)[:, None] * np.exp(-np.exp(t)[:, None] * c[None, :])
    Ht = rng.standard_normal((64, k * (k + 1) // 2)) ** 2
    report("Q = Ht'E                [a1 exp-sum trait product, K = R]", Ht, E)
    Lmsg = rng.standard_normal((p, k)) * rp[:, None]
    report("P1 = (U'Y) Lmsg         [a4, K = p]", np.ascontiguousarray(UtY.T), Lmsg)


if __name__ == "__main__":
    main()
