// CPU oracle in C++ -- TEST INFRASTRUCTURE AND CPU BASELINE ONLY (nothing under sparsefactormixedmodel_b200/ links it).
//
// A second restatement of the reference's Gibbs iteration IN THE REFERENCE'S FORMULATION AND OPERATION ORDER, against
// the OpenBLAS that ships inside SciPy (LP64, `scipy_`-prefixed CBLAS/LAPACKE), so that the CPU baseline of bench.py is
// compiled code calling BLAS/LAPACK exactly where RcppArmadillo would:
//   sample_Lambda_c            R_BSFG/BSFG_functions_c.cpp:86-135   (U'Y_tilde; per trait: FUDi, k x n . n x k, chol, 3 solves)
//   sample_means_c             cpp:43-82                              (Design_U'Y_tilde; per trait gemv U (..))
//   sample_h2s_discrete_c      cpp:196-238
//   sample_F_a_c               cpp:293-339
//   sample_factors_scores_c    cpp:138-163
//   Lambda_prec / resid_Y_prec / E_a_prec / delta / tauh / Plam      R_BSFG/fast_BSFG_sampler.R:235-257, BSFG_functions.R:272-308
//     (E_a_prec through the full t(E_a) %*% Ainv %*% E_a product and its p x p temporary, as R:245 does)
// Variates are injected like in oracle/bsfg_oracle.py, against which tests/test_oracle.py checks this file to 1e-10.
// PARITY UNPINNED in the same sense as the NumPy oracle (no reference outputs exist; see its header).
//
// trait_threads == 1: the reference's order, parallelism only inside BLAS (what R + a threaded BLAS gives).
// trait_threads  > 1: "baseline-opt" of BASELINE.md: the per-trait loops run under OpenMP with single-threaded BLAS.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include <omp.h>

extern "C" {
enum { RowMajor = 101, ColMajor = 102, NoTrans = 111, Trans = 112, Upper = 121, Lower = 122, NonUnit = 131, Left = 141 };
void scipy_cblas_dgemm(int, int, int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
void scipy_cblas_dgemv(int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
void scipy_cblas_dtrsv(int, int, int, int, int, const double*, int, double*, int);
void scipy_cblas_dtrsm(int, int, int, int, int, int, int, double, const double*, int, double*, int);
int scipy_LAPACKE_dpotrf(int, char, int, double*, int);
void scipy_openblas_set_num_threads(int);
int scipy_openblas_get_num_threads(void);
}

namespace {
typedef std::vector<double> vec;
// C (m x n) = alpha op(A) op(B) + beta C, all column-major
inline void gemm(bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
                 double beta, double* C, int ldc) {
    scipy_cblas_dgemm(ColMajor, ta ? Trans : NoTrans, tb ? Trans : NoTrans, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
const double LOG_INV_SQRT_2PI = std::log(1.0 / std::sqrt(2.0 * M_PI));   // cpp:23
}  // namespace

extern "C" int bsfg_ref_blas_threads(int set) {
    if (set > 0) scipy_openblas_set_num_threads(set);
    return scipy_openblas_get_num_threads();
}

extern "C" int bsfg_ref_iteration(
    int n, int p, int k, int r, int b, int H, const double* Y, const double* X, const double* Z1, const double* U1,
    const double* s, const double* U2, const double* s1_2, const double* s2_2, const double* DU, const double* U3,
    const double* s1_3, const double* s2_3, const double* Ainv, const double* h2_priors, const double* pri,
    double* Lambda, double* Lambda_prec, double* Plam, double* F, double* F_a, double* F_h2, double* B, double* E_a,
    double* resid_Y_prec, double* E_a_prec, double* delta, double* tauh, const double* z_Lambda, const double* z_means,
    const double* u_h2, const double* z_Fa, const double* z_F, const double* g_LP, const double* g_resid, const double* g_Ea,
    const double* g_delta, int trait_threads, double* rates_out /* 2p: resid_Y_prec, E_a_prec rates (may be null) */) {
    const int br = b + r;
    const double rr_shape = pri[0], rr_rate = pri[1], ea_shape = pri[2], ea_rate = pri[3], df = pri[4];
    const double d1s = pri[5], d1r = pri[6], d2s = pri[7], d2r = pri[8];
    (void)rr_shape; (void)ea_shape;
    const int blas_threads = scipy_openblas_get_num_threads();
    const int tt = std::max(1, trait_threads);
    auto per_trait_begin = [&]() { if (tt > 1) scipy_openblas_set_num_threads(1); };
    auto per_trait_end = [&]() { if (tt > 1) scipy_openblas_set_num_threads(blas_threads); };
    int bad = 0;

    vec Yt((size_t)n * p);
    // ---- Lambda: Y_tilde = Y - X B                                                     R:189-191 -> cpp:86-135
    std::memcpy(Yt.data(), Y, sizeof(double) * (size_t)n * p);
    if (b > 0) gemm(false, false, n, p, b, -1.0, X, n, B, b, 1.0, Yt.data(), n);
    {
        vec FtU((size_t)k * n), UtY((size_t)n * p);
        gemm(true, false, k, n, n, 1.0, F, n, U1, n, 0.0, FtU.data(), k);                  // FtU = F.t() * U      cpp:107
        gemm(true, false, n, p, n, 1.0, U1, n, Yt.data(), n, 0.0, UtY.data(), n);          // UtY = U.t() * Y_tilde cpp:108
        per_trait_begin();
#pragma omp parallel for num_threads(tt) schedule(dynamic, 8) if (tt > 1)
        for (int j = 0; j < p; j++) {                                                       // cpp:119-132
            vec FUDi((size_t)k * n), Q((size_t)k * k), means(k), y(k);
            const double ep = E_a_prec[j], c = ep / resid_Y_prec[j];
            for (int i = 0; i < n; i++) {                                                   // sweep_times(FtU,2,1/(s+c)) * ep
                const double w = ep * (1.0 / (s[i] + c));
                for (int a = 0; a < k; a++) FUDi[(size_t)i * k + a] = FtU[(size_t)i * k + a] * w;
            }
            scipy_cblas_dgemv(ColMajor, NoTrans, k, n, 1.0, FUDi.data(), k, UtY.data() + (size_t)j * n, 1, 0.0, means.data(), 1);
            gemm(false, true, k, k, n, 1.0, FUDi.data(), k, FtU.data(), k, 0.0, Q.data(), k);   // FUDi * FtU.t()
            for (int a = 0; a < k; a++) Q[(size_t)a * k + a] += Plam[(size_t)a * p + j];         // + diagmat(Plam.row(j))
            if (scipy_LAPACKE_dpotrf(ColMajor, 'U', k, Q.data(), k) != 0) {                      // Llam_t = chol(Qlam)
#pragma omp atomic write
                bad = 1;
                continue;
            }
            scipy_cblas_dtrsv(ColMajor, Upper, Trans, NonUnit, k, Q.data(), k, means.data(), 1);     // vlam = solve(Llam_t.t(), means)
            scipy_cblas_dtrsv(ColMajor, Upper, NoTrans, NonUnit, k, Q.data(), k, means.data(), 1);   // mlam = solve(Llam_t, vlam)
            for (int a = 0; a < k; a++) y[a] = z_Lambda[(size_t)j * k + a];
            scipy_cblas_dtrsv(ColMajor, Upper, NoTrans, NonUnit, k, Q.data(), k, y.data(), 1);       // ylam = solve(Llam_t, Zcol)
            for (int a = 0; a < k; a++) Lambda[(size_t)a * p + j] = y[a] + means[a];
        }
        per_trait_end();
        if (bad) return -2;
    }
    // ---- B, E_a: Y_tilde = Y - F Lambda'                                               R:203-207 -> cpp:43-82
    std::memcpy(Yt.data(), Y, sizeof(double) * (size_t)n * p);
    gemm(false, true, n, p, k, -1.0, F, n, Lambda, p, 1.0, Yt.data(), n);
    {
        vec means((size_t)br * p), loc((size_t)br * p);
        gemm(true, false, br, p, n, 1.0, DU, n, Yt.data(), n, 0.0, means.data(), br);       // Design_U.t() * Y_tilde   cpp:65
        per_trait_begin();
#pragma omp parallel for num_threads(tt) schedule(dynamic, 8) if (tt > 1)
        for (int j = 0; j < p; j++) {                                                       // cpp:75-79
            vec v(br);
            for (int i = 0; i < br; i++) {
                const double d = s1_2[i] * E_a_prec[j] + s2_2[i] * resid_Y_prec[j];
                const double mlam = (means[(size_t)j * br + i] * resid_Y_prec[j]) / d;
                v[i] = mlam + z_means[(size_t)j * br + i] / std::sqrt(d);
            }
            scipy_cblas_dgemv(ColMajor, NoTrans, br, br, 1.0, U2, br, v.data(), 1, 0.0, loc.data() + (size_t)j * br, 1);
        }
        per_trait_end();
        for (int j = 0; j < p; j++) {                                                       // R:206-207
            for (int c = 0; c < b; c++) B[(size_t)j * b + c] = loc[(size_t)j * br + c];
            for (int i = 0; i < r; i++) E_a[(size_t)j * r + i] = loc[(size_t)j * br + b + i];
        }
    }
    // ---- F_h2                                                                          R:221 -> cpp:196-238
    {
        vec sb((size_t)k * n), logps((size_t)k * H);
        gemm(true, false, k, n, n, 1.0, F, n, U1, n, 0.0, sb.data(), k);                     // std_scores_b = F.t() * U
        for (int i = 0; i < H; i++) {
            const double h2 = (double)i / (double)H;
            double det = 0.0;
            if (h2 > 0) for (int l = 0; l < n; l++) det += std::log((s[l] + (1 - h2) / h2) * h2) / 2.0;
            for (int j = 0; j < k; j++) {
                double acc = 0.0;
                for (int l = 0; l < n; l++) {
                    const double x = (h2 > 0) ? 1.0 / std::sqrt(h2) * (sb[(size_t)l * k + j] * (1.0 / std::sqrt(s[l] + (1 - h2) / h2)))
                                              : F[(size_t)j * n + l];
                    acc += LOG_INV_SQRT_2PI - 0.5 * (x * x);
                }
                logps[(size_t)i * k + j] = acc - det + std::log(h2_priors[i]);
            }
        }
        for (int j = 0; j < k; j++) {                                                       // cpp:228-235
            double mx = -INFINITY, sum = 0.0;
            for (int i = 0; i < H; i++) mx = std::max(mx, logps[(size_t)i * k + j]);
            for (int i = 0; i < H; i++) sum += std::exp(logps[(size_t)i * k + j] - mx);
            const double norm = mx + std::log(sum);
            double cum = 0.0;
            int count = 0;
            for (int i = 0; i < H; i++) { cum += std::exp(logps[(size_t)i * k + j] - norm); if (u_h2[j] > cum) count++; }
            F_h2[j] = (double)count / (double)H;
        }
    }
    // ---- F_a                                                                           R:226 -> cpp:293-339
    {
        vec Fs((size_t)n * k), T((size_t)r * k), bb((size_t)r * k), v(r);
        for (int j = 0; j < k; j++) { const double te = 1.0 / (1.0 - F_h2[j]); for (int i = 0; i < n; i++) Fs[(size_t)j * n + i] = F[(size_t)j * n + i] * te; }
        gemm(true, false, r, k, n, 1.0, Z1, n, Fs.data(), n, 0.0, T.data(), r);              // Z_1.t() * sweep_times(F,2,tau_e)
        gemm(true, false, r, k, r, 1.0, U3, r, T.data(), r, 0.0, bb.data(), r);              // b = U.t() * (..)          cpp:311
        for (int j = 0; j < k; j++) {
            const double te = 1.0 / (1.0 - F_h2[j]), tu = 1.0 / F_h2[j];
            double* out = F_a + (size_t)j * r;
            if (te == 1) { for (int i = 0; i < r; i++) out[i] = 0.0; }
            else if (te > 10000) { for (int i = 0; i < r; i++) out[i] = (i < n) ? F[(size_t)j * n + i] : 0.0; }
            else {
                for (int i = 0; i < r; i++) {
                    const double d = s2_3[i] * tu + s1_3[i] * te;
                    v[i] = bb[(size_t)j * r + i] / d + z_Fa[(size_t)j * r + i] / std::sqrt(d);
                }
                scipy_cblas_dgemv(ColMajor, NoTrans, r, r, 1.0, U3, r, v.data(), 1, 0.0, out, 1);
            }
        }
    }
    // ---- F: Y_tilde = Y - X B - Z_1 E_a                                                R:230-232 -> cpp:138-163
    std::memcpy(Yt.data(), Y, sizeof(double) * (size_t)n * p);
    if (b > 0) gemm(false, false, n, p, b, -1.0, X, n, B, b, 1.0, Yt.data(), n);
    gemm(false, false, n, p, r, -1.0, Z1, n, E_a, r, 1.0, Yt.data(), n);                     // dense Z_1 %*% E_a
    {
        vec Lmsg((size_t)p * k), S((size_t)k * k), Meta((size_t)n * k), ZFa((size_t)n * k);
        for (int a = 0; a < k; a++) for (int j = 0; j < p; j++) Lmsg[(size_t)a * p + j] = Lambda[(size_t)a * p + j] * resid_Y_prec[j];
        gemm(true, false, k, k, p, 1.0, Lambda, p, Lmsg.data(), p, 0.0, S.data(), k);
        for (int a = 0; a < k; a++) S[(size_t)a * k + a] += 1.0 / (1.0 - F_h2[a]);
        if (scipy_LAPACKE_dpotrf(ColMajor, 'U', k, S.data(), k) != 0) return -2;            // S = chol(..)               cpp:149
        gemm(false, false, n, k, p, 1.0, Yt.data(), n, Lmsg.data(), p, 0.0, Meta.data(), n);
        gemm(false, false, n, k, r, 1.0, Z1, n, F_a, r, 0.0, ZFa.data(), n);
        for (int a = 0; a < k; a++) { const double te = 1.0 / (1.0 - F_h2[a]); for (int i = 0; i < n; i++) Meta[(size_t)a * n + i] += ZFa[(size_t)a * n + i] * te; }
        // Meta' <- solve(tS, Meta')  <=>  Meta <- Meta S^-1   (right side, upper, no transpose)
        scipy_cblas_dtrsm(ColMajor, 142 /*Right*/, Upper, NoTrans, NonUnit, n, k, 1.0, S.data(), k, Meta.data(), n);
        for (size_t e = 0; e < (size_t)n * k; e++) Meta[e] += z_F[e];
        // F' = solve(S, (Meta + Z)')  <=>  F = (Meta + Z) S^-T
        scipy_cblas_dtrsm(ColMajor, 142, Upper, Trans, NonUnit, n, k, 1.0, S.data(), k, Meta.data(), n);
        std::memcpy(F, Meta.data(), sizeof(double) * (size_t)n * k);
    }
    // ---- Lambda_prec                                                                   R:235-236
    for (int a = 0; a < k; a++)
        for (int j = 0; j < p; j++) {
            const double l = Lambda[(size_t)a * p + j];
            Lambda_prec[(size_t)a * p + j] = g_LP[(size_t)a * p + j] / ((df + (l * l) * tauh[a]) / 2.0);
        }
    // ---- resid_Y_prec: Y_tilde = Y - X B - F Lambda' - Z_1 E_a                          R:239-240
    std::memcpy(Yt.data(), Y, sizeof(double) * (size_t)n * p);
    if (b > 0) gemm(false, false, n, p, b, -1.0, X, n, B, b, 1.0, Yt.data(), n);
    gemm(false, true, n, p, k, -1.0, F, n, Lambda, p, 1.0, Yt.data(), n);
    gemm(false, false, n, p, r, -1.0, Z1, n, E_a, r, 1.0, Yt.data(), n);
    for (int j = 0; j < p; j++) {
        double ss = 0.0;
        for (int i = 0; i < n; i++) { const double v = Yt[(size_t)j * n + i]; ss += v * v; }
        const double rate = rr_rate + 0.5 * ss;
        if (rates_out) rates_out[j] = rate;
        resid_Y_prec[j] = g_resid[j] / rate;
    }
    // ---- E_a_prec: diag(t(E_a) %*% Ainv %*% E_a) through the full p x p product, as R:245 writes it
    {
        vec AE((size_t)r * p), PP((size_t)p * p);
        gemm(false, false, r, p, r, 1.0, Ainv, r, E_a, r, 0.0, AE.data(), r);
        gemm(true, false, p, p, r, 1.0, E_a, r, AE.data(), r, 0.0, PP.data(), p);
        for (int j = 0; j < p; j++) {
            const double rate = ea_rate + 0.5 * PP[(size_t)j * p + j];
            if (rates_out) rates_out[p + j] = rate;
            E_a_prec[j] = g_Ea[j] / rate;
        }
    }
    // ---- delta, tauh, Plam                                                             R:253-257 -> BSFG_functions.R:272-308
    {
        vec colsum(k, 0.0);
        for (int a = 0; a < k; a++) for (int j = 0; j < p; j++) { const double l = Lambda[(size_t)a * p + j]; colsum[a] += Lambda_prec[(size_t)a * p + j] * (l * l); }
        auto cumprod = [&]() { double cp = 1.0; for (int a = 0; a < k; a++) { cp *= delta[a]; tauh[a] = cp; } };
        auto tail = [&](int from) { double acc = 0.0; for (int a = from; a < k; a++) acc += tauh[a] * colsum[a]; return acc; };
        int draw = 0;
        {
            const double rate = d1r + 0.5 * (1.0 / delta[0]) * tail(0);
            (void)d1s;
            delta[0] = g_delta[draw++] / rate;
            cumprod();
        }
        const int hi = k - 1, step = (hi >= 2) ? 1 : -1;
        for (int h = 2; (step > 0) ? (h <= hi) : (h >= hi); h += step) {                   // R's for(h in 2:(k-1))
            const double rate = d2r + 0.5 * (1.0 / delta[h - 1]) * tail(h - 1);
            (void)d2s;
            delta[h - 1] = g_delta[draw++] / rate;
            cumprod();
        }
        for (int a = 0; a < k; a++) for (int j = 0; j < p; j++) Plam[(size_t)a * p + j] = Lambda_prec[(size_t)a * p + j] * tauh[a];
    }
    return 0;
}
