// Device init: the precomputed diagonalisations of fast_BSFG_sampler_init.R:263-325.
//
// Runs once per chain, O(n^3): it is NOT part of the hot path and uses the vendor dense solvers (cuSOLVER syevd /
// potrf / potri, cuBLAS trsm / gemm), loaded with dlopen so that the sweep itself keeps depending on the CUDA runtime
// only.  The lists keep the reference's CONTRACT
//     inv(a I + b Z A Z')           = 1/b U diag(1/(s + a/b)) U'          (BSFG_functions_c.cpp:98-99)
//     inv(a P + b D'D)              = U diag(1/(a s1 + b s2)) U'          (cpp:52-54; fast_BSFG_sampler_init.R:285-325)
// but are computed through symmetric eigen-decompositions of the SPD pencil (D'D, P + D'D) instead of the reference's
// svd(A inv(B)) (`GSVD_2_c`, cpp:27-40), which loses ~6 digits when [X Z_1] is rank deficient;
// U is not unique (signs, degenerate eigenspaces), the contract is what tests/test_init_gpu.py checks.
// Included by bsfg.cu.

#include <dlfcn.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

namespace bsfg {

struct DenseApi {
    void* hs = nullptr; void* hb = nullptr;
    cusolverDnHandle_t solver = nullptr;
    cublasHandle_t blas = nullptr;
    cusolverStatus_t (*Create)(cusolverDnHandle_t*) = nullptr;
    cusolverStatus_t (*SetStream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
    cusolverStatus_t (*SyevdBuf)(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int, const double*, int, const double*, int*) = nullptr;
    cusolverStatus_t (*Syevd)(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int, double*, int, double*, double*, int, int*) = nullptr;
    cusolverStatus_t (*PotrfBuf)(cusolverDnHandle_t, cublasFillMode_t, int, double*, int, int*) = nullptr;
    cusolverStatus_t (*Potrf)(cusolverDnHandle_t, cublasFillMode_t, int, double*, int, double*, int, int*) = nullptr;
    cusolverStatus_t (*PotriBuf)(cusolverDnHandle_t, cublasFillMode_t, int, double*, int, int*) = nullptr;
    cusolverStatus_t (*Potri)(cusolverDnHandle_t, cublasFillMode_t, int, double*, int, double*, int, int*) = nullptr;
    cusolverStatus_t (*GetrfBuf)(cusolverDnHandle_t, int, int, double*, int, int*) = nullptr;
    cusolverStatus_t (*Getrf)(cusolverDnHandle_t, int, int, double*, int, double*, int*, int*) = nullptr;
    cusolverStatus_t (*Getrs)(cusolverDnHandle_t, cublasOperation_t, int, int, const double*, int, const int*, double*, int, int*) = nullptr;
    cusolverStatus_t (*GesvdBuf)(cusolverDnHandle_t, int, int, int*) = nullptr;
    cusolverStatus_t (*Gesvd)(cusolverDnHandle_t, signed char, signed char, int, int, double*, int, double*, double*, int, double*, int,
                              double*, int, double*, int*) = nullptr;
    cublasStatus_t (*BCreate)(cublasHandle_t*) = nullptr;
    cublasStatus_t (*BSetStream)(cublasHandle_t, cudaStream_t) = nullptr;
    cublasStatus_t (*Trsm)(cublasHandle_t, cublasSideMode_t, cublasFillMode_t, cublasOperation_t, cublasDiagType_t, int, int,
                           const double*, const double*, int, double*, int) = nullptr;
    cublasStatus_t (*Gemm)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const double*, const double*, int,
                           const double*, int, const double*, double*, int) = nullptr;
    static void* open_any(const char* const* names) {
        for (int i = 0; names[i]; i++) {
            void* h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
            if (h) return h;
        }
        return nullptr;
    }
    void load(cudaStream_t st) {
        if (!hs) {
            const char* sn[] = {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11", nullptr};
            const char* bn[] = {"libcublas.so.12", "libcublas.so", "/usr/local/cuda/lib64/libcublas.so.12", nullptr};
            hb = open_any(bn);
            hs = open_any(sn);
            if (!hs || !hb) fail(BSFG_ECUDA, "bsfg_init_lists needs libcusolver / libcublas (dlopen failed: %s)", dlerror());
#define BSFG_DSYM(h, field, name) *(void**)(&field) = dlsym(h, name); if (!field) fail(BSFG_ECUDA, "symbol %s missing", name)
            BSFG_DSYM(hs, Create, "cusolverDnCreate"); BSFG_DSYM(hs, SetStream, "cusolverDnSetStream");
            BSFG_DSYM(hs, SyevdBuf, "cusolverDnDsyevd_bufferSize"); BSFG_DSYM(hs, Syevd, "cusolverDnDsyevd");
            BSFG_DSYM(hs, PotrfBuf, "cusolverDnDpotrf_bufferSize"); BSFG_DSYM(hs, Potrf, "cusolverDnDpotrf");
            BSFG_DSYM(hs, PotriBuf, "cusolverDnDpotri_bufferSize"); BSFG_DSYM(hs, Potri, "cusolverDnDpotri");
            BSFG_DSYM(hs, GetrfBuf, "cusolverDnDgetrf_bufferSize"); BSFG_DSYM(hs, Getrf, "cusolverDnDgetrf");
            BSFG_DSYM(hs, Getrs, "cusolverDnDgetrs");
            BSFG_DSYM(hs, GesvdBuf, "cusolverDnDgesvd_bufferSize"); BSFG_DSYM(hs, Gesvd, "cusolverDnDgesvd");
            BSFG_DSYM(hb, BCreate, "cublasCreate_v2"); BSFG_DSYM(hb, BSetStream, "cublasSetStream_v2");
            BSFG_DSYM(hb, Trsm, "cublasDtrsm_v2"); BSFG_DSYM(hb, Gemm, "cublasDgemm_v2");
#undef BSFG_DSYM
            if (Create(&solver) != CUSOLVER_STATUS_SUCCESS) fail(BSFG_ECUDA, "cusolverDnCreate failed");
            if (BCreate(&blas) != CUBLAS_STATUS_SUCCESS) fail(BSFG_ECUDA, "cublasCreate failed");
        }
        SetStream(solver, st);
        BSetStream(blas, st);
    }
};
static DenseApi g_dense;

// ---- small kernels --------------------------------------------------------------------------------
__global__ void init_fill_kernel(double* __restrict__ A, long ld, long rows, long cols, double off, double diag) {
    const long total = rows * cols;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx % rows, j = idx / rows;
        A[j * ld + i] = (i == j) ? diag : off;
    }
}
// A <- (A + A')/2 (uplo == 0) or mirror the lower triangle into the upper one (uplo == 1)
__global__ void init_symmetrize_kernel(double* __restrict__ A, long ld, long n, int mirror_lower) {
    const long total = n * n;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx % n, j = idx / n;
        if (i <= j) continue;                                   // (i > j): strictly lower element, owns the pair
        const double lo = A[j * ld + i], up = A[i * ld + j];
        const double v = mirror_lower ? lo : 0.5 * (lo + up);
        A[j * ld + i] = v; A[i * ld + j] = v;
    }
}
// dst[(r0 + i), (c0 + j)] = src[i, j]
__global__ void init_copy_block_kernel(const double* __restrict__ src, long lds, long rows, long cols, double* __restrict__ dst,
                                       long ldd, long r0, long c0, double scale) {
    const long total = rows * cols;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx % rows, j = idx / rows;
        dst[(c0 + j) * ldd + r0 + i] = scale * src[j * lds + i];
    }
}
__global__ void init_add_kernel(double* __restrict__ dst, long ldd, const double* __restrict__ src, long lds, long rows, long cols) {
    const long total = rows * cols;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx % rows, j = idx / rows;
        dst[j * ldd + i] += src[j * lds + i];
    }
}
// eigenvalues ascending -> descending (columns of V reversed), clipped to [lo, hi]; optional second vector 1 - w
__global__ void init_reverse_kernel(const double* __restrict__ w, const double* __restrict__ V, long ldv, long n, double lo, double hi,
                                    double* __restrict__ w_out, double* __restrict__ one_minus, double* __restrict__ V_out, long ldo) {
    const long total = n * n;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx % n, j = idx / n;
        V_out[j * ldo + i] = V[(n - 1 - j) * ldv + i];
        if (i == 0) {
            const double v = fmin(fmax(w[n - 1 - j], lo), hi);
            w_out[j] = v;
            if (one_minus) one_minus[j] = 1.0 - v;
        }
    }
}

struct InitWork {
    Device& d;
    DVec work; int* info = nullptr;
    explicit InitWork(Device& dev) : d(dev) { BSFG_CUDA(cudaMalloc(&info, sizeof(int))); }
    ~InitWork() { if (info) cudaFree(info); }
    int grid(long total) const { return (int)std::max<long>(1, std::min<long>((total + 255) / 256, (long)d.sms * 8)); }
    void check_info(const char* what) {
        int h = 0;
        BSFG_CUDA(cudaMemcpyAsync(&h, info, sizeof(int), cudaMemcpyDeviceToHost, d.st));
        d.sync();
        if (h != 0) fail(h > 0 ? BSFG_ENOTPD : BSFG_EARG, "%s failed (info = %d)", what, h);
    }
    void need(int lwork) { if ((long)work.n < lwork) work.alloc(lwork); }
    // C (m x n) = op(A) op(B)
    void gemm(bool ta, bool tb, int m, int n, int k, const DMat& A, const DMat& B, DMat& C, double alpha = 1.0, double beta = 0.0) {
        if (g_dense.Gemm(g_dense.blas, ta ? CUBLAS_OP_T : CUBLAS_OP_N, tb ? CUBLAS_OP_T : CUBLAS_OP_N, m, n, k, &alpha, A.p, (int)A.ld,
                         B.p, (int)B.ld, &beta, C.p, (int)C.ld) != CUBLAS_STATUS_SUCCESS)
            fail(BSFG_ECUDA, "cublasDgemm failed");
    }
    // symmetric eigen-decomposition in place: S -> eigenvectors (ascending eigenvalues in w)
    void eigh(DMat& S, DVec& w) {
        const int n = (int)S.rows;
        int lwork = 0;
        if (g_dense.SyevdBuf(g_dense.solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, S.p, (int)S.ld, w.p, &lwork) != CUSOLVER_STATUS_SUCCESS)
            fail(BSFG_ECUDA, "cusolverDnDsyevd_bufferSize failed");
        need(lwork);
        if (g_dense.Syevd(g_dense.solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, S.p, (int)S.ld, w.p, work.p, lwork, info) != CUSOLVER_STATUS_SUCCESS)
            fail(BSFG_ECUDA, "cusolverDnDsyevd failed");
        check_info("syevd");
    }
    void chol_lower(DMat& K) {
        const int n = (int)K.rows;
        int lwork = 0;
        if (g_dense.PotrfBuf(g_dense.solver, CUBLAS_FILL_MODE_LOWER, n, K.p, (int)K.ld, &lwork) != CUSOLVER_STATUS_SUCCESS)
            fail(BSFG_ECUDA, "cusolverDnDpotrf_bufferSize failed");
        need(lwork);
        if (g_dense.Potrf(g_dense.solver, CUBLAS_FILL_MODE_LOWER, n, K.p, (int)K.ld, work.p, lwork, info) != CUSOLVER_STATUS_SUCCESS)
            fail(BSFG_ECUDA, "cusolverDnDpotrf failed");
        check_info("potrf (matrix not positive definite?)");
    }
    // A <- inv(A), A SPD, full symmetric result
    void spd_inverse(DMat& A) {
        const int n = (int)A.rows;
        chol_lower(A);
        int lwork = 0;
        if (g_dense.PotriBuf(g_dense.solver, CUBLAS_FILL_MODE_LOWER, n, A.p, (int)A.ld, &lwork) != CUSOLVER_STATUS_SUCCESS)
            fail(BSFG_ECUDA, "cusolverDnDpotri_bufferSize failed");
        need(lwork);
        if (g_dense.Potri(g_dense.solver, CUBLAS_FILL_MODE_LOWER, n, A.p, (int)A.ld, work.p, lwork, info) != CUSOLVER_STATUS_SUCCESS)
            fail(BSFG_ECUDA, "cusolverDnDpotri failed");
        check_info("potri");
        init_symmetrize_kernel<<<grid((long)n * n), 256, 0, d.st>>>(A.p, A.ld, n, 1);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    // U with U'PU = diag(1 - lam), U'D2 U = diag(lam): L = chol(P + D2), M = L^-1 D2 L^-T = V diag(lam) V', U = L^-T V.
    // P is consumed (becomes L); D2 is consumed (becomes V).  Eigenvalues descending.
    void joint_diag(DMat& P, DMat& D2, DMat& U, DVec& s_P, DVec& s_D) {
        const int m = (int)P.rows;
        init_add_kernel<<<grid((long)m * m), 256, 0, d.st>>>(P.p, P.ld, D2.p, D2.ld, m, m);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        chol_lower(P);
        const double one = 1.0;
        // D2 <- L^-1 D2 ; D2 <- D2 L^-T
        if (g_dense.Trsm(g_dense.blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, m, m, &one, P.p, (int)P.ld,
                         D2.p, (int)D2.ld) != CUBLAS_STATUS_SUCCESS) fail(BSFG_ECUDA, "cublasDtrsm failed");
        if (g_dense.Trsm(g_dense.blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, m, m, &one, P.p, (int)P.ld,
                         D2.p, (int)D2.ld) != CUBLAS_STATUS_SUCCESS) fail(BSFG_ECUDA, "cublasDtrsm failed");
        init_symmetrize_kernel<<<grid((long)m * m), 256, 0, d.st>>>(D2.p, D2.ld, m, 0);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        DVec lam(m);
        eigh(D2, lam);
        init_reverse_kernel<<<grid((long)m * m), 256, 0, d.st>>>(lam.p, D2.p, D2.ld, m, 0.0, 1.0, s_D.p, s_P.p, U.p, U.ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        // U <- L^-T V
        if (g_dense.Trsm(g_dense.blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, m, m, &one, P.p, (int)P.ld,
                         U.p, (int)U.ld) != CUBLAS_STATUS_SUCCESS) fail(BSFG_ECUDA, "cublasDtrsm failed");
        d.sync();
    }
};

// X[, j] *= sqrt(1 + d_j^2);  C_j = d_j / sqrt(1 + d_j^2);  S_j = 1 / sqrt(1 + d_j^2)        cpp:33-37
__global__ void gsvd_finish_kernel(double* __restrict__ X, long ld, long n, const double* __restrict__ dsv, double* __restrict__ C,
                                   double* __restrict__ S) {
    const long total = n * n;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx % n, j = idx / n;
        const double nf = sqrt(1.0 + dsv[j] * dsv[j]);
        X[j * ld + i] *= nf;
        if (i == 0) { C[j] = dsv[j] / nf; S[j] = 1.0 / nf; }
    }
}

}  // namespace bsfg

// GSVD_2_c(A, B) (cpp:27-40), the literal route: svd(A inv(B)) = U diag(d) V', norm = sqrt(1 + d^2), C = d / norm, S = 1 / norm,
// X = (B'V) diag(norm), so that A = U C X' and B = V S X'.  A inv(B) comes from an LU solve (B'M' = A'), the SVD from cuSOLVER gesvd.
extern "C" int bsfg_GSVD_2_c(const double* A, const double* B, int n, double* U_out, double* V_out, double* X_out, double* C_out,
                             double* S_out) {
    BSFG_API_BEGIN
    REQUIRE_PTR(A); REQUIRE_PTR(B);
    if (n <= 0) fail(BSFG_ESHAPE, "GSVD_2_c: n must be positive");
    Device& d = default_device();
    g_dense.load(d.st);
    InitWork w(d);
    DMat dA(n, n), dB(n, n), LU(n, n), Mt(n, n), M(n, n), U(n, n), VT(n, n), V(n, n), X(n, n);
    DVec sv(n), Cd(n), Sd(n), rwork(n);
    dA.upload(A, d.st); dB.upload(B, d.st);
    BSFG_CUDA(cudaMemcpy2DAsync(LU.p, LU.ld * 8, dB.p, dB.ld * 8, (size_t)n * 8, n, cudaMemcpyDeviceToDevice, d.st));
    d.transpose(dA, Mt);                                                       // right-hand side A'
    int* ipiv = nullptr;
    BSFG_CUDA(cudaMalloc(&ipiv, sizeof(int) * (size_t)n));
    struct PivGuard { int* p; ~PivGuard() { cudaFree(p); } } guard{ipiv};
    int lwork = 0;
    if (g_dense.GetrfBuf(g_dense.solver, n, n, LU.p, (int)LU.ld, &lwork) != CUSOLVER_STATUS_SUCCESS) fail(BSFG_ECUDA, "cusolverDnDgetrf_bufferSize failed");
    w.need(lwork);
    if (g_dense.Getrf(g_dense.solver, n, n, LU.p, (int)LU.ld, w.work.p, ipiv, w.info) != CUSOLVER_STATUS_SUCCESS) fail(BSFG_ECUDA, "cusolverDnDgetrf failed");
    w.check_info("GSVD_2_c: inv(B) (B singular?)");
    if (g_dense.Getrs(g_dense.solver, CUBLAS_OP_T, n, n, LU.p, (int)LU.ld, ipiv, Mt.p, (int)Mt.ld, w.info) != CUSOLVER_STATUS_SUCCESS)
        fail(BSFG_ECUDA, "cusolverDnDgetrs failed");
    w.check_info("GSVD_2_c: getrs");
    d.transpose(Mt, M);                                                        // M = A inv(B)
    if (g_dense.GesvdBuf(g_dense.solver, n, n, &lwork) != CUSOLVER_STATUS_SUCCESS) fail(BSFG_ECUDA, "cusolverDnDgesvd_bufferSize failed");
    w.need(lwork);
    if (g_dense.Gesvd(g_dense.solver, 'A', 'A', n, n, M.p, (int)M.ld, sv.p, U.p, (int)U.ld, VT.p, (int)VT.ld, w.work.p, lwork, rwork.p, w.info)
        != CUSOLVER_STATUS_SUCCESS) fail(BSFG_ECUDA, "cusolverDnDgesvd failed");
    w.check_info("GSVD_2_c: svd");
    d.transpose(VT, V);
    d.gemm(1.0, dB, V, 0.0, X);                                                // B'V
    gsvd_finish_kernel<<<w.grid((long)n * n), 256, 0, d.st>>>(X.p, X.ld, n, sv.p, Cd.p, Sd.p);
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
    if (U_out) U.download(U_out, d.st);
    if (V_out) V.download(V_out, d.st);
    if (X_out) X.download(X_out, d.st);
    if (C_out) Cd.download(C_out, d.st);
    if (S_out) Sd.download(S_out, d.st);
    d.sync();
    BSFG_API_END
}

extern "C" int bsfg_init_lists(int n, int r, int b, int r2, const double* Z_1, const double* A, const double* X,
                               const double* Z_2, double b_X_prec, bsfg_init_out* out) {
    BSFG_API_BEGIN
    REQUIRE_PTR(A); REQUIRE_PTR(out);
    if (n <= 0 || r <= 0 || b < 0 || r2 < 0) fail(BSFG_ESHAPE, "bsfg_init_lists: bad sizes");
    if (!Z_1 && r != n) fail(BSFG_ESHAPE, "Z_1 omitted (identity) but r != n");
    if (b > 0) REQUIRE_PTR(X);
    if (r2 > 0) REQUIRE_PTR(Z_2);
    Device& d = default_device();
    g_dense.load(d.st);
    InitWork w(d);
    const int br = b + r;
    DMat dA(r, r), dZ1, Ainv(r, r);
    dA.upload(A, d.st);
    if (Z_1) { dZ1.alloc(n, r); dZ1.upload(Z_1, d.st); }
    // Ainv = solve(A)                                                    fast_BSFG_sampler_init.R:263
    BSFG_CUDA(cudaMemcpy2DAsync(Ainv.p, Ainv.ld * 8, dA.p, dA.ld * 8, (size_t)r * 8, r, cudaMemcpyDeviceToDevice, d.st));
    w.spd_inverse(Ainv);
    if (out->Ainv) Ainv.download(out->Ainv, d.st);

    // list 1: Z_1 A Z_1' = U diag(s) U'                                  init:278-282 (svd of a symmetric PSD matrix)
    {
        DMat S(n, n), U(n, n);
        DVec ev(n), s(n);
        if (Z_1) {
            DMat ZA(n, r);
            w.gemm(false, false, n, r, r, dZ1, dA, ZA);
            w.gemm(false, true, n, n, r, ZA, dZ1, S);
            d.sync();
        } else {
            BSFG_CUDA(cudaMemcpy2DAsync(S.p, S.ld * 8, dA.p, dA.ld * 8, (size_t)n * 8, n, cudaMemcpyDeviceToDevice, d.st));
        }
        init_symmetrize_kernel<<<w.grid((long)n * n), 256, 0, d.st>>>(S.p, S.ld, n, 0);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        w.eigh(S, ev);
        init_reverse_kernel<<<w.grid((long)n * n), 256, 0, d.st>>>(ev.p, S.p, S.ld, n, 0.0, INFINITY, s.p, nullptr, U.p, U.ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        if (out->U1) U.download(out->U1, d.st);
        if (out->s) s.download(out->s, d.st);
        d.sync();
    }
    // list 2: P = bdiag(b_X_prec I_b, Ainv), D = [X Z_1]                 init:285-295
    {
        DMat D(n, br), P(br, br), D2(br, br), U(br, br), DU(n, br);
        DVec s1(br), s2(br);
        if (b > 0) {
            DMat dX(n, b);
            dX.upload(X, d.st);
            init_copy_block_kernel<<<w.grid((long)n * b), 256, 0, d.st>>>(dX.p, dX.ld, n, b, D.p, D.ld, 0, 0, 1.0);
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
            d.sync();
        }
        if (Z_1) init_copy_block_kernel<<<w.grid((long)n * r), 256, 0, d.st>>>(dZ1.p, dZ1.ld, n, r, D.p, D.ld, 0, b, 1.0);
        else init_fill_kernel<<<w.grid((long)n * r), 256, 0, d.st>>>(D.p + (long)b * D.ld, D.ld, n, r, 0.0, 1.0);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        init_fill_kernel<<<w.grid((long)br * br), 256, 0, d.st>>>(P.p, P.ld, br, br, 0.0, b_X_prec);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        init_copy_block_kernel<<<w.grid((long)r * r), 256, 0, d.st>>>(Ainv.p, Ainv.ld, r, r, P.p, P.ld, b, b, 1.0);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        w.gemm(true, false, br, br, n, D, D, D2);
        w.joint_diag(P, D2, U, s1, s2);
        w.gemm(false, false, n, br, br, D, U, DU);
        if (out->U2) U.download(out->U2, d.st);
        if (out->s1_2) s1.download(out->s1_2, d.st);
        if (out->s2_2) s2.download(out->s2_2, d.st);
        if (out->Design_U) DU.download(out->Design_U, d.st);
        d.sync();
    }
    // list 3: a Z_1'Z_1 + b Ainv  (s1 <-> Z'Z, s2 <-> Ainv)               init:315-325
    {
        DMat P(r, r), ZZ(r, r), U(r, r);
        DVec sP(r), sD(r);
        BSFG_CUDA(cudaMemcpy2DAsync(P.p, P.ld * 8, Ainv.p, Ainv.ld * 8, (size_t)r * 8, r, cudaMemcpyDeviceToDevice, d.st));
        if (Z_1) w.gemm(true, false, r, r, n, dZ1, dZ1, ZZ);
        else init_fill_kernel<<<w.grid((long)r * r), 256, 0, d.st>>>(ZZ.p, ZZ.ld, r, r, 0.0, 1.0);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        w.joint_diag(P, ZZ, U, sP, sD);
        if (out->U3) U.download(out->U3, d.st);
        if (out->s1_3) sD.download(out->s1_3, d.st);
        if (out->s2_3) sP.download(out->s2_3, d.st);
        d.sync();
    }
    // list 4 (second random effect): a A_2_inv + b Z_2'Z_2, A_2_inv = I   init:264, 300-311
    if (r2 > 0) {
        DMat dZ2(n, r2), P(r2, r2), ZZ(r2, r2), U(r2, r2), DU(n, r2);
        DVec sP(r2), sD(r2);
        dZ2.upload(Z_2, d.st);
        init_fill_kernel<<<w.grid((long)r2 * r2), 256, 0, d.st>>>(P.p, P.ld, r2, r2, 0.0, 1.0);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        w.gemm(true, false, r2, r2, n, dZ2, dZ2, ZZ);
        w.joint_diag(P, ZZ, U, sP, sD);
        w.gemm(false, false, n, r2, r2, dZ2, U, DU);
        if (out->U4) U.download(out->U4, d.st);
        if (out->s1_4) sP.download(out->s1_4, d.st);
        if (out->s2_4) sD.download(out->s2_4, d.st);
        if (out->Design_U4) DU.download(out->Design_U4, d.st);
        d.sync();
    }
    d.sync();
    BSFG_API_END
}
