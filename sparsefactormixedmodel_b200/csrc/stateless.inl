// Stateless drop-ins: one C-ABI entry per Rcpp export (BSFG_functions_c.cpp) + sample_delta_c.
// Copy in -> kernels -> copy out.  Included by bsfg.cu.

namespace bsfg {

// launch helper: row splits chosen so that the grid is about eight waves of the 4-CTA/SM residency
struct MeansScratch { double* Bpart = nullptr; int* counter = nullptr; long cap_part = 0, cap_cnt = 0; };
inline void launch_means_theta(cudaStream_t st, int sms, MeansScratch& ms, const double* DtYt, long ldd, int br, int p, const double* s1,
                               const double* s2, const double* rp, const double* ep, const VarSrc& z, double* theta, long ldt,
                               double* thetat, long ldtt, const double* Ut, long ldu, int b, double* Bout, long ldb) {
    if (p <= 0 || br <= 0) return;
    const int tiles = ceil_div(p, MEANS_TC);
    int rs = std::max(1, std::min(8, ceil_div(sms * 4 * 8, tiles)));
    int rows_per_cta = (int)round_up(ceil_div(br, rs), 64);
    rs = ceil_div(br, rows_per_cta);
    if (Bout && rs > 1) {
        const long need_part = (long)rs * p * MEANS_BMAX;
        if (ms.cap_part < need_part) {
            if (ms.Bpart) cudaFree(ms.Bpart);
            BSFG_CUDA(cudaMalloc(&ms.Bpart, sizeof(double) * need_part));
            ms.cap_part = need_part;
        }
        if (ms.cap_cnt < tiles) {
            if (ms.counter) cudaFree(ms.counter);
            BSFG_CUDA(cudaMalloc(&ms.counter, sizeof(int) * tiles));
            BSFG_CUDA(cudaMemsetAsync(ms.counter, 0, sizeof(int) * tiles, st));
            ms.cap_cnt = tiles;
        }
    }
    dim3 grid(tiles, rs);
    means_theta_kernel<<<grid, MEANS_THREADS, 0, st>>>(DtYt, ldd, br, p, s1, s2, rp, ep, z, theta, ldt, thetat, ldtt, Ut, ldu, b, Bout, ldb,
                                                       rows_per_cta, ms.Bpart, ms.counter);
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
}

struct Device {
    cudaStream_t st = nullptr;
    int sms = 0;
    DVec ws;          // split-K workspace
    int* d_flag = nullptr;
    DVec chol_ws;         // tile workspace of the large-k Cholesky (k > BSFG_K_SMEM_LIMIT)
    MeansScratch means;   // row-split partial sums of the fused B reduction (launch_means_theta)
    std::map<int, uint16_t*> chol_maps;   // k -> packed-lower index -> tile offset (chol_blocked.cuh)
    static constexpr size_t WS_DOUBLES = (size_t)24 << 20;   // 192 MB of split-K partials
    void init() {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            fail(BSFG_ECUDA, "no CUDA device: libbsfg_b200 has no CPU fallback (%s)", cudaGetErrorString(e));
        int dev = 0;
        BSFG_CUDA(cudaGetDevice(&dev));
        cudaDeviceProp prop;
        BSFG_CUDA(cudaGetDeviceProperties(&prop, dev));
        if (prop.major != 10) fail(BSFG_ECUDA, "libbsfg_b200 is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
        sms = prop.multiProcessorCount;
        BSFG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        ws.alloc(WS_DOUBLES);
        BSFG_CUDA(cudaMalloc(&d_flag, sizeof(int)));
        memset_sync(d_flag, 0, sizeof(int));
    }
    void sync() { BSFG_CUDA(cudaStreamSynchronize(st)); }
    // C = alpha * At' * B + beta * C
    void gemm(int M, int N, int K, double alpha, const double* At, long lda, const double* B, long ldb,
              double beta, double* C, long ldc) {
        dgemm_tn(st, sms, M, N, K, alpha, At, lda, B, ldb, beta, C, ldc, C, ldc, ws.p, WS_DOUBLES);
    }
    void gemm(double alpha, const DMat& At, const DMat& B, double beta, DMat& C) {
        if (At.rows != B.rows || C.rows != At.cols || C.cols != B.cols)
            fail(BSFG_ESHAPE, "gemm shape mismatch: At %ldx%ld B %ldx%ld C %ldx%ld", At.rows, At.cols, B.rows, B.cols, C.rows, C.cols);
        gemm((int)At.cols, (int)B.cols, (int)At.rows, alpha, At.p, At.ld, B.p, B.ld, beta, C.p, C.ld);
    }
    void transpose(const DMat& in, DMat& out) {
        if (out.rows != in.cols || out.cols != in.rows) fail(BSFG_ESHAPE, "transpose shape mismatch");
        if (in.rows == 0 || in.cols == 0) return;
        dim3 grid(ceil_div(in.rows, 32), ceil_div(in.cols, 32)), block(32, 8);
        transpose_kernel<<<grid, block, 0, st>>>(in.p, in.ld, in.rows, in.cols, out.p, out.ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    void scale(const DMat& in, const double* rowscale, const double* colscale, DMat& out) {
        const long total = in.rows * in.cols;
        if (!total) return;
        int blocks = (int)std::min<long>((total + 255) / 256, (long)sms * 16);
        scale_kernel<<<blocks, 256, 0, st>>>(in.p, in.ld, in.rows, in.cols, rowscale, colscale, out.p, out.ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    // table for chol_blocked_kernel: packed-lower pair(a,b) = a(a+1)/2 + b  ->  offset of element (a,b) in the tiled layout
    const uint16_t* chol_map(int k) {
        auto it = chol_maps.find(k);
        if (it != chol_maps.end()) return it->second;
        std::vector<uint16_t> h((size_t)k * (k + 1) / 2);
        for (int a = 0; a < k; a++)
            for (int b = 0; b <= a; b++)
                h[(size_t)a * (a + 1) / 2 + b] = (uint16_t)(cb_tile_id(a >> 3, b >> 3) * 64 + cb_swz(a & 7, b & 7));
        uint16_t* dptr = nullptr;
        BSFG_CUDA(cudaMalloc(&dptr, sizeof(uint16_t) * h.size()));
        BSFG_CUDA(cudaMemcpyAsync(dptr, h.data(), sizeof(uint16_t) * h.size(), cudaMemcpyHostToDevice, st));
        sync();
        chol_maps[k] = dptr;
        return dptr;
    }
    int read_flag_and_clear() {
        int h = 0;
        BSFG_CUDA(cudaMemcpyAsync(&h, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        sync();
        if (h) BSFG_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
        return h;
    }
};

static Device& default_device() {
    static thread_local Device* dev = nullptr;
    if (!dev) {
        Device* d = new Device();
        d->init();
        dev = d;
    }
    return *dev;
}

static VarSrc make_var(const DMat* injected, uint64_t seed, uint32_t site, uint32_t iter, long col_offset, long rows_total) {
    VarSrc v;
    v.ptr = injected ? injected->p : nullptr;
    v.ld = injected ? injected->ld : 0;
    v.rs.seed = seed; v.rs.site = site; v.rs.iter = iter;
    v.col_offset = col_offset; v.rows_total = rows_total;
    return v;
}

// ---- reusable device-level pieces (also used by the sweep) -----------------------------------
// Up to k = 216 a system's tiles fit 227 KB of shared memory (28 block rows incl. the right-hand-side row); above that the
// batched Cholesky keeps them in a global (L2) workspace (chol_blocked_big_kernel) and the F row solves read the factor from
// L2.  The cap itself is the 16 x 32 columns a warp of factor_rows_kernel holds in registers.
constexpr int BSFG_K_SMEM_LIMIT = 216;
constexpr int BSFG_K_LIMIT = 504;
static int kt_for(int k) {
    if (k > BSFG_K_LIMIT) fail(BSFG_ESHAPE, "k = %d exceeds the supported maximum of %d factors", k, BSFG_K_LIMIT);
    if (k <= 32) return 1;
    if (k <= 64) return 2;
    if (k <= 128) return 4;
    if (k <= 256) return 8;
    return 16;
}

// batched Cholesky + solves (chol_blocked.cuh): 4 warps per system up to 13 block rows (k + 1 <= 104), 8 warps above
static void launch_chol_solve(Device& d, const double* Qp, long ldq, int k, int nsys, const double* dadd,
                              long dadd_sys_stride, long dadd_elem_stride, const double* rhs, long ldr,
                              const VarSrc& z, double* out, long ldo, double* Lout) {
    if (nsys <= 0) return;
    kt_for(k);
    const size_t smem = cb_smem_bytes(k, rhs != nullptr);
    if (smem > 227 * 1024) {      // large k: tiles in a global workspace, persistent grid (chol_blocked.cuh, BIG variant)
        const int grid = std::min(nsys, d.sms * 2);
        const long stride = cb_big_ws_doubles(k, rhs != nullptr);
        if ((long)d.chol_ws.n < stride * grid) d.chol_ws.alloc(stride * grid);
        const size_t smem_big = cb_big_smem_bytes(k, rhs != nullptr);
        BSFG_CUDA(cudaFuncSetAttribute(chol_blocked_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big));
        chol_blocked_big_kernel<<<grid, 256, smem_big, d.st>>>(Qp, ldq, k, nsys, dadd, dadd_sys_stride, dadd_elem_stride, rhs, ldr, z, out,
                                                              ldo, Lout, d.d_flag, d.chol_ws.p, stride);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        return;
    }
    const uint16_t* map = d.chol_map(k);
    if (cb_warps(k, rhs != nullptr) == 4) {
        BSFG_CUDA(cudaFuncSetAttribute(chol_blocked_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chol_blocked_kernel<4><<<nsys, 128, smem, d.st>>>(Qp, ldq, k, nsys, map, dadd, dadd_sys_stride, dadd_elem_stride,
                                                          rhs, ldr, z, out, ldo, Lout, d.d_flag);
    } else {
        BSFG_CUDA(cudaFuncSetAttribute(chol_blocked_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chol_blocked_kernel<8><<<nsys, 256, smem, d.st>>>(Qp, ldq, k, nsys, map, dadd, dadd_sys_stride, dadd_elem_stride,
                                                          rhs, ldr, z, out, ldo, Lout, d.d_flag);
    }
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
}

static void launch_factor_rows(Device& d, const double* Lp, int k, int n, const double* YL, long ldy,
                               const double* ZFa, long ldz, const double* tau_e, const VarSrc& z, double* F, long ldf,
                               int row0 = 0 /* data pointers are offset to row0; z is addressed with row0 + i */) {
    if (n <= 0) return;
    // the packed factor is staged in shared memory while it fits (k <= 226); larger factors are read from global memory (L2)
    const size_t lbytes = sizeof(double) * ((size_t)k * (k + 1) / 2);
    const bool l_in_smem = lbytes <= 200 * 1024;
    const size_t smem = l_in_smem ? lbytes : 0;
    const int wpb = 8;
    int blocks = std::min(ceil_div(n, wpb), d.sms * 4);
#define BSFG_LAUNCH_FR(KT)                                                                                     \
    do {                                                                                                       \
        BSFG_CUDA(cudaFuncSetAttribute(factor_rows_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        factor_rows_kernel<KT><<<blocks, wpb * 32, smem, d.st>>>(Lp, k, n, YL, ldy, ZFa, ldz, tau_e, z, row0, F, ldf, l_in_smem ? 1 : 0); \
    } while (0)
    switch (kt_for(k)) {
        case 1: BSFG_LAUNCH_FR(1); break;
        case 2: BSFG_LAUNCH_FR(2); break;
        case 4: BSFG_LAUNCH_FR(4); break;
        case 8: BSFG_LAUNCH_FR(8); break;
        default: BSFG_LAUNCH_FR(16); break;
    }
#undef BSFG_LAUNCH_FR
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
}

// Lambda draw given rotated inputs.  UtF n x k, UtY n x p (rotated data), optional UtX (n x b), B (b x p).
// Writes Lt (k x p).  Scratch: Gt (n x npairs), W, WY (n x p), Q (npairs x p), Mn (k x p) and, for the
// exponential-sum Gram (kernels.cuh), A (n x R), Ht (R x npairs), E (R x p), the device plan and the fallback flag.
enum GramMode { GRAM_DEFAULT = 0, GRAM_EXACT = 1, GRAM_EXPSUM = 2 };
struct GramOpts {
    int mode = GRAM_EXPSUM;      // resolved (never GRAM_DEFAULT)
    int terms = 256;             // R
    double step = 0.28;          // h: relative quadrature error ~1.2e-14
    double s_min = 0.0, s_max = 0.0;
    // stateful sweeps created with gram_terms == 0: R is re-chosen at the start of every bsfg_sweep call from the range
    // of c seen at the end of the previous one (widened 4x on both sides), between 64 and terms_max; the device guard
    // still falls back to the exact GEMM if an iteration needs more
    bool auto_terms = false;
    int terms_max = 256;
};
// nodes needed to cover [s_min + c_min, s_max + c_max] with step h (same formulas as expsum_plan_kernel)
static int expsum_terms_needed(double s_min, double s_max, double c_min, double c_max, double h) {
    const double eps = 1e-16;
    const double x_lo = s_min + c_min, x_hi = s_max + c_max;
    if (!(x_lo > 0.0) || !std::isfinite(x_hi)) return 1 << 30;
    const double t0 = std::log(eps / x_hi), need = std::log(-std::log(eps) / x_lo);
    return (int)std::ceil((need - t0) / h) + 1;
}
static GramOpts g_gram_default;
static GramOpts resolve_gram(int mode, int terms, double step) {
    GramOpts o = g_gram_default;
    if (mode != GRAM_DEFAULT) o.mode = mode;
    if (terms > 0) o.terms = terms;
    if (step > 0.0) o.step = step;
    if (o.mode != GRAM_EXACT && o.mode != GRAM_EXPSUM) fail(BSFG_EARG, "gram_mode must be 0 (default), 1 (exact) or 2 (expsum)");
    if (o.terms < 16 || o.terms > 4096 || o.terms % 16 != 0) fail(BSFG_EARG, "gram_terms must be a multiple of 16 in [16, 4096]");
    if (!(o.step >= 0.05 && o.step <= 0.5)) fail(BSFG_EARG, "gram_step must lie in [0.05, 0.5]");
    return o;
}
static void host_minmax(const double* s, int n, double* lo, double* hi) {
    double a = s[0], b = s[0];
    for (int i = 1; i < n; i++) { a = std::min(a, s[i]); b = std::max(b, s[i]); }
    *lo = a; *hi = b;
}
struct LambdaScratch {
    DMat Gt, W, WY, Q, Mn, A, Ht, E;
    ExpSumPlan* plan = nullptr;
    int* need_exact = nullptr;
    unsigned long long* fallbacks = nullptr;
    cudaEvent_t ev_sub[2] = {nullptr, nullptr};   // after the node product G'A, after the trait product (before the no-op fallback)
    bool sub_timed = false;
    // trait-sharded runs: the node product G'A does not depend on the traits, so its columns (factor pairs) are split
    // over the ranks and all-gathered instead of being recomputed by every rank
    int rank = 0, world = 1;
    std::function<void(double*, size_t)> allgather;   // in place: this rank's chunk sits at buf + rank * chunk_doubles
    std::function<void(double*, size_t)> allreduce_max;
    double* range = nullptr;                          // [-c_min, c_max, bad] (expsum_range_kernel)
    ~LambdaScratch() {
        for (auto& e : ev_sub) if (e) cudaEventDestroy(e);
        if (plan) cudaFree(plan);
        if (need_exact) cudaFree(need_exact);
        if (range) cudaFree(range);
        if (fallbacks) cudaFree(fallbacks);
    }
    void ensure(long n, long p, long k, long R) {
        const long np = k * (k + 1) / 2;
        if (Gt.rows != n || Gt.cols != np) Gt.alloc(n, np);
        if (W.rows != n || W.cols != p) { W.alloc(n, p); WY.alloc(n, p); }
        if (Q.rows != np || Q.cols != p) Q.alloc(np, p);
        if (Mn.rows != k || Mn.cols != p) Mn.alloc(k, p);
        if (R > 0) {
            if (A.rows != n || A.cols != R) A.alloc(n, R);
            const long np_pad = (long)ceil_div(np, world) * world;
            if (Ht.rows != R || Ht.cols != np_pad) Ht.alloc(R, np_pad);
            if (E.rows != R || E.cols != p) E.alloc(R, p);
        }
        if (!plan) {
            BSFG_CUDA(cudaMalloc(&plan, sizeof(ExpSumPlan)));
            memset_sync(plan, 0, sizeof(ExpSumPlan));
            BSFG_CUDA(cudaMalloc(&need_exact, sizeof(int)));
            memset_sync(need_exact, 0, sizeof(int));
            BSFG_CUDA(cudaMalloc(&range, 4 * sizeof(double)));
            memset_sync(range, 0, 4 * sizeof(double));
            BSFG_CUDA(cudaMalloc(&fallbacks, sizeof(unsigned long long)));
            memset_sync(fallbacks, 0, sizeof(unsigned long long));
            for (auto& e : ev_sub) BSFG_CUDA(cudaEventCreate(&e));
        }
    }
};

static void lambda_core(Device& d, LambdaScratch& sc, const GramOpts& go, const DMat& UtF, const DMat& UtY, const double* s,
                        const double* rp, const double* ep, const DMat* UtX, const DMat* B,
                        const double* plam, long plam_sys_stride, long plam_elem_stride, const VarSrc& z,
                        DMat& Lt, cudaEvent_t* ev = nullptr /* 3 marks: before the Gram product, after it, after chol */) {
    const int n = (int)UtF.rows, k = (int)UtF.cols, p = (int)UtY.cols;
    const int np = k * (k + 1) / 2;
    const bool expsum = go.mode == GRAM_EXPSUM;
    const int R = expsum ? go.terms : 0;
    sc.ensure(n, p, k, R);
    if (expsum) {
        expsum_range_kernel<<<1, 1024, 0, d.st>>>(ep, rp, p, sc.range);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        if (sc.world > 1 && sc.allreduce_max) sc.allreduce_max(sc.range, 3);   // one node grid for all ranks
        expsum_plan_kernel<<<1, 32, 0, d.st>>>(sc.range, go.s_min, go.s_max, go.step, R, 0, sc.plan, sc.need_exact, sc.fallbacks);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    {
        // Gram features; a rank of a sharded expsum run builds only its own block of pairs unless the fallback needs all
        int pair_lo = 0, pair_hi = np;
        const bool shard_pairs = expsum && sc.world > 1 && sc.allgather;
        if (shard_pairs) {
            const long chunk = ceil_div(np, sc.world);
            pair_lo = (int)std::min<long>(np, (long)sc.rank * chunk);
            pair_hi = (int)std::min<long>(np, pair_lo + chunk);
        }
        dim3 grid(np, std::min(ceil_div(n, 1024), 16));        // four rows per thread: a quarter of the CTAs of a one-row-per-thread grid
        gram_features_kernel<<<grid, 256, 0, d.st>>>(UtF.p, UtF.ld, n, k, sc.Gt.p, sc.Gt.ld, pair_lo, pair_hi,
                                                     shard_pairs ? sc.need_exact : nullptr);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    {
        const int b = UtX ? (int)UtX->cols : 0;
        lambda_weights_kernel<<<p, 256, 0, d.st>>>(UtY.p, UtY.ld, n, p, s, rp, ep, UtX ? UtX->p : nullptr,
                                                      UtX ? UtX->ld : 0, B ? B->p : nullptr, B ? B->ld : 0, b,
                                                      sc.W.p, sc.W.ld, expsum ? sc.need_exact : nullptr, sc.WY.p, sc.WY.ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    if (expsum) {
        dim3 grid(std::min(ceil_div(n, 256), 64), R);
        expsum_nodes_kernel<<<grid, 256, 0, d.st>>>(sc.plan, s, n, sc.A.p, sc.A.ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        expsum_traits_kernel<<<p, 256, 0, d.st>>>(sc.plan, ep, rp, p, sc.E.p, sc.E.ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    if (ev) BSFG_CUDA(cudaEventRecord(ev[0], d.st));
    if (expsum) {
        // Ht[m, pair] = sum_i A[i,m] f_ia f_ib
        if (sc.world > 1 && sc.allgather) {
            const long chunk = ceil_div(np, sc.world), lo = (long)sc.rank * chunk;
            const long cnt = std::max(0L, std::min(chunk, (long)np - lo));
            if (cnt > 0)
                dgemm_tn(d.st, d.sms, R, (int)cnt, n, 1.0, sc.A.p, sc.A.ld, sc.Gt.p + lo * sc.Gt.ld, sc.Gt.ld, 0.0, nullptr, 0,
                         sc.Ht.p + lo * sc.Ht.ld, sc.Ht.ld, d.ws.p, Device::WS_DOUBLES);
            sc.allgather(sc.Ht.p, (size_t)chunk * (size_t)sc.Ht.ld);
        } else {
            dgemm_tn(d.st, d.sms, R, np, n, 1.0, sc.A.p, sc.A.ld, sc.Gt.p, sc.Gt.ld, 0.0, nullptr, 0, sc.Ht.p, sc.Ht.ld,
                     d.ws.p, Device::WS_DOUBLES);
        }
        if (ev) BSFG_CUDA(cudaEventRecord(sc.ev_sub[0], d.st));
        // Q[pair, j]  = sum_m Ht[m,pair] E[m,j]
        dgemm_tn(d.st, d.sms, np, p, R, 1.0, sc.Ht.p, sc.Ht.ld, sc.E.p, sc.E.ld, 0.0, nullptr, 0, sc.Q.p, sc.Q.ld,
                 d.ws.p, Device::WS_DOUBLES);
        if (ev) { BSFG_CUDA(cudaEventRecord(sc.ev_sub[1], d.st)); sc.sub_timed = true; }
        // exact fallback: a no-op launch unless the plan kernel raised need_exact
        dgemm_tn(d.st, d.sms, np, p, n, 1.0, sc.Gt.p, sc.Gt.ld, sc.W.p, sc.W.ld, 0.0, nullptr, 0, sc.Q.p, sc.Q.ld,
                 nullptr, 0, sc.need_exact);
    } else {
        d.gemm(1.0, sc.Gt, sc.W, 0.0, sc.Q);       // Q[pair, j] = sum_i f_ia f_ib w_ij
    }
    if (ev) BSFG_CUDA(cudaEventRecord(ev[1], d.st));
    d.gemm(1.0, UtF, sc.WY, 0.0, sc.Mn);           // m[a, j]    = sum_i f_ia w_ij y_ij
    launch_chol_solve(d, sc.Q.p, sc.Q.ld, k, p, plam, plam_sys_stride, plam_elem_stride, sc.Mn.p, sc.Mn.ld, z,
                      Lt.p, Lt.ld, nullptr);
    if (ev) BSFG_CUDA(cudaEventRecord(ev[2], d.st));
}

static void check_notpd(Device& d, const char* what) {
    int f = d.read_flag_and_clear();
    if (f) fail(BSFG_ENOTPD, "chol(): decomposition failed in %s (system %d)", what, f - 1);
}

}  // namespace bsfg

using namespace bsfg;

#define BSFG_API_BEGIN try {
#define BSFG_API_END                                   \
    return BSFG_OK;                                    \
    }                                                  \
    catch (const bsfg::Error& e) {                     \
        bsfg::set_last_error("%s", e.what());          \
        return e.code;                                 \
    }                                                  \
    catch (const std::exception& e) {                  \
        bsfg::set_last_error("%s", e.what());          \
        return BSFG_ECUDA;                             \
    }

#define REQUIRE_PTR(p) do { if (!(p)) bsfg::fail(BSFG_EARG, "argument %s is NULL", #p); } while (0)

extern "C" int bsfg_set_gram_default(int mode, int terms, double step) {
    BSFG_API_BEGIN
    if (mode == GRAM_DEFAULT) mode = GRAM_EXPSUM;
    GramOpts saved = g_gram_default;
    g_gram_default = GramOpts();
    try { g_gram_default = resolve_gram(mode, terms, step); }
    catch (...) { g_gram_default = saved; throw; }
    BSFG_API_END
}

extern "C" int bsfg_dgemm_tn_host(int M, int N, int K, double alpha, const double* At, int lda, const double* B,
                                  int ldb, double beta, double* C, int ldc) {
    BSFG_API_BEGIN
    REQUIRE_PTR(At); REQUIRE_PTR(B); REQUIRE_PTR(C);
    if (lda < K || ldb < K || ldc < M) fail(BSFG_ESHAPE, "bsfg_dgemm_tn_host: leading dimension too small");
    Device& d = default_device();
    DMat dA(K, M), dB(K, N), dC(M, N);
    BSFG_CUDA(cudaMemcpy2DAsync(dA.p, dA.ld * 8, At, (size_t)lda * 8, (size_t)K * 8, M, cudaMemcpyHostToDevice, d.st));
    BSFG_CUDA(cudaMemcpy2DAsync(dB.p, dB.ld * 8, B, (size_t)ldb * 8, (size_t)K * 8, N, cudaMemcpyHostToDevice, d.st));
    if (beta != 0.0)
        BSFG_CUDA(cudaMemcpy2DAsync(dC.p, dC.ld * 8, C, (size_t)ldc * 8, (size_t)M * 8, N, cudaMemcpyHostToDevice, d.st));
    d.gemm(alpha, dA, dB, beta, dC);
    BSFG_CUDA(cudaMemcpy2DAsync(C, (size_t)ldc * 8, dC.p, dC.ld * 8, (size_t)M * 8, N, cudaMemcpyDeviceToHost, d.st));
    d.sync();
    BSFG_API_END
}

// ---- sample_Lambda_c (cpp:86-135) -------------------------------------------------------------
extern "C" int bsfg_sample_Lambda_c(const double* Y_tilde, int n, int p, const double* F, int k,
                                    const double* resid_Y_prec, const double* E_a_prec, const double* Plam,
                                    const double* U, const double* s, const double* z, unsigned long long seed,
                                    double* Lambda_out) {
    BSFG_API_BEGIN
    REQUIRE_PTR(Y_tilde); REQUIRE_PTR(F); REQUIRE_PTR(resid_Y_prec); REQUIRE_PTR(E_a_prec); REQUIRE_PTR(Plam);
    REQUIRE_PTR(U); REQUIRE_PTR(s); REQUIRE_PTR(Lambda_out);
    if (n <= 0 || p <= 0 || k <= 0) fail(BSFG_ESHAPE, "sample_Lambda_c: n, p, k must be positive");
    Device& d = default_device();
    DMat dY(n, p), dF(n, k), dU(n, n), dPlam(p, k), dZ, UtF(n, k), UtY(n, p), Lt(k, p), Lam(p, k);
    DVec ds(n), drp(p), dep(p);
    dY.upload(Y_tilde, d.st); dF.upload(F, d.st); dU.upload(U, d.st); dPlam.upload(Plam, d.st);
    ds.upload(s, d.st); drp.upload(resid_Y_prec, d.st); dep.upload(E_a_prec, d.st);
    if (z) { dZ.alloc(k, p); dZ.upload(z, d.st); }
    d.gemm(1.0, dU, dF, 0.0, UtF);       // U'F  = t(FtU), cpp:107
    d.gemm(1.0, dU, dY, 0.0, UtY);       // U'Y_tilde, cpp:108
    LambdaScratch sc;
    GramOpts go = resolve_gram(GRAM_DEFAULT, 0, 0.0);
    host_minmax(s, n, &go.s_min, &go.s_max);
    VarSrc zs = make_var(z ? &dZ : nullptr, seed, SITE_Z_LAMBDA, 0, 0, k);
    lambda_core(d, sc, go, UtF, UtY, ds.p, drp.p, dep.p, nullptr, nullptr, dPlam.p, 1, dPlam.ld, zs, Lt);
    d.transpose(Lt, Lam);
    Lam.download(Lambda_out, d.st);
    check_notpd(d, "sample_Lambda_c");
    BSFG_API_END
}
extern "C" int bsfg_sample_lambda_c(const double* Y_tilde, int n, int p, const double* F, int k,
                                    const double* resid_Y_prec, const double* E_a_prec, const double* Plam,
                                    const double* U, const double* s, const double* z, unsigned long long seed,
                                    double* Lambda_out) {
    return bsfg_sample_Lambda_c(Y_tilde, n, p, F, k, resid_Y_prec, E_a_prec, Plam, U, s, z, seed, Lambda_out);
}

// ---- sample_means_c (cpp:43-82) ---------------------------------------------------------------
extern "C" int bsfg_sample_means_c(const double* Y_tilde, int n, int p, const double* resid_Y_prec,
                                   const double* E_a_prec, const double* U, const double* s1, const double* s2,
                                   const double* Design_U, int br, const double* z, unsigned long long seed,
                                   double* location_sample_out) {
    BSFG_API_BEGIN
    REQUIRE_PTR(Y_tilde); REQUIRE_PTR(resid_Y_prec); REQUIRE_PTR(E_a_prec); REQUIRE_PTR(U); REQUIRE_PTR(s1);
    REQUIRE_PTR(s2); REQUIRE_PTR(Design_U); REQUIRE_PTR(location_sample_out);
    if (n <= 0 || p <= 0 || br <= 0) fail(BSFG_ESHAPE, "sample_means_c: n, p, br must be positive");
    Device& d = default_device();
    DMat dY(n, p), dU(br, br), dUt(br, br), dD(n, br), dZ, DtY(br, p), theta(br, p), out(br, p);
    DVec ds1(br), ds2(br), drp(p), dep(p);
    dY.upload(Y_tilde, d.st); dU.upload(U, d.st); dD.upload(Design_U, d.st);
    ds1.upload(s1, d.st); ds2.upload(s2, d.st); drp.upload(resid_Y_prec, d.st); dep.upload(E_a_prec, d.st);
    if (z) { dZ.alloc(br, p); dZ.upload(z, d.st); }
    d.gemm(1.0, dD, dY, 0.0, DtY);                       // Design_U' Y_tilde, cpp:65
    VarSrc zs = make_var(z ? &dZ : nullptr, seed, SITE_Z_MEANS, 0, 0, round_up(br, 2));   // even: a row pair = one Philox pair
    {
        launch_means_theta(d.st, d.sms, d.means, DtY.p, DtY.ld, br, p, ds1.p, ds2.p, drp.p, dep.p, zs, theta.p, theta.ld, nullptr, 0,
                           nullptr, 0, 0, nullptr, 0);
    }
    d.transpose(dU, dUt);
    d.gemm(1.0, dUt, theta, 0.0, out);                   // U * (mlam + z/sqrt(d)), cpp:78
    out.download(location_sample_out, d.st);
    d.sync();
    BSFG_API_END
}

// ---- sample_factors_scores_c / px (cpp:138-193) -----------------------------------------------
static void factor_scores_impl(const double* Y_tilde, int n, int p, const double* Z_1, int r, const double* Lambda,
                               int k, const double* resid_Y_prec, const double* F_a, const double* F_h2,
                               const double* tau_e_host, const double* z, unsigned long long seed, double* F_out) {
    REQUIRE_PTR(Y_tilde); REQUIRE_PTR(Lambda); REQUIRE_PTR(resid_Y_prec); REQUIRE_PTR(F_a); REQUIRE_PTR(F_out);
    if (!F_h2 && !tau_e_host) fail(BSFG_EARG, "F_h2 / tau_e is NULL");
    if (n <= 0 || p <= 0 || k <= 0 || r < 0) fail(BSFG_ESHAPE, "sample_factors_scores_c: bad sizes");
    Device& d = default_device();
    DMat dY(n, p), dYt(p, n), dL(p, k), Lmsg(p, k), dFa(std::max(r, 1), k), LtL(k, k), YL(n, k), ZFa, dZ, Fo(n, k);
    DVec drp(p), dtau(k), dh2(k), Lp((long)k * (k + 1) / 2);
    dY.upload(Y_tilde, d.st); dL.upload(Lambda, d.st); drp.upload(resid_Y_prec, d.st);
    if (z) { dZ.alloc(n, k); dZ.upload(z, d.st); }
    if (tau_e_host) dtau.upload(tau_e_host, d.st);
    else {
        dh2.upload(F_h2, d.st);
        tau_e_kernel<<<ceil_div(k, 128), 128, 0, d.st>>>(dh2.p, k, dtau.p);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    d.scale(dL, drp.p, nullptr, Lmsg);                   // Lmsg = Lambda * resid_Y_prec (rows), cpp:146
    d.gemm(1.0, dL, Lmsg, 0.0, LtL);                     // Lambda' Lmsg, cpp:149
    d.transpose(dY, dYt);
    d.gemm(1.0, dYt, Lmsg, 0.0, YL);                     // Y_tilde * Lmsg, cpp:152
    if (r > 0) {
        REQUIRE_PTR(Z_1);
        DMat dZ1(n, r), dZ1t(r, n);
        dZ1.upload(Z_1, d.st);
        dFa.upload(F_a, d.st);
        ZFa.alloc(n, k);
        d.transpose(dZ1, dZ1t);
        d.gemm(1.0, dZ1t, dFa, 0.0, ZFa);                // Z_1 * F_a, cpp:152
        d.sync();                                        // dZ1/dZ1t go out of scope
    }
    pack_lower_kernel<<<ceil_div((long)k * (k + 1) / 2, 256), 256, 0, d.st>>>(LtL.p, LtL.ld, k, Lp.p);
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
    DVec Lfac((long)k * (k + 1) / 2);
    VarSrc none = make_var(nullptr, 0, 0, 0, 0, 1);
    launch_chol_solve(d, Lp.p, 0, k, 1, dtau.p, 0, 1, nullptr, 0, none, nullptr, 0, Lfac.p);   // chol(.. + diag(tau_e))
    VarSrc zs = make_var(z ? &dZ : nullptr, seed, SITE_Z_F, 0, 0, n);
    launch_factor_rows(d, Lfac.p, k, n, YL.p, YL.ld, r > 0 ? ZFa.p : nullptr, r > 0 ? ZFa.ld : 0, dtau.p, zs, Fo.p, Fo.ld);
    Fo.download(F_out, d.st);
    check_notpd(d, "sample_factors_scores_c");
}

extern "C" int bsfg_sample_factors_scores_c(const double* Y_tilde, int n, int p, const double* Z_1, int r,
                                            const double* Lambda, int k, const double* resid_Y_prec,
                                            const double* F_a, const double* F_h2, const double* z,
                                            unsigned long long seed, double* F_out) {
    BSFG_API_BEGIN
    REQUIRE_PTR(F_h2);
    factor_scores_impl(Y_tilde, n, p, Z_1, r, Lambda, k, resid_Y_prec, F_a, F_h2, nullptr, z, seed, F_out);
    BSFG_API_END
}
extern "C" int bsfg_sample_px_factors_scores_c(const double* Y_tilde, int n, int p, const double* Z_1, int r,
                                               const double* Lambda_px, int k, const double* resid_Y_prec,
                                               const double* F_a_px, const double* tau_e, const double* px_factor,
                                               const double* z, unsigned long long seed, double* F_px_out) {
    BSFG_API_BEGIN
    (void)px_factor;   // unused in the reference too (cpp:172-176)
    REQUIRE_PTR(tau_e);
    factor_scores_impl(Y_tilde, n, p, Z_1, r, Lambda_px, k, resid_Y_prec, F_a_px, nullptr, tau_e, z, seed, F_px_out);
    BSFG_API_END
}

// ---- sample_h2s_discrete_c (cpp:196-238) ------------------------------------------------------
namespace bsfg {
static void h2s_core(Device& d, const DMat& UtF, const DMat& F, const double* s, int H, const double* prior,
                     const VarSrc& u, double* F_h2, double* log_ps /* k x H ld k or null */, DVec& det, DVec& ll) {
    const int n = (int)F.rows, k = (int)F.cols;
    h2_det_kernel<<<H, 256, 0, d.st>>>(s, n, H, det.p);
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
    {
        const long warps = (long)k * H;
        h2_loglik_kernel<<<ceil_div(warps * 32, 256), 256, 0, d.st>>>(UtF.p, UtF.ld, F.p, F.ld, s, n, k, H, ll.p);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    {
        const int wpb = 4;
        grid_draw_kernel<<<ceil_div(k, wpb), wpb * 32, sizeof(double) * wpb * H, d.st>>>(
            ll.p, k, det.p, 0, 0, prior, k, H, u, F_h2, log_ps, k);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
}
}  // namespace bsfg

extern "C" int bsfg_sample_h2s_discrete_c(const double* F, int n, int k, int h2_divisions, const double* h2_priors,
                                          const double* U, const double* s, const double* u,
                                          unsigned long long seed, double* F_h2_out, double* log_ps_out) {
    BSFG_API_BEGIN
    REQUIRE_PTR(F); REQUIRE_PTR(h2_priors); REQUIRE_PTR(U); REQUIRE_PTR(s); REQUIRE_PTR(F_h2_out);
    const int H = h2_divisions;
    if (n <= 0 || k <= 0 || H <= 0) fail(BSFG_ESHAPE, "sample_h2s_discrete_c: bad sizes");
    Device& d = default_device();
    DMat dF(n, k), dU(n, n), UtF(n, k), dUin;
    DVec ds(n), dprior(H), det(H), dh2(k), dlp((long)k * H), ll((long)k * H);
    dF.upload(F, d.st); dU.upload(U, d.st); ds.upload(s, d.st); dprior.upload(h2_priors, d.st);
    if (u) { dUin.alloc(k, 1); dUin.upload(u, d.st); }
    d.gemm(1.0, dU, dF, 0.0, UtF);                        // t(F.t() * U), cpp:209
    VarSrc us = make_var(u ? &dUin : nullptr, seed, SITE_U_H2, 0, 0, k);
    h2s_core(d, UtF, dF, ds.p, H, dprior.p, us, dh2.p, log_ps_out ? dlp.p : nullptr, det, ll);
    dh2.download(F_h2_out, d.st);
    if (log_ps_out) dlp.download(log_ps_out, d.st);
    d.sync();
    BSFG_API_END
}

// ---- sample_prec_discrete_conditional_c (cpp:241-289) -----------------------------------------
extern "C" int bsfg_sample_prec_discrete_conditional_c(const double* Y, int n, int p, int h2_divisions,
                                                       const double* h2_priors, const double* U, const double* s,
                                                       const double* res_prec, const double* u,
                                                       unsigned long long seed, double* Prec_out, double* log_ps_out) {
    BSFG_API_BEGIN
    REQUIRE_PTR(Y); REQUIRE_PTR(h2_priors); REQUIRE_PTR(U); REQUIRE_PTR(s); REQUIRE_PTR(res_prec); REQUIRE_PTR(Prec_out);
    const int H = h2_divisions;
    if (n <= 0 || p <= 0 || H <= 0) fail(BSFG_ESHAPE, "sample_prec_discrete_conditional_c: bad sizes");
    Device& d = default_device();
    DMat dY(n, p), dU(n, n), UtY(n, p), dUin;
    DVec ds(n), dprior(H), slog(H), drp(p), ll((long)p * H), det((long)p * H), dh2(p), dlp((long)p * H), dout(p);
    dY.upload(Y, d.st); dU.upload(U, d.st); ds.upload(s, d.st); dprior.upload(h2_priors, d.st); drp.upload(res_prec, d.st);
    if (u) { dUin.alloc(p, 1); dUin.upload(u, d.st); }
    d.gemm(1.0, dU, dY, 0.0, UtY);                        // t(Y.t() * U), cpp:258
    prec_logdet_s_kernel<<<H, 256, 0, d.st>>>(ds.p, n, H, slog.p);
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
    {
        const long warps = (long)p * H;
        prec_loglik_kernel<<<ceil_div(warps * 32, 256), 256, 0, d.st>>>(UtY.p, UtY.ld, dY.p, dY.ld, ds.p, slog.p, drp.p,
                                                                       n, p, H, ll.p, det.p);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    VarSrc us = make_var(u ? &dUin : nullptr, seed, SITE_U_PREC, 0, 0, p);
    {
        const int wpb = 4;
        grid_draw_kernel<<<ceil_div(p, wpb), wpb * 32, sizeof(double) * wpb * H, d.st>>>(
            ll.p, p, det.p, p, 1, dprior.p, p, H, us, dh2.p, log_ps_out ? dlp.p : nullptr, p);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    prec_from_h2_kernel<<<ceil_div(p, 256), 256, 0, d.st>>>(drp.p, dh2.p, p, dout.p);
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
    dout.download(Prec_out, d.st);
    if (log_ps_out) dlp.download(log_ps_out, d.st);
    d.sync();
    BSFG_API_END
}

// ---- sample_F_a_c (cpp:293-339) ---------------------------------------------------------------
namespace bsfg {
// b3 = U3' Z_1' F (r x k), v = elementwise, F_a = U3 v, then branch fix-ups.  Z1 may be null (identity).
// Rows [row0, row0 + cnt) of the draw (the whole draw when row0 = 0, cnt = r).  `gather(M)` (may be null) completes a
// row-sharded r x k matrix on every rank; it is called on v before the back-rotation and on Fa at the end.
static void fa_core(Device& d, const DMat& F, const DMat* Z1, const DMat& U3, const DMat& U3t, const double* s1,
                    const double* s2, const double* F_h2, const VarSrc& z, DMat& T1, DMat& b3, DMat& v, DMat& Fa,
                    long row0 = 0, long cnt = -1, const std::function<void(DMat&)>* gather = nullptr) {
    const int n = (int)F.rows, k = (int)F.cols, r = (int)U3.rows;
    if (cnt < 0) cnt = r;
    const DMat* t1 = &F;
    if (Z1) { d.gemm(1.0, *Z1, F, 0.0, T1); t1 = &T1; }    // Z_1' F
    else if (r != n) fail(BSFG_ESHAPE, "Z_1 omitted (identity) but r != n");
    if (cnt > 0) {
        // rows of U'(Z_1'F) <-> columns of U
        dgemm_tn(d.st, d.sms, (int)cnt, k, r, 1.0, U3.p + row0 * U3.ld, U3.ld, t1->p, t1->ld, 0.0, nullptr, 0, b3.p + row0, b3.ld,
                 d.ws.p, Device::WS_DOUBLES);               // U' (Z_1' F), cpp:311
        dim3 grid(std::min(ceil_div(cnt, 256), 64), k);
        fa_v_kernel<<<grid, 256, 0, d.st>>>(b3.p + row0, b3.ld, (int)cnt, k, F_h2, s1 + row0, s2 + row0, z, (int)row0, v.p + row0, v.ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    if (gather) (*gather)(v);
    if (cnt > 0) {
        dgemm_tn(d.st, d.sms, (int)cnt, k, r, 1.0, U3t.p + row0 * U3t.ld, U3t.ld, v.p, v.ld, 0.0, nullptr, 0, Fa.p + row0, Fa.ld,
                 d.ws.p, Device::WS_DOUBLES);               // U * (mlam + z/sqrt(d)), cpp:331
        dim3 grid(std::min(ceil_div(cnt, 256), 64), k);
        fa_fixup_kernel<<<grid, 256, 0, d.st>>>(F.p, F.ld, n, (int)cnt, (int)row0, k, F_h2, Fa.p, Fa.ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    if (gather) (*gather)(Fa);
}
}  // namespace bsfg

extern "C" int bsfg_sample_F_a_c(const double* F, int n, int k, const double* Z_1, int r, const double* F_h2,
                                 const double* U, const double* s1, const double* s2, const double* z,
                                 unsigned long long seed, double* F_a_out) {
    BSFG_API_BEGIN
    REQUIRE_PTR(F); REQUIRE_PTR(Z_1); REQUIRE_PTR(F_h2); REQUIRE_PTR(U); REQUIRE_PTR(s1); REQUIRE_PTR(s2); REQUIRE_PTR(F_a_out);
    if (n <= 0 || k <= 0 || r <= 0) fail(BSFG_ESHAPE, "sample_F_a_c: bad sizes");
    for (int j = 0; j < k; j++)
        if (1.0 / (1.0 - F_h2[j]) > 10000.0 && r != n)
            fail(BSFG_ESHAPE, "sample_F_a_c: tau_e > 10000 copies F[,j] into F_a[,j] (cpp:326-327) but r != n");
    Device& d = default_device();
    DMat dF(n, k), dZ1(n, r), dU(r, r), dUt(r, r), dZ, T1(r, k), b3(r, k), v(r, k), Fa(r, k);
    DVec ds1(r), ds2(r), dh2(k);
    dF.upload(F, d.st); dZ1.upload(Z_1, d.st); dU.upload(U, d.st); ds1.upload(s1, d.st); ds2.upload(s2, d.st); dh2.upload(F_h2, d.st);
    if (z) { dZ.alloc(r, k); dZ.upload(z, d.st); }
    d.transpose(dU, dUt);
    VarSrc zs = make_var(z ? &dZ : nullptr, seed, SITE_Z_FA, 0, 0, r);
    fa_core(d, dF, &dZ1, dU, dUt, ds1.p, ds2.p, dh2.p, zs, T1, b3, v, Fa);
    Fa.download(F_a_out, d.st);
    d.sync();
    BSFG_API_END
}

// ---- update_k (BSFG_functions.R:311-372; px variant BSFG_functions_px.R:301-369): host decides, device applies ----
// One implementation for the sweep's context (sweep.inl: update_k_step) and the stateless bsfg_update_k_c.
struct KAddVariates { VarSrc g_lp, z_lam, z_fa, z_f; double g_delta, u_h2, g_px; bool have_scalars; };
struct KRefs {
    Device* d; const bsfg_priors* pri; uint64_t seed;
    int n, pl, r, kmax; long p_total, p_off; int* k;
    bool z1_identity, px_mode;
    const double* Z1; long ldz1;
    DMat *Lam, *LP, *Lt /* k x pl transposed copy, may be null */, *Plamt, *F, *Fa;
    double *F_h2, *delta, *tauh, *px;
    const double* cnt;        // device, k doubles: #{j: |Lambda[j,h]| < epsilon} over ALL traits (already reduced over ranks)
    int* d_map;               // device scratch, kmax ints
};

// returns true when k changed
static bool update_k_apply(KRefs& s, long i, double uu, const KAddVariates* inj) {
    Device& d = *s.d;
    const bsfg_priors& pri = *s.pri;
    const double prob = 1.0 / exp(pri.b0 + pri.b1 * (double)i);             // R:324
    if (!(uu < prob && i > 200)) return false;                              // R:331
    const int k = *s.k, pl = s.pl, n = s.n, r = s.r;
    std::vector<double> cnt(k);
    BSFG_CUDA(cudaMemcpyAsync(cnt.data(), s.cnt, sizeof(double) * k, cudaMemcpyDeviceToHost, d.st));
    d.sync();
    std::vector<int> vec(k);
    int num = 0;
    bool all_below = true;
    for (int h = 0; h < k; h++) {
        const double lind = cnt[h] / (double)s.p_total;                     // colMeans(abs(Lambda) < epsilon), R:326
        vec[h] = lind >= pri.prop;
        num += vec[h];
        if (!(lind < 0.995)) all_below = false;
    }
    if (i > 20 && num == 0 && all_below && k < 2 * s.p_total) {             // add a column, R:332-341
        if (k + 1 > s.kmax) fail(BSFG_ESHAPE, "update_k wants factor %d but k_max = %d", k + 1, s.kmax);
        const int h = k;
        RngStream rs{s.seed, SITE_K_ADD, (uint32_t)i};
        double g_delta, u_h2, g_px = 1.0;
        if (inj && inj->have_scalars) { g_delta = inj->g_delta; u_h2 = inj->u_h2; g_px = inj->g_px; }
        else {
            g_delta = philox_gamma(rs, 0, pri.delta_2_shape);
            u_h2 = philox_uniform(rs, 1);
            if (s.px_mode) g_px = philox_gamma(rs, (uint64_t)1 << 40, pri.px_shape);
        }
        const double px_new = s.px_mode ? g_px / pri.px_rate : 1.0;         // BSFG_functions_px.R:333
        std::vector<double> delta(k + 1);
        BSFG_CUDA(cudaMemcpyAsync(delta.data(), s.delta, sizeof(double) * k, cudaMemcpyDeviceToHost, d.st));
        d.sync();
        delta[k] = g_delta / pri.delta_2_rate;                              // R:335
        std::vector<double> tauh(k + 1);
        double cp = 1.0;
        for (int q = 0; q <= k; q++) { cp *= delta[q]; tauh[q] = cp; }
        BSFG_CUDA(cudaMemcpyAsync(s.delta, delta.data(), sizeof(double) * (k + 1), cudaMemcpyHostToDevice, d.st));
        BSFG_CUDA(cudaMemcpyAsync(s.tauh, tauh.data(), sizeof(double) * (k + 1), cudaMemcpyHostToDevice, d.st));
        BSFG_CUDA(cudaMemcpyAsync(s.F_h2 + k, &u_h2, sizeof(double), cudaMemcpyHostToDevice, d.st));
        d.sync();   // host vectors go out of scope
        VarSrc g_lp, z_lam, z_fa, z_f;
        if (inj) { g_lp = inj->g_lp; z_lam = inj->z_lam; z_fa = inj->z_fa; z_f = inj->z_f; }
        else {
            // one Philox site per draw family (element = global trait / individual index), scalars on SITE_K_ADD
            g_lp = make_var(nullptr, s.seed, SITE_K_ADD_LP, (uint32_t)i, s.p_off, 1);
            z_lam = make_var(nullptr, s.seed, SITE_K_ADD_ZL, (uint32_t)i, s.p_off, 1);
            z_fa = make_var(nullptr, s.seed, SITE_K_ADD_ZFA, (uint32_t)i, 0, 1);
            z_f = make_var(nullptr, s.seed, SITE_K_ADD_ZF, (uint32_t)i, 0, 1);
        }
        addk_traits_kernel<<<std::min(ceil_div(pl, 256), d.sms * 4), 256, 0, d.st>>>(
            pl, h, pri.Lambda_df, tauh[k], g_lp, z_lam, s.LP->p, s.LP->ld, s.Lam->p, s.Lam->ld, s.Lt ? s.Lt->p : nullptr,
            s.Lt ? s.Lt->ld : 0, s.Plamt->p, s.Plamt->ld, s.px_mode ? 1.0 / sqrt(px_new) : 1.0);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        addk_fa_kernel<<<std::min(ceil_div(r, 256), d.sms * 4), 256, 0, d.st>>>(r, h, u_h2, z_fa, s.Fa->p, s.Fa->ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        addk_f_kernel<<<std::min(ceil_div(n, 256), d.sms * 4), 256, 0, d.st>>>(
            n, r, h, u_h2, s.z1_identity ? nullptr : s.Z1, s.ldz1, s.Fa->p, s.Fa->ld, z_f, s.F->p, s.F->ld,
            s.px_mode ? sqrt(px_new) : 1.0);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        if (s.px_mode) {
            // px_factor[k] and, as BSFG_functions_px.R:328 does, Plam = Lambda_prec * tauh for ALL columns (no px division)
            BSFG_CUDA(cudaMemcpyAsync(s.px + k, &px_new, sizeof(double), cudaMemcpyHostToDevice, d.st));
            d.sync();
            dim3 grid(ceil_div(pl, 32), ceil_div(k + 1, 32)), block(32, 8);
            plam_kernel<<<grid, block, 0, d.st>>>(s.LP->p, s.LP->ld, pl, k + 1, s.tauh, s.Plamt->p, s.Plamt->ld);
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
        }
        *s.k = k + 1;
        return true;
    }
    if (num > 0) {                                                           // drop redundant columns, R:342-359
        std::vector<int> nonred;
        for (int h = 0; h < k; h++) if (!vec[h]) nonred.push_back(h);
        const int knew = (int)nonred.size();
        if (knew < 2) fail(BSFG_ESHAPE, "update_k would leave %d factors; sample_delta needs >= 2", knew);
        std::vector<double> delta(k);
        BSFG_CUDA(cudaMemcpyAsync(delta.data(), s.delta, sizeof(double) * k, cudaMemcpyDeviceToHost, d.st));
        d.sync();
        for (int red = 0; red < k; red++) {                                  // R:347-352
            if (!vec[red]) continue;
            if (red == k - 1) continue;
            delta[red + 1] = delta[red + 1] * delta[red];
        }
        std::vector<double> dnew(knew), tnew(knew);
        double cp = 1.0;
        for (int q = 0; q < knew; q++) { dnew[q] = delta[nonred[q]]; cp *= dnew[q]; tnew[q] = cp; }
        BSFG_CUDA(cudaMemcpyAsync(s.delta, dnew.data(), sizeof(double) * knew, cudaMemcpyHostToDevice, d.st));
        BSFG_CUDA(cudaMemcpyAsync(s.tauh, tnew.data(), sizeof(double) * knew, cudaMemcpyHostToDevice, d.st));
        BSFG_CUDA(cudaMemcpyAsync(s.d_map, nonred.data(), sizeof(int) * knew, cudaMemcpyHostToDevice, d.st));
        d.sync();
        auto cc = [&](DMat& A, int rows) {
            compact_cols_kernel<<<std::min(ceil_div(std::max(rows, 1), 256), d.sms * 4), 256, 0, d.st>>>(A.p, A.ld, rows, knew, s.d_map);
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
        };
        cc(*s.Lam, pl); cc(*s.LP, pl); cc(*s.F, n); cc(*s.Fa, r);
        compact_cols_kernel<<<1, 32, 0, d.st>>>(s.F_h2, 1, 1, knew, s.d_map);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        if (s.px_mode) {
            compact_cols_kernel<<<1, 32, 0, d.st>>>(s.px, 1, 1, knew, s.d_map);          // px_factor[nonred], :351
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
        }
        if (s.Lt) {
            compact_rows_kernel<<<std::min(ceil_div(pl, 256), d.sms * 4), 256, 0, d.st>>>(s.Lt->p, s.Lt->ld, pl, knew, s.d_map);
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
        }
        *s.k = knew;
        {
            dim3 grid(ceil_div(pl, 32), ceil_div(knew, 32)), block(32, 8);
            plam_kernel<<<grid, block, 0, d.st>>>(s.LP->p, s.LP->ld, pl, knew, s.tauh, s.Plamt->p, s.Plamt->ld);   // R:356
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
        }
        return true;
    }
    return false;
}

// ---- update_k_c: the stateless drop-in of the R function (not Rcpp in the reference; SURVEY.md 8b lists it as a new export) ----
extern "C" int bsfg_update_k_c(int n, int p, int r, int k, int k_max, long long i, const bsfg_priors* priors,
                               const double* Z_1, double* Lambda, double* Lambda_prec, double* Plam, double* delta,
                               double* tauh, double* F_h2, double* F, double* F_a, double* px_factor, double* F_px,
                               const bsfg_variates* variates, unsigned long long seed, int* k_out) {
    BSFG_API_BEGIN
    REQUIRE_PTR(priors); REQUIRE_PTR(Lambda); REQUIRE_PTR(Lambda_prec); REQUIRE_PTR(Plam); REQUIRE_PTR(delta); REQUIRE_PTR(tauh);
    REQUIRE_PTR(F_h2); REQUIRE_PTR(F); REQUIRE_PTR(F_a); REQUIRE_PTR(k_out);
    if (n <= 0 || p <= 0 || r <= 0 || k < 1 || k_max < k) fail(BSFG_ESHAPE, "update_k: inconsistent sizes");
    if (!Z_1 && r != n) fail(BSFG_ESHAPE, "update_k: Z_1 = NULL means the identity and requires r == n");
    if ((px_factor == nullptr) != (F_px == nullptr)) fail(BSFG_EARG, "update_k: px_factor and F_px go together");
    Device& d = default_device();
    // in/out arrays have capacity k_max columns with leading dimension = rows; only the first k (then k_out) are meaningful
    DMat Lam(p, k_max), LP(p, k_max), Plamt(k_max, p), Fm(n, k_max), Fa(r, k_max), Z1;
    DVec h2(k_max), dl(k_max), th(k_max), px(k_max), cc(2 * (long)k_max);
    int* d_map = nullptr;
    BSFG_CUDA(cudaMalloc(&d_map, sizeof(int) * (size_t)k_max));
    struct MapGuard { int* p; ~MapGuard() { cudaFree(p); } } guard{d_map};
    auto up = [&](DMat& M, const double* host, long rows) {
        BSFG_CUDA(cudaMemcpy2DAsync(M.p, M.ld * sizeof(double), host, rows * sizeof(double), rows * sizeof(double), k,
                                    cudaMemcpyHostToDevice, d.st));
    };
    up(Lam, Lambda, p); up(LP, Lambda_prec, p); up(Fm, F, n); up(Fa, F_a, r);
    BSFG_CUDA(cudaMemcpyAsync(h2.p, F_h2, sizeof(double) * k, cudaMemcpyHostToDevice, d.st));
    BSFG_CUDA(cudaMemcpyAsync(dl.p, delta, sizeof(double) * k, cudaMemcpyHostToDevice, d.st));
    BSFG_CUDA(cudaMemcpyAsync(th.p, tauh, sizeof(double) * k, cudaMemcpyHostToDevice, d.st));
    DMat Fpx;
    if (px_factor) {
        BSFG_CUDA(cudaMemcpyAsync(px.p, px_factor, sizeof(double) * k, cudaMemcpyHostToDevice, d.st));
        Fpx.alloc(n, k_max);
        up(Fpx, F_px, n);
    }
    if (Z_1) { Z1.alloc(n, r); Z1.upload(Z_1, d.st); }
    {   // Plam stays as passed unless k changes (R returns current_state untouched then)
        DMat Pl(p, k_max);
        up(Pl, Plam, p);
        DMat Plv; Plv.p = Pl.p; Plv.rows = p; Plv.cols = k; Plv.ld = Pl.ld;
        DMat Ptv; Ptv.p = Plamt.p; Ptv.rows = k; Ptv.cols = p; Ptv.ld = Plamt.ld;
        d.transpose(Plv, Ptv);
        d.sync();
    }
    factor_reduce_kernel<<<k, 256, 0, d.st>>>(LP.p, LP.ld, Lam.p, Lam.ld, p, k, priors->epsilon, cc.p, cc.p + k_max);
    BSFG_CHECK_LAUNCH(); g_launch_counter++;

    const uint32_t ii = (uint32_t)i;
    double uu;
    if (variates && variates->u_k) uu = *variates->u_k;
    else { RngStream rs{seed, SITE_U_K, ii}; uu = philox_uniform(rs, 0); }
    KAddVariates ka;
    DVec vk;
    const bool have_add = variates && variates->k_g_lp && variates->k_g_delta && variates->k_z_lambda && variates->k_u_h2 &&
                          variates->k_z_Fa && variates->k_z_F && (!px_factor || variates->k_g_px);
    if (have_add) {
        vk.alloc(2L * p + r + n);
        BSFG_CUDA(cudaMemcpyAsync(vk.p, variates->k_g_lp, sizeof(double) * p, cudaMemcpyHostToDevice, d.st));
        BSFG_CUDA(cudaMemcpyAsync(vk.p + p, variates->k_z_lambda, sizeof(double) * p, cudaMemcpyHostToDevice, d.st));
        BSFG_CUDA(cudaMemcpyAsync(vk.p + 2L * p, variates->k_z_Fa, sizeof(double) * r, cudaMemcpyHostToDevice, d.st));
        BSFG_CUDA(cudaMemcpyAsync(vk.p + 2L * p + r, variates->k_z_F, sizeof(double) * n, cudaMemcpyHostToDevice, d.st));
        d.sync();
        auto mk = [&](double* q) { VarSrc v = make_var(nullptr, 0, 0, 0, 0, 1); v.ptr = q; v.ld = 0; return v; };
        ka.g_lp = mk(vk.p); ka.z_lam = mk(vk.p + p); ka.z_fa = mk(vk.p + 2L * p); ka.z_f = mk(vk.p + 2L * p + r);
        ka.g_delta = *variates->k_g_delta; ka.u_h2 = *variates->k_u_h2; ka.g_px = px_factor ? *variates->k_g_px : 1.0;
        ka.have_scalars = true;
    }
    int knew = k;
    KRefs s;
    s.d = &d; s.pri = priors; s.seed = seed;
    s.n = n; s.pl = p; s.r = r; s.kmax = k_max; s.p_total = p; s.p_off = 0; s.k = &knew;
    // The R functions hold the UNEXPANDED Lambda and F (the sweep's context stores Lambda_px / F_px instead, hence its
    // px_mode multipliers); px_factor / F_px are handled below, as BSFG_functions_px.R:333-334, 351-352 do.
    s.z1_identity = Z_1 == nullptr; s.px_mode = false;
    s.Z1 = Z1.p; s.ldz1 = Z1.ld;
    s.Lam = &Lam; s.LP = &LP; s.Lt = nullptr; s.Plamt = &Plamt; s.F = &Fm; s.Fa = &Fa;
    s.F_h2 = h2.p; s.delta = dl.p; s.tauh = th.p; s.px = px.p;
    s.cnt = cc.p + k_max; s.d_map = d_map;
    const bool changed = update_k_apply(s, (long)i, uu, have_add ? &ka : nullptr);
    if (changed && knew > k) {
        // R recomputes Plam = sweep(Lambda_prec, 2, tauh, '*') over ALL columns in the add arm (R:337)
        dim3 grid(ceil_div(p, 32), ceil_div(knew, 32)), block(32, 8);
        plam_kernel<<<grid, block, 0, d.st>>>(LP.p, LP.ld, p, knew, th.p, Plamt.p, Plamt.ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    if (changed && px_factor) {
        if (knew > k) {
            double g_px;
            if (have_add) g_px = *variates->k_g_px;
            else { RngStream rs{seed, SITE_K_ADD, ii}; g_px = philox_gamma(rs, (uint64_t)1 << 40, priors->px_shape); }
            const double pxv[2] = {g_px / priors->px_rate, sqrt(g_px / priors->px_rate)};
            BSFG_CUDA(cudaMemcpyAsync(px.p + k, &pxv[0], sizeof(double), cudaMemcpyHostToDevice, d.st));
            BSFG_CUDA(cudaMemcpyAsync(cc.p, &pxv[1], sizeof(double), cudaMemcpyHostToDevice, d.st));
            d.sync();
            DMat fin; fin.p = Fm.p + (long)k * Fm.ld; fin.rows = n; fin.cols = 1; fin.ld = Fm.ld;
            DMat fout; fout.p = Fpx.p + (long)k * Fpx.ld; fout.rows = n; fout.cols = 1; fout.ld = Fpx.ld;
            d.scale(fin, nullptr, cc.p, fout);                                  // F_px[,k] = F[,k] sqrt(px_factor[k])
        } else {
            compact_cols_kernel<<<std::min(ceil_div(n, 256), d.sms * 4), 256, 0, d.st>>>(Fpx.p, Fpx.ld, n, knew, d_map);
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
            compact_cols_kernel<<<1, 32, 0, d.st>>>(px.p, 1, 1, knew, d_map);
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
        }
    }
    if (changed) {
        auto down = [&](const DMat& M, double* host, long rows) {
            BSFG_CUDA(cudaMemcpy2DAsync(host, rows * sizeof(double), M.p, M.ld * sizeof(double), rows * sizeof(double), knew,
                                        cudaMemcpyDeviceToHost, d.st));
        };
        DMat Pl(p, k_max);
        DMat Plv; Plv.p = Pl.p; Plv.rows = p; Plv.cols = knew; Plv.ld = Pl.ld;
        DMat Ptv; Ptv.p = Plamt.p; Ptv.rows = knew; Ptv.cols = p; Ptv.ld = Plamt.ld;
        d.transpose(Ptv, Plv);
        down(Lam, Lambda, p); down(LP, Lambda_prec, p); down(Pl, Plam, p); down(Fm, F, n); down(Fa, F_a, r);
        BSFG_CUDA(cudaMemcpyAsync(F_h2, h2.p, sizeof(double) * knew, cudaMemcpyDeviceToHost, d.st));
        BSFG_CUDA(cudaMemcpyAsync(delta, dl.p, sizeof(double) * knew, cudaMemcpyDeviceToHost, d.st));
        BSFG_CUDA(cudaMemcpyAsync(tauh, th.p, sizeof(double) * knew, cudaMemcpyDeviceToHost, d.st));
        if (px_factor) {
            BSFG_CUDA(cudaMemcpyAsync(px_factor, px.p, sizeof(double) * knew, cudaMemcpyDeviceToHost, d.st));
            down(Fpx, F_px, n);
        }
        d.sync();
    }
    d.sync();
    *k_out = knew;
    BSFG_API_END
}

// ---- sample_delta_c (BSFG_functions.R:272-308) ------------------------------------------------
extern "C" int bsfg_sample_delta_c(const double* delta, const double* tauh, int k, const double* Lambda_prec, int p,
                                   double delta_1_shape, double delta_1_rate, double delta_2_shape,
                                   double delta_2_rate, const double* Lambda2, const double* g,
                                   unsigned long long seed, double* delta_out, int* n_draws_out,
                                   double* shapes_out, double* rates_out) {
    BSFG_API_BEGIN
    REQUIRE_PTR(delta); REQUIRE_PTR(Lambda_prec); REQUIRE_PTR(Lambda2); REQUIRE_PTR(delta_out);
    (void)tauh;   // R recomputes tauh = cumprod(delta) after the first draw; the input tauh only enters
                  // the first rate, where it equals cumprod(delta) by the sweep's invariant (R:253-254).
    if (k < 2) fail(BSFG_ESHAPE, "sample_delta: k >= 2 required (R's 2:(k-1) runs off the vector otherwise)");
    if (p <= 0) fail(BSFG_ESHAPE, "sample_delta: p must be positive");
    Device& d = default_device();
    DMat LP(p, k), L2(p, k), dG;
    DVec dd(k), colsum(k), shapes(k + 2), rates(k + 2);
    int* d_nd = nullptr;
    BSFG_CUDA(cudaMalloc(&d_nd, sizeof(int)));
    LP.upload(Lambda_prec, d.st); L2.upload(Lambda2, d.st); dd.upload(delta, d.st);
    if (g) { dG.alloc(k + 2, 1); BSFG_CUDA(cudaMemcpyAsync(dG.p, g, sizeof(double) * (size_t)(k == 2 ? 3 : k - 1), cudaMemcpyHostToDevice, d.st)); }
    delta_colsum_kernel<<<k, 256, 0, d.st>>>(LP.p, LP.ld, L2.p, L2.ld, p, k, 0, colsum.p);
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
    VarSrc gs = make_var(g ? &dG : nullptr, seed, SITE_G_DELTA, 0, 0, k);
    delta_recursion_kernel<<<1, 128, (4 * k + 2) * sizeof(double), d.st>>>(dd.p, k, colsum.p, (double)p, delta_1_shape, delta_1_rate,
                                                               delta_2_shape, delta_2_rate, gs, nullptr, shapes.p, rates.p, d_nd);
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
    dd.download(delta_out, d.st);
    int nd = 0;
    BSFG_CUDA(cudaMemcpyAsync(&nd, d_nd, sizeof(int), cudaMemcpyDeviceToHost, d.st));
    d.sync();
    cudaFree(d_nd);
    if (n_draws_out) *n_draws_out = nd;
    if (shapes_out) BSFG_CUDA(cudaMemcpy(shapes_out, shapes.p, sizeof(double) * nd, cudaMemcpyDeviceToHost));
    if (rates_out) BSFG_CUDA(cudaMemcpy(rates_out, rates.p, sizeof(double) * nd, cudaMemcpyDeviceToHost));
    BSFG_API_END
}
