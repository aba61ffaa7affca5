// Host side of the FP64 GEMM core: tensor-map encoding and launch geometry.
// Included by bsfg.cu (single translation unit).

namespace bsfg {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        BSFG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !ptr) fail(BSFG_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
        fn = (PFN_encodeTiled)ptr;
    }
    return fn;
}

// 2-D map over a K-contiguous operand: dim0 = K (contiguous), dim1 = rows (M or N), box = 16 x box_rows
static void make_operand_map(CUtensorMap* tm, const double* base, long K, long rows, long ld, int box_rows) {
    if (((uintptr_t)base & 15) != 0) fail(BSFG_EARG, "GEMM operand not 16-byte aligned");
    if (ld % 2 != 0) fail(BSFG_EARG, "GEMM operand leading dimension %ld must be even", ld);
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(BSFG_ECUDA, "cuTensorMapEncodeTiled failed (%d) K=%ld rows=%ld ld=%ld", (int)r, K, rows, ld);
}

constexpr int GEMM_STAGES = 6;

static long long g_launch_counter = 0;   // kernels launched by this library (bench.py's gpu_launches)

static bool persistent_enabled() {
    static const bool on = [] { const char* e = getenv("BSFG_GEMM_PERSISTENT"); return !e || atoi(e) != 0; }();
    return on;
}

template <int BM>
static void launch_gemm(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g, dim3 grid, bool streamk) {
    using Cfg = GemmSmem<BM, 128, GEMM_STAGES>;
    // Warp layout WM x (8 / WM): WM = 2 ((BM/2) x 32 warp tiles) for every multiple of 16; the one tile height that is not,
    // BM = 104 (13 row fragments: M = k = 97..104 factors, e.g. the k = 100 of C4 without the 12% zero fill of a 112-row
    // tile), runs as WM = 1 (104 x 16 warp tiles, 15 fragment loads per 26 DMMAs, no predicated DMMA).  The all-rows /
    // all-columns layouts for the OTHER heights were measured slower on B200 (sweep 13.3 -> 14.0 ms at C4: 18 instead of 12
    // fragment loads per 32 DMMAs and a predicated DMMA sequence) and are not instantiated.
    constexpr int WM = (BM % 16 == 0) ? 2 : 1;
    static bool attr_set = false;
    if (!attr_set) {
        BSFG_CUDA(cudaFuncSetAttribute(dgemm_tn_kernel<BM, 128, GEMM_STAGES, WM>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::TOTAL));
        BSFG_CUDA(cudaFuncSetAttribute(dgemm_tn_streamk_kernel<BM, 128, GEMM_STAGES, WM>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::TOTAL));
        attr_set = true;
    }
    // The stream-K decomposition is a separate kernel (gemm_dmma.cuh says why).
    if (streamk) dgemm_tn_streamk_kernel<BM, 128, GEMM_STAGES, WM><<<grid, GEMM_THREADS, Cfg::TOTAL, st>>>(tmA, tmB, g);
    else dgemm_tn_kernel<BM, 128, GEMM_STAGES, WM><<<grid, GEMM_THREADS, Cfg::TOTAL, st>>>(tmA, tmB, g);
}

static void dgemm_tn_impl(cudaStream_t st, int sm_count, int M, int N, int K, double alpha, const double* At, long lda,
                          const double* Bm, long ldb, double beta, const double* Cin, long ldcin, double* C, long ldc,
                          double* ws, size_t ws_doubles, const int* run_if, int trans_out) {
    // row-tile height: the smallest instantiated BM that holds a single short row tile, else 128
    // row-tile height: 128 unless another instantiated height wastes at least 3% fewer rows on zero fill (a single short row
    // tile for M = k factors: 104 rows for k = 100 instead of 112 (measured: sweep 10.30 -> 10.05 ms at C4); two 104-row tiles
    // for k = 200 instead of two of 128)
    static const bool bm104 = [] { const char* e = getenv("BSFG_GEMM_BM104"); return !e || atoi(e) != 0; }();
    int BM = 128;
    {
        const long pad128 = (long)ceil_div(M, 128) * 128;
        long best = pad128;
        for (int cand : {112, 104, 96, 64, 32}) {
            if (cand == 104 && !bm104) continue;
            if (cand < 96 && M > cand) continue;          // the short tiles only as a single row tile
            const long pad = (long)ceil_div(M, cand) * cand;
            if (pad < best && (double)pad <= 0.97 * (double)pad128) { best = pad; BM = cand; }
        }
    }
    CUtensorMap tmA, tmB;
    make_operand_map(&tmA, At, K, M, lda, BM);
    make_operand_map(&tmB, Bm, K, N, ldb, 128);
    GemmArgs g;
    g.M = M; g.N = N; g.K = K;
    g.C = C; g.ldc = ldc;
    g.Cin = (beta != 0.0) ? Cin : nullptr; g.ldcin = ldcin;
    g.alpha = alpha; g.beta = beta;
    g.run_if = run_if;
    g.trans_out = trans_out;
    g.tiles_m = ceil_div(M, BM);
    g.tiles_n = ceil_div(N, 128);
    g.group_n = 8;
    const int tiles = g.tiles_m * g.tiles_n;
    const int ktiles = ceil_div(K, GEMM_BK);
    // split-K, chosen by a small cost model (microseconds): waves of one CTA per SM, 2.1 us per 128x128x16 k-tile at
    // the DMMA rate of one SM (scaled by BM/128), ~8 us of pipeline fill + epilogue per CTA, and for s > 1 the extra pass
    // over the s x M x N partials plus the reduce launch.  E.g. M=k=100, N=p=20000, K=5000 (157 tiles = 2 waves at 53%):
    // s=5 -> ~750 us instead of 1330; Lambda'Lmsg (1 tile, K=20000): s=148.  Partials are summed in a fixed order by
    // splitk_reduce_kernel, so results stay bit-reproducible.
    // the last 64 Ki ints of the workspace are the per-tile arrival counters of the fused split-K reduction
    constexpr size_t COUNTER_DOUBLES = 32768;
    int* counters = nullptr;
    if (ws && ws_doubles > 4 * COUNTER_DOUBLES) {
        ws_doubles -= COUNTER_DOUBLES;
        counters = reinterpret_cast<int*>(ws + ws_doubles);
    }
    int splits = 1;
    if (ws) {
        const double per = (double)M * (double)N;
        const double kt_us = 2.1 * BM / 128.0;
        static const double item_us = [] { const char* e = getenv("BSFG_GEMM_ITEM_US"); return e ? atof(e) : 8.0; }();
        static const double red_scale = [] { const char* e = getenv("BSFG_GEMM_RED_SCALE"); return e ? atof(e) : 1.0; }();
        auto cost = [&](int s) {
            const long ctas = (long)tiles * s;
            const double waves = (double)((ctas + sm_count - 1) / sm_count);
            double c = waves * (ceil_div(ktiles, s) * kt_us + item_us);
            if (s > 1) c += red_scale * (16e-6 * s * per / 5.0 + 5.0);
            return c;
        };
        double best = cost(1);
        const int smax = std::min(std::min(ktiles / 4, 2 * sm_count), (int)std::min<double>(1e9, (double)ws_doubles / per));
        for (int s = 2; s <= smax; s++) {
            const double c = cost(s);
            if (c < 0.95 * best) { best = c; splits = s; }
        }
    }
    g.ktiles_per_split = ceil_div(ktiles, splits);
    splits = ceil_div(ktiles, g.ktiles_per_split);    // no empty splits
    g.splits = splits;
    g.ws = ws;
    // Stream-K (BSFG_GEMM_STREAMK=0 turns it off): when the (tile, split) items of the plan above do not fill whole waves
    // -- 157 column tiles x 5 splits = 5.3 waves of 148 SMs run as 6 -- cut the tiles x k-tiles work into one equal
    // contiguous run per SM instead.  Complete tiles are stored directly; only the (at most two per CTA, when a run is
    // longer than a tile) partial pieces go through the workspace.  Used when it removes at least 4% of idle SM time and the
    // slots fit the workspace; products with many waves of whole tiles keep the plain persistent tiling.
    static const bool sk_ok = [] { const char* e = getenv("BSFG_GEMM_STREAMK"); return !e || atoi(e) != 0; }();
    bool use_sk = false;
    long sk_total = 0;
    if (sk_ok && ws && persistent_enabled()) {
        const long items0 = (long)tiles * splits;
        const double fill = (double)items0 / ((double)sm_count * (double)((items0 + sm_count - 1) / sm_count));
        const long U = (long)tiles * ktiles;
        const long G = std::min<long>(sm_count, std::max<long>(1, U / 4));       // at least 4 k-tiles per CTA
        const int P = (int)((U + G - 1) / G);
        const long slots = (ktiles + P - 1) / P + 1;
        // not for products with few k-tiles per tile (K = k, K = R): they are mostly epilogue, and a partial piece costs a
        // workspace round trip on top of it (measured at N = 8: the K = 100 product of the theta step 0.28 -> 0.39 ms)
        if (fill < 0.96 && ktiles >= 48 && G == sm_count && (double)slots * (double)M * (double)N <= (double)ws_doubles) {
            use_sk = true; sk_total = U;
            g.ktiles_per_split = P;
            g.splits = 1;
        }
    }
    // Fused reduction (the last CTA to deliver a partial of a tile sums them): OFF by default.  Measured at C4 on one B200
    // (profiles/r02c_bench_c4_n1_fused_reduce_{on,off}.json): 11.64 ms per iteration with it, 10.81 ms without -- the summing
    // CTA reads splits x 112 x 128 partials through one SM while its neighbours idle at the end of the wave, which costs
    // more than the 14 reduce launches it saves.  BSFG_GEMM_FUSED_REDUCE=1 turns it on (results are identical: same order).
    static const bool fuse_ok = [] { const char* e = getenv("BSFG_GEMM_FUSED_REDUCE"); return e && atoi(e) != 0; }();
    const bool fused = fuse_ok && !use_sk && splits > 1 && splits <= 16 && counters && tiles <= 65536 && run_if == nullptr;
    g.tile_counter = fused ? counters : nullptr;
    // persistent launch: one CTA per SM walks the (tile, split) items; BSFG_GEMM_PERSISTENT=0 launches one CTA per item
    const bool persistent = persistent_enabled();
    const long items = (long)tiles * splits;
    dim3 grid((unsigned)(use_sk ? ceil_div(sk_total, g.ktiles_per_split) : (persistent ? std::min<long>(items, sm_count) : items)));
    // (Tried in round 2 and removed: a variant of the kernel for the nearly-all-epilogue K = k product that sends every thread's
    // 64 Cin values to a private shared-memory slot with 8-byte cp.async at the start of the tile -- three ring stages + a
    // 128 KB staging tile.  Measured SLOWER at C4: the product 0.91 -> 1.09 ms, theta phase 1.73 -> 1.91 ms.)
    switch (BM) {
        case 32: launch_gemm<32>(st, tmA, tmB, g, grid, use_sk); break;
        case 64: launch_gemm<64>(st, tmA, tmB, g, grid, use_sk); break;
        case 96: launch_gemm<96>(st, tmA, tmB, g, grid, use_sk); break;
        case 104: launch_gemm<104>(st, tmA, tmB, g, grid, use_sk); break;
        case 112: launch_gemm<112>(st, tmA, tmB, g, grid, use_sk); break;
        default: launch_gemm<128>(st, tmA, tmB, g, grid, use_sk); break;
    }
    BSFG_CHECK_LAUNCH();
    g_launch_counter++;
    if (use_sk) {
        const size_t total = (size_t)M * N;
        int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count * 8);
        streamk_reduce_kernel<128><<<blocks, 256, 0, st>>>(g, BM);
        BSFG_CHECK_LAUNCH();
        g_launch_counter++;
    } else if (splits > 1 && !fused) {
        const size_t total = (size_t)M * N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > sm_count * 8) blocks = sm_count * 8;
        splitk_reduce_kernel<<<blocks, 256, 0, st>>>(g);
        BSFG_CHECK_LAUNCH();
        g_launch_counter++;
    }
}

void dgemm_tn(cudaStream_t st, int sm_count, int M, int N, int K, double alpha, const double* At, long lda,
              const double* Bm, long ldb, double beta, const double* Cin, long ldcin, double* C, long ldc,
              double* ws, size_t ws_doubles, const int* run_if) {
    if (M <= 0 || N <= 0) return;
    if (K <= 0) fail(BSFG_ESHAPE, "dgemm_tn: K must be positive");
    // C = At'Bm with a short N (N = k factors) and a long M: run the transposed product Bm'At, whose short dimension is
    // the row tile, and store it transposed
    if (N <= 112 && M > 128)
        dgemm_tn_impl(st, sm_count, N, M, K, alpha, Bm, ldb, At, lda, beta, Cin, ldcin, C, ldc, ws, ws_doubles, run_if, 1);
    else
        dgemm_tn_impl(st, sm_count, M, N, K, alpha, At, lda, Bm, ldb, beta, Cin, ldcin, C, ldc, ws, ws_doubles, run_if, 0);
}

}  // namespace bsfg
