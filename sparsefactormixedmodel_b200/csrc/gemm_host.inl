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
using GemmCfg = GemmSmem<128, 128, GEMM_STAGES>;

static long long g_launch_counter = 0;   // kernels launched by this library (bench.py's gpu_launches)

void dgemm_tn(cudaStream_t st, int sm_count, int M, int N, int K, double alpha, const double* At, long lda,
              const double* Bm, long ldb, double beta, const double* Cin, long ldcin, double* C, long ldc,
              double* ws, size_t ws_doubles, const int* run_if) {
    if (M <= 0 || N <= 0) return;
    if (K <= 0) fail(BSFG_ESHAPE, "dgemm_tn: K must be positive");
    static bool attr_set = false;
    if (!attr_set) {
        BSFG_CUDA(cudaFuncSetAttribute(dgemm_tn_kernel<128, 128, GEMM_STAGES, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg::TOTAL));
        attr_set = true;
    }
    CUtensorMap tmA, tmB;
    make_operand_map(&tmA, At, K, M, lda, 128);
    make_operand_map(&tmB, Bm, K, N, ldb, 128);
    GemmArgs g;
    g.M = M; g.N = N; g.K = K;
    g.C = C; g.ldc = ldc;
    g.Cin = (beta != 0.0) ? Cin : nullptr; g.ldcin = ldcin;
    g.alpha = alpha; g.beta = beta;
    g.run_if = run_if;
    g.tiles_m = ceil_div(M, 128);
    g.tiles_n = ceil_div(N, 128);
    g.group_n = 8;
    const int tiles = g.tiles_m * g.tiles_n;
    const int ktiles = ceil_div(K, GEMM_BK);
    // split-K, chosen by a small cost model (microseconds): waves of one CTA per SM, 2.1 us per 128x128x16 k-tile at
    // the DMMA rate of one SM, ~8 us of pipeline fill + epilogue per CTA, and for s > 1 the extra pass over the
    // s x M x N partials plus the reduce launch.  E.g. M=k=100, N=p=20000, K=5000 (157 tiles = 2 waves at 53%):
    // s=7 -> 870 us instead of 1330; Lambda'Lmsg (1 tile, K=20000): s=148.  Partials are summed in a fixed order by
    // splitk_reduce_kernel, so results stay bit-reproducible.
    int splits = 1;
    if (ws) {
        const double per = (double)M * (double)N;
        const double kt_us = 2.1;
        auto cost = [&](int s) {
            const long ctas = (long)tiles * s;
            const double waves = (double)((ctas + sm_count - 1) / sm_count);
            double c = waves * (ceil_div(ktiles, s) * kt_us + 8.0);
            if (s > 1) c += 16e-6 * s * per / 5.0 + 5.0;
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
    dim3 grid(tiles, splits);
    // WM = 2 (64 x 32 warp tiles) for every shape: the all-rows / all-columns layouts (WM = 1 / 8) that skip the empty
    // fragments of M = k or N = k products evenly were measured SLOWER on B200 (sweep 13.3 -> 14.0 ms at C4: 18 instead
    // of 12 fragment loads per 32 DMMAs and a predicated DMMA sequence), so they are not instantiated.
    dgemm_tn_kernel<128, 128, GEMM_STAGES, 2><<<grid, GEMM_THREADS, GemmCfg::TOTAL, st>>>(tmA, tmB, g);
    BSFG_CHECK_LAUNCH();
    g_launch_counter++;
    if (splits > 1) {
        const size_t total = (size_t)M * N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > sm_count * 8) blocks = sm_count * 8;
        splitk_reduce_kernel<<<blocks, 256, 0, st>>>(g);
        BSFG_CHECK_LAUNCH();
        g_launch_counter++;
    }
}

}  // namespace bsfg
