// Stand-alone experiment, NOT part of the product.
// Run on a B200 in round 2 with S = 6, 7, 8: accuracy and rates in profiles/r02_i8_sliced_dgemm.log.
// Second step after tools/i8_umma_gemm.cu: the whole INT8-slice emulation of an FP64 "TN" GEMM
// (numerics in tools/ozaki_prototype.py) as one stand-alone program, so that accuracy and speed
// of the scheme can be measured on a B200 before anything is wired into the sweep:
//
//   C[M x N] (FP64, row-major) = A[M x K] * B[N x K]',   A and B FP64 with K contiguous (the library's operand form)
//
//   1. slice_rows_kernel: per row, e = ilogb(max|x|) + 2, x 2^-e in (-0.5, 0.5), S signed 7-bit digits by truncation
//      (q_t = trunc(128 r), r <- 128 r - q_t, exact in FP64), written as S int8 planes [t][row][K].
//   2. i8_sliced_gemm_kernel (persistent, 320 threads): for every output tile and every level L = t + u (largest L first,
//      i.e. smallest weight first), all digit pairs of the level and all k-blocks accumulate EXACTLY in one int32 TMEM
//      accumulator (pairs * K * 127^2 < 2^31); the 8 epilogue warps then add (double)acc * 2^(-7 L) into FP64 registers
//      (64 columns per thread) while the MMA warp already works on the next level in the other accumulator stage; after
//      the last level the sums are scaled by 2^(e_row + e_col) and stored.  Pairs with t + u > S + 1 are dropped.
//        warp 0: TMA producer (A digit plane t, B digit plane u; SWIZZLE_128B; 6-stage ring of (128 + 128) x 128 bytes)
//        warp 1: MMA issuer   (tcgen05.mma.cta_group::1.kind::i8, M 128 x N 128 x K 32)
//        warps 2..9: epilogue (tcgen05.ld 32x32b; warp w reads TMEM lanes 32 (w % 4).., columns 64 ((w - 2) / 4)..)
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/i8_sliced_dgemm tools/i8_sliced_dgemm.cu -lcuda
// Run:    tools/i8_sliced_dgemm [S]      (default S = 8 digits; prints the error against a long-double CPU product on small
//                                         shapes, then timings on the sweep's shapes incl. the slicing passes)
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int BM = 128, BN = 128, BK = 128, UMMA_K = 32;
constexpr int STAGES = 6;
constexpr int A_BYTES = BM * BK, B_BYTES = BN * BK, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int ACC_STAGES = 2, TMEM_COLS = ACC_STAGES * BN;    // 256
constexpr int EPI_WARPS = 8, THREADS = 64 + 32 * EPI_WARPS;    // 320
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int DIGIT_BITS = 7;
constexpr int MAX_S = 9;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// see tools/i8_umma_gemm.cu for the field-by-field derivation of the two descriptors
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_i8(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- 1. digit planes -------------------------------------------------------------------------------------------------
// X: rows x K doubles, K contiguous (row pitch ldx).  planes: S x rows_pad x Kp int8 (zero initialised), expo: rows ints.
__global__ void slice_rows_kernel(const double* __restrict__ X, long ldx, int rows, int K, int S, int8_t* __restrict__ planes,
                                  long rows_pad, long Kp, int* __restrict__ expo) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const double* x = X + (long)row * ldx;
    double amax = 0.0;
    for (int k = lane; k < K; k += 32) amax = fmax(amax, fabs(x[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const int e = (amax > 0.0) ? ilogb(amax) + 2 : 0;            // |x| 2^-e < 0.5
    if (lane == 0) expo[row] = e;
    for (int k = lane; k < K; k += 32) {
        double r = scalbn(x[k], -e);
        for (int t = 0; t < S; t++) {
            r *= (double)(1 << DIGIT_BITS);
            const double q = trunc(r);                           // |q| <= 63 here (|r| < 0.5 * 128), always inside int8
            planes[((long)t * rows_pad + row) * Kp + k] = (int8_t)(int)q;
            r -= q;                                              // exact
        }
    }
}

// ---- 2. the sliced GEMM ----------------------------------------------------------------------------------------------
struct Args { int M, N, K, S; long Mpad, Npad; double* C; const int* ea; const int* eb; int tiles_m, tiles_n; };

__global__ void __launch_bounds__(THREADS, 1)
i8_sliced_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Args g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;
    uint64_t* acc_empty = acc_full + ACC_STAGES;
    uint32_t* tmem_base_p = (uint32_t*)(acc_empty + ACC_STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = g.tiles_m * g.tiles_n;
    const int nkb = (g.K + BK - 1) / BK;
    const int S = g.S;
    // levels L = t + u with 1 <= t, u <= S, truncated at L <= S + 1, walked from S + 1 (smallest weight) down to 2

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < ACC_STAGES; s++) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_p)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_p;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t put = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int m0 = (tile % g.tiles_m) * BM, n0 = (tile / g.tiles_m) * BN;
                for (int L = S + 1; L >= 2; L--)
                    for (int t = max(1, L - S); t <= min(S, L - 1); t++) {
                        const int u = L - t;
                        for (int kb = 0; kb < nkb; kb++, put++) {
                            const int s = put % STAGES;
                            mbar_wait(&empty_bar[s], ((put / STAGES) & 1) ^ 1);
                            mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                            uint8_t* sa = smem + s * STAGE_BYTES;
                            tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, (int)((t - 1) * g.Mpad) + m0);
                            tma_load_2d(sa + A_BYTES, &tmB, &full_bar[s], kb * BK, (int)((u - 1) * g.Npad) + n0);
                        }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_i8(BM, BN);
            uint32_t get = 0, lvl_i = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int L = S + 1; L >= 2; L--, lvl_i++) {
                    const int as = lvl_i % ACC_STAGES;
                    mbar_wait(&acc_empty[as], ((lvl_i / ACC_STAGES) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
                    uint32_t first = 1;
                    for (int t = max(1, L - S); t <= min(S, L - 1); t++)
                        for (int kb = 0; kb < nkb; kb++, get++) {
                            const int s = get % STAGES;
                            mbar_wait(&full_bar[s], (get / STAGES) & 1);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                            const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + A_BYTES);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; k++) {
                                umma_i8(tmem_d, da + (uint64_t)(k * (UMMA_K >> 4)), db + (uint64_t)(k * (UMMA_K >> 4)), idesc, first ? 0u : 1u);
                                first = 0;
                            }
                            umma_commit(&empty_bar[s]);
                        }
                    umma_commit(&acc_full[as]);
                }
            }
        }
    } else {
        const int e = warp - 2;                         // 0..7
        const int quad = warp & 3;                      // TMEM lane quadrant this warp may read
        const int half = e >> 2;                        // which 64 of the 128 accumulator columns
        uint32_t lvl_i = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int m0 = (tile % g.tiles_m) * BM, n0 = (tile / g.tiles_m) * BN;
            double acc[64];
#pragma unroll
            for (int j = 0; j < 64; j++) acc[j] = 0.0;
            for (int L = S + 1; L >= 2; L--, lvl_i++) {
                const int as = lvl_i % ACC_STAGES;
                const double w = scalbn(1.0, -DIGIT_BITS * L);
                mbar_wait(&acc_full[as], (lvl_i / ACC_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + half * 64 + c * 32), v);
#pragma unroll
                    for (int j = 0; j < 32; j++) acc[c * 32 + j] = fma((double)(int32_t)v[j], w, acc[c * 32 + j]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[as]);
            }
            const int row = m0 + quad * 32 + lane;
            if (row < g.M) {
                const int er = g.ea[row];
                double* dst = g.C + (size_t)row * g.N + n0 + half * 64;
#pragma unroll
                for (int j = 0; j < 64; j++) {
                    const int col = n0 + half * 64 + j;
                    if (col < g.N) dst[j] = scalbn(acc[j], er + g.eb[col]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// ---- host ------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_fn() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p) { fprintf(stderr, "cuTensorMapEncodeTiled unavailable\n"); exit(1); }
    return (PFN_encodeTiled)p;
}
static CUtensorMap make_map(const int8_t* base, long rows, long Kp, int box_rows) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)Kp};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed (%d)\n", (int)r); exit(1); }
    return tm;
}

struct Sliced { int8_t* planes = nullptr; int* expo = nullptr; long rows_pad = 0, Kp = 0; };
static Sliced slice(const double* dX, int rows, int K, int S, int tile_rows, float* ms_out) {
    Sliced s;
    s.rows_pad = ((long)rows + tile_rows - 1) / tile_rows * tile_rows;
    s.Kp = ((long)K + 15) / 16 * 16;
    CK(cudaMalloc(&s.planes, (size_t)S * s.rows_pad * s.Kp));
    CK(cudaMemset(s.planes, 0, (size_t)S * s.rows_pad * s.Kp));
    CK(cudaMalloc(&s.expo, sizeof(int) * rows));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    slice_rows_kernel<<<(rows * 32 + 255) / 256, 256>>>(dX, K, rows, K, S, s.planes, s.rows_pad, s.Kp, s.expo);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    if (ms_out) CK(cudaEventElapsedTime(ms_out, e0, e1));
    return s;
}

static void run(int M, int N, int K, int S, bool check, int reps) {
    std::vector<double> hA((size_t)M * K), hB((size_t)N * K);
    uint64_t st = 88172645463325252ull + (uint64_t)M * 131 + N * 7 + K;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) * (1.0 / 9007199254740992.0); };
    auto gauss = [&]() { const double u = rnd() + 1e-300, v = rnd(); return std::sqrt(-2.0 * std::log(u)) * std::cos(6.283185307179586 * v); };
    for (size_t i = 0; i < hA.size(); i++) hA[i] = gauss() * std::exp(2.0 * gauss());     // wide dynamic range inside rows
    for (size_t i = 0; i < hB.size(); i++) hB[i] = gauss() * (1.0 + 100.0 * rnd());
    double *dA, *dB, *dC;
    CK(cudaMalloc(&dA, sizeof(double) * hA.size())); CK(cudaMalloc(&dB, sizeof(double) * hB.size()));
    CK(cudaMalloc(&dC, sizeof(double) * (size_t)M * N));
    CK(cudaMemcpy(dA, hA.data(), sizeof(double) * hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), sizeof(double) * hB.size(), cudaMemcpyHostToDevice));
    float msA = 0, msB = 0;
    Sliced sa = slice(dA, M, K, S, BM, &msA), sb = slice(dB, N, K, S, BN, &msB);
    CUtensorMap tmA = make_map(sa.planes, (long)S * sa.rows_pad, sa.Kp, BM), tmB = make_map(sb.planes, (long)S * sb.rows_pad, sb.Kp, BN);
    Args g{M, N, K, S, sa.rows_pad, sb.rows_pad, dC, sa.expo, sb.expo, (M + BM - 1) / BM, (N + BN - 1) / BN};
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = std::min(g.tiles_m * g.tiles_n, sms);
    CK(cudaFuncSetAttribute(i8_sliced_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    i8_sliced_gemm_kernel<<<grid, THREADS, SMEM_BYTES>>>(tmA, tmB, g);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    if (check) {
        std::vector<double> hC((size_t)M * N);
        CK(cudaMemcpy(hC.data(), dC, sizeof(double) * hC.size(), cudaMemcpyDeviceToHost));
        double worst_b = 0, worst_c = 0, cmax = 0;
        for (int i = 0; i < M; i++)
            for (int j = 0; j < N; j++) {
                long double acc = 0, bound = 0;
                for (int k = 0; k < K; k++) {
                    acc += (long double)hA[(size_t)i * K + k] * hB[(size_t)j * K + k];
                    bound += fabsl((long double)hA[(size_t)i * K + k] * hB[(size_t)j * K + k]);
                }
                const double err = (double)fabsl((long double)hC[(size_t)i * N + j] - acc);
                worst_b = std::max(worst_b, err / (double)bound);
                worst_c = std::max(worst_c, err);
                cmax = std::max(cmax, (double)fabsl(acc));
            }
        printf("check %d x %d x %d, S = %d: err/|A||B| = %.2e, err/max|C| = %.2e  (prototype: ~2 decimal digits per slice; S = 8 -> ~1e-14)\n",
               M, N, K, S, worst_b, worst_c / cmax);
    }
    if (reps > 0) {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; r++) i8_sliced_gemm_kernel<<<grid, THREADS, SMEM_BYTES>>>(tmA, tmB, g);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float t; CK(cudaEventElapsedTime(&t, e0, e1));
        const double ms = t / reps, pairs = 0.5 * S * (S + 1);
        printf("%6d x %6d x %6d, S = %d: gemm %.3f ms = %.1f TFLOP/s FP64-equivalent (%.0f TOP/s int8 over %g digit products); slicing A %.3f ms, B %.3f ms\n",
               M, N, K, S, ms, 2.0 * M * N * K / ms * 1e-9, 2.0 * M * N * K * pairs / ms * 1e-9, pairs, msA, msB);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(sa.planes); cudaFree(sb.planes); cudaFree(sa.expo); cudaFree(sb.expo);
}

int main(int argc, char** argv) {
    const int S = argc > 1 ? std::atoi(argv[1]) : 8;
    if (S < 2 || S > MAX_S) { fprintf(stderr, "S must be in 2..%d\n", MAX_S); return 1; }
    run(256, 256, 384, S, true, 0);
    run(200, 300, 272, S, true, 0);            // ragged M, N, K
    run(5120, 20480, 256, S, false, 10);       // exponential-sum trait product
    run(128, 20480, 5120, S, false, 10);       // thin k x p product over K = n
    run(4096, 4096, 4096, S, false, 3);
    return 0;
}
