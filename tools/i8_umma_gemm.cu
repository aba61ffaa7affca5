// Stand-alone experiment, NOT part of the product.
// Run on a B200 in round 2: both self-checks exact, 73.8 / 441.8 / 1691.5 TOP/s on the three timing shapes (profiles/r02_i8_umma_gemm.log).
// Stand-alone microbenchmark of the building block that DESIGN.md section 9 plans for the
// sweep's GEMMs: an INT8 x INT8 -> INT32 GEMM on the tcgen05 tensor pipe with TMEM accumulators, i.e. one "slice
// product" of the INT8-slice emulation of FP64 (tools/ozaki_prototype.py).  It answers two questions before any
// integration work: (1) are the hand-built shared-memory / instruction descriptors right (self-check against a CPU
// product), (2) what INT8 rate does one B200 sustain on the sweep's shapes (K = 256 trait product, K = 5000 thin
// products), which fixes the slice budget.
//
//   C[M x N] (int32, row-major) = A[M x K] (int8, K contiguous) * B[N x K]' (int8, K contiguous)
//
// Structure (persistent, one CTA per SM, 192 threads):
//   warp 0     : TMA producer   -- cp.async.bulk.tensor.2d, SWIZZLE_128B, 4-stage mbarrier ring of (128 + 256) x 128 B
//   warp 1     : MMA issuer     -- one elected lane issues tcgen05.mma.cta_group::1.kind::i8 (M 128, N 256, K 32) x 4 per
//                                  stage; tcgen05.commit releases the stage / publishes the accumulator
//   warps 2..5 : epilogue       -- tcgen05.ld 32x32b (warp w reads TMEM lanes 32 (w % 4) ..), stores int32 rows
//   TMEM       : 512 columns = 2 accumulator stages of 128 lanes x 256 int32, so a tile's epilogue overlaps the next
//                tile's MMAs.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/i8_umma_gemm tools/i8_umma_gemm.cu -lcuda
// Run:    tools/i8_umma_gemm            (self-check on 256 x 512 x 384 and a ragged case, then timings)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int BM = 128, BN = 256, BK = 128;      // bytes == int8 elements along K per stage: one 128-byte swizzle row
constexpr int UMMA_K = 32;                       // kind::i8: 32 elements (256 bits) per instruction
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK, B_BYTES = BN * BK, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int ACC_STAGES = 2, TMEM_COLS = ACC_STAGES * BN;   // 512
constexpr int THREADS = 192;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* align */ + 256 /* barriers, tmem pointer */;

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
// K-major operand tile, rows of 128 bytes, SWIZZLE_128B: 8-row groups 1024 bytes apart (SBO), LBO = 1 (unused for a
// swizzled K-major layout), descriptor version 1, layout type 2 (cute/arch/mma_sm100_desc.hpp: SmemDescriptor).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                              // leading byte offset (>> 4), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset (>> 4), bits [32,46)
    d |= (uint64_t)1 << 46;                              // version = 1 (Blackwell), bits [46,48)
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B, bits [61,64)
    return d;
}
// kind::i8 instruction descriptor (InstrDescriptor): c_format S32 = 2 at [4,6); a/b format INT8 = 1 at [7,10), [10,13);
// K-major A and B (bits 15, 16 zero); N >> 3 at [17,23); M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_i8(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Args { int M, N, K; int32_t* C; int tiles_m, tiles_n; };

__global__ void __launch_bounds__(THREADS, 1)
i8_umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Args g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;          // MMA -> epilogue, one per accumulator stage
    uint64_t* acc_empty = acc_full + ACC_STAGES;      // epilogue -> MMA
    uint32_t* tmem_base_p = (uint32_t*)(acc_empty + ACC_STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = g.tiles_m * g.tiles_n;
    const int nkb = (g.K + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < ACC_STAGES; s++) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }   // 4 epilogue warps
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {      // one warp allocates all 512 TMEM columns (one CTA per SM) and gives up the permit
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_p)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_p;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t put = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int m0 = (t % g.tiles_m) * BM, n0 = (t / g.tiles_m) * BN;
                for (int kb = 0; kb < nkb; kb++, put++) {
                    const int s = put % STAGES;
                    mbar_wait(&empty_bar[s], ((put / STAGES) & 1) ^ 1);
                    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    uint8_t* sa = smem + s * STAGE_BYTES;
                    tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, m0);
                    tma_load_2d(sa + A_BYTES, &tmB, &full_bar[s], kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_i8(BM, BN);
            uint32_t get = 0, tile_i = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, tile_i++) {
                const int as = tile_i % ACC_STAGES;
                mbar_wait(&acc_empty[as], ((tile_i / ACC_STAGES) & 1) ^ 1);        // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);           // column offset in the low 16 bits
                for (int kb = 0; kb < nkb; kb++, get++) {
                    const int s = get % STAGES;
                    mbar_wait(&full_bar[s], (get / STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; k++)                           // advance 32 bytes along K: +2 in the address field
                        umma_i8(tmem_d, da + (uint64_t)(k * (UMMA_K >> 4)), db + (uint64_t)(k * (UMMA_K >> 4)), idesc, (kb | k) != 0);
                    umma_commit(&empty_bar[s]);                                     // stage reusable once these MMAs have read it
                }
                umma_commit(&acc_full[as]);                                         // accumulator complete
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> global =====================
        const int quad = warp & 3;                        // TMEM lanes this warp may touch: [32 quad, 32 quad + 32)
        uint32_t tile_i = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, tile_i++) {
            const int m0 = (t % g.tiles_m) * BM, n0 = (t / g.tiles_m) * BN;
            const int as = tile_i % ACC_STAGES;
            mbar_wait(&acc_full[as], (tile_i / ACC_STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = m0 + quad * 32 + lane;
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + c0);
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
                if (row < g.M) {
                    int32_t* dst = g.C + (size_t)row * g.N + n0 + c0;
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        if (n0 + c0 + j < g.N) dst[j] = (int32_t)v[j];
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// ---- host ---------------------------------------------------------------------------------------------------------
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
static CUtensorMap make_map(const int8_t* base, int rows, int K, int box_rows) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)K};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed (%d)\n", (int)r); exit(1); }
    return tm;
}

static double run(int M, int N, int K, bool check, int reps) {
    if (K % 16 != 0) { fprintf(stderr, "K must be a multiple of 16 (TMA row pitch)\n"); exit(1); }
    std::vector<int8_t> hA((size_t)M * K), hB((size_t)N * K);
    uint32_t st = 12345u + (uint32_t)(M * 31 + N * 7 + K);
    auto rnd = [&]() { st = st * 1664525u + 1013904223u; return (int8_t)((int)((st >> 24) & 0xFF) - 128); };
    for (auto& v : hA) v = rnd();
    for (auto& v : hB) v = rnd();
    int8_t *dA, *dB; int32_t* dC;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dC, sizeof(int32_t) * (size_t)M * N));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xFF, sizeof(int32_t) * (size_t)M * N));
    CUtensorMap tmA = make_map(dA, M, K, BM), tmB = make_map(dB, N, K, BN);
    Args g{M, N, K, dC, (M + BM - 1) / BM, (N + BN - 1) / BN};
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = std::min(g.tiles_m * g.tiles_n, sms);
    CK(cudaFuncSetAttribute(i8_umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    i8_umma_gemm_kernel<<<grid, THREADS, SMEM_BYTES>>>(tmA, tmB, g);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    if (check) {
        std::vector<int32_t> hC((size_t)M * N);
        CK(cudaMemcpy(hC.data(), dC, sizeof(int32_t) * hC.size(), cudaMemcpyDeviceToHost));
        long bad = 0;
        for (int i = 0; i < M; i++)
            for (int j = 0; j < N; j++) {
                int32_t acc = 0;
                for (int k = 0; k < K; k++) acc += (int32_t)hA[(size_t)i * K + k] * (int32_t)hB[(size_t)j * K + k];
                if (acc != hC[(size_t)i * N + j] && bad++ < 5) printf("  mismatch C[%d,%d] = %d, want %d\n", i, j, hC[(size_t)i * N + j], acc);
            }
        printf("check %d x %d x %d: %s (%ld mismatches)\n", M, N, K, bad ? "FAILED" : "ok", bad);
        if (bad) exit(2);
    }
    double ms = 0;
    if (reps > 0) {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; r++) i8_umma_gemm_kernel<<<grid, THREADS, SMEM_BYTES>>>(tmA, tmB, g);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float t; CK(cudaEventElapsedTime(&t, e0, e1)); ms = t / reps;
        printf("%6d x %6d x %6d: %.3f ms  %.1f TOP/s\n", M, N, K, ms, 2.0 * M * N * K / ms * 1e-9);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return ms;
}

int main() {
    run(256, 512, 384, true, 0);          // 2 x 2 tiles, 3 k-blocks
    run(200, 300, 272, true, 0);          // ragged M, N and K (TMA zero fill; partial tiles in the epilogue)
    run(5120, 20480, 256, false, 20);     // the exponential-sum trait product's shape (k = 100: 5050 pairs)
    run(128, 20480, 5120, false, 20);     // a thin k x p product over K = n
    run(8192, 8192, 8192, false, 5);      // square: the INT8 rate itself
    return 0;
}
