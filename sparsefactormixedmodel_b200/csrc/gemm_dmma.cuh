// FP64 GEMM core for sm_100a: DMMA.8x8x4 tiles fed by TMA through an mbarrier ring.
//
//   C[M x N] = alpha * A * B + beta * Cin,   A given as At (K x M, column-major, lda),
//                                            B given as Bm (K x N, column-major, ldb)
// i.e. the BLAS "TN" form: BOTH operands are contiguous along K.  Every contraction of the BSFG
// sweep is arranged to have this form (DESIGN.md "layouts"), so one kernel serves all of them:
// U'Y, U'F (BSFG_functions_c.cpp:107-108), the weighted Gram product behind Qlam (:120-122),
// Design_U'Y (:65), Y_tilde*Lmsg and Lambda'Lmsg (:149-152), and the F*Lambda' cross terms.
//
// Tile = BM x BN x 16 doubles per stage.  A stage's operand tile is [rows][16 doubles] = 128-byte
// rows written by ONE cp.async.bulk.tensor.2d (SWIZZLE_128B).  The 4 k-indices one DMMA consumes
// are {2j, 2j+1, 2j+8, 2j+9} (j = 0..3) rather than {4j..4j+3}: summation order over k is free, and
// with this choice the 16 lanes of a half-warp (4 rows x 4 k) hit 16 distinct 8-byte bank pairs
// under the hardware swizzle (chunk' = chunk ^ (row & 7)), so fragment loads are conflict-free
// LDS.64 with no padding and no repacking.
//
// The kernel is persistent: at most one CTA per SM walks the (tile, split) work items with a fixed stride, and the TMA
// ring runs across items, so an item's epilogue stores overlap the first loads of the next one.
//
// Warp roles: 8 warps, each owning a 64 x 32 sub-tile of the 128 x 128 CTA tile (8 x 4 DMMA
// accumulators = 64 FP64 values = 128 registers per thread).  Thread 0 also issues the TMA loads,
// STAGES-1 k-tiles ahead (a 9th producer warp would round the CTA up to 384 threads for register
// allocation and cap the kernel at 168 registers -> spills in the DMMA loop).  Split-K
// writes raw partial tiles to a workspace that `splitk_reduce_kernel` sums in a FIXED order, so
// results are bit-reproducible run to run and rank to rank (no FP64 atomics).
#pragma once
#include "common.cuh"

namespace bsfg {

constexpr int GEMM_BK = 16;          // doubles per stage row = 128 bytes
constexpr int GEMM_CONSUMERS = 8;    // consumer warps
constexpr int GEMM_THREADS = GEMM_CONSUMERS * 32;   // no dedicated producer warp: 256 threads keep the 255-register budget

struct GemmArgs {
    int M, N, K;
    double* C;
    long ldc;
    const double* Cin;     // may be null (beta ignored) or alias C
    long ldcin;
    double alpha, beta;
    int splits;            // >= 1
    int ktiles_per_split;  // stream-K launches (dgemm_tn_streamk_kernel): k-tiles per CTA run, splits = 1
    double* ws;            // [splits][N][M] raw partials when splits > 1
    int tiles_m, tiles_n;
    int group_n;           // rasterisation: tiles are walked m-fastest inside groups of `group_n` n-tiles
    const int* run_if;     // optional device flag: the whole launch is a no-op when *run_if == 0
    int trans_out;         // 1: the product is stored transposed, element (m, n) at C[m * ldc + n] (and Cin likewise);
                           // lets an N = k product run as the M = k product of the swapped operands (short row tile)
    int* tile_counter;     // != null: split-K partials are summed by the LAST CTA to finish a tile (fixed split order, so the
                           // result is the same as splitk_reduce_kernel's) -- no separate reduce launch; one self-resetting
                           // counter per output tile
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

template <int BM, int BN, int STAGES>
struct GemmSmem {
    static constexpr int A_BYTES = BM * GEMM_BK * 8;
    static constexpr int B_BYTES = BN * GEMM_BK * 8;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;  // + barriers + align slack
};

// ---- the kernel -----------------------------------------------------------------------------
// WM = warps along M (the other 8/WM along N).  WM = 2: (BM/2) x 32 warp tiles, the only layout instantiated (gemm_host.inl
// records why the all-rows / all-columns layouts for M = k or N = k products are not used).  BM < 128 serves products
// whose M is a single short row tile (M = k factors): BM = 112 for k = 100 multiplies 12% of zero fill instead of 28%.
template <int BM, int BN, int STAGES, int WM>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
dgemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs g) {
    static_assert(BN == 128 && BM % (8 * WM) == 0 && BM >= 16 && BM <= 128, "128-column CTA tiles of 16..128 rows");
    static_assert(WM == 1 || WM == 2 || WM == 8, "supported warp layouts");
    constexpr int WN = GEMM_CONSUMERS / WM;
    constexpr int FM = (BM / 8) / WM;      // 8-row fragments per warp
    constexpr int FN = (BN / 8) / WN;      // 8-column fragments per warp
    using S = GemmSmem<BM, BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    if (g.run_if && *g.run_if == 0) return;        // uniform across the grid: a conditional launch (exact-Gram fallback)
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = (uint64_t*)(smem + STAGES * S::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // Work items = (output tile, K split); a CTA walks items blockIdx.x, blockIdx.x + gridDim.x, ... (persistent when the
    // grid is smaller than the item count: one CTA per SM).  Tiles are numbered with a grouped rasterisation so the CTAs
    // that run side by side share operand panels in L2.
    const int tiles = g.tiles_m * g.tiles_n;
    const int work_total = tiles * g.splits;
    const int total_ktiles = (g.K + GEMM_BK - 1) / GEMM_BK;
    struct Item { int m0, n0, kt_begin, nkt, split, tile; };
    auto decode = [&](int w) {
        Item it;
        const int tile = w % tiles;
        it.tile = tile;
        it.split = w / tiles;
        const int group_sz = g.group_n * g.tiles_m;
        const int grp = tile / group_sz;
        const int first_n = grp * g.group_n;
        const int gn = min(g.group_n, g.tiles_n - first_n);
        const int in_grp = tile - grp * group_sz;
        it.n0 = (first_n + in_grp % gn) * BN;
        it.m0 = (in_grp / gn) * BM;
        it.kt_begin = it.split * g.ktiles_per_split;
        it.nkt = max(0, min(total_ktiles, it.kt_begin + g.ktiles_per_split) - it.kt_begin);
        return it;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], GEMM_CONSUMERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ===================== DMMA consumers =====================
    const int gq = lane >> 2;      // row inside an 8x8 fragment (0..7)
    const int q = lane & 3;        // k slot (0..3)
    const int wm0 = (warp % WM) * (FM * 8);
    const int wn0 = (warp / WM) * (FN * 8);

    // byte offset inside a 128-byte row for k-step j: chunk = j + 4*(q>>1), swizzled by (row&7)==gq
    uint32_t koff[4];
#pragma unroll
    for (int j = 0; j < 4; j++) koff[j] = (uint32_t)((((j + 4 * (q >> 1)) ^ gq) << 4) + ((q & 1) << 3));
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t a_row = (uint32_t)((wm0 + gq) * 128);
    const uint32_t b_row = (uint32_t)(S::A_BYTES + (wn0 + gq) * 128);

    // TMA issue is folded into consumer warp 0 (lane 0): it runs STAGES-1 k-tiles ahead of the math, ACROSS work items --
    // the first k-tiles of the next item are requested before this item's epilogue, so its stores overlap their flight.
    // ring_put / ring_get count k-tiles over the CTA's whole life; slot = count % STAGES, phase = (count / STAGES) & 1.
    uint32_t ring_put = 0, ring_get = 0;
    auto issue = [&](const Item& w, int kt) {
        const int s2 = ring_put % STAGES;
        mbar_wait(&empty_bar[s2], ((ring_put / STAGES) & 1) ^ 1);   // passes immediately on first use of a slot
        mbar_expect_tx(&full_bar[s2], S::STAGE_BYTES);
        uint8_t* sa = smem + s2 * S::STAGE_BYTES;
        const int kc = (w.kt_begin + kt) * GEMM_BK;
        tma_load_2d(sa, &tmA, &full_bar[s2], kc, w.m0);
        tma_load_2d(sa + S::A_BYTES, &tmB, &full_bar[s2], kc, w.n0);
        ring_put++;
    };
    int w = blockIdx.x;
    if (w >= work_total) return;
    Item cur = decode(w);
    int issued = 0;                // k-tiles of `cur` already requested (thread 0 only)
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (; issued < STAGES - 1 && issued < cur.nkt; issued++) issue(cur, issued);
    }
    __syncwarp();

    for (;;) {
        // the tile's Cin lines are asked into L2 now: the epilogue of a K = k product (seven k-tiles per tile) otherwise waits
        // four times for HBM while the DMMA pipe idles
        if (g.Cin && g.splits == 1 && !g.trans_out) {
            for (int t = threadIdx.x; t < BN * (BM / 16); t += GEMM_THREADS) {
                const int n = cur.n0 + t / (BM / 16), m = cur.m0 + (t % (BM / 16)) * 16;
                if (n < g.N && m < g.M) asm volatile("prefetch.global.L2 [%0];" ::"l"(g.Cin + (size_t)n * g.ldcin + m));
            }
        }
        double acc[FM][FN][2];
#pragma unroll
        for (int i = 0; i < FM; i++)
#pragma unroll
            for (int j = 0; j < FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int it = 0; it < cur.nkt; it++) {
            const int s = ring_get % STAGES;
            const uint32_t ph = (ring_get / STAGES) & 1;
            if (threadIdx.x == 0 && issued < cur.nkt) { issue(cur, issued); issued++; }
            __syncwarp();
            mbar_wait(&full_bar[s], ph);
            const uint32_t sbase = smem_base + s * S::STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double a[FM], b[FN];
#pragma unroll
                for (int i = 0; i < FM; i++) a[i] = lds_f64(sbase + a_row + i * 8 * 128 + koff[j]);
#pragma unroll
                for (int i = 0; i < FN; i++) b[i] = lds_f64(sbase + b_row + i * 8 * 128 + koff[j]);
                // fragments beyond the matrix edge multiply TMA zero fill: skipping them (predicated DMMAs) was measured
                // slower than issuing them, because the other warps of the CTA do the full work anyway
#pragma unroll
                for (int i = 0; i < FM; i++)
#pragma unroll
                    for (int jj = 0; jj < FN; jj++) dmma884(acc[i][jj][0], acc[i][jj][1], a[i], b[jj]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            ring_get++;
        }

        // next item: request its first k-tiles now (the slots free up as the other warps leave the loop above)
        const int wn = w + (int)gridDim.x;
        const bool more = wn < work_total;
        if (threadIdx.x == 0) {
            issued = 0;
            if (more) {
                const Item nxt = decode(wn);
                for (; issued < STAGES - 1 && issued < nxt.nkt; issued++) issue(nxt, issued);
            }
        }
        __syncwarp();

        // ===================== epilogue =====================
        // thread owns C[m0+wm0+8i+gq][n0+wn0+8j+2q+{0,1}]
        const int m0 = cur.m0, n0 = cur.n0;
        if (g.splits > 1) {
            double* ws = g.ws + (size_t)cur.split * (size_t)g.M * (size_t)g.N;
#pragma unroll
            for (int i = 0; i < FM; i++) {
                const int m = m0 + wm0 + 8 * i + gq;
                if (m >= g.M) continue;
#pragma unroll
                for (int j = 0; j < FN; j++) {
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const int n = n0 + wn0 + 8 * j + 2 * q + c;
                        if (n < g.N) ws[(size_t)n * g.M + m] = acc[i][j][c];
                    }
                }
            }
            if (g.tile_counter) {
                // the last CTA to deliver a partial of this tile sums all of them, in split order
                __shared__ int s_last;
                __threadfence();
                __syncthreads();
                if (threadIdx.x == 0) {
                    const int t = atomicAdd(&g.tile_counter[cur.tile], 1);
                    s_last = (t == g.splits - 1);
                    if (s_last) g.tile_counter[cur.tile] = 0;
                }
                __syncthreads();
                if (s_last) {
                    __threadfence();
                    const size_t total = (size_t)g.M * (size_t)g.N;
                    const int rows = min(BM, g.M - m0), cols = min(BN, g.N - n0);
                    for (int idx = threadIdx.x; idx < rows * cols; idx += GEMM_THREADS) {
                        const int m = m0 + idx % rows, n = n0 + idx / rows;
                        const size_t e = (size_t)n * g.M + m;
                        double a = 0.0;
                        for (int sp = 0; sp < g.splits; sp++) a += __ldcg(g.ws + (size_t)sp * total + e);
                        double v = g.alpha * a;
                        if (g.Cin) v += g.beta * g.Cin[g.trans_out ? (size_t)m * g.ldcin + n : (size_t)n * g.ldcin + m];
                        g.C[g.trans_out ? (size_t)m * g.ldc + n : (size_t)n * g.ldc + m] = v;
                    }
                }
            }
        } else {
            // beta != 0: the 16 Cin values of one 8-column fragment strip are loaded as one batch before any of them is
            // used, so 16 independent 8-byte loads per thread are in flight (a load->FMA->store chain per element made the
            // K = k products, which are all epilogue, 2.5x slower).
#pragma unroll
            for (int j = 0; j < FN; j++) {
                double cin[FM][2];
                if (g.Cin) {
#pragma unroll
                    for (int i = 0; i < FM; i++) {
#pragma unroll
                        for (int c = 0; c < 2; c++) {
                            const int m = m0 + wm0 + 8 * i + gq;
                            const int n = n0 + wn0 + 8 * j + 2 * q + c;
                            cin[i][c] = (m < g.M && n < g.N)
                                            ? g.Cin[g.trans_out ? (size_t)m * g.ldcin + n : (size_t)n * g.ldcin + m] : 0.0;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < FM; i++) {
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const int m = m0 + wm0 + 8 * i + gq;
                        const int n = n0 + wn0 + 8 * j + 2 * q + c;
                        if (m < g.M && n < g.N) {
                            double v = g.alpha * acc[i][j][c];
                            if (g.Cin) v += g.beta * cin[i][c];
                            g.C[g.trans_out ? (size_t)m * g.ldc + n : (size_t)n * g.ldc + m] = v;
                        }
                    }
                }
            }
        }
        if (!more) break;
        w = wn;
        cur = decode(w);       // re-derived after the epilogue rather than kept live across it (register budget)
    }
}

// The stream-K twin of dgemm_tn_kernel: same tiles, ring and DMMA loop; a CTA's work is one contiguous run of the flattened
// (tile, k-tile) space instead of a list of (tile, split) items (g.ktiles_per_split = k-tiles per run, g.splits = 1).  Kept
// as a separate kernel with NO new GemmArgs members: three more parameter words changed ptxas's register allocation of the
// 128 x 128 tiles of dgemm_tn_kernel (250 -> 222 registers) and cost them 6% (trait product 1.59 -> 1.70 ms).
template <int BM, int BN, int STAGES, int WM>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
dgemm_tn_streamk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs g) {
    static_assert(BN == 128 && BM % (8 * WM) == 0 && BM >= 16 && BM <= 128, "128-column CTA tiles of 16..128 rows");
    static_assert(WM == 1 || WM == 2 || WM == 8, "supported warp layouts");
    constexpr bool SK = true;
    constexpr int WN = GEMM_CONSUMERS / WM;
    constexpr int FM = (BM / 8) / WM;      // 8-row fragments per warp
    constexpr int FN = (BN / 8) / WN;      // 8-column fragments per warp
    using S = GemmSmem<BM, BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    if (g.run_if && *g.run_if == 0) return;        // uniform across the grid: a conditional launch (exact-Gram fallback)
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = (uint64_t*)(smem + STAGES * S::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // Work items = (output tile, K split); a CTA walks items blockIdx.x, blockIdx.x + gridDim.x, ... (persistent when the
    // grid is smaller than the item count: one CTA per SM).  Tiles are numbered with a grouped rasterisation so the CTAs
    // that run side by side share operand panels in L2.
    const int tiles = g.tiles_m * g.tiles_n;
    const int work_total = tiles * g.splits;
    const int total_ktiles = (g.K + GEMM_BK - 1) / GEMM_BK;
    struct Item { int m0, n0, kt_begin, nkt, split, tile; };
    auto decode_tiles = [&](int w) {
        Item it;
        const int tile = w % tiles;
        it.tile = tile;
        it.split = w / tiles;
        const int group_sz = g.group_n * g.tiles_m;
        const int grp = tile / group_sz;
        const int first_n = grp * g.group_n;
        const int gn = min(g.group_n, g.tiles_n - first_n);
        const int in_grp = tile - grp * group_sz;
        it.n0 = (first_n + in_grp % gn) * BN;
        it.m0 = (in_grp / gn) * BM;
        it.kt_begin = it.split * g.ktiles_per_split;
        it.nkt = max(0, min(total_ktiles, it.kt_begin + g.ktiles_per_split) - it.kt_begin);
        return it;
    };
    // stream-K: w = index of the run's piece inside this CTA; pieces are consecutive tiles
    auto decode_run = [&](int w) {
        Item it;
        const long lo = (long)blockIdx.x * g.ktiles_per_split, hi = min(lo + (long)g.ktiles_per_split, (long)tiles * total_ktiles);
        const int tile = (int)(lo / total_ktiles) + w;
        const long tb = (long)tile * total_ktiles;
        it.kt_begin = (w == 0) ? (int)(lo - tb) : 0;
        it.nkt = (int)min((long)total_ktiles, hi - tb) - it.kt_begin;          // <= 0: past the end of the run
        it.split = (int)blockIdx.x - (int)(tb / g.ktiles_per_split);                  // workspace slot of a partial piece
        it.tile = tile; it.m0 = 0; it.n0 = 0;
        if (it.nkt <= 0 || tile >= tiles) { it.nkt = 0; return it; }
        const int group_sz = g.group_n * g.tiles_m;
        const int grp = tile / group_sz;
        const int first_n = grp * g.group_n;
        const int gn = min(g.group_n, g.tiles_n - first_n);
        const int in_grp = tile - grp * group_sz;
        it.n0 = (first_n + in_grp % gn) * BN;
        it.m0 = (in_grp / gn) * BM;
        return it;
    };
    auto decode = [&](int w) {
        if constexpr (SK) return decode_run(w);
        else return decode_tiles(w);
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], GEMM_CONSUMERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ===================== DMMA consumers =====================
    const int gq = lane >> 2;      // row inside an 8x8 fragment (0..7)
    const int q = lane & 3;        // k slot (0..3)
    const int wm0 = (warp % WM) * (FM * 8);
    const int wn0 = (warp / WM) * (FN * 8);

    // byte offset inside a 128-byte row for k-step j: chunk = j + 4*(q>>1), swizzled by (row&7)==gq
    uint32_t koff[4];
#pragma unroll
    for (int j = 0; j < 4; j++) koff[j] = (uint32_t)((((j + 4 * (q >> 1)) ^ gq) << 4) + ((q & 1) << 3));
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t a_row = (uint32_t)((wm0 + gq) * 128);
    const uint32_t b_row = (uint32_t)(S::A_BYTES + (wn0 + gq) * 128);

    // TMA issue is folded into consumer warp 0 (lane 0): it runs STAGES-1 k-tiles ahead of the math, ACROSS work items --
    // the first k-tiles of the next item are requested before this item's epilogue, so its stores overlap their flight.
    // ring_put / ring_get count k-tiles over the CTA's whole life; slot = count % STAGES, phase = (count / STAGES) & 1.
    uint32_t ring_put = 0, ring_get = 0;
    auto issue = [&](const Item& w, int kt) {
        const int s2 = ring_put % STAGES;
        mbar_wait(&empty_bar[s2], ((ring_put / STAGES) & 1) ^ 1);   // passes immediately on first use of a slot
        mbar_expect_tx(&full_bar[s2], S::STAGE_BYTES);
        uint8_t* sa = smem + s2 * S::STAGE_BYTES;
        const int kc = (w.kt_begin + kt) * GEMM_BK;
        tma_load_2d(sa, &tmA, &full_bar[s2], kc, w.m0);
        tma_load_2d(sa + S::A_BYTES, &tmB, &full_bar[s2], kc, w.n0);
        ring_put++;
    };
    int w = blockIdx.x;
    if constexpr (SK) w = 0;
    else if (w >= work_total) return;
    Item cur = decode(w);
    if constexpr (SK) { if (cur.nkt <= 0) return; }
    int issued = 0;                // k-tiles of `cur` already requested (thread 0 only)
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (; issued < STAGES - 1 && issued < cur.nkt; issued++) issue(cur, issued);
    }
    __syncwarp();

    for (;;) {
        double acc[FM][FN][2];
#pragma unroll
        for (int i = 0; i < FM; i++)
#pragma unroll
            for (int j = 0; j < FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int it = 0; it < cur.nkt; it++) {
            const int s = ring_get % STAGES;
            const uint32_t ph = (ring_get / STAGES) & 1;
            if (threadIdx.x == 0 && issued < cur.nkt) { issue(cur, issued); issued++; }
            __syncwarp();
            mbar_wait(&full_bar[s], ph);
            const uint32_t sbase = smem_base + s * S::STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double a[FM], b[FN];
#pragma unroll
                for (int i = 0; i < FM; i++) a[i] = lds_f64(sbase + a_row + i * 8 * 128 + koff[j]);
#pragma unroll
                for (int i = 0; i < FN; i++) b[i] = lds_f64(sbase + b_row + i * 8 * 128 + koff[j]);
                // fragments beyond the matrix edge multiply TMA zero fill: skipping them (predicated DMMAs) was measured
                // slower than issuing them, because the other warps of the CTA do the full work anyway
#pragma unroll
                for (int i = 0; i < FM; i++)
#pragma unroll
                    for (int jj = 0; jj < FN; jj++) dmma884(acc[i][jj][0], acc[i][jj][1], a[i], b[jj]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            ring_get++;
        }

        // next item: request its first k-tiles now (the slots free up as the other warps leave the loop above)
        const int wn = SK ? w + 1 : w + (int)gridDim.x;
        const bool more = SK ? decode_run(wn).nkt > 0 : wn < work_total;
        if (threadIdx.x == 0) {
            issued = 0;
            if (more) {
                const Item nxt = decode(wn);
                for (; issued < STAGES - 1 && issued < nxt.nkt; issued++) issue(nxt, issued);
            }
        }
        __syncwarp();

        // ===================== epilogue =====================
        // thread owns C[m0+wm0+8i+gq][n0+wn0+8j+2q+{0,1}]
        const int m0 = cur.m0, n0 = cur.n0;
        if (SK ? cur.nkt < total_ktiles : g.splits > 1) {
            double* ws = g.ws + (size_t)cur.split * (size_t)g.M * (size_t)g.N;
#pragma unroll
            for (int i = 0; i < FM; i++) {
                const int m = m0 + wm0 + 8 * i + gq;
                if (m >= g.M) continue;
#pragma unroll
                for (int j = 0; j < FN; j++) {
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const int n = n0 + wn0 + 8 * j + 2 * q + c;
                        if (n < g.N) ws[(size_t)n * g.M + m] = acc[i][j][c];
                    }
                }
            }
            if (g.tile_counter) {
                // the last CTA to deliver a partial of this tile sums all of them, in split order
                __shared__ int s_last;
                __threadfence();
                __syncthreads();
                if (threadIdx.x == 0) {
                    const int t = atomicAdd(&g.tile_counter[cur.tile], 1);
                    s_last = (t == g.splits - 1);
                    if (s_last) g.tile_counter[cur.tile] = 0;
                }
                __syncthreads();
                if (s_last) {
                    __threadfence();
                    const size_t total = (size_t)g.M * (size_t)g.N;
                    const int rows = min(BM, g.M - m0), cols = min(BN, g.N - n0);
                    for (int idx = threadIdx.x; idx < rows * cols; idx += GEMM_THREADS) {
                        const int m = m0 + idx % rows, n = n0 + idx / rows;
                        const size_t e = (size_t)n * g.M + m;
                        double a = 0.0;
                        for (int sp = 0; sp < g.splits; sp++) a += __ldcg(g.ws + (size_t)sp * total + e);
                        double v = g.alpha * a;
                        if (g.Cin) v += g.beta * g.Cin[g.trans_out ? (size_t)m * g.ldcin + n : (size_t)n * g.ldcin + m];
                        g.C[g.trans_out ? (size_t)m * g.ldc + n : (size_t)n * g.ldc + m] = v;
                    }
                }
            }
        } else {
            // beta != 0: the 16 Cin values of one 8-column fragment strip are loaded as one batch before any of them is
            // used, so 16 independent 8-byte loads per thread are in flight (a load->FMA->store chain per element made the
            // K = k products, which are all epilogue, 2.5x slower).
#pragma unroll
            for (int j = 0; j < FN; j++) {
                double cin[FM][2];
                if (g.Cin) {
#pragma unroll
                    for (int i = 0; i < FM; i++) {
#pragma unroll
                        for (int c = 0; c < 2; c++) {
                            const int m = m0 + wm0 + 8 * i + gq;
                            const int n = n0 + wn0 + 8 * j + 2 * q + c;
                            cin[i][c] = (m < g.M && n < g.N)
                                            ? g.Cin[g.trans_out ? (size_t)m * g.ldcin + n : (size_t)n * g.ldcin + m] : 0.0;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < FM; i++) {
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const int m = m0 + wm0 + 8 * i + gq;
                        const int n = n0 + wn0 + 8 * j + 2 * q + c;
                        if (m < g.M && n < g.N) {
                            double v = g.alpha * acc[i][j][c];
                            if (g.Cin) v += g.beta * cin[i][c];
                            g.C[g.trans_out ? (size_t)m * g.ldc + n : (size_t)n * g.ldc + m] = v;
                        }
                    }
                }
            }
        }
        if (!more) break;
        w = wn;
        cur = decode(w);       // re-derived after the epilogue rather than kept live across it (register budget)
    }
}

__global__ void splitk_reduce_kernel(GemmArgs g) {
    if (g.run_if && *g.run_if == 0) return;
    const size_t total = (size_t)g.M * g.N;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(idx % g.M);
        const int n = (int)(idx / g.M);
        double acc = 0.0;
        for (int s = 0; s < g.splits; s++) acc += g.ws[(size_t)s * total + idx];   // fixed order
        double v = g.alpha * acc;
        if (g.Cin) v += g.beta * g.Cin[g.trans_out ? (size_t)m * g.ldcin + n : (size_t)n * g.ldcin + m];
        g.C[g.trans_out ? (size_t)m * g.ldc + n : (size_t)n * g.ldc + m] = v;
    }
}

// Sums the partial pieces of the tiles a stream-K launch did not finish in one piece: element (m, n) belongs to tile
// t(m, n) (same rasterisation as the kernel); the tile's pieces were produced by CTAs first..last = floor(t KT / P) ..
// floor(((t + 1) KT - 1) / P) and sit in workspace slots 0 .. last - first; one piece means the tile was stored directly.
// Fixed summation order (CTA order), so results are reproducible.
template <int BN>
__global__ void streamk_reduce_kernel(GemmArgs g, int BM) {
    if (g.run_if && *g.run_if == 0) return;
    const size_t total = (size_t)g.M * g.N;
    const int KT = (g.K + GEMM_BK - 1) / GEMM_BK;
    const int group_sz = g.group_n * g.tiles_m;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(idx % g.M), n = (int)(idx / g.M);
        const int mt = m / BM, nt = n / BN;
        const int grp = nt / g.group_n, first_n = grp * g.group_n;
        const int gn = min(g.group_n, g.tiles_n - first_n);
        const long tile = (long)grp * group_sz + (long)mt * gn + (nt - first_n);
        const int first = (int)(tile * KT / g.ktiles_per_split), last = (int)(((tile + 1) * KT - 1) / g.ktiles_per_split);
        if (last == first) continue;
        double acc = 0.0;
        for (int sp = 0; sp <= last - first; sp++) acc += g.ws[(size_t)sp * total + idx];
        double v = g.alpha * acc;
        if (g.Cin) v += g.beta * g.Cin[g.trans_out ? (size_t)m * g.ldcin + n : (size_t)n * g.ldcin + m];
        g.C[g.trans_out ? (size_t)m * g.ldc + n : (size_t)n * g.ldc + m] = v;
    }
}

// ---- host side ----------------------------------------------------------------------------
struct GemmPlan;   // opaque to callers; see gemm_host.cu

// C = alpha * At' * Bm + beta * Cin.  At: K x M (lda), Bm: K x N (ldb); all column-major, lda/ldb
// even, pointers 16-byte aligned.  `ws`/`ws_doubles` is scratch for split-K (may be null: no split).
void dgemm_tn(cudaStream_t st, int sm_count, int M, int N, int K, double alpha, const double* At, long lda,
              const double* Bm, long ldb, double beta, const double* Cin, long ldcin, double* C, long ldc,
              double* ws, size_t ws_doubles, const int* run_if = nullptr);

}  // namespace bsfg
