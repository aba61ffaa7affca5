// Batched k x k Cholesky + the three triangular solves of sample_Lambda_c (BSFG_functions_c.cpp:124-130)
// and sample_factors_scores_c (cpp:149), one CTA per system, as a LEFT-LOOKING BLOCKED factorisation on
// 8 x 8 tiles held in shared memory, with the trailing updates on the FP64 tensor pipe (DMMA.8x8x4).
//
//   Q_j   = unpack(Qp[:, j]) + diag(dadd_j)        (packed lower, pair(a,b) = a(a+1)/2 + b, a >= b)
//   L L'  = Q_j                                     (L = t(chol(Qlam)), cpp:124)
//   out_j = L'^-1 (L^-1 m_j + z_j)                 (= mlam + ylam, cpp:125-130)
//
// Layout in shared memory: the lower triangle of Q as NB(NB+1)/2 tiles (NB = ceil(k/8)), tile (I,J), I >= J,
// at index I(I+1)/2 + J, 64 doubles each, element (r,c) at r*8 + (c ^ ((r&2)<<1)).  The XOR makes the DMMA
// operand fragment loads (lane (g,q) reads element (g, q+4h)) hit 16 distinct 8-byte bank pairs per half warp.
// Rows/columns k..8NB-1 are identity padding.  The right-hand side m rides along as an extra matrix row k (with a
// huge diagonal), so the forward solve v = L^-1 m costs nothing: v' is row k of the factor.
//
// Round-2 structure (what the measurements of profiles/r02_chol_study.md led to).  The kernel is LATENCY bound: only four
// systems fit an SM (55 KB of tiles each), so its duration is (#systems / 4 / #SMs) x the critical path of one system,
// and that path is the chain of k scalar pivots (rsqrt 66 cycles + 2 dependent DFMA = 83 cycles each on B200) plus
// whatever else sits between two pivots.  A DMMA occupies the FP64 pipe of its SM sub-partition for 16 cycles, so pivot
// chains that share a sub-partition with DMMA-issuing warps ran 3x slower than alone.  Hence:
//   * ONE warp per CTA (the "diagonal warp", a different sub-partition from CTA to CTA) owns the pivot chain and does
//     nothing else on the critical path: per block column J it adds the last term L(J,J-1) L(J,J-1)' to the diagonal tile
//     (2 DMMA), factors the 8 x 8 tile in registers, and publishes L_JJ (with 1/diag) for the panel rows;
//   * the other warps ("panel warps") meanwhile run the tensor work of column J -- tile(I,J) -= sum_{m<J} L(I,m) L(J,m)'
//     for I > J and the partial sum (m < J) of the NEXT diagonal tile -- with exact per-warp tile counts (a predicated-off
//     DMMA still costs its 16 pipe cycles) and software-pipelined fragment loads; after the barrier they apply the
//     8-step forward substitution to their panel rows while the diagonal warp inverts L_JJ for the backward solve;
//   * the system is staged with cp.async (all 5050 8-byte copies in flight at once) instead of load -> store round trips;
//   * the backward solve L'x = v + z stays on one warp without block barriers: x_J = inv(L_JJ)' w_J is an 8 x 8
//     mat-vec (no 8-step dependency chain), the update of the earlier rows keeps four independent rows per lane in flight.
// Two block barriers per column.  The per-element map packed-index -> tile offset is a uint16 table built once per k.
#pragma once
#include "kernels.cuh"

namespace bsfg {

__host__ __device__ __forceinline__ int cb_tile_id(int I, int J) { return I * (I + 1) / 2 + J; }
__host__ __device__ __forceinline__ int cb_swz(int r, int c) { return r * 8 + (c ^ ((r & 2) << 1)); }

__device__ __forceinline__ void cb_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// 1/sqrt(x) to ~1 ulp without the special-case handling of CUDA's rsqrt(double): the hardware seed (MUFU.RSQ64H,
// relative error 2^-22) and ONE third-order step y <- y (1 + e/2 + 3 e^2/8), e = 1 - x y^2 (error (5/16) e^3 ~ 2^-67):
// four dependent FP64 operations after the seed.  x > 0 finite is the caller's business (a non-positive pivot is
// flagged before the result is used).
__device__ __forceinline__ double cb_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    const double p = fma(e, 0.375, 0.5);
    const double q = y * e;
    return fma(q, p, y);
}

__device__ __forceinline__ void cb_cp_async8(double* smem_dst, const double* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem_src) : "memory");
}
// block barrier usable from the two warp-specialised code paths (every warp of the CTA executes exactly one of them)
__device__ __forceinline__ void cb_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void cb_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// NT tiles of block column J owned by one panel warp: tile(I_t, J) -= sum_{m<J} L(I_t,m) L(J,m)', I_t = I0 + stride t.
// Accumulators live in the DMMA C fragment, two sets per tile (consecutive DMMAs of a tile are independent), operand
// fragments of step m+1 are loaded before the DMMAs of step m are issued.  B == A (diag = true) gives the partial sum of a
// diagonal tile: tile(I0, I0) -= sum_{m<J} L(I0,m) L(I0,m)'.
template <int NT, bool DIAG>
__device__ __forceinline__ void cb_update_tiles(double* __restrict__ T, int J, int I0, int stride, int fa0, int fa1, int fc) {
    const double* rowJ = T + cb_tile_id(DIAG ? I0 : J, 0) * 64;
    double* rowI[NT];
    double acc[NT][2], acd[NT][2], a0[NT], a1[NT];
#pragma unroll
    for (int t = 0; t < NT; t++) {
        rowI[t] = T + cb_tile_id(I0 + stride * t, 0) * 64;
        acc[t][0] = acc[t][1] = 0.0;
        acd[t][0] = acd[t][1] = 0.0;
        a0[t] = rowI[t][fa0]; a1[t] = rowI[t][fa1];
    }
    double b0 = rowJ[fa0], b1 = rowJ[fa1];
    for (int m = 0; m < J; m++) {
        // step m + 1 <= J still addresses tiles of the same block rows (valid shared memory); its values are unused at m + 1 == J
        const int nx = (m + 1) * 64;
        const double nb0 = rowJ[nx + fa0], nb1 = rowJ[nx + fa1];
        double na0[NT], na1[NT];
#pragma unroll
        for (int t = 0; t < NT; t++) { na0[t] = rowI[t][nx + fa0]; na1[t] = rowI[t][nx + fa1]; }
#pragma unroll
        for (int t = 0; t < NT; t++) {
            cb_dmma(acc[t][0], acc[t][1], a0[t], b0);
            cb_dmma(acd[t][0], acd[t][1], a1[t], b1);
        }
        b0 = nb0; b1 = nb1;
#pragma unroll
        for (int t = 0; t < NT; t++) { a0[t] = na0[t]; a1[t] = na1[t]; }
    }
#pragma unroll
    for (int t = 0; t < NT; t++) {
        double2* c = reinterpret_cast<double2*>(rowI[t] + (DIAG ? I0 : J) * 64 + fc);
        double2 v = *c;
        v.x -= acc[t][0] + acd[t][0]; v.y -= acc[t][1] + acd[t][1];
        *c = v;
    }
}

// shared-memory bytes needed for a k x k system (+1 row when a right-hand side rides along)
inline size_t cb_smem_bytes(int k, bool with_rhs) {
    const int NB = (k + (with_rhs ? 1 : 0) + 7) / 8;
    // tiles + scratch diagonal tile + published diagonal factor + inverted diagonal blocks + solve vector + flag
    return sizeof(double) * ((size_t)NB * (NB + 1) / 2 * 64 + 64 + 64 + (size_t)NB * 64 + (size_t)NB * 8) + 16;
}
// warps per system: 4 while one panel row per panel thread suffices (3 x 32 rows below the first diagonal block), 8 above
inline int cb_warps(int k, bool with_rhs) { return (k + (with_rhs ? 1 : 0) + 7) / 8 <= 13 ? 4 : 8; }

// tools/chol_bench.cu builds this header with -DCB_PROF: clock64() totals of the kernel's stages for one mid-grid CTA
#ifdef CB_PROF
__device__ long long cb_prof_cycles[8][8];      // [warp][stage], plain stores (atomics perturb the timing)
#define CB_T0() long long cb_t_ = clock64(); long long cb_acc_[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define CB_TICK(stage)                                  \
    do {                                                \
        const long long now_ = clock64();               \
        cb_acc_[stage] += now_ - cb_t_;                 \
        cb_t_ = now_;                                   \
    } while (0)
#define CB_FLUSH()                                                                                   \
    do {                                                                                             \
        if (blockIdx.x == gridDim.x / 2 && (threadIdx.x & 31) == 0)                                  \
            for (int s_ = 0; s_ < 8; s_++) cb_prof_cycles[threadIdx.x >> 5][s_] = cb_acc_[s_];       \
    } while (0)
#else
#define CB_T0()
#define CB_TICK(stage)
#define CB_FLUSH()
#endif

// BIG = true (k + 1 > 8 * 27, i.e. the tiles no longer fit 227 KB of shared memory; update_k may grow k towards 2p,
// BSFG_functions.R:332): the same algorithm with the tile array in a per-CTA slice of a global workspace (L2 resident), plain
// loads instead of cp.async for staging, and a persistent grid whose CTAs walk the systems.  Latency bound on L2 instead of
// shared memory -- a correctness path for factor counts the reference allows, not a tuned one.
template <int NW, bool BIG>
__device__ __forceinline__ void
cb_system(double* sm, double* T, const int sys, const double* __restrict__ Qp, long ldq, int k, const uint16_t* __restrict__ tile_off,
          const double* __restrict__ dadd, long dadd_sys_stride, long dadd_elem_stride,
          const double* __restrict__ rhs, long ldr, const VarSrc& zsrc, double* __restrict__ out, long ldo,
          double* __restrict__ Lout, int* __restrict__ fail_flag) {
    constexpr int T_ = NW * 32, PW = NW - 1, PT = PW * 32;
    const bool aug = rhs != nullptr;               // right-hand side stored as matrix row k
    const int kk = k + (aug ? 1 : 0);
    const int NB = (kk + 7) >> 3;
    const int ntile = NB * (NB + 1) / 2;
    const int npairs = k * (k + 1) / 2;
    double* Dtmp = BIG ? sm : T + ntile * 64;     // 64 : the diagonal tile being factored, DMMA fragment -> one row per register set
    double* Lpub = Dtmp + 64;             // 64 : factored diagonal block for the panel rows (tile layout, 1/L[c][c] on the diagonal)
    double* Dinv = Lpub + 64;             // NB * 64 : inverse of each factored diagonal block (row-major r*8+c, c <= r)
    double* xv = Dinv + NB * 64;          // 8 NB : backward-solve vector
    int* s_bad = (int*)(xv + 8 * NB);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    CB_T0();

    // ---- stage the system into tiles: one warp per matrix row, every element its own 8-byte cp.async (no staging
    //      registers, no offset table: the copies of a row are issued back to back).  Only the last block row holds
    //      padding (rows kk..8NB-1, identity) and needs clearing; the upper triangles of the diagonal tiles are never read.
    {
        double2* T2 = reinterpret_cast<double2*>(T + cb_tile_id(NB - 1, 0) * 64);
        for (int idx = tid; idx < NB * 32; idx += T_) T2[idx] = make_double2(0.0, 0.0);
        if (tid == 0) *s_bad = 0;
    }
    __syncthreads();
    {
        const double* q = Qp + (long)sys * ldq;
        for (int a = warp; a < k; a += NW) {
            const double* qa = q + (long)a * (a + 1) / 2;
            double* Ta = T + cb_tile_id(a >> 3, 0) * 64 + (a & 7) * 8;
            const int sw4 = (a & 2) << 1;
            for (int b = lane; b <= a; b += 32) {
                if constexpr (BIG) Ta[(b >> 3) * 64 + ((b & 7) ^ sw4)] = qa[b];
                else cb_cp_async8(Ta + (b >> 3) * 64 + ((b & 7) ^ sw4), qa + b);
            }
        }
        for (int a = kk + tid; a < 8 * NB; a += T_) T[cb_tile_id(a >> 3, a >> 3) * 64 + cb_swz(a & 7, a & 7)] = 1.0;
        if (aug) {
            const int Ik = k >> 3, rk = k & 7;
            for (int j = tid; j < k; j += T_) T[cb_tile_id(Ik, j >> 3) * 64 + cb_swz(rk, j & 7)] = rhs[(long)sys * ldr + j];
            if (tid == 0) T[cb_tile_id(Ik, Ik) * 64 + cb_swz(rk, rk)] = 1e300;   // keeps the pivot of row k positive; its factor is never used
        }
        if constexpr (!BIG) cb_cp_async_wait_all();
    }
    __syncthreads();
    if (dadd) {
        for (int a = tid; a < k; a += T_)
            T[cb_tile_id(a >> 3, a >> 3) * 64 + cb_swz(a & 7, a & 7)] += dadd[(long)sys * dadd_sys_stride + (long)a * dadd_elem_stride];
    }
    __syncthreads();
    CB_TICK(0);

    // ---- blocked left-looking Cholesky ------------------------------------------------------------
    const int gq = lane >> 2, q = lane & 3;
    const int fa0 = cb_swz(gq, q), fa1 = cb_swz(gq, 4 + q);     // operand fragment offsets, k-halves 0 and 1
    const int fc = cb_swz(gq, 2 * q);                           // accumulator pair (even -> 16-byte aligned)
    const int dw = sys % NW;                                    // this CTA's diagonal warp (spread over the sub-partitions)
    const bool is_diag = warp == dw;
    const int pw = warp - (warp > dw ? 1 : 0);                  // panel-warp index 0..PW-1
    const int ptid = pw * 32 + lane;
    const bool want_inv = out != nullptr && aug;
    for (int J = 0; J < NB; J++) {
        double* Tjj = T + cb_tile_id(J, J) * 64;
        if (is_diag) {
            double L[8][8], inv[8];                             // the factored tile, kept in registers for the inverse
            // ---- the pivot chain: last rank-8 term of the diagonal tile, 8 x 8 factorisation, publish ----
            double2 v = *reinterpret_cast<const double2*>(Tjj + fc);
            if (J > 0) {
                const double* rowJ = T + cb_tile_id(J, J - 1) * 64;
                const double a0 = rowJ[fa0], a1 = rowJ[fa1];
                double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                cb_dmma(c0, c1, a0, a0);
                cb_dmma(d0, d1, a1, a1);
                v.x -= c0 + d0; v.y -= c1 + d1;
            }
            *reinterpret_cast<double2*>(Dtmp + fc) = v;
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int c = 0; c <= r; c++) L[r][c] = Dtmp[cb_swz(r, c)];
            bool bad = false;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const double dd = L[c][c];
                if (!(dd > 0.0)) bad = true;
                inv[c] = cb_rsqrt(dd);
                L[c][c] = dd * inv[c];
#pragma unroll
                for (int r = c + 1; r < 8; r++) L[r][c] *= inv[c];
#pragma unroll
                for (int r = c + 1; r < 8; r++)
#pragma unroll
                    for (int l = c + 1; l <= r; l++) L[r][l] -= L[r][c] * L[l][c];
            }
            if (lane < 2) {
                // lane 0 writes L_JJ into its tile (factor output, backward solve), lane 1 the panel copy with 1/L[c][c] on
                // the diagonal; only the lower triangle is ever read, pairs above it carry don't-care values
                double* dst = lane ? Lpub : Tjj;
                if (bad && lane == 0) *s_bad = 1;
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const int sw4 = (r & 2) << 1;
#pragma unroll
                    for (int c = 0; c <= r; c += 2) {
                        const double x0 = (c == r) ? (lane ? inv[r] : L[r][r]) : L[r][c];
                        const double x1 = (c + 1 < r) ? L[r][c + 1] : ((c + 1 == r) ? (lane ? inv[r] : L[r][r]) : 0.0);
                        *reinterpret_cast<double2*>(dst + r * 8 + (c ^ sw4)) = make_double2(x0, x1);
                    }
                }
            }
            CB_TICK(1);
            cb_bar(1, T_);
            CB_TICK(2);
            // ---- slack of the diagonal warp: lane c < 8 computes column c of X = L_JJ^-1 as the forward substitution of
            //      the unit vector e_c (x L' = e_c'), stored row-major for the backward solve ----
            if (want_inv) {
                double x[8];
#pragma unroll
                for (int c = 0; c < 8; c++) x[c] = (c == lane) ? 1.0 : 0.0;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    x[c] *= inv[c];
#pragma unroll
                    for (int l = c + 1; l < 8; l++) x[l] -= x[c] * L[l][c];
                }
                if (lane < 8) {
#pragma unroll
                    for (int rr = 0; rr < 8; rr++) Dinv[J * 64 + rr * 8 + lane] = x[rr];     // X[rr][lane]; zero above the diagonal
                }
            }
        } else {
          if (J > 0) {
            // ---- tensor work of column J: tiles (I,J), I > J, and the partial sum of the next diagonal tile ----
            // items t = 0..NB-J-2 are the tiles I = J+1+t, item NB-J-1 is the diagonal tile J+1 (absent for the last column);
            // panel warp pw takes items pw, pw + PW, ...
            const int ntl = NB - J - 1;                                  // off-diagonal tiles of this column
            const int nt = ntl > pw ? (ntl - pw + PW - 1) / PW : 0;
            int I0 = J + 1 + pw, rem = nt;
            if constexpr (BIG) {      // more than 5 tiles per warp only beyond 28 block rows
                while (rem > 5) { cb_update_tiles<4, false>(T, J, I0, PW, fa0, fa1, fc); I0 += 4 * PW; rem -= 4; }
            }
            switch (rem) {
                case 1: cb_update_tiles<1, false>(T, J, I0, PW, fa0, fa1, fc); break;
                case 2: cb_update_tiles<2, false>(T, J, I0, PW, fa0, fa1, fc); break;
                case 3: cb_update_tiles<3, false>(T, J, I0, PW, fa0, fa1, fc); break;
                case 4: cb_update_tiles<4, false>(T, J, I0, PW, fa0, fa1, fc); break;
                case 5: cb_update_tiles<5, false>(T, J, I0, PW, fa0, fa1, fc); break;
                default: break;
            }
            if (ntl > 0 && ntl % PW == pw) cb_update_tiles<1, true>(T, J, J + 1, 0, fa0, fa1, fc);
            CB_TICK(1);
          }
            cb_bar(1, T_);
            CB_TICK(2);
            // ---- panel rows: x L_JJ' = a, one row per thread ----
            const int nrows = 8 * (NB - J - 1);
            if (ptid < nrows) {
                double P[8][8], pinv[8];
#pragma unroll
                for (int r = 0; r < 8; r++) {
#pragma unroll
                    for (int c = 0; c < r; c++) P[r][c] = Lpub[cb_swz(r, c)];
                    pinv[r] = Lpub[cb_swz(r, r)];
                }
                for (int prow = ptid; prow < nrows; prow += PT) {
                    const int row = 8 * (J + 1) + prow;
                    const int I = row >> 3, r = row & 7;
                    double* mine = T + cb_tile_id(I, J) * 64 + r * 8;
                    const int sw4 = (r & 2) << 1;                      // the row's 16-byte pairs sit at columns (c ^ sw4), c even
                    double x[8];
#pragma unroll
                    for (int c = 0; c < 8; c += 2) {
                        const double2 v2 = *reinterpret_cast<const double2*>(mine + (c ^ sw4));
                        x[c] = v2.x; x[c + 1] = v2.y;
                    }
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        x[c] *= pinv[c];
#pragma unroll
                        for (int l = c + 1; l < 8; l++) x[l] -= x[c] * P[l][c];
                    }
#pragma unroll
                    for (int c = 0; c < 8; c += 2) *reinterpret_cast<double2*>(mine + (c ^ sw4)) = make_double2(x[c], x[c + 1]);
                }
            }
        }
        CB_TICK(3);
        __syncthreads();
        CB_TICK(4);
    }
    if (*s_bad) {
        if (tid == 0) atomicMax(fail_flag, sys + 1);
        return;
    }
    if (Lout) {
        if constexpr (BIG) {      // no uint16 offset table beyond 65535 tile elements
            for (int a = warp; a < k; a += NW)
                for (int b = lane; b <= a; b += 32)
                    Lout[(long)sys * npairs + (long)a * (a + 1) / 2 + b] = T[cb_tile_id(a >> 3, b >> 3) * 64 + cb_swz(a & 7, b & 7)];
        } else {
            for (int idx = tid; idx < npairs; idx += T_) Lout[(long)sys * npairs + idx] = T[tile_off[idx]];
        }
    }
    if (!out || !aug) return;

    // ---- w = v + z with v' = row k of the factor (cpp:125, 128: mlam and ylam share one back-substitution) ---------
    {
        const int Ik = k >> 3, rk = k & 7;
        for (int j = tid; j < 8 * NB; j += T_)
            xv[j] = (j < k) ? T[cb_tile_id(Ik, j >> 3) * 64 + cb_swz(rk, j & 7)] + var_normal(zsrc, j, sys) : 0.0;
    }
    __syncthreads();
    CB_TICK(5);
    if (warp != 0) { CB_FLUSH(); return; }
    // ---- backward solve L' x = w on one warp ------------------------------------------------------------------------
    for (int J = NB - 1; J >= 0; J--) {
        double xr = 0.0;
        if (lane < 8) {
            // x_r = sum_{c >= r} inv(L_JJ)[c][r] w_c  (entries with c < r are stored zeros); two partial sums
            const double* Xb = Dinv + J * 64;
            double xa = 0.0, xb = 0.0;
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                xa += Xb[c * 8 + lane] * xv[8 * J + c];
                xb += Xb[(c + 1) * 8 + lane] * xv[8 * J + c + 1];
            }
            xr = xa + xb;
        }
        __syncwarp();
        if (lane < 8) xv[8 * J + lane] = xr;
        __syncwarp();
        // w_i -= sum_c L[8J + c][i] x_c for every earlier row i < 8J: four independent rows per lane in flight
        double xc[8];
#pragma unroll
        for (int c = 0; c < 8; c++) xc[c] = xv[8 * J + c];
        for (int i0 = lane; i0 < 8 * J; i0 += 128) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            const double* Tt[4];
            int ri[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int i = min(i0 + 32 * t, 8 * J - 1);
                Tt[t] = T + cb_tile_id(J, i >> 3) * 64;
                ri[t] = i & 7;
            }
#pragma unroll
            for (int c = 0; c < 8; c++)
#pragma unroll
                for (int t = 0; t < 4; t++) acc[t] += Tt[t][cb_swz(c, ri[t])] * xc[c];
#pragma unroll
            for (int t = 0; t < 4; t++)
                if (i0 + 32 * t < 8 * J) xv[i0 + 32 * t] -= acc[t];
        }
        __syncwarp();
    }
    for (int i = lane; i < k; i += 32) out[(long)sys * ldo + i] = xv[i];
    CB_TICK(6);
    CB_FLUSH();
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, NW == 4 ? 4 : 1)
chol_blocked_kernel(const double* __restrict__ Qp, long ldq, int k, int nsys, const uint16_t* __restrict__ tile_off,
                    const double* __restrict__ dadd, long dadd_sys_stride, long dadd_elem_stride,
                    const double* __restrict__ rhs, long ldr, VarSrc zsrc, double* __restrict__ out, long ldo,
                    double* __restrict__ Lout, int* __restrict__ fail_flag) {
    extern __shared__ __align__(16) double sm[];
    const int sys = blockIdx.x;
    if (sys >= nsys) return;
    cb_system<NW, false>(sm, sm, sys, Qp, ldq, k, tile_off, dadd, dadd_sys_stride, dadd_elem_stride, rhs, ldr, zsrc, out, ldo, Lout, fail_flag);
}

// shared-memory bytes of the BIG variant: scratch diagonal tile, published factor, inverted diagonal blocks, solve vector, flag
inline size_t cb_big_smem_bytes(int k, bool with_rhs) {
    const int NB = (k + (with_rhs ? 1 : 0) + 7) / 8;
    return sizeof(double) * (64 + 64 + (size_t)NB * 64 + (size_t)NB * 8) + 16;
}
// doubles of global workspace per CTA of the BIG variant (the tile array)
inline long cb_big_ws_doubles(int k, bool with_rhs) {
    const long NB = (k + (with_rhs ? 1 : 0) + 7) / 8;
    return NB * (NB + 1) / 2 * 64;
}
__global__ void __launch_bounds__(256, 2)
chol_blocked_big_kernel(const double* __restrict__ Qp, long ldq, int k, int nsys,
                        const double* __restrict__ dadd, long dadd_sys_stride, long dadd_elem_stride,
                        const double* __restrict__ rhs, long ldr, VarSrc zsrc, double* __restrict__ out, long ldo,
                        double* __restrict__ Lout, int* __restrict__ fail_flag, double* __restrict__ gws, long gws_stride) {
    extern __shared__ __align__(16) double sm[];
    double* T = gws + (long)blockIdx.x * gws_stride;
    for (int sys = blockIdx.x; sys < nsys; sys += gridDim.x) {
        cb_system<8, true>(sm, T, sys, Qp, ldq, k, nullptr, dadd, dadd_sys_stride, dadd_elem_stride, rhs, ldr, zsrc, out, ldo, Lout, fail_flag);
        __syncthreads();          // the backward-solve warp is done with T before the next system is staged
    }
}

}  // namespace bsfg
