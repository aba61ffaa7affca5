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
// Rows/columns k..8NB-1 are identity padding.
//
// Per block column J:
//   phase 1 (all warps, tensor pipe): tile(I,J) -= sum_{m<J} L(I,m) L(J,m)'  for I = J..NB-1.  A warp owns up
//           to 4 tiles (I = J + warp + NW t) and keeps their 8x8 accumulators in the DMMA C fragment (2 doubles
//           per lane per tile); both operands are K-contiguous tile rows, so A and B fragments load identically.
//           The updated diagonal tile is also written to a scratch tile, so that
//   phase 2 (one thread per matrix row) can read it while its owners overwrite the real one: every thread factors
//           the 8x8 diagonal tile redundantly in registers (36 values; no shuffles, no extra barrier) and applies the
//           8-step forward substitution to its own row of the panel below; the 8 owners also store the INVERSE of the
//           factored diagonal block (used by the backward solve).
// 2 barriers per block column.  (Factoring the tile in its 8 owner threads only and letting the panel rows multiply by
// the shared inverse removes 40% of the kernel's instructions but needs a third barrier per column: measured SLOWER on
// B200, Lambda-Cholesky phase 2.53 -> 2.73 ms at C4 -- the kernel is barrier/latency bound, not issue bound.)  The right-hand side m rides along as an extra matrix row k (with a huge diagonal), so
// the forward solve v = L^-1 m costs nothing: v' is row k of the factor.  The backward solve L'x = v + z runs on ONE
// warp without block barriers: x_J = inv(L_JJ)' w_J is an 8x8 mat-vec (no 8-step dependency chain), followed by an
// 8-FMA update of the earlier rows.  (ncu on the first version: ~40% of the stall samples sat on the barriers of the
// two per-row solves, where three of four warps waited for the in-warp substitution of the fourth.)
// The per-element map packed-index -> tile offset is a uint16 table built once per k on the host.
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

// shared-memory bytes needed for a k x k system (+1 row when a right-hand side rides along)
inline size_t cb_smem_bytes(int k, bool with_rhs) {
    const int NB = (k + (with_rhs ? 1 : 0) + 7) / 8;
    // tiles + scratch diagonal tile + inverted diagonal blocks + solve vector + flag
    return sizeof(double) * ((size_t)NB * (NB + 1) / 2 * 64 + 64 + (size_t)NB * 64 + (size_t)NB * 8) + 16;
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, NW == 4 ? 4 : 1)
chol_blocked_kernel(const double* __restrict__ Qp, long ldq, int k, int nsys, const uint16_t* __restrict__ tile_off,
                    const double* __restrict__ dadd, long dadd_sys_stride, long dadd_elem_stride,
                    const double* __restrict__ rhs, long ldr, VarSrc zsrc, double* __restrict__ out, long ldo,
                    double* __restrict__ Lout, int* __restrict__ fail_flag) {
    extern __shared__ __align__(16) double sm[];
    constexpr int T_ = NW * 32;
    const bool aug = rhs != nullptr;               // right-hand side stored as matrix row k
    const int kk = k + (aug ? 1 : 0);
    const int NB = (kk + 7) >> 3;
    const int ntile = NB * (NB + 1) / 2;
    const int npairs = k * (k + 1) / 2;
    double* T = sm;                       // ntile * 64
    double* Dtmp = T + ntile * 64;        // 64 : copy of the diagonal tile being factored
    double* Dinv = Dtmp + 64;             // NB * 64 : inverse of each factored diagonal block (row-major r*8+c, c <= r)
    double* xv = Dinv + NB * 64;          // 8 NB : backward-solve vector
    int* s_bad = (int*)(xv + 8 * NB);
    const int sys = blockIdx.x;
    if (sys >= nsys) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- stage the system into tiles -------------------------------------------------------------
    {
        double2* T2 = reinterpret_cast<double2*>(T);
        for (int idx = tid; idx < ntile * 32; idx += T_) T2[idx] = make_double2(0.0, 0.0);
        if (tid == 0) *s_bad = 0;
    }
    __syncthreads();
    {
        const double* q = Qp + (long)sys * ldq;
#pragma unroll 8
        for (int idx = tid; idx < npairs; idx += T_) T[tile_off[idx]] = q[idx];
        for (int a = kk + tid; a < 8 * NB; a += T_) T[cb_tile_id(a >> 3, a >> 3) * 64 + cb_swz(a & 7, a & 7)] = 1.0;
        if (aug) {
            const int Ik = k >> 3, rk = k & 7;
            for (int j = tid; j < k; j += T_) T[cb_tile_id(Ik, j >> 3) * 64 + cb_swz(rk, j & 7)] = rhs[(long)sys * ldr + j];
            if (tid == 0) T[cb_tile_id(Ik, Ik) * 64 + cb_swz(rk, rk)] = 1e300;   // keeps the pivot of row k positive; its factor is never used
        }
    }
    __syncthreads();
    if (dadd) {
        for (int a = tid; a < k; a += T_)
            T[cb_tile_id(a >> 3, a >> 3) * 64 + cb_swz(a & 7, a & 7)] += dadd[(long)sys * dadd_sys_stride + (long)a * dadd_elem_stride];
    }
    __syncthreads();
    if (tid < 64) Dtmp[tid] = T[tid];     // tile (0,0): nothing to update before its factorisation
    __syncthreads();

    // ---- blocked left-looking Cholesky ------------------------------------------------------------
    const int gq = lane >> 2, q = lane & 3;
    const int fa0 = cb_swz(gq, q), fa1 = cb_swz(gq, 4 + q);     // operand fragment offsets, k-halves 0 and 1
    const int fc = cb_swz(gq, 2 * q);                           // accumulator pair (even -> 16-byte aligned)
    for (int J = 0; J < NB; J++) {
        if (J > 0) {
            const double* rowJ = T + cb_tile_id(J, 0) * 64;
            const double* rowI[4];
            bool act[4];
            double acc[4][2];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int I = J + warp + NW * t;
                act[t] = I < NB;
                rowI[t] = T + cb_tile_id(act[t] ? I : J, 0) * 64;
                acc[t][0] = acc[t][1] = 0.0;
            }
            for (int m = 0; m < J; m++) {
                const double b0 = rowJ[m * 64 + fa0], b1 = rowJ[m * 64 + fa1];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    if (act[t]) {                                   // warp-uniform
                        const double a0 = rowI[t][m * 64 + fa0], a1 = rowI[t][m * 64 + fa1];
                        cb_dmma(acc[t][0], acc[t][1], a0, b0);
                        cb_dmma(acc[t][0], acc[t][1], a1, b1);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < 4; t++) {
                if (act[t]) {
                    double2* c = reinterpret_cast<double2*>(const_cast<double*>(rowI[t]) + J * 64 + fc);
                    double2 v = *c;
                    v.x -= acc[t][0]; v.y -= acc[t][1];
                    *c = v;
                    if (t == 0 && warp == 0) *reinterpret_cast<double2*>(Dtmp + fc) = v;   // I == J: the diagonal tile
                }
            }
            __syncthreads();
        }
        {
            const int row = 8 * J + tid;
            if (row < 8 * NB) {
                double L[8][8];
#pragma unroll
                for (int r = 0; r < 8; r++)
#pragma unroll
                    for (int c = 0; c <= r; c++) L[r][c] = Dtmp[cb_swz(r, c)];
                double inv[8];
                bool bad = false;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const double dd = L[c][c];
                    if (!(dd > 0.0)) bad = true;
                    inv[c] = rsqrt(dd);
                    L[c][c] = dd * inv[c];
#pragma unroll
                    for (int r = c + 1; r < 8; r++) L[r][c] *= inv[c];
#pragma unroll
                    for (int r = c + 1; r < 8; r++)
#pragma unroll
                        for (int l = c + 1; l <= r; l++) L[r][l] -= L[r][c] * L[l][c];
                }
                const int I = row >> 3, r = row & 7;
                double* mine = T + cb_tile_id(I, J) * 64;
                if (I == J) {
                    if (bad && r == 0) *s_bad = 1;
                    // inverse of the factored block, X = L^-1 (lower), one column at a time (8 live values):
                    // X[c][c] = 1/L[c][c], X[rr][c] = -(1/L[rr][rr]) sum_{m=c}^{rr-1} L[rr][m] X[m][c]; this thread keeps row r
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        double xc[8];
                        xc[c] = inv[c];
#pragma unroll
                        for (int rr = c + 1; rr < 8; rr++) {
                            double acc = 0.0;
#pragma unroll
                            for (int m = c; m < rr; m++) acc += L[rr][m] * xc[m];
                            xc[rr] = -inv[rr] * acc;
                        }
                        double mine_x = 0.0;
#pragma unroll
                        for (int rr = c; rr < 8; rr++)
                            if (rr == r) mine_x = xc[rr];
                        Dinv[J * 64 + r * 8 + c] = mine_x;               // zero above the diagonal (c > r)
                    }
#pragma unroll
                    for (int rr = 0; rr < 8; rr++) {
                        if (rr == r) {
#pragma unroll
                            for (int c = 0; c < 8; c++) mine[cb_swz(rr, c)] = (c <= rr) ? L[rr][c] : 0.0;
                        }
                    }
                } else {
                    double x[8];
#pragma unroll
                    for (int c = 0; c < 8; c++) x[c] = mine[cb_swz(r, c)];
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        x[c] *= inv[c];
#pragma unroll
                        for (int l = c + 1; l < 8; l++) x[l] -= x[c] * L[l][c];
                    }
#pragma unroll
                    for (int c = 0; c < 8; c++) mine[cb_swz(r, c)] = x[c];
                }
            }
            __syncthreads();
        }
    }
    if (*s_bad) {
        if (tid == 0) atomicMax(fail_flag, sys + 1);
        return;
    }
    if (Lout) {
        for (int idx = tid; idx < npairs; idx += T_) Lout[(long)sys * npairs + idx] = T[tile_off[idx]];
    }
    if (!out || !aug) return;

    // ---- w = v + z with v' = row k of the factor (cpp:125, 128: mlam and ylam share one back-substitution) ---------
    {
        const int Ik = k >> 3, rk = k & 7;
        for (int j = tid; j < 8 * NB; j += T_)
            xv[j] = (j < k) ? T[cb_tile_id(Ik, j >> 3) * 64 + cb_swz(rk, j & 7)] + var_normal(zsrc, j, sys) : 0.0;
    }
    __syncthreads();
    if (warp != 0) return;
    // ---- backward solve L' x = w on one warp ------------------------------------------------------------------------
    for (int J = NB - 1; J >= 0; J--) {
        double xr = 0.0;
        if (lane < 8) {
            // x_r = sum_{c >= r} inv(L_JJ)[c][r] w_c
            const double* Xb = Dinv + J * 64;
#pragma unroll
            for (int c = 0; c < 8; c++) xr += Xb[c * 8 + lane] * xv[8 * J + c];     // entries with c < r are stored zeros
        }
        __syncwarp();
        if (lane < 8) xv[8 * J + lane] = xr;
        __syncwarp();
        // w_i -= sum_c L[8J + c][i] x_c for every earlier row i < 8J
        for (int i = lane; i < 8 * J; i += 32) {
            const double* Tt = T + cb_tile_id(J, i >> 3) * 64;
            const int ri = i & 7;
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < 8; c++) acc += Tt[cb_swz(c, ri)] * xv[8 * J + c];
            xv[i] -= acc;
        }
        __syncwarp();
    }
    for (int i = lane; i < k; i += 32) out[(long)sys * ldo + i] = xv[i];
}

}  // namespace bsfg
