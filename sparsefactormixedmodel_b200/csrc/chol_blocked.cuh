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
//   phase 2 (one thread per matrix row): every thread factors the updated 8x8 diagonal tile redundantly in
//           registers (36 values; no shuffles, no extra barrier) and applies the 8-step forward substitution to
//           its own row of the panel below.
// Then forward solve, + z, backward solve with one thread per row, 8-wide in-warp substitution on the diagonal
// tile and an 8-FMA panel update per block.  3 barriers per block column in the factorisation, 1 per block in
// each solve.  The per-element map packed-index -> tile offset is a uint16 table built once per k on the host.
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

// shared-memory doubles needed for a k x k system
inline size_t cb_smem_bytes(int k) {
    const int NB = (k + 7) / 8;
    return sizeof(double) * ((size_t)NB * (NB + 1) / 2 * 64 + 2 * (size_t)NB * 8) + 16;
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, NW == 4 ? 4 : 1)
chol_blocked_kernel(const double* __restrict__ Qp, long ldq, int k, int nsys, const uint16_t* __restrict__ tile_off,
                    const double* __restrict__ dadd, long dadd_sys_stride, long dadd_elem_stride,
                    const double* __restrict__ rhs, long ldr, VarSrc zsrc, double* __restrict__ out, long ldo,
                    double* __restrict__ Lout, int* __restrict__ fail_flag) {
    extern __shared__ __align__(16) double sm[];
    constexpr int T_ = NW * 32;
    const int NB = (k + 7) >> 3;
    const int ntile = NB * (NB + 1) / 2;
    const int npairs = k * (k + 1) / 2;
    double* T = sm;                       // ntile * 64
    double* xv = sm + ntile * 64;         // 8 NB : block-solution exchange
    double* invd = xv + 8 * NB;           // 8 NB : 1 / L[i][i]
    int* s_bad = (int*)(invd + 8 * NB);
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
        for (int a = k + tid; a < 8 * NB; a += T_) T[cb_tile_id(a >> 3, a >> 3) * 64 + cb_swz(a & 7, a & 7)] = 1.0;
    }
    __syncthreads();
    if (dadd) {
        for (int a = tid; a < k; a += T_)
            T[cb_tile_id(a >> 3, a >> 3) * 64 + cb_swz(a & 7, a & 7)] += dadd[(long)sys * dadd_sys_stride + (long)a * dadd_elem_stride];
    }
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
                }
            }
            __syncthreads();
        }
        {
            const int row = 8 * J + tid;
            const bool has_row = row < 8 * NB;
            double L[8][8];
            if (has_row) {
                const double* D = T + cb_tile_id(J, J) * 64;
#pragma unroll
                for (int r = 0; r < 8; r++)
#pragma unroll
                    for (int c = 0; c <= r; c++) L[r][c] = D[cb_swz(r, c)];
            }
            __syncthreads();      // every thread holds the diagonal tile before its owners overwrite it
            if (has_row) {
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
#pragma unroll
                    for (int rr = 0; rr < 8; rr++) {
                        if (rr == r) {
#pragma unroll
                            for (int c = 0; c < 8; c++) mine[cb_swz(rr, c)] = (c <= rr) ? L[rr][c] : 0.0;
                            invd[8 * J + rr] = inv[rr];
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
    if (!out) return;

    // ---- solves: one thread per row ------------------------------------------------------------------
    const int I = tid >> 3, r = tid & 7;
    const bool live = tid < 8 * NB;
    double bval = (tid < k) ? rhs[(long)sys * ldr + tid] : 0.0;
    // forward: L v = m
    for (int J = 0; J < NB; J++) {
        if (warp == (J >> 2)) {
            const int base = (8 * J) & 31;
            const bool mineb = live && I == J;
            const double* D = T + cb_tile_id(J, J) * 64;
            double Lr[8];
#pragma unroll
            for (int c = 0; c < 8; c++) Lr[c] = (mineb && c < r) ? D[cb_swz(r, c)] : 0.0;
            const double idg = mineb ? invd[8 * J + r] : 0.0;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if (mineb && r == c) bval *= idg;
                const double xc = __shfl_sync(0xffffffffu, bval, base + c);
                if (mineb && r > c) bval -= Lr[c] * xc;
            }
            if (mineb) xv[tid] = bval;
        }
        __syncthreads();
        if (live && I > J) {
            const double* Tt = T + cb_tile_id(I, J) * 64;
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < 8; c++) acc += Tt[cb_swz(r, c)] * xv[8 * J + c];
            bval -= acc;
        }
    }
    // w = v + z   (cpp:128: ylam solves against Zcol; the two back-substitutions share one pass)
    if (tid < k) bval += var_normal(zsrc, tid, sys);
    __syncthreads();
    // backward: L' x = w
    for (int J = NB - 1; J >= 0; J--) {
        if (warp == (J >> 2)) {
            const int base = (8 * J) & 31;
            const bool mineb = live && I == J;
            const double* D = T + cb_tile_id(J, J) * 64;
            double Lc[8];
#pragma unroll
            for (int c = 0; c < 8; c++) Lc[c] = (mineb && c > r) ? D[cb_swz(c, r)] : 0.0;   // column r of the diagonal tile
            const double idg = mineb ? invd[8 * J + r] : 0.0;
#pragma unroll
            for (int c = 7; c >= 0; c--) {
                if (mineb && r == c) bval *= idg;
                const double xc = __shfl_sync(0xffffffffu, bval, base + c);
                if (mineb && r < c) bval -= Lc[c] * xc;
            }
            if (mineb) xv[tid] = bval;
        }
        __syncthreads();
        if (live && I < J) {
            const double* Tt = T + cb_tile_id(J, I) * 64;
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < 8; c++) acc += Tt[cb_swz(c, r)] * xv[8 * J + c];
            bval -= acc;
        }
    }
    if (tid < k) out[(long)sys * ldo + tid] = bval;
}

}  // namespace bsfg
