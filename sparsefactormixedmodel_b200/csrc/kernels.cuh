// Non-GEMM kernels of the BSFG sweep (sm_100a): element-wise producers, the per-trait k x k
// Cholesky solve, the discrete h2 grid draws, triangular row solves, Gamma draws and reductions.
// All matrices are column-major FP64 with explicit leading dimensions.
#pragma once
#include "common.cuh"

namespace bsfg {

// variate source: injected (ptr != null, column-major with leading dimension ld) or device Philox
struct VarSrc {
    const double* ptr;
    long ld;
    RngStream rs;
    long col_offset;   // global index of local column 0 (trait sharding) for Philox addressing
    long rows_total;   // rows of the full logical block (for Philox element index)
};
__device__ __forceinline__ double var_normal(const VarSrc& v, long row, long col) {
    if (v.ptr) return v.ptr[col * v.ld + row];
    return philox_normal(v.rs, (uint64_t)((col + v.col_offset) * v.rows_total + row));
}
__device__ __forceinline__ double var_uniform(const VarSrc& v, long row, long col) {
    if (v.ptr) return v.ptr[col * v.ld + row];
    return philox_uniform(v.rs, (uint64_t)((col + v.col_offset) * v.rows_total + row));
}
__device__ __forceinline__ double var_gamma(const VarSrc& v, long row, long col, double shape) {
    if (v.ptr) return v.ptr[col * v.ld + row];
    return philox_gamma(v.rs, (uint64_t)((col + v.col_offset) * v.rows_total + row), shape);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum, result valid in thread 0 (and broadcast through smem to all)
__device__ __forceinline__ double block_sum(double v, double* red /* >= 32 doubles */) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (w == 0) {
        t = warp_sum(t);
        if (l == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}

// ------------------------------------------------------------------------------------
// transpose: out (cols x rows, ldo) = in' (rows x cols, ldi), optional row scaling of `in`
// (out[c][r] = in[r][c] * rowscale[r] if rowscale) -- 32x32 tiles through padded smem.
// ------------------------------------------------------------------------------------
__global__ void transpose_kernel(const double* __restrict__ in, long ldi, long rows, long cols,
                                 double* __restrict__ out, long ldo) {
    __shared__ double tile[32][33];
    const long r0 = (long)blockIdx.x * 32, c0 = (long)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        long r = r0 + threadIdx.x, c = c0 + j;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[c * ldi + r];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        long c = c0 + threadIdx.x, r = r0 + j;
        if (r < rows && c < cols) out[r * ldo + c] = tile[threadIdx.x][j];
    }
}

// out[r,c] = in[r,c] * rowscale[r] (rowscale may be null) * colscale[c] (may be null)
__global__ void scale_kernel(const double* __restrict__ in, long ldi, long rows, long cols,
                             const double* __restrict__ rowscale, const double* __restrict__ colscale,
                             double* __restrict__ out, long ldo) {
    const long total = rows * cols;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long r = idx % rows, c = idx / rows;
        double v = in[c * ldi + r];
        if (rowscale) v *= rowscale[r];
        if (colscale) v *= colscale[c];
        out[c * ldo + r] = v;
    }
}

// ------------------------------------------------------------------------------------
// sample_Lambda producers
// ------------------------------------------------------------------------------------
// Khatri-Rao ("Gram feature") build: Gt[i, pair(a,b)] = UtF[i,a] * UtF[i,b], a >= b,
// pair = a(a+1)/2 + b.  The weighted Gram Qlam_j = sum_i w_ij f_i f_i' (cpp:120-122) then is the
// plain GEMM  Q = Gt' W.
// Trait-sharded runs only need this rank's block of pairs [pair_lo, pair_hi) for the node product; the rest is built
// only when the exact fallback will consume the whole of G (*need_all != 0).
__global__ void gram_features_kernel(const double* __restrict__ UtF, long ldf, int n, int k,
                                     double* __restrict__ Gt, long ldg, int pair_lo, int pair_hi,
                                     const int* __restrict__ need_all) {
    const int pair = blockIdx.x;                   // pairs on grid.x: k(k+1)/2 exceeds the 65535 limit of grid.y from k = 362
    if ((pair < pair_lo || pair >= pair_hi) && !(need_all && *need_all != 0)) return;
    int a = (int)((sqrt(8.0 * pair + 1.0) - 1.0) * 0.5);
    while ((a + 1) * (a + 2) / 2 <= pair) a++;
    while (a * (a + 1) / 2 > pair) a--;
    const int b = pair - a * (a + 1) / 2;
    const double* fa = UtF + (long)a * ldf;
    const double* fb = UtF + (long)b * ldf;
    double* g = Gt + (long)pair * ldg;
    // Gt and UtF are DMat storage (even leading dimension, zero pad row): rows as 16-byte pairs
    for (int i = 2 * (blockIdx.y * blockDim.x + threadIdx.x); i < n; i += 2 * gridDim.y * blockDim.x) {
        const double2 x = *reinterpret_cast<const double2*>(fa + i), y = *reinterpret_cast<const double2*>(fb + i);
        *reinterpret_cast<double2*>(g + i) = make_double2(x.x * y.x, x.y * y.y);
    }
}

// W[i,j]  = ep_j / (s_i + ep_j / rp_j)                                   (cpp:120, the sweep weights)
// WY[i,j] = W[i,j] * (UtY[i,j] - sum_c UtX[i,c] B[c,j])                  (U'(Y - X B), R:189 rotated)
// W is only materialised when the exact Gram GEMM will consume it: W != null and (*need_w != 0 or need_w == null).
__global__ void lambda_weights_kernel(const double* __restrict__ UtY, long ldy, int n, int p,
                                      const double* __restrict__ s, const double* __restrict__ rp,
                                      const double* __restrict__ ep, const double* __restrict__ UtX, long ldx,
                                      const double* __restrict__ B, long ldb, int b,
                                      double* __restrict__ W, long ldw, const int* __restrict__ need_w,
                                      double* __restrict__ WY, long ldwy) {
    const int j = blockIdx.x;                      // one CTA per trait: 400 K 256-element blocks were latency bound
    if (j >= p) return;
    const bool write_w = W && (!need_w || *need_w != 0);
    const double epj = ep[j];
    const double c = epj / rp[j];
    double bj[4] = {0.0, 0.0, 0.0, 0.0};
    for (int cc = 0; cc < b && cc < 4; cc++) bj[cc] = B[(long)j * ldb + cc];
    // UtY, UtX, W, WY are DMat storage (even leading dimension, zero pad row): rows are processed as 16-byte pairs
    const double* __restrict__ ycol = UtY + (long)j * ldy;
    double* __restrict__ wycol = WY + (long)j * ldwy;
    double* __restrict__ wcol = write_w ? W + (long)j * ldw : nullptr;
#pragma unroll 4
    for (int i = 2 * threadIdx.x; i < n; i += 2 * blockDim.x) {
        const bool has1 = i + 1 < n;
        const double w0 = epj / (s[i] + c);
        const double w1 = has1 ? epj / (s[i + 1] + c) : 0.0;
        double2 y = *reinterpret_cast<const double2*>(ycol + i);
#pragma unroll
        for (int cc = 0; cc < 4; cc++)
            if (cc < b) {
                const double2 x = *reinterpret_cast<const double2*>(UtX + (long)cc * ldx + i);
                y.x -= x.x * bj[cc]; y.y -= x.y * bj[cc];
            }
        for (int cc = 4; cc < b; cc++) {
            const double bb = B[(long)j * ldb + cc];
            y.x -= UtX[(long)cc * ldx + i] * bb;
            if (has1) y.y -= UtX[(long)cc * ldx + i + 1] * bb;
        }
        if (write_w) *reinterpret_cast<double2*>(wcol + i) = make_double2(w0, w1);
        *reinterpret_cast<double2*>(wycol + i) = make_double2(w0 * y.x, w1 * y.y);
    }
}

// ------------------------------------------------------------------------------------
// Exponential-sum form of the weighted Gram.  The weights of cpp:120 are a Cauchy kernel,
//   w_ij = ep_j / (s_i + c_j),  c_j = ep_j / rp_j,   and   1/x = int exp(-x e^t) e^t dt
// discretised by the trapezoid rule with step h on t_m = t0 + m h, m < R, has RELATIVE error ~2 exp(-pi^2/h)
// (1.2e-14 at h = 0.28) uniformly on [x_lo, x_hi] once the grid covers [log(eps/x_hi), log(-log(eps)/x_lo)], with all
// terms positive (no cancellation):
//   w_ij ~= sum_m A[i,m] E[m,j],   A[i,m] = h e^{t_m} exp(-e^{t_m} s_i),   E[m,j] = ep_j exp(-e^{t_m} c_j)
// so  Q = G' W = (G' A) E : two GEMMs with inner dimension R (256) instead of one with inner dimension n.
// The grid is re-centred every iteration on the device from the current range of c; if R terms cannot cover the
// range (dynamic range > ~1e13) `need_exact` is raised and the exact W / Gram GEMM run instead (they are launched
// unconditionally and exit immediately when the flag is clear), so the path never silently loses accuracy.
// ------------------------------------------------------------------------------------
struct ExpSumPlan {
    double t0, h;
    int R, ok;
    double c_min, c_max;
};
// range[0] = -min_j c_j, range[1] = max_j c_j, range[2] = 1 if some c_j is not a positive finite number (else 0):
// three values that combine over ranks with ONE max-allreduce, so that every rank of a trait-sharded run places the
// same node grid (the node product G'A is then common work that can be split over the ranks).
__global__ void expsum_range_kernel(const double* __restrict__ ep, const double* __restrict__ rp, int p,
                                    double* __restrict__ range) {
    __shared__ double red[32];
    double lo = INFINITY, hi = 0.0;
    bool bad = false;
    for (int j = threadIdx.x; j < p; j += blockDim.x) {
        const double c = ep[j] / rp[j];
        if (!(c > 0.0) || !isfinite(c)) bad = true;
        lo = fmin(lo, c); hi = fmax(hi, c);
    }
    lo = -lo;
    lo = warp_max(lo); hi = warp_max(hi);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = lo;
    __syncthreads();
    double t = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
    t = warp_max(t);
    const double neg_c_min = __shfl_sync(0xffffffffu, t, 0);
    __syncthreads();
    if (l == 0) red[w] = hi;
    __syncthreads();
    t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    t = warp_max(t);
    const double c_max = __shfl_sync(0xffffffffu, t, 0);
    const int any_bad = __syncthreads_or(bad ? 1 : 0);
    if (threadIdx.x == 0) { range[0] = neg_c_min; range[1] = c_max; range[2] = any_bad ? 1.0 : 0.0; }
}
__global__ void expsum_plan_kernel(const double* __restrict__ range, double s_min, double s_max, double h, int R,
                                   int force_exact, ExpSumPlan* __restrict__ plan, int* __restrict__ need_exact,
                                   unsigned long long* __restrict__ fallback_count) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double c_min = -range[0], c_max = range[1];
    const bool any_bad = range[2] != 0.0;
    const double eps = 1e-16;
    const double x_lo = s_min + c_min, x_hi = s_max + c_max;
    const double t0 = log(eps / x_hi);
    const double need = log(-log(eps) / x_lo);
    const bool ok = !force_exact && !any_bad && x_lo > 0.0 && isfinite(x_hi) && isfinite(t0) && isfinite(need) &&
                    (t0 + h * (double)(R - 1) >= need);
    plan->t0 = t0; plan->h = h; plan->R = R; plan->ok = ok ? 1 : 0;
    plan->c_min = c_min; plan->c_max = c_max;
    *need_exact = ok ? 0 : 1;
    if (!ok && !force_exact && fallback_count) *fallback_count += 1ull;
}
// A[i,m] = h e^{t_m} exp(-e^{t_m} s_i)            (n x R, column m contiguous over i)
__global__ void expsum_nodes_kernel(const ExpSumPlan* __restrict__ plan, const double* __restrict__ s, int n,
                                    double* __restrict__ A, long lda) {
    if (!plan->ok) return;
    const int m = blockIdx.y;
    const double a = exp(plan->t0 + plan->h * (double)m);
    const double wgt = plan->h * a;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        A[(long)m * lda + i] = wgt * exp(-a * s[i]);
}
// E[m,j] = ep_j exp(-e^{t_m} c_j)                 (R x p, column j contiguous over m); one CTA per trait
__global__ void expsum_traits_kernel(const ExpSumPlan* __restrict__ plan, const double* __restrict__ ep,
                                     const double* __restrict__ rp, int p, double* __restrict__ E, long lde) {
    if (!plan->ok) return;
    const int j = blockIdx.x;
    if (j >= p) return;
    const double epj = ep[j];
    const double c = epj / rp[j];
    const int R = plan->R;
    for (int m = threadIdx.x; m < R; m += blockDim.x) {
        const double a = exp(plan->t0 + plan->h * (double)m);
        E[(long)j * lde + m] = epj * exp(-a * c);
    }
}

// pack a full symmetric k x k (column-major, ld) into packed-lower order
__global__ void pack_lower_kernel(const double* __restrict__ A, long ld, int k, double* __restrict__ out) {
    const int npairs = k * (k + 1) / 2;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < npairs; idx += gridDim.x * blockDim.x) {
        int a = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while ((a + 1) * (a + 2) / 2 <= idx) a++;
        while (a * (a + 1) / 2 > idx) a--;
        const int b = idx - a * (a + 1) / 2;
        out[idx] = A[(long)b * ld + a];
    }
}

// ------------------------------------------------------------------------------------
// sample_factors_scores row solves (cpp:152-160): for each individual i (a row of rhs),
//   F[i,:] = L'^-1 ( L^-1 rhs[i,:]' + z[i,:]' ),  L packed lower from the k x k Cholesky.
// One warp per row; L staged in shared memory once per CTA.
// rhs[i,h] = YL[i,h] + ZFa[i,h] * tau_e[h]  is formed on the fly.
// ------------------------------------------------------------------------------------
template <int KT>
__global__ void __launch_bounds__(256)
factor_rows_kernel(const double* __restrict__ Lp, int k, int n, const double* __restrict__ YL, long ldy,
                   const double* __restrict__ ZFa, long ldz, const double* __restrict__ tau_e, VarSrc zsrc, int row0,
                   double* __restrict__ Fout, long ldf, int stage_l /* 0: read the factor from global memory (large k) */) {
    extern __shared__ double sm[];
    const int npairs = k * (k + 1) / 2;
    if (stage_l) {
        for (int idx = threadIdx.x; idx < npairs; idx += blockDim.x) sm[idx] = Lp[idx];
        __syncthreads();
    }
    const double* L = stage_l ? sm : Lp;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int row_i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row_i < n; row_i += gridDim.x * warps_per_block) {
        double x[KT];
#pragma unroll
        for (int t = 0; t < KT; t++) {
            const int h = lane + 32 * t;
            x[t] = 0.0;
            if (h < k) {
                double v = YL[(long)h * ldy + row_i];
                if (ZFa) v += ZFa[(long)h * ldz + row_i] * tau_e[h];
                x[t] = v;
            }
        }
        for (int i = 0; i < k; i++) {
            const double* row = L + i * (i + 1) / 2;
            double part = 0.0;
#pragma unroll
            for (int t = 0; t < KT; t++) {
                const int j = lane + 32 * t;
                if (j < i) part += row[j] * x[t];
            }
            part = warp_sum(part);
            const int own = i & 31, ti = i >> 5;
#pragma unroll
            for (int t = 0; t < KT; t++)
                if (t == ti && lane == own) x[t] = (x[t] - part) / row[i];
        }
#pragma unroll
        for (int t = 0; t < KT; t++) {
            const int h = lane + 32 * t;
            if (h < k) x[t] += var_normal(zsrc, row_i + row0, h);      // z is n x k: element (i, h); row0 = first row of this rank's block
        }
        for (int j = k - 1; j >= 0; j--) {
            const double* row = L + j * (j + 1) / 2;
            const int own = j & 31, tj = j >> 5;
            double xj = 0.0;
#pragma unroll
            for (int t = 0; t < KT; t++)
                if (t == tj) xj = x[t];
            xj = __shfl_sync(0xffffffffu, xj, own) / row[j];
#pragma unroll
            for (int t = 0; t < KT; t++) {
                const int i = lane + 32 * t;
                if (t == tj && lane == own) x[t] = xj;
                else if (i < j) x[t] -= row[i] * xj;
            }
        }
#pragma unroll
        for (int t = 0; t < KT; t++) {
            const int h = lane + 32 * t;
            if (h < k) Fout[(long)h * ldf + row_i] = x[t];
        }
    }
}

// ------------------------------------------------------------------------------------
// sample_means element-wise step (cpp:75-78 without the U product):
//   theta[i,j] = (DtYt[i,j] * rp_j) / d + z[i,j] / sqrt(d),  d = s1_i ep_j + s2_i rp_j
// ------------------------------------------------------------------------------------
// One CTA per MEANS_TC = 16 consecutive traits.  A warp covers an 8-row x 16-trait tile per step: lane (rp, cg) owns the row
// pair (2 rp, 2 rp + 1), rp = 0..3, of the two traits 2 cg, 2 cg + 1, cg = 0..7 (MEANS_CPL = 2 traits per lane: measured 7%
// faster than 4 traits x 8 row pairs, 122 instead of 128 registers and no spills), so
//   * the two rows of a pair share one Philox block (one Box-Muller transform per two draws),
//   * DtYt is read and theta written as 16-byte pairs, 4 lanes side by side = 64 contiguous bytes per trait,
//   * theta' (p x br; the K = p operand of theta Lmsg, R:230), when thetat != null, is written from the same registers as
//     16 bytes per lane, 8 lanes side by side = one 128-byte line per row -- no transpose pass over the br x p array.
// When Bout != null the first b (<= MEANS_BMAX) rows of location_sample = U theta (cpp:78), i.e. B
// (fast_BSFG_sampler.R:206), are reduced in the same pass: B[c,j] = sum_i U[c,i] theta[i,j], U given transposed
// (Ut[i,c], column c contiguous); fixed summation order (shuffles, then warps in order).
// DtYt and theta are DMat storage (even leading dimension, zero pad row): the pair access of the last odd row is in bounds.
// Philox draws: element (i, j) has stream index (j + col_offset) rows_total + i; callers pass an even rows_total so that a
// row pair is one aligned Philox pair (an odd rows_total falls back to one block per draw).
constexpr int MEANS_BMAX = 4;
constexpr int MEANS_TC = 16;
#ifndef MEANS_CPL
#define MEANS_CPL 2          // traits per lane: 4 -> 8 row pairs x 4 trait groups per warp, 2 -> 4 row pairs x 8 trait groups
#endif
constexpr int MEANS_NRP = 32 / (MEANS_TC / MEANS_CPL);      // row pairs per warp step
constexpr int MEANS_THREADS = 128;
// gridDim.y CTAs split the rows of a trait tile (rows_per_cta each, a multiple of 64); with Bout they leave partial sums in
// Bpart[gridDim.y][p][MEANS_BMAX] and the last one to finish (counter[blockIdx.x], self-resetting) adds them in split order.
__global__ void __launch_bounds__(MEANS_THREADS, 4)
means_theta_kernel(const double* __restrict__ DtYt, long ldd, int br, int p,
                   const double* __restrict__ s1, const double* __restrict__ s2,
                   const double* __restrict__ rp, const double* __restrict__ ep, VarSrc zsrc,
                   double* __restrict__ theta, long ldt, double* __restrict__ thetat, long ldtt,
                   const double* __restrict__ Ut, long ldu, int b, double* __restrict__ Bout, long ldb,
                   int rows_per_cta, double* __restrict__ Bpart, int* __restrict__ counter) {
    __shared__ double red[MEANS_THREADS / 32][MEANS_TC][MEANS_BMAX];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rpi = lane % MEANS_NRP, cg = lane / MEANS_NRP;
    const int j0 = blockIdx.x * MEANS_TC + cg * MEANS_CPL;          // first of this lane's traits
    const int nc = max(0, min(MEANS_CPL, p - j0));
    double rpj[MEANS_CPL], epj[MEANS_CPL];
    const double* dcol[MEANS_CPL];
    double* tcol[MEANS_CPL];
    uint64_t base[MEANS_CPL];
#pragma unroll
    for (int c = 0; c < MEANS_CPL; c++) {
        const int j = min(j0 + c, p - 1);
        rpj[c] = rp[j]; epj[c] = ep[j];
        dcol[c] = DtYt + (long)j * ldd;
        tcol[c] = theta + (long)j * ldt;
        base[c] = (uint64_t)(j + zsrc.col_offset) * (uint64_t)zsrc.rows_total;
    }
    const bool paired = (zsrc.rows_total & 1) == 0;
    double acc[MEANS_CPL][MEANS_BMAX];
#pragma unroll
    for (int c = 0; c < MEANS_CPL; c++)
#pragma unroll
        for (int cc = 0; cc < MEANS_BMAX; cc++) acc[c][cc] = 0.0;
    const int row_lo = blockIdx.y * rows_per_cta, row_hi = min(br, row_lo + rows_per_cta);
    for (int r0 = row_lo + (warp * MEANS_NRP + rpi) * 2; r0 < row_hi; r0 += (MEANS_THREADS / 32) * MEANS_NRP * 2) {
        const bool has1 = r0 + 1 < br;
        const double s1a = s1[r0], s2a = s2[r0];
        const double s1b = has1 ? s1[r0 + 1] : 1.0, s2b = has1 ? s2[r0 + 1] : 1.0;
        double ua[MEANS_BMAX], ub[MEANS_BMAX];
        if (Bout) {
#pragma unroll
            for (int cc = 0; cc < MEANS_BMAX; cc++) {
                ua[cc] = cc < b ? Ut[(long)cc * ldu + r0] : 0.0;
                ub[cc] = (cc < b && has1) ? Ut[(long)cc * ldu + r0 + 1] : 0.0;
            }
        }
        double2 vin[MEANS_CPL];                          // the four 16-byte loads are in flight before the first Box-Muller starts
#pragma unroll
        for (int c = 0; c < MEANS_CPL; c++) vin[c] = c < nc ? *reinterpret_cast<const double2*>(dcol[c] + r0) : make_double2(0.0, 0.0);
        double th[MEANS_CPL][2];
#pragma unroll
        for (int c = 0; c < MEANS_CPL; c++) {
            th[c][0] = th[c][1] = 0.0;
            if (c < nc) {
                double z0, z1;
                if (zsrc.ptr) {
                    const double* zc = zsrc.ptr + (long)(j0 + c) * zsrc.ld + r0;
                    z0 = zc[0]; z1 = has1 ? zc[1] : 0.0;
                } else if (paired) {
                    philox_normal_pair(zsrc.rs, (base[c] + (uint64_t)r0) >> 1, &z0, &z1);
                } else {
                    z0 = philox_normal(zsrc.rs, base[c] + (uint64_t)r0);
                    z1 = philox_normal(zsrc.rs, base[c] + (uint64_t)r0 + 1);
                }
                const double2 v = vin[c];
                // mean = (v rp)/d, noise = z/sqrt(d) with one reciprocal square root r = d^-1/2 instead of a division, a
                // square root and another division (cpp:76-78)
                const double ra = fast_rsqrt(s1a * epj[c] + s2a * rpj[c]);
                const double rb = fast_rsqrt(s1b * epj[c] + s2b * rpj[c]);
                th[c][0] = ((v.x * rpj[c]) * ra + z0) * ra;
                th[c][1] = has1 ? ((v.y * rpj[c]) * rb + z1) * rb : 0.0;
                *reinterpret_cast<double2*>(tcol[c] + r0) = make_double2(th[c][0], th[c][1]);
                if (Bout) {
#pragma unroll
                    for (int cc = 0; cc < MEANS_BMAX; cc++)
                        if (cc < b) acc[c][cc] += ua[cc] * th[c][0] + ub[cc] * th[c][1];
                }
            }
        }
        if (thetat && nc > 0) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                if (e == 1 && !has1) break;
                double* dst = thetat + (long)(r0 + e) * ldtt + j0;       // ldtt even, j0 % 4 == 0: 16-byte aligned
                if (nc == MEANS_CPL) {
#pragma unroll
                    for (int c = 0; c < MEANS_CPL; c += 2) reinterpret_cast<double2*>(dst)[c >> 1] = make_double2(th[c][e], th[c + 1][e]);
                } else {
                    for (int c = 0; c < nc; c++) dst[c] = th[c][e];
                }
            }
        }
    }
    if (Bout) {
#pragma unroll
        for (int c = 0; c < MEANS_CPL; c++)
#pragma unroll
            for (int cc = 0; cc < MEANS_BMAX; cc++) {
                double v = acc[c][cc];
#pragma unroll
                for (int o = 1; o < MEANS_NRP; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (rpi == 0) red[warp][cg * MEANS_CPL + c][cc] = v;
            }
        __syncthreads();
        const bool split = gridDim.y > 1;
        for (int idx = threadIdx.x; idx < MEANS_TC * MEANS_BMAX; idx += MEANS_THREADS) {
            const int cl = idx / MEANS_BMAX, cc = idx % MEANS_BMAX;
            const int j = blockIdx.x * MEANS_TC + cl;
            if (j < p && cc < b) {
                double v = 0.0;
                for (int w = 0; w < MEANS_THREADS / 32; w++) v += red[w][cl][cc];
                if (split) Bpart[((long)blockIdx.y * p + j) * MEANS_BMAX + cc] = v;
                else Bout[(long)j * ldb + cc] = v;
            }
        }
        if (split) {
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) {
                const int t = atomicAdd(&counter[blockIdx.x], 1);
                s_last = (t == (int)gridDim.y - 1);
                if (s_last) counter[blockIdx.x] = 0;                  // ready for the next launch
            }
            __syncthreads();
            if (s_last) {
                __threadfence();
                for (int idx = threadIdx.x; idx < MEANS_TC * MEANS_BMAX; idx += MEANS_THREADS) {
                    const int cl = idx / MEANS_BMAX, cc = idx % MEANS_BMAX;
                    const int j = blockIdx.x * MEANS_TC + cl;
                    if (j < p && cc < b) {
                        double v = 0.0;
                        for (int y = 0; y < (int)gridDim.y; y++) v += __ldcg(&Bpart[((long)y * p + j) * MEANS_BMAX + cc]);
                        Bout[(long)j * ldb + cc] = v;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// discrete h2 grid (cpp:196-238): three small kernels.
// ------------------------------------------------------------------------------------
// det[i] = sum_l log((s_l + (1-h2)/h2) * h2) / 2, h2 = i/H, i >= 1; det[0] = 0      (cpp:216)
__global__ void h2_det_kernel(const double* __restrict__ s, int n, int H, double* __restrict__ det) {
    __shared__ double red[32];
    const int i = blockIdx.x;
    if (i == 0) {
        if (threadIdx.x == 0) det[0] = 0.0;
        return;
    }
    const double h2 = (double)i / (double)H;
    const double c = (1.0 - h2) / h2;
    double acc = 0.0;
    for (int l = threadIdx.x; l < n; l += blockDim.x) acc += log((s[l] + c) * h2) / 2.0;
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) det[i] = acc;
}

// ss[j,i] = sum_l dnorm_std_log(x_{j,i,l});  x = 1/sqrt(h2) * UtF[l,j] / sqrt(s_l + (1-h2)/h2)  (cpp:215)
// i = 0 uses F itself (cpp:218).  One warp per (factor, grid point).
__global__ void h2_loglik_kernel(const double* __restrict__ UtF, long ldu, const double* __restrict__ F, long ldf,
                                 const double* __restrict__ s, int n, int k, int H,
                                 double* __restrict__ ll /* k x H, ld k */) {
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_global >= k * H) return;
    const int j = warp_global % k, i = warp_global / k;
    const double c0 = -0.91893853320467274178;     // log(1/sqrt(2 pi)), cpp:23
    double acc = 0.0;
    if (i == 0) {
        const double* f = F + (long)j * ldf;
        for (int l = lane; l < n; l += 32) { const double x = f[l]; acc += c0 - 0.5 * (x * x); }
    } else {
        const double h2 = (double)i / (double)H;
        const double c = (1.0 - h2) / h2;
        const double ih = 1.0 / sqrt(h2);
        const double* f = UtF + (long)j * ldu;
        for (int l = lane; l < n; l += 32) {
            const double x = ih * (f[l] * (1.0 / sqrt(s[l] + c)));
            acc += c0 - 0.5 * (x * x);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) ll[(long)i * k + j] = acc;
}

// log_ps[j,i] = ll[j,i] - det[i] (+ extra[j,i]) + log(prior[i]); LSE-normalise; cumsum; count (cpp:226-235)
// One warp per row j; the 50-term normalise/cumsum runs sequentially on lane 0 in grid order,
// like the reference, so an index can only differ when u is within rounding of a cumsum edge.
__global__ void grid_draw_kernel(const double* __restrict__ ll, long ldl, const double* __restrict__ det, long det_ld,
                                 int det_per_row, const double* __restrict__ prior, int rows, int H, VarSrc usrc,
                                 double* __restrict__ out_h2, double* __restrict__ log_ps_out, long ldlp) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= rows) return;
    extern __shared__ double smg[];
    double* lp = smg + (threadIdx.x >> 5) * H;
    double mx = -INFINITY;
    for (int i = lane; i < H; i += 32) {
        const double d = det_per_row ? det[(long)i * det_ld + j] : det[i];
        const double v = ll[(long)i * ldl + j] - d + log(prior[i]);
        lp[i] = v;
        if (log_ps_out) log_ps_out[(long)i * ldlp + j] = v;
        mx = fmax(mx, v);
    }
    mx = warp_max(mx);
    __syncwarp();
    if (lane == 0) {
        double sum = 0.0;
        for (int i = 0; i < H; i++) sum += exp(lp[i] - mx);
        const double norm = mx + log(sum);
        const double u = var_uniform(usrc, j, 0);
        double cum = 0.0;
        int count = 0;
        for (int i = 0; i < H; i++) {
            cum += exp(lp[i] - norm);
            if (u > cum) count++;
        }
        out_h2[j] = (double)count / (double)H;
    }
}

// ------------------------------------------------------------------------------------
// sample_prec_discrete_conditional_c (cpp:241-289) log-likelihood table.  One warp per
// (trait, grid point).  ll[j,i] = sum_l dnorm_std_log(x);  det[j,i] per cpp:270-274.
// ------------------------------------------------------------------------------------
__global__ void prec_logdet_s_kernel(const double* __restrict__ s, int n, int H, double* __restrict__ slog) {
    // slog[i] = sum_l log(s_l + (1-h2)/h2), i >= 1
    __shared__ double red[32];
    const int i = blockIdx.x;
    if (i == 0) { if (threadIdx.x == 0) slog[0] = 0.0; return; }
    const double h2 = (double)i / (double)H;
    const double c = (1.0 - h2) / h2;
    double acc = 0.0;
    for (int l = threadIdx.x; l < n; l += blockDim.x) acc += log(s[l] + c);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) slog[i] = acc;
}

__global__ void prec_loglik_kernel(const double* __restrict__ UtY, long ldu, const double* __restrict__ Y, long ldy,
                                   const double* __restrict__ s, const double* __restrict__ slog,
                                   const double* __restrict__ res_prec, int n, int p, int H,
                                   double* __restrict__ ll /* p x H ld p */, double* __restrict__ det /* p x H ld p */) {
    const long warp_global = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_global >= (long)p * H) return;
    const int j = (int)(warp_global % p), i = (int)(warp_global / p);
    const double c0 = -0.91893853320467274178;
    const double rp = res_prec[j];
    double acc = 0.0;
    if (i == 0) {
        const double sq = sqrt(rp);
        const double* y = Y + (long)j * ldy;
        for (int l = lane; l < n; l += 32) { const double x = y[l] * sq; acc += c0 - 0.5 * (x * x); }
        acc = warp_sum(acc);
        if (lane == 0) {
            ll[(long)i * p + j] = acc;
            det[(long)i * p + j] = (double)(n / 2) * log(1.0 / rp);      // integer n/2, cpp:274
        }
    } else {
        const double h2 = (double)i / (double)H;
        const double c = (1.0 - h2) / h2;
        const double b2 = h2 / (rp * (1.0 - h2));
        const double ib = 1.0 / sqrt(b2);
        const double* y = UtY + (long)j * ldu;
        for (int l = lane; l < n; l += 32) {
            const double x = (y[l] * (1.0 / sqrt(s[l] + c))) * ib;
            acc += c0 - 0.5 * (x * x);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            ll[(long)i * p + j] = acc;
            // sum_l log((s_l + c) * b2)/2 = (slog_i + n log b2)/2     (cpp:270-271, factored)
            det[(long)i * p + j] = 0.5 * (slog[i] + (double)n * log(b2));
        }
    }
}

__global__ void prec_from_h2_kernel(const double* __restrict__ res_prec, const double* __restrict__ h2, int p,
                                    double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < p) out[j] = (res_prec[j] * (1.0 - h2[j])) / h2[j];          // cpp:286 (h2 == 0 -> Inf, kept)
}

// ------------------------------------------------------------------------------------
// sample_F_a element-wise step (cpp:320-333 before the U product):
//   v[i,j] = b3[i,j]*tau_e_j / d + z[i,j]/sqrt(d),  d = s2_i tau_u_j + s1_i tau_e_j
// (b3 = U3' Z_1' F; the column scaling by tau_e of cpp:311 commutes with the products)
// columns with tau_e == 1 or tau_e > 10000 are zeroed here and patched by fa_fixup_kernel.
// ------------------------------------------------------------------------------------
// rows [row0, row0 + r) of the r_total-row problem (row0 > 0: one rank's block of a row-sharded draw); b3, v, s1, s2
// are given already offset to row0, z is addressed with the global row.
__global__ void fa_v_kernel(const double* __restrict__ b3, long ldb, int r, int k, const double* __restrict__ F_h2,
                            const double* __restrict__ s1, const double* __restrict__ s2, VarSrc zsrc, int row0,
                            double* __restrict__ v, long ldv) {
    const int j = blockIdx.y;
    if (j >= k) return;
    const double h2 = F_h2[j];
    const double tau_e = 1.0 / (1.0 - h2), tau_u = 1.0 / h2;
    const bool special = (tau_e == 1.0) || (tau_e > 10000.0);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < r; i += gridDim.x * blockDim.x) {
        double out = 0.0;
        if (!special) {
            const double d = s2[i] * tau_u + s1[i] * tau_e;
            out = (b3[(long)j * ldb + i] * tau_e) / d + var_normal(zsrc, i + row0, j) / sqrt(d);
        }
        v[(long)j * ldv + i] = out;
    }
}
// rows [row0, row0 + r) (F, Fa given un-offset)
__global__ void fa_fixup_kernel(const double* __restrict__ F, long ldf, int n, int r, int row0, int k,
                                const double* __restrict__ F_h2, double* __restrict__ Fa, long lda) {
    const int j = blockIdx.y;
    if (j >= k) return;
    const double tau_e = 1.0 / (1.0 - F_h2[j]);
    if (tau_e == 1.0) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < r; i += gridDim.x * blockDim.x) Fa[(long)j * lda + row0 + i] = 0.0;
    } else if (tau_e > 10000.0) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < r; i += gridDim.x * blockDim.x)
            Fa[(long)j * lda + row0 + i] = (row0 + i < n) ? F[(long)j * ldf + row0 + i] : 0.0;
    }
}

// row-block exchange for the row-sharded replicated draws: pack rows [lo, lo+cnt) of C (rows x cols, ldc) into a dense
// chunk x cols block (zero padded), and scatter `world` such blocks back into C.
__global__ void pack_rows_kernel(const double* __restrict__ C, long ldc, long lo, long cnt, long chunk, long cols,
                                 double* __restrict__ block) {
    const long total = chunk * cols;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx % chunk, j = idx / chunk;
        block[idx] = (i < cnt) ? C[j * ldc + lo + i] : 0.0;
    }
}
__global__ void unpack_rows_kernel(const double* __restrict__ stage, long chunk, long cols, int world, int skip_rank, long rows,
                                   double* __restrict__ C, long ldc) {
    const long per = chunk * cols, total = per * world;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int q = (int)(idx / per);
        if (q == skip_rank) continue;
        const long e = idx - (long)q * per;
        const long i = e % chunk, j = e / chunk, row = (long)q * chunk + i;
        if (row < rows) C[j * ldc + row] = stage[idx];
    }
}

__global__ void tau_e_kernel(const double* __restrict__ F_h2, int k, double* __restrict__ tau_e) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < k) tau_e[j] = 1.0 / (1.0 - F_h2[j]);                          // cpp:147
}

// ------------------------------------------------------------------------------------
// sample_delta (BSFG_functions.R:272-308)
// ------------------------------------------------------------------------------------
// colsum[h] = sum_j Lambda_prec[j,h] * Lambda2[j,h]   (one block per column)
__global__ void delta_colsum_kernel(const double* __restrict__ LP, long ldp, const double* __restrict__ L2, long ldl,
                                    int p, int k, int square_l2, double* __restrict__ colsum) {
    __shared__ double red[32];
    const int h = blockIdx.x;
    double acc = 0.0;
    for (int j = threadIdx.x; j < p; j += blockDim.x) {
        double l = L2[(long)h * ldl + j];
        if (square_l2) l = l * l;
        acc += LP[(long)h * ldp + j] * l;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) colsum[h] = acc;
}
// the recursion of BSFG_functions.R:286-306, one warp, delta/tauh/colsum staged in shared memory.  After delta[h]
// changes, R recomputes tauh = cumprod(delta); entries l < h are unchanged and entries l >= h scale by new/old, which
// is what the lanes apply (relative rounding difference <= k ulp, far inside the 1e-9 contract); the tauh written
// out at the end is the plain sequential cumprod of the final delta.  The draws stay sequential.
// Loop order = R's 2:(k-1) (counts down when k == 2).  Shared memory: 3k doubles.
// The Gamma shapes do not depend on the data (a_1 + pk/2, a_2 + p(k-h+1)/2), so all standard-Gamma variates of the
// recursion are drawn up front by the whole block in parallel (shared memory gdraw[]); the sequential part then only
// divides by the rates.  Shared memory: 4k + 2 doubles.
__global__ void delta_recursion_kernel(double* __restrict__ delta_g, int k, const double* __restrict__ colsum_g,
                                       double n_genes, double d1s, double d1r, double d2s, double d2r,
                                       VarSrc gsrc, double* __restrict__ tauh_out, double* __restrict__ shapes,
                                       double* __restrict__ rates, int* __restrict__ ndraws) {
    if (blockIdx.x != 0) return;
    extern __shared__ double dsm[];
    double* tauh = dsm;
    double* delta = dsm + k;
    double* colsum = dsm + 2 * k;
    double* gdraw = dsm + 3 * k;                 // k + 2 entries, in draw order
    {
        const int hi0 = k - 1, step0 = (hi0 >= 2) ? 1 : -1;
        const int n_loop = (step0 > 0) ? (hi0 - 1) : (2 - hi0 + 1);      // |2:(k-1)| in R's sense
        for (int dr = threadIdx.x; dr < 1 + n_loop; dr += blockDim.x) {
            double shape;
            if (dr == 0) shape = d1s + 0.5 * n_genes * k;
            else {
                const int h = 2 + step0 * (dr - 1);                          // 1-based h of this draw
                shape = d2s + 0.5 * n_genes * (k - h + 1);
            }
            gdraw[dr] = var_gamma(gsrc, dr, 0, shape);
        }
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    for (int h = lane; h < k; h += 32) { delta[h] = delta_g[h]; colsum[h] = colsum_g[h]; }
    __syncwarp();
    if (lane == 0) {
        double cp = 1.0;
        for (int h = 0; h < k; h++) { cp *= delta[h]; tauh[h] = cp; }
    }
    __syncwarp();
    auto tail_sum = [&](int from) {              // sum_{l >= from} tauh[l] * colsum[l]
        double acc = 0.0;
        for (int l = from + lane; l < k; l += 32) acc += tauh[l] * colsum[l];
        return warp_sum(acc);
    };
    auto redraw = [&](int h0, double shape, double prior_rate, int draw) {   // 0-based h0
        const double ts = tail_sum(h0);
        const double old = delta[h0];
        const double rate = prior_rate + 0.5 * (1.0 / old) * ts;
        const double dnew = gdraw[draw] / rate;                              // same value on every lane
        const double ratio = dnew / old;
        __syncwarp();
        for (int l = h0 + lane; l < k; l += 32) tauh[l] *= ratio;
        if (lane == 0) {
            delta[h0] = dnew;
            if (shapes) shapes[draw] = shape;
            if (rates) rates[draw] = rate;
        }
        __syncwarp();
    };
    int draw = 0;
    redraw(0, d1s + 0.5 * n_genes * k, d1r, draw++);
    const int hi = k - 1;
    const int step = (hi >= 2) ? 1 : -1;
    for (int h = 2; (step > 0) ? (h <= hi) : (h >= hi); h += step)       // 1-based h
        redraw(h - 1, d2s + 0.5 * n_genes * (k - h + 1), d2r, draw++);
    for (int h = lane; h < k; h += 32) delta_g[h] = delta[h];
    if (lane == 0) {
        if (tauh_out) {
            double cp = 1.0;
            for (int h = 0; h < k; h++) { cp *= delta[h]; tauh_out[h] = cp; }
        }
        if (ndraws) *ndraws = draw;
    }
}

}  // namespace bsfg

// ====================================================================================
// kernels used only by the stateful sweep
// ====================================================================================
namespace bsfg {

// Lt (k x p, ld) -> Lam (p x k) and Lmsg = Lam * rp (rows)  (cpp:146); 32x32 smem tiles
__global__ void lt_to_lam_kernel(const double* __restrict__ Lt, long ldt, int k, int p, const double* __restrict__ rp,
                                 double* __restrict__ Lam, long ldl, double* __restrict__ Lmsg, long ldm) {
    __shared__ double tile[32][33];
    const int h0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        const int h = h0 + threadIdx.x, j = j0 + jj;
        if (h < k && j < p) tile[jj][threadIdx.x] = Lt[(long)j * ldt + h];
    }
    __syncthreads();
    for (int hh = threadIdx.y; hh < 32; hh += blockDim.y) {
        const int j = j0 + threadIdx.x, h = h0 + hh;
        if (h < k && j < p) {
            const double v = tile[threadIdx.x][hh];
            Lam[(long)h * ldl + j] = v;
            if (Lmsg) Lmsg[(long)h * ldm + j] = v * rp[j];
        }
    }
}

// Plamt[h,j] = LP[j,h] * tauh[h]   (fast_BSFG_sampler.R:257), transposing
__global__ void plam_kernel(const double* __restrict__ LP, long ldp, int p, int k, const double* __restrict__ tauh,
                            double* __restrict__ Plamt, long ldt) {
    __shared__ double tile[32][33];
    const int j0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
    for (int hh = threadIdx.y; hh < 32; hh += blockDim.y) {
        const int j = j0 + threadIdx.x, h = h0 + hh;
        if (j < p && h < k) tile[hh][threadIdx.x] = LP[(long)h * ldp + j] * tauh[h];
    }
    __syncthreads();
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        const int h = h0 + threadIdx.x, j = j0 + jj;
        if (j < p && h < k) Plamt[(long)j * ldt + h] = tile[threadIdx.x][jj];
    }
}

// Lambda_prec[j,h] ~ Gamma((df+1)/2, (df + Lambda[j,h]^2 tauh[h])/2)   (fast_BSFG_sampler.R:235-236)
// g is laid out like the reference's rgamma(p*k) stream: element j + h*p_total.
__global__ void lambda_prec_kernel(const double* __restrict__ Lam, long ldl, int p, int k, const double* __restrict__ tauh,
                                   double df, VarSrc gsrc, double* __restrict__ LP, long ldp,
                                   double* __restrict__ rate_out /* p x k dense or null */) {
    const long total = (long)p * k;
    const double shape = (df + 1.0) / 2.0;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % p), h = (int)(idx / p);
        const double l = Lam[(long)h * ldl + j];
        const double rate = (df + (l * l) * tauh[h]) / 2.0;
        double g;
        if (gsrc.ptr) g = gsrc.ptr[(long)h * gsrc.ld + j];
        else g = philox_gamma(gsrc.rs, (uint64_t)((long)h * gsrc.rows_total + j + gsrc.col_offset), shape);
        LP[(long)h * ldp + j] = g / rate;
        if (rate_out) rate_out[(long)h * p + j] = rate;
    }
}

// per-factor reductions for sample_delta and update_k:
//   colsum[h] = sum_j LP[j,h] Lam[j,h]^2  (BSFG_functions.R:287-292), cnt[h] = #{j: |Lam[j,h]| < eps} (R:326)
__global__ void factor_reduce_kernel(const double* __restrict__ LP, long ldp, const double* __restrict__ Lam, long ldl,
                                     int p, int k, double eps, double* __restrict__ colsum, double* __restrict__ cnt) {
    __shared__ double red[32];
    const int h = blockIdx.x;
    double acc = 0.0, c = 0.0;
    for (int j = threadIdx.x; j < p; j += blockDim.x) {
        const double l = Lam[(long)h * ldl + j];
        acc += LP[(long)h * ldp + j] * (l * l);
        c += (fabs(l) < eps) ? 1.0 : 0.0;
    }
    acc = block_sum(acc, red);
    c = block_sum(c, red);
    if (threadIdx.x == 0) { colsum[h] = acc; cnt[h] = c; }
}

// per-trait reductions + the two Gamma precision draws (fast_BSFG_sampler.R:239-245), one warp per trait.
//   SS_j = yy - 2 th.DtY - 2 l.FtY + q2 + 2 l.FDth + l.G2          (expanded ||y - Design_U th - F l||^2;
//          FDth = F'Design_U th_j, k values per trait instead of the br-vector Design_U'F l_j)
//   q2   = sum s2 th^2 (diag mode)  or  th.(M2 th) (direct mode);  q1 likewise with s1 / M1
//   resid_Y_prec ~ Gamma(a_r + n/2, b_r + SS/2);  E_a_prec ~ Gamma(a_e + r/2, b_e + (q1 - bX sum B^2)/2)
struct TraitReduceArgs {
    const double* theta; long ldth;
    const double* DtY; long lddy;
    const double* FDth; long ldfd;                        // (F' Design_U) theta, k x p
    const double* M1th; const double* M2th; long ldm;     // null in diag mode
    const double* s1; const double* s2;
    const double* Lt; long ldlt;
    const double* FtY; long ldfy;
    const double* G2; long ldg2;
    const double* Bfix; long ldb;
    const double* yy;
    int br, k, b, p;
    double n_half, r_half;
    double rr_shape, rr_rate, re_shape, re_rate, bX;
    double* rp; double* ep;
    double* rr_rate_out; double* re_rate_out;     // optional aux
    // second random effect (null W: none).  W r2 x p in original coordinates; Z2tY = Z_2'Y, Z2DUth = Z_2'Design_U theta,
    // ZZW = Z_2'Z_2 W, A2W = A_2_inv W (all r2 x p), FZ2W = F'Z_2 W (k x p)
    const double* W; long ldw; int r2;
    const double* Z2tY; long ldzy; const double* Z2DUth; long ldzd; const double* ZZW; long ldzz;
    const double* A2W; long lda2; const double* FZ2W; long ldfz;
    double* wp; double rw_shape, rw_rate, r2_half; double* rw_rate_out;
};
__global__ void trait_reduce_kernel(TraitReduceArgs a, VarSrc g_resid, VarSrc g_ea, VarSrc g_w) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= a.p) return;
    const double* th = a.theta + (long)j * a.ldth;
    const double* dy = a.DtY + (long)j * a.lddy;
    double a1 = 0, a2 = 0, q1 = 0, q2 = 0;
    if (a.M1th) {
        const double* m1 = a.M1th + (long)j * a.ldm;
        const double* m2 = a.M2th + (long)j * a.ldm;
        for (int i = lane; i < a.br; i += 32) {
            const double t = th[i];
            a1 += t * dy[i]; q1 += t * m1[i]; q2 += t * m2[i];
        }
    } else {
        for (int i = lane; i < a.br; i += 32) {
            const double t = th[i];
            a1 += t * dy[i]; q1 += a.s1[i] * (t * t); q2 += a.s2[i] * (t * t);
        }
    }
    double l1 = 0, l2 = 0;
    const double* lt = a.Lt + (long)j * a.ldlt;
    const double* fy = a.FtY + (long)j * a.ldfy;
    const double* g2 = a.G2 + (long)j * a.ldg2;
    const double* fd = a.FDth + (long)j * a.ldfd;
    for (int h = lane; h < a.k; h += 32) { const double l = lt[h]; l1 += l * fy[h]; l2 += l * g2[h]; a2 += l * fd[h]; }
    double bb = 0;
    for (int c = lane; c < a.b; c += 32) { const double v = a.Bfix[(long)j * a.ldb + c]; bb += v * v; }
    // second random effect: ||..- Z_2 w||^2 adds  -2 w.Z2'y + w.Z2'Z2 w + 2 w.Z2'Design_U th + 2 l.F'Z2 w   (R:239, 249)
    double w1 = 0, w2 = 0, w3 = 0, w4 = 0, l3 = 0;
    if (a.W) {
        const double* w = a.W + (long)j * a.ldw;
        const double* zy = a.Z2tY + (long)j * a.ldzy;
        const double* zd = a.Z2DUth + (long)j * a.ldzd;
        const double* zz = a.ZZW + (long)j * a.ldzz;
        const double* a2w = a.A2W + (long)j * a.lda2;
        for (int i = lane; i < a.r2; i += 32) {
            const double wi = w[i];
            w1 += wi * zy[i]; w2 += wi * zd[i]; w3 += wi * zz[i]; w4 += wi * a2w[i];
        }
        const double* fz = a.FZ2W + (long)j * a.ldfz;
        for (int h = lane; h < a.k; h += 32) l3 += lt[h] * fz[h];
        w1 = warp_sum(w1); w2 = warp_sum(w2); w3 = warp_sum(w3); w4 = warp_sum(w4); l3 = warp_sum(l3);
    }
    a1 = warp_sum(a1); a2 = warp_sum(a2); q1 = warp_sum(q1); q2 = warp_sum(q2);
    l1 = warp_sum(l1); l2 = warp_sum(l2); bb = warp_sum(bb);
    if (lane == 0) {
        double ss = a.yy[j] - 2.0 * a1 - 2.0 * l1 + q2 + 2.0 * a2 + l2;
        if (a.W) {
            ss += -2.0 * w1 + w3 + 2.0 * w2 + 2.0 * l3;
            const double rate_w = a.rw_rate + 0.5 * w4;
            a.wp[j] = var_gamma(g_w, j, 0, a.rw_shape + a.r2_half) / rate_w;
            if (a.rw_rate_out) a.rw_rate_out[j] = rate_w;
        }
        const double rate_r = a.rr_rate + 0.5 * ss;
        const double rate_e = a.re_rate + 0.5 * (q1 - a.bX * bb);
        a.rp[j] = var_gamma(g_resid, j, 0, a.rr_shape + a.n_half) / rate_r;
        a.ep[j] = var_gamma(g_ea, j, 0, a.re_shape + a.r_half) / rate_e;
        if (a.rr_rate_out) a.rr_rate_out[j] = rate_r;
        if (a.re_rate_out) a.re_rate_out[j] = rate_e;
    }
}

// Missing phenotypes (fast_BSFG_sampler.R:180-184): Y[i,j] = fit[i,j] + z[i,j] / sqrt(rp_j) where Y0[i,j] is NaN, Y0 elsewhere.
// fit = X B + F Lambda' + Z_1 E_a + Z_2 W.  One CTA per trait column.
__global__ void impute_missing_kernel(const double* __restrict__ Y0, long ld0, int n, int p, const double* __restrict__ fit,
                                      long ldf, const double* __restrict__ rp, VarSrc zsrc, double* __restrict__ Y, long ldy) {
    const int j = blockIdx.x;
    if (j >= p) return;
    const double sd = sqrt(1.0 / rp[j]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double y0 = Y0[(long)j * ld0 + i];
        Y[(long)j * ldy + i] = isnan(y0) ? fit[(long)j * ldf + i] + var_normal(zsrc, i, j) * sd : y0;
    }
}
// dst += src (same shape), used to add Z_1 E_a when Z_1 is the identity
__global__ void add_inplace_kernel(double* __restrict__ dst, long ldd, const double* __restrict__ src, long lds, long rows, long cols) {
    const long total = rows * cols;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx % rows, j = idx / rows;
        dst[j * ldd + i] += src[j * lds + i];
    }
}

// ---- parameter-expanded sampler (fast_BSFG_sampler_px.R) ---------------------------------------------------------
// sq[h] = sqrt(px[h]), isq[h] = 1/sqrt(px[h]), ipx[h] = 1/px[h]
__global__ void px_vectors_kernel(const double* __restrict__ px, int k, double* __restrict__ sq, double* __restrict__ isq,
                                  double* __restrict__ ipx) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h < k) { const double s = sqrt(px[h]); sq[h] = s; isq[h] = 1.0 / s; ipx[h] = 1.0 / px[h]; }
}
// tau_e = 1 / (px_factor * (1 - F_h2))                                     fast_BSFG_sampler_px.R:250
__global__ void tau_e_px_kernel(const double* __restrict__ F_h2, const double* __restrict__ px, int k, double* __restrict__ tau_e) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < k) tau_e[j] = 1.0 / (px[j] * (1.0 - F_h2[j]));
}
// px_factor[h] = 1 / rgamma(shape = a + n/2, rate = b + colSums((F_px - Z_1 F_a_px)^2)[h] / 2 / (1 - F_h2[h]))   (:260-263)
// one CTA per factor; ZFa = Z_1 F_a_px (n x k)
__global__ void px_factor_kernel(const double* __restrict__ Fpx, long ldf, const double* __restrict__ ZFa, long ldz, int n, int k,
                                 const double* __restrict__ F_h2, double shape, double prior_rate, VarSrc gsrc,
                                 double* __restrict__ px, double* __restrict__ rate_out) {
    __shared__ double red[32];
    const int h = blockIdx.x;
    if (h >= k) return;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const double e = Fpx[(long)h * ldf + i] - ZFa[(long)h * ldz + i]; acc += e * e; }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
        const double rate = prior_rate + 0.5 * acc / (1.0 - F_h2[h]);
        px[h] = 1.0 / (var_gamma(gsrc, h, 0, shape) / rate);
        if (rate_out) rate_out[h] = rate;
    }
}

// ---- posterior bookkeeping on the device (save_posterior_samples, BSFG_functions.R:375-408) ------------------
// trace slot (rows x kcap, dense) <- current (rows x k, ld), zero padded to kcap columns (the reference pads the sample
// vector with zeros when k grew, :384-391)
__global__ void trace_store_kernel(const double* __restrict__ cur, long ld, long rows, int k, int kcap, double* __restrict__ slot) {
    const long total = rows * kcap;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx % rows, h = idx / rows;
        slot[idx] = (h < k) ? cur[h * ld + i] : 0.0;
    }
}
// running mean exactly as the reference writes it: mean = (mean * (sp_num - 1) + x) / sp_num      (:403-405)
__global__ void mean_update_kernel(double* __restrict__ mean, long ldm, const double* __restrict__ x, long ldx, long rows, long cols,
                                   double sp_num) {
    const long total = rows * cols;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long i = idx % rows, j = idx / rows;
        mean[j * ldm + i] = (mean[j * ldm + i] * (sp_num - 1.0) + x[j * ldx + i]) / sp_num;
    }
}

// yy[j] = sum_i Y[i,j]^2, one warp per column
__global__ void colsumsq_kernel(const double* __restrict__ Y, long ld, int n, int p, double* __restrict__ out) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= p) return;
    double acc = 0;
    for (int i = lane; i < n; i += 32) { const double v = Y[(long)j * ld + i]; acc += v * v; }
    acc = warp_sum(acc);
    if (lane == 0) out[j] = acc;
}

// max |M[i,j] - (i==j ? d[i] : 0)| over a square matrix -> out[0] (atomicMax on the bit pattern; values >= 0)
__global__ void diag_residual_kernel(const double* __restrict__ M, long ld, int m, const double* __restrict__ d,
                                     unsigned long long* __restrict__ out) {
    double mx = 0.0;
    const long total = (long)m * m;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % m), j = (int)(idx / m);
        double v = M[(long)j * ld + i];
        if (i == j) v -= d[i];
        mx = fmax(mx, fabs(v));
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(mx));
}

// Pfull = blockdiag(bX * I_b, Ainv)  (fast_BSFG_sampler_init.R:289)
__global__ void build_prior_prec_kernel(const double* __restrict__ Ainv, long lda, int r, int b, double bX,
                                        double* __restrict__ P, long ldp) {
    const int br = b + r;
    const long total = (long)br * br;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % br), j = (int)(idx / br);
        double v = 0.0;
        if (i < b || j < b) v = (i == j) ? bX : 0.0;
        else v = Ainv[(long)(j - b) * lda + (i - b)];
        P[(long)j * ldp + i] = v;
    }
}

// ---- update_k (BSFG_functions.R:311-372) ----------------------------------------------------
// add arm, trait-indexed part: new column h = k (0-based) of Lambda_prec, Plam, Lambda
__global__ void addk_traits_kernel(int p, int h, double df, double tauh_new, VarSrc g_lp, VarSrc z_lam,
                                   double* __restrict__ LP, long ldp, double* __restrict__ Lam, long ldl,
                                   double* __restrict__ Lt, long ldlt, double* __restrict__ Plamt, long ldpt,
                                   double lam_mult /* 1, or 1/sqrt(px_new): the px sweep stores Lambda_px */) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p; j += gridDim.x * blockDim.x) {
        const double lp = var_gamma(g_lp, j, 0, df / 2.0) / (df / 2.0);     // R:334
        const double plam = lp * tauh_new;                                 // R:337
        const double l = var_normal(z_lam, j, 0) * sqrt(1.0 / plam) * lam_mult;   // R:338
        LP[(long)h * ldp + j] = lp;
        Lam[(long)h * ldl + j] = l;
        if (Lt) Lt[(long)j * ldlt + h] = l;
        Plamt[(long)j * ldpt + h] = plam;
    }
}
// add arm, individual-indexed part: F_a[,k] = z sqrt(h2); F[,k] = Z_1 F_a[,k] + z sqrt(1-h2)   (R:339-341)
__global__ void addk_fa_kernel(int r, int h, double h2, VarSrc z_fa, double* __restrict__ Fa, long lda) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < r; i += gridDim.x * blockDim.x)
        Fa[(long)h * lda + i] = var_normal(z_fa, i, 0) * sqrt(h2);
}
__global__ void addk_f_kernel(int n, int r, int h, double h2, const double* __restrict__ Z1, long ldz,
                              const double* __restrict__ Fa, long lda, VarSrc z_f, double* __restrict__ F, long ldf,
                              double f_mult /* 1, or sqrt(px_new): the px sweep stores F_px */) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double m;
        if (Z1) {
            m = 0.0;
            for (int c = 0; c < r; c++) m += Z1[(long)c * ldz + i] * Fa[(long)h * lda + c];
        } else {
            m = Fa[(long)h * lda + i];
        }
        F[(long)h * ldf + i] = (m + var_normal(z_f, i, 0) * sqrt(1.0 - h2)) * f_mult;
    }
}
// drop arm: keep columns map[0..knew) (ascending) of a column-major matrix, in place
__global__ void compact_cols_kernel(double* __restrict__ A, long ld, int rows, int knew, const int* __restrict__ map) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x)
        for (int h = 0; h < knew; h++) {
            const int src = map[h];
            if (src != h) A[(long)h * ld + i] = A[(long)src * ld + i];
        }
}
// same for the rows of a (k x p) column-major matrix
__global__ void compact_rows_kernel(double* __restrict__ A, long ld, int p, int knew, const int* __restrict__ map) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p; j += gridDim.x * blockDim.x)
        for (int h = 0; h < knew; h++) {
            const int src = map[h];
            if (src != h) A[(long)j * ld + h] = A[(long)j * ld + src];
        }
}

}  // namespace bsfg
