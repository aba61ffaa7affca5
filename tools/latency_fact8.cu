// cycles of one redundant-per-thread 8x8 Cholesky (the phase-2 body of chol_blocked_kernel), alone on an SM
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double cb_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    return y;
}
template <int MODE>
__global__ void fact8(double* out, long long* cyc, int iters) {
    __shared__ double D[64];
    __shared__ double P[128 * 8];
    if (threadIdx.x < 64) { int r = threadIdx.x >> 3, c = threadIdx.x & 7; D[threadIdx.x] = (r == c) ? 10.0 + r : 0.1 * (r + c); }
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) P[i] = 0.01 * i;
    __syncthreads();
    long long t0 = clock64();
    double carry = 0;
    for (int it = 0; it < iters; it++) {
        double L[8][8];
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c <= r; c++) L[r][c] = D[r * 8 + c];
        L[0][0] += carry * 1e-30;
        double inv[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const double dd = L[c][c];
            inv[c] = cb_rsqrt(dd);
            L[c][c] = dd * inv[c];
#pragma unroll
            for (int r = c + 1; r < 8; r++) L[r][c] *= inv[c];
#pragma unroll
            for (int r = c + 1; r < 8; r++)
#pragma unroll
                for (int l = c + 1; l <= r; l++) L[r][l] -= L[r][c] * L[l][c];
        }
        if (MODE == 0) { carry = L[7][7] + inv[3]; }
        if (MODE == 1) {     // + panel substitution of this thread's row
            double x[8];
#pragma unroll
            for (int c = 0; c < 8; c++) x[c] = P[threadIdx.x * 8 + c];
#pragma unroll
            for (int c = 0; c < 8; c++) {
                x[c] *= inv[c];
#pragma unroll
                for (int l = c + 1; l < 8; l++) x[l] -= x[c] * L[l][c];
            }
#pragma unroll
            for (int c = 0; c < 8; c++) P[threadIdx.x * 8 + c] = x[c] * 0.5 + 1.0;
            carry = x[7];
        }
        if (MODE == 2) { __syncthreads(); carry = L[7][7]; if (threadIdx.x == 0) D[1] = 0.1 + 1e-20 * carry; __syncthreads(); }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = carry;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024 * 256); cudaMalloc(&cyc, 8);
    const int iters = 512;
    for (int threads : {32, 128}) {
#define RUN(M, name) { fact8<M><<<1, threads>>>(out, cyc, iters); fact8<M><<<1, threads>>>(out, cyc, iters); long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
        printf("%3d threads  %-40s %8.1f cycles / iteration\n", threads, name, (double)h / iters); }
        RUN(0, "8x8 factorisation (36 LDS + chain)")
        RUN(1, "factorisation + row substitution + smem")
        RUN(2, "factorisation + 2 barriers")
    }
    for (int ctas : {2, 4}) {
        fact8<1><<<ctas * 148, 128>>>(out, cyc, iters); fact8<1><<<ctas * 148, 128>>>(out, cyc, iters);
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%d CTAs/SM x 128 threads: factorisation + substitution %8.1f cycles / iteration\n", ctas, (double)h / iters);
    }
    return 0;
}
