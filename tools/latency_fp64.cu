// Dependent-chain latencies of the FP64 building blocks of the batched Cholesky on sm_100a (one warp per CTA, clock64).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/latency_fp64 tools/latency_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    return y;
}
template <int OP>
__global__ void lat(double* out, long long* cyc, int iters, double seed) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 8 + 8) % 1024);
    __syncthreads();
    double x = seed + threadIdx.x * 1e-12, a = 1.0000001, b = 1e-9, c1 = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (OP == 0) x = fma(x, a, b);
        if (OP == 1) { asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(x), "+d"(c1) : "d"(a), "d"(b)); }
        if (OP == 2) x = rsqrt(x) + 1.5;
        if (OP == 3) x = fast_rsqrt(x) + 1.5;
        if (OP == 4) x = 1.0 / x + 1.5;
        if (OP == 5) x = sqrt(x) + 1.5;
        if (OP == 6) x = sm[(int)x];                       // dependent LDS.64 (+ F2I)
        if (OP == 7) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
        if (OP == 8) { __syncthreads(); }
        if (OP == 9) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31]; __syncwarp(); }   // smem round trip
        if (OP == 10) { asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(x), "+d"(c1) : "d"(a), "d"(b));
                        x = x * a; }                      // DMMA -> DMUL dependent (pipe switch)
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + c1;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024 * 256); cudaMalloc(&cyc, 8);
    const char* nm[] = {"DFMA dep", "DMMA dep (acc)", "rsqrt(double)+add", "fast_rsqrt+add", "1/x + add", "sqrt + add", "LDS.64 dep (+F2I)", "SHFL f64", "__syncthreads (4 warps)",
                        "STS->LDS round trip", "DMMA -> DMUL dep"};
    const int iters = 4096;
    for (int threads : {32, 128}) {
        printf("-- %d threads per CTA, 1 CTA --\n", threads);
#define RUN(OP)                                                                                        \
    {                                                                                                  \
        lat<OP><<<1, threads>>>(out, cyc, iters, 2.0); lat<OP><<<1, threads>>>(out, cyc, iters, 2.0);   \
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                    \
        printf("%-26s %7.1f cycles / op\n", nm[OP], (double)h / iters);                                 \
    }
        RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10)
    }
    return 0;
}
