// Shared host/device helpers for libbsfg_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>

#include "../../include/bsfg_b200.h"

namespace bsfg {

// ------------------------------------------------------------------------------------
// error plumbing: every failure becomes a negative bsfg_status + a thread-local message,
// the C-ABI analogue of Armadillo's exceptions -> Rcpp stop() (BSFG_functions_c.cpp, all exports).
// ------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const char* fmt, ...);
const char* get_last_error();

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error(code, buf);
}

#define BSFG_CUDA(expr)                                                                     \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess)                                                             \
            ::bsfg::fail(BSFG_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                         __FILE__, __LINE__);                                               \
    } while (0)

#define BSFG_CHECK_LAUNCH() BSFG_CUDA(cudaGetLastError())

// cudaMemset runs on the legacy default stream and may return before the fill has happened; the library works on a
// non-blocking stream, which the legacy stream does NOT order against.  Every allocation-time clear therefore waits
// for the fill, or a later kernel/copy on the library stream could be overwritten by a late memset.
inline void memset_sync(void* p, int value, size_t bytes) {
    BSFG_CUDA(cudaMemset(p, value, bytes));
    BSFG_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
}

inline long round_up(long x, long m) { return (x + m - 1) / m * m; }
inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------
// Device matrix: column-major FP64, leading dimension padded to an even number of doubles so
// every column starts 16-byte aligned (a TMA global-stride requirement).  Padding rows are zero.
// ------------------------------------------------------------------------------------
struct DMat {
    double* p = nullptr;
    long rows = 0, cols = 0, ld = 0;
    bool owner = false;
    DMat() = default;
    DMat(long r, long c) { alloc(r, c); }
    DMat(const DMat&) = delete;
    DMat& operator=(const DMat&) = delete;
    DMat(DMat&& o) noexcept { *this = std::move(o); }
    DMat& operator=(DMat&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p; rows = o.rows; cols = o.cols; ld = o.ld; owner = o.owner;
            o.p = nullptr; o.owner = false; o.rows = o.cols = o.ld = 0;
        }
        return *this;
    }
    ~DMat() { release(); }
    void release() {
        if (owner && p) cudaFree(p);
        p = nullptr; owner = false;
    }
    void alloc(long r, long c) {
        release();
        rows = r; cols = c; ld = round_up(r > 0 ? r : 1, 2);
        size_t bytes = sizeof(double) * (size_t)ld * (size_t)(c > 0 ? c : 1);
        BSFG_CUDA(cudaMalloc(&p, bytes));
        owner = true;
        memset_sync(p, 0, bytes);
    }
    // capacity-preserving logical resize (cols only); used for factor-indexed state under update_k
    size_t bytes() const { return sizeof(double) * (size_t)ld * (size_t)(cols > 0 ? cols : 1); }
    void upload(const double* host, cudaStream_t st) {   // host is dense column-major rows x cols
        if (rows == 0 || cols == 0) return;
        BSFG_CUDA(cudaMemcpy2DAsync(p, ld * sizeof(double), host, rows * sizeof(double),
                                    rows * sizeof(double), cols, cudaMemcpyHostToDevice, st));
    }
    void download(double* host, cudaStream_t st) const {
        if (rows == 0 || cols == 0) return;
        BSFG_CUDA(cudaMemcpy2DAsync(host, rows * sizeof(double), p, ld * sizeof(double),
                                    rows * sizeof(double), cols, cudaMemcpyDeviceToHost, st));
    }
    void zero(cudaStream_t st) { BSFG_CUDA(cudaMemsetAsync(p, 0, bytes(), st)); }
};

struct DVec {
    double* p = nullptr;
    long n = 0;
    DVec() = default;
    explicit DVec(long n_) { alloc(n_); }
    DVec(const DVec&) = delete;
    DVec& operator=(const DVec&) = delete;
    DVec(DVec&& o) noexcept { p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    DVec& operator=(DVec&& o) noexcept {
        if (this != &o) { if (p) cudaFree(p); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DVec() { if (p) cudaFree(p); }
    void alloc(long n_) {
        if (p) cudaFree(p);
        n = n_;
        size_t bytes = sizeof(double) * (size_t)(n > 0 ? n : 1);
        BSFG_CUDA(cudaMalloc(&p, bytes));
        memset_sync(p, 0, bytes);
    }
    void upload(const double* host, cudaStream_t st) {
        if (n) BSFG_CUDA(cudaMemcpyAsync(p, host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    void download(double* host, cudaStream_t st) const {
        if (n) BSFG_CUDA(cudaMemcpyAsync(host, p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
};

// ------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al. 2011), usable on host and device.  Streams are
// addressed as (seed; site, iteration, element) so a draw does not depend on how traits are
// sharded over GPUs and replicated draws are bit-identical on every rank.
// ------------------------------------------------------------------------------------
struct Philox {
    uint32_t key[2];
    __host__ __device__ Philox(uint64_t seed) { key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32); }
    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
        uint64_t prod = (uint64_t)a * (uint64_t)b;
        hi = (uint32_t)(prod >> 32);
        lo = (uint32_t)prod;
    }
    __host__ __device__ inline void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t out[4]) const {
        uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
        for (int r = 0; r < 10; r++) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, c0, hi0, lo0);
            mulhilo(0xCD9E8D57u, c2, hi1, lo1);
            uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

// two 32-bit words -> U(0,1) double strictly inside (0,1): the top 52 bits become the mantissa of a double in [1, 2)
// (two shifts and an OR instead of a 64-bit integer -> double conversion), minus 1, plus half a grid step:
// values (m + 1/2) 2^-52, m = 0 .. 2^52 - 1.
__host__ __device__ inline double u01_from_bits(uint32_t hi, uint32_t lo) {
    const uint32_t mh = 0x3FF00000u | (hi >> 12);
    const uint32_t ml = (hi << 20) | (lo >> 12);
#ifdef __CUDA_ARCH__
    const double d = __hiloint2double((int)mh, (int)ml);
#else
    const uint64_t bits = ((uint64_t)mh << 32) | ml;
    double d;
    memcpy(&d, &bits, sizeof d);
#endif
    return (d - 1.0) + 1.1102230246251565e-16;      // + 2^-53
}

// 1/sqrt(x) to ~1 ulp without the special-case handling of CUDA's rsqrt(double): the hardware seed (MUFU.RSQ64H,
// relative error 2^-22) and ONE third-order step y <- y (1 + e/2 + 3 e^2/8), e = 1 - x y^2 (error (5/16) e^3 ~ 2^-67):
// four dependent FP64 operations after the seed.  x > 0 finite and normal is the caller's business.
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    const double p = fma(e, 0.375, 0.5);
    const double q = y * e;
    return fma(q, p, y);
}

// Box-Muller radius sqrt(-2 ln u) for u in (0, 1): on the device x * rsqrt(x) with the 5-operation reciprocal square root
// (x = -2 ln u lies in [2e-16, 74]: positive, normal), on the host the library call.
__host__ __device__ inline double bm_radius(double u) {
    const double x = -2.0 * log(u);
#ifdef __CUDA_ARCH__
    return x * fast_rsqrt(x);
#else
    return sqrt(x);
#endif
}

// draw sites (the `site` word of the Philox counter); one per reference RNG call site
enum Site : uint32_t {
    SITE_Z_LAMBDA = 1,     // randn(k,p)            BSFG_functions_c.cpp:110
    SITE_Z_MEANS = 2,      // randn(br,p)           :68
    SITE_U_H2 = 3,         // randu(1) per factor   :232
    SITE_Z_FA = 4,         // rnorm(r*k)            :316-318
    SITE_Z_F = 5,          // randn(n,k)            :154
    SITE_G_LAMBDA_PREC = 6,  // rgamma(p*k)         fast_BSFG_sampler.R:236
    SITE_G_RESID = 7,      // rgamma(p)             :240
    SITE_G_EA = 8,         // rgamma(p)             :245
    SITE_G_DELTA = 9,      // rgamma(1) x k         BSFG_functions.R:293,303
    SITE_U_K = 10,         // runif(1)              BSFG_functions.R:325
    SITE_K_ADD = 11,       // add-a-factor draws    BSFG_functions.R:334-341
    SITE_U_PREC = 12,      // randu(1) per trait    BSFG_functions_c.cpp:282
    SITE_Z_MEANS_W = 13,   // randn(r2,p)           :68 (second call, fast_BSFG_sampler.R:214)
    SITE_G_W = 14,         // rgamma(p)             fast_BSFG_sampler.R:249
    SITE_Z_Y = 15,         // rnorm(p*n)            fast_BSFG_sampler.R:182 (missing phenotypes)
    SITE_G_PX = 16,        // rgamma(k)             fast_BSFG_sampler_px.R:263
    // the vector draws of the add-a-factor arm each have their own site: Gamma draws address Philox blocks by element and
    // normal draws by element PAIR, so two families on one site would share blocks (SITE_K_ADD keeps the three scalars)
    SITE_K_ADD_LP = 17,    // rgamma(p)             BSFG_functions.R:336
    SITE_K_ADD_ZL = 18,    // rnorm(p)              :338
    SITE_K_ADD_ZFA = 19,   // rnorm(r)              :340
    SITE_K_ADD_ZF = 20     // rnorm(n)              :341
};

struct RngStream {
    uint64_t seed;
    uint32_t site;
    uint32_t iter;
};

// N(0,1) for element `idx` of a stream: elements 2q and 2q+1 are the cosine and sine branches of one Box-Muller
// transform on Philox block q, so a thread that owns an aligned pair pays for one block; element-addressed, so the
// value is independent of launch geometry and of how traits are sharded.
__host__ __device__ inline void philox_normal_pair(const RngStream& s, uint64_t pair, double* z_even, double* z_odd) {
    Philox ph(s.seed);
    uint32_t o[4];
    ph((uint32_t)pair, (uint32_t)(pair >> 32), s.site, s.iter, o);
    const double u1 = u01_from_bits(o[0], o[1]);
    const double u2 = u01_from_bits(o[2], o[3]);
    const double r = bm_radius(u1);
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    *z_even = r * cs;
    *z_odd = r * sn;
}
__host__ __device__ inline double philox_normal(const RngStream& s, uint64_t idx) {
    Philox ph(s.seed);
    uint32_t o[4];
    const uint64_t pair = idx >> 1;
    ph((uint32_t)pair, (uint32_t)(pair >> 32), s.site, s.iter, o);
    const double u1 = u01_from_bits(o[0], o[1]);
    const double u2 = u01_from_bits(o[2], o[3]);
    const double r = bm_radius(u1);
    return (idx & 1) ? r * sinpi(2.0 * u2) : r * cospi(2.0 * u2);
}

__host__ __device__ inline double philox_uniform(const RngStream& s, uint64_t idx) {
    Philox ph(s.seed);
    uint32_t o[4];
    ph((uint32_t)idx, (uint32_t)(idx >> 32), s.site, s.iter, o);
    return u01_from_bits(o[0], o[1]);
}

// Standard Gamma(shape,1), shape >= 1: Marsaglia-Tsang (2000) squeeze, one Philox block per
// attempt (attempt number folded into the top bits of the element index).
__host__ __device__ inline double philox_gamma(const RngStream& s, uint64_t idx, double shape) {
    Philox ph(s.seed);
    double boost = 1.0;
    double a = shape;
    uint32_t o[4];
    if (a < 1.0) {   // Gamma(a) = Gamma(a+1) * U^(1/a)
        ph((uint32_t)idx, (uint32_t)(idx >> 32) | 0x80000000u, s.site, s.iter, o);
        boost = pow(u01_from_bits(o[0], o[1]), 1.0 / a);
        a += 1.0;
    }
    const double d = a - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    for (uint32_t attempt = 0; attempt < 64; attempt++) {
        ph((uint32_t)idx, (uint32_t)(idx >> 32) | (attempt << 24), s.site, s.iter, o);
        double u1 = u01_from_bits(o[0], o[1]);
        double u2 = u01_from_bits(o[2], o[3]);
        double x = bm_radius(u1) * cospi(2.0 * u2);
        // the acceptance uniform comes from a second block so it is independent of x
        uint32_t o2[4];
        ph((uint32_t)idx, (uint32_t)(idx >> 32) | (attempt << 24) | 0x40000000u, s.site, s.iter, o2);
        double u = u01_from_bits(o2[0], o2[1]);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return boost * d * v;
    }
    return boost * d;   // unreachable in practice (acceptance > 0.95 per attempt)
}

}  // namespace bsfg
