// RcppArmadillo.h -- STAND-IN header, TEST INFRASTRUCTURE ONLY (never part of the product library).
//
// Purpose: compile the reference's own `/root/reference/R_BSFG/BSFG_functions_c.cpp` UNMODIFIED, from where it
// lies, into `oracle/_ref/libbsfg_reference.so` (recipe: oracle/ref_shim/Makefile), so that the NumPy oracle and the
// CUDA path are checked against the reference's code itself and not against a restatement of it.
//
// The reference includes <RcppArmadillo.h> (cpp:1) = R headers + Rcpp + Armadillo; none of them exist in this image.
// This file supplies the closed vocabulary that one source file uses, with the documented Armadillo / Rcpp semantics:
//
//   arma::  mat / vec / rowvec / uvec (column-major, FP64), .n_rows/.n_cols/.n_elem, (i) / (i,j), .t(), .col(j), .row(j)
//           + - % / (element-wise, scalar on either side), * (matrix product: dgemm/dgemv of the BLAS below),
//           trans, repmat, reshape, diagmat, ones, zeros, randn, randu, log, exp, sqrt, sum (vector -> scalar, matrix
//           -> along dim), max, cumsum(X, dim), operator> (-> umat), find, chol (upper, throws "chol(): decomposition
//           failed"), solve (triangular systems by substitution, general by LU -- what Armadillo's solve() does after
//           its structure detection), inv, svd (full, divide and conquer = Armadillo's default "dc").
//   Rcpp::  List (create / ["name"]), _["name"] = value, as<mat>/as<vec>, Environment("package:stats")["rnorm"],
//           Function call.
//
// Evaluation is eager (every operator returns a dense object); the two places where Armadillo's expression templates
// change the COST materially are kept: X.t() is a lazy view that products and solve() consume as a BLAS/LAPACK
// transpose flag, and objects are moved, not copied, out of functions.
//
// Linear algebra back end: the OpenBLAS + LAPACK bundled with SciPy (LP64, symbols prefixed `scipy_`) -- the same class
// of library R links Armadillo to.
//
// Random numbers: arma::randn / arma::randu / R's stats::rnorm draw from INJECTABLE streams (refshim::normals,
// refshim::uniforms; set through ref_exports.cpp) consumed in Armadillo's/R's column-major fill order, so a test can
// hand the reference code and the device path the very same N(0,1)/U(0,1) variates.  With no stream installed they
// fall back to an internal generator (timing runs).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

extern "C" {
void scipy_cblas_dgemm(int, int, int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
void scipy_cblas_dgemv(int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
int scipy_LAPACKE_dpotrf(int, char, int, double*, int);
int scipy_LAPACKE_dtrtrs(int, char, char, char, int, int, const double*, int, double*, int);
int scipy_LAPACKE_dgesv(int, int, int, double*, int, int*, double*, int);
int scipy_LAPACKE_dgetrf(int, int, int, double*, int, int*);
int scipy_LAPACKE_dgetri(int, int, double*, int, const int*);
int scipy_LAPACKE_dgesdd(int, char, int, int, double*, int, double*, double*, int, double*, int);
}

namespace refshim {

// ---- injectable variate streams -----------------------------------------------------------------------------------
struct Stream {
    const double* p = nullptr;
    size_t n = 0, pos = 0;
};
inline Stream normals, uniforms;
inline uint64_t fallback_state[2] = {0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};

inline uint64_t fallback_next() {  // xoroshiro128+
    uint64_t s0 = fallback_state[0], s1 = fallback_state[1], r = s0 + s1;
    s1 ^= s0;
    fallback_state[0] = ((s0 << 24) | (s0 >> 40)) ^ s1 ^ (s1 << 16);
    fallback_state[1] = (s1 << 37) | (s1 >> 27);
    return r;
}
inline double fallback_unif() { return ((fallback_next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

inline double next_uniform() {
    if (uniforms.p) {
        if (uniforms.pos >= uniforms.n) throw std::runtime_error("refshim: injected U(0,1) stream exhausted");
        return uniforms.p[uniforms.pos++];
    }
    return fallback_unif();
}
inline double next_normal() {
    if (normals.p) {
        if (normals.pos >= normals.n) throw std::runtime_error("refshim: injected N(0,1) stream exhausted");
        return normals.p[normals.pos++];
    }
    return std::sqrt(-2.0 * std::log(fallback_unif())) * std::cos(6.283185307179586 * fallback_unif());
}

enum { ColMajor = 102, NoTrans = 111, Trans = 112 };

}  // namespace refshim

namespace arma {

typedef unsigned long long uword;

// Kind: 0 = mat, 1 = vec (n x 1), 2 = rowvec (1 x n).  The kind decides what Armadillo's shape-dependent functions do
// (sum / max of a vector expression give a scalar, of a matrix a row of column results).
template <int K> struct Dense;
template <int K> struct Tr;
typedef Dense<0> mat;
typedef Dense<1> vec;
typedef Dense<2> rowvec;
typedef Dense<1> colvec;

template <int K> struct TransKind { static const int value = K == 1 ? 2 : (K == 2 ? 1 : 0); };
template <int A, int B> struct SameKind { static const int value = (A == B) ? A : 0; };
template <int A, int B> struct ProdKind { static const int value = (B == 1 && A != 2) ? 1 : ((A == 2 && B != 1) ? 2 : 0); };

struct colview;
struct rowview;

template <int K> struct Dense {
    uword n_rows = 0, n_cols = 0, n_elem = 0;
    std::unique_ptr<double[]> mem;

    Dense() {
        if (K == 1) n_cols = 1;
        if (K == 2) n_rows = 1;
    }
    Dense(uword r, uword c) { init(r, c); }
    explicit Dense(uword n) { K == 2 ? init(1, n) : init(n, 1); }
    Dense(const Dense& o) { copy_from(o); }
    Dense(Dense&& o) noexcept : n_rows(o.n_rows), n_cols(o.n_cols), n_elem(o.n_elem), mem(std::move(o.mem)) { o.reset(); }
    template <int J> Dense(const Dense<J>& o) { check_layout(o.n_rows, o.n_cols); copy_from(o); }
    template <int J> Dense(Dense<J>&& o) {
        check_layout(o.n_rows, o.n_cols);
        n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem; mem = std::move(o.mem);
        o.reset();
    }
    template <int J> Dense(const Tr<J>& t);  // materialise a lazy transpose
    Dense& operator=(const Dense& o) { if ((const void*)this != (const void*)&o) copy_from(o); return *this; }
    Dense& operator=(Dense&& o) noexcept {
        n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem; mem = std::move(o.mem);
        o.reset();
        return *this;
    }

    void reset() { n_rows = (K == 2); n_cols = (K == 1); n_elem = 0; mem.reset(); }
    void init(uword r, uword c) {
        check_layout(r, c);
        n_rows = r; n_cols = c; n_elem = r * c;
        mem.reset(n_elem ? new double[n_elem] : nullptr);
    }
    static void check_layout(uword r, uword c) {
        if (K == 1 && c != 1) throw std::logic_error("Mat::init(): requested size is not compatible with column vector layout");
        if (K == 2 && r != 1) throw std::logic_error("Mat::init(): requested size is not compatible with row vector layout");
    }
    template <int J> void copy_from(const Dense<J>& o) {
        if (n_elem != o.n_elem || !mem) mem.reset(o.n_elem ? new double[o.n_elem] : nullptr);
        n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem;
        if (n_elem) std::memcpy(mem.get(), o.mem.get(), n_elem * sizeof(double));
    }
    void fill(double v) { std::fill(mem.get(), mem.get() + n_elem, v); }

    double* memptr() { return mem.get(); }
    const double* memptr() const { return mem.get(); }
    double& operator()(uword i) { bounds(i < n_elem); return mem[i]; }
    const double& operator()(uword i) const { bounds(i < n_elem); return mem[i]; }
    double& operator()(uword i, uword j) { bounds(i < n_rows && j < n_cols); return mem[i + j * n_rows]; }
    const double& operator()(uword i, uword j) const { bounds(i < n_rows && j < n_cols); return mem[i + j * n_rows]; }
    static void bounds(bool ok) { if (!ok) throw std::logic_error("Mat::operator(): index out of bounds"); }

    Tr<K> t() const;
    colview col(uword j);
    const colview col(uword j) const;
    rowview row(uword i);
    const rowview row(uword i) const;
};

// X.t(): a view; products / solve() turn it into a transpose flag, anything else materialises it.
template <int K> struct Tr {
    const Dense<K>& src;
    explicit Tr(const Dense<K>& s) : src(s) {}
    Dense<TransKind<K>::value> eval() const {
        Dense<TransKind<K>::value> out(src.n_cols, src.n_rows);
        const uword R = src.n_rows, C = src.n_cols;
        const double* a = src.memptr();
        double* o = out.memptr();
        if (R == 1 || C == 1) {
            if (src.n_elem) std::memcpy(o, a, src.n_elem * sizeof(double));
        } else {
            const uword T = 32;  // blocked: the n x p transposes of the prec / h2 functions are memory-bound
            for (uword jb = 0; jb < C; jb += T)
                for (uword ib = 0; ib < R; ib += T)
                    for (uword j = jb; j < std::min(C, jb + T); ++j)
                        for (uword i = ib; i < std::min(R, ib + T); ++i) o[j + i * C] = a[i + j * R];
        }
        return out;
    }
};
template <int K> template <int J> Dense<K>::Dense(const Tr<J>& t) {
    Dense<TransKind<J>::value> e = t.eval();
    check_layout(e.n_rows, e.n_cols);
    n_rows = e.n_rows; n_cols = e.n_cols; n_elem = e.n_elem; mem = std::move(e.mem);
}
template <int K> Tr<K> Dense<K>::t() const { return Tr<K>(*this); }
template <int K> Dense<TransKind<K>::value> trans(const Dense<K>& x) { return Tr<K>(x).eval(); }

// .col(j) / .row(j): a copy of the column/row that writes back to its parent on assignment (subview semantics).
struct colview : vec {
    double* dst;  // parent storage of this column (contiguous)
    colview(double* d, uword n) : vec(n), dst(d) { if (n) std::memcpy(memptr(), d, n * sizeof(double)); }
    template <int J> colview& operator=(const Dense<J>& o) { return assign(o.memptr(), o.n_rows, o.n_cols); }
    colview& operator=(const colview& o) { return assign(o.memptr(), o.n_rows, o.n_cols); }
    template <int J> colview& operator=(const Tr<J>& o) { Dense<TransKind<J>::value> e = o.eval(); return assign(e.memptr(), e.n_rows, e.n_cols); }
    colview& assign(const double* src, uword r, uword c) {
        if (r != n_rows || c != 1) throw std::logic_error("copy into submatrix: incompatible matrix dimensions");
        if (r) { std::memcpy(dst, src, r * sizeof(double)); std::memcpy(memptr(), src, r * sizeof(double)); }
        return *this;
    }
};
struct rowview : rowvec {
    double* dst;  // parent element (i, 0)
    uword ld;
    rowview(double* d, uword n, uword ld_) : rowvec(n), dst(d), ld(ld_) { for (uword j = 0; j < n; ++j) mem[j] = d[j * ld]; }
    template <int J> rowview& operator=(const Dense<J>& o) { return assign(o.memptr(), o.n_rows, o.n_cols); }
    rowview& operator=(const rowview& o) { return assign(o.memptr(), o.n_rows, o.n_cols); }
    template <int J> rowview& operator=(const Tr<J>& o) { Dense<TransKind<J>::value> e = o.eval(); return assign(e.memptr(), e.n_rows, e.n_cols); }
    rowview& assign(const double* src, uword r, uword c) {
        if (r != 1 || c != n_cols) throw std::logic_error("copy into submatrix: incompatible matrix dimensions");
        for (uword j = 0; j < c; ++j) { dst[j * ld] = src[j]; mem[j] = src[j]; }
        return *this;
    }
};
template <int K> colview Dense<K>::col(uword j) { bounds(j < n_cols); return colview(mem.get() + j * n_rows, n_rows); }
template <int K> const colview Dense<K>::col(uword j) const { bounds(j < n_cols); return colview(mem.get() + j * n_rows, n_rows); }
template <int K> rowview Dense<K>::row(uword i) { bounds(i < n_rows); return rowview(mem.get() + i, n_cols, n_rows); }
template <int K> const rowview Dense<K>::row(uword i) const { bounds(i < n_rows); return rowview(mem.get() + i, n_cols, n_rows); }

// ---- element-wise arithmetic --------------------------------------------------------------------------------------
inline void same_size(uword r1, uword c1, uword r2, uword c2, const char* what) {
    if (r1 != r2 || c1 != c2)
        throw std::logic_error(std::string(what) + ": incompatible matrix dimensions: " + std::to_string(r1) + "x" + std::to_string(c1) +
                               " and " + std::to_string(r2) + "x" + std::to_string(c2));
}

#define REFSHIM_EWISE(OP, NAME)                                                                                          \
    template <int A, int B> Dense<SameKind<A, B>::value> operator OP(const Dense<A>& x, const Dense<B>& y) {             \
        same_size(x.n_rows, x.n_cols, y.n_rows, y.n_cols, NAME);                                                         \
        Dense<SameKind<A, B>::value> o(x.n_rows, x.n_cols);                                                              \
        const double *a = x.memptr(), *b = y.memptr();                                                                   \
        double* c = o.memptr();                                                                                          \
        for (uword i = 0; i < x.n_elem; ++i) c[i] = REFSHIM_APPLY(a[i], b[i], OP);                                        \
        return o;                                                                                                        \
    }                                                                                                                    \
    template <int A> Dense<A> operator OP(const Dense<A>& x, double s) {                                                 \
        Dense<A> o(x.n_rows, x.n_cols);                                                                                  \
        const double* a = x.memptr();                                                                                    \
        double* c = o.memptr();                                                                                          \
        for (uword i = 0; i < x.n_elem; ++i) c[i] = REFSHIM_APPLY(a[i], s, OP);                                           \
        return o;                                                                                                        \
    }                                                                                                                    \
    template <int A> Dense<A> operator OP(double s, const Dense<A>& x) {                                                 \
        Dense<A> o(x.n_rows, x.n_cols);                                                                                  \
        const double* a = x.memptr();                                                                                    \
        double* c = o.memptr();                                                                                          \
        for (uword i = 0; i < x.n_elem; ++i) c[i] = REFSHIM_APPLY(s, a[i], OP);                                           \
        return o;                                                                                                        \
    }                                                                                                                    \
    template <int A, int B> Dense<SameKind<TransKind<A>::value, TransKind<B>::value>::value> operator OP(const Tr<A>& x, const Tr<B>& y) { \
        return x.eval() OP y.eval();                                                                                     \
    }                                                                                                                    \
    template <int A, int B> Dense<SameKind<A, TransKind<B>::value>::value> operator OP(const Dense<A>& x, const Tr<B>& y) { \
        return x OP y.eval();                                                                                            \
    }                                                                                                                    \
    template <int A, int B> Dense<SameKind<TransKind<A>::value, B>::value> operator OP(const Tr<A>& x, const Dense<B>& y) { \
        return x.eval() OP y;                                                                                            \
    }

#define REFSHIM_APPLY(a, b, OP) ((a)OP(b))
REFSHIM_EWISE(+, "addition")
REFSHIM_EWISE(-, "subtraction")
REFSHIM_EWISE(/, "element-wise division")
#undef REFSHIM_APPLY
// Schur product `%`; `scalar * X` / `X * scalar` are the scalar forms of it.
template <int A, int B> Dense<SameKind<A, B>::value> operator%(const Dense<A>& x, const Dense<B>& y) {
    same_size(x.n_rows, x.n_cols, y.n_rows, y.n_cols, "element-wise multiplication");
    Dense<SameKind<A, B>::value> o(x.n_rows, x.n_cols);
    const double *a = x.memptr(), *b = y.memptr();
    double* c = o.memptr();
    for (uword i = 0; i < x.n_elem; ++i) c[i] = a[i] * b[i];
    return o;
}
template <int A> Dense<A> operator*(const Dense<A>& x, double s) {
    Dense<A> o(x.n_rows, x.n_cols);
    const double* a = x.memptr();
    double* c = o.memptr();
    for (uword i = 0; i < x.n_elem; ++i) c[i] = a[i] * s;
    return o;
}
template <int A> Dense<A> operator*(double s, const Dense<A>& x) { return x * s; }
#undef REFSHIM_EWISE

#define REFSHIM_MAP(NAME, EXPR)                                                                 \
    template <int A> Dense<A> NAME(const Dense<A>& x) {                                         \
        Dense<A> o(x.n_rows, x.n_cols);                                                         \
        const double* a = x.memptr();                                                           \
        double* c = o.memptr();                                                                 \
        for (uword i = 0; i < x.n_elem; ++i) { const double v = a[i]; c[i] = (EXPR); }          \
        return o;                                                                               \
    }
REFSHIM_MAP(log, std::log(v))
REFSHIM_MAP(exp, std::exp(v))
REFSHIM_MAP(sqrt, std::sqrt(v))
#undef REFSHIM_MAP

// ---- matrix product ----------------------------------------------------------------------------------------------
inline void gemm_into(double* C, bool ta, bool tb, uword m, uword n, uword k, const double* A, uword lda, const double* B, uword ldb) {
    using namespace refshim;
    if (m == 0 || n == 0) return;
    if (k == 0) { std::fill(C, C + m * n, 0.0); return; }
    if (n == 1 && !tb) {  // matrix * column: dgemv, as Armadillo dispatches it
        scipy_cblas_dgemv(ColMajor, ta ? Trans : NoTrans, ta ? (int)k : (int)m, ta ? (int)m : (int)k, 1.0, A, (int)lda, B, 1, 0.0, C, 1);
        return;
    }
    scipy_cblas_dgemm(ColMajor, ta ? Trans : NoTrans, tb ? Trans : NoTrans, (int)m, (int)n, (int)k, 1.0, A, (int)lda, B, (int)ldb, 0.0, C, (int)m);
}
inline void prod_check(uword ac, uword br) {
    if (ac != br) throw std::logic_error("matrix multiplication: incompatible matrix dimensions");
}
template <int A, int B> Dense<ProdKind<A, B>::value> operator*(const Dense<A>& x, const Dense<B>& y) {
    prod_check(x.n_cols, y.n_rows);
    Dense<ProdKind<A, B>::value> o(x.n_rows, y.n_cols);
    gemm_into(o.memptr(), false, false, x.n_rows, y.n_cols, x.n_cols, x.memptr(), x.n_rows, y.memptr(), y.n_rows);
    return o;
}
template <int A, int B> Dense<ProdKind<TransKind<A>::value, B>::value> operator*(const Tr<A>& x, const Dense<B>& y) {
    prod_check(x.src.n_rows, y.n_rows);
    Dense<ProdKind<TransKind<A>::value, B>::value> o(x.src.n_cols, y.n_cols);
    gemm_into(o.memptr(), true, false, x.src.n_cols, y.n_cols, x.src.n_rows, x.src.memptr(), x.src.n_rows, y.memptr(), y.n_rows);
    return o;
}
template <int A, int B> Dense<ProdKind<A, TransKind<B>::value>::value> operator*(const Dense<A>& x, const Tr<B>& y) {
    prod_check(x.n_cols, y.src.n_cols);
    Dense<ProdKind<A, TransKind<B>::value>::value> o(x.n_rows, y.src.n_rows);
    gemm_into(o.memptr(), false, true, x.n_rows, y.src.n_rows, x.n_cols, x.memptr(), x.n_rows, y.src.memptr(), y.src.n_rows);
    return o;
}
template <int A, int B> Dense<ProdKind<TransKind<A>::value, TransKind<B>::value>::value> operator*(const Tr<A>& x, const Tr<B>& y) {
    prod_check(x.src.n_rows, y.src.n_cols);
    Dense<ProdKind<TransKind<A>::value, TransKind<B>::value>::value> o(x.src.n_cols, y.src.n_rows);
    gemm_into(o.memptr(), true, true, x.src.n_cols, y.src.n_rows, x.src.n_rows, x.src.memptr(), x.src.n_rows, y.src.memptr(), y.src.n_rows);
    return o;
}

// ---- generators ----------------------------------------------------------------------------------------------------
inline vec ones(uword n) { vec o(n); o.fill(1.0); return o; }
inline mat ones(uword r, uword c) { mat o(r, c); o.fill(1.0); return o; }
inline vec zeros(uword n) { vec o(n); o.fill(0.0); return o; }
inline mat zeros(uword r, uword c) { mat o(r, c); o.fill(0.0); return o; }
inline mat randn(uword r, uword c) {  // column-major fill, like Armadillo's and like reshape(rnorm(r*c), r, c)
    mat o(r, c);
    for (uword i = 0; i < o.n_elem; ++i) o.mem[i] = refshim::next_normal();
    return o;
}
inline vec randn(uword n) { vec o(n); for (uword i = 0; i < n; ++i) o.mem[i] = refshim::next_normal(); return o; }
inline vec randu(uword n) { vec o(n); for (uword i = 0; i < n; ++i) o.mem[i] = refshim::next_uniform(); return o; }
inline mat randu(uword r, uword c) { mat o(r, c); for (uword i = 0; i < o.n_elem; ++i) o.mem[i] = refshim::next_uniform(); return o; }

// ---- shape functions -----------------------------------------------------------------------------------------------
template <int A> mat repmat(const Dense<A>& x, uword rr, uword cc) {
    mat o(x.n_rows * rr, x.n_cols * cc);
    const uword R = x.n_rows, C = x.n_cols, OR = o.n_rows;
    for (uword cj = 0; cj < cc; ++cj)
        for (uword j = 0; j < C; ++j) {
            double* dst = o.memptr() + (cj * C + j) * OR;
            const double* src = x.memptr() + j * R;
            for (uword ri = 0; ri < rr; ++ri) std::memcpy(dst + ri * R, src, R * sizeof(double));
        }
    return o;
}
template <int A> mat repmat(const Tr<A>& x, uword rr, uword cc) { return repmat(x.eval(), rr, cc); }
template <int A> mat reshape(const Dense<A>& x, uword r, uword c) {
    mat o(r, c);
    const uword m = std::min<uword>(o.n_elem, x.n_elem);
    if (m) std::memcpy(o.memptr(), x.memptr(), m * sizeof(double));
    for (uword i = m; i < o.n_elem; ++i) o.mem[i] = 0.0;
    return o;
}
template <int A> mat diagmat(const Dense<A>& x) {
    if (x.n_rows == 1 || x.n_cols == 1) {
        mat o = zeros(x.n_elem, x.n_elem);
        for (uword i = 0; i < x.n_elem; ++i) o.mem[i + i * x.n_elem] = x.mem[i];
        return o;
    }
    mat o = zeros(x.n_rows, x.n_cols);
    for (uword i = 0; i < std::min(x.n_rows, x.n_cols); ++i) o.mem[i + i * x.n_rows] = x.mem[i + i * x.n_rows];
    return o;
}

// ---- reductions ------------------------------------------------------------------------------------------------------
inline double sum(double x) { return x; }
inline double sum(const vec& x) { double s = 0; for (uword i = 0; i < x.n_elem; ++i) s += x.mem[i]; return s; }
inline double sum(const rowvec& x) { double s = 0; for (uword i = 0; i < x.n_elem; ++i) s += x.mem[i]; return s; }
inline mat sum(const mat& x, uword dim) {
    if (dim == 0) {
        mat o(1, x.n_cols);
        for (uword j = 0; j < x.n_cols; ++j) {
            double s = 0;
            const double* a = x.memptr() + j * x.n_rows;
            for (uword i = 0; i < x.n_rows; ++i) s += a[i];
            o.mem[j] = s;
        }
        return o;
    }
    mat o = zeros(x.n_rows, 1);  // dim 1: row sums, accumulated column by column (Armadillo's order)
    for (uword j = 0; j < x.n_cols; ++j) {
        const double* a = x.memptr() + j * x.n_rows;
        for (uword i = 0; i < x.n_rows; ++i) o.mem[i] += a[i];
    }
    return o;
}
inline mat sum(const mat& x) { return sum(x, 0); }
inline double max(const vec& x) { double m = x(0); for (uword i = 1; i < x.n_elem; ++i) m = std::max(m, x.mem[i]); return m; }
inline double max(const rowvec& x) { double m = x(0); for (uword i = 1; i < x.n_elem; ++i) m = std::max(m, x.mem[i]); return m; }
template <int A> Dense<A> cumsum(const Dense<A>& x, uword dim) {
    Dense<A> o(x.n_rows, x.n_cols);
    if (dim == 0) {
        for (uword j = 0; j < x.n_cols; ++j) {
            double s = 0;
            for (uword i = 0; i < x.n_rows; ++i) { s += x.mem[i + j * x.n_rows]; o.mem[i + j * x.n_rows] = s; }
        }
    } else {
        for (uword i = 0; i < x.n_rows; ++i) {
            double s = 0;
            for (uword j = 0; j < x.n_cols; ++j) { s += x.mem[i + j * x.n_rows]; o.mem[i + j * x.n_rows] = s; }
        }
    }
    return o;
}

// ---- relational / find ---------------------------------------------------------------------------------------------
struct umat {
    uword n_rows = 0, n_cols = 0, n_elem = 0;
    std::vector<uword> mem;
};
struct uvec {
    uword n_rows = 0, n_cols = 1, n_elem = 0;
    std::vector<uword> mem;
    uword operator()(uword i) const { return mem.at(i); }
};
template <int A, int B> umat operator>(const Dense<A>& x, const Dense<B>& y) {
    same_size(x.n_rows, x.n_cols, y.n_rows, y.n_cols, "operator>");
    umat o;
    o.n_rows = x.n_rows; o.n_cols = x.n_cols; o.n_elem = x.n_elem;
    o.mem.resize(x.n_elem);
    for (uword i = 0; i < x.n_elem; ++i) o.mem[i] = x.mem[i] > y.mem[i];
    return o;
}
inline uvec find(const umat& x) {
    uvec o;
    for (uword i = 0; i < x.n_elem; ++i) if (x.mem[i]) o.mem.push_back(i);
    o.n_elem = o.n_rows = o.mem.size();
    return o;
}

// ---- decompositions / solvers --------------------------------------------------------------------------------------
inline mat chol(const mat& Q) {  // upper R with R'R = Q (Armadillo's default layout), LAPACK dpotrf
    if (Q.n_rows != Q.n_cols) throw std::logic_error("chol(): given matrix must be square sized");
    mat R(Q);
    const int n = (int)Q.n_rows;
    if (n && scipy_LAPACKE_dpotrf(refshim::ColMajor, 'U', n, R.memptr(), n) != 0) throw std::runtime_error("chol(): decomposition failed");
    for (int j = 0; j < n; ++j)
        for (int i = j + 1; i < n; ++i) R.mem[i + (size_t)j * n] = 0.0;
    return R;
}
// 0 = general, 1 = upper triangular, 2 = lower triangular (Armadillo's solve() runs this detection on square systems)
inline int triangular_kind(const mat& A) {
    const uword n = A.n_rows;
    bool up = true, lo = true;
    for (uword j = 0; j < n && (up || lo); ++j) {
        const double* a = A.memptr() + j * n;
        for (uword i = 0; i < j && lo; ++i) if (a[i] != 0.0) lo = false;
        for (uword i = j + 1; i < n && up; ++i) if (a[i] != 0.0) up = false;
    }
    return up ? 1 : (lo ? 2 : 0);
}
template <int B> Dense<B> solve_impl(const mat& A, bool transposed, const Dense<B>& rhs) {
    if (A.n_rows != A.n_cols) throw std::logic_error("solve(): only square systems are supported by this stand-in");
    if (A.n_rows != rhs.n_rows) throw std::logic_error("solve(): number of rows in the given matrices must be the same");
    const int n = (int)A.n_rows, nrhs = (int)rhs.n_cols;
    Dense<B> X(rhs);
    if (n == 0 || nrhs == 0) return X;
    const int tk = triangular_kind(A);
    int info;
    if (tk) {
        info = scipy_LAPACKE_dtrtrs(refshim::ColMajor, tk == 1 ? 'U' : 'L', transposed ? 'T' : 'N', 'N', n, nrhs, A.memptr(), n, X.memptr(), n);
    } else {
        mat LU = transposed ? mat(Tr<0>(A)) : mat(A);
        std::vector<int> piv(n);
        info = scipy_LAPACKE_dgesv(refshim::ColMajor, n, nrhs, LU.memptr(), n, piv.data(), X.memptr(), n);
    }
    if (info != 0) throw std::runtime_error("solve(): solution not found");
    return X;
}
template <int B> Dense<B> solve(const mat& A, const Dense<B>& rhs) { return solve_impl(A, false, rhs); }
template <int B> Dense<B> solve(const Tr<0>& A, const Dense<B>& rhs) { return solve_impl(A.src, true, rhs); }
template <int B> Dense<TransKind<B>::value> solve(const mat& A, const Tr<B>& rhs) { return solve_impl(A, false, rhs.eval()); }
template <int B> Dense<TransKind<B>::value> solve(const Tr<0>& A, const Tr<B>& rhs) { return solve_impl(A.src, true, rhs.eval()); }

inline mat inv(const mat& A) {
    if (A.n_rows != A.n_cols) throw std::logic_error("inv(): given matrix must be square sized");
    mat X(A);
    const int n = (int)A.n_rows;
    if (!n) return X;
    std::vector<int> piv(n);
    if (scipy_LAPACKE_dgetrf(refshim::ColMajor, n, n, X.memptr(), n, piv.data()) != 0 ||
        scipy_LAPACKE_dgetri(refshim::ColMajor, n, X.memptr(), n, piv.data()) != 0)
        throw std::runtime_error("inv(): matrix is singular");
    return X;
}
// full SVD X = U diag(d) V' by divide and conquer (Armadillo's default method "dc" = LAPACK dgesdd, jobz 'A')
inline bool svd(mat& U, vec& d, mat& V, const mat& X) {
    const int m = (int)X.n_rows, n = (int)X.n_cols;
    mat A(X), Vt(n, n);
    U.init(m, m);
    d.init(std::min(m, n), 1);
    if (m == 0 || n == 0) { V.init(n, n); return true; }
    if (scipy_LAPACKE_dgesdd(refshim::ColMajor, 'A', m, n, A.memptr(), m, d.memptr(), U.memptr(), m, Vt.memptr(), n) != 0)
        throw std::runtime_error("svd(): decomposition failed");
    V = Tr<0>(Vt).eval();
    return true;
}

}  // namespace arma

// ---- Rcpp: just enough of List / as<> / Environment / Function -----------------------------------------------------
namespace Rcpp {

struct Object {  // what an R numeric vector / matrix SEXP carries
    arma::mat value;
    bool is_vector = false;
};
template <class T> T as(const Object& o);
template <> inline arma::mat as<arma::mat>(const Object& o) { return o.value; }
template <> inline arma::vec as<arma::vec>(const Object& o) {
    if (o.value.n_cols != 1 && o.value.n_rows != 1) throw std::logic_error("as<vec>(): object is not a vector");
    arma::vec v(o.value.n_elem);
    if (v.n_elem) std::memcpy(v.memptr(), o.value.memptr(), v.n_elem * sizeof(double));
    return v;
}

struct NamedObject {
    std::string name;
    Object obj;
};
struct Argument {
    std::string name;
    template <int K> NamedObject operator=(const arma::Dense<K>& x) const { return NamedObject{name, Object{arma::mat(x), K != 0}}; }
};
struct NamedPlaceHolder {
    Argument operator[](const std::string& n) const { return Argument{n}; }
};
static const NamedPlaceHolder _;

class List {
  public:
    std::vector<NamedObject> items;
    template <class... Args> static List create(const Args&... a) {
        List l;
        (l.items.push_back(a), ...);
        return l;
    }
    const Object& operator[](const std::string& n) const {
        for (const auto& it : items) if (it.name == n) return it.obj;
        throw std::out_of_range("Index out of bounds: [index='" + n + "'].");
    }
    void set(const std::string& n, const arma::mat& m, bool is_vector) { items.push_back(NamedObject{n, Object{m, is_vector}}); }
};

class Function {
  public:
    std::string name;
    explicit Function(std::string n = "") : name(std::move(n)) {}
    Object operator()(double count) const {  // stats::rnorm(n): the next n N(0,1) of R's stream, as a numeric vector
        if (name != "rnorm") throw std::runtime_error("refshim: R function '" + name + "' is not available");
        Object o;
        o.value = arma::mat(arma::randn((arma::uword)count));
        o.is_vector = true;
        return o;
    }
};
class Environment {
  public:
    std::string name;
    Environment(const char* n) : name(n) {}
    Function operator[](const std::string& f) const {
        if (name != "package:stats") throw std::runtime_error("refshim: environment '" + name + "' is not available");
        return Function(f);
    }
};

}  // namespace Rcpp
