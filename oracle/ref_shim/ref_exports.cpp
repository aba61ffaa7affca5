// ref_exports.cpp -- TEST INFRASTRUCTURE ONLY.  Flat C entry points around the reference's own Rcpp exports.
//
// The functions called here are DEFINED in `/root/reference/R_BSFG/BSFG_functions_c.cpp`, which the Makefile next to
// this file compiles unmodified (as its own translation unit) against the stand-in `RcppArmadillo.h`.  This file only
// declares their prototypes (copied signatures: cpp:27, 43-46, 86-91, 138-143, 166-172, 196-199, 241-245, 293-296),
// converts flat column-major buffers to arma objects and installs the injected variate streams.  It plays the role of
// the `sourceCpp`-generated `.Call` wrappers (BEGIN_RCPP / END_RCPP: exceptions become an error code + message).
#include <RcppArmadillo.h>

using arma::mat;
using arma::vec;
using Rcpp::List;

List GSVD_2_c(mat A, mat B);
mat sample_means_c(mat Y_tilde, vec resid_Y_prec, vec E_a_prec, List invert_aPXA_bDesignDesignT);
mat sample_Lambda_c(mat Y_tilde, mat F, vec resid_Y_prec, vec E_a_prec, mat Plam, List invert_aI_bZAZ);
mat sample_factors_scores_c(mat Y_tilde, mat Z_1, mat Lambda, vec resid_Y_prec, mat F_a, vec F_h2);
mat sample_px_factors_scores_c(mat Y_tilde, mat Z_1, mat Lambda_px, vec resid_Y_prec, mat F_a_px, vec tau_e, vec px_factor);
vec sample_h2s_discrete_c(mat F, int h2_divisions, vec h2_priors, List invert_aI_bZAZ);
vec sample_prec_discrete_conditional_c(mat Y, int h2_divisions, vec h2_priors, List invert_aI_bZAZ, vec res_prec);
mat sample_F_a_c(mat F, mat Z_1, vec F_h2, List invert_aZZt_Ainv);

extern "C" {
void scipy_openblas_set_num_threads(int);
int scipy_openblas_get_num_threads(void);
}

namespace {

thread_local std::string g_err;

mat M(const double* p, long r, long c) {
    mat m((arma::uword)r, (arma::uword)c);
    if (m.n_elem) std::memcpy(m.memptr(), p, m.n_elem * sizeof(double));
    return m;
}
vec V(const double* p, long n) {
    vec v((arma::uword)n);
    if (n) std::memcpy(v.memptr(), p, n * sizeof(double));
    return v;
}
template <int K> void out(const arma::Dense<K>& x, double* dst, long expect_elems) {
    if ((long)x.n_elem != expect_elems) throw std::logic_error("ref_exports: result has " + std::to_string(x.n_elem) + " elements, caller expected " + std::to_string(expect_elems));
    if (x.n_elem) std::memcpy(dst, x.memptr(), x.n_elem * sizeof(double));
}
struct Streams {  // install the injected variates for the duration of one call; leftover variates are an error
    Streams(const double* z, long nz, const double* u, long nu) {
        refshim::normals = refshim::Stream{z, z ? (size_t)nz : 0, 0};
        refshim::uniforms = refshim::Stream{u, u ? (size_t)nu : 0, 0};
    }
    void finish() {
        if (refshim::normals.p && refshim::normals.pos != refshim::normals.n) throw std::logic_error("ref_exports: injected N(0,1) stream not fully consumed");
        if (refshim::uniforms.p && refshim::uniforms.pos != refshim::uniforms.n) throw std::logic_error("ref_exports: injected U(0,1) stream not fully consumed");
    }
    ~Streams() { refshim::normals = refshim::Stream(); refshim::uniforms = refshim::Stream(); }
};
template <class Fn> int guarded(Fn&& fn) {
    try {
        fn();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // namespace

extern "C" {

const char* ref_last_error(void) { return g_err.c_str(); }
int ref_blas_threads(int set) {
    if (set > 0) scipy_openblas_set_num_threads(set);
    return scipy_openblas_get_num_threads();
}
void ref_seed_fallback_rng(unsigned long long seed) {
    refshim::fallback_state[0] = 0x9E3779B97F4A7C15ull ^ seed;
    refshim::fallback_state[1] = 0xD1B54A32D192ED03ull + 0x632BE59BD9B4E019ull * (seed | 1);
}

// GSVD_2_c(A, B), cpp:27-40.  A, B: n x n.  Outputs U, Vout, X, C, S: n x n each.
int ref_GSVD_2_c(const double* A, const double* B, long n, double* U, double* Vout, double* X, double* C, double* S) {
    return guarded([&] {
        List l = GSVD_2_c(M(A, n, n), M(B, n, n));
        out(Rcpp::as<mat>(l["U"]), U, n * n);
        out(Rcpp::as<mat>(l["V"]), Vout, n * n);
        out(Rcpp::as<mat>(l["X"]), X, n * n);
        out(Rcpp::as<mat>(l["C"]), C, n * n);
        out(Rcpp::as<mat>(l["S"]), S, n * n);
    });
}

// sample_means_c, cpp:43-82.  z: br*p normals (randn(br,p), cpp:68) or NULL.  out: br x p.
int ref_sample_means_c(const double* Y_tilde, long n, long p, const double* resid_Y_prec, const double* E_a_prec, const double* U,
                       const double* s1, const double* s2, const double* Design_U, long br, const double* z, double* result) {
    return guarded([&] {
        Streams st(z, br * p, nullptr, 0);
        List l;
        l.set("U", M(U, br, br), false);
        l.set("s1", M(s1, br, 1), true);
        l.set("s2", M(s2, br, 1), true);
        l.set("Design_U", M(Design_U, n, br), false);
        mat r = sample_means_c(M(Y_tilde, n, p), V(resid_Y_prec, p), V(E_a_prec, p), l);
        st.finish();
        out(r, result, br * p);
    });
}

// sample_Lambda_c, cpp:86-135.  z: k*p normals (randn(k,p), cpp:110) or NULL.  out: p x k.
int ref_sample_Lambda_c(const double* Y_tilde, long n, long p, const double* F, long k, const double* resid_Y_prec, const double* E_a_prec,
                        const double* Plam, const double* U, const double* s, const double* z, double* result) {
    return guarded([&] {
        Streams st(z, k * p, nullptr, 0);
        List l;
        l.set("U", M(U, n, n), false);
        l.set("s", M(s, n, 1), true);
        mat r = sample_Lambda_c(M(Y_tilde, n, p), M(F, n, k), V(resid_Y_prec, p), V(E_a_prec, p), M(Plam, p, k), l);
        st.finish();
        out(r, result, p * k);
    });
}

// sample_factors_scores_c, cpp:138-163.  z: n*k normals (randn(n,k), cpp:154) or NULL.  out: n x k.
int ref_sample_factors_scores_c(const double* Y_tilde, long n, long p, const double* Z_1, long r, const double* Lambda, long k,
                                const double* resid_Y_prec, const double* F_a, const double* F_h2, const double* z, double* result) {
    return guarded([&] {
        Streams st(z, n * k, nullptr, 0);
        mat o = sample_factors_scores_c(M(Y_tilde, n, p), M(Z_1, n, r), M(Lambda, p, k), V(resid_Y_prec, p), M(F_a, r, k), V(F_h2, k));
        st.finish();
        out(o, result, n * k);
    });
}

// sample_px_factors_scores_c, cpp:166-193.
int ref_sample_px_factors_scores_c(const double* Y_tilde, long n, long p, const double* Z_1, long r, const double* Lambda_px, long k,
                                   const double* resid_Y_prec, const double* F_a_px, const double* tau_e, const double* px_factor,
                                   const double* z, double* result) {
    return guarded([&] {
        Streams st(z, n * k, nullptr, 0);
        mat o = sample_px_factors_scores_c(M(Y_tilde, n, p), M(Z_1, n, r), M(Lambda_px, p, k), V(resid_Y_prec, p), M(F_a_px, r, k), V(tau_e, k), V(px_factor, k));
        st.finish();
        out(o, result, n * k);
    });
}

// sample_h2s_discrete_c, cpp:196-238.  u: k uniforms (one randu(1) per factor, cpp:232) or NULL.  out: k.
int ref_sample_h2s_discrete_c(const double* F, long n, long k, int h2_divisions, const double* h2_priors, const double* U, const double* s,
                              const double* u, double* result) {
    return guarded([&] {
        Streams st(nullptr, 0, u, k);
        List l;
        l.set("U", M(U, n, n), false);
        l.set("s", M(s, n, 1), true);
        vec o = sample_h2s_discrete_c(M(F, n, k), h2_divisions, V(h2_priors, h2_divisions), l);
        st.finish();
        out(o, result, k);
    });
}

// sample_prec_discrete_conditional_c, cpp:241-289.  u: p uniforms (cpp:282) or NULL.  out: p.
int ref_sample_prec_discrete_conditional_c(const double* Y, long n, long p, int h2_divisions, const double* h2_priors, const double* U,
                                           const double* s, const double* res_prec, const double* u, double* result) {
    return guarded([&] {
        Streams st(nullptr, 0, u, p);
        List l;
        l.set("U", M(U, n, n), false);
        l.set("s", M(s, n, 1), true);
        vec o = sample_prec_discrete_conditional_c(M(Y, n, p), h2_divisions, V(h2_priors, h2_divisions), l, V(res_prec, p));
        st.finish();
        out(o, result, p);
    });
}

// sample_F_a_c, cpp:293-339.  z: r*k normals (stats::rnorm(r*k) reshaped r x k, cpp:316-319) or NULL.  out: r x k.
int ref_sample_F_a_c(const double* F, long n, long k, const double* Z_1, long r, const double* F_h2, const double* U, const double* s1,
                     const double* s2, const double* z, double* result) {
    return guarded([&] {
        Streams st(z, r * k, nullptr, 0);
        List l;
        l.set("U", M(U, r, r), false);
        l.set("s1", M(s1, r, 1), true);
        l.set("s2", M(s2, r, 1), true);
        mat o = sample_F_a_c(M(F, n, k), M(Z_1, n, r), V(F_h2, k), l);
        st.finish();
        out(o, result, r * k);
    });
}

}  // extern "C"
