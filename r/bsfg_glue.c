/* R <-> libbsfg_b200 glue: the .Call shims a maintainer of deruncie/SparseFactorMixedModel adds so that
 * `model_setup.R:15`  sourceCpp('BSFG_functions_c.cpp')  can be replaced by  source('bsfg_b200.R').
 *
 * Build where R exists (this repository's build container has no R, so this file is shipped as source and is
 * exercised through the identical C ABI from Python/ctypes in tests/):
 *     R CMD SHLIB bsfg_glue.c -I<repo>/include -L<repo>/sparsefactormixedmodel_b200 -lbsfg_b200 -o bsfg_glue.so
 *
 * Every shim mirrors one Rcpp export of R_BSFG/BSFG_functions_c.cpp: same argument order, R numeric
 * matrices (column-major doubles) passed straight through as host pointers, `List` arguments flattened to
 * their members, result returned as a fresh R matrix; `vec` results come back as k x 1 / p x 1 matrices like
 * RcppArmadillo's wrap().  A trailing `variates` argument (R NULL or a numeric matrix filled by rnorm/runif in
 * the reference's layout) selects injected variates (the testing hook of BSFG_functions_c.cpp:4,69-72) or the
 * device Philox stream keyed by `seed`.  Errors become R conditions, like Rcpp's BEGIN_RCPP/END_RCPP.
 */
#include <R.h>
#include <Rinternals.h>
#include "bsfg_b200.h"

static void chk(int rc) { if (rc != BSFG_OK) Rf_error("libbsfg_b200 (%d): %s", rc, bsfg_last_error()); }
static const double* opt(SEXP x) { return Rf_isNull(x) ? NULL : REAL(x); }
static SEXP member(SEXP list, const char* name) {
    SEXP names = Rf_getAttrib(list, R_NamesSymbol);
    for (R_xlen_t i = 0; i < XLENGTH(list); i++)
        if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
    Rf_error("list member '%s' missing", name);
    return R_NilValue;
}
static unsigned long long seed_of(SEXP s) { return (unsigned long long)Rf_asReal(s); }

/* sample_Lambda_c(Y_tilde,F,resid_Y_prec,E_a_prec,Plam,invert_aI_bZAZ)        BSFG_functions_c.cpp:86-91 */
SEXP R_bsfg_sample_Lambda_c(SEXP Y, SEXP F, SEXP rp, SEXP ep, SEXP Plam, SEXP inv, SEXP z, SEXP seed) {
    int n = Rf_nrows(Y), p = Rf_ncols(Y), k = Rf_ncols(F);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, p, k));
    chk(bsfg_sample_Lambda_c(REAL(Y), n, p, REAL(F), k, REAL(rp), REAL(ep), REAL(Plam), REAL(member(inv, "U")),
                             REAL(member(inv, "s")), opt(z), seed_of(seed), REAL(out)));
    UNPROTECT(1);
    return out;
}
/* sample_means_c(Y_tilde,resid_Y_prec,E_a_prec,invert_aPXA_bDesignDesignT)                     cpp:43-46 */
SEXP R_bsfg_sample_means_c(SEXP Y, SEXP rp, SEXP ep, SEXP inv, SEXP z, SEXP seed) {
    int n = Rf_nrows(Y), p = Rf_ncols(Y);
    SEXP DU = member(inv, "Design_U");
    int br = Rf_ncols(DU);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, br, p));
    chk(bsfg_sample_means_c(REAL(Y), n, p, REAL(rp), REAL(ep), REAL(member(inv, "U")), REAL(member(inv, "s1")),
                            REAL(member(inv, "s2")), REAL(DU), br, opt(z), seed_of(seed), REAL(out)));
    UNPROTECT(1);
    return out;
}
/* sample_factors_scores_c(Y_tilde,Z_1,Lambda,resid_Y_prec,F_a,F_h2)                           cpp:138-143 */
SEXP R_bsfg_sample_factors_scores_c(SEXP Y, SEXP Z1, SEXP Lam, SEXP rp, SEXP Fa, SEXP h2, SEXP z, SEXP seed) {
    int n = Rf_nrows(Y), p = Rf_ncols(Y), r = Rf_ncols(Z1), k = Rf_ncols(Lam);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, k));
    chk(bsfg_sample_factors_scores_c(REAL(Y), n, p, REAL(Z1), r, REAL(Lam), k, REAL(rp), REAL(Fa), REAL(h2), opt(z),
                                     seed_of(seed), REAL(out)));
    UNPROTECT(1);
    return out;
}
/* sample_px_factors_scores_c(Y_tilde,Z_1,Lambda_px,resid_Y_prec,F_a_px,tau_e,px_factor)       cpp:166-172 */
SEXP R_bsfg_sample_px_factors_scores_c(SEXP Y, SEXP Z1, SEXP Lam, SEXP rp, SEXP Fa, SEXP tau_e, SEXP px, SEXP z, SEXP seed) {
    int n = Rf_nrows(Y), p = Rf_ncols(Y), r = Rf_ncols(Z1), k = Rf_ncols(Lam);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, k));
    chk(bsfg_sample_px_factors_scores_c(REAL(Y), n, p, REAL(Z1), r, REAL(Lam), k, REAL(rp), REAL(Fa), REAL(tau_e),
                                        REAL(px), opt(z), seed_of(seed), REAL(out)));
    UNPROTECT(1);
    return out;
}
/* sample_h2s_discrete_c(F,h2_divisions,h2_priors,invert_aI_bZAZ)                              cpp:196-199 */
SEXP R_bsfg_sample_h2s_discrete_c(SEXP F, SEXP H, SEXP pri, SEXP inv, SEXP u, SEXP seed) {
    int n = Rf_nrows(F), k = Rf_ncols(F);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, k, 1));
    chk(bsfg_sample_h2s_discrete_c(REAL(F), n, k, Rf_asInteger(H), REAL(pri), REAL(member(inv, "U")),
                                   REAL(member(inv, "s")), opt(u), seed_of(seed), REAL(out), NULL));
    UNPROTECT(1);
    return out;
}
/* sample_prec_discrete_conditional_c(Y,h2_divisions,h2_priors,invert_aI_bZAZ,res_prec)        cpp:241-245 */
SEXP R_bsfg_sample_prec_discrete_conditional_c(SEXP Y, SEXP H, SEXP pri, SEXP inv, SEXP rp, SEXP u, SEXP seed) {
    int n = Rf_nrows(Y), p = Rf_ncols(Y);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, p, 1));
    chk(bsfg_sample_prec_discrete_conditional_c(REAL(Y), n, p, Rf_asInteger(H), REAL(pri), REAL(member(inv, "U")),
                                                REAL(member(inv, "s")), REAL(rp), opt(u), seed_of(seed), REAL(out), NULL));
    UNPROTECT(1);
    return out;
}
/* sample_F_a_c(F,Z_1,F_h2,invert_aZZt_Ainv)                                                   cpp:293-296 */
SEXP R_bsfg_sample_F_a_c(SEXP F, SEXP Z1, SEXP h2, SEXP inv, SEXP z, SEXP seed) {
    int n = Rf_nrows(F), k = Rf_ncols(F), r = Rf_ncols(Z1);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, r, k));
    chk(bsfg_sample_F_a_c(REAL(F), n, k, REAL(Z1), r, REAL(h2), REAL(member(inv, "U")), REAL(member(inv, "s1")),
                          REAL(member(inv, "s2")), opt(z), seed_of(seed), REAL(out)));
    UNPROTECT(1);
    return out;
}
/* sample_delta(delta,tauh,Lambda_prec,d1s,d1r,d2s,d2r,Lambda2)                       BSFG_functions.R:272 */
SEXP R_bsfg_sample_delta_c(SEXP delta, SEXP tauh, SEXP LP, SEXP d1s, SEXP d1r, SEXP d2s, SEXP d2r, SEXP L2, SEXP g, SEXP seed) {
    int p = Rf_nrows(LP), k = Rf_ncols(LP);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, k));
    chk(bsfg_sample_delta_c(REAL(delta), REAL(tauh), k, REAL(LP), p, Rf_asReal(d1s), Rf_asReal(d1r), Rf_asReal(d2s),
                            Rf_asReal(d2r), REAL(L2), opt(g), seed_of(seed), REAL(out), NULL, NULL, NULL));
    UNPROTECT(1);
    return out;
}

/* GSVD_2_c(A,B)                                                                      cpp:27-40 */
SEXP R_bsfg_GSVD_2_c(SEXP A, SEXP B) {
    int n = Rf_ncols(B);
    const char* names[] = {"U", "V", "X", "C", "S", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SEXP U = PROTECT(Rf_allocMatrix(REALSXP, n, n)), V = PROTECT(Rf_allocMatrix(REALSXP, n, n)), X = PROTECT(Rf_allocMatrix(REALSXP, n, n));
    SEXP Cm = PROTECT(Rf_allocMatrix(REALSXP, n, n)), Sm = PROTECT(Rf_allocMatrix(REALSXP, n, n));
    double* c = (double*)R_alloc(n, sizeof(double)); double* sd = (double*)R_alloc(n, sizeof(double));
    chk(bsfg_GSVD_2_c(REAL(A), REAL(B), n, REAL(U), REAL(V), REAL(X), c, sd));
    memset(REAL(Cm), 0, sizeof(double) * (size_t)n * n); memset(REAL(Sm), 0, sizeof(double) * (size_t)n * n);
    for (int i = 0; i < n; i++) { REAL(Cm)[(size_t)i * n + i] = c[i]; REAL(Sm)[(size_t)i * n + i] = sd[i]; }   /* diagmat, cpp:35-36 */
    SET_VECTOR_ELT(out, 0, U); SET_VECTOR_ELT(out, 1, V); SET_VECTOR_ELT(out, 2, X); SET_VECTOR_ELT(out, 3, Cm); SET_VECTOR_ELT(out, 4, Sm);
    UNPROTECT(6);
    return out;
}
/* update_k: the R closure pads the factor-indexed members to k_max = k + 1 columns (fresh copies), this call edits them in
 * place and returns the new k; pri as in R_bsfg_create.                                      BSFG_functions.R:311-372 */
SEXP R_bsfg_update_k_c(SEXP dims, SEXP pri, SEXP Z1, SEXP Lambda, SEXP LP, SEXP Plam, SEXP delta, SEXP tauh, SEXP F_h2, SEXP F,
                       SEXP F_a, SEXP px_factor, SEXP F_px, SEXP seed) {
    int* d = INTEGER(dims);            /* c(n,p,r,k,k_max,nrun) */
    double* q = REAL(pri);
    bsfg_priors pr = {q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], q[9], NULL, q[10], q[11], q[12], q[13],
                      q[14], q[15], q[16], q[17]};
    int k_out = 0;
    chk(bsfg_update_k_c(d[0], d[1], d[2], d[3], d[4], (long long)d[5], &pr, opt(Z1), REAL(Lambda), REAL(LP), REAL(Plam), REAL(delta),
                        REAL(tauh), REAL(F_h2), REAL(F), REAL(F_a), opt(px_factor), opt(F_px), NULL, seed_of(seed), &k_out));
    return Rf_ScalarInteger(k_out);
}

/* ---- device-resident sweep: fast_BSFG_sampler.R:176-270 in one call --------------------------------------- */
static void ctx_finalizer(SEXP ptr) {
    bsfg_ctx* c = (bsfg_ctx*)R_ExternalPtrAddr(ptr);
    if (c) { bsfg_destroy(c); R_ClearExternalPtr(ptr); }
}
/* dims = c(n,p,k,k_max,r,b,h2_divisions); pri = c(resid shape,rate, E_a shape,rate, Lambda_df, delta_1 shape,rate,
 * delta_2 shape,rate, b_X_prec, b0,b1,epsilon,prop); lists as in BSFG_state$run_variables */
SEXP R_bsfg_create(SEXP dims, SEXP pri, SEXP h2_priors, SEXP Y, SEXP X, SEXP Z1, SEXP inv1, SEXP inv2, SEXP inv3,
                   SEXP Ainv, SEXP z1_identity, SEXP seed, SEXP Z2, SEXP inv4) {
    int* d = INTEGER(dims);            /* c(n,p,k,k_max,r,b,h2_divisions,r2) */
    double* q = REAL(pri);             /* ..., prop, W_prec_shape, W_prec_rate */
    bsfg_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.n = d[0]; cfg.p_total = d[1]; cfg.p_local = d[1]; cfg.k = d[2]; cfg.k_max = d[3]; cfg.r = d[4]; cfg.b = d[5];
    cfg.h2_divisions = d[6]; cfg.r2 = d[7]; cfg.world = 1; cfg.seed = seed_of(seed); cfg.z1_is_identity = Rf_asLogical(z1_identity);
    bsfg_priors pr = {q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], q[9], REAL(h2_priors), q[10], q[11], q[12], q[13],
                      q[14], q[15]};
    bsfg_ctx* c = NULL;
    chk(bsfg_create(&cfg, &c));
    SEXP ptr = PROTECT(R_MakeExternalPtr(c, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
    chk(bsfg_set_priors(c, &pr));
    bsfg_constants k = {REAL(Y), d[5] > 0 ? REAL(X) : NULL, Rf_asLogical(z1_identity) ? NULL : REAL(Z1),
                        REAL(member(inv1, "U")), REAL(member(inv1, "s")),
                        REAL(member(inv2, "U")), REAL(member(inv2, "s1")), REAL(member(inv2, "s2")), REAL(member(inv2, "Design_U")),
                        REAL(member(inv3, "U")), REAL(member(inv3, "s1")), REAL(member(inv3, "s2")), opt(Ainv),
                        d[7] > 0 ? REAL(Z2) : NULL, d[7] > 0 ? REAL(member(inv4, "U")) : NULL,
                        d[7] > 0 ? REAL(member(inv4, "s1")) : NULL, d[7] > 0 ? REAL(member(inv4, "s2")) : NULL,
                        NULL /* A_2_inv = diag(1, r2), fast_BSFG_sampler_init.R:264 */};
    chk(bsfg_upload_constants(c, &k));
    UNPROTECT(1);
    return ptr;
}
static void fill_state(bsfg_state* st, SEXP cs, int k, int nrun) {
    st->Lambda = REAL(member(cs, "Lambda")); st->Lambda_prec = REAL(member(cs, "Lambda_prec")); st->Plam = REAL(member(cs, "Plam"));
    st->F = REAL(member(cs, "F")); st->F_a = REAL(member(cs, "F_a")); st->F_h2 = REAL(member(cs, "F_h2"));
    st->B = REAL(member(cs, "B")); st->E_a = REAL(member(cs, "E_a"));
    st->resid_Y_prec = REAL(member(cs, "resid_Y_prec")); st->E_a_prec = REAL(member(cs, "E_a_prec"));
    st->delta = REAL(member(cs, "delta")); st->tauh = REAL(member(cs, "tauh"));
    st->W = REAL(member(cs, "W")); st->W_prec = REAL(member(cs, "W_prec"));      /* ignored by the library when r2 == 0 */
    st->k = k; st->nrun = nrun;
}
SEXP R_bsfg_set_state(SEXP ptr, SEXP cs) {
    bsfg_state st;
    memset(&st, 0, sizeof st);      /* px_factor / F_px stay NULL: bsfg_set_state reads them as "not a px chain" */
    fill_state(&st, cs, Rf_ncols(member(cs, "F")), Rf_asInteger(member(cs, "nrun")));
    chk(bsfg_set_state((bsfg_ctx*)R_ExternalPtrAddr(ptr), &st));
    return R_NilValue;
}
SEXP R_bsfg_sweep(SEXP ptr, SEXP n_iter) {
    chk(bsfg_sweep((bsfg_ctx*)R_ExternalPtrAddr(ptr), Rf_asInteger(n_iter)));
    return R_NilValue;
}
/* returns c(k, nrun); the caller allocates `cs` with k columns (two calls: first with cs = NULL to learn k) */
SEXP R_bsfg_get_state(SEXP ptr, SEXP cs) {
    bsfg_state st;
    memset(&st, 0, sizeof st);
    if (!Rf_isNull(cs)) fill_state(&st, cs, 0, 0);
    chk(bsfg_get_state((bsfg_ctx*)R_ExternalPtrAddr(ptr), &st));
    SEXP out = PROTECT(Rf_allocVector(INTSXP, 2));
    INTEGER(out)[0] = st.k; INTEGER(out)[1] = st.nrun;
    UNPROTECT(1);
    return out;
}
