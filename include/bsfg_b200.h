/* libbsfg_b200 -- C ABI of the B200-native BSFG Gibbs sweep.
 *
 * Drop-in boundary for the Rcpp exports of deruncie/SparseFactorMixedModel
 * (R_BSFG/BSFG_functions_c.cpp) and for the R-only sweep steps `sample_delta` / `update_k`
 * (R_BSFG/BSFG_functions.R:272-372) plus the loop body of `fast_BSFG_sampler`
 * (R_BSFG/fast_BSFG_sampler.R:176-270).
 *
 * Conventions (identical to what Rcpp/Armadillo hands the reference functions):
 *   - every matrix is FP64, dense, COLUMN-MAJOR, leading dimension == number of rows;
 *   - every pointer in a `bsfg_<name>_c` signature is a HOST pointer owned by the caller; the
 *     library copies in, computes on the current CUDA device, copies out, and keeps nothing;
 *   - R `List` arguments are flattened to their members (U,s / U,s1,s2,Design_U);
 *   - variates: a non-NULL pointer injects the caller's N(0,1)/U(0,1)/Gamma(shape,1) draws, laid
 *     out exactly as the reference fills them (see each function); NULL draws them on the device
 *     from Philox4x32-10 keyed by `seed`;
 *   - return value: 0 (BSFG_OK) or a negative bsfg_status; `bsfg_last_error()` returns the message
 *     (thread-local).  There is NO CPU fallback: without a CUDA device every compute entry point
 *     returns BSFG_ECUDA.
 */
#ifndef BSFG_B200_H
#define BSFG_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bsfg_status {
    BSFG_OK = 0,
    BSFG_ECUDA = -1,   /* CUDA runtime/driver error (includes "no device")                     */
    BSFG_ENOTPD = -2,  /* chol(): decomposition failed (Armadillo would throw); see last_error */
    BSFG_ESHAPE = -3,  /* size mismatch / unsupported size                                      */
    BSFG_ENCCL = -4,   /* NCCL error or NCCL library not loadable                               */
    BSFG_EARG = -5,    /* NULL where data is required, bad enum, ...                            */
    BSFG_ESTATE = -6   /* call order violated (e.g. sweep before upload)                        */
} bsfg_status;

int bsfg_version(void);                 /* 100*major + minor */
const char* bsfg_last_error(void);
int bsfg_device_count(void);            /* number of visible CUDA devices (0 if none) */
long long bsfg_launch_count(void);      /* CUDA kernels launched by this library in this process so far */
int bsfg_k_limit(void);                /* largest supported number of factors k (and k_max); update_k may grow k towards 2p, BSFG_functions.R:332 */

/* ------------------------------------------------------------------------------------------
 * Stateless drop-ins: one per Rcpp export on the hot path.
 * ------------------------------------------------------------------------------------------ */

/* sample_Lambda_c(Y_tilde,F,resid_Y_prec,E_a_prec,Plam,invert_aI_bZAZ)  BSFG_functions_c.cpp:86-135
 * Y_tilde n x p, F n x k, resid_Y_prec p, E_a_prec p, Plam p x k, U n x n, s n.
 * z: k x p N(0,1), column j = trait j (= `randn(k,p)`, cpp:110), or NULL.  Lambda_out: p x k.     */
int bsfg_sample_Lambda_c(const double* Y_tilde, int n, int p, const double* F, int k,
                         const double* resid_Y_prec, const double* E_a_prec, const double* Plam,
                         const double* U, const double* s, const double* z, unsigned long long seed,
                         double* Lambda_out);
/* spelling used by BASELINE.json's north_star; same function */
int bsfg_sample_lambda_c(const double* Y_tilde, int n, int p, const double* F, int k,
                         const double* resid_Y_prec, const double* E_a_prec, const double* Plam,
                         const double* U, const double* s, const double* z, unsigned long long seed,
                         double* Lambda_out);

/* sample_means_c(Y_tilde,resid_Y_prec,E_a_prec,invert_aPXA_bDesignDesignT)        cpp:43-82
 * U br x br, s1 br, s2 br, Design_U n x br.  z: br x p (= `randn(br,p)`, cpp:68) or NULL.
 * location_sample_out: br x p (rows 0..b-1 = B, rest = E_a; fast_BSFG_sampler.R:206-207).          */
int bsfg_sample_means_c(const double* Y_tilde, int n, int p, const double* resid_Y_prec,
                        const double* E_a_prec, const double* U, const double* s1, const double* s2,
                        const double* Design_U, int br, const double* z, unsigned long long seed,
                        double* location_sample_out);

/* sample_factors_scores_c(Y_tilde,Z_1,Lambda,resid_Y_prec,F_a,F_h2)              cpp:138-163
 * Z_1 n x r, Lambda p x k, F_a r x k, F_h2 k.  z: n x k (= `randn(n,k)`, cpp:154) or NULL.  F_out n x k. */
int bsfg_sample_factors_scores_c(const double* Y_tilde, int n, int p, const double* Z_1, int r,
                                 const double* Lambda, int k, const double* resid_Y_prec,
                                 const double* F_a, const double* F_h2, const double* z,
                                 unsigned long long seed, double* F_out);

/* sample_px_factors_scores_c(Y_tilde,Z_1,Lambda_px,resid_Y_prec,F_a_px,tau_e,px_factor)  cpp:166-193
 * identical to the above with tau_e supplied; px_factor is accepted and unused, like the reference. */
int bsfg_sample_px_factors_scores_c(const double* Y_tilde, int n, int p, const double* Z_1, int r,
                                    const double* Lambda_px, int k, const double* resid_Y_prec,
                                    const double* F_a_px, const double* tau_e, const double* px_factor,
                                    const double* z, unsigned long long seed, double* F_px_out);

/* sample_h2s_discrete_c(F,h2_divisions,h2_priors,invert_aI_bZAZ)                 cpp:196-238
 * u: k U(0,1), one per factor in factor order (cpp:232) or NULL.  F_h2_out: k values in {0,1/H,..}.
 * log_ps_out (optional, may be NULL): k x H table of the unnormalised log posteriors (cpp:226),
 * exported so callers can apply the "two grid log-likelihoods within 1e-9" tie rule.             */
int bsfg_sample_h2s_discrete_c(const double* F, int n, int k, int h2_divisions, const double* h2_priors,
                               const double* U, const double* s, const double* u,
                               unsigned long long seed, double* F_h2_out, double* log_ps_out);

/* sample_prec_discrete_conditional_c(Y,h2_divisions,h2_priors,invert_aI_bZAZ,res_prec)  cpp:241-289
 * u: p U(0,1), one per trait (cpp:282) or NULL.  Prec_out: p.  log_ps_out optional p x H.
 * Keeps the C++ integer division n/2 of cpp:274.                                                  */
int bsfg_sample_prec_discrete_conditional_c(const double* Y, int n, int p, int h2_divisions,
                                            const double* h2_priors, const double* U, const double* s,
                                            const double* res_prec, const double* u,
                                            unsigned long long seed, double* Prec_out, double* log_ps_out);

/* sample_F_a_c(F,Z_1,F_h2,invert_aZZt_Ainv)                                       cpp:293-339
 * U r x r, s1 r, s2 r.  z: r x k (= `reshape(rnorm(r*k),r,k)`, cpp:316-319) or NULL; consumed for all
 * k columns.  Branches of cpp:324-332 kept (tau_e==1 -> 0; tau_e>10000 -> F[,j], needs r==n).     */
int bsfg_sample_F_a_c(const double* F, int n, int k, const double* Z_1, int r, const double* F_h2,
                      const double* U, const double* s1, const double* s2, const double* z,
                      unsigned long long seed, double* F_a_out);

/* sample_delta (R only, BSFG_functions.R:272-308); device version named sample_delta_c.
 * Lambda_prec, Lambda2: p x k.  g: standard Gamma(shape,1) variates in draw order -- g[0] for
 * delta[1], then one per h visited by R's `for(h in 2:(k-1))` -- or NULL.  n_draws_out (optional)
 * receives the number of draws; shapes_out/rates_out (optional, k entries each) the Gamma parameters
 * of every draw, which is what parity is judged on when draws come from the device.               */
int bsfg_sample_delta_c(const double* delta, const double* tauh, int k, const double* Lambda_prec, int p,
                        double delta_1_shape, double delta_1_rate, double delta_2_shape,
                        double delta_2_rate, const double* Lambda2, const double* g,
                        unsigned long long seed, double* delta_out, int* n_draws_out,
                        double* shapes_out, double* rates_out);

/* GSVD_2_c(A,B)                                                                    cpp:27-40
 * A, B n x n (B invertible).  The reference's literal route svd(A inv(B)): returns U, V (n x n, orthogonal), X (n x n) and
 * the DIAGONALS of C and S (n each; the R closure rebuilds diagmat) with A = U C X', B = V S X', C^2 + S^2 = I.  Init-only in
 * the reference (fast_BSFG_sampler_init.R:289,303,320); bsfg_init_lists is the faster, better-conditioned way to the same
 * lists.  Any output may be NULL.  Uses cuSOLVER (getrf/getrs/gesvd) through dlopen.                                    */
int bsfg_GSVD_2_c(const double* A, const double* B, int n, double* U_out, double* V_out, double* X_out, double* C_out,
                  double* S_out);

/* Process-wide default for the weighted-Gram evaluation used by bsfg_sample_Lambda_c and by contexts created with
 * gram_mode == 0 (see bsfg_config): mode 1 exact, 2 exponential sum (library default); terms/step 0 = keep.      */
int bsfg_set_gram_default(int mode, int terms, double step);

/* FP64 GEMM core used by everything above, exposed for testing and for `bench.py`'s roofline leg:
 * C (M x N, ldc) = alpha * At' * B + beta * C, At is K x M (lda), B is K x N (ldb); HOST pointers. */
int bsfg_dgemm_tn_host(int M, int N, int K, double alpha, const double* At, int lda, const double* B,
                       int ldb, double beta, double* C, int ldc);

/* ------------------------------------------------------------------------------------------
 * Stateful sweep: the loop body of fast_BSFG_sampler (fast_BSFG_sampler.R:176-270) with the state
 * resident on the device and the data pre-rotated.  One context per process/GPU.
 * ------------------------------------------------------------------------------------------ */
typedef struct bsfg_ctx bsfg_ctx;

typedef struct bsfg_config {
    int n;          /* individuals                                                        */
    int p_total;    /* traits over all ranks                                              */
    int p_offset;   /* first trait owned by this rank                                     */
    int p_local;    /* traits owned by this rank                                          */
    int k;          /* current number of factors                                          */
    int k_max;      /* capacity for update_k growth (>= k)                                */
    int r;          /* ncol(Z_1)                                                          */
    int b;          /* ncol(X) (>= 0)                                                     */
    int h2_divisions;
    int rank, world;         /* trait-sharding geometry (world == 1: no NCCL needed)      */
    unsigned long long seed; /* Philox key                                                */
    int z1_is_identity;      /* 1: Z_1 == I_n (r == n), Z_1 pointer may be NULL           */
    int mode;                /* 0 auto, 1 force "direct" residual/E_a forms, 2 force "diag" */
    /* weighted-Gram evaluation for the Lambda draw: 0 library default, 1 exact DMMA GEMM
     * over the n individuals, 2 exponential-sum factorisation (relative error ~1e-14, exact fallback on device
     * when `gram_terms` nodes cannot cover the dynamic range).  gram_terms / gram_step: 0 = default (256 / 0.28).
     * gram_terms = -1: adaptive -- the node count is re-chosen (64..256) at the start of every bsfg_sweep call from
     * the range seen at the end of the previous call; ~3% faster at C4, but the chain then depends (at the 1e-13
     * level) on where calls are split, so the default keeps a fixed count and a restarted chain is bit-identical to
     * a continuous one.                                                                                          */
    int gram_mode;
    int gram_terms;
    double gram_step;
    int r2;                  /* ncol(Z_2): levels of the second random effect W (0: none)          */
} bsfg_config;

typedef struct bsfg_priors {           /* model_setup.R:49-63, fast_BSFG_sampler_init.R:114 */
    double resid_Y_prec_shape, resid_Y_prec_rate;
    double E_a_prec_shape, E_a_prec_rate;
    double Lambda_df;
    double delta_1_shape, delta_1_rate, delta_2_shape, delta_2_rate;
    double b_X_prec;                   /* scalar on the diagonal of priors$b_X_prec */
    const double* h2_priors_factors;   /* h2_divisions entries */
    double b0, b1, epsilon, prop;      /* run_parameters used by update_k */
    double W_prec_shape, W_prec_rate;  /* second random effect (model_setup.R:53-54); used when r2 > 0 */
    double px_shape, px_rate;          /* parameter-expanded sampler (model_setup_px.R:55-56); used by bsfg_sweep_px */
} bsfg_priors;

/* Constants of one chain (host pointers; Y and everything trait-indexed holds THIS RANK's columns).  NaN entries of
 * Y are missing phenotypes (is.na(Y), fast_BSFG_sampler_init.R:69): they are re-imputed at the start of every
 * iteration (fast_BSFG_sampler.R:180-184) and the rotated data are rebuilt, which costs two n x n x p products per
 * iteration; set_state must then also supply Lambda and E_a (the imputation conditions on them).
 * Y n x p_local, X n x b, Z_1 n x r (NULL if identity), invert_aI_bZAZ {U n x n, s n},
 * invert_aPXA_bDesignDesignT {U br x br, s1, s2, Design_U n x br}, invert_aZZt_Ainv {U r x r, s1, s2},
 * Ainv r x r (used only by the "direct" E_a_prec form; may be NULL in "diag" mode).               */
typedef struct bsfg_constants {
    const double* Y;
    const double* X;
    const double* Z_1;
    const double* U1; const double* s;
    const double* U2; const double* s1_2; const double* s2_2; const double* Design_U;
    const double* U3; const double* s1_3; const double* s2_3;
    const double* Ainv;
    /* second random effect (r2 > 0): Z_2 n x r2, invert_aPXA_bDesignDesignT_rand2 {U r2 x r2, s1, s2}
     * (fast_BSFG_sampler_init.R:300-311; its Design_U = Z_2 U is recomputed on the device), A_2_inv r2 x r2
     * (NULL = identity, init:264).                                                                     */
    const double* Z_2;
    const double* U4; const double* s1_4; const double* s2_4;
    const double* A_2_inv;
} bsfg_constants;

/* current_state (fast_BSFG_sampler_init.R:238-254).  Trait-indexed members hold this rank's rows /
 * columns: Lambda, Lambda_prec, Plam p_local x k; B b x p_local; E_a r x p_local; precisions p_local.
 * On set_state only what the sweep conditions on before resampling is required: F, B, resid_Y_prec,
 * E_a_prec, Plam, delta, tauh (others may be NULL).  On get_state NULL members are skipped.        */
typedef struct bsfg_state {
    double* Lambda; double* Lambda_prec; double* Plam;
    double* F; double* F_a; double* F_h2;
    double* B; double* E_a;
    double* resid_Y_prec; double* E_a_prec;
    double* delta; double* tauh;
    int k;       /* in: number of factor columns supplied; out: current k */
    int nrun;    /* iteration counter i (current_state$nrun) */
    double* W; double* W_prec;   /* r2 x p_local, p_local (required on set_state when r2 > 0) */
    /* parameter-expanded sampler (fast_BSFG_sampler_px.R): px_factor k, F_px n x k.  Supplying both on set_state puts
     * the context in px mode (F is then ignored: F = F_px / sqrt(px_factor)); get_state returns them when non-NULL. */
    double* px_factor; double* F_px;
    /* multi-GPU set_state only (world > 1): 1 = the replicated members (F or F_px, F_a, F_h2, delta, tauh, px_factor) are
     * read from RANK 0's call and handed to the other ranks over NCCL (NVLink) instead of one PCIe upload per GPU; the
     * other ranks may pass NULL for them.  Every rank must make the call with the same flag, k and nrun.  0: each rank
     * uploads what it is given.  Ignored by get_state (NULL members are skipped there: only one rank needs to read the
     * replicated members back, they are bit-identical). */
    int bcast_replicated;
} bsfg_state;

/* One iteration's injected variates (NULL members fall back to device Philox).
 * z_Lambda k x p_local; z_means br x p_local; u_h2 k; z_Fa r x k; z_F n x k;
 * g_Lambda_prec p_local x k ~ Gamma((df+1)/2,1); g_resid p_local ~ Gamma(a+n/2,1);
 * g_Ea p_local ~ Gamma(a+r/2,1); g_delta: as bsfg_sample_delta_c; u_k: runif(1) of update_k.       */
typedef struct bsfg_variates {
    const double* z_Lambda; const double* z_means; const double* u_h2; const double* z_Fa;
    const double* z_F; const double* g_Lambda_prec; const double* g_resid; const double* g_Ea;
    const double* g_delta; const double* u_k;
    /* add-a-factor arm of update_k, R draw order BSFG_functions.R:334-341 */
    const double* k_g_lp; const double* k_g_delta; const double* k_z_lambda; const double* k_u_h2;
    const double* k_z_Fa; const double* k_z_F;
    /* second random effect: z_W r2 x p_local (= randn(br,p) of the second sample_means_c call,
     * fast_BSFG_sampler.R:214), g_W p_local ~ Gamma(a + r2/2, 1) (R:249) */
    const double* z_W; const double* g_W;
    /* missing phenotypes (NaN in Y): z_Y n x p_local N(0,1); element (i,j) is entry i*p + j of the reference's
     * rnorm(p*n) stream, which it lays out byrow (fast_BSFG_sampler.R:182) */
    const double* z_Y;
    /* parameter-expanded sampler: g_px k ~ Gamma(px_shape + n/2, 1) (fast_BSFG_sampler_px.R:263); k_g_px scalar
     * ~ Gamma(px_shape, 1) for the add-a-factor arm (BSFG_functions_px.R:333) */
    const double* g_px; const double* k_g_px;
} bsfg_variates;

/* update_k(current_state, priors, run_parameters, data_matrices)  (R only, BSFG_functions.R:311-372; with px_factor and
 * F_px (n x k_max) non-NULL the variant of BSFG_functions_px.R:301-369); device version named update_k_c.
 * Factor-indexed arrays are IN/OUT with capacity k_max columns and leading dimension = rows: Lambda, Lambda_prec, Plam
 * p x k_max; F n x k_max; F_a r x k_max; F_h2, delta, tauh, px_factor k_max entries; the first k (on return *k_out) columns
 * are meaningful, and nothing is written when k does not change.  i = current_state$nrun (R:323).  Z_1 n x r, or NULL for
 * the identity (r == n).  variates (or NULL members -> Philox keyed by (seed, i)): u_k (R:325) and the add-a-factor arm
 * k_g_lp p, k_g_delta, k_z_lambda p, k_u_h2, k_z_Fa r, k_z_F n (R:334-341; k_g_px with px_factor).                      */
int bsfg_update_k_c(int n, int p, int r, int k, int k_max, long long i, const bsfg_priors* priors, const double* Z_1,
                    double* Lambda, double* Lambda_prec, double* Plam, double* delta, double* tauh, double* F_h2, double* F,
                    double* F_a, double* px_factor, double* F_px, const bsfg_variates* variates, unsigned long long seed,
                    int* k_out);

/* Gamma (shape, rate) parameters of the last iteration's precision draws, for parity checks. */
typedef struct bsfg_aux {
    double* Lambda_prec_rate;   /* p_local x k */
    double* resid_Y_prec_rate;  /* p_local */
    double* E_a_prec_rate;      /* p_local */
    double* delta_rate;         /* k (draw order) */
    double* h2_log_ps;          /* k x H */
    double* W_prec_rate;        /* p_local (r2 > 0) */
    double* Y_imputed;          /* n x p_local: Y after the last iteration's imputation step (only with missing data) */
    double* px_rate;            /* k: rate of the px_factor draw (px mode) */
} bsfg_aux;

int bsfg_create(const bsfg_config* cfg, bsfg_ctx** out);
int bsfg_destroy(bsfg_ctx* ctx);
int bsfg_set_priors(bsfg_ctx* ctx, const bsfg_priors* pr);
int bsfg_upload_constants(bsfg_ctx* ctx, const bsfg_constants* c);
int bsfg_set_state(bsfg_ctx* ctx, const bsfg_state* st);
int bsfg_get_state(bsfg_ctx* ctx, bsfg_state* st);
int bsfg_get_aux(bsfg_ctx* ctx, bsfg_aux* aux);
/* n_iter Gibbs iterations with device Philox draws (the production path). */
int bsfg_sweep(bsfg_ctx* ctx, int n_iter);
/* exactly one iteration with injected variates (the parity path). */
int bsfg_sweep_injected(bsfg_ctx* ctx, const bsfg_variates* v);
/* Parameter-expanded sampler: the loop body of fast_BSFG_sampler_px.R:196-311 (F Lambda' = F_px Lambda_px' with a
 * redundant per-factor scale px_factor that is resampled every iteration; update_k of BSFG_functions_px.R:301-369).
 * The state must have been set with px_factor and F_px.  bsfg_sweep_px: n_iter iterations with device Philox draws;
 * bsfg_sweep_px_injected: one iteration with injected variates.                                                      */
int bsfg_sweep_px(bsfg_ctx* ctx, int n_iter);
int bsfg_sweep_px_injected(bsfg_ctx* ctx, const bsfg_variates* v);
/* one iteration of the fixed-Lambda sampler (fast_BSFG_sampler_fixedlambda.R:159-236): Lambda (p_local x k, k = the
 * state's k) is a stored posterior draw; B/E_a, W, F_h2, F_a, F and the precisions are resampled, Lambda, Lambda_prec,
 * delta and k are not.  The (B,E_a) step also removes X B like that driver does (its line 180).  variates may be NULL
 * (device Philox) or carry the members used by those steps.                                                         */
int bsfg_sweep_fixed_lambda(bsfg_ctx* ctx, const double* Lambda, int k, const bsfg_variates* variates);
/* ---- posterior bookkeeping on the device: save_posterior_samples (BSFG_functions.R:375-408) without a device->host
 * copy per saved sample.  bsfg_posterior_begin reserves `capacity` sample slots (and clears earlier ones);
 * bsfg_sweep_collect runs n_iter iterations and, after every iteration i with (i - burn) %% thin == 0 && i > burn
 * (fast_BSFG_sampler.R:272), stores sample sp_num = (i - burn)/thin: full traces of Lambda, F, F_a, delta, F_h2 and
 * the precisions, running means of B, E_a, W.  bsfg_get_posterior copies out what was collected:
 *   traces, one column per sample in slot order: Lambda (p_local*k_max), F (n*k_max), F_a (r*k_max), delta (k_max),
 *   F_h2 (k_max) -- each sample is the column-major vec of the (rows x k) matrix, zero padded to k_max columns like the
 *   reference pads when k grew (:384-391); resid_Y_prec, E_a_prec, W_prec (p_local);
 *   means: B (b x p_local), E_a (r x p_local), W (r2 x p_local); sp_num[s] = the sample number of slot s, k[s] = the
 *   number of factors it had.
 * NULL members are skipped.                                                                                         */
typedef struct bsfg_posterior {
    double* Lambda; double* F; double* F_a; double* delta; double* F_h2;
    double* resid_Y_prec; double* E_a_prec; double* W_prec;
    double* B; double* E_a; double* W;
    int* sp_num; int* k;
} bsfg_posterior;
int bsfg_posterior_begin(bsfg_ctx* ctx, int capacity);
int bsfg_sweep_collect(bsfg_ctx* ctx, int n_iter, int burn, int thin);
int bsfg_posterior_count(bsfg_ctx* ctx, int* n_samples, int* k_max);
int bsfg_get_posterior(bsfg_ctx* ctx, bsfg_posterior* out);

/* which residual / E_a_prec formulation the context selected (1 direct, 2 diag) and the measured
 * ||U2' P U2 - diag(s1)||_max, ||Design_U' Design_U - diag(s2)||_max that drove the choice.        */
int bsfg_get_mode(bsfg_ctx* ctx, int* mode, double* resid_s1, double* resid_s2);
/* weighted-Gram evaluation in use, number of iterations that fell back to the exact GEMM because the exponential
 * sum could not cover the range of E_a_prec/resid_Y_prec, that range at the last iteration, and (expsum mode) the
 * device milliseconds of the two GEMMs of the last timed iteration: G'A (nodes) and (G'A)'E (traits).              */
int bsfg_get_gram_stats(bsfg_ctx* ctx, int* mode, int* terms, double* step, long long* fallbacks,
                        double* c_min, double* c_max, double* ms_nodes, double* ms_traits);
/* timing of the last bsfg_sweep call: total device milliseconds (CUDA events on the sweep stream),
 * kernel launches issued, and per-phase milliseconds (array of BSFG_N_PHASES doubles, may be NULL) */
#define BSFG_N_PHASES 12
int bsfg_get_timing(bsfg_ctx* ctx, double* total_ms, long long* launches, double* phase_ms);

/* ------------------------------------------------------------------------------------------
 * Device init: the precomputed lists of fast_BSFG_sampler_init.R:263-325 (Ainv, svd(Z_1 A Z_1'), the three GSVD
 * diagonalisations and their Design_U), computed on the GPU with cuSOLVER/cuBLAS (loaded on first use).  Same list
 * contract as the reference (inv(aP + bD'D) = U diag(1/(a s1 + b s2)) U'), obtained through symmetric
 * eigen-decompositions instead of GSVD_2_c's svd(A inv(B)) (cpp:27-40); U is not unique, the contract is.
 * Host pointers; Z_1 NULL = identity (r == n); X n x b; Z_2 n x r2 (r2 = 0: none); NULL outputs are skipped.
 * Sizes: Ainv r x r; U1 n x n, s n; U2 br x br, s1_2, s2_2 br, Design_U n x br; U3 r x r, s1_3, s2_3 r;
 * U4 r2 x r2, s1_4, s2_4 r2, Design_U4 n x r2.
 * ------------------------------------------------------------------------------------------ */
typedef struct bsfg_init_out {
    double* Ainv;
    double* U1; double* s;
    double* U2; double* s1_2; double* s2_2; double* Design_U;
    double* U3; double* s1_3; double* s2_3;
    double* U4; double* s1_4; double* s2_4; double* Design_U4;
} bsfg_init_out;
int bsfg_init_lists(int n, int r, int b, int r2, const double* Z_1, const double* A, const double* X,
                    const double* Z_2, double b_X_prec, bsfg_init_out* out);

/* multi-GPU: rank 0 calls bsfg_nccl_unique_id, broadcasts the 128 bytes out of band (the Python host
 * uses torch.distributed), every rank calls bsfg_comm_init before bsfg_sweep.                    */
int bsfg_nccl_unique_id(char id_out[128]);
int bsfg_comm_init(bsfg_ctx* ctx, const char id[128]);

#ifdef __cplusplus
}
#endif
#endif /* BSFG_B200_H */
