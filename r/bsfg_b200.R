# Drop-in replacement for  sourceCpp('BSFG_functions_c.cpp')  (R_BSFG/model_setup.R:15).
#   source('bsfg_b200.R')            # after dyn.load of the glue; defines the same eight closures
# Same names and argument order as the Rcpp exports (BSFG_functions_c.cpp:26,42,85,137,165,195,240,292), so
# fast_BSFG_sampler.R:191-232 runs unmodified.  Extra trailing arguments, all optional:
#   variates = NULL   device Philox draws (production);  or a numeric matrix in the reference's layout, e.g.
#                     matrix(rnorm(k*p), k, p) -- the testing hook of BSFG_functions_c.cpp:111-114
#   seed              Philox key for this call (default: drawn from R's RNG so set.seed() governs the chain)
dyn.load(file.path(Sys.getenv('BSFG_B200_HOME', '.'), 'sparsefactormixedmodel_b200', 'libbsfg_b200.so'), local = FALSE)
dyn.load(file.path(Sys.getenv('BSFG_B200_HOME', '.'), 'r', 'bsfg_glue.so'))
.bsfg_seed <- function() floor(runif(1) * 2^52)

sample_Lambda_c <- function(Y_tilde, F, resid_Y_prec, E_a_prec, Plam, invert_aI_bZAZ, variates = NULL, seed = .bsfg_seed())
  .Call('R_bsfg_sample_Lambda_c', Y_tilde, F, as.double(resid_Y_prec), as.double(E_a_prec), Plam, invert_aI_bZAZ, variates, seed)
sample_lambda_c <- sample_Lambda_c
sample_means_c <- function(Y_tilde, resid_Y_prec, E_a_prec, invert_aPXA_bDesignDesignT, variates = NULL, seed = .bsfg_seed())
  .Call('R_bsfg_sample_means_c', Y_tilde, as.double(resid_Y_prec), as.double(E_a_prec), invert_aPXA_bDesignDesignT, variates, seed)
sample_factors_scores_c <- function(Y_tilde, Z_1, Lambda, resid_Y_prec, F_a, F_h2, variates = NULL, seed = .bsfg_seed())
  .Call('R_bsfg_sample_factors_scores_c', Y_tilde, Z_1, Lambda, as.double(resid_Y_prec), F_a, as.double(F_h2), variates, seed)
sample_px_factors_scores_c <- function(Y_tilde, Z_1, Lambda_px, resid_Y_prec, F_a_px, tau_e, px_factor, variates = NULL, seed = .bsfg_seed())
  .Call('R_bsfg_sample_px_factors_scores_c', Y_tilde, Z_1, Lambda_px, as.double(resid_Y_prec), F_a_px, as.double(tau_e), as.double(px_factor), variates, seed)
sample_h2s_discrete_c <- function(F, h2_divisions, h2_priors, invert_aI_bZAZ, variates = NULL, seed = .bsfg_seed())
  .Call('R_bsfg_sample_h2s_discrete_c', F, as.integer(h2_divisions), as.double(h2_priors), invert_aI_bZAZ, variates, seed)
sample_prec_discrete_conditional_c <- function(Y, h2_divisions, h2_priors, invert_aI_bZAZ, res_prec, variates = NULL, seed = .bsfg_seed())
  .Call('R_bsfg_sample_prec_discrete_conditional_c', Y, as.integer(h2_divisions), as.double(h2_priors), invert_aI_bZAZ, as.double(res_prec), variates, seed)
sample_F_a_c <- function(F, Z_1, F_h2, invert_aZZt_Ainv, variates = NULL, seed = .bsfg_seed())
  .Call('R_bsfg_sample_F_a_c', F, Z_1, as.double(F_h2), invert_aZZt_Ainv, variates, seed)
# device version of the R-only sample_delta (BSFG_functions.R:272); same argument order
sample_delta_c <- function(delta, tauh, Lambda_prec, delta_1_shape, delta_1_rate, delta_2_shape, delta_2_rate, Lambda2, variates = NULL, seed = .bsfg_seed())
  .Call('R_bsfg_sample_delta_c', as.double(delta), as.double(tauh), Lambda_prec, delta_1_shape, delta_1_rate, delta_2_shape, delta_2_rate, Lambda2, variates, seed)
# GSVD_2_c(A, B) (BSFG_functions_c.cpp:27-40): list(U, V, X, C, S)
GSVD_2_c <- function(A, B) .Call('R_bsfg_GSVD_2_c', A, B)
# device version of the R-only update_k (BSFG_functions.R:311-372; BSFG_functions_px.R:301-369 when current_state has px_factor);
# same argument list, returns current_state
update_k_c <- function(current_state, priors, run_parameters, data_matrices, seed = .bsfg_seed()) {
  cs <- current_state; k <- ncol(cs$F); n <- nrow(cs$F); p <- nrow(cs$Lambda); r <- nrow(cs$F_a)
  pad <- function(m) cbind(m, 0) + 0; padv <- function(v) c(as.double(v), 0)
  b <- list(Lambda = pad(cs$Lambda), Lambda_prec = pad(cs$Lambda_prec), Plam = pad(cs$Plam), F = pad(cs$F), F_a = pad(cs$F_a),
            delta = padv(cs$delta), tauh = padv(cs$tauh), F_h2 = padv(cs$F_h2))
  px <- !is.null(cs$px_factor)
  if (px) { b$px_factor <- padv(cs$px_factor); b$F_px <- pad(cs$F_px) }
  pr <- priors; rp <- run_parameters
  pri <- c(pr$resid_Y_prec_shape, pr$resid_Y_prec_rate, pr$E_a_prec_shape, pr$E_a_prec_rate, pr$Lambda_df,
           pr$delta_1_shape, pr$delta_1_rate, pr$delta_2_shape, pr$delta_2_rate, pr$b_X_prec[1], rp$b0, rp$b1, rp$epsilon, rp$prop,
           2, 0.1, if (px) pr$px_shape else 0.5, if (px) pr$px_rate else 0.5)
  Z_1 <- data_matrices$Z_1
  ident <- nrow(Z_1) == ncol(Z_1) && all(Z_1 == diag(nrow(Z_1)))
  kn <- .Call('R_bsfg_update_k_c', as.integer(c(n, p, r, k, k + 1, cs$nrun)), pri, if (ident) NULL else Z_1, b$Lambda, b$Lambda_prec,
              b$Plam, b$delta, b$tauh, b$F_h2, b$F, b$F_a, b$px_factor, b$F_px, seed)
  if (kn != k) {
    for (nm in c('Lambda', 'Lambda_prec', 'Plam', 'F', 'F_a', if (px) 'F_px')) cs[[nm]] <- b[[nm]][, 1:kn, drop = FALSE]
    for (nm in c('delta', 'tauh', if (px) 'px_factor')) cs[[nm]] <- b[[nm]][1:kn]
    cs$F_h2 <- matrix(b$F_h2[1:kn], kn, 1); cs$k <- kn
  }
  cs
}

# fast_BSFG_sampler with the loop body (fast_BSFG_sampler.R:176-270) on the GPU.  Prologue/epilogue of the
# reference driver (state unpacking :56-152, Posterior growth :159-169, save/return :311-346) stay as they are;
# only the `for(i in start_i+(1:n_samples))` loop is replaced: the state is uploaded once, iterated on the device
# in chunks that end on thinning boundaries, and read back where the reference calls save_posterior_samples.
bsfg_sweep_loop <- function(BSFG_state, n_samples, k_max = NULL) {
  dm <- BSFG_state$data_matrices; rv <- BSFG_state$run_variables; pr <- BSFG_state$priors
  rp <- BSFG_state$run_parameters; cs <- BSFG_state$current_state
  k <- ncol(cs$F); if (is.null(k_max)) k_max <- max(2 * k, k + 8)
  dims <- as.integer(c(rv$n, rv$p, k, k_max, rv$r, rv$b, rp$h2_divisions, rv$r2))
  pri <- c(pr$resid_Y_prec_shape, pr$resid_Y_prec_rate, pr$E_a_prec_shape, pr$E_a_prec_rate, pr$Lambda_df,
           pr$delta_1_shape, pr$delta_1_rate, pr$delta_2_shape, pr$delta_2_rate, pr$b_X_prec[1], rp$b0, rp$b1, rp$epsilon, rp$prop,
           pr$W_prec_shape, pr$W_prec_rate)
  ident <- nrow(dm$Z_1) == ncol(dm$Z_1) && all(dm$Z_1 == diag(nrow(dm$Z_1)))
  ctx <- .Call('R_bsfg_create', dims, pri, as.double(pr$h2_priors_factors), dm$Y, dm$X, dm$Z_1, rv$invert_aI_bZAZ,
               rv$invert_aPXA_bDesignDesignT, rv$invert_aZZt_Ainv, rv$Ainv, ident, .bsfg_seed(), dm$Z_2,
               rv$invert_aPXA_bDesignDesignT_rand2)
  .Call('R_bsfg_set_state', ctx, cs)
  i <- cs$nrun; end <- i + n_samples; sp_num <- ncol(BSFG_state$Posterior$Lambda)
  pull <- function() {
    kn <- .Call('R_bsfg_get_state', ctx, NULL)
    k <- kn[1]
    out <- list(Lambda = matrix(0, rv$p, k), Lambda_prec = matrix(0, rv$p, k), Plam = matrix(0, rv$p, k), F = matrix(0, rv$n, k),
                F_a = matrix(0, rv$r, k), F_h2 = matrix(0, k, 1), B = matrix(0, rv$b, rv$p), E_a = matrix(0, rv$r, rv$p),
                resid_Y_prec = numeric(rv$p), E_a_prec = numeric(rv$p), delta = numeric(k), tauh = numeric(k),
                W = matrix(0, rv$r2, rv$p), W_prec = numeric(rv$p))
    kn <- .Call('R_bsfg_get_state', ctx, out)                     # fills the preallocated members in place
    out$nrun <- kn[2]
    out
  }
  while (i < end) {
    nxt <- end
    j <- max(i + 1, rp$burn + 1); while ((j - rp$burn) %% rp$thin != 0) j <- j + 1
    if (j <= end) nxt <- j
    .Call('R_bsfg_sweep', ctx, as.integer(nxt - i)); i <- nxt
    if ((i - rp$burn) %% rp$thin == 0 && i > rp$burn) {            # fast_BSFG_sampler.R:272-279
      sp_num <- (i - rp$burn) / rp$thin
      s <- pull()                                                  # BSFG_functions.R:375, signature unchanged
      BSFG_state$Posterior <- save_posterior_samples(sp_num, BSFG_state$Posterior, s$Lambda, s$F, s$F_a, s$B, s$W, s$E_a,
                                                     s$delta, s$F_h2, s$resid_Y_prec, s$E_a_prec, s$W_prec)
    }
  }
  BSFG_state$current_state <- pull()
  BSFG_state
}
