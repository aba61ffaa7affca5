/* The stateless entry points of the mock (no header on purpose: old-style definitions that accept any argument list). */
int bsfg_sample_Lambda_c() { return 0; }
int bsfg_sample_means_c() { return 0; }
int bsfg_sample_factors_scores_c() { return 0; }
int bsfg_sample_px_factors_scores_c() { return 0; }
int bsfg_sample_h2s_discrete_c() { return 0; }
int bsfg_sample_prec_discrete_conditional_c() { return 0; }
int bsfg_sample_F_a_c() { return 0; }
int bsfg_sample_delta_c() { return 0; }
int bsfg_GSVD_2_c() { return 0; }
int bsfg_update_k_c() { return 0; }
