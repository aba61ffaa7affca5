"""sparsefactormixedmodel_b200 -- B200-native BSFG Gibbs sweep behind the reference's Rcpp interface.

The compute lives in `libbsfg_b200.so` (hand-written sm_100a CUDA, C ABI in include/bsfg_b200.h).
This package is the Python host-side mirror of the reference's `_c` functions and of the
`fast_BSFG_sampler` driver contract.  Importing the package does not need a GPU; calling any sampler
without the built library or without a CUDA device raises.
"""
from ._lib import BSFGError, LIB_PATH, load  # noqa: F401
from .samplers import (GSVD_2_c, dgemm_tn, update_k_c, sample_delta_c, sample_F_a_c, sample_factors_scores_c,  # noqa: F401
                       sample_h2s_discrete_c, sample_Lambda_c, sample_lambda_c, sample_means_c,
                       sample_prec_discrete_conditional_c, sample_px_factors_scores_c)
from .driver import BSFGSweep, fast_BSFG_sampler, init_lists, set_gram_default  # noqa: F401


def k_limit():
    """Largest supported number of factors (k and k_max)."""
    return int(load().bsfg_k_limit())
