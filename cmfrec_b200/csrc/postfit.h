// Matrices the reference leaves behind for later predictions on new users when
// precompute_for_predictions is set (reference src/collective.c:8935-9255 explicit, :10055-10130 implicit).
// Small dense k x k work done once after the fit, on the host.
#pragma once
#include "cmf_types.h"

namespace cmfb200 {

struct PostfitExplicit {
    const real_t *B; const real_t *biasB; int_t n; int kk;
    bool user_bias, item_bias;
    real_t lam, lam_bias; bool scale_lam;
    real_t *B_plus_bias;        // [n x (kk+1)] or null
    real_t *BtB;                // [(kk+ub) x (kk+ub)] upper triangle written, or null
    real_t *TransBtBinvBt;      // [n x (kk+ub)] or null
};

int postfit_explicit(const PostfitExplicit &a);
int postfit_implicit(const real_t *B, int_t n, int kk, real_t lam, real_t *BtB);

// in-place Cholesky solve of S X = R for `nrhs` right-hand sides stored as rows of R (each of length d);
// S is d x d symmetric, upper triangle (row-major) is read.  Returns nonzero if S is not positive definite.
int host_spd_solve_rows(std::size_t d, const real_t *S_upper, real_t *R, std::size_t nrhs);

}  // namespace cmfb200
