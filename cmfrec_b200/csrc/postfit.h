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
    // the collective model (dense side information and / or implicit features; k_user = k_item = k_main = 0)
    const real_t *C = nullptr; int p = 0; real_t w_user = 1;           // [p x kk]
    const real_t *Bi = nullptr; bool implicit_features = false; real_t w_implicit = 1;   // [n x kk]
    bool scale_lam_sideinfo = false;
    real_t *BiTBi = nullptr;            // [kk x kk]   w_implicit Bi^T Bi, upper triangle
    real_t *TransCtCinvCt = nullptr;    // [p x kk]    C (C^T C + lam (scale_lam ? p : 1) / w_user I)^-1
    real_t *CtCw = nullptr;             // [kk x kk]   w_user C^T C, upper triangle
    real_t *BeTBeChol = nullptr;        // [(kk+ub)^2] Cholesky factor of BtB + CtCw + BiTBi + regulariser, upper triangle
};

int postfit_explicit(const PostfitExplicit &a);
// implicit model: BtB + lam I; with side information also BeTBe = w_user C^T C + BtB + lam I and its Cholesky factor
// last_was_cg: the last A update ran CG -- the reference then assembles BeTBe from an UNSCALED C^T C (its scaling line is
// guarded by `w_user == 1.`, src/collective.c:10078-10079), which is reproduced
int postfit_implicit(const real_t *B, int_t n, int kk, real_t lam, real_t *BtB, const real_t *C = nullptr, int p = 0,
                     real_t w_user = 1, real_t *BeTBe = nullptr, real_t *BeTBeChol = nullptr, bool last_was_cg = false);

// in-place Cholesky solve of S X = R for `nrhs` right-hand sides stored as rows of R (each of length d);
// S is d x d symmetric, upper triangle (row-major) is read.  Returns nonzero if S is not positive definite.
int host_spd_solve_rows(std::size_t d, const real_t *S_upper, real_t *R, std::size_t nrhs);

}  // namespace cmfb200
