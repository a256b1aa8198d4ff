// Argument blocks of the two ALS fit entry points; field names follow the reference's parameter names
// (reference src/cmfrec.h:1851-1921) so that capi.cu is a 1:1 forwarding layer.
#pragma once
#include "cmf_types.h"

namespace cmfb200 {

struct ExplicitArgs {
    real_t *biasA, *biasB, *A, *B, *C, *D, *Ai, *Bi;
    bool add_implicit_features, reset_values;
    int_t seed;
    real_t *glob_mean, *U_colmeans, *I_colmeans;
    int_t m, n, k;
    int_t *ixA, *ixB;
    real_t *X;
    size_t nnz;
    real_t *Xfull, *weight;
    bool user_bias, item_bias, center;
    real_t lam, *lam_unique, l1_lam, *l1_lam_unique;
    bool scale_lam, scale_lam_sideinfo, scale_bias_const;
    real_t *scaling_biasA, *scaling_biasB;
    real_t *U; int_t m_u, p;
    real_t *II; int_t n_i, q;
    int_t *U_row, *U_col; real_t *U_sp; size_t nnz_U;
    int_t *I_row, *I_col; real_t *I_sp; size_t nnz_I;
    bool NA_as_zero_X, NA_as_zero_U, NA_as_zero_I;
    int_t k_main, k_user, k_item;
    real_t w_main, w_user, w_item, w_implicit;
    int_t niter; int nthreads;
    bool verbose, handle_interrupt, use_cg;
    int_t max_cg_steps;
    bool precondition_cg, finalize_chol, nonneg;
    int_t max_cd_steps;
    bool nonneg_C, nonneg_D, precompute_for_predictions, include_all_X;
    real_t *B_plus_bias, *precomputedBtB, *precomputedTransBtBinvBt, *precomputedBtXbias, *precomputedBeTBeChol,
        *precomputedBiTBi, *precomputedTransCtCinvCt, *precomputedCtCw, *precomputedCtUbias;
};

struct ImplicitArgs {
    real_t *A, *B, *C, *D;
    bool reset_values;
    int_t seed;
    real_t *U_colmeans, *I_colmeans;
    int_t m, n, k;
    int_t *ixA, *ixB;
    real_t *X;
    size_t nnz;
    real_t lam, *lam_unique, l1_lam, *l1_lam_unique;
    real_t *U; int_t m_u, p;
    real_t *II; int_t n_i, q;
    int_t *U_row, *U_col; real_t *U_sp; size_t nnz_U;
    int_t *I_row, *I_col; real_t *I_sp; size_t nnz_I;
    bool NA_as_zero_U, NA_as_zero_I;
    int_t k_main, k_user, k_item;
    real_t w_main, w_user, w_item;
    real_t *w_main_multiplier;
    real_t alpha;
    bool adjust_weight, apply_log_transf;
    int_t niter; int nthreads;
    bool verbose, handle_interrupt, use_cg;
    int_t max_cg_steps;
    bool precondition_cg, finalize_chol, nonneg;
    int_t max_cd_steps;
    bool nonneg_C, nonneg_D, precompute_for_predictions;
    real_t *precomputedBtB, *precomputedBeTBe, *precomputedBeTBeChol, *precomputedCtUbias;
};

int fit_explicit(const ExplicitArgs &a);
int fit_implicit(const ImplicitArgs &a);

}  // namespace cmfb200
