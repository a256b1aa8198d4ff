// fit_most_popular on the GPU (reference src/common.c:5371-5699, 5703-6102).
#pragma once
#include "cmf_types.h"
namespace cmfb200 {
int most_popular(real_t *biasA, real_t *biasB, real_t *glob_mean, real_t lam_user, real_t lam_item, bool scale_lam,
                 bool scale_bias_const, real_t alpha, int_t m, int_t n, int_t *ixA, int_t *ixB, real_t *X, size_t nnz,
                 real_t *Xfull, real_t *weight, bool implicit, bool adjust_weight, bool apply_log_transf, bool nonneg,
                 bool NA_as_zero, real_t *w_main_multiplier, int nthreads);
}
