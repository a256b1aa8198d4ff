// topN on the GPU (reference src/common.c:5127-5369).
#pragma once
#include "cmf_types.h"
namespace cmfb200 {
int top_n(real_t *a_vec, int_t k_user, real_t *B, int_t k_item, real_t *biasB, real_t glob_mean, real_t biasA, int_t k,
          int_t k_main, int_t *include_ix, int_t n_include, int_t *exclude_ix, int_t n_exclude, int_t *outp_ix,
          real_t *outp_score, int_t n_top, int_t n, int nthreads);
}
