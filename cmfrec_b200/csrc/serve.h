// Batched serving on the GPU: predictions for (row, column) pairs and top-N items for many users at once, from factor
// matrices resident in HBM.  reference predict_multiple src/common.c:5066-5112, topN src/common.c:5127-5369 and their
// collective-model wrappers src/collective.c:11546-11614, 11797-11862.
#pragma once
#include "cmf_types.h"
namespace cmfb200 {
struct ServeState;
ServeState *serve_create(const real_t *A, int_t m, int_t k_user, const real_t *B, int_t n, int_t k_item, const real_t *biasA,
                         const real_t *biasB, real_t glob_mean, int_t k, int_t k_main, int *rc);
void serve_destroy(ServeState *s);
// out[i] = <A[row[i]], B[col[i]]> + biasA + biasB + glob_mean; NaN when an index is out of range (host arrays)
int serve_predict(ServeState *s, const int_t *row, const int_t *col, size_t n_predict, real_t *out);
// top-n items of each listed user, excluding the user's seen items (CSR over the listed users; null = nothing excluded);
// out_ix [n_users x n_top], out_score [n_users x n_top] or null.  Ties go to the lower item id.
int serve_topn(ServeState *s, const int_t *users, int_t n_users, const size_t *seen_ptr, const int_t *seen_idx, int_t n_top,
               int_t *out_ix, real_t *out_score, float *ms_device);
// host-pointer entry points with the reference's argument lists
int predict_multiple_host(real_t *A, int_t k_user, real_t *B, int_t k_item, real_t *biasA, real_t *biasB, real_t glob_mean, int_t k,
                          int_t k_main, int_t m, int_t n, int_t *predA, int_t *predB, size_t nnz, real_t *outp);
}
