// Host-side preparation routines (see host_prep.cpp for the reference lines each one follows).
#pragma once
#include "cmf_types.h"

namespace cmfb200 {

void seed_state(int_t seed, uint64_t state[4]);
void fill_normal(real_t *out, size_t n, uint64_t state[4]);
void fill_uniform(real_t *out, size_t n, uint64_t state[4]);
void random_init(real_t *A, size_t sizeA, real_t *B, size_t sizeB, int_t seed, bool normal, int nthreads = 1);

void coo_to_csr_and_csc(const int_t *row, const int_t *col, const real_t *val, int_t m, int_t n, size_t nnz,
                        size_t *csr_p, int_t *csr_i, real_t *csr_v, size_t *csc_p, int_t *csc_i, real_t *csc_v);

real_t global_mean(const real_t *X, size_t nnz, int nthreads);

void init_biases_twosided(int_t m, int_t n, const size_t *csr_p, const int_t *csr_i, const real_t *csr_v,
                          const size_t *csc_p, const int_t *csc_i, const real_t *csc_v, real_t lam_user,
                          real_t lam_item, bool scale_lam, bool nonneg, real_t *biasA, real_t *biasB, int nthreads);

void init_biases_onesided(int_t m, const size_t *csr_p, const real_t *csr_v, real_t lam, bool scale_lam, bool nonneg,
                          real_t *bias);

}  // namespace cmfb200
