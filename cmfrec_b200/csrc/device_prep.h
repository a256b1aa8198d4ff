// Device-side ingestion (see device_prep.cu).  All pointers are device pointers.
#pragma once
#include <cuda_runtime.h>
#include "als.h"

namespace cmfb200 {

// stable compression of COO triplets by `major` index: ptr[nmajor+1], idx/out[nnz]
int device_compress(const int_t *d_major, const int_t *d_minor, const real_t *d_val, size_t nnz, int_t nmajor,
                    size_t *d_ptr, int_t *d_idx, real_t *d_out, cudaStream_t stream);
int device_degree_order(const size_t *d_ptr, int_t rows, int_t *d_order, std::vector<int_t> &deg_sorted, size_t *nnz_total,
                        cudaStream_t stream);
int device_all_positive(const real_t *d_x, size_t n, bool *all_positive, cudaStream_t stream);
int device_subtract(real_t *d_x, size_t n, real_t mu, cudaStream_t stream);
int device_init_biases_twosided(int_t m, int_t n, const size_t *csr_p, const int_t *csr_i, const real_t *csr_v,
                                const size_t *csc_p, const int_t *csc_i, const real_t *csc_v, real_t lam_user, real_t lam_item,
                                bool scale_lam, real_t *d_biasA, real_t *d_biasB, cudaStream_t stream);
int device_init_biases_onesided(int_t rows, const size_t *ptr, const real_t *val, real_t lam, bool scale_lam, real_t *d_bias,
                                cudaStream_t stream);

}  // namespace cmfb200
