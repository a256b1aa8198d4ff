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
// row dealing for the multi-GPU path (all device-side; see als.cu: setup_from_coo)
int device_deal_rows(const int_t *d_order, int_t rows, int world, int_t block, int_t *d_to_dev, int_t *d_to_old, cudaStream_t stream);
int device_extract_block(const size_t *ptr_full, const int_t *idx_full, const real_t *val_full, const int_t *to_old,
                         const int_t *other_to_dev, int_t row_begin, int_t row_end, int_t n_local, int_t rows_padded,
                         DevBuf<size_t> &ptr_local, DevBuf<int_t> &idx_local, DevBuf<real_t> &val_local, size_t *nnz_local,
                         cudaStream_t stream);
// dst[to_dev[r]] = src[r] for kk columns (to_dev null = identity) and the inverse
int device_scatter_rows(const real_t *src, int lds, int_t rows, int kk, const int_t *to_dev, real_t *dst, int ldd, cudaStream_t stream);
int device_gather_rows_back(const real_t *src, int lds, int_t rows, int kk, const int_t *to_dev, real_t *dst, int ldd, cudaStream_t stream);
int device_all_positive(const real_t *d_x, size_t n, bool *all_positive, cudaStream_t stream);
int device_subtract(real_t *d_x, size_t n, real_t mu, cudaStream_t stream);
int device_init_biases_twosided(int_t m, int_t n, const size_t *csr_p, const int_t *csr_i, const real_t *csr_v,
                                const size_t *csc_p, const int_t *csc_i, const real_t *csc_v, real_t lam_user, real_t lam_item,
                                bool scale_lam, real_t *d_biasA, real_t *d_biasB, cudaStream_t stream);
int device_init_biases_onesided(int_t rows, const size_t *ptr, const real_t *val, real_t lam, bool scale_lam, real_t *d_bias,
                                cudaStream_t stream);

}  // namespace cmfb200
