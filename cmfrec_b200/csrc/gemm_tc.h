// C[M x N] = A[M x K] B[N x K]^T (+ row_term[m] + col_term[n] + add_const) on the tcgen05 tensor cores (gemm_tc.cu; fp32 library).
#pragma once
#include <cuda_runtime.h>
#include "cmf_types.h"
namespace cmfb200 {
// device pointers; row_term / col_term may be null.  Returns 0, 1 (CUDA error) or 3 (not covered: fp64 library, K < 1, grid limits)
int launch_gemm_nt_tc(const real_t *A, int lda, long long M, const real_t *B, int ldb, long long N, int K, real_t *C, long long ldc,
                      const real_t *row_term, const real_t *col_term, real_t add_const, cudaStream_t stream);
}
