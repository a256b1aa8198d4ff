// Small dense / sparse-dense building blocks for the models with side information (see dense_small.cu).
#pragma once
#include <vector>
#include <cuda_runtime.h>
#include "sweep.h"

namespace cmfb200 {

// out[ncx x ncy] = X[:, :ncx]^T Y[:, :ncy] over `rows` rows (split over row slices, fixed-order reduction)
size_t xty_workspace_elems(int ncx, int ncy);
int launch_xty(const real_t *X, int ldx, int ncx, const real_t *Y, int ldy, int ncy, int_t rows, real_t *out, real_t *workspace,
               cudaStream_t stream);
// out[r, :nc] = alpha * M[r, :p] S[:p, :nc] (+ out)
int launch_rows_times_small(const real_t *M, int ldm, int p, const real_t *S, int lds, int nc, real_t alpha, bool accumulate,
                            real_t *out, int ldo, int_t rows, cudaStream_t stream);
// Y[r, :kk] = alpha * sum over the stored entries of row r of F[col, :kk] (+ Y)     (reference tgemm_sp_dense with
// all-ones values, src/helpers.c:1135)
int launch_spmm_ones(const CsrView &X, const SweepPlan &plan, const real_t *F, int ldf, int kk, real_t alpha, bool accumulate,
                     real_t *Y, int ldy, cudaStream_t stream);
// Cholesky factor (lower, row-major) of a small SPD matrix given on the host; computed in double
int spd_factor_host(int d, const real_t *S_host, std::vector<real_t> &L_host);
// L_dev := Cholesky factor (lower, row-major) of S_dev + lam*I, all on the device (single thread block)
int launch_spd_factor(const real_t *S_dev, int d, real_t lam, real_t *L_dev, cudaStream_t stream);
// every row of R := (L L^T)^-1 row
int launch_tri_solve_rows(const real_t *L_dev, int d, real_t *R, int ldr, int_t rows, cudaStream_t stream);
// out = alpha*x + beta*y (either may be null)
int launch_axpby(int n, real_t alpha, const real_t *x, real_t beta, const real_t *y, real_t *out, cudaStream_t stream);

}  // namespace cmfb200
