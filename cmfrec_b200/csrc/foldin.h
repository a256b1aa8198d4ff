// Factors for NEW rows given a fitted model ("fold-in") and the matrices precomputed for it, on the GPU:
//   reference factors_collective_explicit_multiple src/collective.c:10865, factors_collective_implicit_multiple :11176
//   (per row: collective_factors_warm :3555 -> factors_closed_form, collective_factors_warm_implicit :3966 ->
//   factors_implicit_chol), precompute_collective_explicit :10209, precompute_collective_implicit :10487.
// The per-row solves are the exact (Cholesky) half-sweep of the fit with B fixed: the same kernels, driven by a CSR of the
// new rows.  Covered: sparse X (COO or CSR), no side information; everything else is refused with code 2.
#pragma once
#include "cmf_types.h"
namespace cmfb200 {
struct FoldinExplicitArgs {
    real_t *A, *biasA; int_t m;
    const int_t *ixA, *ixB; const real_t *X; size_t nnz;
    const size_t *Xcsr_p; const int_t *Xcsr_i; const real_t *Xcsr;
    const real_t *B, *biasB; int_t n, n_max; bool include_all_X;
    real_t glob_mean;
    int_t k, k_main;
    real_t lam; const real_t *lam_unique;
    bool scale_lam, scale_lam_sideinfo, scale_bias_const; real_t scaling_biasA;
    real_t w_main;
};
int foldin_explicit(const FoldinExplicitArgs &a);
struct FoldinImplicitArgs {
    real_t *A; int_t m;
    const int_t *ixA, *ixB; const real_t *X; size_t nnz;
    const size_t *Xcsr_p; const int_t *Xcsr_i; const real_t *Xcsr;
    const real_t *B; int_t n;
    int_t k, k_main;
    real_t lam, alpha, w_main, w_main_multiplier;
    bool apply_log_transf;
};
int foldin_implicit(const FoldinImplicitArgs &a);
int foldin_refuse(const char *what);
}
