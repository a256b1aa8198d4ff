// Device-level half-sweeps: one call updates every row of one factor matrix given the other one.
// They are what the reference's optimizeA (src/common.c:2742, sparse "Case 4" :3209-3302) and
// optimizeA_implicit (src/common.c:3305-3421) do with an OpenMP loop over rows.
//
// Device layout of a factor matrix F with kk latent coordinates: row-major [rows x ld], ld = cmf_ld_for(kk)
// (whole 128-byte lines); columns 0..kk-1 hold the coordinates, the rest is zero padding.  The biases of a side
// live in a separate array indexed by row.
#pragma once
#include <cuda_runtime.h>
#include "cmf_types.h"

namespace cmfb200 {

struct CsrView {            // device pointers
    const size_t *ptr;      // [rows+1]
    const int_t *idx;       // [nnz]   column = row index into the opposing factor
    const real_t *val;      // [nnz]
};

struct SweepPlan {          // device pointers, built once per CSR by build_sweep_plan()
    const int_t *order;     // rows sorted by decreasing degree (rows without entries excluded)
    int_t n_rows;           // length of order
    int_t n_long;           // the first n_long rows of order get one whole thread block each
    int_t n_huge;           // the first n_huge (<= n_long) rows get a whole cluster of thread blocks each
    const int_t *host_deg;  // HOST copy of the stored-entry counts of `order` (descending); null = not available
};

struct CgSweepParams {
    real_t *F; int ldF;           // factor being solved (in/out: CG is warm-started)
    const real_t *G; int ldG;     // opposing factor
    real_t *Fbias;                // biases of the solved side [rows] (in/out), used when solve_bias
    const real_t *Gbias;          // biases of the opposing side, used when center_opp
    int kk;                       // shared latent coordinates
    CsrView X;
    SweepPlan plan;
    real_t lam, lam_last;
    bool scale_lam, scale_bias_const;
    bool solve_bias;              // the solved row has a bias coordinate (Fbias; its opposing value is 1)
    bool center_opp;              // subtract the opposing row's bias (Gbias) from every x before use
    bool bias_start_one;          // start the bias coordinate from 1.0 instead of the stored bias
    bool last_coord_special;      // exact solves only: the LAST coordinate is regularised by lam_last even without a bias
                                  // (the reference's fold-in without user bias, src/collective.c:3780-3796 -> common.c:718-721)
    int max_cg_steps;
    const real_t *gram;           // constant [kk x kk] matrix (row-major, full symmetric) added to every row's system:
                                  // implicit model: G^T G; explicit model with side info / implicit features: Q
    const real_t *qvec; int ldq;  // explicit + side info: per-row vector [rows x ldq] added to the right-hand side
    bool solve_all_rows;          // explicit + side info: rows without stored entries are solved too
    bool values_positive;         // implicit model: every stored value is > 0 (lets the tensor-core sweep use sqrt(x) weights)
    void *debug;                  // developer builds (-DNM_TIMING): device buffer for per-role cycle counts, else null
    cudaStream_t side_stream;     // optional second stream for the long-row kernel (caller orders it around the sweep)
    // CG sweeps, optional: the n_hot most gathered opposing rows (hot_rows, device) are kept in a shared-memory table by
    // every thread block; hot_idx = X.idx with the table slot + 1 packed into bits 20..30 (needs fewer than 2^20 opposing
    // rows and n_hot <= 2047).  Null: every gather goes to L2.
    const int_t *hot_idx;
    const int_t *hot_rows;
    int n_hot;
};

// returns 0, or 2 when kk is outside the supported range
int launch_explicit_cg_sweep(const CgSweepParams &p, cudaStream_t stream);
int launch_implicit_cg_sweep(const CgSweepParams &p, cudaStream_t stream);

// Same sweeps with every row's gathered block resident in shared memory across the CG passes, teams of 1-8 warps
// or clusters of 2-8 thread blocks per row (sweep_cg_resident.cu).  Return 3 when the shape is not covered.
// *n_launches (optional) is incremented by the number of kernels launched.
int launch_explicit_cg_sweep_resident(const CgSweepParams &p, cudaStream_t stream, int *n_launches);
int launch_implicit_cg_sweep_resident(const CgSweepParams &p, cudaStream_t stream, int *n_launches);

// Exact per-row solves (normal equations + Cholesky); same parameter block, `max_cg_steps` and
// `bias_start_one` only matter for rows without entries.  reference: factors_closed_form sparse branch
// src/common.c:978-1013 + 1058-1070, factors_implicit_chol src/common.c:2063-2126.
int launch_explicit_chol_sweep(const CgSweepParams &p, cudaStream_t stream);
int launch_implicit_chol_sweep(const CgSweepParams &p, cudaStream_t stream);
// the same with every row's normal matrix built on the tensor cores (sweep_nm.cu; fp32, padded widths 64 / 128);
// return 3 when the shape is not covered (nothing launched).  The launchers above try these first (CMFB200_NM=0: never).
int launch_explicit_chol_sweep_nm(const CgSweepParams &p, cudaStream_t stream);
int launch_implicit_chol_sweep_nm(const CgSweepParams &p, cudaStream_t stream);
// the same on the FP64 tensor cores (sweep_chol_dmma.cu; fp64 library, k + bias + 1 <= 136); 3 = not covered
int launch_explicit_chol_sweep_dmma(const CgSweepParams &p, cudaStream_t stream);
int launch_implicit_chol_sweep_dmma(const CgSweepParams &p, cudaStream_t stream);
// explicit / collective model, k <= 64: the reference's truncated CG run on the tensor-core-built normal matrix (one
// gather per stored entry instead of one per CG pass); 3 = not covered
int launch_explicit_cg_sweep_nm(const CgSweepParams &p, cudaStream_t stream);

// gram[kk x kk] = G[:, :kk]^T G[:, :kk] over `rows` rows (full symmetric storage);
// workspace must hold gram_workspace_elems(kk) elements.
size_t gram_workspace_elems(int kk);
int launch_gram(const real_t *G, int ldG, int_t rows, int kk, real_t *gram, real_t *workspace, cudaStream_t stream);

int max_supported_k();

}  // namespace cmfb200
