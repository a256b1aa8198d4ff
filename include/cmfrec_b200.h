/* cmfrec_b200 -- C ABI of the B200-native ALS solver.
 *
 * Two shared libraries export this same set of symbols, one per floating-point type, exactly like the
 * reference builds one extension module per type (reference setup.py:389-412):
 *     libcmfrec_b200_f64.so   real_t = double   (default)
 *     libcmfrec_b200_f32.so   real_t = float    (compiled with -DUSE_FLOAT)
 * int_t is a 32-bit int, index pointers of compressed matrices are size_t, all dense matrices are row-major
 * (reference include/cmfrec.h.in:201-205, src/cmfrec.h:232-305).
 *
 * PART 1 are drop-in replacements: same names, same argument lists, same return codes as the reference
 * (0 = ok, 1 = out of memory / device failure, 2 = invalid or unsupported input, 3 = interrupted); host pointers in,
 * host pointers out, caller allocates every output.  A build of the reference's Cython shim links against them
 * unchanged.
 * PART 2 exposes the device-resident pieces the fits are made of, for callers that keep data in HBM between
 * calls (multi-GPU drivers, benchmarks, per-function parity tests).  Pointers marked DEVICE are CUDA device
 * pointers on the current device; everything else is host memory.
 */
#ifndef CMFREC_B200_H
#define CMFREC_B200_H

#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef USE_FLOAT
typedef float cmf_real_t;
#else
typedef double cmf_real_t;
#endif
#ifndef CMFREC_B200_NO_SHORT_TYPES
#define real_t cmf_real_t
#define int_t int
#endif

/* ------------------------------------------------------------------------------------------------------------
 * PART 1 -- reference entry points
 * ---------------------------------------------------------------------------------------------------------- */

/* replaces fit_collective_explicit_als, reference src/cmfrec.h:1851-1892 (body src/collective.c:7263-9370) */
int_t fit_collective_explicit_als(
    real_t *biasA, real_t *biasB,
    real_t *A, real_t *B,
    real_t *C, real_t *D,
    real_t *Ai, real_t *Bi,
    bool add_implicit_features,
    bool reset_values, int_t seed,
    real_t *glob_mean,
    real_t *U_colmeans, real_t *I_colmeans,
    int_t m, int_t n, int_t k,
    int_t ixA[], int_t ixB[], real_t *X, size_t nnz,
    real_t *Xfull,
    real_t *weight,
    bool user_bias, bool item_bias, bool center,
    real_t lam, real_t *lam_unique,
    real_t l1_lam, real_t *l1_lam_unique,
    bool scale_lam, bool scale_lam_sideinfo, bool scale_bias_const,
    real_t *scaling_biasA, real_t *scaling_biasB,
    real_t *U, int_t m_u, int_t p,
    real_t *II, int_t n_i, int_t q,
    int_t U_row[], int_t U_col[], real_t *U_sp, size_t nnz_U,
    int_t I_row[], int_t I_col[], real_t *I_sp, size_t nnz_I,
    bool NA_as_zero_X, bool NA_as_zero_U, bool NA_as_zero_I,
    int_t k_main, int_t k_user, int_t k_item,
    real_t w_main, real_t w_user, real_t w_item, real_t w_implicit,
    int_t niter, int nthreads,
    bool verbose, bool handle_interrupt,
    bool use_cg, int_t max_cg_steps, bool precondition_cg, bool finalize_chol,
    bool nonneg, int_t max_cd_steps, bool nonneg_C, bool nonneg_D,
    bool precompute_for_predictions,
    bool include_all_X,
    real_t *B_plus_bias,
    real_t *precomputedBtB,
    real_t *precomputedTransBtBinvBt,
    real_t *precomputedBtXbias,
    real_t *precomputedBeTBeChol,
    real_t *precomputedBiTBi,
    real_t *precomputedTransCtCinvCt,
    real_t *precomputedCtCw,
    real_t *precomputedCtUbias);

/* replaces fit_collective_implicit_als, reference src/cmfrec.h:1893-1921 (body src/collective.c:9375-10207) */
int_t fit_collective_implicit_als(
    real_t *A, real_t *B,
    real_t *C, real_t *D,
    bool reset_values, int_t seed,
    real_t *U_colmeans, real_t *I_colmeans,
    int_t m, int_t n, int_t k,
    int_t ixA[], int_t ixB[], real_t *X, size_t nnz,
    real_t lam, real_t *lam_unique,
    real_t l1_lam, real_t *l1_lam_unique,
    real_t *U, int_t m_u, int_t p,
    real_t *II, int_t n_i, int_t q,
    int_t U_row[], int_t U_col[], real_t *U_sp, size_t nnz_U,
    int_t I_row[], int_t I_col[], real_t *I_sp, size_t nnz_I,
    bool NA_as_zero_U, bool NA_as_zero_I,
    int_t k_main, int_t k_user, int_t k_item,
    real_t w_main, real_t w_user, real_t w_item,
    real_t *w_main_multiplier,
    real_t alpha, bool adjust_weight, bool apply_log_transf,
    int_t niter, int nthreads,
    bool verbose, bool handle_interrupt,
    bool use_cg, int_t max_cg_steps, bool precondition_cg, bool finalize_chol,
    bool nonneg, int_t max_cd_steps, bool nonneg_C, bool nonneg_D,
    bool precompute_for_predictions,
    real_t *precomputedBtB,
    real_t *precomputedBeTBe,
    real_t *precomputedBeTBeChol,
    real_t *precomputedCtUbias);

/* replaces fit_most_popular, reference src/cmfrec.h:1164-1179 (body src/common.c:5371-5699) */
int_t fit_most_popular(
    real_t *biasA, real_t *biasB,
    real_t *glob_mean,
    real_t lam_user, real_t lam_item,
    bool scale_lam, bool scale_bias_const,
    real_t alpha,
    int_t m, int_t n,
    int_t ixA[], int_t ixB[], real_t *X, size_t nnz,
    real_t *Xfull,
    real_t *weight,
    bool implicit, bool adjust_weight, bool apply_log_transf,
    bool nonneg, bool NA_as_zero,
    real_t *w_main_multiplier,
    int nthreads);

/* replaces topN, reference src/cmfrec.h:1152-1163 (body src/common.c:5127-5369) */
int_t topN(
    real_t *a_vec, int_t k_user,
    real_t *B, int_t k_item,
    real_t *biasB,
    real_t glob_mean, real_t biasA,
    int_t k, int_t k_main,
    int_t *include_ix, int_t n_include,
    int_t *exclude_ix, int_t n_exclude,
    int_t *outp_ix, real_t *outp_score,
    int_t n_top, int_t n, int nthreads);

/* replaces predict_multiple, reference src/cmfrec.h (body src/common.c:5066-5112): one dot product + biases + mean per
 * (row, column) pair, NaN where an id is outside [0, m) x [0, n); m == 0 / n == 0: ids are trusted */
int_t predict_multiple(
    real_t *A, int_t k_user,
    real_t *B, int_t k_item,
    real_t *biasA, real_t *biasB,
    real_t glob_mean,
    int_t k, int_t k_main,
    int_t m, int_t n,
    int_t predA[], int_t predB[], size_t nnz,
    real_t *outp,
    int nthreads);

/* replaces predict_X_old_collective_explicit / _implicit, reference src/cmfrec.h:2192-2210 (src/collective.c:11797-11862) */
int_t predict_X_old_collective_explicit(
    int_t row[], int_t col[], real_t *predicted, size_t n_predict,
    real_t *A, real_t *biasA,
    real_t *B, real_t *biasB,
    real_t glob_mean,
    int_t k, int_t k_user, int_t k_item, int_t k_main,
    int_t m, int_t n_max,
    int nthreads);
int_t predict_X_old_collective_implicit(
    int_t row[], int_t col[], real_t *predicted, size_t n_predict,
    real_t *A,
    real_t *B,
    int_t k, int_t k_user, int_t k_item, int_t k_main,
    int_t m, int_t n,
    int nthreads);

/* replaces topN_old_collective_explicit / _implicit, reference src/cmfrec.h:2104-2127 (src/collective.c:11546-11614) */
int_t topN_old_collective_explicit(
    real_t *a_vec, real_t a_bias,
    real_t *A, real_t *biasA, int_t row_index,
    real_t *B,
    real_t *biasB,
    real_t glob_mean,
    int_t k, int_t k_user, int_t k_item, int_t k_main,
    int_t *include_ix, int_t n_include,
    int_t *exclude_ix, int_t n_exclude,
    int_t *outp_ix, real_t *outp_score,
    int_t n_top, int_t n, int_t n_max, bool include_all_X, int nthreads);
int_t topN_old_collective_implicit(
    real_t *a_vec,
    real_t *A, int_t row_index,
    real_t *B,
    int_t k, int_t k_user, int_t k_item, int_t k_main,
    int_t *include_ix, int_t n_include,
    int_t *exclude_ix, int_t n_exclude,
    int_t *outp_ix, real_t *outp_score,
    int_t n_top, int_t n, int nthreads);

/* replaces factors_collective_explicit_multiple, reference src/cmfrec.h:2012-2047 (body src/collective.c:10865-11174): factors
 * (and biases) of NEW rows given the fitted item factors.  GPU path: sparse X of the new rows (COO or CSR) without side
 * information -- the exact (Cholesky) half-sweep of the fit with B fixed; any other argument combination returns 2. */
int_t factors_collective_explicit_multiple(
    real_t *A, real_t *biasA, int_t m,
    real_t *U, int_t m_u, int_t p,
    bool NA_as_zero_U, bool NA_as_zero_X,
    bool nonneg,
    int_t U_row[], int_t U_col[], real_t *U_sp, size_t nnz_U,
    size_t U_csr_p[], int_t U_csr_i[], real_t *U_csr,
    real_t *Ub, int_t m_ubin, int_t pbin,
    real_t *C, real_t *Cb,
    real_t glob_mean, real_t *biasB,
    real_t *U_colmeans,
    real_t *X, int_t ixA[], int_t ixB[], size_t nnz,
    size_t *Xcsr_p, int_t *Xcsr_i, real_t *Xcsr,
    real_t *Xfull, int_t n,
    real_t *weight,
    real_t *B,
    real_t *Bi, bool add_implicit_features,
    int_t k, int_t k_user, int_t k_item, int_t k_main,
    real_t lam, real_t *lam_unique,
    real_t l1_lam, real_t *l1_lam_unique,
    bool scale_lam, bool scale_lam_sideinfo,
    bool scale_bias_const, real_t scaling_biasA,
    real_t w_main, real_t w_user, real_t w_implicit,
    int_t n_max, bool include_all_X,
    real_t *BtB,
    real_t *TransBtBinvBt,
    real_t *BtXbias,
    real_t *BeTBeChol,
    real_t *BiTBi,
    real_t *TransCtCinvCt,
    real_t *CtCw,
    real_t *CtUbias,
    real_t *B_plus_bias,
    int nthreads);

/* replaces factors_collective_implicit_multiple, reference src/cmfrec.h:2048-2070 (body src/collective.c:11176-11330) */
int_t factors_collective_implicit_multiple(
    real_t *A, int_t m,
    real_t *U, int_t m_u, int_t p,
    bool NA_as_zero_U,
    bool nonneg,
    int_t U_row[], int_t U_col[], real_t *U_sp, size_t nnz_U,
    size_t U_csr_p[], int_t U_csr_i[], real_t *U_csr,
    real_t *X, int_t ixA[], int_t ixB[], size_t nnz,
    size_t *Xcsr_p, int_t *Xcsr_i, real_t *Xcsr,
    real_t *B, int_t n,
    real_t *C,
    real_t *U_colmeans,
    int_t k, int_t k_user, int_t k_item, int_t k_main,
    real_t lam, real_t l1_lam, real_t alpha, real_t w_main, real_t w_user,
    real_t w_main_multiplier,
    bool apply_log_transf,
    real_t *BeTBe,
    real_t *BtB,
    real_t *BeTBeChol,
    real_t *CtUbias,
    int nthreads);

/* replaces precompute_collective_explicit, reference src/cmfrec.h:1922-1945 (body src/collective.c:10209-10485) */
int_t precompute_collective_explicit(
    real_t *B, int_t n, int_t n_max, bool include_all_X,
    real_t *C, int_t p,
    real_t *Bi, bool add_implicit_features,
    real_t *biasB, real_t glob_mean, bool NA_as_zero_X,
    real_t *U_colmeans, bool NA_as_zero_U,
    int_t k, int_t k_user, int_t k_item, int_t k_main,
    bool user_bias,
    bool nonneg,
    real_t lam, real_t *lam_unique,
    bool scale_lam, bool scale_lam_sideinfo,
    bool scale_bias_const, real_t scaling_biasA,
    real_t w_main, real_t w_user, real_t w_implicit,
    real_t *B_plus_bias,
    real_t *BtB,
    real_t *TransBtBinvBt,
    real_t *BtXbias,
    real_t *BeTBeChol,
    real_t *BiTBi,
    real_t *TransCtCinvCt,
    real_t *CtCw,
    real_t *CtUbias);

/* replaces precompute_collective_implicit, reference src/cmfrec.h:1946-1958 (body src/collective.c:10487-10573) */
int_t precompute_collective_implicit(
    real_t *B, int_t n,
    real_t *C, int_t p,
    real_t *U_colmeans, bool NA_as_zero_U,
    int_t k, int_t k_user, int_t k_item, int_t k_main,
    real_t lam, real_t w_main, real_t w_user, real_t w_main_multiplier,
    bool nonneg,
    bool extra_precision,
    real_t *BtB,
    real_t *BeTBe,
    real_t *BeTBeChol,
    real_t *CtUbias);

/* replaces get_has_openmp, reference src/cmfrec.h:646 (helpers.c:1817) */
bool get_has_openmp(void);

/* ------------------------------------------------------------------------------------------------------------
 * PART 2 -- building blocks (cmfb200_ prefix)
 * ---------------------------------------------------------------------------------------------------------- */

/* "f32" or "f64": which real_t this library was built for */
const char *cmfb200_real_name(void);
/* number of CUDA devices visible (0 = none: every compute entry point then fails with code 1) */
int cmfb200_device_count(void);

/* --- host-side preparation, bit-exact counterparts of reference helpers --------------------------------- */
/* random_parallel, reference src/helpers.c:930-1043 (A from the seed state, B from the jumped state) */
void cmfb200_random_init(real_t *A, size_t sizeA, real_t *B, size_t sizeB, int_t seed, bool normal);
/* the same matrices filled by `nthreads` threads (what the fits call): the generator is jumped to the start of every
 * 2^15-draw chunk, the chunks are sampled independently and stitched in order -- values identical to the call above */
void cmfb200_random_init_threads(real_t *A, size_t sizeA, real_t *B, size_t sizeB, int_t seed, bool normal, int nthreads);
/* coo_to_csr_and_csc, reference src/helpers.c:1375-1491 (no weights) */
void cmfb200_coo_to_csr_and_csc(const int_t *Xrow, const int_t *Xcol, const real_t *Xval, int_t m, int_t n, size_t nnz,
                                size_t *csr_p, int_t *csr_i, real_t *csr_v,
                                size_t *csc_p, int_t *csc_i, real_t *csc_v);
/* the mean step of calc_mean_and_center, reference src/common.c:3494-3513 + 3603-3604 (sparse, unweighted) */
real_t cmfb200_global_mean(const real_t *X, size_t nnz, int nthreads);
/* initialize_biases_twosided, reference src/common.c:4410 (sparse branches :4643-4669, :4799-4825) */
void cmfb200_init_biases_twosided(int_t m, int_t n,
                                  const size_t *csr_p, const int_t *csr_i, const real_t *csr_v,
                                  const size_t *csc_p, const int_t *csc_i, const real_t *csc_v,
                                  real_t lam_user, real_t lam_item, bool scale_lam, bool nonneg,
                                  real_t *biasA, real_t *biasB, int nthreads);

/* --- device-resident ALS state ---------------------------------------------------------------------------
 * Holds both orientations of X, both factor matrices and workspaces in HBM.  One state per GPU/process; with
 * world > 1 every rank passes the same matrices, keeps only its own block of rows of each orientation, and the
 * freshly solved blocks are all-gathered over NCCL after every half-sweep.                                   */
/* gram[kk x kk] = G^T G (full symmetric, row-major) for a dense host matrix G [rows x kk] -- the cblas_tsyrk of the
 * implicit half-sweep (reference src/common.c:3328, src/collective.c:6276).  fp32 library: tcgen05 tensor cores with
 * the 3xTF32 split when the padded row width is 64, 128 or 256 floats.  `repeats` > 0 additionally times that many
 * launches with CUDA events (*ms_per_launch, optional). */
int cmfb200_gram(const real_t *G, int_t rows, int kk, real_t *gram, int repeats, float *ms_per_launch);

/* Device buffers are recycled between calls through the device's default memory pool (CMFB200_POOL=0 disables it);
 * this hands the cached memory back to the driver. */
void cmfb200_trim_pool(void);

/* Batched serving from factors kept in HBM (serve.cu).  A [m x (k_user+k+k_main)], B [n x (k_item+k+k_main)] row-major as the
 * fit entry points return them; biases may be NULL.  cmfb200_serve_predict = predict_multiple for many pairs;
 * cmfb200_serve_topn = topN for MANY users in one call: scores of the listed users against all items on the tensor cores
 * (fp32) / DFMA (fp64), the users' seen items (CSR over the listed users: seen_ptr [n_users + 1], seen_idx; NULL = none)
 * excluded, then an exact per-user radix select; out_ix [n_users x n_top], out_score likewise or NULL; n_top <= 2048.
 * Equal scores rank the lower item id first.  ms_device (optional): device time of the call. */
void *cmfb200_serve_create(const real_t *A, int_t m, int_t k_user, const real_t *B, int_t n, int_t k_item, const real_t *biasA,
                           const real_t *biasB, real_t glob_mean, int_t k, int_t k_main, int *rc_out);
void cmfb200_serve_destroy(void *h);
int cmfb200_serve_predict(void *h, const int_t *row, const int_t *col, size_t n_predict, real_t *out);
int cmfb200_serve_topn(void *h, const int_t *users, int_t n_users, const size_t *seen_ptr, const int_t *seen_idx, int_t n_top,
                       int_t *out_ix, real_t *out_score, float *ms_device);
/* C[M x N] = A[M x K] B[N x K]^T on the tcgen05 tensor cores with the 3xTF32 split (gemm_tc.cu; fp32 library, returns 3 in the
 * fp64 one); host buffers; repeats > 0 also times the launch */
int cmfb200_gemm_nt(const real_t *A, int lda, int M, const real_t *B, int ldb, int N, int K, real_t *C, int repeats, float *ms_per_launch);
/* Test aid: fills the shared memory of every SM with `pattern` (e.g. 0x7fc00000 = NaN), so that a kernel launched next
 * that reads shared memory it never wrote produces visibly wrong results instead of depending on what ran before. */
int cmfb200_debug_poison_smem(unsigned pattern);

typedef struct cmfb200_als cmfb200_als;

typedef struct cmfb200_als_options {
    int implicit;              /* 0: explicit-feedback model (optimizeA), 1: implicit (optimizeA_implicit) */
    int_t m, n, k;             /* k = number of latent coordinates solved (reference k + k_main) */
    int user_bias, item_bias;  /* explicit only */
    real_t lam_A, lam_B;       /* regulariser for rows of A / rows of B (after division by w_main) */
    real_t lam_biasA, lam_biasB;
    int scale_lam;
    int max_cg_steps;
    int rank, world;           /* world > 1 needs nccl_id */
    const void *nccl_id;       /* 128 bytes from cmfb200_nccl_unique_id, identical on all ranks; states created with an id
                                * seen before in this process reuse its communicator */
    void *stream;              /* cudaStream_t to enqueue on (NULL = default stream) */
} cmfb200_als_options;

int cmfb200_nccl_unique_id(void *out128);
/* Multi-GPU behind the reference-named entry points (which carry no communicator argument): after this call every
 * fit_collective_explicit_als / fit_collective_implicit_als of the process runs as rank `rank` of `world` ranks, one process
 * per GPU.  Every rank passes the same arguments and receives the full factors; X is ingested on every device, the rows
 * are dealt to the ranks there.  nccl_id128: the 128 bytes of cmfb200_nccl_unique_id from rank 0, identical on all ranks.
 * world <= 1 returns to single-GPU operation. */
int cmfb200_set_world(int rank, int world, const void *nccl_id128);

/* How rows are dealt to ranks: row r of the caller's numbering lives at device row to_device_row[r]; rank q owns
 * device rows [q*block, (q+1)*block).  Rows are dealt round-robin in order of decreasing number of stored entries
 * (equal block sizes, near-equal entry counts); identity when world == 1.  Pure host function. */
int cmfb200_partition_rows(const size_t *indptr, int_t rows, int world, int_t *to_device_row, int_t *block);

/* X given as CSR and CSC with identical entries (values already centred / scaled the way the model wants) */
int cmfb200_als_create(cmfb200_als **out, const cmfb200_als_options *opt,
                       const size_t *csr_p, const int_t *csr_i, const real_t *csr_v,
                       const size_t *csc_p, const int_t *csc_i, const real_t *csc_v);
/* the same from COO triplets that are ALREADY ON THE DEVICE (device pointers; d_X is transformed in place to (x - mu) * scale): both
 * orientations are built -- and with world > 1 dealt to the ranks -- on the GPU.  nnz <= INT_MAX. */
int cmfb200_als_create_from_device_coo(cmfb200_als **out, const cmfb200_als_options *opt, const int_t *d_ixA, const int_t *d_ixB,
                                       real_t *d_X, size_t nnz, real_t mu, real_t scale);
/* A ~ U(0, scale) from a counter-based hash of (seed, element) on the device (identical on every rank), B and biases zero */
int cmfb200_als_random_factors(cmfb200_als *s, unsigned long long seed, real_t scale);
void cmfb200_als_destroy(cmfb200_als *s);
/* factors in caller numbering, row-major [m x k] / [n x k]; bias arrays may be NULL */
int cmfb200_als_set_factors(cmfb200_als *s, const real_t *A, const real_t *biasA, const real_t *B, const real_t *biasB);
int cmfb200_als_get_factors(cmfb200_als *s, real_t *A, real_t *biasA, real_t *B, real_t *biasB);
/* one half-sweep: which = 0 updates B from A (reference optimizeA on the CSC), 1 updates A from B.
 * `iter` is the 0-based ALS iteration (decides the bias warm start, reference src/collective.c:8538-8545).
 * solver: 0 = conjugate gradient, 1 = Cholesky.  Enqueues on the state's stream; includes the all-gather. */
int cmfb200_als_half_sweep(cmfb200_als *s, int which, int iter, int solver);
/* n_iters full iterations (B then A), the last one of `niter_total` switched to Cholesky when finalize_chol */
int cmfb200_als_iterate(cmfb200_als *s, int first_iter, int n_iters, int niter_total, int use_cg, int finalize_chol);
/* same, bracketed by CUDA events recorded on the state's stream; *elapsed_ms = device time of the n_iters */
int cmfb200_als_timed_iterate(cmfb200_als *s, int first_iter, int n_iters, int niter_total, int use_cg,
                              int finalize_chol, float *elapsed_ms);
/* Turn an explicit-feedback state (single GPU) into the model with dense side information and/or implicit features:
 * U_centred [m x p], I_centred [n x q] host matrices without missing values (NULL = absent), weights and regularisers
 * as the reference uses them after dividing by w_main (lam_C = lam/w_user * (scale_lam ? m : 1), ...; reference
 * src/collective.c:8358-8530).  cmfb200_als_iterate then runs the C, D, Bi, Ai, B, A order of src/collective.c:8342-8876. */
int cmfb200_als_attach_collective(cmfb200_als *s, const real_t *U_centred, int p, const real_t *I_centred, int q,
                                  int add_implicit_features, real_t w_user, real_t w_item, real_t w_implicit, real_t lam_C,
                                  real_t lam_D, real_t lam_Bi, real_t lam_Ai);
/* C [p x k], D [q x k], Ai [m x k], Bi [n x k] (NULL to skip) */
int cmfb200_als_get_collective(cmfb200_als *s, real_t *C, real_t *D, real_t *Ai, real_t *Bi);
/* per-launch timing of the row-solve kernel: when on, every half-sweep brackets its solve kernel with CUDA events
 * on the state's stream; read_profile synchronises, returns the summed kernel time and launch count for
 * which = 0 (B sweeps) / 1 (A sweeps) since the last read, and clears the record. */
void cmfb200_als_set_profile(cmfb200_als *s, int on);
int cmfb200_als_read_profile(cmfb200_als *s, int which, double *total_ms, long long *count);
int cmfb200_als_sync(cmfb200_als *s);
/* kernels launched by this state so far */
long long cmfb200_als_launch_count(const cmfb200_als *s);
/* stored entries / rows held by this rank (for throughput accounting) */
void cmfb200_als_local_counts(const cmfb200_als *s, size_t *nnz_rows_A, size_t *nnz_rows_B, int_t *rows_A, int_t *rows_B);

#ifndef CMFREC_B200_NO_SHORT_TYPES
#undef real_t
#undef int_t
#endif

#ifdef __cplusplus
}
#endif
#endif /* CMFREC_B200_H */
