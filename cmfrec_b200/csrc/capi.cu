// extern "C" surface of libcmfrec_b200_{f32,f64}.so -- see include/cmfrec_b200.h.
// Pure forwarding: argument lists are the reference's (src/cmfrec.h), bodies live in fit.cu, als.cu,
// host_prep.cpp, popular.cpp and topn.cu.
#include "../../include/cmfrec_b200.h"
#include <cstdio>
#include <cstring>
#include "als.h"
#include "collective.h"
#include "fit.h"
#include "host_prep.h"
#include "nccl_link.h"
#include "popular.h"
#include "topn.h"
#include "serve.h"
#include "foldin.h"
#include "postfit.h"
#include "gemm_tc.h"
#include <cmath>

using namespace cmfb200;

struct cmfb200_als {
    AlsState st;
};

extern "C" {

int fit_collective_explicit_als(
    real_t *biasA, real_t *biasB, real_t *A, real_t *B, real_t *C, real_t *D, real_t *Ai, real_t *Bi,
    bool add_implicit_features, bool reset_values, int_t seed, real_t *glob_mean, real_t *U_colmeans,
    real_t *I_colmeans, int_t m, int_t n, int_t k, int_t ixA[], int_t ixB[], real_t *X, size_t nnz, real_t *Xfull,
    real_t *weight, bool user_bias, bool item_bias, bool center, real_t lam, real_t *lam_unique, real_t l1_lam,
    real_t *l1_lam_unique, bool scale_lam, bool scale_lam_sideinfo, bool scale_bias_const, real_t *scaling_biasA,
    real_t *scaling_biasB, real_t *U, int_t m_u, int_t p, real_t *II, int_t n_i, int_t q, int_t U_row[], int_t U_col[],
    real_t *U_sp, size_t nnz_U, int_t I_row[], int_t I_col[], real_t *I_sp, size_t nnz_I, bool NA_as_zero_X,
    bool NA_as_zero_U, bool NA_as_zero_I, int_t k_main, int_t k_user, int_t k_item, real_t w_main, real_t w_user,
    real_t w_item, real_t w_implicit, int_t niter, int nthreads, bool verbose, bool handle_interrupt, bool use_cg,
    int_t max_cg_steps, bool precondition_cg, bool finalize_chol, bool nonneg, int_t max_cd_steps, bool nonneg_C,
    bool nonneg_D, bool precompute_for_predictions, bool include_all_X, real_t *B_plus_bias, real_t *precomputedBtB,
    real_t *precomputedTransBtBinvBt, real_t *precomputedBtXbias, real_t *precomputedBeTBeChol,
    real_t *precomputedBiTBi, real_t *precomputedTransCtCinvCt, real_t *precomputedCtCw, real_t *precomputedCtUbias)
{
    ExplicitArgs a{biasA, biasB, A, B, C, D, Ai, Bi, add_implicit_features, reset_values, seed, glob_mean, U_colmeans,
                   I_colmeans, m, n, k, ixA, ixB, X, nnz, Xfull, weight, user_bias, item_bias, center, lam, lam_unique,
                   l1_lam, l1_lam_unique, scale_lam, scale_lam_sideinfo, scale_bias_const, scaling_biasA,
                   scaling_biasB, U, m_u, p, II, n_i, q, U_row, U_col, U_sp, nnz_U, I_row, I_col, I_sp, nnz_I,
                   NA_as_zero_X, NA_as_zero_U, NA_as_zero_I, k_main, k_user, k_item, w_main, w_user, w_item,
                   w_implicit, niter, nthreads, verbose, handle_interrupt, use_cg, max_cg_steps, precondition_cg,
                   finalize_chol, nonneg, max_cd_steps, nonneg_C, nonneg_D, precompute_for_predictions, include_all_X,
                   B_plus_bias, precomputedBtB, precomputedTransBtBinvBt, precomputedBtXbias, precomputedBeTBeChol,
                   precomputedBiTBi, precomputedTransCtCinvCt, precomputedCtCw, precomputedCtUbias};
    return fit_explicit(a);
}

int fit_collective_implicit_als(
    real_t *A, real_t *B, real_t *C, real_t *D, bool reset_values, int_t seed, real_t *U_colmeans, real_t *I_colmeans,
    int_t m, int_t n, int_t k, int_t ixA[], int_t ixB[], real_t *X, size_t nnz, real_t lam, real_t *lam_unique,
    real_t l1_lam, real_t *l1_lam_unique, real_t *U, int_t m_u, int_t p, real_t *II, int_t n_i, int_t q, int_t U_row[],
    int_t U_col[], real_t *U_sp, size_t nnz_U, int_t I_row[], int_t I_col[], real_t *I_sp, size_t nnz_I,
    bool NA_as_zero_U, bool NA_as_zero_I, int_t k_main, int_t k_user, int_t k_item, real_t w_main, real_t w_user,
    real_t w_item, real_t *w_main_multiplier, real_t alpha, bool adjust_weight, bool apply_log_transf, int_t niter,
    int nthreads, bool verbose, bool handle_interrupt, bool use_cg, int_t max_cg_steps, bool precondition_cg,
    bool finalize_chol, bool nonneg, int_t max_cd_steps, bool nonneg_C, bool nonneg_D, bool precompute_for_predictions,
    real_t *precomputedBtB, real_t *precomputedBeTBe, real_t *precomputedBeTBeChol, real_t *precomputedCtUbias)
{
    ImplicitArgs a{A, B, C, D, reset_values, seed, U_colmeans, I_colmeans, m, n, k, ixA, ixB, X, nnz, lam, lam_unique,
                   l1_lam, l1_lam_unique, U, m_u, p, II, n_i, q, U_row, U_col, U_sp, nnz_U, I_row, I_col, I_sp, nnz_I,
                   NA_as_zero_U, NA_as_zero_I, k_main, k_user, k_item, w_main, w_user, w_item, w_main_multiplier,
                   alpha, adjust_weight, apply_log_transf, niter, nthreads, verbose, handle_interrupt, use_cg,
                   max_cg_steps, precondition_cg, finalize_chol, nonneg, max_cd_steps, nonneg_C, nonneg_D,
                   precompute_for_predictions, precomputedBtB, precomputedBeTBe, precomputedBeTBeChol,
                   precomputedCtUbias};
    return fit_implicit(a);
}

int fit_most_popular(real_t *biasA, real_t *biasB, real_t *glob_mean, real_t lam_user, real_t lam_item, bool scale_lam,
                     bool scale_bias_const, real_t alpha, int_t m, int_t n, int_t ixA[], int_t ixB[], real_t *X,
                     size_t nnz, real_t *Xfull, real_t *weight, bool implicit, bool adjust_weight,
                     bool apply_log_transf, bool nonneg, bool NA_as_zero, real_t *w_main_multiplier, int nthreads)
{
    return most_popular(biasA, biasB, glob_mean, lam_user, lam_item, scale_lam, scale_bias_const, alpha, m, n, ixA, ixB,
                        X, nnz, Xfull, weight, implicit, adjust_weight, apply_log_transf, nonneg, NA_as_zero,
                        w_main_multiplier, nthreads);
}

int topN(real_t *a_vec, int_t k_user, real_t *B, int_t k_item, real_t *biasB, real_t glob_mean, real_t biasA, int_t k,
         int_t k_main, int_t *include_ix, int_t n_include, int_t *exclude_ix, int_t n_exclude, int_t *outp_ix,
         real_t *outp_score, int_t n_top, int_t n, int nthreads)
{
    return top_n(a_vec, k_user, B, k_item, biasB, glob_mean, biasA, k, k_main, include_ix, n_include, exclude_ix,
                 n_exclude, outp_ix, outp_score, n_top, n, nthreads);
}

int predict_multiple(real_t *A, int_t k_user, real_t *B, int_t k_item, real_t *biasA, real_t *biasB, real_t glob_mean, int_t k,
                     int_t k_main, int_t m, int_t n, int_t predA[], int_t predB[], size_t nnz, real_t *outp, int nthreads)
{
    (void)nthreads;
    return predict_multiple_host(A, k_user, B, k_item, biasA, biasB, glob_mean, k, k_main, m, n, predA, predB, nnz, outp);
}

// reference src/collective.c:11797-11835: unknown ids fall back to mean + whatever bias is known
int predict_X_old_collective_explicit(int_t row[], int_t col[], real_t *predicted, size_t n_predict, real_t *A, real_t *biasA,
                                      real_t *B, real_t *biasB, real_t glob_mean, int_t k, int_t k_user, int_t k_item, int_t k_main,
                                      int_t m, int_t n_max, int nthreads)
{
    (void)nthreads;
    const int rc = predict_multiple_host(A, k_user, B, k_item, biasA, biasB, glob_mean, k, k_main, m, n_max, row, col, n_predict, predicted);
    if (rc) return rc;
    for (size_t ix = 0; ix < n_predict; ix++)
        if (std::isnan(predicted[ix]))
            predicted[ix] = glob_mean + ((biasA != nullptr && row[ix] < m) ? biasA[row[ix]] : real_t(0)) +
                            ((biasB != nullptr && col[ix] < n_max) ? biasB[col[ix]] : real_t(0));
    return 0;
}

// reference src/collective.c:11837-11862
int predict_X_old_collective_implicit(int_t row[], int_t col[], real_t *predicted, size_t n_predict, real_t *A, real_t *B, int_t k,
                                      int_t k_user, int_t k_item, int_t k_main, int_t m, int_t n, int nthreads)
{
    (void)nthreads;
    return predict_multiple_host(A, k_user, B, k_item, nullptr, nullptr, real_t(0), k, k_main, m, n, row, col, n_predict, predicted);
}

// reference src/collective.c:11546-11587
int topN_old_collective_explicit(real_t *a_vec, real_t a_bias, real_t *A, real_t *biasA, int_t row_index, real_t *B, real_t *biasB,
                                 real_t glob_mean, int_t k, int_t k_user, int_t k_item, int_t k_main, int_t *include_ix, int_t n_include,
                                 int_t *exclude_ix, int_t n_exclude, int_t *outp_ix, real_t *outp_score, int_t n_top, int_t n, int_t n_max,
                                 bool include_all_X, int nthreads)
{
    if (include_all_X || n == 0) n = n_max;
    if (a_vec != nullptr)
        return top_n(a_vec, k_user, B, k_item, biasB, glob_mean, a_bias, k, k_main, include_ix, n_include, exclude_ix, n_exclude, outp_ix,
                     outp_score, n_top, n, nthreads);
    return top_n(A + (size_t)row_index * (size_t)(k_user + k + k_main), k_user, B, k_item, biasB, glob_mean,
                 biasA == nullptr ? real_t(0) : biasA[row_index], k, k_main, include_ix, n_include, exclude_ix, n_exclude, outp_ix, outp_score,
                 n_top, n, nthreads);
}

// reference src/collective.c:11589-11614
int topN_old_collective_implicit(real_t *a_vec, real_t *A, int_t row_index, real_t *B, int_t k, int_t k_user, int_t k_item, int_t k_main,
                                 int_t *include_ix, int_t n_include, int_t *exclude_ix, int_t n_exclude, int_t *outp_ix, real_t *outp_score,
                                 int_t n_top, int_t n, int nthreads)
{
    return topN_old_collective_explicit(a_vec, real_t(0), A, nullptr, row_index, B, nullptr, real_t(0), k, k_user, k_item, k_main, include_ix,
                                        n_include, exclude_ix, n_exclude, outp_ix, outp_score, n_top, n, n, false, nthreads);
}

// reference src/collective.c:10865-11174.  Covered on the GPU: sparse X (COO or CSR) of new rows, no side information
int factors_collective_explicit_multiple(
    real_t *A, real_t *biasA, int_t m, real_t *U, int_t m_u, int_t p, bool NA_as_zero_U, bool NA_as_zero_X, bool nonneg,
    int_t U_row[], int_t U_col[], real_t *U_sp, size_t nnz_U, size_t U_csr_p[], int_t U_csr_i[], real_t *U_csr, real_t *Ub,
    int_t m_ubin, int_t pbin, real_t *C, real_t *Cb, real_t glob_mean, real_t *biasB, real_t *U_colmeans, real_t *X, int_t ixA[],
    int_t ixB[], size_t nnz, size_t *Xcsr_p, int_t *Xcsr_i, real_t *Xcsr, real_t *Xfull, int_t n, real_t *weight, real_t *B,
    real_t *Bi, bool add_implicit_features, int_t k, int_t k_user, int_t k_item, int_t k_main, real_t lam, real_t *lam_unique,
    real_t l1_lam, real_t *l1_lam_unique, bool scale_lam, bool scale_lam_sideinfo, bool scale_bias_const, real_t scaling_biasA,
    real_t w_main, real_t w_user, real_t w_implicit, int_t n_max, bool include_all_X, real_t *BtB, real_t *TransBtBinvBt,
    real_t *BtXbias, real_t *BeTBeChol, real_t *BiTBi, real_t *TransCtCinvCt, real_t *CtCw, real_t *CtUbias, real_t *B_plus_bias,
    int nthreads)
{
    (void)m_u; (void)p; (void)m_ubin; (void)pbin; (void)C; (void)Cb; (void)U_colmeans; (void)Bi; (void)w_user; (void)w_implicit;
    (void)BtB; (void)TransBtBinvBt; (void)BtXbias; (void)BeTBeChol; (void)BiTBi; (void)TransCtCinvCt; (void)CtCw; (void)CtUbias;
    (void)B_plus_bias; (void)nthreads; (void)U_row; (void)U_col; (void)U_csr_i;
    if (U || U_sp || U_csr_p || U_csr || nnz_U || Ub) return foldin_refuse("side information for new rows");
    if (NA_as_zero_U || NA_as_zero_X) return foldin_refuse("NA_as_zero for new rows");
    if (Xfull) return foldin_refuse("dense X for new rows");
    if (weight) return foldin_refuse("observation weights for new rows");
    if (add_implicit_features) return foldin_refuse("implicit features for new rows");
    if (nonneg || l1_lam != 0 || l1_lam_unique) return foldin_refuse("non-negativity / L1 for new rows");
    if (k_user || k_item) return foldin_refuse("k_user / k_item for new rows");
    FoldinExplicitArgs a{A, biasA, m, ixA, ixB, X, nnz, Xcsr_p, Xcsr_i, Xcsr, B, biasB, n, n_max, include_all_X, glob_mean, k, k_main,
                         lam, lam_unique, scale_lam, scale_lam_sideinfo, scale_bias_const, scaling_biasA, w_main};
    return foldin_explicit(a);
}

// reference src/collective.c:11176-11330
int factors_collective_implicit_multiple(
    real_t *A, int_t m, real_t *U, int_t m_u, int_t p, bool NA_as_zero_U, bool nonneg, int_t U_row[], int_t U_col[], real_t *U_sp,
    size_t nnz_U, size_t U_csr_p[], int_t U_csr_i[], real_t *U_csr, real_t *X, int_t ixA[], int_t ixB[], size_t nnz, size_t *Xcsr_p,
    int_t *Xcsr_i, real_t *Xcsr, real_t *B, int_t n, real_t *C, real_t *U_colmeans, int_t k, int_t k_user, int_t k_item, int_t k_main,
    real_t lam, real_t l1_lam, real_t alpha, real_t w_main, real_t w_user, real_t w_main_multiplier, bool apply_log_transf,
    real_t *BeTBe, real_t *BtB, real_t *BeTBeChol, real_t *CtUbias, int nthreads)
{
    (void)m_u; (void)p; (void)C; (void)U_colmeans; (void)w_user; (void)BeTBe; (void)BtB; (void)BeTBeChol; (void)CtUbias; (void)nthreads;
    (void)U_row; (void)U_col; (void)U_csr_i;
    if (U || U_sp || U_csr_p || U_csr || nnz_U) return foldin_refuse("side information for new rows");
    if (NA_as_zero_U) return foldin_refuse("NA_as_zero for new rows");
    if (nonneg || l1_lam != 0) return foldin_refuse("non-negativity / L1 for new rows");
    if (k_user || k_item) return foldin_refuse("k_user / k_item for new rows");
    FoldinImplicitArgs a{A, m, ixA, ixB, X, nnz, Xcsr_p, Xcsr_i, Xcsr, B, n, k, k_main, lam, alpha, w_main, w_main_multiplier, apply_log_transf};
    return foldin_implicit(a);
}

// reference src/collective.c:10209-10485: the matrices kept for predictions on new rows
int precompute_collective_explicit(
    real_t *B, int_t n, int_t n_max, bool include_all_X, real_t *C, int_t p, real_t *Bi, bool add_implicit_features, real_t *biasB,
    real_t glob_mean, bool NA_as_zero_X, real_t *U_colmeans, bool NA_as_zero_U, int_t k, int_t k_user, int_t k_item, int_t k_main,
    bool user_bias, bool nonneg, real_t lam, real_t *lam_unique, bool scale_lam, bool scale_lam_sideinfo, bool scale_bias_const,
    real_t scaling_biasA, real_t w_main, real_t w_user, real_t w_implicit, real_t *B_plus_bias, real_t *BtB, real_t *TransBtBinvBt,
    real_t *BtXbias, real_t *BeTBeChol, real_t *BiTBi, real_t *TransCtCinvCt, real_t *CtCw, real_t *CtUbias)
{
    (void)glob_mean; (void)U_colmeans; (void)scaling_biasA; (void)BtXbias; (void)CtUbias; (void)nonneg;
    if (NA_as_zero_X || NA_as_zero_U) return foldin_refuse("NA_as_zero in precompute_collective_explicit");
    if (k_user || k_item) return foldin_refuse("k_user / k_item in precompute_collective_explicit");
    if ((scale_lam || scale_lam_sideinfo) && scale_bias_const && user_bias) return foldin_refuse("scale_bias_const");
    if ((C || add_implicit_features) && k_main) return foldin_refuse("k_main together with side information / implicit features");
    if (!B || n < 1) return 2;
    if (n_max == 0) n_max = n;
    if (include_all_X) n = n_max;
    real_t lam_main = lam_unique ? lam_unique[2] : lam, lam_bias = lam_unique ? lam_unique[user_bias ? 0 : 2] : lam;
    if (w_main != 1) {
        lam_main /= w_main; lam_bias /= w_main; w_user /= w_main; w_implicit /= w_main;
    }
    PostfitExplicit pf;
    pf.B = B; pf.biasB = biasB; pf.n = n; pf.kk = k + k_main;
    pf.user_bias = user_bias; pf.item_bias = biasB != nullptr;
    pf.lam = lam_main; pf.lam_bias = lam_bias; pf.scale_lam = scale_lam || scale_lam_sideinfo;
    pf.B_plus_bias = B_plus_bias; pf.BtB = BtB; pf.TransBtBinvBt = TransBtBinvBt;
    if (C || add_implicit_features) {
        pf.C = C; pf.p = C ? p : 0; pf.w_user = w_user;
        pf.Bi = add_implicit_features ? Bi : nullptr; pf.implicit_features = add_implicit_features; pf.w_implicit = w_implicit;
        pf.scale_lam_sideinfo = scale_lam_sideinfo;
        pf.BiTBi = BiTBi; pf.TransCtCinvCt = TransCtCinvCt; pf.CtCw = CtCw; pf.BeTBeChol = BeTBeChol;
    }
    return postfit_explicit(pf);
}

// reference src/collective.c:10487-10573
int precompute_collective_implicit(real_t *B, int_t n, real_t *C, int_t p, real_t *U_colmeans, bool NA_as_zero_U, int_t k, int_t k_user,
                                   int_t k_item, int_t k_main, real_t lam, real_t w_main, real_t w_user, real_t w_main_multiplier,
                                   bool nonneg, bool extra_precision, real_t *BtB, real_t *BeTBe, real_t *BeTBeChol, real_t *CtUbias)
{
    (void)U_colmeans; (void)nonneg; (void)extra_precision; (void)CtUbias;
    if (NA_as_zero_U) return foldin_refuse("NA_as_zero_U in precompute_collective_implicit");
    if (k_user || k_item) return foldin_refuse("k_user / k_item in precompute_collective_implicit");
    if (!B || n < 1 || !BtB) return 2;
    if (w_main_multiplier != 1) w_main *= w_main_multiplier;
    if (w_main != 1) {
        lam /= w_main;
        w_user /= w_main;
    }
    return postfit_implicit(B, n, k + k_main, lam, BtB, p ? C : nullptr, p, w_user, BeTBe, BeTBeChol, false);
}

// ---- batched serving from factors resident in HBM (serve.cu)
void *cmfb200_serve_create(const real_t *A, int_t m, int_t k_user, const real_t *B, int_t n, int_t k_item, const real_t *biasA,
                           const real_t *biasB, real_t glob_mean, int_t k, int_t k_main, int *rc_out)
{
    int rc = 0;
    cmfb200::ServeState *s = cmfb200::serve_create(A, m, k_user, B, n, k_item, biasA, biasB, glob_mean, k, k_main, &rc);
    if (rc_out) *rc_out = rc;
    return s;
}
void cmfb200_serve_destroy(void *h) { cmfb200::serve_destroy(static_cast<cmfb200::ServeState *>(h)); }
int cmfb200_serve_predict(void *h, const int_t *row, const int_t *col, size_t n_predict, real_t *out)
{
    return h ? cmfb200::serve_predict(static_cast<cmfb200::ServeState *>(h), row, col, n_predict, out) : 2;
}
int cmfb200_serve_topn(void *h, const int_t *users, int_t n_users, const size_t *seen_ptr, const int_t *seen_idx, int_t n_top,
                       int_t *out_ix, real_t *out_score, float *ms_device)
{
    return h ? cmfb200::serve_topn(static_cast<cmfb200::ServeState *>(h), users, n_users, seen_ptr, seen_idx, n_top, out_ix, out_score, ms_device)
             : 2;
}

// C[M x N] = A[M x K] B[N x K]^T on the tensor cores, host buffers in / out (test and measurement aid for gemm_tc.cu)
int cmfb200_gemm_nt(const real_t *A, int lda, int M, const real_t *B, int ldb, int N, int K, real_t *C, int repeats, float *ms_per_launch)
{
    using namespace cmfb200;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return 1;
    DevBuf<real_t> dA, dB, dC;
    if (!dA.alloc((size_t)M * lda + 4) || !dB.alloc((size_t)N * ldb + 4) || !dC.alloc((size_t)M * N)) return 1;
    cudaMemcpy(dA.p, A, (size_t)M * lda * sizeof(real_t), cudaMemcpyHostToDevice);
    cudaMemcpy(dB.p, B, (size_t)N * ldb * sizeof(real_t), cudaMemcpyHostToDevice);
    int rc = launch_gemm_nt_tc(dA.p, lda, M, dB.p, ldb, N, K, dC.p, N, nullptr, nullptr, real_t(0), nullptr);
    if (rc) return rc;
    if (repeats > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, nullptr);
        for (int i = 0; i < repeats; i++) launch_gemm_nt_tc(dA.p, lda, M, dB.p, ldb, N, K, dC.p, N, nullptr, nullptr, real_t(0), nullptr);
        cudaEventRecord(e1, nullptr);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms_per_launch) *ms_per_launch = ms / repeats;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    if (cudaMemcpy(C, dC.p, (size_t)M * N * sizeof(real_t), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    return 0;
}

bool get_has_openmp(void)
{
#ifdef _OPENMP
    return true;
#else
    return false;
#endif
}

const char *cmfb200_real_name(void) { return CMF_REAL_NAME; }

int cmfb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void cmfb200_random_init(real_t *A, size_t sizeA, real_t *B, size_t sizeB, int_t seed, bool normal)
{
    random_init(A, sizeA, B, sizeB, seed, normal);
}

void cmfb200_random_init_threads(real_t *A, size_t sizeA, real_t *B, size_t sizeB, int_t seed, bool normal, int nthreads)
{
    random_init(A, sizeA, B, sizeB, seed, normal, nthreads);
}

void cmfb200_coo_to_csr_and_csc(const int_t *Xrow, const int_t *Xcol, const real_t *Xval, int_t m, int_t n, size_t nnz,
                                size_t *csr_p, int_t *csr_i, real_t *csr_v, size_t *csc_p, int_t *csc_i, real_t *csc_v)
{
    coo_to_csr_and_csc(Xrow, Xcol, Xval, m, n, nnz, csr_p, csr_i, csr_v, csc_p, csc_i, csc_v);
}

real_t cmfb200_global_mean(const real_t *X, size_t nnz, int nthreads) { return global_mean(X, nnz, nthreads); }

void cmfb200_init_biases_twosided(int_t m, int_t n, const size_t *csr_p, const int_t *csr_i, const real_t *csr_v,
                                  const size_t *csc_p, const int_t *csc_i, const real_t *csc_v, real_t lam_user,
                                  real_t lam_item, bool scale_lam, bool nonneg, real_t *biasA, real_t *biasB,
                                  int nthreads)
{
    init_biases_twosided(m, n, csr_p, csr_i, csr_v, csc_p, csc_i, csc_v, lam_user, lam_item, scale_lam, nonneg, biasA,
                         biasB, nthreads);
}

int cmfb200_partition_rows(const size_t *indptr, int_t rows, int world, int_t *to_device_row, int_t *block)
{
    if (!indptr || rows < 0 || world < 1 || !to_device_row) return 2;
    Renumbering ren;
    build_renumbering(indptr, rows, world, ren);
    for (int_t r = 0; r < rows; r++) to_device_row[r] = ren.to_dev[r];
    if (block) *block = ren.block;
    return 0;
}

int cmfb200_nccl_unique_id(void *out128) { return NcclLink::unique_id(out128); }

int cmfb200_als_create(cmfb200_als **out, const cmfb200_als_options *opt, const size_t *csr_p, const int_t *csr_i,
                       const real_t *csr_v, const size_t *csc_p, const int_t *csc_i, const real_t *csc_v)
{
    if (!out || !opt) return 2;
    *out = nullptr;
    if (cmfb200_device_count() < 1) {
        std::fprintf(stderr, "cmfrec_b200: no CUDA device available; this library has no CPU path.\n");
        return 1;
    }
    AlsConfig c;
    c.implicit = opt->implicit != 0;
    c.m = opt->m; c.n = opt->n; c.kk = opt->k;
    c.user_bias = !c.implicit && opt->user_bias; c.item_bias = !c.implicit && opt->item_bias;
    c.lam_A = opt->lam_A; c.lam_B = opt->lam_B; c.lam_biasA = opt->lam_biasA; c.lam_biasB = opt->lam_biasB;
    c.scale_lam = opt->scale_lam != 0;
    c.max_cg_steps = opt->max_cg_steps;
    c.rank = opt->rank; c.world = opt->world < 1 ? 1 : opt->world;
    cmfb200_als *s = new cmfb200_als();
    int rc = s->st.setup(c, csr_p, csr_i, csr_v, csc_p, csc_i, csc_v, (cudaStream_t)opt->stream, opt->nccl_id);
    if (rc) { delete s; return rc; }
    *out = s;
    return 0;
}

// The same state from COO triplets ALREADY ON THE DEVICE (ixA / ixB / X: device pointers; X is multiplied by `scale` and, for
// the explicit model, reduced by `mu`, IN PLACE): both orientations are built, and with world > 1 dealt, on the GPU.  This is how
// workloads too large to stage on the host are set up (BASELINE config 5: 10^9 entries generated on the device).
int cmfb200_als_create_from_device_coo(cmfb200_als **out, const cmfb200_als_options *opt, const int_t *d_ixA, const int_t *d_ixB,
                                       real_t *d_X, size_t nnz, real_t mu, real_t scale)
{
    if (!out || !opt || (nnz && (!d_ixA || !d_ixB || !d_X))) return 2;
    *out = nullptr;
    if (cmfb200_device_count() < 1) {
        std::fprintf(stderr, "cmfrec_b200: no CUDA device available; this library has no CPU path.\n");
        return 1;
    }
    if (nnz > (size_t)2147483647) return 2;
    AlsConfig c;
    c.implicit = opt->implicit != 0;
    c.m = opt->m; c.n = opt->n; c.kk = opt->k;
    c.user_bias = !c.implicit && opt->user_bias; c.item_bias = !c.implicit && opt->item_bias;
    c.lam_A = opt->lam_A; c.lam_B = opt->lam_B; c.lam_biasA = opt->lam_biasA; c.lam_biasB = opt->lam_biasB;
    c.scale_lam = opt->scale_lam != 0;
    c.max_cg_steps = opt->max_cg_steps;
    c.rank = opt->rank; c.world = opt->world < 1 ? 1 : opt->world;
    cmfb200_als *s = new cmfb200_als();
    int rc = s->st.setup_from_coo(c, d_ixA, d_ixB, d_X, nnz, mu, scale, (cudaStream_t)opt->stream, nullptr, opt->nccl_id, nullptr, true);
    if (rc) { delete s; return rc; }
    *out = s;
    return 0;
}

// A filled on the device with uniform values in (0, scale) that depend only on (seed, row, column) -- identical replicas on
// every rank --, B and the biases zero: a starting point for benchmarks whose factors never exist on the host
int cmfb200_als_random_factors(cmfb200_als *s, unsigned long long seed, real_t scale)
{
    if (!s) return 2;
    return s->st.random_factors(seed, scale);
}

void cmfb200_als_destroy(cmfb200_als *s) { delete s; }

int cmfb200_als_set_factors(cmfb200_als *s, const real_t *A, const real_t *biasA, const real_t *B, const real_t *biasB)
{
    return s->st.upload_factors(A, s->st.cfg.kk, biasA, B, s->st.cfg.kk, biasB);
}

int cmfb200_als_get_factors(cmfb200_als *s, real_t *A, real_t *biasA, real_t *B, real_t *biasB)
{
    return s->st.download_factors(A, s->st.cfg.kk, biasA, B, s->st.cfg.kk, biasB);
}

int cmfb200_als_half_sweep(cmfb200_als *s, int which, int iter, int solver)
{
    int rc = s->st.half_sweep(which, iter, solver);
    if (rc) return rc;
    return s->st.exchange(which);
}

int cmfb200_als_iterate(cmfb200_als *s, int first_iter, int n_iters, int niter_total, int use_cg, int finalize_chol)
{
    return s->st.iterate(first_iter, n_iters, niter_total, use_cg != 0, finalize_chol != 0);
}

int cmfb200_als_timed_iterate(cmfb200_als *s, int first_iter, int n_iters, int niter_total, int use_cg,
                              int finalize_chol, float *elapsed_ms)
{
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return 1;
    cudaEventRecord(e0, s->st.stream);
    int rc = s->st.iterate(first_iter, n_iters, niter_total, use_cg != 0, finalize_chol != 0);
    cudaEventRecord(e1, s->st.stream);
    if (cudaEventSynchronize(e1) != cudaSuccess) rc = rc ? rc : 1;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (elapsed_ms) *elapsed_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

void cmfb200_als_set_profile(cmfb200_als *s, int on) { s->st.profile = on != 0; }

int cmfb200_als_read_profile(cmfb200_als *s, int which, double *total_ms, long long *count)
{
    return s->st.read_profile(which, total_ms, count);
}

int cmfb200_als_attach_collective(cmfb200_als *s, const real_t *U_centred, int p, const real_t *I_centred, int q,
                                  int add_implicit_features, real_t w_user, real_t w_item, real_t w_implicit, real_t lam_C,
                                  real_t lam_D, real_t lam_Bi, real_t lam_Ai)
{
    if (!s || s->st.cfg.implicit || s->st.coll) return 2;
    if (s->st.cfg.world != 1 && (U_centred || I_centred)) return 2;   // sharded: implicit features only (collective.cu)
    CollectiveConfig cc;
    cc.p = U_centred ? p : 0;
    cc.q = I_centred ? q : 0;
    cc.implicit_features = add_implicit_features != 0;
    cc.w_user = w_user; cc.w_item = w_item; cc.w_implicit = w_implicit;
    cc.lam_C = lam_C; cc.lam_D = lam_D; cc.lam_Bi = lam_Bi; cc.lam_Ai = lam_Ai;
    CollectiveState *cs = new CollectiveState();
    int rc = cs->setup(&s->st, cc, U_centred, I_centred);
    if (rc) { delete cs; return rc; }
    s->st.coll = cs;
    return 0;
}

int cmfb200_als_get_collective(cmfb200_als *s, real_t *C, real_t *D, real_t *Ai, real_t *Bi)
{
    if (!s || !s->st.coll) return 2;
    return s->st.coll->download(C, D, Ai, Bi);
}

// gram[kk x kk] = G^T G for a dense host matrix G [rows x kk] (reference: the cblas_tsyrk call, src/common.c:3328).
// The device copy uses the factor layout (rows padded to whole 128-byte lines); `repeats` launches are timed with CUDA
// events and their mean is returned in *ms_per_launch (optional).
int cmfb200_gram(const real_t *G, int_t rows, int kk, real_t *gram, int repeats, float *ms_per_launch)
{
    using namespace cmfb200;
    if (!G || !gram || rows < 1 || kk < 1) return 2;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return 1;
    const int ld = cmf_ld_for(kk);
    DevBuf<real_t> dG, dgram, ws;
    if (!dG.alloc((size_t)rows * ld) || !dgram.alloc((size_t)kk * kk) || !ws.alloc(gram_workspace_elems(kk))) return 1;
    cudaStream_t st = nullptr;
    cudaMemsetAsync(dG.p, 0, dG.n * sizeof(real_t), st);
    if (cudaMemcpy2DAsync(dG.p, (size_t)ld * sizeof(real_t), G, (size_t)kk * sizeof(real_t), (size_t)kk * sizeof(real_t), rows,
                          cudaMemcpyHostToDevice, st) != cudaSuccess)
        return 1;
    int rc = launch_gram(dG.p, ld, rows, kk, dgram.p, ws.p, st);
    if (rc) return rc;
    if (repeats > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        for (int i = 0; i < repeats && rc == 0; i++) rc = launch_gram(dG.p, ld, rows, kk, dgram.p, ws.p, st);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms_per_launch) *ms_per_launch = ms / repeats;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        if (rc) return rc;
    }
    if (cudaMemcpyAsync(gram, dgram.p, (size_t)kk * kk * sizeof(real_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
    return cudaStreamSynchronize(st) == cudaSuccess ? 0 : 1;
}

void cmfb200_trim_pool(void) { cmfb200::devbuf_trim_pool(); }

// Every later fit_collective_*_als call of this process runs as rank `rank` of `world` (one process per GPU, the same
// arguments on every rank, every rank receives the full result); world = 1 switches back.
int cmfb200_set_world(int rank, int world, const void *nccl_id128)
{
    cmfb200::WorldSetting &w = cmfb200::world_setting();
    if (world <= 1) {
        w.rank = 0;
        w.world = 1;
        return 0;
    }
    if (rank < 0 || rank >= world || !nccl_id128) return 2;
    w.rank = rank;
    w.world = world;
    std::memcpy(w.nccl_id, nccl_id128, 128);
    return 0;
}

namespace {
__global__ void poison_smem_kernel(unsigned pattern)
{
    extern __shared__ unsigned poison_words[];
    const unsigned nwords = 220u * 1024u / 4u;
    for (unsigned i = threadIdx.x; i < nwords; i += blockDim.x) poison_words[i] = pattern;
    __syncthreads();
    if (poison_words[(threadIdx.x * 7u) % nwords] != pattern) __trap();   // keeps the stores alive
}
}  // namespace

int cmfb200_debug_poison_smem(unsigned pattern)
{
    if (cudaFuncSetAttribute(poison_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) return 1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    poison_smem_kernel<<<sms, 256, 220 * 1024>>>(pattern);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}

int cmfb200_als_sync(cmfb200_als *s) { return cudaStreamSynchronize(s->st.stream) == cudaSuccess ? 0 : 1; }

long long cmfb200_als_launch_count(const cmfb200_als *s) { return s->st.launches; }

void cmfb200_als_local_counts(const cmfb200_als *s, size_t *nnzA, size_t *nnzB, int_t *rowsA, int_t *rowsB)
{
    if (nnzA) *nnzA = s->st.byA.nnz_local;
    if (nnzB) *nnzB = s->st.byB.nnz_local;
    if (rowsA) *rowsA = s->st.byA.n_order;
    if (rowsB) *rowsB = s->st.byB.n_order;
}

}  // extern "C"
