// Explicit-feedback fits WITH dense side information (U, I) and/or implicit features, on top of AlsState.
//
// reference: fit_collective_explicit_als loop body, src/collective.c:8334-8898 — update order C, D, Bi, Ai, B, A:
//   C  = (A^T A + lamC I)^-1 A^T Uc           optimizeA Case 1 (dense, do_B)        src/common.c:2793-2900
//   Bi = (A^T A + lamI I)^-1 sum_{i in j} a_i  optimizeA Case 3 (sparse, NA as zero) src/common.c:3117-3203
//   B, A: optimizeA_collective "general case"                                        src/collective.c:5566-5968
//         row system  (sum_e g g^T + Q + Lambda) a = sum_e x g + q_i   with
//         Q = w_user C^T C + w_implicit Bi^T Bi,   q_i = w_user C^T u_i + w_implicit sum_e Bi_e
//         (collective_block_cg :2134, collective_closed_form_block :1223).
// Supported here: dense U / I without missing values, m_u == m, n_i == n, k_user = k_item = k_main = 0.
#include "collective.h"
#include "dense_small.h"
#include "postfit.h"
#include "nccl_link.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace cmfb200 {

int CollectiveState::setup(AlsState *state, const CollectiveConfig &c, const real_t *Uc_host, const real_t *Ic_host)
{
    st = state;
    cc = c;
    const int k = st->cfg.kk;
    const int_t m = st->cfg.m, n = st->cfg.n;
    cudaStream_t s = st->stream;
    const int ldq = cmf_ld_for(k);
    // more than one GPU: implicit features only (Ai / Bi are replicated like A / B, device numbering, every rank solves
    // its own block of rows and the blocks are all-gathered); dense side information is not sharded yet
    if (st->cfg.world > 1 && (cc.p > 0 || cc.q > 0)) return 2;
    const size_t mp = (size_t)st->renA.rows_padded, np_ = (size_t)st->renB.rows_padded;
    if (cc.p > 0) {
        if (!Uc.alloc((size_t)m * cc.p) || !C.alloc((size_t)cc.p * k)) return 1;
        if (cudaMemcpyAsync(Uc.p, Uc_host, Uc.n * sizeof(real_t), cudaMemcpyHostToDevice, s) != cudaSuccess) return 1;
        if (cudaMemsetAsync(C.p, 0, C.n * sizeof(real_t), s) != cudaSuccess) return 1;
    }
    if (cc.q > 0) {
        if (!Ic.alloc((size_t)n * cc.q) || !D.alloc((size_t)cc.q * k)) return 1;
        if (cudaMemcpyAsync(Ic.p, Ic_host, Ic.n * sizeof(real_t), cudaMemcpyHostToDevice, s) != cudaSuccess) return 1;
        if (cudaMemsetAsync(D.p, 0, D.n * sizeof(real_t), s) != cudaSuccess) return 1;
    }
    if (cc.implicit_features) {
        if (!Ai.alloc(mp * k) || !Bi.alloc(np_ * k)) return 1;
        if (cudaMemsetAsync(Ai.p, 0, Ai.n * sizeof(real_t), s) != cudaSuccess) return 1;
        if (cudaMemsetAsync(Bi.p, 0, Bi.n * sizeof(real_t), s) != cudaSuccess) return 1;
    }
    const int pmax = std::max(std::max(cc.p, cc.q), k);
    if (!QA.alloc((size_t)k * k) || !QB.alloc((size_t)k * k) || !G1.alloc((size_t)k * k) || !G2.alloc((size_t)k * k) ||
        !T1.alloc((size_t)pmax * k) || !Ldev.alloc((size_t)k * k) || !ws.alloc(std::max(xty_workspace_elems(pmax, k), gram_workspace_elems(k))) ||
        !qA.alloc(mp * ldq) || !qB.alloc(np_ * ldq))
        return 1;
    st->extra_ldq[0] = st->extra_ldq[1] = ldq;
    return cudaStreamSynchronize(s) == cudaSuccess ? 0 : 1;
}

// side-information factor: Cout[p x k] = ((F^T F + lam I)^-1 F^T S)^T, F = factor [rows x ld] (first k columns), S [rows x p]
int CollectiveState::update_side_factor(const real_t *F, int ldF, int_t rows, const real_t *S, int p, real_t lam, real_t *Cout)
{
    const int k = st->cfg.kk;
    cudaStream_t s = st->stream;
    int rc = launch_gram(F, ldF, rows, k, G1.p, ws.p, s);   // F^T F (tcgen05 where gram_tc.cu covers the shape)
    if (rc) return rc;
    if ((rc = launch_xty(S, p, p, F, ldF, k, rows, T1.p, ws.p, s))) return rc;       // S^T F : [p x k]
    // each of the p rows of S^T F is a right-hand side of the k x k system (posv with p right-hand sides)
    if ((rc = launch_spd_factor(G1.p, k, lam, Ldev.p, s))) return rc;
    if ((rc = launch_tri_solve_rows(Ldev.p, k, T1.p, k, p, s))) return rc;
    st->launches += 6;
    return cudaMemcpyAsync(Cout, T1.p, (size_t)p * k * sizeof(real_t), cudaMemcpyDeviceToDevice, s) == cudaSuccess ? 0 : 1;
}

// implicit-features factor: Out[r] = (F^T F + lam I)^-1 sum_{e in row r of X} F[col_e]
int CollectiveState::update_implicit_factor(const DeviceSide &side, const real_t *F, int ldF, int_t rowsF, real_t lam, real_t *Out)
{
    const int k = st->cfg.kk;
    cudaStream_t s = st->stream;
    int rc = launch_gram(F, ldF, rowsF, k, G1.p, ws.p, s);
    if (rc) return rc;
    if ((rc = launch_spd_factor(G1.p, k, lam, Ldev.p, s))) return rc;
    if ((rc = launch_spmm_ones(side.view(), side.plan(), F, ldF, k, real_t(1), false, Out, k, s))) return rc;
    // this rank's rows are the contiguous block [row_begin, row_begin + n_order) of the device numbering (all rows on one GPU)
    if ((rc = launch_tri_solve_rows(Ldev.p, k, Out + (size_t)side.row_begin * k, k, side.n_order, s))) return rc;
    st->launches += 5;
    if (st->link) {
        if ((rc = st->link->all_gather_inplace(Out, (size_t)side.block * k * sizeof(real_t), s))) return rc;
        st->launches += 1;
    }
    return 0;
}

// Q and q for one side.  sideinfo: S [rows x p], Cfac [p x k], weight w_side; implicit: factor Fi of the opposing side
int CollectiveState::build_extras(int which, const DeviceSide &side, int_t rows, const real_t *S, int p, const real_t *Cfac,
                                  real_t w_side, const real_t *Fi_opp, int_t rows_opp)
{
    const int k = st->cfg.kk;
    cudaStream_t s = st->stream;
    real_t *Q = which ? QA.p : QB.p;
    real_t *qv = which ? qA.p : qB.p;
    const int ldq = st->extra_ldq[which];
    int rc = 0;
    const bool has_side = p > 0, has_imp = cc.implicit_features;
    if (has_side && (rc = launch_xty(Cfac, k, k, Cfac, k, k, p, G1.p, ws.p, s))) return rc;   // p rows only
    if (has_imp && (rc = launch_gram(Fi_opp, k, rows_opp, k, G2.p, ws.p, s))) return rc;
    if ((rc = launch_axpby(k * k, w_side, has_side ? G1.p : nullptr, cc.w_implicit, has_imp ? G2.p : nullptr, Q, s))) return rc;
    bool acc = false;
    if (has_side) {
        if ((rc = launch_rows_times_small(S, p, p, Cfac, k, k, w_side, false, qv, ldq, rows, s))) return rc;
        acc = true;
    }
    if (has_imp) {
        if ((rc = launch_spmm_ones(side.view(), side.plan(), Fi_opp, k, k, cc.w_implicit, acc, qv, ldq, s))) return rc;
        acc = true;
    }
    st->launches += 7;
    st->extraQ[which] = Q;
    st->extraq[which] = acc ? qv : nullptr;
    st->extra_all_rows[which] = has_side;
    return 0;
}

int CollectiveState::iteration(int it, int solver)
{
    const int_t m = st->cfg.m, n = st->cfg.n;
    int rc;
    auto say = [&](const char *what) {
        if (st->verbose) {
            std::printf("%s", what);
            std::fflush(stdout);
        }
    };
    auto done = [&]() -> int {
        if (st->verbose && cudaStreamSynchronize(st->stream) != cudaSuccess) return 1;
        say(" done\n");
        return 0;
    };
    // C and D from the current A and B
    if (cc.p > 0) {
        if (stop_flag()) return 3;
        say("Updating C ...");
        if ((rc = update_side_factor(st->A.p, st->ldA, m, Uc.p, cc.p, cc.lam_C, C.p))) return rc;
        if ((rc = done())) return rc;
    }
    if (cc.q > 0) {
        if (stop_flag()) return 3;
        say("Updating D ...");
        if ((rc = update_side_factor(st->B.p, st->ldB, n, Ic.p, cc.q, cc.lam_D, D.p))) return rc;
        if ((rc = done())) return rc;
    }
    // Bi from A, Ai from B (both before B and A are touched)
    if (cc.implicit_features) {
        if (stop_flag()) return 3;
        say("Updating Bi...");
        if ((rc = update_implicit_factor(st->byB, st->A.p, st->ldA, st->renA.rows_padded, cc.lam_Bi, Bi.p))) return rc;
        if ((rc = done())) return rc;
        if (stop_flag()) return 3;
        say("Updating Ai...");
        if ((rc = update_implicit_factor(st->byA, st->B.p, st->ldB, st->renB.rows_padded, cc.lam_Ai, Ai.p))) return rc;
        if ((rc = done())) return rc;
    }
    // B given A (extras use D and Ai), then A given the new B (extras use C and Bi)
    if (stop_flag()) return 3;
    say("Updating B ...");
    if ((rc = build_extras(0, st->byB, n, Ic.p, cc.q, D.p, cc.w_item, Ai.p, st->renA.rows_padded))) return rc;
    if ((rc = st->half_sweep(0, it, solver))) return rc;
    if ((rc = st->exchange(0))) return rc;
    if ((rc = done())) return rc;
    if (stop_flag()) return 3;
    say("Updating A ...");
    if ((rc = build_extras(1, st->byA, m, Uc.p, cc.p, C.p, cc.w_user, Bi.p, st->renB.rows_padded))) return rc;
    if ((rc = st->half_sweep(1, it, solver))) return rc;
    if ((rc = st->exchange(1))) return rc;
    return done();
}

int CollectiveState::iterate(int niter, bool use_cg, bool finalize_chol)
{
    for (int it = 0; it < niter; it++) {
        const bool cg_now = use_cg && !(finalize_chol && it == niter - 1);
        if (int rc = iteration(it, cg_now ? 0 : 1)) return rc;
    }
    return 0;
}

int CollectiveState::download(real_t *hC, real_t *hD, real_t *hAi, real_t *hBi)
{
    cudaStream_t s = st->stream;
    if (hC && C.n) cudaMemcpyAsync(hC, C.p, C.n * sizeof(real_t), cudaMemcpyDeviceToHost, s);
    if (hD && D.n) cudaMemcpyAsync(hD, D.p, D.n * sizeof(real_t), cudaMemcpyDeviceToHost, s);
    const int k = st->cfg.kk;
    if (hAi && Ai.n && st->download_matrix(1, Ai.p, k, hAi, k)) return 1;   // back to the caller's row numbering
    if (hBi && Bi.n && st->download_matrix(0, Bi.p, k, hBi, k)) return 1;
    return cudaStreamSynchronize(s) == cudaSuccess ? 0 : 1;
}

// column means and centred copy of a dense side-information matrix, exactly as center_by_cols does for a dense
// matrix without missing values (src/common.c:4940-4947, 4983-4985): sums accumulated in real_t row by row
int center_side_info(const real_t *S, int_t rows, int p, real_t *colmeans, std::vector<real_t> &centred)
{
    for (size_t i = 0; i < (size_t)rows * p; i++)
        if (std::isnan(S[i])) return 2;
    std::vector<real_t> mean(p, real_t(0));
    for (int_t r = 0; r < rows; r++)
        for (int c = 0; c < p; c++) mean[c] += S[(size_t)r * p + c];
    for (int c = 0; c < p; c++) mean[c] = (real_t)((double)mean[c] / (double)rows);
    centred.resize((size_t)rows * p);
    for (int_t r = 0; r < rows; r++)
        for (int c = 0; c < p; c++) centred[(size_t)r * p + c] = S[(size_t)r * p + c] - mean[c];
    if (colmeans) std::memcpy(colmeans, mean.data(), (size_t)p * sizeof(real_t));
    return 0;
}

}  // namespace cmfb200
