// Fold-in of new rows on the GPU; see foldin.h.  The row solves are AlsState::half_sweep(which = 1, solver = Cholesky)
// with the item factors fixed: sweep_nm.cu (fp32, tcgen05) / sweep_chol_dmma.cu (fp64, DMMA) / sweep_chol.cu.
#include "foldin.h"
#include "als.h"
#include <cmath>
#include <cstdio>
#include <limits>
#include <vector>

namespace cmfb200 {

int foldin_refuse(const char *what)
{
    std::fprintf(stderr, "cmfrec_b200: %s is not supported by the GPU path (no CPU fallback).\n", what);
    return 2;
}

namespace {

int device_ready()
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        std::fprintf(stderr, "cmfrec_b200: no CUDA device available; this library has no CPU path.\n");
        return 1;
    }
    return 0;
}

// COO triplets of the new rows with admissible column ids; rows holding an inadmissible id are flagged (the reference
// answers NaN for them, check_sparse_indices src/collective.c:10633-10643) and their entries dropped
struct NewRows {
    std::vector<int_t> row, col;
    std::vector<real_t> val;
    std::vector<int_t> count;      // admissible entries per row
    std::vector<char> bad;
};

void collect(int_t m, int_t n, const int_t *ixA, const int_t *ixB, const real_t *X, size_t nnz, const size_t *csr_p, const int_t *csr_i,
             const real_t *csr_v, NewRows &o)
{
    o.count.assign((size_t)m, 0);
    o.bad.assign((size_t)m, 0);
    const size_t total = csr_p ? csr_p[m] : nnz;
    o.row.reserve(total); o.col.reserve(total); o.val.reserve(total);
    auto add = [&](int_t r, int_t c, real_t v) {
        if (r < 0 || r >= m) return;
        if (c < 0 || c >= n) { o.bad[r] = 1; return; }
        o.row.push_back(r); o.col.push_back(c); o.val.push_back(v);
        o.count[r]++;
    };
    if (csr_p) {
        for (int_t r = 0; r < m; r++)
            for (size_t e = csr_p[r]; e < csr_p[r + 1]; e++) add(r, csr_i[e], csr_v[e]);
    } else {
        for (size_t e = 0; e < nnz; e++) add(ixA[e], ixB[e], X[e]);
    }
}

}  // namespace

int foldin_explicit(const FoldinExplicitArgs &a)
{
    if (a.m < 1) return 0;
    if (!a.A || !a.B || a.k + a.k_main < 1) return 2;
    if (!a.Xcsr_p && a.nnz && (!a.ixA || !a.ixB || !a.X)) return 2;
    if (int rc = device_ready()) return rc;
    const int kk = a.k + a.k_main;
    const int_t n = (a.include_all_X || a.n == 0) ? a.n_max : a.n;
    const bool user_bias = a.biasA != nullptr;
    // regularisation as factors_collective_explicit_single prepares it (src/collective.c:10609-10629, w_main at :3700-3708)
    real_t lam = a.lam, lam_bias = a.lam;
    if (a.lam_unique) {
        lam_bias = a.lam_unique[user_bias ? 0 : 2];
        lam = a.lam_unique[2];
    }
    bool scale_bias_const = user_bias && a.scale_bias_const;
    const bool scale_lam = a.scale_lam || a.scale_lam_sideinfo;
    if (scale_lam && scale_bias_const) lam_bias *= a.scaling_biasA;
    if (a.w_main != 1) {
        lam /= a.w_main;
        lam_bias /= a.w_main;
    }

    NewRows nr;
    collect(a.m, n, a.ixA, a.ixB, a.X, a.nnz, a.Xcsr_p, a.Xcsr_i, a.Xcsr, nr);
    std::vector<real_t> A0((size_t)a.m * kk, real_t(0)), bias0;
    if (user_bias) bias0.assign((size_t)a.m, real_t(0));
    if (!nr.val.empty()) {
        AlsConfig cfg;
        cfg.implicit = false;
        cfg.m = a.m; cfg.n = n; cfg.kk = kk;
        cfg.user_bias = user_bias; cfg.item_bias = a.biasB != nullptr;
        cfg.lam_A = lam; cfg.lam_B = lam; cfg.lam_biasA = lam_bias; cfg.lam_biasB = lam;
        cfg.scale_lam = scale_lam;
        cfg.scale_bias_const = scale_lam && scale_bias_const;
        if (!user_bias && scale_lam) {
            // without a user bias the reference hands `scale_lam` over in the place of scale_bias_const and `lam` in the place of
            // lam_last (src/collective.c:3780-3796), so the LAST latent coordinate keeps the unscaled lam (src/common.c:718-721)
            cfg.last_coord_special = true;
            cfg.scale_bias_const = true;
            cfg.lam_biasA = lam;
        }
        AlsState st;
        int rc = st.setup_from_coo(cfg, nr.row.data(), nr.col.data(), nr.val.data(), nr.val.size(), a.glob_mean, real_t(1), nullptr);
        if (rc) return rc == 2 ? foldin_refuse("this value of k") : rc;
        if ((rc = st.upload_factors(A0.data(), kk, user_bias ? bias0.data() : nullptr, a.B, kk, a.biasB))) return rc;
        if ((rc = st.half_sweep(1, 0, 1))) return rc == 2 ? foldin_refuse("this value of k") : rc;
        if ((rc = st.download_factors(A0.data(), kk, user_bias ? bias0.data() : nullptr, nullptr, kk, nullptr))) return rc;
        if (cudaStreamSynchronize(nullptr) != cudaSuccess) return 1;
    }
    const real_t nan = std::numeric_limits<real_t>::quiet_NaN();
    for (int_t r = 0; r < a.m; r++) {
        real_t *dst = a.A + (size_t)r * kk;
        if (nr.bad[r]) {
            for (int c = 0; c < kk; c++) dst[c] = nan;
            if (user_bias) a.biasA[r] = nan;
        } else if (nr.count[r] == 0) {          // no information at all: zeros (src/collective.c:3653-3670)
            for (int c = 0; c < kk; c++) dst[c] = 0;
            if (user_bias) a.biasA[r] = 0;
        } else {
            for (int c = 0; c < kk; c++) dst[c] = A0[(size_t)r * kk + c];
            if (user_bias) a.biasA[r] = bias0[r];
        }
    }
    return 0;
}

int foldin_implicit(const FoldinImplicitArgs &a)
{
    if (a.m < 1) return 0;
    if (!a.A || !a.B || a.k + a.k_main < 1 || a.n < 1) return 2;
    if (!a.Xcsr_p && a.nnz && (!a.ixA || !a.ixB || !a.X)) return 2;
    if (int rc = device_ready()) return rc;
    const int kk = a.k + a.k_main;
    real_t lam = a.lam;
    const real_t w_main = a.w_main * a.w_main_multiplier;   // src/collective.c:4002-4006
    if (w_main != 1) lam /= w_main;
    NewRows nr;
    collect(a.m, a.n, a.ixA, a.ixB, a.X, a.nnz, a.Xcsr_p, a.Xcsr_i, a.Xcsr, nr);
    if (a.apply_log_transf)
        for (real_t &v : nr.val) v = std::log(v);
    std::vector<real_t> A0((size_t)a.m * kk, real_t(0));
    if (!nr.val.empty()) {
        AlsConfig cfg;
        cfg.implicit = true;
        cfg.m = a.m; cfg.n = a.n; cfg.kk = kk;
        cfg.lam_A = lam; cfg.lam_B = lam;
        AlsState st;
        int rc = st.setup_from_coo(cfg, nr.row.data(), nr.col.data(), nr.val.data(), nr.val.size(), real_t(0), a.alpha, nullptr);
        if (rc) return rc == 2 ? foldin_refuse("this value of k") : rc;
        if ((rc = st.upload_coordinates(A0.data(), a.B))) return rc;
        if ((rc = st.half_sweep(1, 0, 1))) return rc == 2 ? foldin_refuse("this value of k") : rc;
        if ((rc = st.download_factors(A0.data(), kk, nullptr, nullptr, kk, nullptr))) return rc;
        if (cudaStreamSynchronize(nullptr) != cudaSuccess) return 1;
    }
    const real_t nan = std::numeric_limits<real_t>::quiet_NaN();
    for (int_t r = 0; r < a.m; r++) {
        real_t *dst = a.A + (size_t)r * kk;
        for (int c = 0; c < kk; c++) dst[c] = nr.bad[r] ? nan : (nr.count[r] == 0 ? real_t(0) : A0[(size_t)r * kk + c]);
    }
    return 0;
}

}  // namespace cmfb200
