#include "postfit.h"
#include <cmath>
#include <cstring>
#include <vector>

namespace cmfb200 {

// upper triangle of M^T M for a row-major [rows x d] matrix with row stride ld, accumulated in double
static void upper_gram(const real_t *M, size_t rows, int d, size_t ld, real_t *out)
{
    std::vector<double> acc((size_t)d * d, 0.);
#pragma omp parallel
    {
        std::vector<double> loc((size_t)d * d, 0.);
#pragma omp for schedule(static) nowait
        for (long long r = 0; r < (long long)rows; r++) {
            const real_t *x = M + (size_t)r * ld;
            for (int i = 0; i < d; i++) {
                const double xi = x[i];
                double *row = loc.data() + (size_t)i * d;
                for (int j = i; j < d; j++) row[j] += xi * (double)x[j];
            }
        }
#pragma omp critical
        for (size_t t = 0; t < acc.size(); t++) acc[t] += loc[t];
    }
    for (int i = 0; i < d; i++)
        for (int j = i; j < d; j++) out[(size_t)i * d + j] = (real_t)acc[(size_t)i * d + j];
}

int host_spd_solve_rows(size_t d, const real_t *S_upper, real_t *R, size_t nrhs)
{
    // factor S = L L^T in double, L stored row-major lower
    std::vector<double> Lm(d * d, 0.);
    for (size_t i = 0; i < d; i++)
        for (size_t j = 0; j <= i; j++) Lm[i * d + j] = S_upper[j * d + i];
    for (size_t j = 0; j < d; j++) {
        double s = Lm[j * d + j];
        for (size_t t = 0; t < j; t++) s -= Lm[j * d + t] * Lm[j * d + t];
        if (!(s > 0)) return 1;
        const double dj = std::sqrt(s);
        Lm[j * d + j] = dj;
        for (size_t i = j + 1; i < d; i++) {
            double v = Lm[i * d + j];
            for (size_t t = 0; t < j; t++) v -= Lm[i * d + t] * Lm[j * d + t];
            Lm[i * d + j] = v / dj;
        }
    }
#pragma omp parallel for schedule(static)
    for (long long r = 0; r < (long long)nrhs; r++) {
        real_t *x = R + (size_t)r * d;
        std::vector<double> y(d);
        for (size_t i = 0; i < d; i++) {
            double v = x[i];
            for (size_t t = 0; t < i; t++) v -= Lm[i * d + t] * y[t];
            y[i] = v / Lm[i * d + i];
        }
        for (size_t ii = d; ii-- > 0;) {
            double v = y[ii];
            for (size_t t = ii + 1; t < d; t++) v -= Lm[t * d + ii] * y[t];
            y[ii] = v / Lm[ii * d + ii];
        }
        for (size_t i = 0; i < d; i++) x[i] = (real_t)y[i];
    }
    return 0;
}

// upper-triangular Cholesky factor (U^T U = S, what LAPACK's potrf('L') leaves in a row-major array) of the symmetric
// matrix whose upper triangle is S_upper; the strictly lower part of `out` is left as it is.  NaN-filled when S is not
// positive definite (the reference would return LAPACK's partial result; nothing downstream can use either).
static void upper_cholesky(size_t d, const real_t *S_upper, real_t *out)
{
    std::vector<double> U(d * d, 0.);
    bool ok = true;
    for (size_t i = 0; i < d && ok; i++) {
        for (size_t j = i; j < d; j++) {
            double v = S_upper[i * d + j];
            for (size_t t = 0; t < i; t++) v -= U[t * d + i] * U[t * d + j];
            if (j == i) {
                if (!(v > 0)) { ok = false; break; }
                U[i * d + i] = std::sqrt(v);
            } else {
                U[i * d + j] = v / U[i * d + i];
            }
        }
    }
    for (size_t i = 0; i < d; i++)
        for (size_t j = i; j < d; j++) out[i * d + j] = ok ? (real_t)U[i * d + j] : (real_t)NAN;
}

int postfit_explicit(const PostfitExplicit &a)
{
    const int kk = a.kk;
    const bool has_bias = a.user_bias || a.item_bias;
    const int ldb = kk + (has_bias ? 1 : 0);
    const int d = kk + (a.user_bias ? 1 : 0);
    // The reference's working copy of B: factors plus one extra column that ends the fit holding 1.0 when
    // users have a bias (it multiplies the user bias) and the item bias otherwise
    // (src/collective.c:8723-8736).  That buffer IS B_plus_bias when the caller passes one.
    std::vector<real_t> own;
    real_t *Bb = a.B_plus_bias;
    if (!has_bias) {
        Bb = nullptr;
    } else if (!Bb) {
        own.resize((size_t)a.n * ldb);
        Bb = own.data();
    }
    if (has_bias) {
        for (int_t r = 0; r < a.n; r++) {
            std::memcpy(Bb + (size_t)r * ldb, a.B + (size_t)r * kk, (size_t)kk * sizeof(real_t));
            Bb[(size_t)r * ldb + kk] = a.user_bias ? real_t(1) : (a.biasB ? a.biasB[r] : real_t(0));
        }
    }
    const real_t *M = has_bias ? Bb : a.B;
    // BtB is needed by everything below whether or not the caller asked for it
    std::vector<real_t> own_btb;
    real_t *BtB = a.BtB;
    if (!BtB) {
        own_btb.assign((size_t)d * d, real_t(0));
        BtB = own_btb.data();
    }
    upper_gram(M, a.n, d, ldb, BtB);
    const real_t mult = a.scale_lam ? (real_t)a.n : real_t(1);
    int rc = 0;
    if (a.TransBtBinvBt && !a.implicit_features) {            // src/collective.c:9001
        std::vector<real_t> S((size_t)d * d);
        std::memcpy(S.data(), BtB, S.size() * sizeof(real_t));
        for (int i = 0; i < d; i++) S[(size_t)i * d + i] += a.lam * mult;
        if (a.user_bias && a.lam_bias != a.lam) S[(size_t)d * d - 1] += (a.lam_bias - a.lam) * mult;
        for (int_t r = 0; r < a.n; r++)
            std::memcpy(a.TransBtBinvBt + (size_t)r * d, M + (size_t)r * ldb, (size_t)d * sizeof(real_t));
        if (host_spd_solve_rows(d, S.data(), a.TransBtBinvBt, a.n)) {
            for (size_t t = 0; t < (size_t)a.n * d; t++) a.TransBtBinvBt[t] = (real_t)NAN;   // like a failed posv
            rc = 0;
        }
    }
    if (a.implicit_features && a.Bi && a.BiTBi) {               // src/collective.c:8975-8981
        upper_gram(a.Bi, a.n, kk, kk, a.BiTBi);
        for (int i = 0; i < kk; i++)
            for (int j = i; j < kk; j++) a.BiTBi[(size_t)i * kk + j] *= a.w_implicit;
    }
    std::vector<real_t> ctc;                                       // C^T C, unscaled, upper
    if (a.p > 0 && a.C) {
        ctc.assign((size_t)kk * kk, real_t(0));
        upper_gram(a.C, a.p, kk, kk, ctc.data());
        if (a.TransCtCinvCt && !a.implicit_features) {          // src/collective.c:9084-9150
            std::vector<real_t> S(ctc);
            const real_t reg = a.lam * (a.scale_lam ? (real_t)a.p : real_t(1)) / a.w_user;
            for (int i = 0; i < kk; i++) S[(size_t)i * kk + i] += reg;
            std::memcpy(a.TransCtCinvCt, a.C, (size_t)a.p * kk * sizeof(real_t));
            if (host_spd_solve_rows(kk, S.data(), a.TransCtCinvCt, a.p))
                for (size_t t = 0; t < (size_t)a.p * kk; t++) a.TransCtCinvCt[t] = (real_t)NAN;
        }
        if (a.CtCw)
            for (int i = 0; i < kk; i++)
                for (int j = i; j < kk; j++) a.CtCw[(size_t)i * kk + j] = a.w_user * ctc[(size_t)i * kk + j];
    }
    if (a.BeTBeChol && (a.p > 0 || a.implicit_features)) {      // src/collective.c:9165-9235
        std::vector<real_t> S((size_t)d * d, real_t(0));
        for (int i = 0; i < d; i++)
            for (int j = i; j < d; j++) S[(size_t)i * d + j] = BtB[(size_t)i * d + j];
        if (a.p > 0 && a.C)
            for (int i = 0; i < kk; i++)
                for (int j = i; j < kk; j++) S[(size_t)i * d + j] += a.w_user * ctc[(size_t)i * kk + j];
        if (a.implicit_features && a.Bi) {
            std::vector<real_t> bitbi((size_t)kk * kk, real_t(0));
            upper_gram(a.Bi, a.n, kk, kk, bitbi.data());
            for (int i = 0; i < kk; i++)
                for (int j = i; j < kk; j++) S[(size_t)i * d + j] += a.w_implicit * bitbi[(size_t)i * kk + j];
        }
        const real_t m2 = a.scale_lam_sideinfo ? (real_t)(a.p + a.n) : mult;
        for (int i = 0; i < d; i++) S[(size_t)i * d + i] += a.lam * m2;
        if (a.user_bias && a.lam_bias != a.lam) S[(size_t)d * d - 1] += (a.lam_bias - a.lam) * m2;
        for (size_t t = 0; t < (size_t)d * d; t++) a.BeTBeChol[t] = real_t(0);
        upper_cholesky(d, S.data(), a.BeTBeChol);
    }
    return rc;
}

int postfit_implicit(const real_t *B, int_t n, int kk, real_t lam, real_t *BtB, const real_t *C, int p, real_t w_user,
                     real_t *BeTBe, real_t *BeTBeChol, bool last_was_cg)
{
    if (last_was_cg) w_user = 1;      // reference quirk, see postfit.h
    std::vector<real_t> own;
    if (!BtB) {
        own.assign((size_t)kk * kk, real_t(0));
        BtB = own.data();
    }
    upper_gram(B, n, kk, kk, BtB);
    for (int i = 0; i < kk; i++) BtB[(size_t)i * kk + i] += lam;
    if (C && p > 0 && BeTBe) {                                    // src/collective.c:10073-10105
        for (size_t t = 0; t < (size_t)kk * kk; t++) BeTBe[t] = real_t(0);
        upper_gram(C, p, kk, kk, BeTBe);
        for (int i = 0; i < kk; i++)
            for (int j = i; j < kk; j++) BeTBe[(size_t)i * kk + j] = w_user * BeTBe[(size_t)i * kk + j] + BtB[(size_t)i * kk + j];
        if (BeTBeChol) {
            for (size_t t = 0; t < (size_t)kk * kk; t++) BeTBeChol[t] = real_t(0);
            upper_cholesky(kk, BeTBe, BeTBeChol);
        }
    }
    return 0;
}

}  // namespace cmfb200
