/* TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
 *
 * Plain sequential C restatement of the reference's ALS path (david-cortes/cmfrec v3.5.1), used only as a
 * checker by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg.  Nothing under cmfrec_b200/ may
 * call into this file.  Every function cites the reference lines it restates.  No BLAS: dot/axpy/symv/posv are
 * written out as loops in the reference's order of operations (nonzeros in CSR order, coordinates in index
 * order), so results agree with the reference build to summation-order noise.
 *
 * Parity pin: checked against the reference itself (oracle/_ref, built from /root/reference/src by
 * oracle/Makefile) in tests/test_oracle_vs_reference.py, and against the committed golden vectors in
 * tests/golden/ that were generated from that same reference build (tools/make_golden.py).
 *
 * Build: gcc -O2 -std=c99 -ffp-contract=off [-DUSE_FLOAT] -shared -fPIC cmf_oracle.c -lm
 */
#include <math.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifdef USE_FLOAT
typedef float real_t;
#define EPS_T FLT_EPSILON
#else
typedef double real_t;
#define EPS_T DBL_EPSILON
#endif

/* ---------------------------------------------------------------------------------------------------------
 * RNG: splitmix64 -> xoshiro256++ ; ziggurat normal / uniform          reference src/helpers.c:505-568, 653-864
 * ------------------------------------------------------------------------------------------------------- */
#define ZIG_TYPE_ki_double uint64_t
#define ZIG_TYPE_wi_double double
#define ZIG_TYPE_fi_double double
#define ZIG_TYPE_ki_float uint32_t
#define ZIG_TYPE_wi_float float
#define ZIG_TYPE_fi_float float
#define ZIG_BEGIN(name) __attribute__((unused)) static const ZIG_TYPE_##name T_##name[256] = {
#define ZIG_END(name) };
#define ZIG_U(i, v) v,
#define ZIG_F(i, v) v,
#include "../cmfrec_b200/csrc/zig_tables.inc"

static uint64_t sm64(uint64_t s)
{
    uint64_t z = s + 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static uint64_t rol(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t xo(uint64_t s[4])
{
    uint64_t r = rol(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rol(s[3], 45);
    return r;
}
static void xo_jump(uint64_t s[4])          /* src/helpers.c:541-568 */
{
    static const uint64_t J[4] = {0x180ec6d33cfd0abaULL, 0xd5a61266f0c9392cULL, 0xa9582618e03fc9aaULL, 0x39abdc4529b1661cULL};
    uint64_t a[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++)
        for (int b = 0; b < 64; b++) {
            if (J[i] & (1ULL << b)) { a[0] ^= s[0]; a[1] ^= s[1]; a[2] ^= s[2]; a[3] ^= s[3]; }
            xo(s);
        }
    memcpy(s, a, sizeof(a));
}

#ifndef USE_FLOAT
static void rnorm(real_t *seq, size_t n, uint64_t st[4])      /* src/helpers.c:653-721 */
{
    size_t ix = 0;
    while (ix < n) {
        uint64_t rnd = xo(st);
        unsigned rect = rnd & 255; rnd >>= 8;
        unsigned sign = rnd & 1; rnd >>= 4;
        double x = rnd * T_wi_double[rect];
        if (rnd < T_ki_double[rect]) seq[ix++] = sign ? x : -x;
        else if (rect != 0) {
            uint64_t r2 = xo(st);
            double u = ((double)(r2 >> 12) + 0.5) * 0x1.0p-52;
            if (u * (T_fi_double[rect - 1] - T_fi_double[rect]) < exp(-0.5 * x * x) - T_fi_double[rect])
                seq[ix++] = sign ? x : -x;
        }
    }
    for (size_t i = 0; i < n; i++) seq[i] *= 0x1.0p-7;
}
static void runif(real_t *seq, size_t n, uint64_t st[4])      /* src/helpers.c:723-732 */
{
    for (size_t i = 0; i < n; i++) seq[i] = ((double)(xo(st) >> 12) + 0.5) * 0x1.0p-59;
}
#else
static void rnorm(real_t *seq, size_t n, uint64_t st[4])      /* src/helpers.c:750-836 */
{
    uint64_t big = 0; bool reuse = false; size_t ix = 0;
    while (ix < n) {
        uint32_t rnd;
        if (reuse) { reuse = false; rnd = (uint32_t)big; } else { big = xo(st); reuse = true; rnd = big & 0xffffffff; big >>= 32; }
        unsigned rect = rnd & 255; rnd >>= 8;
        unsigned sign = rnd & 1; rnd >>= 1;
        float x = rnd * T_wi_float[rect];
        if (rnd < T_ki_float[rect]) seq[ix++] = sign ? x : -x;
        else {
            if (reuse) { reuse = false; rnd = (uint32_t)big; } else { big = xo(st); reuse = true; rnd = big & 0xffffffff; big >>= 32; }
            float u = ((float)(rnd >> 9) + 0.5f) * 0x1.0p-23f;
            /* rect == 0 reads one element before the fi table in the reference: that is the last wi entry */
            float fprev = rect ? T_fi_float[rect - 1] : T_wi_float[255];
            if (u * (fprev - T_fi_float[rect]) < expf(-0.5f * x * x) - T_fi_float[rect]) seq[ix++] = sign ? x : -x;
        }
    }
    for (size_t i = 0; i < n; i++) seq[i] *= 0x1.0p-7f;
}
static void runif(real_t *seq, size_t n, uint64_t st[4])      /* src/helpers.c:838-864 */
{
    size_t lim = n >> 1;
    for (size_t i = 0; i < lim; i++) {
        uint64_t r = xo(st);
        seq[2 * i] = ((float)(r & 0x7fffff) + 0.5f) * 0x1.0p-30f;
        seq[2 * i + 1] = ((float)(r >> 41) + 0.5f) * 0x1.0p-30f;
    }
    if ((lim << 1) < n) { uint64_t r = xo(st); if (lim) seq[lim - 1] = ((float)(r & 0x7fffff) + 0.5f) * 0x1.0p-30f; }
}
#endif

/* src/helpers.c:877-1043 (seed_state + random_parallel, incl. its one-bucket-per-array behaviour) */
void oracle_random_init(real_t *A, size_t sizeA, real_t *B, size_t sizeB, int seed, bool normal)
{
    uint64_t st[4], stB[4];
    st[0] = sm64((uint64_t)(int64_t)seed); st[1] = sm64(st[0]); st[2] = sm64(st[1]); st[3] = sm64(st[2]);
    if (sizeA + sizeB <= ((size_t)1 << 18)) {
        if (sizeA) rnorm(A, sizeA, st);
        if (sizeB) rnorm(B, sizeB, st);
        return;
    }
    memcpy(stB, st, sizeof(st));
    if (sizeA && sizeB) xo_jump(stB);
    if (sizeA) { if (normal) rnorm(A, sizeA, st); else runif(A, sizeA, st); }
    if (sizeB) { if (normal) rnorm(B, sizeB, stB); else runif(B, sizeB, stB); }
}

/* ---------------------------------------------------------------------------------------------------------
 * COO -> CSR (stable)                                                           reference src/helpers.c:1375-1447
 * ------------------------------------------------------------------------------------------------------- */
void oracle_coo_to_csr(const int *row, const int *col, const real_t *val, int m, size_t nnz, size_t *p, int *ix, real_t *v)
{
    memset(p, 0, ((size_t)m + 1) * sizeof(size_t));
    for (size_t e = 0; e < nnz; e++) p[row[e] + 1]++;
    for (int r = 0; r < m; r++) p[r + 1] += p[r];
    int *cnt = (int *)calloc(m > 0 ? m : 1, sizeof(int));
    for (size_t e = 0; e < nnz; e++) {
        size_t d = p[row[e]] + cnt[row[e]]++;
        v[d] = val[e]; ix[d] = col[e];
    }
    free(cnt);
}

/* ---------------------------------------------------------------------------------------------------------
 * global mean                                                            reference src/common.c:3494-3513, 3603
 * ------------------------------------------------------------------------------------------------------- */
real_t oracle_global_mean(const real_t *X, size_t nnz, int nthreads)
{
    double s = 0; real_t out;
    if (nthreads >= 8) { for (size_t e = 0; e < nnz; e++) s += X[e]; out = (real_t)(s / (double)nnz); }
    else { size_t c = 0; for (size_t e = 0; e < nnz; e++) s += (X[e] - s) / (double)(++c); out = (real_t)s; }
#ifdef USE_FLOAT
    if (fabsf(out) < sqrtf(EPS_T)) out = 0;
#else
    if (fabs(out) < sqrt(EPS_T)) out = 0;
#endif
    return out;
}

/* ---------------------------------------------------------------------------------------------------------
 * two-sided bias initialisation, sparse X                       reference src/common.c:4410-4909 (:4643, :4799)
 * ------------------------------------------------------------------------------------------------------- */
void oracle_init_biases_twosided(int m, int n, const size_t *rp, const int *ri, const real_t *rv, const size_t *cp,
                                 const int *ci, const real_t *cv, real_t lam_user, real_t lam_item, bool scale_lam,
                                 real_t *biasA, real_t *biasB)
{
    if (fabs(lam_user) < EPS_T) lam_user = EPS_T;
    if (fabs(lam_item) < EPS_T) lam_item = EPS_T;
    memset(biasA, 0, (size_t)m * sizeof(real_t));
    memset(biasB, 0, (size_t)n * sizeof(real_t));
    for (int it = 0; it < 5; it++) {
        for (int c = 0; c < n; c++) {
            double b = 0;
            for (size_t t = cp[c]; t < cp[c + 1]; t++) { real_t d = cv[t] - biasA[ci[t]]; b += ((double)d - b) / (double)(t - cp[c] + 1); }
            size_t cnt = cp[c + 1] - cp[c];
            b *= (double)cnt / ((double)cnt + lam_item * (scale_lam ? (double)(cnt > 1 ? cnt : 1) : 1.));
            biasB[c] = (real_t)b;
        }
        for (int r = 0; r < m; r++) {
            double b = 0;
            for (size_t t = rp[r]; t < rp[r + 1]; t++) { real_t d = rv[t] - biasB[ri[t]]; b += ((double)d - b) / (double)(t - rp[r] + 1); }
            size_t cnt = rp[r + 1] - rp[r];
            if (cnt) b *= (double)cnt / ((double)cnt + lam_user * (scale_lam ? (double)cnt : 1.));
            biasA[r] = (real_t)b;
        }
    }
}

/* one-sided: shrunken row means                                          reference src/common.c:4266-4289 */
void oracle_init_biases_onesided(int m, const size_t *rp, const real_t *rv, real_t lam, bool scale_lam, real_t *bias)
{
    if (fabs(lam) < EPS_T) lam = EPS_T;
    for (int r = 0; r < m; r++) {
        double b = 0;
        for (size_t t = rp[r]; t < rp[r + 1]; t++) b += ((double)rv[t] - b) / (double)(t - rp[r] + 1);
        size_t cnt = rp[r + 1] - rp[r];
        b *= (double)cnt / ((double)cnt + lam * (scale_lam ? (double)(cnt > 1 ? cnt : 1) : 1.));
        bias[r] = (real_t)b;
    }
}

/* ---------------------------------------------------------------------------------------------------------
 * small dense helpers
 * ------------------------------------------------------------------------------------------------------- */
static real_t dot(int k, const real_t *x, const real_t *y) { real_t s = 0; for (int i = 0; i < k; i++) s += x[i] * y[i]; return s; }
static void axpy(int k, real_t a, const real_t *x, real_t *y) { for (int i = 0; i < k; i++) y[i] += a * x[i]; }

/* solve S a = b in place (b), S symmetric positive definite k x k, upper triangle (row-major) read; S destroyed */
static void spd_solve(int k, real_t *S, real_t *b)
{
    /* Cholesky S = U^T U on the upper triangle */
    for (int j = 0; j < k; j++) {
        real_t d = S[j * k + j];
        for (int t = 0; t < j; t++) d -= S[t * k + j] * S[t * k + j];
        d = (real_t)sqrt(d);
        S[j * k + j] = d;
        for (int c = j + 1; c < k; c++) {
            real_t v = S[j * k + c];
            for (int t = 0; t < j; t++) v -= S[t * k + j] * S[t * k + c];
            S[j * k + c] = v / d;
        }
    }
    for (int i = 0; i < k; i++) { real_t v = b[i]; for (int t = 0; t < i; t++) v -= S[t * k + i] * b[t]; b[i] = v / S[i * k + i]; }
    for (int i = k - 1; i >= 0; i--) { real_t v = b[i]; for (int t = i + 1; t < k; t++) v -= S[i * k + t] * b[t]; b[i] = v / S[i * k + i]; }
}

/* ---------------------------------------------------------------------------------------------------------
 * row solvers
 * ------------------------------------------------------------------------------------------------------- */
/* factors_explicit_cg                                                       reference src/common.c:1098-1188 */
static void row_explicit_cg(real_t *a, int k, const real_t *B, int ldb, const real_t *Xa, const int *ixB, size_t nnz,
                            real_t lam, real_t lam_last, int steps, real_t *buf)
{
    real_t *Ap = buf, *p = buf + k, *r = buf + 2 * k;
    memset(r, 0, k * sizeof(real_t));
    for (size_t e = 0; e < nnz; e++) {
        const real_t *b = B + (size_t)ixB[e] * ldb;
        real_t coef = dot(k, b, a) - Xa[e];
        axpy(k, -coef, b, r);
    }
    axpy(k, -lam, a, r);
    if (lam != lam_last) r[k - 1] -= (lam_last - lam) * a[k - 1];
    real_t r_old = dot(k, r, r);
    if (r_old <= 1e-12) return;
    memcpy(p, r, k * sizeof(real_t));
    for (int s = 0; s < steps; s++) {
        memset(Ap, 0, k * sizeof(real_t));
        for (size_t e = 0; e < nnz; e++) { const real_t *b = B + (size_t)ixB[e] * ldb; axpy(k, dot(k, b, p), b, Ap); }
        axpy(k, lam, p, Ap);
        if (lam != lam_last) Ap[k - 1] += (lam_last - lam) * p[k - 1];
        real_t al = r_old / dot(k, p, Ap);
        axpy(k, al, p, a);
        axpy(k, -al, Ap, r);
        real_t r_new = dot(k, r, r);
        if (r_new <= 1e-8) break;
        real_t be = r_new / r_old;
        for (int i = 0; i < k; i++) p[i] = be * p[i] + r[i];
        r_old = r_new;
    }
}

/* factors_closed_form, sparse Cholesky branch                 reference src/common.c:978-1013 + 1058-1070 */
static void row_explicit_chol(real_t *a, int k, const real_t *B, int ldb, const real_t *Xa, const int *ixB, size_t nnz,
                              real_t lam, real_t lam_last, real_t *buf)
{
    real_t *S = buf;
    memset(a, 0, k * sizeof(real_t));
    for (size_t e = 0; e < nnz; e++) axpy(k, Xa[e], B + (size_t)ixB[e] * ldb, a);
    memset(S, 0, (size_t)k * k * sizeof(real_t));
    for (size_t e = 0; e < nnz; e++) {
        const real_t *b = B + (size_t)ixB[e] * ldb;
        for (int i = 0; i < k; i++) for (int j = i; j < k; j++) S[i * k + j] += b[i] * b[j];
    }
    for (int i = 0; i < k - 1; i++) S[i * k + i] += lam;
    S[k * k - 1] += lam_last;
    spd_solve(k, S, a);
}

/* factors_implicit_cg                                                       reference src/common.c:1914-1986 */
static void row_implicit_cg(real_t *a, int k, const real_t *B, size_t ldb, const real_t *Xa, const int *ixB, size_t nnz,
                            real_t lam, const real_t *BtB, int steps, real_t *buf)
{
    real_t *Ap = buf, *r = buf + k, *p = buf + 2 * k;
    for (int i = 0; i < k; i++) { real_t s = 0; for (int j = 0; j < k; j++) s += BtB[(i <= j) ? i * k + j : j * k + i] * a[j]; r[i] = -s; }
    for (size_t e = 0; e < nnz; e++) {
        const real_t *b = B + (size_t)ixB[e] * ldb;
        real_t coef = dot(k, b, a);
        axpy(k, -(coef - 1) * Xa[e] - coef, b, r);
    }
    axpy(k, -lam, a, r);
    memcpy(p, r, k * sizeof(real_t));
    real_t r_old = dot(k, r, r);
    if (r_old <= 1e-12) return;
    for (int s = 0; s < steps; s++) {
        for (int i = 0; i < k; i++) { real_t t = 0; for (int j = 0; j < k; j++) t += BtB[(i <= j) ? i * k + j : j * k + i] * p[j]; Ap[i] = t; }
        for (size_t e = 0; e < nnz; e++) {
            const real_t *b = B + (size_t)ixB[e] * ldb;
            real_t coef = dot(k, b, p);
            axpy(k, coef * (Xa[e] - 1) + coef, b, Ap);
        }
        axpy(k, lam, p, Ap);
        real_t al = r_old / dot(k, Ap, p);
        axpy(k, al, p, a);
        axpy(k, -al, Ap, r);
        real_t r_new = dot(k, r, r);
        if (r_new <= 1e-8) break;
        real_t be = r_new / r_old;
        for (int i = 0; i < k; i++) p[i] = be * p[i] + r[i];
        r_old = r_new;
    }
}

/* factors_implicit_chol (BtB already carries +lam on its diagonal)          reference src/common.c:2063-2126 */
static void row_implicit_chol(real_t *a, int k, const real_t *B, size_t ldb, const real_t *Xa, const int *ixB, size_t nnz,
                              const real_t *BtB_lam, real_t *buf)
{
    real_t *S = buf;
    for (size_t e = 0; e < nnz; e++) axpy(k, Xa[e] + 1, B + (size_t)ixB[e] * ldb, a);
    memset(S, 0, (size_t)k * k * sizeof(real_t));
    for (size_t e = 0; e < nnz; e++) {
        const real_t *b = B + (size_t)ixB[e] * ldb;
        for (int i = 0; i < k; i++) for (int j = i; j < k; j++) S[i * k + j] += Xa[e] * b[i] * b[j];
    }
    for (int i = 0; i < k; i++) for (int j = i; j < k; j++) S[i * k + j] += BtB_lam[i * k + j];
    spd_solve(k, S, a);
}

/* ---------------------------------------------------------------------------------------------------------
 * half-sweeps
 * ------------------------------------------------------------------------------------------------------- */
/* optimizeA, "Case 4" (sparse X, missing = unknown, no weights)                reference src/common.c:3209-3302
 * + the scale_lam rule of factors_closed_form                                  reference src/common.c:679-723 */
void oracle_optimizeA(real_t *A, int lda, const real_t *B, int ldb, int m, int k, const size_t *Xp, const int *Xi,
                      const real_t *Xv, real_t lam, real_t lam_last, bool scale_lam, bool use_cg, int max_cg_steps)
{
    real_t *buf = (real_t *)malloc(((size_t)k * k + 3 * (size_t)k) * sizeof(real_t));
    for (int r = 0; r < m; r++) {
        size_t nnz = Xp[r + 1] - Xp[r];
        if (!nnz) continue;                                       /* rows without entries keep their value */
        real_t l = lam, ll = lam_last;
        if (scale_lam) { l *= (real_t)nnz; ll *= (real_t)nnz; }
        if (use_cg) row_explicit_cg(A + (size_t)r * lda, k, B, ldb, Xv + Xp[r], Xi + Xp[r], nnz, l, ll, max_cg_steps, buf);
        else row_explicit_chol(A + (size_t)r * lda, k, B, ldb, Xv + Xp[r], Xi + Xp[r], nnz, l, ll, buf);
    }
    free(buf);
}

/* optimizeA_implicit                                                           reference src/common.c:3305-3421 */
void oracle_optimizeA_implicit(real_t *A, size_t lda, const real_t *B, size_t ldb, int m, int n, int k, const size_t *Xp,
                               const int *Xi, const real_t *Xv, real_t lam, bool use_cg, int max_cg_steps)
{
    real_t *BtB = (real_t *)calloc((size_t)k * k, sizeof(real_t));
    real_t *buf = (real_t *)malloc(((size_t)k * k + 3 * (size_t)k) * sizeof(real_t));
    for (int r = 0; r < n; r++) {
        const real_t *b = B + (size_t)r * ldb;
        for (int i = 0; i < k; i++) for (int j = i; j < k; j++) BtB[i * k + j] += b[i] * b[j];
    }
    if (!use_cg) {
        for (int i = 0; i < k; i++) BtB[i * k + i] += lam;
        for (int r = 0; r < m; r++) memset(A + (size_t)r * lda, 0, k * sizeof(real_t));
    }
    for (int r = 0; r < m; r++) {
        size_t nnz = Xp[r + 1] - Xp[r];
        if (!nnz) continue;
        if (use_cg) row_implicit_cg(A + (size_t)r * lda, k, B, ldb, Xv + Xp[r], Xi + Xp[r], nnz, lam, BtB, max_cg_steps, buf);
        else row_implicit_chol(A + (size_t)r * lda, k, B, ldb, Xv + Xp[r], Xi + Xp[r], nnz, BtB, buf);
    }
    free(buf); free(BtB);
}

/* ---------------------------------------------------------------------------------------------------------
 * whole fits (no side information)
 * ------------------------------------------------------------------------------------------------------- */
/* fit_collective_explicit_als: centring :7555, CSR/CSC :7593, bias init :8164-8226, random init :8241-8274,
 * bias columns :8283-8317, loop :8334-8898 (B then A, bias column juggling :8538-8546, 8723-8736, re-centring
 * :8566-8571, 8750-8755), copy back :8908-8933.  w_main folding :7497-7521.   reference src/collective.c */
int oracle_fit_explicit(real_t *biasA, real_t *biasB, real_t *A, real_t *B, int seed, real_t *glob_mean, int m, int n, int k,
                        const int *ixA, const int *ixB, const real_t *X, size_t nnz, bool user_bias, bool item_bias,
                        bool center, real_t lam, const real_t *lam_unique, bool scale_lam, real_t w_main, int niter,
                        int nthreads, bool use_cg, int max_cg_steps, bool finalize_chol)
{
    real_t lu[6];
    for (int i = 0; i < 6; i++) lu[i] = lam_unique ? lam_unique[i] : lam;
    if (w_main != 1) for (int i = 0; i < 6; i++) lu[i] /= w_main;
    if (!use_cg) finalize_chol = false;
    bool has_bias = user_bias || item_bias;
    int ld = k + (has_bias ? 1 : 0);

    real_t *Xc = (real_t *)malloc(nnz * sizeof(real_t));
    memcpy(Xc, X, nnz * sizeof(real_t));
    real_t mu = 0;
    if (center) { mu = oracle_global_mean(X, nnz, nthreads); if (mu != 0) for (size_t e = 0; e < nnz; e++) Xc[e] -= mu; }
    *glob_mean = mu;
    size_t *rp = (size_t *)malloc(((size_t)m + 1) * sizeof(size_t)), *cp = (size_t *)malloc(((size_t)n + 1) * sizeof(size_t));
    int *ri = (int *)malloc(nnz * sizeof(int)), *ci = (int *)malloc(nnz * sizeof(int));
    real_t *rv0 = (real_t *)malloc(nnz * sizeof(real_t)), *cv0 = (real_t *)malloc(nnz * sizeof(real_t));
    real_t *rv = (real_t *)malloc(nnz * sizeof(real_t)), *cv = (real_t *)malloc(nnz * sizeof(real_t));
    oracle_coo_to_csr(ixA, ixB, Xc, m, nnz, rp, ri, rv0);
    oracle_coo_to_csr(ixB, ixA, Xc, n, nnz, cp, ci, cv0);
    memcpy(rv, rv0, nnz * sizeof(real_t)); memcpy(cv, cv0, nnz * sizeof(real_t));
    free(Xc);

    if (user_bias && item_bias) oracle_init_biases_twosided(m, n, rp, ri, rv0, cp, ci, cv0, lu[0], lu[1], scale_lam, biasA, biasB);
    else if (user_bias) oracle_init_biases_onesided(m, rp, rv0, lu[0], scale_lam, biasA);
    else if (item_bias && use_cg) oracle_init_biases_onesided(n, cp, cv0, lu[1], scale_lam, biasB);

    oracle_random_init(A, (size_t)m * k, NULL, 0, seed, true);
    if (use_cg) memset(B, 0, (size_t)n * k * sizeof(real_t));

    real_t *Ab = (real_t *)malloc((size_t)m * ld * sizeof(real_t)), *Bb = (real_t *)malloc((size_t)n * ld * sizeof(real_t));
    for (int r = 0; r < m; r++) { memcpy(Ab + (size_t)r * ld, A + (size_t)r * k, k * sizeof(real_t)); if (has_bias) Ab[(size_t)r * ld + k] = user_bias ? biasA[r] : 1; }
    for (int r = 0; r < n; r++) { memcpy(Bb + (size_t)r * ld, B + (size_t)r * k, k * sizeof(real_t)); if (has_bias) Bb[(size_t)r * ld + k] = item_bias ? biasB[r] : 1; }

    for (int it = 0; it < niter; it++) {
        if (it == niter - 1 && use_cg && finalize_chol) use_cg = false;
        if (item_bias) for (int r = 0; r < m; r++) Ab[(size_t)r * ld + k] = 1;
        if (user_bias) for (size_t e = 0; e < nnz; e++) cv[e] = cv0[e] - biasA[ci[e]];
        oracle_optimizeA(Bb, ld, Ab, ld, n, k + (item_bias ? 1 : 0), cp, ci, cv, lu[3], lu[item_bias ? 1 : 3], scale_lam, use_cg, max_cg_steps);
        if (item_bias) for (int r = 0; r < n; r++) biasB[r] = Bb[(size_t)r * ld + k];
        if (user_bias) for (int r = 0; r < n; r++) Bb[(size_t)r * ld + k] = 1;
        if (item_bias) for (size_t e = 0; e < nnz; e++) rv[e] = rv0[e] - biasB[ri[e]];
        oracle_optimizeA(Ab, ld, Bb, ld, m, k + (user_bias ? 1 : 0), rp, ri, rv, lu[2], lu[user_bias ? 0 : 2], scale_lam, use_cg, max_cg_steps);
        if (user_bias) for (int r = 0; r < m; r++) biasA[r] = Ab[(size_t)r * ld + k];
    }
    for (int r = 0; r < m; r++) memcpy(A + (size_t)r * k, Ab + (size_t)r * ld, k * sizeof(real_t));
    for (int r = 0; r < n; r++) memcpy(B + (size_t)r * k, Bb + (size_t)r * ld, k * sizeof(real_t));
    free(Ab); free(Bb); free(rp); free(cp); free(ri); free(ci); free(rv0); free(cv0); free(rv); free(cv);
    return 0;
}

/* fit_collective_implicit_als: value transform :9578-9599, CSR/CSC :9600, init :9750-9774, weights :9776-9811,
 * loop :9827-10040 (B then A).                                                    reference src/collective.c */
int oracle_fit_implicit(real_t *A, real_t *B, int seed, int m, int n, int k, const int *ixA, const int *ixB, const real_t *X,
                        size_t nnz, real_t lam, real_t w_main, real_t *w_main_multiplier, real_t alpha, bool adjust_weight,
                        bool apply_log_transf, int niter, bool use_cg, int max_cg_steps, bool finalize_chol)
{
    if (!use_cg) finalize_chol = false;
    real_t *Xs = (real_t *)malloc(nnz * sizeof(real_t));
    memcpy(Xs, X, nnz * sizeof(real_t));
#ifdef USE_FLOAT
    if (apply_log_transf) for (size_t e = 0; e < nnz; e++) Xs[e] = logf(Xs[e]);
#else
    if (apply_log_transf) for (size_t e = 0; e < nnz; e++) Xs[e] = log(Xs[e]);
#endif
    if (alpha != 1) for (size_t e = 0; e < nnz; e++) Xs[e] *= alpha;
    size_t *rp = (size_t *)malloc(((size_t)m + 1) * sizeof(size_t)), *cp = (size_t *)malloc(((size_t)n + 1) * sizeof(size_t));
    int *ri = (int *)malloc(nnz * sizeof(int)), *ci = (int *)malloc(nnz * sizeof(int));
    real_t *rv = (real_t *)malloc(nnz * sizeof(real_t)), *cv = (real_t *)malloc(nnz * sizeof(real_t));
    oracle_coo_to_csr(ixA, ixB, Xs, m, nnz, rp, ri, rv);
    oracle_coo_to_csr(ixB, ixA, Xs, n, nnz, cp, ci, cv);
    free(Xs);
    oracle_random_init(A, (size_t)m * k, NULL, 0, seed, false);
    if (use_cg) memset(B, 0, (size_t)n * k * sizeof(real_t));
    *w_main_multiplier = 1;
    if (adjust_weight) { *w_main_multiplier = (real_t)((long double)nnz / (long double)((size_t)m * (size_t)n)); w_main *= *w_main_multiplier; }
    if (w_main != 1) lam /= w_main;
    for (int it = 0; it < niter; it++) {
        if (it == niter - 1 && use_cg && finalize_chol) use_cg = false;
        oracle_optimizeA_implicit(B, k, A, k, n, m, k, cp, ci, cv, lam, use_cg, max_cg_steps);
        oracle_optimizeA_implicit(A, k, B, k, m, n, k, rp, ri, rv, lam, use_cg, max_cg_steps);
    }
    free(rp); free(cp); free(ri); free(ci); free(rv); free(cv);
    return 0;
}
