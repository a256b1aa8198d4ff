// fit_most_popular on the GPU: the bias-only ("most popular") model.
//   reference fit_most_popular            src/common.c:5371-5699
//   reference fit_most_popular_internal   src/common.c:5703-6102   (alternating closed-form bias updates,
//                                          6 rounds item-then-user; implicit: one closed form per item)
//   reference initialize_biases           src/common.c:3651-4127   (centring; item-only model :3923-3958)
//
// The reference accumulates every bias as a running sum over the COO entries in input order, in real_t.
// Here X is compressed by row and by column on the host (stable, so the entries of one row/column keep their
// COO order) and one GPU thread walks one row/column sequentially with the same operations in the same order,
// so float results are reproduced exactly; the final division is done in double like the reference does.
// Supported: sparse COO input without observation weights and without NA_as_zero; anything else is refused.
#include "popular.h"
#include "als.h"
#include "host_prep.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace cmfb200 {

namespace {

// out[r] = ( sum_t (val[t] - other[idx[t]]) ) / (cnt + lam * (scale ? cnt : 1)),  NaN -> 0, optional clamp at 0
template <typename T>
__global__ void bias_update_kernel(int_t rows, const size_t *__restrict__ ptr, const int_t *__restrict__ idx,
                                   const T *__restrict__ val, const T *__restrict__ other, T lam, bool scale_lam,
                                   bool nonneg, T *__restrict__ out)
{
    const int_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const size_t b = ptr[r], e = ptr[r + 1];
    T acc = T(0);
    if (other) {
        for (size_t t = b; t < e; t++) acc = acc + (val[t] - other[idx[t]]);
    } else {
        for (size_t t = b; t < e; t++) acc = acc + val[t];
    }
    const double cnt = (double)(e - b);
    const double denom = __dadd_rn(cnt, __dmul_rn((double)lam, scale_lam ? cnt : 1.0));
    T res = (T)__ddiv_rn((double)acc, denom);
    if (isnan(res)) res = T(0);
    if (nonneg && !(res >= T(0))) res = T(0);
    out[r] = res;
}

// implicit, no user bias: S = sum_t (x_t + 1);  bias = alpha*S / (alpha*S + (m - cnt) + lam)
template <typename T>
__global__ void implicit_popularity_kernel(int_t cols, const size_t *__restrict__ ptr, const T *__restrict__ val,
                                           T alpha, T lam, int_t m, T *__restrict__ out)
{
    const int_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const size_t b = ptr[c], e = ptr[c + 1];
    T s = T(0);
    for (size_t t = b; t < e; t++) s = s + (val[t] + T(1));
    const T as = alpha * s;
    const double denom = __dadd_rn(__dadd_rn((double)as, (double)(m - (int_t)(e - b))), (double)lam);
    out[c] = (T)__ddiv_rn((double)as, denom);
}

struct DevCsr {
    DevBuf<size_t> p;
    DevBuf<int_t> i;
    DevBuf<real_t> v;
    bool upload(const std::vector<size_t> &hp, const std::vector<int_t> &hi, const std::vector<real_t> &hv)
    {
        if (!p.alloc(hp.size()) || !i.alloc(hi.size() ? hi.size() : 1) || !v.alloc(hv.size() ? hv.size() : 1)) return false;
        cudaMemcpy(p.p, hp.data(), hp.size() * sizeof(size_t), cudaMemcpyHostToDevice);
        if (!hi.empty()) cudaMemcpy(i.p, hi.data(), hi.size() * sizeof(int_t), cudaMemcpyHostToDevice);
        if (!hv.empty()) cudaMemcpy(v.p, hv.data(), hv.size() * sizeof(real_t), cudaMemcpyHostToDevice);
        return cudaGetLastError() == cudaSuccess;
    }
};

int refuse(const char *what)
{
    std::fprintf(stderr, "cmfrec_b200: fit_most_popular: %s is not supported by the GPU path (no CPU fallback).\n", what);
    return 2;
}

}  // namespace

int most_popular(real_t *biasA, real_t *biasB, real_t *glob_mean, real_t lam_user, real_t lam_item, bool scale_lam,
                 bool scale_bias_const, real_t alpha, int_t m, int_t n, int_t *ixA, int_t *ixB, real_t *X, size_t nnz,
                 real_t *Xfull, real_t *weight, bool implicit, bool adjust_weight, bool apply_log_transf, bool nonneg,
                 bool NA_as_zero, real_t *w_main_multiplier, int nthreads)
{
    // argument normalisation exactly as the reference (src/common.c:5389-5437)
    if (implicit) {
        NA_as_zero = false;
        scale_lam = false;
        weight = nullptr;
        if (Xfull) return 2;
    } else {
        adjust_weight = false;
        apply_log_transf = false;
    }
    if (!scale_lam) scale_bias_const = false;
    if (Xfull) return refuse("dense X");
    if (weight) return refuse("observation weights");
    if (NA_as_zero) return refuse("NA_as_zero");
    if (implicit && biasA) return refuse("implicit feedback with user biases");
    if (!implicit && !biasA) return refuse("explicit feedback without user biases");
    if (scale_bias_const) return refuse("scale_bias_const");
    if (!biasB || m < 1 || n < 1) return 2;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        std::fprintf(stderr, "cmfrec_b200: no CUDA device available; this library has no CPU path.\n");
        return 1;
    }
    if (w_main_multiplier) *w_main_multiplier = 1;
    if (glob_mean) *glob_mean = 0;

    std::vector<real_t> Xw(X, X + nnz);
    const int threads = 128;

    if (implicit) {
        // src/common.c:5733-5808
        if (apply_log_transf)
            for (size_t e = 0; e < nnz; e++) Xw[e] = std::log(Xw[e]);
        std::vector<size_t> cp((size_t)n + 1), rp((size_t)m + 1);
        std::vector<int_t> ci(nnz), ri(nnz);
        std::vector<real_t> cv(nnz), rv(nnz);
        coo_to_csr_and_csc(ixA, ixB, Xw.data(), m, n, nnz, rp.data(), ri.data(), rv.data(), cp.data(), ci.data(), cv.data());
        if (adjust_weight) {
            const real_t mult = (real_t)((long double)nnz / ((long double)m * (long double)n));
            if (w_main_multiplier) *w_main_multiplier = mult;
            lam_item /= mult;
        }
        DevCsr csc;
        DevBuf<real_t> dB;
        if (!csc.upload(cp, ci, cv) || !dB.alloc(n)) return 1;
        implicit_popularity_kernel<real_t><<<(n + threads - 1) / threads, threads>>>(n, csc.p.p, csc.v.p, alpha, lam_item, m, dB.p);
        if (cudaMemcpy(biasB, dB.p, (size_t)n * sizeof(real_t), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
        return 0;
    }

    // explicit feedback, user + item biases (src/common.c:5810-5841 -> initialize_biases: centring only;
    // then the alternating updates of src/common.c:5931-6087, sparse unweighted branch :6037-6085)
    real_t mu = 0;
    if (glob_mean) {
        mu = global_mean(X, nnz, nthreads);
        *glob_mean = mu;
        if (mu != 0)
            for (size_t e = 0; e < nnz; e++) Xw[e] -= mu;
    }
    std::vector<size_t> cp((size_t)n + 1), rp((size_t)m + 1);
    std::vector<int_t> ci(nnz), ri(nnz);
    std::vector<real_t> cv(nnz), rv(nnz);
    coo_to_csr_and_csc(ixA, ixB, Xw.data(), m, n, nnz, rp.data(), ri.data(), rv.data(), cp.data(), ci.data(), cv.data());
    DevCsr csr, csc;
    DevBuf<real_t> dA, dB;
    if (!csr.upload(rp, ri, rv) || !csc.upload(cp, ci, cv) || !dA.alloc(m) || !dB.alloc(n)) return 1;
    cudaMemset(dA.p, 0, (size_t)m * sizeof(real_t));
    cudaMemset(dB.p, 0, (size_t)n * sizeof(real_t));
    const int rounds = nonneg ? 15 : 5;
    for (int it = 0; it <= rounds; it++) {
        bias_update_kernel<real_t><<<(n + threads - 1) / threads, threads>>>(n, csc.p.p, csc.i.p, csc.v.p, dA.p, lam_item,
                                                                              scale_lam, nonneg, dB.p);
        bias_update_kernel<real_t><<<(m + threads - 1) / threads, threads>>>(m, csr.p.p, csr.i.p, csr.v.p, dB.p, lam_user,
                                                                              scale_lam, nonneg, dA.p);
    }
    if (cudaMemcpy(biasA, dA.p, (size_t)m * sizeof(real_t), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    if (cudaMemcpy(biasB, dB.p, (size_t)n * sizeof(real_t), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    return 0;
}

}  // namespace cmfb200
