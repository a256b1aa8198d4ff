// topN on the GPU: rank the items for one user.
//   reference topN  src/common.c:5127-5369  (scores = B a + biasB, then a partial argsort in decreasing order;
//   `include_ix` restricts the candidates, `exclude_ix` removes some; scores returned with glob_mean + biasA added)
//
// One warp scores one candidate item (coalesced read of its factor row, shuffle reduction); the candidates are
// then ordered by a stable device radix sort on the score keys (CUB, descending), so that among exactly equal
// scores the lower candidate position comes first.  Argument checks and return codes are the reference's.
#include "topn.h"
#include "als.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

namespace cmfb200 {

namespace {

template <typename T>
__global__ void score_candidates_kernel(const T *__restrict__ B, int ldb, const T *__restrict__ a, int k_pred,
                                        const T *__restrict__ biasB, const int_t *__restrict__ cand, int_t ncand,
                                        T *__restrict__ score, int_t *__restrict__ pos)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= ncand) return;
    const int_t item = cand ? cand[warp] : warp;
    const T *row = B + (size_t)item * ldb;
    T s = T(0);
    for (int c = lane; c < k_pred; c += 32) s = fma(row[c], a[c], s);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) {
        score[warp] = s + (biasB ? biasB[item] : T(0));
        pos[warp] = warp;
    }
}

}  // namespace

int top_n(real_t *a_vec, int_t k_user, real_t *B, int_t k_item, real_t *biasB, real_t glob_mean, real_t biasA, int_t k,
          int_t k_main, int_t *include_ix, int_t n_include, int_t *exclude_ix, int_t n_exclude, int_t *outp_ix,
          real_t *outp_score, int_t n_top, int_t n, int nthreads)
{
    (void)nthreads;
    // ---- argument checks (src/common.c:5140-5196)
    int retval = 0;
    if (include_ix != nullptr && exclude_ix != nullptr) retval = 2;
    if (n_top == 0) retval = 2;
    if (n_exclude > n - n_top) retval = 2;
    if (n_include > n) retval = 2;
    if (include_ix)
        for (int_t i = 0; i < n_include; i++)
            if (include_ix[i] < 0 || include_ix[i] >= n) { retval = 2; break; }
    if (exclude_ix)
        for (int_t i = 0; i < n_exclude; i++)
            if (exclude_ix[i] < 0 || exclude_ix[i] >= n) { retval = 2; break; }
    for (int_t i = 0; i < k_user + k + k_main; i++)
        if (std::isnan(a_vec[i])) { retval = 2; break; }
    if (std::isnan(biasA)) retval = 2;
    if (retval) return retval;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        std::fprintf(stderr, "cmfrec_b200: no CUDA device available; this library has no CPU path.\n");
        return 1;
    }

    const int k_pred = k + k_main;
    const int ldb = k_item + k + k_main;
    // candidate list: include_ix as given; otherwise all items (minus the excluded ones, in increasing id order)
    std::vector<int_t> cand_host;
    const int_t *cand = nullptr;
    int_t ncand = n;
    if (include_ix) {
        cand = include_ix;
        ncand = n_include;
    } else if (exclude_ix && n_exclude > 0) {
        std::vector<char> drop((size_t)n, 0);
        for (int_t i = 0; i < n_exclude; i++) drop[exclude_ix[i]] = 1;
        cand_host.reserve(n);
        for (int_t i = 0; i < n; i++)
            if (!drop[i]) cand_host.push_back(i);
        cand = cand_host.data();
        ncand = (int_t)cand_host.size();
    }
    if (ncand < n_top) return 2;

    DevBuf<real_t> dB, da, dbias, dscore, dscore_sorted;
    DevBuf<int_t> dcand, dpos, dpos_sorted;
    DevBuf<unsigned char> dtemp;
    if (!dB.alloc((size_t)n * ldb) || !da.alloc(k_pred) || !dscore.alloc(ncand) || !dscore_sorted.alloc(ncand) ||
        !dpos.alloc(ncand) || !dpos_sorted.alloc(ncand))
        return 1;
    cudaMemcpy(dB.p, B, (size_t)n * ldb * sizeof(real_t), cudaMemcpyHostToDevice);
    cudaMemcpy(da.p, a_vec + k_user, (size_t)k_pred * sizeof(real_t), cudaMemcpyHostToDevice);
    if (biasB) {
        if (!dbias.alloc(n)) return 1;
        cudaMemcpy(dbias.p, biasB, (size_t)n * sizeof(real_t), cudaMemcpyHostToDevice);
    }
    if (cand) {
        if (!dcand.alloc(ncand)) return 1;
        cudaMemcpy(dcand.p, cand, (size_t)ncand * sizeof(int_t), cudaMemcpyHostToDevice);
    }
    const int threads = 256;
    const long long blocks = ((long long)ncand * 32 + threads - 1) / threads;
    score_candidates_kernel<real_t><<<(unsigned)blocks, threads>>>(dB.p + k_item, ldb, da.p, k_pred, biasB ? dbias.p : nullptr,
                                                                   cand ? dcand.p : nullptr, ncand, dscore.p, dpos.p);
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, temp_bytes, dscore.p, dscore_sorted.p, dpos.p, dpos_sorted.p, ncand);
    if (!dtemp.alloc(temp_bytes ? temp_bytes : 1)) return 1;
    cub::DeviceRadixSort::SortPairsDescending(dtemp.p, temp_bytes, dscore.p, dscore_sorted.p, dpos.p, dpos_sorted.p, ncand);
    std::vector<int_t> top_pos(n_top);
    std::vector<real_t> top_score(n_top);
    if (cudaMemcpy(top_pos.data(), dpos_sorted.p, (size_t)n_top * sizeof(int_t), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    if (cudaMemcpy(top_score.data(), dscore_sorted.p, (size_t)n_top * sizeof(real_t), cudaMemcpyDeviceToHost) != cudaSuccess)
        return 1;
    const real_t shift = glob_mean + biasA;
    for (int_t i = 0; i < n_top; i++) {
        outp_ix[i] = cand ? cand[top_pos[i]] : top_pos[i];
        if (outp_score) outp_score[i] = top_score[i] + shift;
    }
    return 0;
}

}  // namespace cmfb200
