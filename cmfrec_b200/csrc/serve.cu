// Batched serving: predictions for (row, column) pairs and top-N for many users at once, factors resident in HBM.
//   reference predict_multiple  src/common.c:5066-5112   (dot product + biases + mean per pair, NaN for unknown ids)
//   reference topN              src/common.c:5127-5369   (scores = B a + biasB, partial argsort, one user per call)
//
// top-N for a batch of users:  scores[users x items] = A_batch B^T + biasB  (fp32: tcgen05 tensor cores, gemm_tc.cu;
// fp64: DFMA kernel below), seen items masked out from a CSR, then ONE THREAD BLOCK PER USER selects the n_top best
// with an exact radix select on the composite key (score, lower item id first) -- histogram passes over the row of
// scores (L2 / HBM-bound), the survivors sorted in shared memory.  The ranking is a total order, so the result does not
// depend on thread scheduling.
#include "serve.h"
#include "gemm_tc.h"
#include "als.h"
#include <cmath>
#include <cstdio>
#include <limits>
#include <algorithm>
#include <vector>

namespace cmfb200 {

struct ServeState {
    int_t m = 0, n = 0, k_user = 0, k_item = 0, k = 0, k_main = 0;
    int lda = 0, ldb = 0, k_pred = 0;
    real_t glob_mean = 0;
    DevBuf<real_t> A, B, biasA, biasB;
    bool has_biasA = false, has_biasB = false;
    cudaStream_t stream = nullptr;
};

namespace {

constexpr int kSelThreads = 256;
constexpr int kSelMaxTop = 2048;

template <typename T>
__global__ void predict_pairs_kernel(const T *__restrict__ A, int lda, const T *__restrict__ B, int ldb, const T *__restrict__ biasA,
                                     const T *__restrict__ biasB, T glob_mean, int k_pred, int_t m, int_t n, const int_t *__restrict__ row,
                                     const int_t *__restrict__ col, size_t n_predict, T *__restrict__ out)
{
    // 8 lanes per pair: every load instruction of a group covers a contiguous run of the two factor rows
    const size_t pair = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int l = threadIdx.x & 7;
    const bool live = pair < n_predict;
    int_t r = 0, c = 0;
    if (live) {
        r = row[pair];
        c = col[pair];
    }
    const bool ok = live && r >= 0 && r < m && c >= 0 && c < n;
    T s = T(0);
    if (ok) {
        const T *a = A + (size_t)r * lda, *b = B + (size_t)c * ldb;
        for (int j = l; j < k_pred; j += 8) s = fma(a[j], b[j], s);
    }
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (live && l == 0)
        out[pair] = ok ? s + (biasA ? biasA[r] : T(0)) + (biasB ? biasB[c] : T(0)) + glob_mean : std::numeric_limits<T>::quiet_NaN();
}

// fp64 (and fallback) scores: one thread per item, 8 users per block share every loaded row of B
template <typename T>
__global__ void scores_fma_kernel(const T *__restrict__ A, int lda, const int_t *__restrict__ users, int_t n_users, const T *__restrict__ B,
                                  int ldb, int_t n, int k_pred, const T *__restrict__ biasB, T *__restrict__ S)
{
    constexpr int UB = 8;
    extern __shared__ unsigned char serve_smem[];
    T *as = reinterpret_cast<T *>(serve_smem);   // [UB][k_pred]
    const int_t u0 = blockIdx.y * UB;
    for (int i = threadIdx.x; i < UB * k_pred; i += blockDim.x) {
        const int uu = i / k_pred, j = i % k_pred;
        as[i] = (u0 + uu < n_users) ? A[(size_t)users[u0 + uu] * lda + j] : T(0);
    }
    __syncthreads();
    const int_t item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n) return;
    T acc[UB];
#pragma unroll
    for (int uu = 0; uu < UB; uu++) acc[uu] = T(0);
    const T *b = B + (size_t)item * ldb;
    for (int j = 0; j < k_pred; j++) {
        const T bv = b[j];
#pragma unroll
        for (int uu = 0; uu < UB; uu++) acc[uu] = fma(as[uu * k_pred + j], bv, acc[uu]);
    }
    const T bb = biasB ? biasB[item] : T(0);
#pragma unroll
    for (int uu = 0; uu < UB; uu++)
        if (u0 + uu < n_users) S[(size_t)(u0 + uu) * n + item] = acc[uu] + bb;
}

// the listed users' factor rows, contiguous (the A operand of the tensor-core product)
template <typename T>
__global__ void gather_rows_kernel(const T *__restrict__ A, int lda, const int_t *__restrict__ users, int_t n_users, int k_pred, T *__restrict__ out,
                                   int ldo)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_users * ldo) return;
    const int_t u = (int_t)(i / ldo);
    const int j = (int)(i % ldo);
    out[i] = j < k_pred ? A[(size_t)users[u] * lda + j] : T(0);
}

template <typename T>
__global__ void mask_seen_kernel(T *__restrict__ S, int_t n, const size_t *__restrict__ seen_ptr, const int_t *__restrict__ seen_idx, int_t u_first,
                                 int_t n_users)
{
    // one warp per user of the chunk
    const int_t u = (int_t)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (u >= n_users) return;
    const size_t base0 = seen_ptr[0];   // seen_idx holds the listed users' entries only, starting with the first user's
    for (size_t e = seen_ptr[u_first + u] + lane; e < seen_ptr[u_first + u + 1]; e += 32) {
        const int_t it = seen_idx[e - base0];
        if (it >= 0 && it < n) S[(size_t)u * n + it] = -std::numeric_limits<T>::infinity();
    }
}

// order-preserving map of a floating-point score to an unsigned key (larger key = larger score)
__device__ __forceinline__ uint64_t sortable(float x)
{
    const uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? (uint32_t)~b : (b | 0x80000000u);
}
__device__ __forceinline__ uint64_t sortable(double x)
{
    const uint64_t b = (uint64_t)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// One block per user: the n_top largest (score, lower id first) of S[user][0..n), sorted, with `shift` added to the scores.
template <typename T>
__global__ void __launch_bounds__(kSelThreads)
topn_select_kernel(const T *__restrict__ S, int_t n, int n_top, const T *__restrict__ biasA, const int_t *__restrict__ users, T glob_mean,
                   int_t *__restrict__ out_ix, T *__restrict__ out_score)
{
    constexpr int SB = (int)sizeof(T);           // score bytes; the composite key has SB + 4 bytes (item id last)
    constexpr int NPASS = SB + 4;
    __shared__ int hist[256];
    __shared__ int sel_digit, sel_remaining, n_cand;
    __shared__ uint64_t cand_key[kSelMaxTop];
    __shared__ uint32_t cand_id[kSelMaxTop];
    const int tid = threadIdx.x, lane = tid & 31;
    const T *row = S + (size_t)blockIdx.x * n;

    uint64_t pref_s = 0, mask_s = 0;             // bits of the score key fixed so far
    uint32_t pref_i = 0, mask_i = 0;             // bits of ~id fixed so far
    int remaining = n_top;                       // how many are still to be taken among the elements matching the prefix
    bool done = false;                           // every element matching the prefix is taken
    for (int pass = 0; pass < NPASS && !done; pass++) {
        for (int i = tid; i < 256; i += kSelThreads) hist[i] = 0;
        __syncthreads();
        const bool in_score = pass < SB;
        const int shift = in_score ? 8 * (SB - 1 - pass) : 8 * (3 - (pass - SB));
        for (int_t i0 = 0; i0 < n; i0 += kSelThreads) {
            const int_t i = i0 + tid;
            bool match = false;
            int digit = 0;
            if (i < n) {
                const uint64_t ks = sortable(row[i]);
                const uint32_t ki = ~(uint32_t)i;
                match = ((ks ^ pref_s) & mask_s) == 0 && ((ki ^ pref_i) & mask_i) == 0;
                digit = in_score ? (int)((ks >> shift) & 255u) : (int)((ki >> shift) & 255u);
            }
            // warp-aggregated histogram update: scores share their leading bytes, plain atomics would serialise
            const unsigned act = __ballot_sync(0xffffffffu, match);
            if (match) {
                const unsigned same = __match_any_sync(act, digit);
                if (lane == __ffs(same) - 1) atomicAdd(&hist[digit], __popc(same));
            }
        }
        __syncthreads();
        if (tid == 0) {
            int cum = 0, d = 255;
            for (; d > 0; d--) {
                if (cum + hist[d] >= remaining) break;
                cum += hist[d];
            }
            sel_digit = d;
            sel_remaining = remaining - cum;     // to be taken from bin d
            n_cand = 0;
        }
        __syncthreads();
        const int d = sel_digit;
        remaining = sel_remaining;
        done = hist[d] == remaining;
        if (in_score) {
            pref_s |= (uint64_t)d << shift;
            mask_s |= (uint64_t)255u << shift;
        } else {
            pref_i |= (uint32_t)d << shift;
            mask_i |= 255u << shift;
        }
        __syncthreads();
    }
    // ---- collect: every element whose key is >= the prefix on the fixed bits
    if (tid == 0) n_cand = 0;
    __syncthreads();
    for (int_t i0 = 0; i0 < n; i0 += kSelThreads) {
        const int_t i = i0 + tid;
        if (i < n) {
            const uint64_t ks = sortable(row[i]);
            const uint32_t ki = ~(uint32_t)i;
            const uint64_t a = ks & mask_s, b = pref_s & mask_s;
            const bool take = a > b || (a == b && (ki & mask_i) >= (pref_i & mask_i));
            if (take) {
                const int slot = atomicAdd(&n_cand, 1);
                if (slot < kSelMaxTop) {
                    cand_key[slot] = ks;
                    cand_id[slot] = (uint32_t)i;
                }
            }
        }
    }
    __syncthreads();
    const int nc = min(n_cand, kSelMaxTop);      // == n_top by construction
    // ---- sort the survivors: rank by counting (n_top is small), descending key, ascending id
    for (int c = tid; c < nc; c += kSelThreads) {
        const uint64_t kc = cand_key[c];
        const uint32_t ic = cand_id[c];
        int rank = 0;
        for (int o = 0; o < nc; o++) {
            const uint64_t ko = cand_key[o];
            rank += (ko > kc || (ko == kc && cand_id[o] < ic)) ? 1 : 0;
        }
        if (rank < n_top) {
            out_ix[(size_t)blockIdx.x * n_top + rank] = (int_t)ic;
            if (out_score) {
                const T shift_score = glob_mean + (biasA ? biasA[users[blockIdx.x]] : T(0));
                out_score[(size_t)blockIdx.x * n_top + rank] = row[ic] + shift_score;
            }
        }
    }
}

bool have_device()
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        std::fprintf(stderr, "cmfrec_b200: no CUDA device available; this library has no CPU path.\n");
        return false;
    }
    return true;
}

}  // namespace

ServeState *serve_create(const real_t *A, int_t m, int_t k_user, const real_t *B, int_t n, int_t k_item, const real_t *biasA,
                         const real_t *biasB, real_t glob_mean, int_t k, int_t k_main, int *rc)
{
    *rc = 0;
    if (!have_device()) { *rc = 1; return nullptr; }
    if (m < 0 || n < 1 || k + k_main < 1 || (m > 0 && !A) || !B) { *rc = 2; return nullptr; }
    ServeState *s = new ServeState();
    s->m = m; s->n = n; s->k_user = k_user; s->k_item = k_item; s->k = k; s->k_main = k_main;
    s->lda = k_user + k + k_main; s->ldb = k_item + k + k_main; s->k_pred = k + k_main;
    s->glob_mean = glob_mean;
    bool ok = s->A.alloc((size_t)m * s->lda + 4) && s->B.alloc((size_t)n * s->ldb + 4);
    if (ok && m > 0) ok = cudaMemcpy(s->A.p, A, (size_t)m * s->lda * sizeof(real_t), cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok) ok = cudaMemcpy(s->B.p, B, (size_t)n * s->ldb * sizeof(real_t), cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok && biasA && m > 0) {
        ok = s->biasA.alloc(m) && cudaMemcpy(s->biasA.p, biasA, (size_t)m * sizeof(real_t), cudaMemcpyHostToDevice) == cudaSuccess;
        s->has_biasA = true;
    }
    if (ok && biasB) {
        ok = s->biasB.alloc(n) && cudaMemcpy(s->biasB.p, biasB, (size_t)n * sizeof(real_t), cudaMemcpyHostToDevice) == cudaSuccess;
        s->has_biasB = true;
    }
    if (!ok) {
        cudaGetLastError();
        delete s;
        *rc = 1;
        return nullptr;
    }
    return s;
}

void serve_destroy(ServeState *s) { delete s; }

int serve_predict(ServeState *s, const int_t *row, const int_t *col, size_t n_predict, real_t *out)
{
    if (n_predict == 0) return 0;
    DevBuf<int_t> drow, dcol;
    DevBuf<real_t> dout;
    if (!drow.alloc(n_predict) || !dcol.alloc(n_predict) || !dout.alloc(n_predict)) return 1;
    cudaMemcpyAsync(drow.p, row, n_predict * sizeof(int_t), cudaMemcpyHostToDevice, s->stream);
    cudaMemcpyAsync(dcol.p, col, n_predict * sizeof(int_t), cudaMemcpyHostToDevice, s->stream);
    const int threads = 256;
    const size_t blocks = (n_predict * 8 + threads - 1) / threads;
    predict_pairs_kernel<real_t><<<(unsigned)blocks, threads, 0, s->stream>>>(
        s->A.p + s->k_user, s->lda, s->B.p + s->k_item, s->ldb, s->has_biasA ? s->biasA.p : nullptr, s->has_biasB ? s->biasB.p : nullptr,
        s->glob_mean, s->k_pred, s->m, s->n, drow.p, dcol.p, n_predict, dout.p);
    if (cudaMemcpyAsync(out, dout.p, n_predict * sizeof(real_t), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) return 1;
    return cudaStreamSynchronize(s->stream) == cudaSuccess ? 0 : 1;
}

int serve_topn(ServeState *s, const int_t *users, int_t n_users, const size_t *seen_ptr, const int_t *seen_idx, int_t n_top,
               int_t *out_ix, real_t *out_score, float *ms_device)
{
    if (n_users < 1) return 0;
    if (n_top < 1 || n_top > s->n || n_top > kSelMaxTop || !users || !out_ix) return 2;
    for (int_t u = 0; u < n_users; u++) {
        if (users[u] < 0 || users[u] >= s->m) return 2;
        if (seen_ptr && (long long)(seen_ptr[u + 1] - seen_ptr[u]) > (long long)s->n - n_top) return 2;   // src/common.c:5148
    }
    const int_t n = s->n;
    // users per chunk: at most 2^26 scores (256 MB in fp32) in flight
    int_t chunk = (int_t)std::max<long long>(1, std::min<long long>(n_users, (1ll << 26) / n));
    const int ldg = ((s->k_pred + 3) / 4) * 4;
    DevBuf<real_t> dS, dAg, dscore;
    DevBuf<int_t> dusers, dix, dseen;
    DevBuf<size_t> dptr;
    if (!dS.alloc((size_t)chunk * n) || !dAg.alloc((size_t)chunk * ldg + 4) || !dusers.alloc(n_users) || !dix.alloc((size_t)chunk * n_top) ||
        (out_score && !dscore.alloc((size_t)chunk * n_top)))
        return 1;
    cudaMemcpyAsync(dusers.p, users, (size_t)n_users * sizeof(int_t), cudaMemcpyHostToDevice, s->stream);
    if (seen_ptr) {
        if (!dptr.alloc((size_t)n_users + 1)) return 1;
        cudaMemcpyAsync(dptr.p, seen_ptr, ((size_t)n_users + 1) * sizeof(size_t), cudaMemcpyHostToDevice, s->stream);
        const size_t total = seen_ptr[n_users] - seen_ptr[0];
        if (total > 0) {
            if (!seen_idx || !dseen.alloc(total)) return 1;
            cudaMemcpyAsync(dseen.p, seen_idx + seen_ptr[0], total * sizeof(int_t), cudaMemcpyHostToDevice, s->stream);
        }
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms_device) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, s->stream);
    }
    const real_t *biasB = s->has_biasB ? s->biasB.p : nullptr;
    for (int_t u0 = 0; u0 < n_users; u0 += chunk) {
        const int_t nu = std::min<int_t>(chunk, n_users - u0);
        int rc = 3;
#ifdef USE_FLOAT
        {
            const size_t tot = (size_t)nu * ldg;
            gather_rows_kernel<real_t><<<(unsigned)((tot + 255) / 256), 256, 0, s->stream>>>(s->A.p + s->k_user, s->lda, dusers.p + u0, nu, s->k_pred,
                                                                                             dAg.p, ldg);
            rc = launch_gemm_nt_tc(dAg.p, ldg, nu, s->B.p + s->k_item, s->ldb, n, s->k_pred, dS.p, n, nullptr, biasB, real_t(0), s->stream);
        }
#endif
        if (rc == 3) {
            const dim3 grid((unsigned)((n + 255) / 256), (unsigned)((nu + 7) / 8));
            scores_fma_kernel<real_t><<<grid, 256, (size_t)8 * s->k_pred * sizeof(real_t), s->stream>>>(
                s->A.p + s->k_user, s->lda, dusers.p + u0, nu, s->B.p + s->k_item, s->ldb, n, s->k_pred, biasB, dS.p);
            rc = cudaGetLastError() == cudaSuccess ? 0 : 1;
        }
        if (rc) return rc;
        if (seen_ptr && dseen.p)
            mask_seen_kernel<real_t><<<(unsigned)(((size_t)nu * 32 + 255) / 256), 256, 0, s->stream>>>(dS.p, n, dptr.p, dseen.p, u0, nu);
        topn_select_kernel<real_t><<<(unsigned)nu, kSelThreads, 0, s->stream>>>(dS.p, n, n_top, s->has_biasA ? s->biasA.p : nullptr, dusers.p + u0,
                                                                               s->glob_mean, dix.p, out_score ? dscore.p : nullptr);
        if (cudaMemcpyAsync(out_ix + (size_t)u0 * n_top, dix.p, (size_t)nu * n_top * sizeof(int_t), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess)
            return 1;
        if (out_score &&
            cudaMemcpyAsync(out_score + (size_t)u0 * n_top, dscore.p, (size_t)nu * n_top * sizeof(real_t), cudaMemcpyDeviceToHost, s->stream) !=
                cudaSuccess)
            return 1;
        if (cudaStreamSynchronize(s->stream) != cudaSuccess) return 1;   // dix / dscore are reused by the next chunk
    }
    if (ms_device) {
        cudaEventRecord(e1, s->stream);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(ms_device, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int predict_multiple_host(real_t *A, int_t k_user, real_t *B, int_t k_item, real_t *biasA, real_t *biasB, real_t glob_mean, int_t k,
                          int_t k_main, int_t m, int_t n, int_t *predA, int_t *predB, size_t nnz, real_t *outp)
{
    if (nnz == 0) return 0;
    // m == 0 / n == 0 mean "unknown, trust the indices" in the reference (src/common.c:5088-5089): size the upload by the largest id
    int_t mm = m, nn = n;
    if (mm == 0 || nn == 0) {
        for (size_t i = 0; i < nnz; i++) {
            if (m == 0 && predA[i] + 1 > mm) mm = predA[i] + 1;
            if (n == 0 && predB[i] + 1 > nn) nn = predB[i] + 1;
        }
    }
    int rc = 0;
    if (2 * nnz < (size_t)mm + (size_t)nn) {
        // few pairs against big matrices: ship only the rows that are asked for (pair i -> packed row i of both sides);
        // unknown ids keep their place as NaN
        const int lda = k_user + k + k_main, ldb = k_item + k + k_main;
        std::vector<real_t> pa(nnz * (size_t)lda, real_t(0)), pb(nnz * (size_t)ldb, real_t(0)), ba, bb;
        std::vector<int_t> ia(nnz), ib(nnz);
        if (biasA) ba.assign(nnz, real_t(0));
        if (biasB) bb.assign(nnz, real_t(0));
        for (size_t i = 0; i < nnz; i++) {
            const bool ok = predA[i] >= 0 && predA[i] < mm && predB[i] >= 0 && predB[i] < nn;
            ia[i] = ib[i] = ok ? (int_t)i : -1;
            if (!ok) continue;
            std::copy(A + (size_t)predA[i] * lda, A + (size_t)predA[i] * lda + lda, pa.begin() + i * lda);
            std::copy(B + (size_t)predB[i] * ldb, B + (size_t)predB[i] * ldb + ldb, pb.begin() + i * ldb);
            if (biasA) ba[i] = biasA[predA[i]];
            if (biasB) bb[i] = biasB[predB[i]];
        }
        ServeState *s = serve_create(pa.data(), (int_t)nnz, k_user, pb.data(), (int_t)nnz, k_item, biasA ? ba.data() : nullptr,
                                     biasB ? bb.data() : nullptr, glob_mean, k, k_main, &rc);
        if (!s) return rc;
        rc = serve_predict(s, ia.data(), ib.data(), nnz, outp);
        serve_destroy(s);
        return rc;
    }
    ServeState *s = serve_create(A, mm, k_user, B, nn, k_item, biasA, biasB, glob_mean, k, k_main, &rc);
    if (!s) return rc;
    rc = serve_predict(s, predA, predB, nnz, outp);
    serve_destroy(s);
    return rc;
}

}  // namespace cmfb200
