// Device-side ingestion for single-GPU fits: what the reference does on the host before the alternation starts,
// done in HBM so that only the raw COO triplets cross PCIe.
//
//   COO -> CSR / CSC     reference coo_to_csr_and_csc, src/helpers.c:1375-1491: a STABLE counting sort (entries of a
//                        row keep their COO order).  Here: stable LSD radix sort of (major index, position) pairs
//                        (CUB DeviceRadixSort), then a gather; row pointers from a histogram + exclusive scan.
//                        The integer outputs are identical to the reference's, entry for entry.
//   centring             x - glob_mean in real_t                         src/common.c:3632-3644
//   bias initialisation  initialize_biases_twosided / _onesided          src/common.c:4643-4669, 4799-4825, 4266-4289
//                        the reference's sequential chain (running mean in double, residual formed in real_t) in its
//                        order, so results are bit-identical: one warp per short row, one block per long row with
//                        producer warps feeding the chain through shared memory (bias_sweep_kernel).
#include "device_prep.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace cmfb200 {

namespace {

// NAME=0 in the environment switches an optional path off (a test and measurement aid)
bool env_flag_off(const char *name)
{
    const char *e = std::getenv(name);
    return e && std::atoi(e) == 0;
}

__global__ void iota_kernel(uint32_t *out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i;
}

__global__ void histogram_kernel(const int_t *__restrict__ major, size_t nnz, unsigned long long *__restrict__ counts)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) atomicAdd(&counts[major[i]], 1ULL);
}

// any index outside [0, limit) raises the flag (the compression below would write out of bounds / mis-sort it)
__global__ void index_range_kernel(const int_t *__restrict__ ix, size_t nnz, int_t limit, int *__restrict__ flag)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) {
        const int_t v = ix[i];
        if (v < 0 || v >= limit) *flag = 1;
    }
}

template <typename T>
__global__ void gather_kernel(const uint32_t *__restrict__ perm, const int_t *__restrict__ minor, const T *__restrict__ val,
                              size_t nnz, int_t *__restrict__ out_idx, T *__restrict__ out_val)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) {
        const uint32_t src = perm[i];
        out_idx[i] = minor[src];
        out_val[i] = val[src];
    }
}

template <typename T> __global__ void subtract_kernel(T *x, size_t n, T mu)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = x[i] - mu;
}

// one sweep of the bias initialisation over one orientation: bias[r] = shrunken running mean of (x - other[idx]).
// The running mean  mean += (resid - mean) / count  is a strictly sequential chain in the reference's order (the result
// is bit-identical to the reference's scalar loop), so a sweep takes as long as the longest row's chain: five dependent
// FP64 operations per stored entry.  The division is the correctly rounded quotient from a correctly rounded reciprocal
// (q0 = d * y, rem = fma(-q0, n, d) exact, q = fma(rem, y, q0): Markstein's final step, exact for integer n), which
// keeps the IEEE division routine off the chain.
//
// Two roles in one launch:
//  * rows of fewer than kBiasLongMin entries: one WARP per row.  The lanes fetch 32 consecutive entries (coalesced values
//    and indices, gathered biases), form the residuals in real_t and the reciprocals of the running counts, and park
//    them in shared memory; then every lane walks the same chain over the 32 parked entries.
//  * longer rows: one BLOCK per row (the first n_long_blocks blocks of the grid, which the hardware dispatches first).
//    Warps 1..7 are producers: they turn tile t+1 of the row (kBiasTile entries) into (residual, reciprocal) pairs in
//    shared memory while warp 0 walks the chain over tile t -- nothing but the five chain operations and broadcast
//    shared-memory reads is left in the consumer's loop (the guard on the operand range is integer work on the
//    exponent field).  Rows of at least kBiasHugeMin entries are taken first.
constexpr int kBiasWarps = 8;
constexpr int kBiasLongMin = 1024;
constexpr int kBiasHugeMin = 8192;
constexpr int kBiasProducers = (kBiasWarps - 1) * 32;
constexpr int kBiasPerProducer = 4;
constexpr int kBiasTile = kBiasProducers * kBiasPerProducer;   // 896 entries = 14 KB per buffer

// rows with at least kBiasLongMin entries: the huge ones from the front of `list`, the others from its back;
// count[0] / count[1] = how many of each
__global__ void bias_long_rows_kernel(int_t rows, const size_t *__restrict__ ptr, int_t *__restrict__ list, int *__restrict__ count)
{
    const int_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const size_t len = ptr[r + 1] - ptr[r];
    if (len >= (size_t)kBiasHugeMin) list[atomicAdd(&count[0], 1)] = r;
    else if (len >= (size_t)kBiasLongMin) list[rows - 1 - atomicAdd(&count[1], 1)] = r;
}

// exponent field of d for the range guard; an exact zero is harmless and reads as 1023
__device__ __forceinline__ int bias_guard_exponent(double d)
{
    const unsigned hi = (unsigned)__double2hiint(d) & 0x7fffffffu;
    const unsigned lo = (unsigned)__double2loint(d);
    return (hi | lo) == 0u ? 1023 : (int)(hi >> 20);
}
// operands for which the three-step quotient is not guaranteed exact (|d| below 2^-923 or above 2^917, denormals,
// infinities, NaN: never seen on ratings) send the row to the IEEE division routine
__device__ __forceinline__ bool bias_guard_suspicious(int emin, int emax) { return emin < 100 || emax > 1940; }

template <typename T>
__device__ __forceinline__ double bias_row_exact(size_t b, size_t e, const int_t *__restrict__ idx, const T *__restrict__ val,
                                                 const T *__restrict__ other)
{
    double mean = 0.;
    for (size_t t = b; t < e; t++) {
        const T resid = other ? (T)(val[t] - other[idx[t]]) : val[t];
        mean = __dadd_rn(mean, __ddiv_rn(__dsub_rn((double)resid, mean), (double)(t - b + 1)));
    }
    return mean;
}

template <typename T>
__device__ __forceinline__ void bias_row_finish(int_t r, size_t cnt, double mean, T lam, bool scale_lam, bool clamp_count_to_one,
                                                bool shrink_empty, T *__restrict__ out)
{
    if (cnt > 0 || shrink_empty) {
        const double c = (double)cnt;
        const double mult = scale_lam ? (clamp_count_to_one ? (double)(cnt > 1 ? cnt : 1) : c) : 1.;
        mean = __dmul_rn(mean, __ddiv_rn(c, __dadd_rn(c, __dmul_rn((double)lam, mult))));
    }
    out[r] = (T)mean;
}

template <typename T>
__global__ void __launch_bounds__(kBiasWarps * 32)
bias_sweep_kernel(int_t rows, const size_t *__restrict__ ptr, const int_t *__restrict__ idx, const T *__restrict__ val,
                  const T *__restrict__ other, T lam, bool scale_lam, bool clamp_count_to_one, bool shrink_empty,
                  T *__restrict__ out, const int_t *__restrict__ long_list, const int *__restrict__ long_count,
                  int n_long_blocks)
{
    __shared__ double2 s_tile[2][kBiasTile];   // long rows: (residual, reciprocal of the running count)
    __shared__ double s_resid[kBiasWarps][32], s_rcp[kBiasWarps][32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if ((int)blockIdx.x < n_long_blocks) {
        // ---------------------------------------------------------------- one block per long row
        const int n_huge = long_count[0], n_all = n_huge + long_count[1];
        for (int li = blockIdx.x; li < n_all; li += n_long_blocks) {
            const int_t r = li < n_huge ? long_list[li] : long_list[rows - 1 - (li - n_huge)];
            const size_t b = ptr[r], e = ptr[r + 1], len = e - b;
            const int ntiles = (int)((len + kBiasTile - 1) / kBiasTile);
            auto produce = [&](int t, int buf) {
                const int pt = threadIdx.x - 32;
                const size_t base = b + (size_t)t * kBiasTile;
                int_t col[kBiasPerProducer];
                T x[kBiasPerProducer], o[kBiasPerProducer];
#pragma unroll
                for (int j = 0; j < kBiasPerProducer; j++) {
                    const size_t pos = base + j * kBiasProducers + pt;
                    col[j] = 0;
                    x[j] = T(0);
                    if (pos < e) {
                        x[j] = val[pos];
                        if (other) col[j] = idx[pos];
                    }
                }
#pragma unroll
                for (int j = 0; j < kBiasPerProducer; j++) {
                    const size_t pos = base + j * kBiasProducers + pt;
                    o[j] = (other && pos < e) ? other[col[j]] : T(0);
                }
#pragma unroll
                for (int j = 0; j < kBiasPerProducer; j++) {
                    const size_t pos = base + j * kBiasProducers + pt;
                    double2 v = make_double2(0., 0.);
                    if (pos < e) {
                        v.x = (double)(other ? (T)(x[j] - o[j]) : x[j]);
                        v.y = __drcp_rn((double)(pos - b + 1));
                    }
                    s_tile[buf][j * kBiasProducers + pt] = v;
                }
            };
            if (w > 0) produce(0, 0);
            __syncthreads();
            double mean = 0.;
            int emin = 1023, emax = 1023;
            for (int t = 0; t < ntiles; t++) {
                if (w > 0) {
                    if (t + 1 < ntiles) produce(t + 1, (t + 1) & 1);
                } else {
                    const size_t done = (size_t)t * kBiasTile;
                    const int n = len - done < (size_t)kBiasTile ? (int)(len - done) : kBiasTile;
                    const double2 *tile = s_tile[t & 1];
                    double cnt = (double)done;
#pragma unroll 8
                    for (int u = 0; u < n; u++) {
                        const double2 v = tile[u];
                        cnt += 1.;
                        const double d = __dsub_rn(v.x, mean);
                        const double q0 = __dmul_rn(d, v.y);
                        const double rem = __fma_rn(-q0, cnt, d);
                        mean = __dadd_rn(mean, __fma_rn(rem, v.y, q0));
                        const int ex = bias_guard_exponent(d);
                        emin = min(emin, ex);
                        emax = max(emax, ex);
                    }
                }
                __syncthreads();
            }
            if (w == 0) {
                if (bias_guard_suspicious(emin, emax)) mean = bias_row_exact(b, e, idx, val, other);   // warp-uniform
                if (lane == 0) bias_row_finish(r, len, mean, lam, scale_lam, clamp_count_to_one, shrink_empty, out);
            }
        }
        return;
    }

    // -------------------------------------------------------------------- one warp per short row
    const int_t r = ((int_t)blockIdx.x - n_long_blocks) * kBiasWarps + w;
    if (r >= rows) return;
    const size_t b = ptr[r], e = ptr[r + 1];
    if (n_long_blocks > 0 && e - b >= (size_t)kBiasLongMin) return;
    double mean = 0.;
    // this lane's entry of the 32-entry chunk starting at t0: residual (as the reference forms it, in real_t) and the
    // reciprocal of its running count
    auto fetch = [&](size_t t0, double &resid, double &rcp) {
        const size_t t = t0 + lane;
        resid = 0.;
        rcp = 0.;
        if (t < e) {
            resid = (double)(other ? (T)(val[t] - other[idx[t]]) : val[t]);
            rcp = __drcp_rn((double)(t - b + 1));
        }
    };
    // the fetches run two chunks ahead of the chain
    int emin = 1023, emax = 1023;
    double res1, rcp1, res2, rcp2;
    fetch(b, res1, rcp1);
    fetch(b + 32, res2, rcp2);
    for (size_t t0 = b; t0 < e; t0 += 32) {
        s_resid[w][lane] = res1;
        s_rcp[w][lane] = rcp1;
        res1 = res2;
        rcp1 = rcp2;
        fetch(t0 + 64, res2, rcp2);
        __syncwarp();
        const int n = e - t0 < 32 ? (int)(e - t0) : 32;
        double cnt = (double)(t0 - b);
#pragma unroll 4
        for (int u = 0; u < n; u++) {
            cnt += 1.;
            const double d = __dsub_rn(s_resid[w][u], mean), y = s_rcp[w][u];
            const double q0 = __dmul_rn(d, y);
            const double rem = __fma_rn(-q0, cnt, d);
            mean = __dadd_rn(mean, __fma_rn(rem, y, q0));
            const int ex = bias_guard_exponent(d);
            emin = min(emin, ex);
            emax = max(emax, ex);
        }
        __syncwarp();
    }
    // warp-uniform: every lane walked the same chain
    if (bias_guard_suspicious(emin, emax)) mean = bias_row_exact(b, e, idx, val, other);
    if (lane == 0) bias_row_finish(r, e - b, mean, lam, scale_lam, clamp_count_to_one, shrink_empty, out);
}

}  // namespace

int device_compress(const int_t *d_major, const int_t *d_minor, const real_t *d_val, size_t nnz, int_t nmajor,
                    size_t *d_ptr, int_t *d_idx, real_t *d_out, cudaStream_t stream)
{
    const int threads = 256;
    // entry positions are carried as 32-bit integers through the sort: more entries than that is refused loudly
    // (the reference takes a size_t count; a silent truncation here would build a corrupt matrix)
    if (nnz > (size_t)0x7fffffff) {
        std::fprintf(stderr, "cmfrec_b200: more than 2^31-1 stored entries per device are not supported.\n");
        return 2;
    }
    const unsigned blocks = (unsigned)((nnz + threads - 1) / threads);
    if (nnz) {
        DevBuf<int> flag;
        if (!flag.alloc(1)) return 1;
        cudaMemsetAsync(flag.p, 0, sizeof(int), stream);
        index_range_kernel<<<blocks, threads, 0, stream>>>(d_major, nnz, nmajor, flag.p);
        int bad = 0;
        if (cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
            cudaStreamSynchronize(stream) != cudaSuccess)
            return 1;
        if (bad) {
            std::fprintf(stderr, "cmfrec_b200: row / column index outside of [0, %d).\n", (int)nmajor);
            return 2;
        }
    }
    // row pointers
    DevBuf<unsigned long long> counts;
    if (!counts.alloc((size_t)nmajor + 1)) return 1;
    cudaMemsetAsync(counts.p, 0, ((size_t)nmajor + 1) * sizeof(unsigned long long), stream);
    if (nnz) histogram_kernel<<<blocks, threads, 0, stream>>>(d_major, nnz, counts.p);
    size_t tb = 0;
    static_assert(sizeof(size_t) == sizeof(unsigned long long), "size_t must be 64-bit");
    cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.p, reinterpret_cast<unsigned long long *>(d_ptr), nmajor + 1, stream);
    DevBuf<unsigned char> tmp;
    if (!tmp.alloc(tb ? tb : 1)) return 1;
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, counts.p, reinterpret_cast<unsigned long long *>(d_ptr), nmajor + 1, stream);
    if (!nnz) return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : 1;
    // stable sort of positions by major index
    DevBuf<uint32_t> pos_in, pos_out;
    DevBuf<int_t> keys_out;
    if (!pos_in.alloc(nnz) || !pos_out.alloc(nnz) || !keys_out.alloc(nnz)) return 1;
    iota_kernel<<<blocks, threads, 0, stream>>>(pos_in.p, nnz);
    int bits = 1;
    while (bits < 31 && ((int_t)1 << bits) < nmajor) bits++;
    size_t sb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sb, d_major, keys_out.p, pos_in.p, pos_out.p, (int)nnz, 0, bits, stream);
    DevBuf<unsigned char> stmp;
    if (!stmp.alloc(sb ? sb : 1)) return 1;
    cub::DeviceRadixSort::SortPairs(stmp.p, sb, d_major, keys_out.p, pos_in.p, pos_out.p, (int)nnz, 0, bits, stream);
    gather_kernel<real_t><<<blocks, threads, 0, stream>>>(pos_out.p, d_minor, d_val, nnz, d_idx, d_out);
    if (cudaStreamSynchronize(stream) != cudaSuccess) return 1;   // temporaries are released on return
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

namespace {
__global__ void any_nonpositive_kernel(const real_t *__restrict__ x, size_t n, int *__restrict__ flag)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(x[i] > real_t(0))) *flag = 1;
}
}  // namespace

// *all_positive = every stored value is > 0 (synchronises the stream)
int device_all_positive(const real_t *d_x, size_t n, bool *all_positive, cudaStream_t stream)
{
    *all_positive = true;
    if (!n) return 0;
    DevBuf<int> flag;
    if (!flag.alloc(1)) return 1;
    cudaMemsetAsync(flag.p, 0, sizeof(int), stream);
    any_nonpositive_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_x, n, flag.p);
    int bad = 0;
    if (cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess)
        return 1;
    *all_positive = bad == 0;
    return 0;
}

int device_subtract(real_t *d_x, size_t n, real_t mu, cudaStream_t stream)
{
    if (!n) return 0;
    subtract_kernel<real_t><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_x, n, mu);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

namespace {
__global__ void degree_kernel(const size_t *__restrict__ ptr, int_t rows, uint32_t *__restrict__ deg, int_t *__restrict__ ids)
{
    const int_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) {
        deg[i] = (uint32_t)(ptr[i + 1] - ptr[i]);
        ids[i] = i;
    }
}
}  // namespace

// rows sorted by decreasing number of stored entries, ties in increasing row order (a stable sort, as the host
// version in als.cu): d_order on the device, the sorted counts and the total on the host
int device_degree_order(const size_t *d_ptr, int_t rows, int_t *d_order, std::vector<int_t> &deg_sorted, size_t *nnz_total,
                        cudaStream_t stream)
{
    deg_sorted.assign((size_t)rows, 0);
    if (cudaMemcpyAsync(nnz_total, d_ptr + rows, sizeof(size_t), cudaMemcpyDeviceToHost, stream) != cudaSuccess) return 1;
    if (rows < 1) return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : 1;
    DevBuf<uint32_t> deg_in, deg_out;
    DevBuf<int_t> ids;
    if (!deg_in.alloc(rows) || !deg_out.alloc(rows) || !ids.alloc(rows)) return 1;
    degree_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(d_ptr, rows, deg_in.p, ids.p);
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, deg_in.p, deg_out.p, ids.p, d_order, (int)rows, 0, 32, stream);
    DevBuf<unsigned char> tmp;
    if (!tmp.alloc(bytes ? bytes : 1)) return 1;
    cub::DeviceRadixSort::SortPairsDescending(tmp.p, bytes, deg_in.p, deg_out.p, ids.p, d_order, (int)rows, 0, 32, stream);
    static_assert(sizeof(int_t) == sizeof(uint32_t), "the counts are downloaded as they are");
    if (cudaMemcpyAsync(deg_sorted.data(), deg_out.p, (size_t)rows * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream) != cudaSuccess)
        return 1;
    return cudaStreamSynchronize(stream) == cudaSuccess && cudaGetLastError() == cudaSuccess ? 0 : 1;
}

namespace {
// sorted position s of a row -> device row (s % world) * block + s / world: rows are dealt round-robin in order of
// decreasing degree (equal blocks for the all-gather, near-equal numbers of stored entries)
__global__ void deal_kernel(const int_t *__restrict__ order, int_t rows, int world, int_t block, int_t *__restrict__ to_dev,
                            int_t *__restrict__ to_old)
{
    const int_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < rows) {
        const int_t r = order[s];
        const int_t dev = (s % world) * block + s / world;
        to_dev[r] = dev;
        to_old[dev] = r;
    }
}
__global__ void fill_kernel(int_t *x, size_t n, int_t v)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}
// number of stored entries of every device row held by this rank (0 for the rows of other ranks and for padding)
__global__ void local_len_kernel(const size_t *__restrict__ ptr_full, const int_t *__restrict__ to_old, int_t row_begin,
                                 int_t row_end, int_t rows_padded, unsigned long long *__restrict__ len)
{
    const int_t dev = blockIdx.x * blockDim.x + threadIdx.x;
    if (dev <= rows_padded) {
        unsigned long long l = 0;
        if (dev >= row_begin && dev < row_end && dev < rows_padded) {
            const int_t old = to_old[dev];
            if (old >= 0) l = ptr_full[old + 1] - ptr_full[old];
        }
        len[dev] = l;
    }
}
// one warp per local row: copy its entries, column ids renumbered into the other side's device numbering
__global__ void gather_rows_kernel(const size_t *__restrict__ ptr_full, const int_t *__restrict__ idx_full,
                                   const real_t *__restrict__ val_full, const int_t *__restrict__ to_old,
                                   const int_t *__restrict__ other_to_dev, int_t row_begin, int_t n_local,
                                   const size_t *__restrict__ ptr_local, int_t *__restrict__ idx_local,
                                   real_t *__restrict__ val_local)
{
    const int_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n_local) return;
    const int_t dev = row_begin + i;
    const int_t old = to_old[dev];
    if (old < 0) return;
    const size_t src = ptr_full[old], cnt = ptr_full[old + 1] - src, dst = ptr_local[dev];
    for (size_t e = lane; e < cnt; e += 32) {
        idx_local[dst + e] = other_to_dev[idx_full[src + e]];
        val_local[dst + e] = val_full[src + e];
    }
}
// rows of a dense host-numbered matrix -> device numbering with padded stride (and back)
__global__ void scatter_rows_kernel(const real_t *__restrict__ src, int lds, int_t rows, int kk, const int_t *__restrict__ to_dev,
                                    real_t *__restrict__ dst, int ldd)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)rows * kk) return;
    const int_t r = (int_t)(t / kk);
    const int c = (int)(t - (size_t)r * kk);
    dst[(size_t)(to_dev ? to_dev[r] : r) * ldd + c] = src[(size_t)r * lds + c];
}
__global__ void gather_back_kernel(const real_t *__restrict__ src, int lds, int_t rows, int kk, const int_t *__restrict__ to_dev,
                                   real_t *__restrict__ dst, int ldd)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)rows * kk) return;
    const int_t r = (int_t)(t / kk);
    const int c = (int)(t - (size_t)r * kk);
    dst[(size_t)r * ldd + c] = src[(size_t)(to_dev ? to_dev[r] : r) * lds + c];
}
}  // namespace

// d_order: rows by decreasing degree (device_degree_order).  Fills to_dev [rows] and to_old [world * block] (-1 = padding).
int device_deal_rows(const int_t *d_order, int_t rows, int world, int_t block, int_t *d_to_dev, int_t *d_to_old, cudaStream_t stream)
{
    const size_t padded = (size_t)world * block;
    fill_kernel<<<(unsigned)((padded + 255) / 256), 256, 0, stream>>>(d_to_old, padded, (int_t)-1);
    if (rows > 0) deal_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(d_order, rows, world, block, d_to_dev, d_to_old);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// This rank's block of one orientation out of the full matrix: ptr_local [rows_padded + 1] (device row ids; rows of other
// ranks are empty), idx / val of the local entries with column ids in the other side's device numbering.
// *nnz_local receives the number of local entries (synchronises).  idx_local / val_local are allocated here.
int device_extract_block(const size_t *ptr_full, const int_t *idx_full, const real_t *val_full, const int_t *to_old,
                         const int_t *other_to_dev, int_t row_begin, int_t row_end, int_t n_local, int_t rows_padded,
                         DevBuf<size_t> &ptr_local, DevBuf<int_t> &idx_local, DevBuf<real_t> &val_local, size_t *nnz_local,
                         cudaStream_t stream)
{
    DevBuf<unsigned long long> len;
    if (!len.alloc((size_t)rows_padded + 1) || !ptr_local.alloc((size_t)rows_padded + 1)) return 1;
    local_len_kernel<<<(rows_padded + 1 + 255) / 256, 256, 0, stream>>>(ptr_full, to_old, row_begin, row_end, rows_padded, len.p);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, len.p, reinterpret_cast<unsigned long long *>(ptr_local.p), rows_padded + 1, stream);
    DevBuf<unsigned char> tmp;
    if (!tmp.alloc(tb ? tb : 1)) return 1;
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, len.p, reinterpret_cast<unsigned long long *>(ptr_local.p), rows_padded + 1, stream);
    size_t total = 0;
    if (cudaMemcpyAsync(&total, ptr_local.p + rows_padded, sizeof(size_t), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess)
        return 1;
    *nnz_local = total;
    if (!idx_local.alloc(total ? total : 1) || !val_local.alloc(total ? total : 1)) return 1;
    if (n_local > 0 && total > 0) {
        const long long threads = (long long)n_local * 32;
        gather_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(ptr_full, idx_full, val_full, to_old, other_to_dev,
                                                                                   row_begin, n_local, ptr_local.p, idx_local.p, val_local.p);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int device_scatter_rows(const real_t *src, int lds, int_t rows, int kk, const int_t *to_dev, real_t *dst, int ldd, cudaStream_t stream)
{
    const size_t n = (size_t)rows * kk;
    if (n) scatter_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, lds, rows, kk, to_dev, dst, ldd);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
int device_gather_rows_back(const real_t *src, int lds, int_t rows, int kk, const int_t *to_dev, real_t *dst, int ldd, cudaStream_t stream)
{
    const size_t n = (size_t)rows * kk;
    if (n) gather_back_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, lds, rows, kk, to_dev, dst, ldd);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// the long-row list of one orientation (see bias_sweep_kernel) and how many blocks take rows from it
struct BiasLongRows {
    DevBuf<int_t> list;
    DevBuf<int> count;
    int blocks = 0;
    int build(int_t rows, const size_t *ptr, cudaStream_t stream)
    {
        blocks = 0;
        if (rows < 1 || env_flag_off("CMFB200_BIAS_LONG")) return 0;
        if (!list.alloc((size_t)rows) || !count.alloc(2)) return 1;
        if (cudaMemsetAsync(count.p, 0, 2 * sizeof(int), stream) != cudaSuccess) return 1;
        bias_long_rows_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(rows, ptr, list.p, count.p);
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        blocks = std::min<long long>((long long)sms * 4, rows);
        return cudaGetLastError() == cudaSuccess ? 0 : 1;
    }
};

template <typename T>
static void launch_bias_sweep(int_t rows, const size_t *ptr, const int_t *idx, const T *val, const T *other, T lam, bool scale_lam,
                              bool clamp_count_to_one, bool shrink_empty, T *out, const BiasLongRows &lr, cudaStream_t stream)
{
    const unsigned grid = (unsigned)lr.blocks + (unsigned)((rows + kBiasWarps - 1) / kBiasWarps);
    bias_sweep_kernel<T><<<grid, kBiasWarps * 32, 0, stream>>>(rows, ptr, idx, val, other, lam, scale_lam, clamp_count_to_one,
                                                               shrink_empty, out, lr.list.p, lr.count.p, lr.blocks);
}

int device_init_biases_twosided(int_t m, int_t n, const size_t *csr_p, const int_t *csr_i, const real_t *csr_v,
                                const size_t *csc_p, const int_t *csc_i, const real_t *csc_v, real_t lam_user, real_t lam_item,
                                bool scale_lam, real_t *d_biasA, real_t *d_biasB, cudaStream_t stream)
{
    if (fabs((double)lam_user) < (double)CMF_EPS) lam_user = CMF_EPS;
    if (fabs((double)lam_item) < (double)CMF_EPS) lam_item = CMF_EPS;
    if (cudaMemsetAsync(d_biasA, 0, (size_t)m * sizeof(real_t), stream) != cudaSuccess ||
        cudaMemsetAsync(d_biasB, 0, (size_t)n * sizeof(real_t), stream) != cudaSuccess)
        return 1;
    BiasLongRows long_items, long_users;
    if (long_items.build(n, csc_p, stream) || long_users.build(m, csr_p, stream)) return 1;

    for (int s = 0; s < 5; s++) {
        launch_bias_sweep<real_t>(n, csc_p, csc_i, csc_v, d_biasA, lam_item, scale_lam, true, true, d_biasB, long_items, stream);
        launch_bias_sweep<real_t>(m, csr_p, csr_i, csr_v, d_biasB, lam_user, scale_lam, false, false, d_biasA, long_users, stream);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int device_init_biases_onesided(int_t rows, const size_t *ptr, const real_t *val, real_t lam, bool scale_lam, real_t *d_bias,
                                cudaStream_t stream)
{
    if (fabs((double)lam) < (double)CMF_EPS) lam = CMF_EPS;
    BiasLongRows lr;
    if (lr.build(rows, ptr, stream)) return 1;
    launch_bias_sweep<real_t>(rows, ptr, nullptr, val, nullptr, lam, scale_lam, true, true, d_bias, lr, stream);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace cmfb200
