// Small dense / sparse-dense building blocks of the models with side information and implicit features
// (reference optimizeA Case 1 src/common.c:2793-2900, Case 3 :3117-3203; optimizeA_collective general case
// src/collective.c:5566-5968).  None of these dominate an iteration; they are written for clarity and determinism.
#include "dense_small.h"
#include "device_utils.cuh"
#include <cstdint>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace cmfb200 {

namespace {

constexpr int TILE = 64, PANEL = 16, MAX_SLICES = 64;

// partial[slice][ncx][ncy] += X[rows of slice, :ncx]^T  Y[rows of slice, :ncy]   (64x64 tiles, 4x4 per thread)
template <typename T>
__global__ void __launch_bounds__(256) xty_partial_kernel(const T *__restrict__ X, int ldx, int ncx, const T *__restrict__ Y,
                                                          int ldy, int ncy, int_t rows, int nslices, T *__restrict__ partial)
{
    __shared__ T sa[PANEL][TILE + 1];
    __shared__ T sb[PANEL][TILE + 1];
    const int ti = blockIdx.x, tj = blockIdx.y, slice = blockIdx.z;
    const long long per = ((long long)rows + nslices - 1) / nslices;
    const long long r_begin = (long long)slice * per;
    long long r_end = r_begin + per;
    if (r_end > rows) r_end = rows;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = T(0);
    for (long long r0 = r_begin; r0 < r_end; r0 += PANEL) {
        for (int i = threadIdx.x; i < PANEL * TILE; i += 256) {
            const int rr = i / TILE, cc = i % TILE;
            const long long r = r0 + rr;
            const int ca = ti * TILE + cc, cb = tj * TILE + cc;
            const bool ok = r < r_end;
            sa[rr][cc] = (ok && ca < ncx) ? X[(size_t)r * ldx + ca] : T(0);
            sb[rr][cc] = (ok && cb < ncy) ? Y[(size_t)r * ldy + cb] : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < PANEL; rr++) {
            T av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) av[i] = sa[rr][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; j++) bv[j] = sb[rr][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    T *out = partial + (size_t)slice * ncx * ncy;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int a = ti * TILE + ty + 16 * i, b = tj * TILE + tx + 16 * j;
            if (a < ncx && b < ncy) out[(size_t)a * ncy + b] = acc[i][j];
        }
}

template <typename T> __global__ void sum_slices_kernel(const T *__restrict__ partial, int total, int nslices, T *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    T s = T(0);
    for (int sl = 0; sl < nslices; sl++) s += partial[(size_t)sl * total + i];
    out[i] = s;
}

// out[r, :nc] = alpha * M[r, :p] . S[:p, :nc]  (+ out if accumulate); one warp per row, S cached in shared memory
template <typename T>
__global__ void rows_times_small_kernel(const T *__restrict__ M, int ldm, int p, const T *__restrict__ S, int lds, int nc,
                                        T alpha, bool accumulate, T *__restrict__ out, int ldo, int_t rows)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *Ss = reinterpret_cast<T *>(smem_raw);
    for (int i = threadIdx.x; i < p * nc; i += blockDim.x) Ss[i] = S[(size_t)(i / nc) * lds + (i % nc)];
    __syncthreads();
    const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int_t r = blockIdx.x * warps + w; r < rows; r += gridDim.x * warps) {
        const T *mrow = M + (size_t)r * ldm;
        for (int c = lane; c < nc; c += 32) {
            T s = T(0);
            for (int q = 0; q < p; q++) s = fma(mrow[q], Ss[q * nc + c], s);
            T *o = out + (size_t)r * ldo + c;
            *o = accumulate ? fma(alpha, s, *o) : alpha * s;
        }
    }
}

// Y[r, :kk] = alpha * sum_{e in row r} F[idx_e, :kk]  (+ Y if accumulate); warp per row, block per long row
template <typename T>
__global__ void __launch_bounds__(256) spmm_ones_kernel(const size_t *__restrict__ ptr, const int_t *__restrict__ idx,
                                                        const int_t *__restrict__ order, int_t n_rows, int_t n_long,
                                                        const T *__restrict__ F, int ldf, int kk, T alpha, bool accumulate,
                                                        T *__restrict__ Y, int ldy)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *red = reinterpret_cast<T *>(smem_raw);   // [8][kk]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_slots = n_long + (n_rows - n_long + 7) / 8;
    for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        const bool team = slot < n_long;
        const int i = team ? slot : n_long + (slot - n_long) * 8 + w;
        if (i >= n_rows) continue;   // warp-uniform; team slots never take this branch
        const int_t row = order[i];
        const size_t b = ptr[row], e = ptr[row + 1];
        for (int c0 = 0; c0 < kk; c0 += 32) {
            const int c = c0 + lane;
            T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
            size_t t = b + (team ? w : 0);
            const size_t step = team ? 8 : 1;
            if (c < kk) {
                for (; t + 3 * step < e; t += 4 * step) {
                    s0 += F[(size_t)idx[t] * ldf + c];
                    s1 += F[(size_t)idx[t + step] * ldf + c];
                    s2 += F[(size_t)idx[t + 2 * step] * ldf + c];
                    s3 += F[(size_t)idx[t + 3 * step] * ldf + c];
                }
                for (; t < e; t += step) s0 += F[(size_t)idx[t] * ldf + c];
            }
            T s = (s0 + s1) + (s2 + s3);
            if (team) {
                __syncthreads();
                if (c < kk) red[w * kk + c] = s;
                __syncthreads();
                if (w == 0 && c < kk) {
                    s = T(0);
                    for (int ww = 0; ww < 8; ww++) s += red[ww * kk + c];
                }
            }
            if ((!team || w == 0) && c < kk) {
                T *o = Y + (size_t)row * ldy + c;
                *o = accumulate ? fma(alpha, s, *o) : alpha * s;
            }
        }
        if (team) __syncthreads();
    }
}

// R[r, :d] := solution of (L L^T) x = R[r, :d]; L lower-triangular [d x d] row-major; one warp per row
template <typename T>
__global__ void tri_solve_rows_kernel(const T *__restrict__ Lmat, int d, T *__restrict__ R, int ldr, int_t rows)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *Ls = reinterpret_cast<T *>(smem_raw);                  // [d][d+1]
    const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T *ys = Ls + (size_t)d * (d + 1) + (size_t)w * d;         // [warps][d]
    for (int i = threadIdx.x; i < d * d; i += blockDim.x) Ls[(i / d) * (d + 1) + (i % d)] = Lmat[i];
    __syncthreads();
    for (int_t r = blockIdx.x * warps + w; r < rows; r += gridDim.x * warps) {
        T *x = R + (size_t)r * ldr;
        for (int i = lane; i < d; i += 32) ys[i] = x[i];
        __syncwarp();
        for (int i = 0; i < d; i++) {                         // forward: L y = b
            T s = T(0);
            for (int t = lane; t < i; t += 32) s = fma(Ls[i * (d + 1) + t], ys[t], s);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) ys[i] = (ys[i] - s) / Ls[i * (d + 1) + i];
            __syncwarp();
        }
        for (int i = d - 1; i >= 0; i--) {                    // backward: L^T x = y
            T s = T(0);
            for (int t = i + 1 + lane; t < d; t += 32) s = fma(Ls[t * (d + 1) + i], ys[t], s);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) ys[i] = (ys[i] - s) / Ls[i * (d + 1) + i];
            __syncwarp();
        }
        for (int i = lane; i < d; i += 32) x[i] = ys[i];
        __syncwarp();
    }
}

// The same solve with one LANE per row: a warp takes 32 rows, loads them coalesced and parks them transposed in shared
// memory (ys[i][row], row stride 33: conflict-free both ways), then every lane runs the two substitutions of its own row
// with the elements of L broadcast from shared memory -- no shuffles, no per-step synchronisation.  Used when L and the
// parked rows fit shared memory (launch_tri_solve_rows picks the number of warps per block).
template <typename T>
__global__ void tri_solve_lanes_kernel(const T *__restrict__ Lmat, int d, T *__restrict__ R, int ldr, int_t rows)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *Ls = reinterpret_cast<T *>(smem_raw);                  // [d][d+1]; the diagonal holds its reciprocal-free value
    const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T *ys = Ls + (size_t)d * (d + 1) + (size_t)w * d * 33;    // [warps][d][33]
    for (int i = threadIdx.x; i < d * d; i += blockDim.x) Ls[(i / d) * (d + 1) + (i % d)] = Lmat[i];
    __syncthreads();
    for (long long r0 = ((long long)blockIdx.x * warps + w) * 32; r0 < rows; r0 += (long long)gridDim.x * warps * 32) {
        const int nr = rows - r0 < 32 ? (int)(rows - r0) : 32;
        for (int rr = 0; rr < nr; rr++) {
            const T *x = R + (size_t)(r0 + rr) * ldr;
            for (int i = lane; i < d; i += 32) ys[i * 33 + rr] = x[i];
        }
        __syncwarp();
        if (lane < nr) {
            T *y = ys + lane;
            for (int i = 0; i < d; i++) {                     // forward: L y = b
                const T *Li = Ls + i * (d + 1);
                T s0 = T(0), s1 = T(0);
                int t = 0;
                for (; t + 1 < i; t += 2) {
                    s0 = fma(Li[t], y[t * 33], s0);
                    s1 = fma(Li[t + 1], y[(t + 1) * 33], s1);
                }
                if (t < i) s0 = fma(Li[t], y[t * 33], s0);
                y[i * 33] = (y[i * 33] - (s0 + s1)) / Li[i];
            }
            for (int i = d - 1; i >= 0; i--) {                // backward: L^T x = y
                T s0 = T(0), s1 = T(0);
                int t = i + 1;
                for (; t + 1 < d; t += 2) {
                    s0 = fma(Ls[t * (d + 1) + i], y[t * 33], s0);
                    s1 = fma(Ls[(t + 1) * (d + 1) + i], y[(t + 1) * 33], s1);
                }
                if (t < d) s0 = fma(Ls[t * (d + 1) + i], y[t * 33], s0);
                y[i * 33] = (y[i * 33] - (s0 + s1)) / Ls[i * (d + 1) + i];
            }
        }
        __syncwarp();
        for (int rr = 0; rr < nr; rr++) {
            T *x = R + (size_t)(r0 + rr) * ldr;
            for (int i = lane; i < d; i += 32) x[i] = ys[i * 33 + rr];
        }
        __syncwarp();
    }
}

// L := Cholesky factor (lower, row-major, zeros above the diagonal) of S + lam*I; one thread block; S full symmetric [d x d].
// Works in place in global memory (the matrix is tiny and stays in L1/L2).
template <typename T> __global__ void __launch_bounds__(256) spd_factor_kernel(const T *__restrict__ S, int d, T lam, T *__restrict__ Lout)
{
    __shared__ T dj_s;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < d * d; i += nt) {
        const int r = i / d, c = i % d;
        Lout[i] = (c <= r) ? S[i] + ((c == r) ? lam : T(0)) : T(0);
    }
    __syncthreads();
    for (int j = 0; j < d; j++) {
        if (tid == 0) {
            const T v = sqrt(Lout[j * d + j]);
            Lout[j * d + j] = v;
            dj_s = v;
        }
        __syncthreads();
        const T inv = T(1) / dj_s;
        for (int i = j + 1 + tid; i < d; i += nt) Lout[i * d + j] *= inv;
        __syncthreads();
        // trailing update of the lower triangle: (i, c) with j < c <= i
        const int rem = d - j - 1;
        for (int e = tid; e < rem * rem; e += nt) {
            const int i = j + 1 + e / rem, c = j + 1 + e % rem;
            if (c <= i) Lout[i * d + c] = fma(-Lout[i * d + j], Lout[c * d + j], Lout[i * d + c]);
        }
        __syncthreads();
    }
}

// The same factorisation (same operations on every element in the same order: identical results) with the matrix in shared
// memory, row stride d + 1 (column accesses conflict-free); the working copy in global memory cost 87 us at d = 64.
template <typename T> __global__ void __launch_bounds__(256) spd_factor_smem_kernel(const T *__restrict__ S, int d, T lam, T *__restrict__ Lout)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *W = reinterpret_cast<T *>(smem_raw);   // [d][d + 1]
    __shared__ T dj_s;
    const int tid = threadIdx.x, nt = blockDim.x, ld = d + 1;
    for (int i = tid; i < d * d; i += nt) {
        const int r = i / d, c = i % d;
        W[r * ld + c] = (c <= r) ? S[i] + ((c == r) ? lam : T(0)) : T(0);
    }
    __syncthreads();
    for (int j = 0; j < d; j++) {
        if (tid == 0) {
            const T v = sqrt(W[j * ld + j]);
            W[j * ld + j] = v;
            dj_s = v;
        }
        __syncthreads();
        const T inv = T(1) / dj_s;
        for (int i = j + 1 + tid; i < d; i += nt) W[i * ld + j] *= inv;
        __syncthreads();
        // trailing update of the lower triangle, 16 x 16 threads over (row, column): no integer division in the loop
        // (the linear index of the global-memory version spent 71 us at d = 64 mostly on it)
        const int ty = tid >> 4, tx = tid & 15;
        for (int i = j + 1 + ty; i < d; i += 16) {
            const T lij = W[i * ld + j];
            for (int c = j + 1 + tx; c <= i; c += 16) W[i * ld + c] = fma(-lij, W[c * ld + j], W[i * ld + c]);
        }
        __syncthreads();
    }
    for (int i = tid; i < d * d; i += nt) Lout[i] = W[(i / d) * ld + (i % d)];
}

template <typename T> __global__ void axpby_kernel(int n, T alpha, const T *x, T beta, const T *y, T *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (x ? alpha * x[i] : T(0)) + (y ? beta * y[i] : T(0));
}

int sm_count()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace

size_t xty_workspace_elems(int ncx, int ncy) { return (size_t)MAX_SLICES * ncx * ncy; }

int launch_xty(const real_t *X, int ldx, int ncx, const real_t *Y, int ldy, int ncy, int_t rows, real_t *out, real_t *workspace,
               cudaStream_t stream)
{
    if (ncx < 1 || ncy < 1) return 2;
    const int tx = (ncx + TILE - 1) / TILE, ty = (ncy + TILE - 1) / TILE;
    int ns = (2 * sm_count() + tx * ty - 1) / (tx * ty);
    if (ns > MAX_SLICES) ns = MAX_SLICES;
    const long long by_rows = ((long long)rows + 4 * PANEL - 1) / (4 * PANEL);
    if (ns > by_rows) ns = (int)by_rows;
    if (ns < 1) ns = 1;
    xty_partial_kernel<real_t><<<dim3(tx, ty, ns), 256, 0, stream>>>(X, ldx, ncx, Y, ldy, ncy, rows, ns, workspace);
    const int total = ncx * ncy;
    sum_slices_kernel<real_t><<<(total + 255) / 256, 256, 0, stream>>>(workspace, total, ns, out);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_rows_times_small(const real_t *M, int ldm, int p, const real_t *S, int lds, int nc, real_t alpha, bool accumulate,
                            real_t *out, int ldo, int_t rows, cudaStream_t stream)
{
    if (rows < 1 || p < 1) return 0;
    const size_t smem = (size_t)p * nc * sizeof(real_t);
    auto kern = rows_times_small_kernel<real_t>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 2;
    }
    int blocks = (int)std::min<long long>(((long long)rows + 7) / 8, 4LL * sm_count());
    kern<<<blocks, 256, smem, stream>>>(M, ldm, p, S, lds, nc, alpha, accumulate, out, ldo, rows);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

namespace {

// Same product with 16-byte gathers: 8 lanes read one whole 128-byte line of an opposing row per load instruction,
// 4 entries per step, the (column) indices of 32 entries fetched coalesced and handed round with shuffles; every lane
// keeps NV vector accumulators, the four groups of a warp are combined with shuffles at the end of the row.
// Needs 16-byte aligned rows of F (ldf a multiple of the vector width) and kk <= 8 * NV * (16 / sizeof(T)).
template <typename T, int NV>
__global__ void __launch_bounds__(256) spmm_ones_vec_kernel(const size_t *__restrict__ ptr, const int_t *__restrict__ idx,
                                                            const int_t *__restrict__ order, int_t n_rows, int_t n_long,
                                                            const T *__restrict__ F, int ldf, int kk, T alpha, bool accumulate,
                                                            T *__restrict__ Y, int ldy)
{
    constexpr int VN = 16 / (int)sizeof(T);
    typedef typename VecOf<T>::type Vec;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *red = reinterpret_cast<T *>(smem_raw);   // [8][8 * NV * VN]
    constexpr int KP = 8 * NV * VN;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, l = lane & 7;
    const int pieces = (kk + VN - 1) / VN;
    const int n_slots = n_long + (n_rows - n_long + 7) / 8;
    for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        const bool team = slot < n_long;
        const int i = team ? slot : n_long + (slot - n_long) * 8 + w;
        if (i >= n_rows) continue;   // warp-uniform; team slots never take this branch
        const int_t row = order[i];
        const size_t b = ptr[row], e = ptr[row + 1];
        T acc[NV][VN];
#pragma unroll
        for (int v = 0; v < NV; v++)
#pragma unroll
            for (int q = 0; q < VN; q++) acc[v][q] = T(0);
        // 32-entry chunks: this warp's chunks are w, w + 8, ... of a team row, all of a warp row
        for (size_t t0 = b + (team ? (size_t)w * 32 : 0); t0 < e; t0 += team ? 256 : 32) {
            const size_t t = t0 + lane;
            const int col_r = t < e ? idx[t] : -1;
            const int left = e - t0 < 32 ? (int)(e - t0) : 32;
#pragma unroll 4
            for (int st = 0; st < 8; st++) {
                if (st * 4 >= left) break;   // warp-uniform
                const int col = __shfl_sync(CMF_FULL_MASK, col_r, st * 4 + g);
                if (col >= 0) {
                    const Vec *frow = reinterpret_cast<const Vec *>(F + (size_t)col * (size_t)ldf);
#pragma unroll
                    for (int v = 0; v < NV; v++) {
                        if (v * 8 + l < pieces) {
                            const Vec x = __ldg(frow + v * 8 + l);
                            const T *px = reinterpret_cast<const T *>(&x);
#pragma unroll
                            for (int q = 0; q < VN; q++) acc[v][q] += px[q];
                        }
                    }
                }
            }
        }
        // combine the four groups of the warp
#pragma unroll
        for (int v = 0; v < NV; v++)
#pragma unroll
            for (int q = 0; q < VN; q++) {
                acc[v][q] += __shfl_xor_sync(CMF_FULL_MASK, acc[v][q], 8);
                acc[v][q] += __shfl_xor_sync(CMF_FULL_MASK, acc[v][q], 16);
            }
        if (team) {
            __syncthreads();
            if (g == 0) {
#pragma unroll
                for (int v = 0; v < NV; v++)
#pragma unroll
                    for (int q = 0; q < VN; q++) red[w * KP + (v * 8 + l) * VN + q] = acc[v][q];
            }
            __syncthreads();
            if (w == 0 && g == 0) {
#pragma unroll
                for (int v = 0; v < NV; v++)
#pragma unroll
                    for (int q = 0; q < VN; q++) {
                        T s = T(0);
                        for (int ww = 0; ww < 8; ww++) s += red[ww * KP + (v * 8 + l) * VN + q];
                        acc[v][q] = s;
                    }
            }
        }
        if ((!team || w == 0) && g == 0) {
#pragma unroll
            for (int v = 0; v < NV; v++)
#pragma unroll
                for (int q = 0; q < VN; q++) {
                    const int c = (v * 8 + l) * VN + q;
                    if (c < kk) {
                        T *o = Y + (size_t)row * ldy + c;
                        *o = accumulate ? fma(alpha, acc[v][q], *o) : alpha * acc[v][q];
                    }
                }
        }
        if (team) __syncthreads();
    }
}

}  // namespace

int launch_spmm_ones(const CsrView &X, const SweepPlan &plan, const real_t *F, int ldf, int kk, real_t alpha, bool accumulate,
                     real_t *Y, int ldy, cudaStream_t stream)
{
    if (plan.n_rows < 1) return 0;
    const int n_slots = plan.n_long + (plan.n_rows - plan.n_long + 7) / 8;
    {
        constexpr int VN = 16 / (int)sizeof(real_t);
        const int pieces = (kk + VN - 1) / VN;
        const bool aligned = ldf % VN == 0 && ((uintptr_t)F & 15u) == 0 && pieces * VN <= ldf;
        if (aligned && pieces <= 32) {
            const int blocks_v = std::min(n_slots, 8 * sm_count());
            if (pieces <= 8) {
                spmm_ones_vec_kernel<real_t, 1><<<blocks_v, 256, (size_t)8 * 8 * 1 * VN * sizeof(real_t), stream>>>(
                    X.ptr, X.idx, plan.order, plan.n_rows, plan.n_long, F, ldf, kk, alpha, accumulate, Y, ldy);
            } else if (pieces <= 16) {
                spmm_ones_vec_kernel<real_t, 2><<<blocks_v, 256, (size_t)8 * 8 * 2 * VN * sizeof(real_t), stream>>>(
                    X.ptr, X.idx, plan.order, plan.n_rows, plan.n_long, F, ldf, kk, alpha, accumulate, Y, ldy);
            } else {
                spmm_ones_vec_kernel<real_t, 4><<<blocks_v, 256, (size_t)8 * 8 * 4 * VN * sizeof(real_t), stream>>>(
                    X.ptr, X.idx, plan.order, plan.n_rows, plan.n_long, F, ldf, kk, alpha, accumulate, Y, ldy);
            }
            return cudaGetLastError() == cudaSuccess ? 0 : 1;
        }
    }
    int blocks = std::min(n_slots, 8 * sm_count());
    spmm_ones_kernel<real_t><<<blocks, 256, (size_t)8 * kk * sizeof(real_t), stream>>>(X.ptr, X.idx, plan.order, plan.n_rows,
                                                                                        plan.n_long, F, ldf, kk, alpha,
                                                                                        accumulate, Y, ldy);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int spd_factor_host(int d, const real_t *S_host, std::vector<real_t> &L_host)
{
    std::vector<double> Lm((size_t)d * d, 0.);
    for (int i = 0; i < d; i++)
        for (int j = 0; j <= i; j++) Lm[(size_t)i * d + j] = S_host[(size_t)i * d + j];
    for (int j = 0; j < d; j++) {
        double s = Lm[(size_t)j * d + j];
        for (int t = 0; t < j; t++) s -= Lm[(size_t)j * d + t] * Lm[(size_t)j * d + t];
        if (!(s > 0)) return 1;
        const double dj = std::sqrt(s);
        Lm[(size_t)j * d + j] = dj;
        for (int i = j + 1; i < d; i++) {
            double v = Lm[(size_t)i * d + j];
            for (int t = 0; t < j; t++) v -= Lm[(size_t)i * d + t] * Lm[(size_t)j * d + t];
            Lm[(size_t)i * d + j] = v / dj;
        }
    }
    L_host.resize((size_t)d * d);
    for (size_t i = 0; i < L_host.size(); i++) L_host[i] = (real_t)Lm[i];
    return 0;
}

int launch_spd_factor(const real_t *S_dev, int d, real_t lam, real_t *L_dev, cudaStream_t stream)
{
    const size_t smem = (size_t)d * (d + 1) * sizeof(real_t);
    if (smem <= 200 * 1024) {
        auto kern = spd_factor_smem_kernel<real_t>;
        if (smem <= 48 * 1024 || cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
            kern<<<1, 256, smem, stream>>>(S_dev, d, lam, L_dev);
            return cudaGetLastError() == cudaSuccess ? 0 : 1;
        }
        cudaGetLastError();
    }
    spd_factor_kernel<real_t><<<1, 256, 0, stream>>>(S_dev, d, lam, L_dev);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_tri_solve_rows(const real_t *L_dev, int d, real_t *R, int ldr, int_t rows, cudaStream_t stream)
{
    if (rows < 1) return 0;
    // many rows: one lane per row while L and 32 parked rows per warp fit shared memory
    static const bool lanes_on = [] { const char *e = std::getenv("CMFB200_TRI_LANES"); return !e || std::atoi(e) != 0; }();
    if (lanes_on && rows >= 256) {
        for (int warps = 8; warps >= 2; warps >>= 1) {
            const size_t smem = ((size_t)d * (d + 1) + (size_t)warps * d * 33) * sizeof(real_t);
            const long long groups = ((long long)rows + 31) / 32;
            if (smem > 200 * 1024 || (warps > 2 && (groups + warps - 1) / warps < sm_count())) continue;   // fill the SMs first
            auto kern = tri_solve_lanes_kernel<real_t>;
            if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                cudaGetLastError();
                break;
            }
            const int blocks = (int)std::min<long long>((groups + warps - 1) / warps, 4LL * sm_count());
            kern<<<blocks, warps * 32, smem, stream>>>(L_dev, d, R, ldr, rows);
            return cudaGetLastError() == cudaSuccess ? 0 : 1;
        }
    }
    const int warps = 8;
    const size_t smem = ((size_t)d * (d + 1) + (size_t)warps * d) * sizeof(real_t);
    auto kern = tri_solve_rows_kernel<real_t>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 2;
    }
    int blocks = (int)std::min<long long>(((long long)rows + warps - 1) / warps, 4LL * sm_count());
    kern<<<blocks, warps * 32, smem, stream>>>(L_dev, d, R, ldr, rows);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_axpby(int n, real_t alpha, const real_t *x, real_t beta, const real_t *y, real_t *out, cudaStream_t stream)
{
    axpby_kernel<real_t><<<(n + 255) / 256, 256, 0, stream>>>(n, alpha, x, beta, y, out);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace cmfb200
