// Gram matrix of a factor: gram = G[:, :kk]^T G[:, :kk]  (reference: the cblas_tsyrk call sites, e.g.
// src/common.c:3328 for the implicit half-sweep; the reference keeps only the upper triangle, here the
// full symmetric matrix is written so that the row solvers can read rows contiguously).
//
// Split-K over row slices: block (tile, slice) accumulates one 64x64 output tile over its slice of rows in
// registers (4x4 per thread) from shared-memory staged 16-row panels, writes it to a per-slice partial,
// and a second kernel sums the partials in a fixed order, so the result is run-to-run deterministic.
#include "sweep.h"
#include <cstdlib>

namespace cmfb200 {

// gram_tc.cu: per-slice partials on the tensor cores (fp32 library, row widths of 64 / 128 / 256 floats)
int launch_gram_partials_tc(const real_t *G, int ldG, int_t rows, int kk, int max_slices, real_t *partial, int *nslices_out,
                            cudaStream_t stream);

namespace {

constexpr int TILE = 64;
constexpr int PANEL = 16;
constexpr int MAX_SLICES = 148;   // one slice per SM in the tensor-core path (gram_tc.cu)

template <typename T>
__global__ void __launch_bounds__(256) gram_partial_kernel(const T *__restrict__ G, int ldG, int_t rows, int kk,
                                                           int ntile, int nslices, T *__restrict__ partial)
{
    __shared__ T sa[PANEL][TILE + 1];
    __shared__ T sb[PANEL][TILE + 1];
    // upper-triangular tile pair from a linear index
    int ti = 0, tj = 0;
    {
        int t = blockIdx.x;
        for (ti = 0; ti < ntile; ti++) {
            const int span = ntile - ti;
            if (t < span) { tj = ti + t; break; }
            t -= span;
        }
    }
    const int slice = blockIdx.y;
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
        // stage PANEL rows of the two column blocks (coalesced along columns)
        for (int i = threadIdx.x; i < PANEL * TILE; i += 256) {
            const int rr = i / TILE, cc = i % TILE;
            const long long r = r0 + rr;
            const int ca = ti * TILE + cc, cb = tj * TILE + cc;
            const bool ok = r < r_end;
            sa[rr][cc] = (ok && ca < kk) ? G[(size_t)r * ldG + ca] : T(0);
            sb[rr][cc] = (ok && cb < kk) ? G[(size_t)r * ldG + cb] : T(0);
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
    T *out = partial + (size_t)slice * kk * kk;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int a = ti * TILE + ty + 16 * i, b = tj * TILE + tx + 16 * j;
            if (a < kk && b < kk) {
                out[(size_t)a * kk + b] = acc[i][j];
                if (ti != tj) out[(size_t)b * kk + a] = acc[i][j];
            }
        }
}

template <typename T>
__global__ void gram_reduce_kernel(const T *__restrict__ partial, int kk, int nslices, T *__restrict__ gram)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kk * kk) return;
    const int a = i / kk, b = i % kk;
    // take the value from the upper triangle for both (a,b) and (b,a): exactly symmetric output
    const int src = (a <= b) ? i : b * kk + a;
    T s = T(0);
    for (int sl = 0; sl < nslices; sl++) s += partial[(size_t)sl * kk * kk + src];
    gram[i] = s;
}

int slices_for(int_t rows, int kk)
{
    const int ntile = (kk + TILE - 1) / TILE;
    const int npairs = ntile * (ntile + 1) / 2;
    int s = (2 * 148 + npairs - 1) / npairs;
    if (s > MAX_SLICES) s = MAX_SLICES;
    const long long max_by_rows = ((long long)rows + 4 * PANEL - 1) / (4 * PANEL);
    if (s > max_by_rows) s = (int)max_by_rows;
    if (s < 1) s = 1;
    return s;
}

}  // namespace

size_t gram_workspace_elems(int kk) { return (size_t)MAX_SLICES * kk * kk; }

int launch_gram(const real_t *G, int ldG, int_t rows, int kk, real_t *gram, real_t *workspace, cudaStream_t stream)
{
    if (kk < 1) return 2;
    const int total = kk * kk;
    {
        // tensor-core partials where the shape is covered, same fixed-order reduction
        static const bool use_tc = [] { const char *e = std::getenv("CMFB200_GRAM_TC"); return !e || std::atoi(e) != 0; }();
        int ns_tc = 0;
        const int rc = use_tc ? launch_gram_partials_tc(G, ldG, rows, kk, MAX_SLICES, workspace, &ns_tc, stream) : 3;
        if (rc == 0) {
            gram_reduce_kernel<real_t><<<(total + 255) / 256, 256, 0, stream>>>(workspace, kk, ns_tc, gram);
            return cudaGetLastError() == cudaSuccess ? 0 : 1;
        }
        if (rc != 3) return rc;
    }
    const int ntile = (kk + TILE - 1) / TILE;
    const int npairs = ntile * (ntile + 1) / 2;
    const int ns = slices_for(rows, kk);
    dim3 grid(npairs, ns);
    gram_partial_kernel<real_t><<<grid, 256, 0, stream>>>(G, ldG, rows, kk, ntile, ns, workspace);
    gram_reduce_kernel<real_t><<<(total + 255) / 256, 256, 0, stream>>>(workspace, kk, ns, gram);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace cmfb200
