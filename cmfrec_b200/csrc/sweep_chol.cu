// Exact half-sweeps: per row, build the normal equations from the row's stored entries and solve them with a
// Cholesky factorisation, all inside one thread block.
//
//   explicit:  M = sum_e g_e g_e^T + diag(lam .. lam, lam_last),  rhs = sum_e x_e g_e      (g_e = opposing row,
//              extended by a 1 when the solved side has a bias; x_e already reduced by the opposing bias)
//              reference factors_closed_form, sparse branch, src/common.c:978-1013 + 1058-1070
//   implicit:  M = G^T G + lam I + sum_e x_e g_e g_e^T,            rhs = sum_e (x_e + 1) g_e
//              reference factors_implicit_chol src/common.c:2063-2126 (G^T G + lam I prepared by
//              optimizeA_implicit src/common.c:3328-3335)
//
// The kd x kd matrix (kd = k [+1]) is accumulated in registers as 4x4 tiles of its upper triangle, spread over
// the 256 threads of the block (each stored entry is staged once in shared memory, 16 at a time, and read by
// every thread), then written to shared memory and factorised there (right-looking, two barriers per column);
// the two triangular solves are done by one warp.  Rows without entries are left as the reference leaves them.
#include "cg_row.cuh"
#include <cstdlib>

namespace cmfb200 {

namespace {

// four consecutive elements from 16-byte aligned shared memory with vector loads
__device__ __forceinline__ void load4(const float *p, float (&v)[4])
{
    const float4 t = *reinterpret_cast<const float4 *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const double *p, double (&v)[4])
{
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

constexpr int NB = 32;       // stored entries staged per round

// NT threads per block, TPT register tiles per thread; MG: the kd x kd matrix does not fit in shared memory
// and lives in a per-block slice of a global workspace instead (it stays in L2).
template <typename T, int NT, int TPT, int MODEL, bool MG>
__global__ void __launch_bounds__(NT) chol_sweep_kernel(const CgSweepParams p, int kd, int kdp, int ntile_rows,
                                                        T *__restrict__ workspace)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *M = MG ? workspace + (size_t)blockIdx.x * kd * kdp : reinterpret_cast<T *>(smem_raw);   // [kd][kdp]
    T *gs = MG ? reinterpret_cast<T *>(smem_raw) : M + (size_t)kd * kdp;   // [NB][kdp] staged opposing rows
    T *wgt = gs + NB * kdp;                          // [NB]        matrix weight of each staged entry
    T *xr = wgt + NB;                                // [NB]        right-hand-side weight
    T *rhs = xr + NB;                                // [kdp]
    T *colbuf = rhs + kdp;                           // [kdp]       current column of L
    T *diag = colbuf + kdp;                          // [kdp]       diagonal of L
    unsigned short *tile_map = reinterpret_cast<unsigned short *>(diag + kdp);   // [ntiles][2]

    constexpr bool IMPLICIT = MODEL == kModelImplicit;
    const int tid = threadIdx.x;
    const int kk = p.kk;
    const bool hb = !IMPLICIT && p.solve_bias;
    const int ntiles = ntile_rows * (ntile_rows + 1) / 2;
    for (int t = tid; t < ntiles; t += NT) {
        int ti = 0, rem = t;
        while (rem >= ntile_rows - ti) { rem -= ntile_rows - ti; ti++; }
        tile_map[2 * t] = (unsigned short)ti;
        tile_map[2 * t + 1] = (unsigned short)(ti + rem);
    }
    for (int i = tid; i < NB * kdp; i += NT) gs[i] = T(0);   // columns beyond the copied pieces stay zero
    __syncthreads();
    int my_ti[TPT], my_tj[TPT];
#pragma unroll
    for (int s = 0; s < TPT; s++) {
        const int t = tid + s * NT;
        my_ti[s] = t < ntiles ? tile_map[2 * t] : -1;
        my_tj[s] = t < ntiles ? tile_map[2 * t + 1] : 0;
    }

    for (int slot = blockIdx.x; slot < p.plan.n_rows; slot += gridDim.x) {
        const int row = p.plan.order[slot];
        const size_t beg = p.X.ptr[row];
        const int nnz = (int)(p.X.ptr[row + 1] - beg);
        T *frow = p.F + (size_t)row * (size_t)p.ldF;
        if (nnz <= 0 && !(MODEL != kModelExplicit && p.solve_all_rows)) {
            if (IMPLICIT || MODEL == kModelCollective) {
                // implicit: A := 0 up front (src/common.c:3334); collective without any information: zeroed too
                for (int c = tid; c < kk; c += NT) frow[c] = T(0);
                if (MODEL == kModelCollective && hb && tid == 0) p.Fbias[row] = T(0);
            } else if (hb && p.bias_start_one && tid == 0) {
                p.Fbias[row] = T(1);                                     // see sweep_cg.cu
            }
            continue;
        }
        T lam = p.lam, lam_last = p.lam_last;
        if (!IMPLICIT && p.scale_lam && nnz > 0) {
            lam *= (T)nnz;
            if (!p.scale_bias_const) lam_last *= (T)nnz;
        }

        T acc[TPT][4][4];
#pragma unroll
        for (int s = 0; s < TPT; s++)
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[s][i][j] = T(0);
        T rhs_acc = T(0);

        for (int e0 = 0; e0 < nnz; e0 += NB) {
            const int nb = min(NB, nnz - e0);
            // ---- stage nb opposing rows (coalesced along the row)
            if (tid < nb) {
                const int col = p.X.idx[beg + e0 + tid];
                const T x = p.X.val[beg + e0 + tid];
                if (IMPLICIT) {
                    wgt[tid] = x;
                    xr[tid] = x + T(1);
                } else {
                    const T ob = p.center_opp ? __ldg(p.Gbias + col) : T(0);
                    wgt[tid] = T(1);
                    xr[tid] = x - ob;
                }
            }
            // 16 threads per staged row, 16-byte pieces (the padding columns of G are zero and its rows are whole
            // 128-byte lines, so whole pieces can be copied; the bias column, which may share the last piece, is patched
            // in the register before the store)
            {
                constexpr int VN = 16 / (int)sizeof(T);
                typedef typename VecOf<T>::type Vec;
                const int ppr = (kk + VN - 1) / VN;
                for (int b = tid >> 4; b < nb; b += NT / 16) {
                    const int col = p.X.idx[beg + e0 + b];
                    const T *grow = p.G + (size_t)col * (size_t)p.ldG;
                    for (int pc = tid & 15; pc < ppr; pc += 16) {
                        Vec v = __ldg(reinterpret_cast<const Vec *>(grow) + pc);
                        if (hb && pc * VN <= kk && kk < (pc + 1) * VN) reinterpret_cast<T *>(&v)[kk - pc * VN] = T(1);
                        *reinterpret_cast<Vec *>(gs + (size_t)b * kdp + pc * VN) = v;
                    }
                }
                // the bias column when it starts a piece of its own (never written by the copy above)
                if (hb && kk % VN == 0 && tid < nb) gs[(size_t)tid * kdp + kk] = T(1);
            }
            __syncthreads();
            // ---- rank-nb update of the register tiles and of the right-hand side
            for (int b = 0; b < nb; b++) {
                const T *g = gs + b * kdp;
                const T wb = wgt[b];
#pragma unroll
                for (int s = 0; s < TPT; s++) {
                    if (my_ti[s] >= 0) {
                        T gi[4], gj[4];
                        load4(g + 4 * my_ti[s], gi);
                        load4(g + 4 * my_tj[s], gj);
                        if (IMPLICIT) {
#pragma unroll
                            for (int i = 0; i < 4; i++) gi[i] *= wb;
                        }
#pragma unroll
                        for (int i = 0; i < 4; i++)
#pragma unroll
                            for (int j = 0; j < 4; j++) acc[s][i][j] = fma(gi[i], gj[j], acc[s][i][j]);
                    }
                }
                if (tid < kd) rhs_acc = fma(xr[b], g[tid], rhs_acc);
            }
            __syncthreads();
        }

        // ---- assemble M (full symmetric) in shared memory
#pragma unroll
        for (int s = 0; s < TPT; s++) {
            if (my_ti[s] >= 0) {
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int a = 4 * my_ti[s] + i, b = 4 * my_tj[s] + j;
                        if (a < kd && b < kd) {
                            T v = acc[s][i][j];
                            if (MODEL != kModelExplicit && p.gram && a < kk && b < kk) v += p.gram[(size_t)a * kk + b];
                            if (a == b) v += ((hb || p.last_coord_special) && a == kd - 1) ? lam_last : lam;
                            if (a <= b) {
                                M[b * kdp + a] = v;   // lower triangle is what the factorisation uses
                                if (my_ti[s] != my_tj[s] || a != b) M[a * kdp + b] = v;
                            }
                        }
                    }
            }
        }
        if (tid < kd) {
            if (MODEL != kModelExplicit && p.qvec && tid < kk) rhs_acc += p.qvec[(size_t)row * (size_t)p.ldq + tid];
            rhs[tid] = rhs_acc;
        }
        __syncthreads();

        // ---- Cholesky, lower, left-looking (Cholesky-Crout): column j of L from the rows of L already known,
        //      L[i][j] = (M[i][j] - sum_{t<j} L[i][t] L[j][t]) / L[j][j].  Four adjacent lanes share one row's dot product
        //      (a quarter of the range each, two shuffles), so that one round covers NT / 4 rows; with a row stride of
        //      4 (mod 32) elements the 8 rows x 4 parts of a warp fall on 32 different banks.  One sixth of the
        //      instructions of the right-looking update with its short inner loops (profiles/README.md).
        const int warp = tid >> 5, lane = tid & 31;
        {
            constexpr int PARTS = 4;
            const int part = tid & (PARTS - 1);
            for (int j = 0; j < kd; j++) {
                const T *Lj = M + (size_t)j * kdp;
                for (int base = j; base < kd; base += NT / PARTS) {   // block-uniform
                    const int i = base + tid / PARTS;
                    const bool active = i < kd;
                    T s0 = T(0), s1 = T(0);
                    if (active) {
                        const T *Li = M + (size_t)i * kdp;
                        int t = part;
                        for (; t + PARTS < j; t += 2 * PARTS) {
                            s0 = fma(Li[t], Lj[t], s0);
                            s1 = fma(Li[t + PARTS], Lj[t + PARTS], s1);
                        }
                        if (t < j) s0 = fma(Li[t], Lj[t], s0);
                    }
                    T sum = s0 + s1;
                    sum += __shfl_xor_sync(CMF_FULL_MASK, sum, 1);
                    sum += __shfl_xor_sync(CMF_FULL_MASK, sum, 2);
                    if (active && part == 0) colbuf[i] = M[(size_t)i * kdp + j] - sum;
                }
                __syncthreads();
                const T d = sqrt(colbuf[j]);
                const T inv = T(1) / d;
                for (int i = j + 1 + tid; i < kd; i += NT) M[(size_t)i * kdp + j] = colbuf[i] * inv;
                if (tid == 0) diag[j] = d;
                __syncthreads();
            }
        }

        // ---- L y = rhs, L^T a = y  (one warp; lane owns entries lane, lane+32, ...)
        if (warp == 0) {
            for (int j = 0; j < kd; j++) {
                T yj = rhs[j] / diag[j];
                __syncwarp();
                if (lane == 0) rhs[j] = yj;
                for (int i = j + 1 + lane; i < kd; i += 32) rhs[i] = fma(-M[i * kdp + j], yj, rhs[i]);
                __syncwarp();
            }
            for (int j = kd - 1; j >= 0; j--) {
                T aj = rhs[j] / diag[j];
                __syncwarp();
                if (lane == 0) rhs[j] = aj;
                for (int i = lane; i < j; i += 32) rhs[i] = fma(-M[j * kdp + i], aj, rhs[i]);
                __syncwarp();
            }
            for (int c = lane; c < kd; c += 32) {
                if (c < kk) frow[c] = rhs[c];
                else p.Fbias[row] = rhs[c];
            }
        }
        __syncthreads();
    }
}

template <typename T> size_t chol_smem_bytes(int kd, int kdp, int ntile_rows, bool matrix_in_smem)
{
    const size_t ntiles = (size_t)ntile_rows * (ntile_rows + 1) / 2;
    return ((matrix_in_smem ? (size_t)kd * kdp : 0) + (size_t)NB * kdp + 2 * NB + 3 * (size_t)kdp) * sizeof(T) +
           ntiles * 2 * sizeof(unsigned short) + 16;
}

// per-device scratch for the MG variant, grown on demand and kept for the life of the process
template <typename T> T *chol_workspace(size_t elems)
{
    thread_local T *buf[64] = {nullptr};   // per host thread: concurrent fits never share scratch
    thread_local size_t cap[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return nullptr;
    if (cap[dev] < elems) {
        if (buf[dev]) cudaFree(buf[dev]);
        buf[dev] = nullptr;
        cap[dev] = 0;
        if (cudaMalloc((void **)&buf[dev], elems * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        cap[dev] = elems;
    }
    return buf[dev];
}

template <typename T, int NT, int TPT, int MODEL, bool MG>
int launch_tpt(const CgSweepParams &p, int kd, int kdp, int ntile_rows, cudaStream_t stream)
{
    const size_t smem = chol_smem_bytes<T>(kd, kdp, ntile_rows, !MG);
    auto kern = chol_sweep_kernel<T, NT, TPT, MODEL, MG>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 2;
    }
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
    if (occ < 1) return 2;
    long long grid = (long long)sms * occ;
    if (grid > p.plan.n_rows) grid = p.plan.n_rows;
    if (grid < 1) return 0;
    T *ws = nullptr;
    if (MG) {
        ws = chol_workspace<T>((size_t)grid * kd * kdp);
        if (!ws) return 1;
    }
    kern<<<(unsigned)grid, NT, smem, stream>>>(p, kd, kdp, ntile_rows, ws);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

template <int MODEL> int dispatch_chol(const CgSweepParams &p, cudaStream_t stream)
{
    typedef real_t T;
    const int kd = p.kk + ((MODEL != kModelImplicit && p.solve_bias) ? 1 : 0);
    const int ntile_rows = (kd + 3) / 4;
    int kdp = ntile_rows * 4;
    if ((kdp & 31) == 0) kdp += 4;   // keep consecutive rows of M off the same banks
    const int ntiles = ntile_rows * (ntile_rows + 1) / 2;
    const bool in_smem = chol_smem_bytes<T>(kd, kdp, ntile_rows, true) <= 220 * 1024;
    const int tpt256 = (ntiles + 255) / 256;
    const int tpt512 = (ntiles + 511) / 512;
    if (in_smem) {
        if (ntiles <= 160) return launch_tpt<T, 160, 1, MODEL, false>(p, kd, kdp, ntile_rows, stream);   // k = 64 (+ bias): 153 tiles
        if (tpt256 <= 1) return launch_tpt<T, 256, 1, MODEL, false>(p, kd, kdp, ntile_rows, stream);
        if (tpt256 <= 2) return launch_tpt<T, 256, 2, MODEL, false>(p, kd, kdp, ntile_rows, stream);
        if (tpt256 <= 3) return launch_tpt<T, 256, 3, MODEL, false>(p, kd, kdp, ntile_rows, stream);
        if (tpt512 <= 2) return launch_tpt<T, 512, 2, MODEL, false>(p, kd, kdp, ntile_rows, stream);
        if (tpt512 <= 3) return launch_tpt<T, 512, 3, MODEL, false>(p, kd, kdp, ntile_rows, stream);
#ifdef USE_FLOAT
        if (tpt512 <= 5) return launch_tpt<T, 512, 5, MODEL, false>(p, kd, kdp, ntile_rows, stream);
#endif
        return 2;
    }
    if (tpt512 <= 2) return launch_tpt<T, 512, 2, MODEL, true>(p, kd, kdp, ntile_rows, stream);
    if (tpt512 <= 3) return launch_tpt<T, 512, 3, MODEL, true>(p, kd, kdp, ntile_rows, stream);
#ifdef USE_FLOAT
    if (tpt512 <= 5) return launch_tpt<T, 512, 5, MODEL, true>(p, kd, kdp, ntile_rows, stream);
#endif
    return 2;   // fp32: k up to ~280, fp64: k up to ~215
}

}  // namespace

static bool nm_enabled()
{
    const char *e = std::getenv("CMFB200_NM");
    return e ? std::atoi(e) != 0 : true;
}

static bool dmma_enabled()
{
    const char *e = std::getenv("CMFB200_DMMA");
    return e ? std::atoi(e) != 0 : true;
}

int launch_explicit_chol_sweep(const CgSweepParams &p, cudaStream_t stream)
{
    if (nm_enabled() && !p.last_coord_special) {
        const int rc = launch_explicit_chol_sweep_nm(p, stream);
        if (rc != 3) return rc;
    }
    if (dmma_enabled()) {
        const int rc = launch_explicit_chol_sweep_dmma(p, stream);
        if (rc != 3) return rc;
    }
    return (p.gram || p.qvec || p.solve_all_rows) ? dispatch_chol<kModelCollective>(p, stream)
                                                  : dispatch_chol<kModelExplicit>(p, stream);
}
int launch_implicit_chol_sweep(const CgSweepParams &p, cudaStream_t stream)
{
    if (nm_enabled()) {
        const int rc = launch_implicit_chol_sweep_nm(p, stream);
        if (rc != 3) return rc;
    }
    if (dmma_enabled()) {
        const int rc = launch_implicit_chol_sweep_dmma(p, stream);
        if (rc != 3) return rc;
    }
    return dispatch_chol<kModelImplicit>(p, stream);
}

}  // namespace cmfb200
