// Conjugate-gradient half-sweeps on the GPU.
//
// One launch updates every row of one factor matrix.  A row is solved either by one warp or, for the
// few very long rows, by a whole thread block.  Inside a warp the 32 lanes are split into 32/L groups of
// L lanes; a group owns one stored entry (nonzero) of the row at a time and its L lanes each hold C of the
// k latent coordinates, so that the opposing-factor row of that entry is fetched with 16-byte loads that
// are contiguous across the group.  The per-entry dot product is an L-lane shuffle reduction, the axpy is
// lane-local, and the per-pass sums over entries are combined across groups (shuffles) and, for
// block-per-row, across warps (shared memory).  All CG vectors live in registers.
//
// Arithmetic follows the reference's single-row solvers step for step:
//   explicit:  factors_explicit_cg   reference src/common.c:1098-1188 (called from factors_closed_form
//              :631 with the scale_lam rule of :679-723, under optimizeA Case 4 :3259-3299)
//   implicit:  factors_implicit_cg   reference src/common.c:1914-1986 (under optimizeA_implicit :3349)
// including the absolute thresholds 1e-12 / 1e-8 on ||r||^2, warm start from the current row, the bias
// coordinate handled as the last coordinate with its own regulariser (lam_last), and the way the
// reference's implicit residual is written (coefficient -(c-1)x - c).  The subtraction of the opposing
// side's bias from x, which the reference does by rewriting the CSR values before every half-sweep
// (src/collective.c:8566-8571, 8750-8755), is fused into the gather: the opposing row carries its bias in
// the slot after its k coordinates.
#include "cg_row.cuh"
#include <cstdlib>

namespace cmfb200 {

namespace {

constexpr int kWarpsPerBlock = 8;

// Entries of the row read straight from global memory (L2 for the opposing factor): every pass re-gathers.
// The team's warps take 32-entry chunks round-robin; inside a warp each of the 32/L groups owns one entry at a time.
template <typename T, int C, int L, int TW> struct DirectGather {
    static constexpr int G = 32 / L;
    const CgSweepParams &p;
    size_t beg;
    int nnz, lane, wt, g, l;
    int first_chunk, chunk_stride;   // which 32-entry chunks this warp takes

    __device__ __forceinline__ DirectGather(const CgSweepParams &p_, int warp_in_team)
        : p(p_), beg(0), nnz(0), wt(warp_in_team), first_chunk(warp_in_team), chunk_stride(TW)
    {
        lane = threadIdx.x & 31;
        g = lane / L;
        l = lane % L;
    }

    template <int KIND>
    __device__ __forceinline__ void pass(const T (&vec)[C], T vecb, T (&acc)[C], T &accb) const
    {
        const int nchunks = (nnz + 31) >> 5;
        const int kk = p.kk;
        for (int ch = first_chunk; ch < nchunks; ch += chunk_stride) {
            const int e = ch * 32 + lane;
            int col_r = -1;
            T x_r = T(0);
            if (e < nnz) {
                col_r = p.X.idx[beg + e];
                x_r = p.X.val[beg + e];
            }
            const int left = nnz - ch * 32;  // entries in this chunk (may exceed 32)
#pragma unroll 4
            for (int t = 0; t < L; t++) {
                if (t * G >= left) break;  // warp-uniform
                const int item = t * G + g;
                const int col = __shfl_sync(CMF_FULL_MASK, col_r, item);
                T x = __shfl_sync(CMF_FULL_MASK, x_r, item);
                const bool valid = col >= 0;
                const T *grow = p.G + (size_t)(valid ? col : 0) * (size_t)p.ldG;
                T v[C];
                gather_row<T, C, L>(grow, l, p.ldG, valid, v);
                if (KIND == kExplicitResidual && p.center_opp && valid) x -= __ldg(p.Gbias + col);
                T d0 = T(0), d1 = T(0);
#pragma unroll
                for (int j = 0; j < C; j += 2) {
                    d0 = fma(v[j], vec[j], d0);
                    if (j + 1 < C) d1 = fma(v[j + 1], vec[j + 1], d1);
                }
                T d = group_sum<L>(d0 + d1);
                d += vecb;  // opposing value of the bias coordinate is 1 (vecb is 0 when there is none)
                T coef = entry_coef<KIND>(d, x);
                if (!valid) coef = T(0);
#pragma unroll
                for (int j = 0; j < C; j++) acc[j] = fma(coef, v[j], acc[j]);
                accb += coef;
            }
        }
    }
};

template <typename T, int C, int L, int MODEL, bool GRAM_SMEM, int MINB = 2>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MINB) cg_sweep_kernel(const CgSweepParams p)
{
    typedef Layout<T, C, L> Lay;
    constexpr int W = kWarpsPerBlock;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch_team = reinterpret_cast<T *>(smem_raw);                         // block-per-row scratch
    T *scratch_warp = scratch_team + TeamScratch<T, C, L, W>::elems();         // [W] warp-per-row scratch
    T *gram_sm = scratch_warp + W * TeamScratch<T, C, L, 1>::elems();          // [kk][KP] (implicit, if it fits)
    const T *gram = p.gram;
    if constexpr (MODEL != kModelExplicit && GRAM_SMEM) {
        const int kk = p.kk;
        for (int i = threadIdx.x; i < kk * Lay::KP; i += blockDim.x) {
            const int d = i / Lay::KP, c = i % Lay::KP;
            gram_sm[i] = (c < kk) ? p.gram[(size_t)d * kk + c] : T(0);
        }
        gram = gram_sm;
        __syncthreads();
    }
    const int w = threadIdx.x >> 5;
    const int n_long = p.plan.n_long;
    const int n_short_slots = (p.plan.n_rows - n_long + W - 1) / W;
    const int n_slots = n_long + n_short_slots;
    for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        if (slot < n_long) {
            // whole block on one row
            const int row = p.plan.order[slot];
            const size_t beg = p.X.ptr[row];
            const int nnz = (int)(p.X.ptr[row + 1] - beg);
            CgRow<T, C, L, MODEL, W, GRAM_SMEM> s(p, scratch_team, gram, w, 0);
            if (nnz > 0 || (MODEL != kModelExplicit && p.solve_all_rows)) {
                DirectGather<T, C, L, W> gat(p, w);
                gat.beg = beg; gat.nnz = nnz;
                s.solve(row, nnz, gat);
            } else {
                s.empty_row(row);
            }
            __syncthreads();
        } else {
            const int i = n_long + (slot - n_long) * W + w;
            if (i < p.plan.n_rows) {
                const int row = p.plan.order[i];
                const size_t beg = p.X.ptr[row];
                const int nnz = (int)(p.X.ptr[row + 1] - beg);
                CgRow<T, C, L, MODEL, 1, GRAM_SMEM> s(p, scratch_warp + w * TeamScratch<T, C, L, 1>::elems(), gram, 0, 0);
                if (nnz > 0 || (MODEL != kModelExplicit && p.solve_all_rows)) {
                    DirectGather<T, C, L, 1> gat(p, 0);
                    gat.beg = beg; gat.nnz = nnz;
                    s.solve(row, nnz, gat);
                } else {
                    s.empty_row(row);
                }
            }
        }
    }
}

// Rows with very many stored entries: one row per CLUSTER of kClusterSize thread blocks, so that the longest
// rows do not leave the rest of the GPU idle at the end of the sweep.  Entries are dealt over all warps of the
// cluster; the per-pass sums are combined through distributed shared memory (cg_row.cuh, CL > 1).
constexpr int kClusterSize = 8;

template <typename T, int C, int L, int MODEL>
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kWarpsPerBlock * 32)
    cg_sweep_cluster_kernel(const CgSweepParams p, int n_huge)
{
    namespace cg = cooperative_groups;
    constexpr int W = kWarpsPerBlock;
    typedef TeamScratch<T, C, L, W> Scr;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw);
    T *cl_buf = scratch + Scr::elems();                 // [2][RED_STRIDE]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int w = threadIdx.x >> 5;
    const int n_clusters = gridDim.x / kClusterSize;
    for (int slot = blockIdx.x / kClusterSize; slot < n_huge; slot += n_clusters) {
        const int row = p.plan.order[slot];
        const size_t beg = p.X.ptr[row];
        const int nnz = (int)(p.X.ptr[row + 1] - beg);
        CgRow<T, C, L, MODEL, W, false, kClusterSize> s(p, scratch, p.gram, w, 0);
        s.cl_buf = cl_buf;
        DirectGather<T, C, L, W> gat(p, w);
        gat.beg = beg; gat.nnz = nnz;
        gat.first_chunk = rank * W + w;
        gat.chunk_stride = kClusterSize * W;
        s.cl_rank = rank;   // only block 0 of the cluster writes the row back
        s.solve(row, nnz, gat);
        cluster.sync();
    }
}

template <typename T, int C, int L, int MODEL>
int launch_cluster_cfg(const CgSweepParams &p, int n_huge, cudaStream_t stream)
{
    if (n_huge <= 0) return 0;
    constexpr int W = kWarpsPerBlock;
    typedef TeamScratch<T, C, L, W> Scr;
    const size_t smem = (size_t)(Scr::elems() + 2 * Scr::RED_STRIDE) * sizeof(T);
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int clusters = n_huge;
    const int max_clusters = 2 * (sms / kClusterSize);
    if (clusters > max_clusters) clusters = max_clusters;
    auto kern = cg_sweep_cluster_kernel<T, C, L, MODEL>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<clusters * kClusterSize, W * 32, smem, stream>>>(p, n_huge);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

template <typename T, int C, int L, int MODEL, int MINB = 2>
int launch_cfg(const CgSweepParams &p_in, cudaStream_t stream)
{
    // the first n_huge rows of the (degree-sorted) order go to the cluster kernel, the rest to the main kernel
    CgSweepParams p = p_in;
    int n_huge = p.plan.n_huge < p.plan.n_long ? p.plan.n_huge : p.plan.n_long;
    if (n_huge > 0) {
        int rc = launch_cluster_cfg<T, C, L, MODEL>(p_in, n_huge, p_in.side_stream ? p_in.side_stream : stream);
        if (rc) return rc;
        p.plan.order += n_huge;
        p.plan.n_rows -= n_huge;
        p.plan.n_long -= n_huge;
    }
    typedef Layout<T, C, L> Lay;
    constexpr int W = kWarpsPerBlock;
    size_t smem = (size_t)(TeamScratch<T, C, L, W>::elems() + W * TeamScratch<T, C, L, 1>::elems()) * sizeof(T);
    const size_t gram_bytes = MODEL != kModelExplicit ? (size_t)p.kk * Lay::KP * sizeof(T) : 0;
    const bool gram_in_smem = MODEL != kModelExplicit && (smem + gram_bytes <= 100 * 1024);
    if (gram_in_smem) smem += gram_bytes;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_long = p.plan.n_long;
    const int n_slots = n_long + (p.plan.n_rows - n_long + W - 1) / W;
    if (n_slots <= 0) return 0;
    auto kern = gram_in_smem ? cg_sweep_kernel<T, C, L, MODEL, true, MINB> : cg_sweep_kernel<T, C, L, MODEL, false, MINB>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, W * 32, smem);
    if (occ < 1) occ = 1;
    // persistent grid: a whole number of waves (SM count x resident blocks), never more blocks than slots
    {
        const char *e = std::getenv("CMFB200_OCC");
        if (e && std::atoi(e) > 0 && std::atoi(e) < occ) occ = std::atoi(e);
    }
    long long grid = (long long)sms * occ;
    if (grid > n_slots) grid = n_slots;
    kern<<<(unsigned)grid, W * 32, smem, stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

template <int MODEL> int dispatch(const CgSweepParams &p, cudaStream_t stream)
{
    const int kk = p.kk;
    if (kk < 1) return 2;
#ifdef USE_FLOAT
    if (kk <= 16) return launch_cfg<float, 4, 4, MODEL>(p, stream);
    if (kk <= 32) return launch_cfg<float, 8, 4, MODEL>(p, stream);
    if (kk <= 64) {
        const char *e = std::getenv("CMFB200_CFG64");
        const int v = e ? std::atoi(e) : 0;
        if (v == 1) return launch_cfg<float, 8, 8, MODEL>(p, stream);
        if (v == 2) return launch_cfg<float, 4, 16, MODEL>(p, stream);
        if (v == 3) return launch_cfg<float, 16, 4, MODEL, 3>(p, stream);
        if (v == 4) return launch_cfg<float, 16, 4, MODEL, 4>(p, stream);
        if (v == 5) return launch_cfg<float, 8, 8, MODEL, 4>(p, stream);
        if (v == 6) return launch_cfg<float, 8, 8, MODEL, 3>(p, stream);
        return launch_cfg<float, 16, 4, MODEL>(p, stream);   // measured fastest: 8 entries in flight per warp
    }
    if (kk <= 128) return launch_cfg<float, 8, 16, MODEL>(p, stream);
    if (kk <= 256) return launch_cfg<float, 8, 32, MODEL>(p, stream);
    if (kk <= 512) return launch_cfg<float, 16, 32, MODEL>(p, stream);
#else
    if (kk <= 16) return launch_cfg<double, 4, 4, MODEL>(p, stream);
    if (kk <= 32) return launch_cfg<double, 4, 8, MODEL>(p, stream);
    if (kk <= 64) return launch_cfg<double, 4, 16, MODEL>(p, stream);
    if (kk <= 128) return launch_cfg<double, 4, 32, MODEL>(p, stream);
    if (kk <= 256) return launch_cfg<double, 8, 32, MODEL>(p, stream);
    if (kk <= 512) return launch_cfg<double, 16, 32, MODEL>(p, stream);
#endif
    return 2;
}

}  // namespace

int max_supported_k() { return 512; }

int launch_explicit_cg_sweep(const CgSweepParams &p, cudaStream_t stream)
{
    return (p.gram || p.qvec || p.solve_all_rows) ? dispatch<kModelCollective>(p, stream) : dispatch<kModelExplicit>(p, stream);
}
int launch_implicit_cg_sweep(const CgSweepParams &p, cudaStream_t stream) { return dispatch<kModelImplicit>(p, stream); }

}  // namespace cmfb200
