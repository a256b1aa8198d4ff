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
#include "sweep.h"
#include "device_utils.cuh"
#include <cstdlib>

namespace cmfb200 {

namespace {

constexpr int kWarpsPerBlock = 8;

template <typename T, int C, int L> struct Layout {
    static constexpr int VN = (C % VecOf<T>::N == 0) ? VecOf<T>::N : 1;
    static constexpr int KP = C * L;  // padded number of coordinates handled by a group
    // column owned by lane-in-group l, register j
    __device__ __forceinline__ static int col(int l, int j) { return ((j / VN) * L + l) * VN + (j % VN); }
};

// fetch this lane's C coordinates of one opposing row (columns >= ld read as zero)
template <typename T, int C, int L>
__device__ __forceinline__ void gather_row(const T *row, int l, int ld, bool valid, T (&v)[C])
{
    typedef Layout<T, C, L> Lay;
    if constexpr (Lay::VN > 1) {
#pragma unroll
        for (int q = 0; q < C / Lay::VN; q++) {
            const int c = (q * L + l) * Lay::VN;
            if (valid && c < ld) {
                ldg_vec(row + c, &v[q * Lay::VN]);
            } else {
#pragma unroll
                for (int e = 0; e < Lay::VN; e++) v[q * Lay::VN + e] = T(0);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < C; j++) {
            const int c = j * L + l;
            v[j] = (valid && c < ld) ? __ldg(row + c) : T(0);
        }
    }
}

enum PassKind { kExplicitResidual = 0, kExplicitAp = 1, kImplicitResidual = 2, kImplicitAp = 3 };

template <typename T, int C, int L, bool IMPLICIT, bool TEAM, bool GRAM_SMEM> struct RowSolver {
    typedef Layout<T, C, L> Lay;
    static constexpr int G = 32 / L;
    static constexpr int W = kWarpsPerBlock;
    static constexpr int RED_STRIDE = Lay::KP + 4;

    const CgSweepParams &p;
    T *red;          // TEAM: [2][W][RED_STRIDE]
    T *vec_sm;       // implicit: KP entries, per team (TEAM) or per warp
    const T *gram;   // implicit: shared-memory copy (row stride KP) or global (row stride kk)
    int lane, w, g, l;
    int phase;

    __device__ __forceinline__ RowSolver(const CgSweepParams &p_, T *red_, T *vec_sm_, const T *gram_)
        : p(p_), red(red_), vec_sm(vec_sm_), gram(gram_), phase(0)
    {
        lane = threadIdx.x & 31;
        w = threadIdx.x >> 5;
        g = lane / L;
        l = lane % L;
    }

    __device__ __forceinline__ void team_sync() const
    {
        if constexpr (TEAM) __syncthreads(); else __syncwarp();
    }

    // acc[j] (and accb) hold this lane's partial sums over the entries its group processed; on return every
    // lane of the team holds the totals.
    __device__ __forceinline__ void combine(T (&acc)[C], T &accb)
    {
#pragma unroll
        for (int j = 0; j < C; j++) acc[j] = across_groups_sum<L>(acc[j]);
        accb = across_groups_sum<L>(accb);
        if constexpr (TEAM) {
            T *buf = red + phase * (W * RED_STRIDE);
            if (g == 0) {
#pragma unroll
                for (int j = 0; j < C; j++) buf[w * RED_STRIDE + Lay::col(l, j)] = acc[j];
                if (l == 0) buf[w * RED_STRIDE + Lay::KP] = accb;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < C; j++) {
                T s = T(0);
#pragma unroll
                for (int ww = 0; ww < W; ww++) s += buf[ww * RED_STRIDE + Lay::col(l, j)];
                acc[j] = s;
            }
            T sb = T(0);
#pragma unroll
            for (int ww = 0; ww < W; ww++) sb += buf[ww * RED_STRIDE + Lay::KP];
            accb = sb;
            phase ^= 1;
        }
    }

    __device__ __forceinline__ T dot_full(const T (&x)[C], const T (&y)[C], T xb, T yb) const
    {
        T s = T(0);
#pragma unroll
        for (int j = 0; j < C; j++) s = fma(x[j], y[j], s);
        s = group_sum<L>(s);
        return fma(xb, yb, s);
    }

    // one pass over the stored entries of the row: acc += sum_e coef_e * g_e, accb += sum_e coef_e
    template <int KIND>
    __device__ __forceinline__ void sparse_pass(size_t beg, int nnz, const T (&vec)[C], T vecb, T (&acc)[C], T &accb) const
    {
        const int nchunks = (nnz + 31) >> 5;
        const int first = TEAM ? w : 0;
        const int stride = TEAM ? W : 1;
        const int kk = p.kk;
        for (int ch = first; ch < nchunks; ch += stride) {
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
                const T x = __shfl_sync(CMF_FULL_MASK, x_r, item);
                const bool valid = col >= 0;
                const T *grow = p.G + (size_t)(valid ? col : 0) * (size_t)p.ldG;
                T v[C];
                gather_row<T, C, L>(grow, l, p.ldG, valid, v);
                T ob = T(0);
                if (p.center_opp && valid) ob = __ldg(grow + kk);
                T d = T(0);
#pragma unroll
                for (int j = 0; j < C; j++) d = fma(v[j], vec[j], d);
                d = group_sum<L>(d);
                d += vecb;  // opposing value of the bias coordinate is 1 (vecb is 0 when there is none)
                T coef;
                if constexpr (KIND == kExplicitResidual) coef = (x - ob) - d;
                else if constexpr (KIND == kExplicitAp) coef = d;
                else if constexpr (KIND == kImplicitResidual) coef = -(d - T(1)) * x - d;
                else coef = d * (x - T(1)) + d;
                if (!valid) coef = T(0);
#pragma unroll
                for (int j = 0; j < C; j++) acc[j] = fma(coef, v[j], acc[j]);
                accb += coef;
            }
        }
    }

    // acc += sign * gram * vec, rows of gram distributed over the groups of the team
    __device__ __forceinline__ void gram_matvec(const T (&vec)[C], T sign, T (&acc)[C])
    {
        const int kk = p.kk;
        team_sync();  // previous readers of vec_sm are done
        if (g == 0 && (!TEAM || w == 0)) {
#pragma unroll
            for (int j = 0; j < C; j++) vec_sm[Lay::col(l, j)] = vec[j];
        }
        team_sync();
        const int ngroups = TEAM ? W * G : G;
        const int gg = TEAM ? w * G + g : g;
        for (int d = gg; d < kk; d += ngroups) {
            const T s = sign * vec_sm[d];
            if constexpr (GRAM_SMEM) {
                const T *mrow = gram + (size_t)d * Lay::KP;
#pragma unroll
                for (int j = 0; j < C; j++) acc[j] = fma(mrow[Lay::col(l, j)], s, acc[j]);
            } else {
                const T *mrow = gram + (size_t)d * kk;
#pragma unroll
                for (int j = 0; j < C; j++) {
                    const int c = Lay::col(l, j);
                    if (c < kk) acc[j] = fma(__ldg(mrow + c), s, acc[j]);
                }
            }
        }
    }

    __device__ void solve(int row)
    {
        const size_t beg = p.X.ptr[row];
        const int nnz = (int)(p.X.ptr[row + 1] - beg);
        const int kk = p.kk;
        T *frow = p.F + (size_t)row * (size_t)p.ldF;
        if (nnz <= 0) {
            // rows without entries are skipped by the reference and keep whatever their storage holds, which
            // for the bias column is the 1.0 written there before the sweep (src/collective.c:8538-8542)
            if (!IMPLICIT && p.solve_bias && p.bias_start_one && lane == 0 && (!TEAM || w == 0)) frow[kk] = T(1);
            return;
        }

        T a[C], r[C], pv[C], acc[C];
        T ab = T(0), rb = T(0), pb = T(0), accb = T(0);
#pragma unroll
        for (int j = 0; j < C; j++) {
            const int c = Lay::col(l, j);
            a[j] = (c < kk) ? frow[c] : T(0);
        }
        const bool hb = !IMPLICIT && p.solve_bias;
        if (hb) ab = p.bias_start_one ? T(1) : frow[kk];

        T lam = p.lam, lam_last = p.lam_last;
        if (!IMPLICIT && p.scale_lam) {
            lam *= (T)nnz;
            if (!p.scale_bias_const) lam_last *= (T)nnz;
        }

        // ---- residual at the starting point
#pragma unroll
        for (int j = 0; j < C; j++) acc[j] = T(0);
        accb = T(0);
        if constexpr (IMPLICIT) gram_matvec(a, T(-1), acc);
        sparse_pass<IMPLICIT ? kImplicitResidual : kExplicitResidual>(beg, nnz, a, ab, acc, accb);
        combine(acc, accb);
#pragma unroll
        for (int j = 0; j < C; j++) {
            const int c = Lay::col(l, j);
            r[j] = (c < kk) ? fma(-lam, a[j], acc[j]) : T(0);
        }
        if (hb) {
            rb = fma(-lam, ab, accb);
            if (lam != lam_last) rb -= (lam_last - lam) * ab;
        }
        T r_old = dot_full(r, r, rb, rb);
        bool changed = false;
        if (!(r_old <= T(1e-12))) {
#pragma unroll
            for (int j = 0; j < C; j++) pv[j] = r[j];
            pb = rb;
            for (int step = 0; step < p.max_cg_steps; step++) {
#pragma unroll
                for (int j = 0; j < C; j++) acc[j] = T(0);
                accb = T(0);
                if constexpr (IMPLICIT) gram_matvec(pv, T(1), acc);
                sparse_pass<IMPLICIT ? kImplicitAp : kExplicitAp>(beg, nnz, pv, pb, acc, accb);
                combine(acc, accb);
#pragma unroll
                for (int j = 0; j < C; j++) {
                    const int c = Lay::col(l, j);
                    acc[j] = (c < kk) ? fma(lam, pv[j], acc[j]) : T(0);
                }
                if (hb) {
                    accb = fma(lam, pb, accb);
                    if (lam != lam_last) accb += (lam_last - lam) * pb;
                } else {
                    accb = T(0);
                }
                const T alpha = r_old / dot_full(pv, acc, pb, accb);
#pragma unroll
                for (int j = 0; j < C; j++) {
                    a[j] = fma(alpha, pv[j], a[j]);
                    r[j] = fma(-alpha, acc[j], r[j]);
                }
                ab = fma(alpha, pb, ab);
                rb = fma(-alpha, accb, rb);
                changed = true;
                const T r_new = dot_full(r, r, rb, rb);
                if (r_new <= T(1e-8)) break;
                const T beta = r_new / r_old;
#pragma unroll
                for (int j = 0; j < C; j++) pv[j] = fma(beta, pv[j], r[j]);
                pb = fma(beta, pb, rb);
                r_old = r_new;
            }
        }
        // A row that exits before the first step is left exactly as it was, except that a bias coordinate
        // restarted from 1.0 is what the reference leaves in the matrix.
        if (g == 0 && (!TEAM || w == 0)) {
            if (changed) {
#pragma unroll
                for (int j = 0; j < C; j++) {
                    const int c = Lay::col(l, j);
                    if (c < kk) frow[c] = a[j];
                }
            }
            if (hb && l == 0 && (changed || p.bias_start_one)) frow[kk] = ab;
        }
        if constexpr (TEAM) __syncthreads();
    }
};

template <typename T, int C, int L, bool IMPLICIT, bool GRAM_SMEM>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) cg_sweep_kernel(const CgSweepParams p)
{
    typedef Layout<T, C, L> Lay;
    constexpr int W = kWarpsPerBlock;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *red = reinterpret_cast<T *>(smem_raw);                                  // [2][W][KP+4]
    T *vec_sm = red + 2 * W * (Lay::KP + 4);                                   // [W][KP]
    T *gram_sm = vec_sm + W * Lay::KP;                                         // [kk][KP] (implicit, if it fits)
    const T *gram = p.gram;
    if constexpr (IMPLICIT && GRAM_SMEM) {
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
            RowSolver<T, C, L, IMPLICIT, true, GRAM_SMEM> s(p, red, vec_sm, gram);
            s.solve(p.plan.order[slot]);
        } else {
            const int i = n_long + (slot - n_long) * W + w;
            if (i < p.plan.n_rows) {
                RowSolver<T, C, L, IMPLICIT, false, GRAM_SMEM> s(p, red, vec_sm + w * Lay::KP, gram);
                s.solve(p.plan.order[i]);
            }
        }
    }
}

template <typename T, int C, int L, bool IMPLICIT>
int launch_cfg(const CgSweepParams &p, cudaStream_t stream)
{
    typedef Layout<T, C, L> Lay;
    constexpr int W = kWarpsPerBlock;
    size_t smem = (size_t)(2 * W * (Lay::KP + 4) + W * Lay::KP) * sizeof(T);
    const size_t gram_bytes = IMPLICIT ? (size_t)p.kk * Lay::KP * sizeof(T) : 0;
    const bool gram_in_smem = IMPLICIT && (smem + gram_bytes <= 100 * 1024);
    if (gram_in_smem) smem += gram_bytes;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_long = p.plan.n_long;
    const int n_slots = n_long + (p.plan.n_rows - n_long + W - 1) / W;
    if (n_slots <= 0) return 0;
    auto kern = gram_in_smem ? cg_sweep_kernel<T, C, L, IMPLICIT, true> : cg_sweep_kernel<T, C, L, IMPLICIT, false>;
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

template <bool IMPLICIT> int dispatch(const CgSweepParams &p, cudaStream_t stream)
{
    const int kk = p.kk;
    if (kk < 1) return 2;
#ifdef USE_FLOAT
    if (kk <= 16) return launch_cfg<float, 4, 4, IMPLICIT>(p, stream);
    if (kk <= 32) return launch_cfg<float, 8, 4, IMPLICIT>(p, stream);
    if (kk <= 64) {
        const char *e = std::getenv("CMFB200_CFG64");
        const int v = e ? std::atoi(e) : 0;
        if (v == 1) return launch_cfg<float, 8, 8, IMPLICIT>(p, stream);
        if (v == 2) return launch_cfg<float, 4, 16, IMPLICIT>(p, stream);
        return launch_cfg<float, 16, 4, IMPLICIT>(p, stream);   // measured fastest: 8 entries in flight per warp
    }
    if (kk <= 128) return launch_cfg<float, 8, 16, IMPLICIT>(p, stream);
    if (kk <= 256) return launch_cfg<float, 8, 32, IMPLICIT>(p, stream);
    if (kk <= 512) return launch_cfg<float, 16, 32, IMPLICIT>(p, stream);
#else
    if (kk <= 16) return launch_cfg<double, 4, 4, IMPLICIT>(p, stream);
    if (kk <= 32) return launch_cfg<double, 4, 8, IMPLICIT>(p, stream);
    if (kk <= 64) return launch_cfg<double, 4, 16, IMPLICIT>(p, stream);
    if (kk <= 128) return launch_cfg<double, 4, 32, IMPLICIT>(p, stream);
    if (kk <= 256) return launch_cfg<double, 8, 32, IMPLICIT>(p, stream);
    if (kk <= 512) return launch_cfg<double, 16, 32, IMPLICIT>(p, stream);
#endif
    return 2;
}

}  // namespace

int max_supported_k() { return 512; }

int launch_explicit_cg_sweep(const CgSweepParams &p, cudaStream_t stream) { return dispatch<false>(p, stream); }
int launch_implicit_cg_sweep(const CgSweepParams &p, cudaStream_t stream) { return dispatch<true>(p, stream); }

}  // namespace cmfb200
