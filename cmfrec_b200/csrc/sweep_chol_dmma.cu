// Exact (Cholesky) half-sweep of the fp64 library with every row's normal matrix AND its factorisation on the FP64
// tensor cores (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4 -- the native shape of sm_100a).
//
//   explicit:  M = sum_e g_e g_e^T + diag(lam .. lam, lam_last),  rhs = sum_e x_e g_e
//              reference factors_closed_form, sparse branch, src/common.c:978-1013 + 1058-1070
//   implicit:  M = G^T G + lam I + sum_e x_e g_e g_e^T,            rhs = sum_e (x_e + 1) g_e
//              reference factors_implicit_chol src/common.c:2063-2126
//   collective: + Q, + q_i   (reference collective_closed_form_block, src/collective.c:1223-1847)
//
// The matrix is EXTENDED by the right-hand side as one more row (row kd), so the forward substitution L y = rhs comes
// out of the factorisation as row kd of L.  Its lower triangle is cut into 8 x 8 tiles (NTR tile rows).  Two kernels
// per batch of rows, connected by a tile workspace in global memory (it mostly lives in L2):
//
//   BUILD  (one thread block per row, 2 per SM at k = 128): the row's opposing rows are staged 32 at a time (cp.async,
//      16 bytes, double buffered) into shared memory as [entry][column] with a row stride of 4 (mod 16) doubles, which
//      makes the A and the B fragment loads (4 entries x 8 columns per warp) conflict-free; one DMMA per tile and 4
//      entries, the tiles being the accumulator fragments in registers (warp w owns tile rows w and NTR-1-w);
//      regulariser, constant matrix and per-row vector are added in the registers; tiles are written as they sit in
//      the accumulators.  Bound by the tensor pipe.
//   FACTOR (one WARP per matrix, no block-level synchronisation at all, many matrices in flight per SM): blocked
//      left-looking Cholesky by tile columns.  Column j: load its tiles, subtract L(i,p) L(j,p)^T for the finished
//      columns p < j (2 DMMAs per tile pair, operands streamed from the workspace in operand-fragment order),
//      factorise the 8 x 8 diagonal tile with shuffles (one reciprocal square root per pivot) and invert it, panel
//      solve as a product with the inverse (2 DMMAs per tile), store.  Then the blocked backward substitution
//      L^T a = y, again as DMMAs on vectors, and the write-back.
//
// Roofline: FP64 tensor throughput, nnz * (kd+1)^2 + rows * (kd+1)^3 / 3 flop per half-sweep against the measured
// DGEMM peak (profiles/r2_fp_peaks.json).  See DESIGN.md.
#include "cg_row.cuh"
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace cmfb200 {

#ifndef USE_FLOAT

namespace {

constexpr int DM_NB = 32;    // stored entries staged per chunk

__device__ __forceinline__ void dm_mma(double (&c)[2], double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}
__device__ __forceinline__ void dm_cp16(double *dst, const double *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void dm_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void dm_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NTR> struct DmCfg {
    static constexpr int W = (NTR + 1) / 2;          // warps
    static constexpr int NT = W * 32;
    static constexpr int ROWS = NTR * 8;
    static constexpr int LD = NTR * 8 + 4;           // staging row stride (doubles), 4 or 12 (mod 16)
    static constexpr int NTILES = NTR * (NTR + 1) / 2;
    static constexpr int USIZE = 2 * DM_NB * LD;
    static constexpr size_t smem_bytes() { return (size_t)(USIZE + 2 * DM_NB) * sizeof(double); }
    // tile (i, j), j <= i, of matrix m in the workspace: 64 doubles at
    __host__ __device__ static constexpr int tile_at(int i, int j) { return (i * (i + 1) / 2 + j) * 64; }
};

// ---- per-warp pieces, specialised on the warp's first tile row R1 so that every tile <-> register-slot relation is a
// compile-time constant: slot s holds tile (R2, s) for s <= R2 and tile (R1, NTR - s) above (R2 = NTR - 1 - R1 >= R1)
template <int NTR, int R1> struct DmWarp {
    static constexpr int R2 = NTR - 1 - R1;
    static constexpr bool TWO = R1 < R2;                          // the middle warp of an odd NTR owns one tile row only
    __host__ __device__ static constexpr bool used(int s) { return s <= R2 || TWO; }
    __host__ __device__ static constexpr bool first(int s) { return s <= R2; }
    __host__ __device__ static constexpr int trow(int s) { return s <= R2 ? R2 : R1; }
    __host__ __device__ static constexpr int tcol(int s) { return s <= R2 ? s : NTR - s; }

    // acc += sum over `ksteps` groups of 4 staged entries of  a_row (x) b_col
    template <bool IMPLICIT>
    __device__ __forceinline__ static void syrk(double (&acc)[NTR + 1][2], const double *gs, int ld, const double *wgt, int ksteps,
                                                int g, int q, int kd, bool on1, bool on2)
    {
        for (int ks = 0; ks < ksteps; ks++) {
            const double *er = gs + (4 * ks + q) * ld + g;
            double a2 = on2 ? er[8 * R2] : 0.0, a1 = (TWO && on1) ? er[8 * R1] : 0.0;
            if (IMPLICIT) {
                const double w = wgt[4 * ks + q];
                if (8 * R2 + g != kd) a2 *= w;
                if (8 * R1 + g != kd) a1 *= w;
            }
#pragma unroll
            for (int s = 0; s <= NTR; s++)
                if (used(s)) dm_mma(acc[s], first(s) ? a2 : a1, er[8 * tcol(s)]);
        }
    }
};

// run f.template operator()<R1>() for the R1 equal to `warp`
template <int NTR, int R, typename F> __device__ __forceinline__ void dm_dispatch(int warp, F &&f)
{
    if (warp == R) {
        f(DmWarp<NTR, R>());
    } else if constexpr (R + 1 < DmCfg<NTR>::W) {
        dm_dispatch<NTR, R + 1>(warp, static_cast<F &&>(f));
    }
}

constexpr int DM_SEG = 4096;   // stored entries per unit of a long row

// Long rows (the first n_split rows of a batch, rows being sorted by decreasing length) are cut into segments of DM_SEG
// entries so that no single thread block carries a 40,000-entry row: prefix[i] = first unit of row i (device array,
// n_split + 1 entries); the rows behind them are one unit each.
struct DmSplit {
    const int *prefix;
    int n_split, units_split;
    __device__ __forceinline__ void locate(int unit, int &ms, int &seg) const
    {
        if (unit >= units_split) {
            ms = n_split + (unit - units_split);
            seg = 0;
            return;
        }
        int lo = 0, hi = n_split - 1;   // last i with prefix[i] <= unit
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (prefix[mid] <= unit) lo = mid;
            else hi = mid - 1;
        }
        ms = lo;
        seg = unit - prefix[lo];
    }
    __device__ __forceinline__ void units_of(int ms, int &first, int &count) const
    {
        if (ms >= n_split) {
            first = units_split + (ms - n_split);
            count = 1;
        } else {
            first = prefix[ms];
            count = prefix[ms + 1] - first;
        }
    }
};

// ======================================================================= BUILD
// rows plan.order[slot0 .. slot0 + nslots) -> tiles of their extended normal matrices in ws[slot - slot0]
template <int NTR, int MODEL, int BPS>
__global__ void __launch_bounds__(DmCfg<NTR>::NT, BPS)
chol_dmma_build_kernel(const CgSweepParams p, int kd, int slot0, int nunits, DmSplit sp, double *__restrict__ ws)
{
    typedef DmCfg<NTR> Cfg;
    constexpr int NT = Cfg::NT, LD = Cfg::LD, ROWS = Cfg::ROWS;
    constexpr bool IMPLICIT = MODEL == kModelImplicit;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *U = reinterpret_cast<double *>(smem_raw);   // staging [2][NB][LD]
    double *wgt = U + Cfg::USIZE;                       // [2][NB] matrix weight of the staged entries (implicit)

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(CMF_FULL_MASK, tid >> 5, 0);
    const int g = lane >> 2, q = lane & 3;
    const int r1 = warp, r2 = NTR - 1 - warp;           // tile rows of this warp (r1 <= r2)
    const int kk = p.kk;
    const bool hb = !IMPLICIT && p.solve_bias;
    const int p_last = kd >> 3;                         // tile row that holds the right-hand side
    const bool row2_on = r2 <= p_last;
    const bool row1_on = r1 < r2 && r1 <= p_last;
    const int ppr = kk >> 1;                            // whole 16-byte pieces per opposing row copied asynchronously
    const bool kodd = (kk & 1) != 0;
    const int zc0 = kd + 1, zc1 = ROWS;                 // staging columns that must read as zero (every slot is read)

    for (int i = tid; i < 2 * DM_NB * (zc1 - zc0); i += NT) {
        const int b = i / (zc1 - zc0), c = zc0 + i % (zc1 - zc0);
        U[b * LD + c] = 0.0;
    }
    __syncthreads();

    // a unit = one row, or one segment of DM_SEG stored entries of a long row (their partial tiles are summed by the factor kernel)
    for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
        int ms, seg;
        sp.locate(unit, ms, seg);
        const int row = p.plan.order[slot0 + ms];
        const size_t beg = p.X.ptr[row] + (size_t)seg * DM_SEG;
        const int nnz_row = (int)(p.X.ptr[row + 1] - p.X.ptr[row]);
        // only the rows in front of the batch are cut into segments; every other row is one unit whatever its length
        const int nnz = ms < sp.n_split ? min(DM_SEG, nnz_row - seg * DM_SEG) : nnz_row;
        if (nnz_row <= 0 && !(MODEL != kModelExplicit && p.solve_all_rows)) continue;   // the factor kernel deals with it
        double lam = p.lam, lam_last = p.lam_last;
        if (!IMPLICIT && p.scale_lam && nnz_row > 0) {
            lam *= (double)nnz_row;
            if (!p.scale_bias_const) lam_last *= (double)nnz_row;
        }

        double acc[NTR + 1][2];
#pragma unroll
        for (int s = 0; s <= NTR; s++) acc[s][0] = acc[s][1] = 0.0;

        // ---- 1. normal matrix: chunks of DM_NB stored entries, double buffered
        const int nchunks = (nnz + DM_NB - 1) / DM_NB;
        double pxr = 0.0, pw = 0.0, plast = 0.0;   // the patch thread's prefetched values of the chunk in flight
        auto issue = [&](int c) {
            const int e0 = c * DM_NB, nb = min(DM_NB, nnz - e0);
            double *gs = U + (c & 1) * DM_NB * LD;
            for (int b = tid >> 4; b < nb; b += NT / 16) {
                const int col = __ldg(p.X.idx + beg + e0 + b);
                const double *src = p.G + (size_t)col * (size_t)p.ldG;
                double *dst = gs + b * LD;
                for (int pc = tid & 15; pc < ppr; pc += 16) dm_cp16(dst + 2 * pc, src + 2 * pc);
            }
            if (tid < nb) {
                const int col = __ldg(p.X.idx + beg + e0 + tid);
                const double x = __ldg(p.X.val + beg + e0 + tid);
                if (IMPLICIT) {
                    pw = x;
                    pxr = x + 1.0;
                } else {
                    pxr = x - (p.center_opp ? __ldg(p.Gbias + col) : 0.0);
                }
                if (kodd) plast = __ldg(p.G + (size_t)col * (size_t)p.ldG + (kk - 1));
            }
            dm_commit();
        };
        if (nchunks > 0) issue(0);
        for (int c = 0; c < nchunks; c++) {
            const int e0 = c * DM_NB, nb = min(DM_NB, nnz - e0);
            double *gs = U + (c & 1) * DM_NB * LD;
            const double cxr = pxr, cw = pw, clast = plast;   // chunk c's patch values, before the prefetch overwrites them
            if (c + 1 < nchunks) {
                issue(c + 1);
                dm_wait<1>();
            } else {
                dm_wait<0>();
            }
            if (tid < nb) {
                double *grow = gs + tid * LD;
                if (kodd) grow[kk - 1] = clast;
                if (hb) grow[kk] = 1.0;
                grow[kd] = cxr;
                if (IMPLICIT) wgt[(c & 1) * DM_NB + tid] = cw;
            }
            const int nb4 = (nb + 3) & ~3;
            for (int i = tid; i < (nb4 - nb) * (kd + 1); i += NT) gs[(nb + i / (kd + 1)) * LD + i % (kd + 1)] = 0.0;
            if (IMPLICIT && tid < nb4 - nb) wgt[(c & 1) * DM_NB + nb + tid] = 0.0;
            __syncthreads();
            if (row2_on || row1_on)
                dm_dispatch<NTR, 0>(warp, [&](auto wt) {
                    decltype(wt)::template syrk<IMPLICIT>(acc, gs, LD, wgt + (c & 1) * DM_NB, nb4 >> 2, g, q, kd, row1_on, row2_on);
                });
            __syncthreads();
        }

        // ---- 2. regulariser, constant matrix, per-row vector; tiles out as they sit in the accumulators
        double *wm = ws + (size_t)unit * (Cfg::NTILES * 64);
        const bool lead = seg == 0;   // regulariser, constant matrix and per-row vector go to the first segment only
#pragma unroll
        for (int s = 0; s <= NTR; s++) {
            const bool first = s < NTR && s <= r2;
            if (first ? row2_on : row1_on) {
                const int ti = first ? r2 : r1, tj = first ? s : NTR - s;
                const int a = 8 * ti + g;
                const int b0 = 8 * tj + 2 * q;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int b = b0 + e;
                    if (!lead) continue;
                    if (a < kd && b < kd) {
                        double v = acc[s][e];
                        if (MODEL != kModelExplicit && p.gram && a < kk && b < kk) v += __ldg(p.gram + (size_t)a * kk + b);
                        if (a == b) v += ((hb || p.last_coord_special) && a == kd - 1) ? lam_last : lam;
                        acc[s][e] = v;
                    } else if (MODEL != kModelExplicit && a == kd && b < kk && p.qvec) {
                        acc[s][e] += __ldg(p.qvec + (size_t)row * (size_t)p.ldq + b);
                    }
                }
                *reinterpret_cast<double2 *>(wm + Cfg::tile_at(ti, tj) + 2 * lane) = make_double2(acc[s][0], acc[s][1]);
            }
        }
    }
}

// ======================================================================= FACTOR
// Tile storage orders inside the workspace (64 doubles per tile):
//   "C order"  element (r, c) at 2 * (4 r + c / 2) + c % 2   -- the accumulator fragment: lane 4r + c/2 holds (r, 2q), (r, 2q+1)
//   "A order"  element (r, c) at 2 * (4 r + c % 4) + c / 4   -- the operand fragment: lane 4r + c%4 holds (r, q), (r, q+4)
// BUILD writes C order; FACTOR overwrites each tile below the diagonal with L in A order and each diagonal tile with
// the INVERSE of its Cholesky factor in A order.
__device__ __forceinline__ int dm_a_pos(int r, int c) { return 2 * (4 * r + (c & 3)) + (c >> 2); }

// accumulator-fragment pair (lane (g, q) holds columns 2q, 2q+1 of its row) -> operand-fragment pair (columns q, q+4)
__device__ __forceinline__ void dm_c_to_a(double c0, double c1, int lane, double &lo, double &hi)
{
    const int q = lane & 3, base = lane & ~3;
    const double l0 = __shfl_sync(CMF_FULL_MASK, c0, base | (q >> 1)), l1 = __shfl_sync(CMF_FULL_MASK, c1, base | (q >> 1));
    const double h0 = __shfl_sync(CMF_FULL_MASK, c0, base | 2 | (q >> 1)), h1 = __shfl_sync(CMF_FULL_MASK, c1, base | 2 | (q >> 1));
    lo = (q & 1) ? l1 : l0;
    hi = (q & 1) ? h1 : h0;
}

constexpr int DM_FW = 4;            // warps (= matrices in flight) per block of the factor kernel
#ifndef DM_FMINB
#define DM_FMINB 4                  // factor blocks per SM the register allocation aims at
#endif
constexpr int DM_SCR = 160;         // doubles of shared scratch per warp

template <int NTR, int MODEL>
__global__ void __launch_bounds__(DM_FW * 32, DM_FMINB)
chol_dmma_factor_kernel(const CgSweepParams p, int kd, int slot0, int nslots, DmSplit sp, double *ws)
{
    typedef DmCfg<NTR> Cfg;
    constexpr bool IMPLICIT = MODEL == kModelImplicit;
    __shared__ __align__(16) double scratch[DM_FW][DM_SCR];
    const int lane = threadIdx.x & 31;
    const int wib = __shfl_sync(CMF_FULL_MASK, threadIdx.x >> 5, 0);
    double *sc = scratch[wib];          // [8][10] raw diagonal tile, rows padded
    double *sl = sc + 80;               // [8][8]  inverse of its factor
    double *yl = sl + 64;               // [8]     right-hand-side row inside the last diagonal tile
    const int g = lane >> 2, q = lane & 3;
    const int kk = p.kk;
    const bool hb = !IMPLICIT && p.solve_bias;
    const int p_last = kd >> 3, gk = kd & 7;   // the right-hand side is row gk of tile row p_last

    for (int ms = blockIdx.x * DM_FW + wib; ms < nslots; ms += gridDim.x * DM_FW) {
        const int row = p.plan.order[slot0 + ms];
        const int nnz = (int)(p.X.ptr[row + 1] - p.X.ptr[row]);
        double *frow = p.F + (size_t)row * (size_t)p.ldF;
        if (nnz <= 0 && !(MODEL != kModelExplicit && p.solve_all_rows)) {
            if (IMPLICIT || MODEL == kModelCollective) {
                for (int c = lane; c < kk; c += 32) frow[c] = 0.0;
                if (MODEL == kModelCollective && hb && lane == 0) p.Fbias[row] = 0.0;
            } else if (hb && p.bias_start_one && lane == 0) {
                p.Fbias[row] = 1.0;
            }
            continue;
        }
        int u0, nseg;
        sp.units_of(ms, u0, nseg);
        double *wm = ws + (size_t)u0 * (Cfg::NTILES * 64);   // L overwrites the first unit's tiles

        for (int j = 0; j <= p_last; j++) {
            // ---- column j of the extended matrix
            double t[NTR][2];
#pragma unroll
            for (int i = 0; i < NTR; i++) {
                if (i >= j && i <= p_last) {
                    const double2 v = *reinterpret_cast<const double2 *>(wm + Cfg::tile_at(i, j) + 2 * lane);
                    t[i][0] = v.x;
                    t[i][1] = v.y;
                    for (int sgm = 1; sgm < nseg; sgm++) {   // the partial tiles of a long row's other segments
                        const double2 u = *reinterpret_cast<const double2 *>(wm + (size_t)sgm * (Cfg::NTILES * 64) + Cfg::tile_at(i, j) + 2 * lane);
                        t[i][0] += u.x;
                        t[i][1] += u.y;
                    }
                } else {
                    t[i][0] = t[i][1] = 0.0;
                }
            }
            // ---- minus the finished columns
            for (int pc = 0; pc < j; pc++) {
                const double2 bj = *reinterpret_cast<const double2 *>(wm + Cfg::tile_at(j, pc) + 2 * lane);
#pragma unroll
                for (int i = 0; i < NTR; i++) {
                    if (i >= j && i <= p_last) {
                        const double2 ai = *reinterpret_cast<const double2 *>(wm + Cfg::tile_at(i, pc) + 2 * lane);
                        dm_mma(t[i], -ai.x, bj.x);
                        dm_mma(t[i], -ai.y, bj.y);
                    }
                }
            }
            // ---- diagonal tile: factorise (lane r mod 8 holds row r) and invert
            __syncwarp();
#pragma unroll
            for (int i = 0; i < NTR; i++)
                if (i == j) *reinterpret_cast<double2 *>(sc + g * 10 + 2 * q) = make_double2(t[i][0], t[i][1]);
            __syncwarp();
            {
                const int r = lane & 7;
                double d[8], li[8], inv_r = 0.0;
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(sc + r * 10 + c);
                    d[c] = v.x;
                    d[c + 1] = v.y;
                }
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const double dcc = __shfl_sync(CMF_FULL_MASK, d[c], c);
                    const double inv = (8 * j + c < kd) ? rsqrt(dcc) : 0.0;
                    const double lc = d[c] * inv;
                    d[c] = lc;
                    if (r == c) inv_r = inv;
#pragma unroll
                    for (int tt = c + 1; tt < 8; tt++) {
                        const double ltc = __shfl_sync(CMF_FULL_MASK, lc, tt);
                        d[tt] = fma(-lc, ltc, d[tt]);
                    }
                }
                // inverse of the lower-triangular factor, row r in lane r:  Linv[r][c] = -inv_r * sum_{c<=t<r} L[r][t] Linv[t][c]
#pragma unroll
                for (int c = 0; c < 8; c++) li[c] = 0.0;
#pragma unroll
                for (int tt = 0; tt < 8; tt++) {
                    // row tt is final once the rows above it have been folded in
                    if (r == tt) {
#pragma unroll
                        for (int c = 0; c < 8; c++) li[c] = (c < tt) ? -inv_r * li[c] : ((c == tt) ? inv_r : 0.0);
                    }
#pragma unroll
                    for (int c = 0; c <= tt; c++) {
                        const double v = __shfl_sync(CMF_FULL_MASK, li[c], tt);
                        if (r > tt) li[c] = fma(d[tt], v, li[c]);
                    }
                }
                if (lane < 8) {
#pragma unroll
                    for (int c = 0; c < 8; c++) sl[r * 8 + c] = li[c];
                    if (j == p_last && r == gk) {
#pragma unroll
                        for (int c = 0; c < 8; c++) yl[c] = (c < gk) ? d[c] : 0.0;
                    }
                }
            }
            __syncwarp();
            const double b_lo = sl[g * 8 + q], b_hi = sl[g * 8 + q + 4];   // B[k][n] = Linv[n][k]
            // the diagonal tile's slot keeps the inverse, in operand order
            *reinterpret_cast<double2 *>(wm + Cfg::tile_at(j, j) + 2 * lane) = make_double2(b_lo, b_hi);
            // ---- panel solve  L(i, j) = T(i, j) Linv^T,  stored in operand order
#pragma unroll
            for (int i = 0; i < NTR; i++) {
                if (i > j && i <= p_last) {
                    double a_lo, a_hi, x[2] = {0.0, 0.0};
                    dm_c_to_a(t[i][0], t[i][1], lane, a_lo, a_hi);
                    dm_mma(x, a_lo, b_lo);
                    dm_mma(x, a_hi, b_hi);
                    double *dst = wm + Cfg::tile_at(i, j);
                    dst[dm_a_pos(g, 2 * q)] = x[0];
                    dst[dm_a_pos(g, 2 * q + 1)] = x[1];
                }
            }
            __syncwarp();   // this lane's later loads of the tiles just stored by other lanes
            __threadfence_block();
        }

        // ---- L^T a = y: z holds, replicated in every lane group, the columns 2q, 2q+1 of each block of 8 unknowns
        double z[NTR][2];
#pragma unroll
        for (int it = 0; it < NTR; it++) {
            if (it < p_last) {
                const double *src = wm + Cfg::tile_at(p_last, it);
                z[it][0] = src[dm_a_pos(gk, 2 * q)];
                z[it][1] = src[dm_a_pos(gk, 2 * q + 1)];
            } else if (it == p_last) {
                z[it][0] = yl[2 * q];
                z[it][1] = yl[2 * q + 1];
            } else {
                z[it][0] = z[it][1] = 0.0;
            }
        }
        for (int pb = p_last; pb >= 0; pb--) {
            double zb0 = 0.0, zb1 = 0.0;
#pragma unroll
            for (int it = 0; it < NTR; it++)
                if (it == pb) {
                    zb0 = z[it][0];
                    zb1 = z[it][1];
                }
            // a_blk^T = z_blk^T Linv(pb):  A rows = z_blk, B[k][n] = Linv[k][n] (transposed read of the operand-order tile)
            double a_lo, a_hi, ab[2] = {0.0, 0.0};
            dm_c_to_a(zb0, zb1, lane, a_lo, a_hi);
            {
                const double *src = wm + Cfg::tile_at(pb, pb);
                dm_mma(ab, a_lo, src[dm_a_pos(q, g)]);
                dm_mma(ab, a_hi, src[dm_a_pos(q + 4, g)]);
            }
            if (g == 0) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int col = 8 * pb + 2 * q + e;
                    if (col < kk) frow[col] = ab[e];
                    else if (col < kd) p.Fbias[row] = ab[e];
                }
            }
            dm_c_to_a(ab[0], ab[1], lane, a_lo, a_hi);
            // z(it)^T -= a_blk^T L(pb, it)
#pragma unroll
            for (int it = 0; it < NTR; it++) {
                if (it < pb) {
                    const double *src = wm + Cfg::tile_at(pb, it);
                    dm_mma(z[it], -a_lo, src[dm_a_pos(q, g)]);
                    dm_mma(z[it], -a_hi, src[dm_a_pos(q + 4, g)]);
                }
            }
        }
    }
}

// per-device tile workspace, grown on demand and kept for the life of the process
double *dmma_workspace(size_t elems)
{
    thread_local double *buf[64] = {nullptr};   // per host thread: concurrent fits never share scratch
    thread_local size_t cap[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return nullptr;
    if (cap[dev] < elems) {
        if (buf[dev]) cudaFree(buf[dev]);
        buf[dev] = nullptr;
        cap[dev] = 0;
        if (cudaMalloc((void **)&buf[dev], elems * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        cap[dev] = elems;
    }
    return buf[dev];
}

int dmma_env(const char *name, int dflt)
{
    const char *e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

template <int NTR, int MODEL, int BPS> int launch_dmma(const CgSweepParams &p, int kd, cudaStream_t stream)
{
    typedef DmCfg<NTR> Cfg;
    auto build = chol_dmma_build_kernel<NTR, MODEL, BPS>;
    auto factor = chol_dmma_factor_kernel<NTR, MODEL>;
    const size_t smem = Cfg::smem_bytes();
    if (cudaFuncSetAttribute(build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 3;
    }
    int dev = 0, sms = 148, occ = 1, occf = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, build, Cfg::NT, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occf, factor, DM_FW * 32, 0);
    if (occ < 1 || occf < 1) return 3;
    const int fcap = dmma_env("CMFB200_DMMA_FBLOCKS", 8);   // factor blocks per SM: bounds the tiles in flight to what L2 holds
    if (occf > fcap) occf = fcap;
    const int n = p.plan.n_rows;
    if (n < 1) return 0;
    const size_t per = (size_t)Cfg::NTILES * 64;
    int batch = dmma_env("CMFB200_DMMA_BATCH", 16384);   // units per batch
    if (batch < 64) batch = 64;
    const int_t *deg = p.plan.host_deg;                  // descending; null: no splitting
    const bool split_on = deg && dmma_env("CMFB200_DMMA_SPLIT", 1) != 0;   // config 3 on one GPU: 32.6 ms with, 36.9 ms without (profiles/README.md)
    thread_local int *d_prefix[64] = {nullptr};
    thread_local int d_prefix_cap[64] = {0};
    if (dev < 0 || dev >= 64) return 3;
    std::vector<int> prefix;
    int s0 = 0;
    while (s0 < n) {
        // rows [s0, s0 + ns): the leading long rows cut into units of DM_SEG entries, at most `batch` units in all
        prefix.clear();
        prefix.push_back(0);
        int ns = 0, units = 0;
        while (s0 + ns < n && split_on && deg[s0 + ns] > DM_SEG) {
            const int u = (deg[s0 + ns] + DM_SEG - 1) / DM_SEG;
            if (ns > 0 && units + u > batch) break;
            units += u;
            prefix.push_back(units);
            ns++;
        }
        const int n_split = ns, units_split = units;
        if (s0 + ns < n && !(split_on && deg[s0 + ns] > DM_SEG)) {
            const int more = std::min(n - s0 - ns, std::max(0, batch - units));
            ns += more;
            units += more;
        }
        if (ns == 0) return 1;
        double *ws = dmma_workspace((size_t)std::max(units, batch) * per);
        if (!ws) return 1;
        DmSplit sp{nullptr, n_split, units_split};
        if (n_split > 0) {
            if (d_prefix_cap[dev] < n_split + 1) {
                if (d_prefix[dev]) cudaFree(d_prefix[dev]);
                d_prefix[dev] = nullptr;
                d_prefix_cap[dev] = 0;
                if (cudaMalloc((void **)&d_prefix[dev], (size_t)(n_split + 1024) * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return 1; }
                d_prefix_cap[dev] = n_split + 1024;
            }
            // the host vector is reused by the next batch: the copy must have left it before that
            if (cudaMemcpyAsync(d_prefix[dev], prefix.data(), (size_t)(n_split + 1) * sizeof(int), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
                cudaStreamSynchronize(stream) != cudaSuccess)
                return 1;
            sp.prefix = d_prefix[dev];
        }
        long long gb = (long long)sms * occ, gf = (long long)sms * occf;
        if (gb > units) gb = units;
        if (gf > (ns + DM_FW - 1) / DM_FW) gf = (ns + DM_FW - 1) / DM_FW;
        build<<<(unsigned)gb, Cfg::NT, smem, stream>>>(p, kd, s0, units, sp, ws);
        factor<<<(unsigned)gf, DM_FW * 32, 0, stream>>>(p, kd, s0, ns, sp, ws);
        s0 += ns;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

template <int MODEL> int dispatch_dmma(const CgSweepParams &p, cudaStream_t stream)
{
    const int kd = p.kk + ((MODEL != kModelImplicit && p.solve_bias) ? 1 : 0);
    const int need = (kd + 1 + 7) / 8;   // tile rows of the extended matrix
    if (need <= 3) return launch_dmma<3, MODEL, 8>(p, kd, stream);
    if (need <= 5) return launch_dmma<5, MODEL, 6>(p, kd, stream);
    if (need <= 9) return launch_dmma<9, MODEL, 4>(p, kd, stream);
    if (need <= 13) return launch_dmma<13, MODEL, 2>(p, kd, stream);
    if (need <= 17) return launch_dmma<17, MODEL, 2>(p, kd, stream);
    return 3;
}

}  // namespace

int launch_explicit_chol_sweep_dmma(const CgSweepParams &p, cudaStream_t stream)
{
    return (p.gram || p.qvec || p.solve_all_rows) ? dispatch_dmma<kModelCollective>(p, stream)
                                                  : dispatch_dmma<kModelExplicit>(p, stream);
}
int launch_implicit_chol_sweep_dmma(const CgSweepParams &p, cudaStream_t stream)
{
    return dispatch_dmma<kModelImplicit>(p, stream);
}

#else   // the fp32 library builds its normal matrices on tcgen05 (sweep_nm.cu)

int launch_explicit_chol_sweep_dmma(const CgSweepParams &, cudaStream_t) { return 3; }
int launch_implicit_chol_sweep_dmma(const CgSweepParams &, cudaStream_t) { return 3; }

#endif

}  // namespace cmfb200
