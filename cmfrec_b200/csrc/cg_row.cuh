// The conjugate-gradient solve of ONE factor row by a team of TW warps, independent of where the opposing
// rows come from.  The source of the row's stored entries is a policy object ("Gather") with one method
//
//     template <int KIND> void pass(const T (&vec)[C], T vecb, T (&acc)[C], T &accb);
//
// which must add, for every stored entry e of the row,   coef_e * g_e  into acc  and  coef_e  into accb,
// where g_e is the opposing row of entry e (this lane's C coordinates of it), d_e = <g_e, vec> + vecb and
// coef_e = entry_coef<KIND>(d_e, x_e).  Two policies exist: direct gathers from global memory / L2
// (sweep_cg.cu) and rows staged in shared memory by bulk async copies (sweep_cg_staged.cu).
//
// Arithmetic: reference factors_explicit_cg (src/common.c:1098-1188) and factors_implicit_cg
// (src/common.c:1914-1986); see sweep_cg.cu for the full list of reproduced details.
#pragma once
#include "sweep.h"
#include "device_utils.cuh"
#include <cooperative_groups.h>

namespace cmfb200 {

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

// sum_j v[j] * w[j] as (even-index sum) + (odd-index sum); float uses the packed FMA of sm_100 (FFMA2: two
// independent IEEE fmas per instruction, so the result is bit-identical to the scalar form and half the issue slots)
template <int C> __device__ __forceinline__ float dot_pairs(const float (&v)[C], const float (&w)[C])
{
    static_assert(C % 2 == 0, "pairs");
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < C; j += 2) s = __ffma2_rn(make_float2(v[j], v[j + 1]), make_float2(w[j], w[j + 1]), s);
    return s.x + s.y;
}
template <int C> __device__ __forceinline__ double dot_pairs(const double (&v)[C], const double (&w)[C])
{
    double d0 = 0.0, d1 = 0.0;
#pragma unroll
    for (int j = 0; j < C; j += 2) {
        d0 = fma(v[j], w[j], d0);
        if (j + 1 < C) d1 = fma(v[j + 1], w[j + 1], d1);
    }
    return d0 + d1;
}
// acc += coef * v
template <int C> __device__ __forceinline__ void axpy_pairs(float coef, const float (&v)[C], float (&acc)[C])
{
    const float2 cc = make_float2(coef, coef);
#pragma unroll
    for (int j = 0; j < C; j += 2) {
        const float2 r = __ffma2_rn(cc, make_float2(v[j], v[j + 1]), make_float2(acc[j], acc[j + 1]));
        acc[j] = r.x;
        acc[j + 1] = r.y;
    }
}
template <int C> __device__ __forceinline__ void axpy_pairs(double coef, const double (&v)[C], double (&acc)[C])
{
#pragma unroll
    for (int j = 0; j < C; j++) acc[j] = fma(coef, v[j], acc[j]);
}

enum PassKind { kExplicitResidual = 0, kExplicitAp = 1, kImplicitResidual = 2, kImplicitAp = 3 };

// x is the stored value (explicit: already reduced by the opposing bias), d the current prediction
template <int KIND, typename T> __device__ __forceinline__ T entry_coef(T d, T x)
{
    if constexpr (KIND == kExplicitResidual) return x - d;
    else if constexpr (KIND == kExplicitAp) return d;
    else if constexpr (KIND == kImplicitResidual) return -(d - T(1)) * x - d;
    else return d * (x - T(1)) + d;
}

template <int TW> __device__ __forceinline__ void team_barrier(int bar_id)
{
    if constexpr (TW == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(TW * 32) : "memory");
}

// Shared-memory scratch of one team: reduction buffer [2][TW][KP+4] (TW > 1 only) and the broadcast vector
// [KP] used by the implicit model's Gram product.
template <typename T, int C, int L, int TW> struct TeamScratch {
    static constexpr int KP = Layout<T, C, L>::KP;
    static constexpr int RED_STRIDE = KP + 4;
    static constexpr int elems() { return (TW > 1 ? 2 * TW * RED_STRIDE : 0) + KP; }
};

// CL > 1: the team spans the CL thread blocks of a cluster (TW warps in each); per-pass totals are then also
// combined across the blocks through distributed shared memory (`cl_buf`, [2][KP+4] in every block).
// MODEL: 0 = explicit feedback; 1 = implicit feedback (constant matrix = Gram of the opposing factor);
//        2 = explicit feedback with side information / implicit features: a constant matrix Q (p.gram) is added to
//            every row's system and a per-row vector q (p.qvec) to its right-hand side
//            (reference collective_block_cg, src/collective.c:2134-2902: Q = w_user C^T C + w_implicit Bi^T Bi,
//             q = w_user C^T u_i + w_implicit sum_e Bi_e).
constexpr int kModelExplicit = 0, kModelImplicit = 1, kModelCollective = 2;

template <typename T, int C, int L, int MODEL, int TW, bool GRAM_SMEM, int CL = 1, bool COOP = false> struct CgRow {
    static constexpr bool IMPLICIT = MODEL == kModelImplicit;
    static constexpr bool HAS_Q = MODEL != kModelExplicit;
    static_assert(CL == 1 || TW > 1, "a cluster team is made of whole thread blocks");
    typedef Layout<T, C, L> Lay;
    typedef TeamScratch<T, C, L, TW> Scr;
    static constexpr int G = 32 / L;

    const CgSweepParams &p;
    T *red;          // [2][TW][RED_STRIDE]
    T *vec_sm;       // [KP]
    const T *gram;   // shared-memory copy (row stride KP) or global (row stride kk)
    int lane, wt, g, l, bar_id;
    int phase;
    T *cl_buf = nullptr;   // CL > 1 only
    int cl_phase = 0;
    int cl_rank = 0;       // rank of this block in the cluster (CL > 1)
    // COOP (one warp per row, constant matrix too large for shared memory, 256-thread blocks whose 8 warps work on 8 rows in
    // lockstep): the 8 rows' products with the constant matrix are computed TOGETHER, thread c taking column c for all 8
    // vectors, so the matrix is read once per 8 rows and pass instead of once per row and pass (k = 256: 256 KB each time)
    static constexpr int kCoopRows = 8, kCoopBar = 8;
    T *coop_base = nullptr;   // scratch stripes of the block's 8 warps: [8][coop_stride], vector at 0, product at KP
    int coop_stride = 0;

    __device__ __forceinline__ CgRow(const CgSweepParams &p_, T *scratch, const T *gram_, int warp_in_team, int bar_id_)
        : p(p_), red(scratch), vec_sm(scratch + (TW > 1 ? 2 * TW * Scr::RED_STRIDE : 0)), gram(gram_), wt(warp_in_team),
          bar_id(bar_id_), phase(0)
    {
        lane = threadIdx.x & 31;
        g = lane / L;
        l = lane % L;
    }

    __device__ __forceinline__ void sync() const { team_barrier<TW>(bar_id); }

    // partial sums held by every lane -> totals in every lane of the team.
    // Warps first combine their groups with shuffles; the TW per-warp partials are then summed column-parallel
    // (one thread per coordinate) through shared memory, and for a cluster the CL per-block totals likewise
    // through distributed shared memory -- never every thread reading every partial.
    __device__ __forceinline__ void combine(T (&acc)[C], T &accb)
    {
#pragma unroll
        for (int j = 0; j < C; j++) acc[j] = across_groups_sum<L>(acc[j]);
        accb = across_groups_sum<L>(accb);
        if constexpr (TW > 1) {
            T *partial = red;                                      // [TW][RED_STRIDE]
            T *total = red + (TW + phase) * Scr::RED_STRIDE;       // [RED_STRIDE], double-buffered: cluster peers read it
            if (g == 0) {
#pragma unroll
                for (int j = 0; j < C; j++) partial[wt * Scr::RED_STRIDE + Lay::col(l, j)] = acc[j];
                if (l == 0) partial[wt * Scr::RED_STRIDE + Lay::KP] = accb;
            }
            sync();
            const int tid = wt * 32 + lane;
            for (int c = tid; c <= Lay::KP; c += TW * 32) {
                T s = T(0);
#pragma unroll
                for (int ww = 0; ww < TW; ww++) s += partial[ww * Scr::RED_STRIDE + c];
                total[c] = s;
            }
            sync();
            const T *src = total;
            if constexpr (CL > 1) {
                namespace cg = cooperative_groups;
                cg::cluster_group cluster = cg::this_cluster();
                cluster.sync();
                T *ctot = cl_buf + phase * Scr::RED_STRIDE;
                for (int c = tid; c <= Lay::KP; c += TW * 32) {
                    T s = T(0);
                    for (int r = 0; r < CL; r++) s += cluster.map_shared_rank(total, r)[c];
                    ctot[c] = s;
                }
                sync();
                src = ctot;
            }
#pragma unroll
            for (int j = 0; j < C; j++) acc[j] = src[Lay::col(l, j)];
            accb = src[Lay::KP];
            phase ^= 1;
        }
    }

    __device__ __forceinline__ T dot_full(const T (&x)[C], const T (&y)[C], T xb, T yb) const
    {
        T s0 = T(0), s1 = T(0);
#pragma unroll
        for (int j = 0; j < C; j += 2) {
            s0 = fma(x[j], y[j], s0);
            if (j + 1 < C) s1 = fma(x[j + 1], y[j + 1], s1);
        }
        T s = group_sum<L>(s0 + s1);
        return fma(xb, yb, s);
    }

    // acc += sign * gram * vec, rows of gram distributed over the groups of the team
    __device__ __forceinline__ void gram_matvec(const T (&vec)[C], T sign, T (&acc)[C])
    {
        const int kk = p.kk;
        sync();  // previous readers of vec_sm are done
        if (g == 0 && wt == 0) {
#pragma unroll
            for (int j = 0; j < C; j++) vec_sm[Lay::col(l, j)] = vec[j];
        }
        sync();
        const int ngroups = CL * TW * G;
        const int gg = (cl_rank * TW + wt) * G + g;
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

    // acc += sign * gram * vec for the 8 rows the block's warps hold; EVERY thread of the block must call it (warps without a
    // live row pass active = false)
    __device__ __forceinline__ void coop_gram(const T (&vec)[C], bool active, T sign, T (&acc)[C])
    {
        constexpr int KP = Lay::KP;
        constexpr int PARTS = (kCoopRows * 32) / KP;     // threads per column: the rows of the matrix are split between them
        static_assert(!COOP || ((KP == 256 || KP == 128) && TW == 1 && CL == 1 && !GRAM_SMEM), "COOP layout");
        constexpr int VN = VecOf<T>::N;
        typedef typename VecOf<T>::type Vec;
        const int kk = p.kk, tid = threadIdx.x, w = tid >> 5;
        T *mine = coop_base + (size_t)w * coop_stride;
        if (g == 0) {
#pragma unroll
            for (int j = 0; j < C; j++) mine[Lay::col(l, j)] = active ? vec[j] : T(0);
        }
        asm volatile("bar.sync %0, %1;" ::"r"(kCoopBar), "r"(kCoopRows * 32) : "memory");
        const int c = tid % KP, part = tid / KP;
        if (c < kk) {
            T y[kCoopRows];
#pragma unroll
            for (int r = 0; r < kCoopRows; r++) y[r] = T(0);
            const T *gcol = gram + c;
            // this thread's share of the rows of the matrix, in whole vectors
            const int per = ((kk + PARTS * VN - 1) / (PARTS * VN)) * VN;
            int d = part * per;
            const int dend = min(kk, d + per);
            for (; d + VN <= dend; d += VN) {
                T gv[VN];
#pragma unroll
                for (int e = 0; e < VN; e++) gv[e] = __ldg(gcol + (size_t)(d + e) * kk);
#pragma unroll
                for (int r = 0; r < kCoopRows; r++) {
                    const Vec vv = *reinterpret_cast<const Vec *>(coop_base + (size_t)r * coop_stride + d);
                    const T *pv = reinterpret_cast<const T *>(&vv);
#pragma unroll
                    for (int e = 0; e < VN; e++) y[r] = fma(gv[e], pv[e], y[r]);
                }
            }
            for (; d < dend; d++) {
                const T gv = __ldg(gcol + (size_t)d * kk);
#pragma unroll
                for (int r = 0; r < kCoopRows; r++) y[r] = fma(gv, coop_base[(size_t)r * coop_stride + d], y[r]);
            }
#pragma unroll
            for (int r = 0; r < kCoopRows; r++) coop_base[(size_t)r * coop_stride + KP * (1 + part) + c] = y[r];
        }
        asm volatile("bar.sync %0, %1;" ::"r"(kCoopBar), "r"(kCoopRows * 32) : "memory");
        if (active) {
#pragma unroll
            for (int j = 0; j < C; j++) {
                const int cc = Lay::col(l, j);
                if (cc < kk) {
                    T t = mine[KP + cc];
#pragma unroll
                    for (int q = 1; q < PARTS; q++) t += mine[KP * (1 + q) + cc];
                    acc[j] = fma(sign, t, acc[j]);
                }
            }
        }
    }

    // a warp of a COOP block that has no row to solve in this slot still takes part in the block's products
    __device__ __forceinline__ void coop_idle()
    {
        T z[C], acc[C];
#pragma unroll
        for (int j = 0; j < C; j++) z[j] = acc[j] = T(0);
        for (int i = 0; i < 1 + p.max_cg_steps; i++) coop_gram(z, false, T(1), acc);
    }

    // The same solve with the constant-matrix products shared by the block (COOP): every warp makes exactly
    // 1 + max_cg_steps calls of coop_gram, rows that have met an exit threshold go on as passengers without touching
    // their state -- the iterates are those of solve().
    template <typename Gather> __device__ void solve_coop(int row, int nnz, Gather &gather)
    {
        const int kk = p.kk;
        T *frow = p.F + (size_t)row * (size_t)p.ldF;
        T a[C], r[C], pv[C], acc[C];
        T ab = T(0), rb = T(0), pb = T(0), accb = T(0);
#pragma unroll
        for (int j = 0; j < C; j++) {
            const int c = Lay::col(l, j);
            a[j] = (c < kk) ? frow[c] : T(0);
        }
        const bool hb = !IMPLICIT && p.solve_bias;
        if (hb) ab = p.bias_start_one ? T(1) : p.Fbias[row];
        T lam = p.lam, lam_last = p.lam_last;
        if (!IMPLICIT && p.scale_lam && nnz > 0) {
            lam *= (T)nnz;
            if (!p.scale_bias_const) lam_last *= (T)nnz;
        }
#pragma unroll
        for (int j = 0; j < C; j++) acc[j] = T(0);
        accb = T(0);
        coop_gram(a, true, T(-1), acc);
        gather.template pass<IMPLICIT ? kImplicitResidual : kExplicitResidual>(a, ab, acc, accb);
        combine(acc, accb);
#pragma unroll
        for (int j = 0; j < C; j++) {
            const int c = Lay::col(l, j);
            r[j] = (c < kk) ? fma(-lam, a[j], acc[j]) : T(0);
            if (p.qvec && c < kk) r[j] += p.qvec[(size_t)row * (size_t)p.ldq + c];
        }
        if (hb) {
            rb = fma(-lam, ab, accb);
            if (lam != lam_last) rb -= (lam_last - lam) * ab;
        }
        T r_old = dot_full(r, r, rb, rb);
        bool changed = false;
        bool active = !(r_old <= T(1e-12));
#pragma unroll
        for (int j = 0; j < C; j++) pv[j] = r[j];
        pb = rb;
        for (int step = 0; step < p.max_cg_steps; step++) {
#pragma unroll
            for (int j = 0; j < C; j++) acc[j] = T(0);
            accb = T(0);
            coop_gram(pv, active, T(1), acc);
            if (active) {   // warp-uniform
                gather.template pass<IMPLICIT ? kImplicitAp : kExplicitAp>(pv, pb, acc, accb);
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
                if (r_new <= T(1e-8)) {
                    active = false;
                } else {
                    const T beta = r_new / r_old;
#pragma unroll
                    for (int j = 0; j < C; j++) pv[j] = fma(beta, pv[j], r[j]);
                    pb = fma(beta, pb, rb);
                    r_old = r_new;
                }
            }
        }
        if (g == 0 && wt == 0) {
            if (changed) {
#pragma unroll
                for (int j = 0; j < C; j++) {
                    const int c = Lay::col(l, j);
                    if (c < kk) frow[c] = a[j];
                }
            }
            if (hb && l == 0 && (changed || p.bias_start_one)) p.Fbias[row] = ab;
        }
    }

    // Solve one row with at least one stored entry.  `gather` is positioned on that row.
    template <typename Gather> __device__ void solve(int row, int nnz, Gather &gather)
    {
        const int kk = p.kk;
        T *frow = p.F + (size_t)row * (size_t)p.ldF;

        T a[C], r[C], pv[C], acc[C];
        T ab = T(0), rb = T(0), pb = T(0), accb = T(0);
#pragma unroll
        for (int j = 0; j < C; j++) {
            const int c = Lay::col(l, j);
            a[j] = (c < kk) ? frow[c] : T(0);
        }
        const bool hb = !IMPLICIT && p.solve_bias;
        if (hb) ab = p.bias_start_one ? T(1) : p.Fbias[row];

        T lam = p.lam, lam_last = p.lam_last;
        if (!IMPLICIT && p.scale_lam && nnz > 0) {   // rows without entries (models with side information only) keep lam as is
            lam *= (T)nnz;
            if (!p.scale_bias_const) lam_last *= (T)nnz;
        }

        // ---- residual at the starting point
#pragma unroll
        for (int j = 0; j < C; j++) acc[j] = T(0);
        accb = T(0);
        if constexpr (HAS_Q) gram_matvec(a, T(-1), acc);
        gather.template pass<IMPLICIT ? kImplicitResidual : kExplicitResidual>(a, ab, acc, accb);
        combine(acc, accb);
#pragma unroll
        for (int j = 0; j < C; j++) {
            const int c = Lay::col(l, j);
            r[j] = (c < kk) ? fma(-lam, a[j], acc[j]) : T(0);
            if constexpr (MODEL != kModelExplicit) {   // side information: explicit + collective model, or implicit with U / I
                if (p.qvec && c < kk) r[j] += p.qvec[(size_t)row * (size_t)p.ldq + c];
            }
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
                if constexpr (HAS_Q) gram_matvec(pv, T(1), acc);
                gather.template pass<IMPLICIT ? kImplicitAp : kExplicitAp>(pv, pb, acc, accb);
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
        if (g == 0 && wt == 0 && cl_rank == 0) {
            if (changed) {
#pragma unroll
                for (int j = 0; j < C; j++) {
                    const int c = Lay::col(l, j);
                    if (c < kk) frow[c] = a[j];
                }
            }
            if (hb && l == 0 && (changed || p.bias_start_one)) p.Fbias[row] = ab;
        }
    }

    // rows without entries are skipped by the reference and keep whatever their storage holds, which for the
    // bias column is the 1.0 written there before the sweep (src/collective.c:8538-8542)
    __device__ __forceinline__ void empty_row(int row) const
    {
        if constexpr (MODEL == kModelCollective) {
            // reached only when the row has neither entries nor side information: the reference zeroes it
            // (collective_closed_form_block "zero_out", src/collective.c:1259-1270)
            if (g == 0 && wt == 0) {
                T *frow = p.F + (size_t)row * (size_t)p.ldF;
#pragma unroll
                for (int j = 0; j < C; j++) {
                    const int c = Lay::col(l, j);
                    if (c < p.kk) frow[c] = T(0);
                }
                if (l == 0 && p.solve_bias) p.Fbias[row] = T(0);
            }
        } else {
            if (!IMPLICIT && p.solve_bias && p.bias_start_one && lane == 0 && wt == 0)
                p.Fbias[row] = T(1);
        }
    }
};

}  // namespace cmfb200
