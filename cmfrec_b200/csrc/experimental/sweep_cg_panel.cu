// Conjugate-gradient half-sweep, "panel" variant: the gathered opposing-factor rows of a factor row (its PANEL,
// [stored entries x k]) are copied from L2 into shared memory exactly once and all 1 + max_cg_steps passes of the
// CG read them from there -- the single-gather traffic model of SURVEY.md 8(d).
//
// What differs from sweep_cg_resident.cu (same idea, measured slower than the direct kernel, profiles/README.md):
//   * the pass over the panel has no predicates and no streaming branch: a warp's share of the panel is padded to
//     whole steps with zero rows whose stored value and "one" flag are zero, so that every form of the per-entry
//     coefficient vanishes on the padding; the loop is software-pipelined two steps deep (LDS.128 of step s+1 in
//     flight while step s runs its FFMA2 chain);
//   * the per-pass sums are combined across the 32/L groups of a warp by a TRANSPOSED shuffle reduction (each stage
//     halves the number of live registers: 14 shuffles instead of 48 at k=64), which leaves every lane owning C/G
//     columns of the total, and the CG vector algebra (a, r, p, the two dot products) is carried out in that
//     distributed form -- 3 x C/G registers instead of 3 x C -- with only the multiplicand of the next pass
//     broadcast back to all groups through the warp's shared-memory stripe;
//   * the team size (1, 2, 4 or 8 warps of a thread block) is a run-time value: ONE instantiation of the row solver
//     per model instead of four, one named barrier per pass for teams of more than one warp (partials are
//     double-buffered by pass parity and every warp sums the team's partials for its own columns).
// Rows longer than a thread block's panel capacity go to clusters of 2-16 thread blocks (per-block totals exchanged
// through distributed shared memory); rows longer than the largest cluster holds are left to the direct kernel.
//
// Algebra: reference factors_explicit_cg (src/common.c:1098-1188), factors_implicit_cg (src/common.c:1914-1986),
// collective_block_cg (src/collective.c:2134-2902); details reproduced as listed in sweep_cg.cu / cg_row.cuh.
#include "cg_row.cuh"
#include <algorithm>
#include <cstdint>
#include <cstdlib>

namespace cmfb200 {

namespace {

constexpr int kPW = 8;   // warps per thread block

__device__ __forceinline__ void panel_cp_async_16(void *smem_dst, const void *gsrc)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void panel_cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void named_barrier(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(const void *local_smem, int rank)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(local_smem);
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
template <typename T> __device__ __forceinline__ T ld_cluster(uint32_t addr);
template <> __device__ __forceinline__ float ld_cluster<float>(uint32_t addr)
{
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
template <> __device__ __forceinline__ double ld_cluster<double>(uint32_t addr)
{
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Shapes
// ---------------------------------------------------------------------------------------------------------
template <typename T, int C, int L> struct Panel {
    typedef Layout<T, C, L> Lay;
    typedef typename VecOf<T>::type Vec;
    static constexpr int VN = VecOf<T>::N;
    static_assert(Lay::VN == VN, "the lane layout must be made of 16-byte pieces");
    static constexpr int G = 32 / L;              // entries per step of a warp
    static constexpr int KP = Lay::KP;
    static_assert(C % G == 0, "the transposed reduction needs C to be a multiple of the number of groups");
    static constexpr int OWN = C / G;             // columns of a reduced vector owned by one lane
    // 4-lane groups read 64 bytes each and two of them share a quarter-warp: with a row stride of 16 (mod 128) bytes
    // and the two groups taking entries G/2 apart their reads fall on disjoint banks.  Wider groups read whole
    // 128-byte bank rows and never conflict.
    static constexpr int PAD = (L == 4) ? VN : 0;
    static constexpr int RS = KP + PAD;           // elements between panel rows
    static constexpr int U = KP / VN;             // 16-byte pieces per panel row
    static constexpr int ENTRY = RS + 2;          // panel row + {stored value, one}
    // per-warp stripe: partial / total vectors of the pass, double-buffered by pass parity, + the broadcast vector
    static constexpr int PART = KP + 4;           // [KP] columns, [KP] bias coordinate, padding
    static constexpr int STRIPE = 2 * PART + KP;
};

// entry taken by group g in a step: groups 2i and 2i+1 (same quarter-warp when L == 4) take entries G/2 apart
template <int L> __device__ __forceinline__ int group_entry(int g)
{
    constexpr int G = 32 / L;
    return (L == 4) ? ((g >> 1) + (G / 2) * (g & 1)) : g;
}

// ---------------------------------------------------------------------------------------------------------
// One row solved by a team of `tw` warps (all in one thread block) or, CL > 1, by all kPW * CL warps of a cluster.
// ---------------------------------------------------------------------------------------------------------
template <typename T, int C, int L, int MODEL, bool GRAM_SMEM, int CL> struct PanelRow {
    typedef Panel<T, C, L> P;
    typedef Layout<T, C, L> Lay;
    typedef typename P::Vec Vec;
    static constexpr bool IMPLICIT = MODEL == kModelImplicit;
    static constexpr bool HAS_Q = MODEL != kModelExplicit;
    static constexpr int G = P::G, OWN = P::OWN, VN = P::VN, KP = P::KP, RS = P::RS;

    const CgSweepParams &p;
    T *rows;            // this warp's panel region [cap][RS]
    T *xs;              // [cap][2]: stored value (explicit: already reduced by the opposing bias), 1.0
    T *stripe0;         // stripe of the first warp of the team
    T *stripe;          // this warp's stripe
    const T *gram;
    T *cl_tot;          // CL > 1: [2][PART] per-block totals read by the cluster peers
    int cap, lane, g, l, gi;
    int wt, tw;         // index in the team (block-local) / warps of the team in this block
    int bar_id;
    int cl_rank;
    int j0;             // first owned register index
    int parity;

    __device__ __forceinline__ PanelRow(const CgSweepParams &p_, T *region, int cap_, T *stripe0_, const T *gram_, int wt_, int tw_,
                                        int bar_id_)
        : p(p_), rows(region), xs(region + (size_t)cap_ * RS), stripe0(stripe0_), stripe(stripe0_ + (size_t)wt_ * P::STRIPE),
          gram(gram_), cl_tot(nullptr), cap(cap_), wt(wt_), tw(tw_), bar_id(bar_id_), cl_rank(0), parity(0)
    {
        lane = threadIdx.x & 31;
        g = lane / L;
        l = lane % L;
        gi = group_entry<L>(g);
        // owned registers after the transposed reduction: the stage with lane mask m keeps the upper half when (lane & m)
        int j = 0, n = C;
#pragma unroll
        for (int m = 16; m >= L; m >>= 1) {
            n >>= 1;
            if (lane & m) j += n;
        }
        j0 = j;
    }

    __device__ __forceinline__ void team_sync() const
    {
        if (tw == 1) __syncwarp();
        else named_barrier(bar_id, tw * 32);
    }

    // ---- staging: copy the opposing rows of this warp's `mine` entries (starting at `beg`) into its region
    __device__ __forceinline__ int stage(size_t beg, int mine)
    {
        __syncwarp();
        const int ldG = p.ldG;
        const int nsteps = (mine + G - 1) / G;
        for (int i = 0; i * 32 < mine; i++) {
            const int e = i * 32 + lane;
            int col = -1;
            if (e < mine) {
                col = p.X.idx[beg + e];
                T x = p.X.val[beg + e];
                if (p.center_opp) x -= __ldg(p.Gbias + col);
                xs[2 * e] = x;
                xs[2 * e + 1] = T(1);
            }
            if constexpr (P::U <= 32) {
                // 32 / U entries per round: lane -> (entry lane / U of the round, piece lane % U)
                constexpr int EPR = 32 / P::U;
                const int part = lane % P::U;
                const bool in_row = part * VN < ldG;
                T *dst = rows + (size_t)(i * 32 + lane / P::U) * RS + part * VN;
                const T *src = p.G + part * VN;
#pragma unroll 4
                for (int j = 0; j < P::U; j++) {
                    const int c = __shfl_sync(CMF_FULL_MASK, col, j * EPR + lane / P::U);
                    if (c >= 0 && in_row) panel_cp_async_16(dst + (size_t)(j * EPR) * RS, src + (size_t)c * (size_t)ldG);
                }
            } else {
#pragma unroll 4
                for (int j = 0; j < P::U; j++) {
                    const int u = j * 32 + lane;
                    const int ent = u / P::U, part = u % P::U;
                    const int c = __shfl_sync(CMF_FULL_MASK, col, ent);
                    if (c >= 0 && part * VN < ldG)
                        panel_cp_async_16(rows + (size_t)(i * 32 + ent) * RS + part * VN, p.G + (size_t)c * (size_t)ldG + part * VN);
                }
            }
        }
        // padding up to whole steps: zero rows, zero value, zero "one"
        {
            const int tail_end = nsteps * G;
            for (int u = mine * RS + lane; u < tail_end * RS; u += 32) rows[u] = T(0);
            for (int u = 2 * mine + lane; u < 2 * tail_end; u += 32) xs[u] = T(0);
        }
        panel_cp_async_wait_all();
        __syncwarp();
        return nsteps;
    }

    // once per kernel: columns the staging never writes (>= ldG) must read as zero
    __device__ __forceinline__ void clear_region()
    {
        for (int u = lane; u < cap * P::ENTRY; u += 32) rows[u] = T(0);
        __syncwarp();
    }

    __device__ __forceinline__ void load_step(int s, T (&v)[C], T &x, T &one) const
    {
        const int slot = s * G + gi;
        const T *srow = rows + (size_t)slot * RS;
#pragma unroll
        for (int q = 0; q < C / VN; q++) {
            const Vec vv = *reinterpret_cast<const Vec *>(srow + (q * L + l) * VN);
            const T *pv = reinterpret_cast<const T *>(&vv);
#pragma unroll
            for (int e2 = 0; e2 < VN; e2++) v[q * VN + e2] = pv[e2];
        }
        x = xs[2 * slot];
        one = xs[2 * slot + 1];
    }

    template <int KIND>
    __device__ __forceinline__ void step(const T (&v)[C], T x, T one, const T (&vec)[C], T vecb, T (&acc)[C], T &accb) const
    {
        T d = group_sum<L>(dot_pairs<C>(v, vec));
        d = fma(vecb, one, d);   // opposing value of the bias coordinate is 1 (0 on the padding)
        const T coef = entry_coef<KIND>(d, x);
        axpy_pairs<C>(coef, v, acc);
        accb = fma(coef, one, accb);
    }

    // acc += sum_e coef_e g_e, accb += sum_e coef_e over this warp's share of the panel
    template <int KIND>
    __device__ __forceinline__ void pass(int nsteps, const T (&vec)[C], T vecb, T (&acc)[C], T &accb) const
    {
        if (nsteps <= 0) return;
        T v0[C], v1[C], x0, x1, o0, o1;
        load_step(0, v0, x0, o0);
        int s = 0;
        for (; s + 2 < nsteps; s += 2) {
            load_step(s + 1, v1, x1, o1);
            step<KIND>(v0, x0, o0, vec, vecb, acc, accb);
            load_step(s + 2, v0, x0, o0);
            step<KIND>(v1, x1, o1, vec, vecb, acc, accb);
        }
        if (s + 1 < nsteps) {
            load_step(s + 1, v1, x1, o1);
            step<KIND>(v0, x0, o0, vec, vecb, acc, accb);
            step<KIND>(v1, x1, o1, vec, vecb, acc, accb);
        } else {
            step<KIND>(v0, x0, o0, vec, vecb, acc, accb);
        }
    }

    // acc += sign * gram * vec; the rows of gram are dealt over all groups of the team; vec is read from the warp's
    // broadcast stripe (it holds the multiplicand of the coming pass)
    __device__ __forceinline__ void gram_matvec(T sign, T (&acc)[C]) const
    {
        const int kk = p.kk;
        const T *vec_sm = stripe + 2 * P::PART;
        const int ngroups = CL * tw * G;
        const int gg = (cl_rank * tw + wt) * G + g;
        for (int d = gg; d < kk; d += ngroups) {
            const T s = sign * vec_sm[d];
            if constexpr (GRAM_SMEM) {
                const T *mrow = gram + (size_t)d * KP;
#pragma unroll
                for (int q = 0; q < C / VN; q++) {
                    const Vec vv = *reinterpret_cast<const Vec *>(mrow + (q * L + l) * VN);
                    const T *pv = reinterpret_cast<const T *>(&vv);
#pragma unroll
                    for (int e2 = 0; e2 < VN; e2++) acc[q * VN + e2] = fma(pv[e2], s, acc[q * VN + e2]);
                }
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

    // Per-lane partial sums (replicated layout: every group holds C registers) -> the team's totals of this lane's
    // OWN columns, and the total of the bias coordinate in every lane.
    __device__ __forceinline__ void reduce(const T (&acc)[C], T accb, T (&tot)[OWN], T &totb)
    {
        // transposed reduction across the G groups of the warp
        T cur[C];
#pragma unroll
        for (int j = 0; j < C; j++) cur[j] = acc[j];
        int n = C;
#pragma unroll
        for (int m = 16; m >= L; m >>= 1) {
            n >>= 1;
            const bool up = (lane & m) != 0;
#pragma unroll
            for (int j = 0; j < C / 2; j++) {
                if (j < n) {
                    const T keep = up ? cur[j + n] : cur[j];
                    const T send = up ? cur[j] : cur[j + n];
                    cur[j] = keep + __shfl_xor_sync(CMF_FULL_MASK, send, m);
                }
            }
        }
#pragma unroll
        for (int m = 16; m >= L; m >>= 1) accb += __shfl_xor_sync(CMF_FULL_MASK, accb, m);
        if (tw == 1 && CL == 1) {
#pragma unroll
            for (int jj = 0; jj < OWN; jj++) tot[jj] = cur[jj];
            totb = accb;
            return;
        }
        // team: every warp publishes its partial of every column, then sums the team's partials of its own columns
        T *mine = stripe + parity * P::PART;
#pragma unroll
        for (int jj = 0; jj < OWN; jj++) mine[own_col(jj)] = cur[jj];
        if (lane == 0) mine[KP] = accb;
        team_sync();
#pragma unroll
        for (int jj = 0; jj < OWN; jj++) tot[jj] = T(0);
        totb = T(0);
        for (int ww = 0; ww < tw; ww++) {
            const T *theirs = stripe0 + (size_t)ww * P::STRIPE + parity * P::PART;
#pragma unroll
            for (int jj = 0; jj < OWN; jj++) tot[jj] += theirs[own_col(jj)];
            totb += theirs[KP];
        }
        if constexpr (CL > 1) {
            // block totals -> cl_tot (written by warp 0), exchanged through distributed shared memory
            T *bt = cl_tot + parity * P::PART;
            if (wt == 0) {
#pragma unroll
                for (int jj = 0; jj < OWN; jj++) bt[own_col(jj)] = tot[jj];
                if (lane == 0) bt[KP] = totb;
            }
            cluster_barrier();
#pragma unroll
            for (int jj = 0; jj < OWN; jj++) tot[jj] = T(0);
            totb = T(0);
            for (int r = 0; r < CL; r++) {
                const uint32_t base = map_to_cta(bt, r);
#pragma unroll
                for (int jj = 0; jj < OWN; jj++) tot[jj] += ld_cluster<T>(base + (uint32_t)(own_col(jj) * sizeof(T)));
                totb += ld_cluster<T>(base + (uint32_t)(KP * sizeof(T)));
            }
        }
        parity ^= 1;
    }

    __device__ __forceinline__ int own_col(int jj) const { return Lay::col(l, j0 + jj); }

    // sum over all columns of x*y given each lane's owned columns, plus the bias coordinate's product
    __device__ __forceinline__ T dot_owned(const T (&x)[OWN], const T (&y)[OWN], T xb, T yb) const
    {
        T s = T(0);
#pragma unroll
        for (int jj = 0; jj < OWN; jj++) s = fma(x[jj], y[jj], s);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(CMF_FULL_MASK, s, off);
        return fma(xb, yb, s);
    }

    // owned columns -> all C registers of every group, through the warp's broadcast stripe (also read by gram_matvec)
    __device__ __forceinline__ void broadcast(const T (&own)[OWN], T (&full)[C]) const
    {
        T *vec_sm = stripe + 2 * P::PART;
        __syncwarp();   // earlier readers of vec_sm are done
#pragma unroll
        for (int jj = 0; jj < OWN; jj++) vec_sm[own_col(jj)] = own[jj];
        __syncwarp();
#pragma unroll
        for (int q = 0; q < C / VN; q++) {
            const Vec vv = *reinterpret_cast<const Vec *>(vec_sm + (q * L + l) * VN);
            const T *pv = reinterpret_cast<const T *>(&vv);
#pragma unroll
            for (int e2 = 0; e2 < VN; e2++) full[q * VN + e2] = pv[e2];
        }
    }

    // Solve one row; this warp's share of the panel (nsteps steps) is already staged.  nnz = entries of the whole row.
    __device__ void solve(int row, int nnz, int nsteps)
    {
        const int kk = p.kk;
        T *frow = p.F + (size_t)row * (size_t)p.ldF;
        T a[OWN], r[OWN], pv[OWN], tot[OWN];
        T vec[C], acc[C];
        T ab = T(0), rb = T(0), pb = T(0), accb = T(0), totb = T(0);
#pragma unroll
        for (int jj = 0; jj < OWN; jj++) {
            const int c = own_col(jj);
            a[jj] = (c < kk) ? frow[c] : T(0);
        }
        const bool hb = !IMPLICIT && p.solve_bias;
        if (hb) ab = p.bias_start_one ? T(1) : p.Fbias[row];

        T lam = p.lam, lam_last = p.lam_last;
        if (!IMPLICIT && p.scale_lam && nnz > 0) {   // rows without entries (collective model only) keep lam as is
            lam *= (T)nnz;
            if (!p.scale_bias_const) lam_last *= (T)nnz;
        }

        // ---- residual at the starting point
        broadcast(a, vec);
#pragma unroll
        for (int j = 0; j < C; j++) acc[j] = T(0);
        accb = T(0);
        if constexpr (HAS_Q) gram_matvec(T(-1), acc);
        pass<IMPLICIT ? kImplicitResidual : kExplicitResidual>(nsteps, vec, ab, acc, accb);
        reduce(acc, accb, tot, totb);
#pragma unroll
        for (int jj = 0; jj < OWN; jj++) {
            const int c = own_col(jj);
            r[jj] = (c < kk) ? fma(-lam, a[jj], tot[jj]) : T(0);
            if constexpr (MODEL == kModelCollective) {
                if (p.qvec && c < kk) r[jj] += p.qvec[(size_t)row * (size_t)p.ldq + c];
            }
        }
        if (hb) {
            rb = fma(-lam, ab, totb);
            if (lam != lam_last) rb -= (lam_last - lam) * ab;
        }
        T r_old = dot_owned(r, r, rb, rb);
        bool changed = false;
        if (!(r_old <= T(1e-12))) {
#pragma unroll
            for (int jj = 0; jj < OWN; jj++) pv[jj] = r[jj];
            pb = rb;
            for (int it = 0; it < p.max_cg_steps; it++) {
                broadcast(pv, vec);
#pragma unroll
                for (int j = 0; j < C; j++) acc[j] = T(0);
                accb = T(0);
                if constexpr (HAS_Q) gram_matvec(T(1), acc);
                pass<IMPLICIT ? kImplicitAp : kExplicitAp>(nsteps, vec, pb, acc, accb);
                reduce(acc, accb, tot, totb);
#pragma unroll
                for (int jj = 0; jj < OWN; jj++) {
                    const int c = own_col(jj);
                    tot[jj] = (c < kk) ? fma(lam, pv[jj], tot[jj]) : T(0);
                }
                if (hb) {
                    totb = fma(lam, pb, totb);
                    if (lam != lam_last) totb += (lam_last - lam) * pb;
                } else {
                    totb = T(0);
                }
                const T alpha = r_old / dot_owned(pv, tot, pb, totb);
#pragma unroll
                for (int jj = 0; jj < OWN; jj++) {
                    a[jj] = fma(alpha, pv[jj], a[jj]);
                    r[jj] = fma(-alpha, tot[jj], r[jj]);
                }
                ab = fma(alpha, pb, ab);
                rb = fma(-alpha, totb, rb);
                changed = true;
                const T r_new = dot_owned(r, r, rb, rb);
                if (r_new <= T(1e-8)) break;
                const T beta = r_new / r_old;
#pragma unroll
                for (int jj = 0; jj < OWN; jj++) pv[jj] = fma(beta, pv[jj], r[jj]);
                pb = fma(beta, pb, rb);
                r_old = r_new;
            }
        }
        // A row that exits before the first step is left exactly as it was, except that a bias coordinate restarted
        // from 1.0 is what the reference leaves in the matrix.
        if (wt == 0 && cl_rank == 0) {
            if (changed) {
#pragma unroll
                for (int jj = 0; jj < OWN; jj++) {
                    const int c = own_col(jj);
                    if (c < kk) frow[c] = a[jj];
                }
            }
            if (hb && lane == 0 && (changed || p.bias_start_one)) p.Fbias[row] = ab;
        }
    }

    // rows without entries: see CgRow::empty_row (cg_row.cuh)
    __device__ __forceinline__ void empty_row(int row) const
    {
        if constexpr (MODEL == kModelCollective) {
            if (wt == 0 && cl_rank == 0) {
                T *frow = p.F + (size_t)row * (size_t)p.ldF;
#pragma unroll
                for (int jj = 0; jj < OWN; jj++) {
                    const int c = own_col(jj);
                    if (c < p.kk) frow[c] = T(0);
                }
                if (lane == 0 && p.solve_bias) p.Fbias[row] = T(0);
            }
        } else {
            if (!IMPLICIT && p.solve_bias && p.bias_start_one && lane == 0 && wt == 0 && cl_rank == 0) p.Fbias[row] = T(1);
        }
    }
};

// this warp's share of a row dealt over `nw` warps in contiguous equal shares (a multiple of G entries each)
template <int G> __device__ __forceinline__ void warp_share(int row_nnz, int gw, int nw, int &first, int &mine)
{
    const int per = ((row_nnz + nw * G - 1) / (nw * G)) * G;
    first = gw * per;
    mine = row_nnz - first;
    if (mine > per) mine = per;
    if (mine < 0) {
        mine = 0;
        first = 0;
    }
}

// ---------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------
struct PanelPlan {
    // positions in the degree-sorted row list: [b8,b4) one row per thread block, [b4,b2) 4-warp teams,
    // [b2,b1) 2-warp teams, [b1,bend) one warp per row
    int b8, b4, b2, b1, bend;
    int s8, s4, s2, n_slots;   // cumulative slot counts (a slot = one thread block's worth of rows)
    int cap;                   // panel entries per warp
};

template <typename T, int C, int L, int MODEL, bool GRAM_SMEM>
__global__ void __launch_bounds__(kPW * 32, 2) cg_panel_kernel(const CgSweepParams p, const PanelPlan pp)
{
    typedef Panel<T, C, L> P;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *stripes = reinterpret_cast<T *>(smem_raw);
    T *gram_sm = stripes + kPW * P::STRIPE;
    T *regions = gram_sm;
    const T *gram = p.gram;
    if constexpr (MODEL != kModelExplicit && GRAM_SMEM) {
        const int kk = p.kk;
        for (int i = threadIdx.x; i < kk * P::KP; i += blockDim.x) {
            const int d = i / P::KP, c = i % P::KP;
            gram_sm[i] = (c < kk) ? p.gram[(size_t)d * kk + c] : T(0);
        }
        gram = gram_sm;
        regions = gram_sm + (size_t)kk * P::KP;
        __syncthreads();
    }
    const int w = threadIdx.x >> 5;
    T *region = regions + (size_t)w * pp.cap * P::ENTRY;
    PanelRow<T, C, L, MODEL, GRAM_SMEM, 1>(p, region, pp.cap, stripes, gram, 0, 1, 0).clear_region();

    // team size and order-list position of the row this warp works on in a slot (-1: none)
    auto decode = [&](int slot, int &tw) -> int {
        tw = 1;
        if (slot >= pp.n_slots) return -1;
        if (slot < pp.s8) {
            tw = 8;
            return pp.b8 + slot;
        }
        if (slot < pp.s4) {
            tw = 4;
            const int ri = pp.b4 + (slot - pp.s8) * 2 + (w >> 2);
            return ri < pp.b2 ? ri : -1;
        }
        if (slot < pp.s2) {
            tw = 2;
            const int ri = pp.b2 + (slot - pp.s4) * 4 + (w >> 1);
            return ri < pp.b1 ? ri : -1;
        }
        const int ri = pp.b1 + (slot - pp.s2) * 8 + w;
        return ri < pp.bend ? ri : -1;
    };
    const int step = gridDim.x;
    int tw0, tw1;
    int ri = decode(blockIdx.x, tw0);
    int row0 = ri >= 0 ? p.plan.order[ri] : -1;
    ri = decode(blockIdx.x + step, tw1);
    int row1 = ri >= 0 ? p.plan.order[ri] : -1;
    for (int slot = blockIdx.x; slot < pp.n_slots; slot += step) {
        size_t beg = 0, end = 0;
        if (row0 >= 0) {
            beg = p.X.ptr[row0];
            end = p.X.ptr[row0 + 1];
        }
        int tw2;
        ri = decode(slot + 2 * step, tw2);
        const int row2 = ri >= 0 ? p.plan.order[ri] : -1;   // fetched two slots ahead: not waited for
        if (row0 >= 0) {
            const int tw = tw0;
            const int team = w / tw, wt = w % tw;
            // one named barrier per (team size, team): a block's teams run ahead of each other by whole slots
            const int bar_id = (tw == 8) ? 1 : (tw == 4) ? 2 + team : (tw == 2) ? 4 + team : 0;
            PanelRow<T, C, L, MODEL, GRAM_SMEM, 1> s(p, region, pp.cap, stripes + (size_t)(team * tw) * P::STRIPE, gram, wt, tw, bar_id);
            const int nnz = (int)(end - beg);
            if (nnz > 0 || (MODEL == kModelCollective && p.solve_all_rows)) {
                int first, mine;
                warp_share<P::G>(nnz, wt, tw, first, mine);
                const int nsteps = s.stage(beg + first, mine);
                s.solve(row0, nnz, nsteps);
            } else {
                s.empty_row(row0);
            }
            s.team_sync();   // nobody of the team reuses stripes before everybody is done with the row
        }
        row0 = row1;
        tw0 = tw1;
        row1 = row2;
        tw1 = tw2;
    }
}

// rows [first, first + count) of the degree-sorted list, one row per cluster of CL thread blocks
template <typename T, int C, int L, int MODEL, int CL>
__global__ void __launch_bounds__(kPW * 32, 2) cg_panel_cluster_kernel(const CgSweepParams p, int first, int count, int cap)
{
    typedef Panel<T, C, L> P;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *stripes = reinterpret_cast<T *>(smem_raw);
    T *cl_tot = stripes + kPW * P::STRIPE;          // [2][PART]
    T *regions = cl_tot + 2 * P::PART;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int w = threadIdx.x >> 5;
    T *region = regions + (size_t)w * cap * P::ENTRY;
    PanelRow<T, C, L, MODEL, false, CL> s(p, region, cap, stripes, p.gram, w, kPW, 1);
    s.cl_tot = cl_tot;
    s.cl_rank = (int)rank;
    s.clear_region();
    const int n_clusters = gridDim.x / CL;
    for (int slot = blockIdx.x / CL; slot < count; slot += n_clusters) {
        const int row = p.plan.order[first + slot];
        const size_t beg = p.X.ptr[row];
        const int nnz = (int)(p.X.ptr[row + 1] - beg);
        int off, mine;
        warp_share<P::G>(nnz, (int)rank * kPW + w, CL * kPW, off, mine);
        const int nsteps = s.stage(beg + off, mine);
        s.solve(row, nnz, nsteps);
        // peers may still be reading this block's totals of the last pass
        cluster_barrier();
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
int penv(const char *name, int dflt)
{
    const char *e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

struct PanelStreams {
    cudaStream_t s[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ok = false;
    PanelStreams()
    {
        ok = true;
        for (int i = 0; i < 5; i++) {
            ok = ok && cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming) == cudaSuccess;
        }
        ok = ok && cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) == cudaSuccess;
    }
};

PanelStreams &panel_streams()
{
    static PanelStreams ss;   // one process drives one GPU
    return ss;
}

// number of rows of the (descending) degree list with more than `x` stored entries
int rows_longer_than(const int_t *deg, int n, long long x)
{
    return (int)(std::lower_bound(deg, deg + n, x, [](int_t d, long long v) { return (long long)d > v; }) - deg);
}

// how many clusters of CL thread blocks can be resident at once (0: this cluster size cannot be launched)
template <typename T, int C, int L, int MODEL, int CL> int probe_panel_cluster(size_t smem)
{
    auto kern = cg_panel_cluster_kernel<T, C, L, MODEL, CL>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (CL > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL, 1, 1);
    cfg.blockDim = dim3(kPW * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return 0;
    }
    return max_clusters;
}

template <typename T, int C, int L, int MODEL, int CL>
int launch_panel_cluster(const CgSweepParams &p, int first, int count, int cap, size_t smem, int max_clusters, cudaStream_t stream)
{
    auto kern = cg_panel_cluster_kernel<T, C, L, MODEL, CL>;
    cudaLaunchConfig_t cfg = {};
    const int clusters = count < max_clusters ? count : max_clusters;
    cfg.gridDim = dim3(clusters * CL, 1, 1);
    cfg.blockDim = dim3(kPW * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, p, first, count, cap) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    return 0;
}

template <typename T, int C, int L, int MODEL>
int launch_panel_cfg(const CgSweepParams &p, cudaStream_t stream, int *n_launches, int *n_covered_from)
{
    typedef Panel<T, C, L> P;
    const int n_rows = p.plan.n_rows;
    const int_t *deg = p.plan.host_deg;
    if (!deg) return 3;
    if (n_rows <= 0) return 0;

    int dev = 0, sms = 148, smem_sm = 233472, smem_optin = 232448;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    // two thread blocks per SM; the driver reserves 1 KB per block
    size_t per_block = (size_t)smem_sm / 2 - 1024;
    if (per_block > (size_t)smem_optin) per_block = (size_t)smem_optin;

    const size_t entry_bytes = (size_t)P::ENTRY * sizeof(T);
    const size_t fixed1 = (size_t)kPW * P::STRIPE * sizeof(T);
    size_t gram_bytes = MODEL != kModelExplicit ? (size_t)p.kk * P::KP * sizeof(T) : 0;
    const bool gram_smem = MODEL != kModelExplicit && gram_bytes <= per_block / 3;
    if (!gram_smem) gram_bytes = 0;
    if (fixed1 + gram_bytes >= per_block) return 3;
    const int cap1 = (int)((per_block - fixed1 - gram_bytes) / (kPW * entry_bytes)) / P::G * P::G;
    const size_t fixedC = fixed1 + (size_t)2 * P::PART * sizeof(T);
    const int capC = (int)((per_block - fixedC) / (kPW * entry_bytes)) / P::G * P::G;
    if (cap1 < 2 * P::G || capC < 2 * P::G) return 3;   // rows too wide for a shared-memory panel: use the direct kernel

    // buckets of the degree-sorted row list, longest rows first
    const long long blk = (long long)kPW * cap1, cblk = (long long)kPW * capC;
    const size_t smemC = fixedC + (size_t)kPW * capC * entry_bytes;
    // largest usable cluster (16 is a non-portable size: probed), per cluster size the number that can be resident
    static int maxc[4] = {-1, -1, -1, -1};   // clusters of 16, 8, 4, 2 (one process drives one GPU, one shape at a time)
    static size_t maxc_smem = 0;
    if (maxc[0] < 0 || maxc_smem != smemC) {
        maxc[0] = probe_panel_cluster<T, C, L, MODEL, 16>(smemC);
        maxc[1] = probe_panel_cluster<T, C, L, MODEL, 8>(smemC);
        maxc[2] = probe_panel_cluster<T, C, L, MODEL, 4>(smemC);
        maxc[3] = probe_panel_cluster<T, C, L, MODEL, 2>(smemC);
        maxc_smem = smemC;
    }
    int max_cl = penv("CMFB200_PANEL_CLUSTERS", 1) != 0 ? penv("CMFB200_PANEL_MAXCL", 16) : 1;
    if (max_cl >= 16 && maxc[0] < 1) max_cl = 8;
    if (max_cl >= 8 && maxc[1] < 1) max_cl = 4;
    if (max_cl >= 4 && maxc[2] < 1) max_cl = 2;
    if (max_cl >= 2 && maxc[3] < 1) max_cl = 1;
    // bounds[i] = rows longer than what a cluster of (16, 8, 4, 2) blocks / one block holds; rows [0, bounds[0]) are
    // longer than the largest cluster holds: the direct kernel streams them
    int bounds[5];
    bounds[0] = rows_longer_than(deg, n_rows, max_cl >= 16 ? 16 * cblk : max_cl >= 8 ? 8 * cblk : max_cl >= 4 ? 4 * cblk : max_cl >= 2 ? 2 * cblk : blk);
    bounds[1] = max_cl >= 16 ? rows_longer_than(deg, n_rows, 8 * cblk) : bounds[0];
    bounds[2] = max_cl >= 8 ? rows_longer_than(deg, n_rows, 4 * cblk) : bounds[1];
    bounds[3] = max_cl >= 4 ? rows_longer_than(deg, n_rows, 2 * cblk) : bounds[2];
    bounds[4] = max_cl >= 2 ? rows_longer_than(deg, n_rows, blk) : bounds[3];
    for (int i = 1; i < 5; i++)
        if (bounds[i] < bounds[i - 1]) bounds[i] = bounds[i - 1];
    const int n_direct = bounds[0];

    PanelPlan pp;
    pp.cap = cap1;
    pp.b8 = bounds[4];
    pp.b4 = std::max(pp.b8, rows_longer_than(deg, n_rows, (long long)4 * cap1));
    pp.b2 = std::max(pp.b4, rows_longer_than(deg, n_rows, (long long)2 * cap1));
    pp.b1 = std::max(pp.b2, rows_longer_than(deg, n_rows, (long long)cap1));
    pp.bend = n_rows;
    pp.s8 = pp.b4 - pp.b8;
    pp.s4 = pp.s8 + (pp.b2 - pp.b4 + 1) / 2;
    pp.s2 = pp.s4 + (pp.b1 - pp.b2 + 3) / 4;
    pp.n_slots = pp.s2 + (pp.bend - pp.b1 + 7) / 8;

    PanelStreams &ss = panel_streams();
    if (!ss.ok) return 1;
    bool forked = false;
    auto fork_to = [&](int i) {
        if (!forked) {
            cudaEventRecord(ss.fork, stream);
            forked = true;
        }
        cudaStreamWaitEvent(ss.s[i], ss.fork, 0);
    };
    bool used[5] = {false, false, false, false, false};
    // rows too long for any cluster: direct kernel on the leading sub-range of the order list
    if (n_direct > 0) {
        CgSweepParams pd = p;
        pd.plan.n_rows = n_direct;
        pd.plan.n_long = std::min<int_t>(p.plan.n_long, n_direct);
        pd.plan.n_huge = std::min<int_t>(p.plan.n_huge, n_direct);
        pd.side_stream = nullptr;
        fork_to(4);
        // the cached kernel (first entries of every share resident, the rest streamed) where it covers the shape,
        // else the direct kernel
        int nl = 0;
        int rc = MODEL == kModelImplicit ? launch_implicit_cg_sweep_resident(pd, ss.s[4], &nl) : launch_explicit_cg_sweep_resident(pd, ss.s[4], &nl);
        if (rc == 3) {
            pd.plan.host_deg = nullptr;
            rc = MODEL == kModelImplicit ? launch_implicit_cg_sweep(pd, ss.s[4]) : launch_explicit_cg_sweep(pd, ss.s[4]);
            nl = pd.plan.n_huge > 0 ? 2 : 1;
        }
        if (rc) return rc;
        cudaEventRecord(ss.join[4], ss.s[4]);
        used[4] = true;
        if (n_launches) (*n_launches) += nl;
    }
    for (int i = 0; i < 4; i++) {
        const int cfirst = bounds[i], ccount = bounds[i + 1] - bounds[i];
        if (ccount <= 0) continue;
        fork_to(i);
        int rc;
        if (i == 0) rc = launch_panel_cluster<T, C, L, MODEL, 16>(p, cfirst, ccount, capC, smemC, maxc[0], ss.s[i]);
        else if (i == 1) rc = launch_panel_cluster<T, C, L, MODEL, 8>(p, cfirst, ccount, capC, smemC, maxc[1], ss.s[i]);
        else if (i == 2) rc = launch_panel_cluster<T, C, L, MODEL, 4>(p, cfirst, ccount, capC, smemC, maxc[2], ss.s[i]);
        else rc = launch_panel_cluster<T, C, L, MODEL, 2>(p, cfirst, ccount, capC, smemC, maxc[3], ss.s[i]);
        if (rc) return rc;
        cudaEventRecord(ss.join[i], ss.s[i]);
        used[i] = true;
        if (n_launches) (*n_launches)++;
    }
    if (pp.n_slots > 0) {
        const size_t smem1 = fixed1 + gram_bytes + (size_t)kPW * cap1 * entry_bytes;
        auto kern = gram_smem ? cg_panel_kernel<T, C, L, MODEL, true> : cg_panel_kernel<T, C, L, MODEL, false>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1) != cudaSuccess) return 1;
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kPW * 32, smem1);
        if (occ < 1) return 1;
        long long grid = (long long)sms * occ;
        if (grid > pp.n_slots) grid = pp.n_slots;
        kern<<<(unsigned)grid, kPW * 32, smem1, stream>>>(p, pp);
        if (cudaGetLastError() != cudaSuccess) return 1;
        if (n_launches) (*n_launches)++;
    }
    for (int i = 0; i < 5; i++)
        if (used[i]) cudaStreamWaitEvent(stream, ss.join[i], 0);
    (void)n_covered_from;
    return 0;
}

template <int MODEL> int dispatch_panel(const CgSweepParams &p, cudaStream_t stream, int *n_launches)
{
    const int kk = p.kk;
    if (kk < 1) return 2;
#ifdef USE_FLOAT
    if (kk <= 16) return 3;
    if (kk <= 32) return launch_panel_cfg<float, 8, 4, MODEL>(p, stream, n_launches, nullptr);
    if (kk <= 64) return launch_panel_cfg<float, 16, 4, MODEL>(p, stream, n_launches, nullptr);
    if (kk <= 128) return launch_panel_cfg<float, 8, 16, MODEL>(p, stream, n_launches, nullptr);
#else
    if (kk <= 16) return 3;
    if (kk <= 32) return launch_panel_cfg<double, 4, 8, MODEL>(p, stream, n_launches, nullptr);
    if (kk <= 64) return launch_panel_cfg<double, 4, 16, MODEL>(p, stream, n_launches, nullptr);
#endif
    return 3;
}

}  // namespace

// 0 = launched, 3 = this shape is not covered (nothing was launched: use another variant), other = error
int launch_explicit_cg_sweep_panel(const CgSweepParams &p, cudaStream_t stream, int *n_launches)
{
    return (p.gram || p.qvec || p.solve_all_rows) ? dispatch_panel<kModelCollective>(p, stream, n_launches)
                                                  : dispatch_panel<kModelExplicit>(p, stream, n_launches);
}
int launch_implicit_cg_sweep_panel(const CgSweepParams &p, cudaStream_t stream, int *n_launches)
{
    return dispatch_panel<kModelImplicit>(p, stream, n_launches);
}

}  // namespace cmfb200
