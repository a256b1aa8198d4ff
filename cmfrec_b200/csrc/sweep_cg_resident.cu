// Conjugate-gradient half-sweep with the gathered opposing-factor rows RESIDENT IN SHARED MEMORY.
//
// The CG of one row makes 1 + max_cg_steps passes over the row's stored entries, and every pass needs the
// opposing-factor row of every entry.  The direct kernel (sweep_cg.cu) re-gathers them from L2 on every pass
// and is bound by the L1/LSU wavefront rate and by L2 bandwidth (profiles/).  Here every warp copies the
// opposing rows of ITS entries into its own slice of shared memory ONCE (16-byte cp.async, SASS LDGSTS, L1
// bypassed) and all passes read shared memory: the single-gather traffic model of SURVEY.md 8(d).
//
// A row is solved by a team whose size is picked from the row's number of stored entries so that the row fits
// the team's shared memory:  1, 2, 4 or 8 warps of one thread block, or a cluster of 2, 4 or 8 thread blocks
// (per-pass sums combined through distributed shared memory).  Entries are dealt to the warps of a team in
// contiguous equal shares; the assignment is the same in every pass, so a warp only ever reads what it
// staged itself and staging needs no block-level synchronisation.  Entries beyond a warp's capacity (only in
// rows longer than the largest team holds) are streamed from L2 on every pass as in the direct kernel.
//
// The CG algebra is cg_row.cuh (shared with the direct kernel): reference factors_explicit_cg
// (src/common.c:1098-1188), factors_implicit_cg (src/common.c:1914-1986), collective_block_cg
// (src/collective.c:2134-2902).
#include "cg_row.cuh"
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>

#ifndef CMF_TM_ENABLE
#define CMF_TM_ENABLE 0   // tensor memory as a second cache tier: measured neutral (profiles/README.md), compiled out; make EXTRA=-DCMF_TM_ENABLE=1
#endif
#ifndef CMF_RES_DEPTH
#define CMF_RES_DEPTH 4   // steps the streamed gathers run ahead of the arithmetic (8-lane layouts)
#endif

namespace cmfb200 {

namespace {

constexpr int kW = 8;   // warps per thread block

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gsrc)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ---- tensor memory as a second cache tier (fp32, 8 coordinates per lane): the 256 KB of TMEM of an SM are idle in this
// kernel, a tcgen05.ld costs ~12 cycles against the several hundred of an L2 gather, and it does not go through the
// LSU pipe the gathers saturate.  A warp reaches the 32 TMEM lanes of its quarter (warp id mod 4); thread t keeps its 8
// registers of a step in 8 consecutive columns of lane t.
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// four steps (32 columns) at once
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[4][8])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int t = 0; t < 4; t++)
#pragma unroll
        for (int e = 0; e < 8; e++) v[t][e] = __uint_as_float(r[t * 8 + e]);
}
__device__ __forceinline__ void tmem_st8(uint32_t, const double (&)[4]) {}      // fp64 layouts do not use the tier
template <int C> __device__ __forceinline__ void tmem_st8(uint32_t, const double (&)[C]) {}
template <int C> __device__ __forceinline__ void tmem_st8(uint32_t, const float (&)[C]) {}
constexpr int kTmemColsPerBlock = 256;   // two blocks per SM share the 512 columns
constexpr int kTmemColsPerWarp = 128;    // warps w and w + 4 share a lane quarter

// ---------------------------------------------------------------------------------------------------------
// Gather policy: this warp's entries come from its shared-memory region (first `cap` of them) or from L2
// ---------------------------------------------------------------------------------------------------------
// STREAM = false compiles the L2 path out (teams whose rows always fit)
// HOT: the opposing rows that most entries point at (the first p.n_hot rows of the opposing side's degree order) sit in a
// block-wide shared-memory table; p.hot_idx carries, per stored entry, the column with the table slot + 1 packed into bits
// 20..30 (0 = not in the table), and the streamed gathers read such rows from the table instead of L2 (one generic load
// serves both address spaces, so the pipeline of the streamed part is unchanged)
constexpr int kHotShift = 20;
constexpr int kHotColMask = (1 << kHotShift) - 1;
template <typename T, int C, int L, bool STREAM = true, bool HOT = false> struct ResidentGather {
    typedef Layout<T, C, L> Lay;
    typedef typename VecOf<T>::type Vec;
    static constexpr int G = 32 / L;
    static constexpr int VN = VecOf<T>::N;     // elements per 16 bytes
    static_assert(Lay::VN == VN, "the lane layout must be made of 16-byte pieces");
    static constexpr int KP = Lay::KP;
    // 4-lane groups read 64 bytes each and two of them share a quarter-warp: with a row stride of 16 (mod 32)
    // bytes and the two groups taking entries 4 apart their reads fall on disjoint banks.  Wider groups read
    // whole 128-byte bank rows and never conflict.
    static constexpr int PAD = (L == 4) ? VN : 0;
    static constexpr int RS = KP + PAD;        // elements between staged rows
    static constexpr int U = KP / VN;          // 16-byte pieces per staged row
    // staged row + {stored value, "one"}: the resident part of a share is padded to whole steps of G entries with
    // zero rows whose value and "one" are zero, which makes every form of the per-entry coefficient vanish there,
    // so the pass over the resident part carries no predicates at all
    static constexpr size_t kEntryBytes = (size_t)(RS + 2) * sizeof(T);

    const CgSweepParams &p;
    T *rows;          // [cap][RS]
    T *xs;            // [cap][2]   stored value (explicit model: already reduced by the opposing bias), 1.0
    int cap;          // resident entries of this warp (multiple of 8)
    size_t beg;       // first entry of this warp's share of the row
    int nnz;          // entries in this warp's share
    int gw, GW;       // index of this warp in its team / number of warps in the team (all blocks of a cluster)
    int lane, g, l, gi;
    // tensor-memory tier: the first tm_chunks full 32-entry chunks BEYOND the resident part are kept in TMEM by the residual
    // pass and read from there by the later passes (fp32, C == 8, L <= 16 only; 0 = off)
    static constexpr bool TM_OK = CMF_TM_ENABLE && sizeof(T) == 4 && C == 8 && L <= 16;
    static constexpr int TM_CHUNK_COLS = L * 8;   // a chunk is L steps of 8 columns
    uint32_t tm_addr = 0;
    int tm_chunks = 0;
    const T *hot = nullptr;   // HOT: the table, [p.n_hot][KP]

    __device__ __forceinline__ ResidentGather(const CgSweepParams &p_, T *region, int cap_, int gw_, int GW_)
        : p(p_), rows(region), xs(region + (size_t)cap_ * RS), cap(cap_), beg(0), nnz(0), gw(gw_), GW(GW_)
    {
        lane = threadIdx.x & 31;
        g = lane / L;
        l = lane % L;
        gi = (L == 4) ? ((g >> 1) + (G / 2) * (g & 1)) : g;
    }

    // The row's entries are dealt to the team's warps in contiguous, equal shares (a multiple of 8 entries each,
    // so a row of at most GW * cap entries is resident in full)
    __device__ __forceinline__ void begin(size_t row_beg, int row_nnz)
    {
        const int per = ((row_nnz + GW * 8 - 1) / (GW * 8)) * 8;
        const int first = gw * per;
        int mine = row_nnz - first;
        if (mine > per) mine = per;
        if (mine < 0) mine = 0;
        beg = row_beg + (size_t)(mine > 0 ? first : 0);
        nnz = mine;
    }

    // copy the opposing rows of this warp's first `cap` entries into its region
    __device__ __forceinline__ void stage()
    {
        __syncwarp();
        const int ldG = p.ldG;
        const int res = nnz < cap ? nnz : cap;
        for (int i = 0; i * 32 < res; i++) {
            const int e = i * 32 + lane;
            int col = -1;
            if (e < res) {
                col = p.X.idx[beg + e];
                T x = p.X.val[beg + e];
                if (p.center_opp) x -= __ldg(p.Gbias + col);
                xs[2 * e] = x;
                xs[2 * e + 1] = T(1);
            }
            if constexpr (U <= 32) {
                // 32 / U entries per round: lane -> (entry lane / U of the round, piece lane % U)
                constexpr int EPR = 32 / U;
                const int part = lane % U;
                const bool in_row = part * VN < ldG;
                T *dst = rows + (size_t)(i * 32 + lane / U) * RS + part * VN;
                const T *src = p.G + part * VN;
#pragma unroll 4
                for (int j = 0; j < U; j++) {
                    const int c = __shfl_sync(CMF_FULL_MASK, col, j * EPR + lane / U);
                    if (c >= 0 && in_row) cp_async_16(dst + (size_t)(j * EPR) * RS, src + (size_t)c * (size_t)ldG);
                }
            } else {
#pragma unroll 4
                for (int j = 0; j < U; j++) {
                    const int u = j * 32 + lane;
                    const int ent = u / U, part = u % U;
                    const int c = __shfl_sync(CMF_FULL_MASK, col, ent);
                    if (c >= 0 && part * VN < ldG)
                        cp_async_16(rows + (size_t)(i * 32 + ent) * RS + part * VN, p.G + (size_t)c * (size_t)ldG + part * VN);
                }
            }
        }
        // padding of the resident part up to whole steps of G entries: zero rows, zero value, zero "one"
        // (everything else a pass reads was either staged above or zeroed when the kernel started)
        {
            const int tail_end = (res + G - 1) / G * G;
            for (int u = res * RS + lane; u < tail_end * RS; u += 32) rows[u] = T(0);
            for (int u = 2 * res + lane; u < 2 * tail_end; u += 32) xs[u] = T(0);
        }
        cp_async_commit_wait_all();
        __syncwarp();
    }

    // once per kernel: columns the staging never writes (>= ldG) must read as zero
    __device__ __forceinline__ void clear_region()
    {
        for (int u = lane; u < cap * (RS + 2); u += 32) rows[u] = T(0);
        __syncwarp();
    }

    // ---- one step = G entries, one per group
    __device__ __forceinline__ void load_resident(int s, T (&v)[C], T &x, T &one) const
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

    // entry `ent` of a 32-entry chunk whose (column, value) pairs sit one per lane; FULLW: the opposing rows are at
    // least KP wide, so that no piece needs a bounds check
    template <bool FULLW>
    __device__ __forceinline__ void load_streamed(int col_r, T x_r, int ent, T (&v)[C], T &x, T &one) const
    {
        const int col = __shfl_sync(CMF_FULL_MASK, col_r, ent);
        x = __shfl_sync(CMF_FULL_MASK, x_r, ent);
        const bool ok = col >= 0;
        one = ok ? T(1) : T(0);
        if constexpr (HOT) {
            const int slot1 = ok ? (col >> kHotShift) : 0;
            const T *grow = slot1 ? hot + (size_t)(slot1 - 1) * KP + l * VN
                                  : p.G + (size_t)(ok ? (col & kHotColMask) : 0) * (size_t)p.ldG + l * VN;
#pragma unroll
            for (int q = 0; q < C / VN; q++) {
                if (FULLW || slot1 || (q * L + l) * VN < p.ldG) {
                    const Vec vv = *reinterpret_cast<const Vec *>(grow + q * L * VN);   // generic: shared or global
                    const T *pv = reinterpret_cast<const T *>(&vv);
#pragma unroll
                    for (int e2 = 0; e2 < VN; e2++) v[q * VN + e2] = pv[e2];
                } else {
#pragma unroll
                    for (int e2 = 0; e2 < VN; e2++) v[q * VN + e2] = T(0);
                }
            }
            return;
        }
        const T *grow = p.G + (size_t)(ok ? col : 0) * (size_t)p.ldG + l * VN;
#pragma unroll
        for (int q = 0; q < C / VN; q++) {
            if (FULLW || (q * L + l) * VN < p.ldG) {
                ldg_vec(grow + q * L * VN, &v[q * VN]);
            } else {
#pragma unroll
                for (int e2 = 0; e2 < VN; e2++) v[q * VN + e2] = T(0);
            }
        }
    }

    template <int KIND>
    __device__ __forceinline__ void step(const T (&v)[C], T x, T one, const T (&vec)[C], T vecb, T (&acc)[C], T &accb) const
    {
        T d = group_sum<L>(dot_pairs<C>(v, vec));
        d = fma(vecb, one, d);   // opposing value of the bias coordinate is 1 (vecb is 0 when there is none)
        T coef = entry_coef<KIND>(d, x);
        if (one == T(0)) coef = T(0);   // padding / past the end of a streamed chunk
        axpy_pairs<C>(coef, v, acc);
        accb += coef;
    }

    // (column, value) of entry e0 + lane of this warp's share, -1 / 0 past its end
    template <int KIND> __device__ __forceinline__ void load_chunk(int e0, int &col_r, T &x_r) const
    {
        const int e = e0 + lane;
        col_r = -1;
        x_r = T(0);
        if (e < nnz) {
            col_r = HOT ? p.hot_idx[beg + e] : p.X.idx[beg + e];
            x_r = p.X.val[beg + e];
            if (KIND == kExplicitResidual && p.center_opp) x_r -= __ldg(p.Gbias + (HOT ? (col_r & kHotColMask) : col_r));
        }
    }

    // Entries beyond the resident part, 32 at a time.  The gathers run D steps ahead of the arithmetic (D rotating
    // register buffers, no load issued twice) and the (column, value) pairs of the next chunk are fetched while the
    // current one is processed: the streamed part is bound by L2 latency, i.e. by the bytes in flight per warp.
    static constexpr int D = (C <= 8 && L >= 4) ? (CMF_RES_DEPTH < L ? CMF_RES_DEPTH : L) : 2;
    // one full 32-entry chunk streamed from L2: L steps in a straight line, every buffer index a compile-time constant;
    // STORE: the rows are also left in tensor memory (chunk `ch` of the tier) for the later passes
    template <int KIND, bool FULLW, bool STORE>
    __device__ __forceinline__ void chunk_streamed(int col_r, T x_r, int ch, const T (&vec)[C], T vecb, T (&acc)[C], T &accb) const
    {
        T v[D][C], x[D], o[D];
#pragma unroll
        for (int d = 0; d < D - 1; d++) load_streamed<FULLW>(col_r, x_r, d * G + gi, v[d], x[d], o[d]);
#pragma unroll
        for (int t = 0; t < L; t++) {
            if (t + D - 1 < L)
                load_streamed<FULLW>(col_r, x_r, (t + D - 1) * G + gi, v[(t + D - 1) % D], x[(t + D - 1) % D], o[(t + D - 1) % D]);
            step<KIND>(v[t % D], x[t % D], o[t % D], vec, vecb, acc, accb);
            if constexpr (STORE) tmem_st8(tm_addr + (uint32_t)(ch * TM_CHUNK_COLS + t * 8), v[t % D]);
        }
    }

    template <int KIND, bool FULLW>
    __device__ __forceinline__ void pass_streamed(const T (&vec)[C], T vecb, T (&acc)[C], T &accb) const
    {
        constexpr bool FIRST = KIND == kExplicitResidual || KIND == kImplicitResidual;   // the pass that fills the TMEM tier
        int col_r, col_n = -1;
        T x_r, x_n = T(0);
        load_chunk<KIND>(cap, col_r, x_r);
        int e0 = cap;
        if constexpr (TM_OK) {
            // ---- the first tm_chunks full chunks: streamed and stored by the residual pass, read from tensor memory afterwards
            for (int ch = 0; ch < tm_chunks && nnz - e0 >= 32; ch++, e0 += 32) {
                if (e0 + 32 < nnz) load_chunk<KIND>(e0 + 32, col_n, x_n);
                if constexpr (FIRST) {
                    chunk_streamed<KIND, FULLW, true>(col_r, x_r, ch, vec, vecb, acc, accb);
                } else {
#pragma unroll
                    for (int h = 0; h < L / 4; h++) {   // four steps per tcgen05.ld
                        float v4[4][8];
                        tmem_ld32(tm_addr + (uint32_t)(ch * TM_CHUNK_COLS + h * 32), v4);
#pragma unroll
                        for (int t4 = 0; t4 < 4; t4++) {
                            const int ent = (h * 4 + t4) * G + gi;
                            const int col = __shfl_sync(CMF_FULL_MASK, col_r, ent);
                            const T x = __shfl_sync(CMF_FULL_MASK, x_r, ent);
                            step<KIND>(v4[t4], x, col >= 0 ? T(1) : T(0), vec, vecb, acc, accb);
                        }
                    }
                }
                col_r = col_n;
                x_r = x_n;
                col_n = -1;
                x_n = T(0);
            }
            if (FIRST && tm_chunks > 0) tmem_wait_st();
        }
        // ---- everything else from L2 on every pass
        for (; e0 < nnz; e0 += 32) {
            if (e0 + 32 < nnz) load_chunk<KIND>(e0 + 32, col_n, x_n);
            const int left = nnz - e0;
            if (left >= 32) {
                chunk_streamed<KIND, FULLW, false>(col_r, x_r, 0, vec, vecb, acc, accb);
            } else {
                // last, partial chunk of the share
                const int nst = (left + G - 1) / G;
                for (int t = 0; t < nst; t++) {
                    T v[C], x, o;
                    load_streamed<FULLW>(col_r, x_r, t * G + gi, v, x, o);
                    step<KIND>(v, x, o, vec, vecb, acc, accb);
                }
            }
            col_r = col_n;
            x_r = x_n;
        }
    }

    template <int KIND>
    __device__ __forceinline__ void pass(const T (&vec)[C], T vecb, T (&acc)[C], T &accb) const
    {
        // resident part: software-pipelined two steps deep, no predicates
        const int res = nnz < cap ? nnz : cap;
        const int nsteps = (res + G - 1) / G;
        if (nsteps > 0) {
            T v0[C], v1[C], x0, x1, o0, o1;
            load_resident(0, v0, x0, o0);
            int s = 0;
            for (; s + 2 < nsteps; s += 2) {
                load_resident(s + 1, v1, x1, o1);
                step<KIND>(v0, x0, o0, vec, vecb, acc, accb);
                load_resident(s + 2, v0, x0, o0);
                step<KIND>(v1, x1, o1, vec, vecb, acc, accb);
            }
            if (s + 1 < nsteps) {
                load_resident(s + 1, v1, x1, o1);
                step<KIND>(v0, x0, o0, vec, vecb, acc, accb);
                step<KIND>(v1, x1, o1, vec, vecb, acc, accb);
            } else {
                step<KIND>(v0, x0, o0, vec, vecb, acc, accb);
            }
        }
        if constexpr (STREAM) {
            if (nnz > cap) {   // warp-uniform
                if (p.ldG >= KP) pass_streamed<KIND, true>(vec, vecb, acc, accb);
                else pass_streamed<KIND, false>(vec, vecb, acc, accb);
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------
struct ResidentPlan {
    // positions in the degree-sorted row list: [b8,b4) 8-warp teams, [b4,b2) 4-warp teams, [b2,b1) 2-warp teams,
    // [b1,bend) one warp per row
    int b8, b4, b2, b1, bend;
    // cumulative slot counts (a slot = one thread block's worth of rows)
    int s8, s4, s2, n_slots;
    int cap;   // resident entries per warp
    int tmem;  // use tensor memory as a second cache tier (at most two blocks per SM then)
};

template <typename T, int C, int L> struct ResidentSmem {
    typedef Layout<T, C, L> Lay;
    static constexpr int RED_STRIDE = Lay::KP + 4;
    static constexpr int STRIPE = 2 * RED_STRIDE + Lay::KP;   // scratch per warp; a team of TW warps owns TW stripes
};

// one warp per row with the constant matrix in global memory and rows as wide as the block has threads: the block's 8 warps
// share every product with the constant matrix (CgRow::coop_gram)
template <typename T, int C, int L, int MODEL, bool GRAM_SMEM, int TW> struct CoopMode {
    static constexpr int KP = Layout<T, C, L>::KP;
    // the 8 vectors and the (256 / KP) x 8 partial products must fit the 8 scratch stripes
    static constexpr bool value = TW == 1 && MODEL != kModelExplicit && !GRAM_SMEM && (KP == 256 || KP == 128) &&
                                  KP + (kW * 32 / KP) * KP <= ResidentSmem<T, C, L>::STRIPE;
};

template <typename T, int C, int L, int MODEL, bool GRAM_SMEM, int TW, bool HOT = false>
__device__ __forceinline__ void resident_row(const CgSweepParams &p, int row, size_t beg, int nnz, T *stripes, const T *gram,
                                             T *region, int cap, int w, uint32_t tm_addr, int tm_chunks, const T *hot = nullptr)
{
    typedef ResidentGather<T, C, L, TW == 8 || TW == 1, HOT> Gat;   // 2- and 4-warp teams only get rows that fit
    constexpr bool COOP = CoopMode<T, C, L, MODEL, GRAM_SMEM, TW>::value;
    const int team = w / TW, wt = w % TW;
    // one named barrier per (team size, team): a block's teams run ahead of each other by whole slots
    const int bar_id = (TW == 8) ? 1 : (TW == 4) ? 2 + team : (TW == 2) ? 4 + team : 0;
    CgRow<T, C, L, MODEL, TW, GRAM_SMEM, 1, COOP> s(p, stripes + (size_t)(team * TW) * ResidentSmem<T, C, L>::STRIPE, gram, wt, bar_id);
    if constexpr (COOP) {
        s.coop_base = stripes;
        s.coop_stride = ResidentSmem<T, C, L>::STRIPE;
    }
    if (row >= 0 && (nnz > 0 || (MODEL != kModelExplicit && p.solve_all_rows))) {
        Gat gat(p, region, cap, wt, TW);
        gat.tm_addr = tm_addr;
        gat.tm_chunks = tm_chunks;
        gat.hot = hot;
        gat.begin(beg, nnz);
        gat.stage();
        if constexpr (COOP) s.solve_coop(row, nnz, gat);
        else s.solve(row, nnz, gat);
    } else {
        if constexpr (COOP) s.coop_idle();
        if (row >= 0) s.empty_row(row);
    }
    team_barrier<TW>(bar_id);   // nobody of the team reuses scratch or regions before everybody is done with the row
}

template <typename T, int C, int L, int MODEL, bool GRAM_SMEM, int MINB = 2, bool HOT = false>
__global__ void __launch_bounds__(kW * 32, MINB) cg_resident_kernel(const CgSweepParams p, const ResidentPlan rp)
{
    typedef Layout<T, C, L> Lay;
    typedef ResidentGather<T, C, L> Gat;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *stripes = reinterpret_cast<T *>(smem_raw);
    T *gram_sm = stripes + kW * ResidentSmem<T, C, L>::STRIPE;
    T *regions = gram_sm;
    const T *gram = p.gram;
    if constexpr (MODEL != kModelExplicit && GRAM_SMEM) {
        const int kk = p.kk;
        for (int i = threadIdx.x; i < kk * Lay::KP; i += blockDim.x) {
            const int d = i / Lay::KP, c = i % Lay::KP;
            gram_sm[i] = (c < kk) ? p.gram[(size_t)d * kk + c] : T(0);
        }
        gram = gram_sm;
        regions = gram_sm + (size_t)kk * Lay::KP;
        __syncthreads();
    }
    const int w = threadIdx.x >> 5;
    T *region = regions + (size_t)w * rp.cap * (Gat::RS + 2);
    Gat(p, region, rp.cap, 0, 1).clear_region();
    // the table of the most gathered opposing rows, behind the warps' regions (columns >= ldG read as zero)
    const T *hot = nullptr;
    if constexpr (HOT) {
        T *table = regions + (size_t)kW * rp.cap * (Gat::RS + 2);
        constexpr int U = Gat::KP / Gat::VN;
        typedef typename VecOf<T>::type Vec;
        for (int i = threadIdx.x; i < p.n_hot * U; i += blockDim.x) {
            const int slot = i / U, piece = i % U;
            Vec v;
            T *pv = reinterpret_cast<T *>(&v);
#pragma unroll
            for (int e2 = 0; e2 < Gat::VN; e2++) pv[e2] = T(0);
            if (piece * Gat::VN < p.ldG) ldg_vec(p.G + (size_t)p.hot_rows[slot] * (size_t)p.ldG + piece * Gat::VN, pv);
            *reinterpret_cast<Vec *>(table + (size_t)slot * Gat::KP + piece * Gat::VN) = v;
        }
        hot = table;
        __syncthreads();
    }
    // tensor-memory tier (see tmem_st8): 256 columns per block, 128 per warp
    __shared__ uint32_t tmem_slot;
    uint32_t tm_addr = 0;
    int tm_chunks = 0;
    if constexpr (Gat::TM_OK) {
        // A kernel that contains tcgen05.alloc holds the SM's allocation permit from the moment a block starts: no second block
        // of it becomes resident until the permit is relinquished (measured: without this the kernel runs one block per SM at
        // half the speed) -- so relinquish it even when the tier is switched off at run time.
        if (w == 0) {
            if (rp.tmem)
                asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                                 (uint32_t)__cvta_generic_to_shared(&tmem_slot)),
                             "r"((uint32_t)kTmemColsPerBlock)
                             : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        if (rp.tmem) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tm_addr = tmem_slot + ((uint32_t)((w & 3) * 32) << 16) + (uint32_t)((w >> 2) * kTmemColsPerWarp);
            tm_chunks = kTmemColsPerWarp / Gat::TM_CHUNK_COLS;
        }
    }

    // order-list position of the row this warp works on in a slot (-1: none)
    auto decode = [&](int slot) -> int {
        if (slot >= rp.n_slots) return -1;
        if (slot < rp.s8) return rp.b8 + slot;
        if (slot < rp.s4) {
            const int ri = rp.b4 + (slot - rp.s8) * 2 + (w >> 2);
            return ri < rp.b2 ? ri : -1;
        }
        if (slot < rp.s2) {
            const int ri = rp.b2 + (slot - rp.s4) * 4 + (w >> 1);
            return ri < rp.b1 ? ri : -1;
        }
        const int ri = rp.b1 + (slot - rp.s2) * 8 + w;
        return ri < rp.bend ? ri : -1;
    };
    // the row id is fetched two slots ahead and its extent one slot ahead, so that neither load is waited for
    const int step = gridDim.x;
    int slot = blockIdx.x;
    int ri = decode(slot);
    int row0 = ri >= 0 ? p.plan.order[ri] : -1;
    ri = decode(slot + step);
    int row1 = ri >= 0 ? p.plan.order[ri] : -1;
    size_t beg0 = 0, end0 = 0;
    if (row0 >= 0) {
        beg0 = p.X.ptr[row0];
        end0 = p.X.ptr[row0 + 1];
    }
    for (; slot < rp.n_slots; slot += step) {
        ri = decode(slot + 2 * step);
        const int row2 = ri >= 0 ? p.plan.order[ri] : -1;
        size_t beg1 = 0, end1 = 0;
        if (row1 >= 0) {
            beg1 = p.X.ptr[row1];
            end1 = p.X.ptr[row1 + 1];
        }
        if (row0 >= 0 || (CoopMode<T, C, L, MODEL, GRAM_SMEM, 1>::value && slot >= rp.s2)) {
            const int nnz = (int)(end0 - beg0);
            if (slot < rp.s8) resident_row<T, C, L, MODEL, GRAM_SMEM, 8, HOT>(p, row0, beg0, nnz, stripes, gram, region, rp.cap, w, tm_addr, tm_chunks, hot);
#ifdef CMF_RES_TEAMS   // 2- and 4-warp teams (CMFB200_RES_MODE=0, measured slower): compiled on request only
            else if (slot < rp.s4) resident_row<T, C, L, MODEL, GRAM_SMEM, 4>(p, row0, beg0, nnz, stripes, gram, region, rp.cap, w, tm_addr, tm_chunks);
            else if (slot < rp.s2) resident_row<T, C, L, MODEL, GRAM_SMEM, 2>(p, row0, beg0, nnz, stripes, gram, region, rp.cap, w, tm_addr, tm_chunks);
#endif
            else resident_row<T, C, L, MODEL, GRAM_SMEM, 1, HOT>(p, row0, beg0, nnz, stripes, gram, region, rp.cap, w, tm_addr, tm_chunks, hot);
        }
        row0 = row1;
        beg0 = beg1;
        end0 = end1;
        row1 = row2;
    }
    if constexpr (Gat::TM_OK) {
        if (rp.tmem) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (w == 0)
                asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"((uint32_t)kTmemColsPerBlock) : "memory");
        }
    }
}

// rows [first, first + count) of the degree-sorted list, one row per cluster of CL thread blocks
template <typename T, int C, int L, int MODEL, int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(kW * 32, 2)
    cg_resident_cluster_kernel(const CgSweepParams p, int first, int count, int cap)
{
    namespace cg = cooperative_groups;
    typedef TeamScratch<T, C, L, kW> Scr;
    typedef ResidentGather<T, C, L> Gat;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw);
    T *cl_buf = scratch + Scr::elems();                     // [2][RED_STRIDE]
    T *regions = cl_buf + 2 * Scr::RED_STRIDE;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int w = threadIdx.x >> 5;
    T *region = regions + (size_t)w * cap * (Gat::RS + 2);
    Gat(p, region, cap, 0, 1).clear_region();
    const int n_clusters = gridDim.x / CL;
    int slot = blockIdx.x / CL;
    int row0 = slot < count ? p.plan.order[first + slot] : -1;
    int row1 = slot + n_clusters < count ? p.plan.order[first + slot + n_clusters] : -1;
    size_t beg0 = 0, end0 = 0;
    if (row0 >= 0) {
        beg0 = p.X.ptr[row0];
        end0 = p.X.ptr[row0 + 1];
    }
    for (; slot < count; slot += n_clusters) {
        const int row2 = slot + 2 * n_clusters < count ? p.plan.order[first + slot + 2 * n_clusters] : -1;
        size_t beg1 = 0, end1 = 0;
        if (row1 >= 0) {
            beg1 = p.X.ptr[row1];
            end1 = p.X.ptr[row1 + 1];
        }
        const int nnz = (int)(end0 - beg0);
        CgRow<T, C, L, MODEL, kW, false, CL> s(p, scratch, p.gram, w, 0);
        s.cl_buf = cl_buf;
        s.cl_rank = rank;   // only block 0 of the cluster writes the row back
        Gat gat(p, region, cap, rank * kW + w, CL * kW);
        gat.begin(beg0, nnz);
        gat.stage();
        s.solve(row0, nnz, gat);
        cluster.sync();
        row0 = row1;
        beg0 = beg1;
        end0 = end1;
        row1 = row2;
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
int env_int(const char *name, int dflt)
{
    const char *e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

struct SideStreams {
    cudaStream_t s[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[3] = {nullptr, nullptr, nullptr};
    bool ok = false;
    SideStreams()
    {
        ok = true;
        for (int i = 0; i < 3; i++) {
            ok = ok && cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming) == cudaSuccess;
        }
        ok = ok && cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) == cudaSuccess;
    }
};

// one set per calling host thread and device: concurrent fits from different host threads, or fits on several devices of
// one process, never share streams or events (the sets live as long as their thread)
SideStreams &side_streams()
{
    thread_local std::map<int, SideStreams> per_device;
    int dev = 0;
    cudaGetDevice(&dev);
    return per_device[dev];
}

// number of rows of the (descending) degree list with more than `x` stored entries
int count_gt(const int_t *deg, int n, long long x)
{
    return (int)(std::lower_bound(deg, deg + n, x, [](int_t d, long long v) { return (long long)d > v; }) - deg);
}

template <typename T, int C, int L, int MODEL, int CL>
int launch_cluster(const CgSweepParams &p, int first, int count, int cap, size_t smem, cudaStream_t stream)
{
    auto kern = cg_resident_cluster_kernel<T, C, L, MODEL, CL>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL, 1, 1);
    cfg.blockDim = dim3(kW * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
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
        return 1;
    }
    const int clusters = count < max_clusters ? count : max_clusters;
    kern<<<clusters * CL, kW * 32, smem, stream>>>(p, first, count, cap);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// which layouts carry the hot-table variant of the kernel: fp32, 32 < k <= 128, the explicit and the implicit model -- and only
// in builds made with EXTRA=-DCMF_RES_HOT_ENABLE (then switched on with CMFB200_RES_HOT=1).  Measured at ML10M shape, where the
// 256 most popular of the 10,677 items hold 19 % of the entries: A sweep 0.957 -> 1.159 ms (128 rows: 1.088): the generic
// loads, the unpacking and the smaller per-warp cache cost more than a fifth of the gathers served from shared memory saves
// (profiles/README.md), so the default build leaves it out.
template <typename T, int C, int L, int MODEL>
#ifdef CMF_RES_HOT_ENABLE
constexpr bool kHotCompiled = sizeof(T) == 4 && C == 8 && (L == 8 || L == 16) && MODEL != kModelCollective;
#else
constexpr bool kHotCompiled = false;
#endif

template <typename T, int C, int L, int MODEL, int MINB = 2>
int launch_resident_cfg(const CgSweepParams &p, cudaStream_t stream, int *n_launches)
{
    typedef Layout<T, C, L> Lay;
    typedef ResidentGather<T, C, L> Gat;
    typedef ResidentSmem<T, C, L> SM;
    typedef TeamScratch<T, C, L, kW> Scr;
    const int n_rows = p.plan.n_rows;
    const int_t *deg = p.plan.host_deg;
    if (!deg) return 3;
    if (n_rows <= 0) return 0;

    int dev = 0, sms = 148, smem_sm = 233472, smem_optin = 232448;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    // thread blocks per SM (default two); the driver reserves 1 KB per block
    const int bps = env_int("CMFB200_RES_BPS", MINB) == 1 ? 1 : MINB;
    size_t per_block = (size_t)smem_sm / bps - 1024 - 64;   // 64: the kernel's few static words (TMEM slot) must fit beside it
    if (per_block > (size_t)smem_optin) per_block = (size_t)smem_optin;

    // mode 1 (default, measured fastest): one warp per row with a shared-memory cache; mode 0: teams sized so that
    // whole rows are resident (1-8 warps, clusters of 2-8 blocks) -- fewer L2 reads but the per-pass team overhead
    // costs more instructions than the gathers it saves (profiles/README.md)
#ifdef CMF_RES_TEAMS
    const int mode = env_int("CMFB200_RES_MODE", 1);
#else
    const int mode = 1;   // the team variant is not compiled in (make EXTRA=-DCMF_RES_TEAMS)
#endif
    const size_t fixed1 = (size_t)kW * SM::STRIPE * sizeof(T);
    size_t gram_bytes = MODEL != kModelExplicit ? (size_t)p.kk * Lay::KP * sizeof(T) : 0;
    // The constant matrix goes to shared memory whenever it fits (every row reads all of it on every pass); what is
    // left is the cache of gathered rows.  In mode 1 the cache may be small or empty (wide rows): the rest of every
    // share is streamed through the pipelined gather, and shared memory that is not asked for stays L1.
    // (KP = 128 with the matrix in GLOBAL memory is served by the block-shared products, CgRow::coop_gram, and leaves the
    // 64 KB to the row cache: CMFB200_RES_GRAM_GLOBAL=1)
    const bool coop_capable = (Lay::KP == 128 || Lay::KP == 256);
    const bool gram_smem = MODEL != kModelExplicit && !(coop_capable && env_int("CMFB200_RES_GRAM_GLOBAL", 0) != 0) &&
                           (mode == 1 ? fixed1 + gram_bytes + 1024 <= per_block : gram_bytes <= per_block / 3);
    if (!gram_smem) gram_bytes = 0;
    if (fixed1 + gram_bytes >= per_block) return 3;
    // the table of hot opposing rows (mode 1, two blocks per SM): taken out of what the per-warp caches would get
    size_t hot_bytes = 0;
    if (kHotCompiled<T, C, L, MODEL> && mode == 1 && p.hot_idx && p.hot_rows && p.n_hot > 0) {
        hot_bytes = (size_t)p.n_hot * Lay::KP * sizeof(T);
        if (fixed1 + gram_bytes + hot_bytes + 1024 > per_block) hot_bytes = 0;
    }
    int cap1 = (int)((per_block - fixed1 - gram_bytes - hot_bytes) / (kW * Gat::kEntryBytes)) & ~7;
    const size_t fixedC = (size_t)(Scr::elems() + 2 * Scr::RED_STRIDE) * sizeof(T);
    int capC = (int)((per_block - fixedC) / (kW * Gat::kEntryBytes)) & ~7;
    if (mode != 1 && (cap1 < 8 || capC < 8)) return 3;   // teams need whole rows resident
    if (mode == 1) {
        // a cache of fewer than 16 entries per warp does not pay for its staging: stream everything, keep the L1
        const int min_cap = env_int("CMFB200_RES_MINCAP", 16);
        if (cap1 < min_cap) cap1 = 0;
        if (capC < min_cap) capC = 0;
    }

    // buckets of the degree-sorted row list, largest rows first
    const int pct8 = env_int("CMFB200_RES_OVF8", 100);   // % of an 8-warp team's capacity a block-level row may have
    const bool use_clusters = env_int("CMFB200_RES_CLUSTERS", 1) != 0;
    const long long cap_block = (long long)kW * cap1 * pct8 / 100;
    int nC8 = 0, nC4 = 0, nC2 = 0;
    if (use_clusters) {
        nC8 = count_gt(deg, n_rows, (long long)4 * kW * capC);
        nC4 = count_gt(deg, n_rows, (long long)2 * kW * capC);
        nC2 = count_gt(deg, n_rows, cap_block);
        if (nC4 < nC8) nC4 = nC8;
        if (nC2 < nC4) nC2 = nC4;
    }
    ResidentPlan rp;
    rp.cap = cap1;
    rp.tmem = (Gat::TM_OK && env_int("CMFB200_RES_TMEM", 1) != 0) ? 1 : 0;
    if (mode == 1) {
        // one warp per row as in the direct kernel, the first cap1 entries of every row resident, the rest streamed;
        // rows of >= 1024 entries get a thread block, rows of >= 8192 a cluster of 8
        static const int t_cluster = env_int("CMFB200_RES_T_CLUSTER", 8192), t_block = env_int("CMFB200_RES_T_BLOCK", 1024);
        nC8 = use_clusters ? count_gt(deg, n_rows, t_cluster - 1) : 0;
        nC4 = nC2 = nC8;
        rp.b8 = nC8;
        rp.b4 = rp.b2 = rp.b1 = std::max(rp.b8, count_gt(deg, n_rows, t_block - 1));
    } else {
        rp.b8 = nC2;
        rp.b4 = std::max(rp.b8, count_gt(deg, n_rows, (long long)4 * cap1));
        rp.b2 = std::max(rp.b4, count_gt(deg, n_rows, (long long)2 * cap1));
        rp.b1 = std::max(rp.b2, count_gt(deg, n_rows, (long long)cap1));
    }
    rp.bend = n_rows;
    rp.s8 = rp.b4 - rp.b8;
    rp.s4 = rp.s8 + (rp.b2 - rp.b4 + 1) / 2;
    rp.s2 = rp.s4 + (rp.b1 - rp.b2 + 3) / 4;
    rp.n_slots = rp.s2 + (rp.bend - rp.b1 + 7) / 8;

    SideStreams &ss = side_streams();
    if (!ss.ok) return 1;
    const size_t smemC = fixedC + (size_t)kW * capC * Gat::kEntryBytes;
    const int cl_first[3] = {0, nC8, nC4}, cl_count[3] = {nC8, nC4 - nC8, nC2 - nC4};
    bool forked = false;
    for (int i = 0; i < 3; i++) {
        if (cl_count[i] <= 0) continue;
        if (!forked) {
            cudaEventRecord(ss.fork, stream);
            forked = true;
        }
        cudaStreamWaitEvent(ss.s[i], ss.fork, 0);
        int rc;
        if (i == 0) rc = launch_cluster<T, C, L, MODEL, 8>(p, cl_first[i], cl_count[i], capC, smemC, ss.s[i]);
        else if (i == 1) rc = launch_cluster<T, C, L, MODEL, 4>(p, cl_first[i], cl_count[i], capC, smemC, ss.s[i]);
        else rc = launch_cluster<T, C, L, MODEL, 2>(p, cl_first[i], cl_count[i], capC, smemC, ss.s[i]);
        if (rc) return rc;
        cudaEventRecord(ss.join[i], ss.s[i]);
        if (n_launches) (*n_launches)++;
    }
    if (rp.n_slots > 0) {
        const size_t smem1 = fixed1 + gram_bytes + (size_t)kW * cap1 * Gat::kEntryBytes + hot_bytes;
        auto kern = gram_smem ? cg_resident_kernel<T, C, L, MODEL, true, MINB> : cg_resident_kernel<T, C, L, MODEL, false, MINB>;
        if constexpr (kHotCompiled<T, C, L, MODEL>) {
            if (hot_bytes)
                kern = gram_smem ? cg_resident_kernel<T, C, L, MODEL, true, MINB, true> : cg_resident_kernel<T, C, L, MODEL, false, MINB, true>;
        }
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1) != cudaSuccess) return 1;
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kW * 32, smem1);
        if (occ < 1) return 1;
        // the occupancy calculator answers 1 for any kernel that contains tcgen05.alloc (it cannot know how many columns a block
        // takes); registers (128 x 256) and shared memory (<= half an SM) are sized for two blocks, which take 256 columns each
        if (Gat::TM_OK) occ = env_int("CMFB200_RES_TMEM_BPS", 512 / kTmemColsPerBlock);
        if (env_int("CMFB200_RES_DEBUG", 0)) std::fprintf(stderr, "[resident] occ=%d smem1=%zu per_block=%zu cap1=%d tmem=%d\n", occ, smem1, per_block, cap1, rp.tmem);
        long long grid = (long long)sms * occ;
        if (grid > rp.n_slots) grid = rp.n_slots;
        kern<<<(unsigned)grid, kW * 32, smem1, stream>>>(p, rp);
        if (cudaGetLastError() != cudaSuccess) return 1;
        if (n_launches) (*n_launches)++;
    }
    for (int i = 0; i < 3; i++)
        if (cl_count[i] > 0) cudaStreamWaitEvent(stream, ss.join[i], 0);
    return 0;
}

template <int MODEL> int dispatch_resident(const CgSweepParams &p, cudaStream_t stream, int *n_launches)
{
    const int kk = p.kk;
    if (kk < 1) return 2;
#ifdef USE_FLOAT
    if (kk <= 16) return launch_resident_cfg<float, 4, 4, MODEL>(p, stream, n_launches);
    if (kk <= 32) return launch_resident_cfg<float, 8, 4, MODEL>(p, stream, n_launches);
    if (kk <= 64) {
        // 8 lanes per entry read whole 128-byte lines (half the L1 wavefronts of the 4-lane layout on streamed gathers)
        // (measured: 1.84 vs 1.87 ms / iteration at ML10M shape, 5.67 vs 7.13 ms at LastFM shape; a third thread block
        // per SM at 80 registers spills and loses: 2.15 / 7.7 ms -- profiles/README.md)
        const int v = env_int("CMFB200_RES_CFG64", 1);
        if (v == 0) return launch_resident_cfg<float, 16, 4, MODEL>(p, stream, n_launches);
        return launch_resident_cfg<float, 8, 8, MODEL>(p, stream, n_launches);
    }
    if (kk <= 128) return launch_resident_cfg<float, 8, 16, MODEL>(p, stream, n_launches);
    if (kk <= 256) return launch_resident_cfg<float, 8, 32, MODEL>(p, stream, n_launches);
#else
    if (kk <= 16) return launch_resident_cfg<double, 4, 4, MODEL>(p, stream, n_launches);
    if (kk <= 32) return launch_resident_cfg<double, 4, 8, MODEL>(p, stream, n_launches);
    if (kk <= 64) return launch_resident_cfg<double, 4, 16, MODEL>(p, stream, n_launches);
    if (kk <= 128) return launch_resident_cfg<double, 4, 32, MODEL>(p, stream, n_launches);
#endif
    return 3;
}

}  // namespace

// One translation unit per model (the Makefile compiles this file three times with -DCMF_RES_MODEL=0|1|2: the
// instantiations of one model take minutes to compile); sweep_cg_dispatch.cu routes to them.
// 0 = launched, 3 = this shape is not covered (nothing was launched: use the direct kernel), other = error
#ifndef CMF_RES_MODEL
#error "compile with -DCMF_RES_MODEL=0 (explicit), 1 (implicit) or 2 (collective)"
#endif
#if CMF_RES_MODEL == 0
int resident_sweep_explicit(const CgSweepParams &p, cudaStream_t stream, int *n_launches)
{
    return dispatch_resident<kModelExplicit>(p, stream, n_launches);
}
#elif CMF_RES_MODEL == 1
int resident_sweep_implicit(const CgSweepParams &p, cudaStream_t stream, int *n_launches)
{
    return dispatch_resident<kModelImplicit>(p, stream, n_launches);
}
#else
int resident_sweep_collective(const CgSweepParams &p, cudaStream_t stream, int *n_launches)
{
    return dispatch_resident<kModelCollective>(p, stream, n_launches);
}
#endif

}  // namespace cmfb200
