// Conjugate-gradient half-sweep with the row's gathered opposing-factor rows STAGED IN SHARED MEMORY.
//
// A row is solved by a team of TW warps (1, 4 or 16, picked by the row's number of stored entries).  The team
// owns a ring of NS slots in shared memory; a slot holds one chunk of 16*TW stored entries: for each entry the
// complete opposing-factor row (k coordinates + its bias slot), fetched by ONE bulk asynchronous copy
// (cp.async.bulk, the TMA engine's non-tensor mode; SASS UBLKCP) issued by the lane that owns the entry and
// completing on the slot's mbarrier, plus the entry's value.  Issue runs up to NS chunks ahead of use, so many
// KB per warp are in flight with no registers held.  When the whole row fits in the ring (nnz <= NS*16*TW) it
// is gathered exactly once and the 1 + max_cg_steps passes of the CG all read shared memory: the single-gather
// traffic model of SURVEY.md 8(d).  Longer rows stream through the ring once per pass.
//
// The CG algebra itself is cg_row.cuh (shared with the direct-gather kernel in sweep_cg.cu).
#include "cg_row.cuh"
#include <cstdint>
#include <cstdlib>

namespace cmfb200 {

namespace {

// ---------------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned), completion on `bar`
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int kEntriesPerWarp(int G) { return G > 16 ? G : 16; }

// ---------------------------------------------------------------------------------------------------------
// Gather policy: entries come from the team's shared-memory ring
// ---------------------------------------------------------------------------------------------------------
template <typename T, int C, int L, int TW> struct StagedGather {
    typedef Layout<T, C, L> Lay;
    static constexpr int G = 32 / L;
    static constexpr int EPW = kEntriesPerWarp(G);   // entries per warp per chunk
    static constexpr int SUB = EPW / G;              // sub-iterations per chunk
    static constexpr int CH = EPW * TW;              // entries per chunk (= per ring slot)

    const CgSweepParams &p;
    T *rows;             // [NS][CH][RS]
    T *xs;               // [NS][CH]
    uint64_t *full;      // [NS]  count = TW*EPW arrivals + transaction bytes
    uint64_t *empty;     // [NS]  count = TW arrivals
    int NS, RS;
    int lane, wt, g, l;
    // ring state, identical in every thread of the team
    int head;                        // slot the next issued chunk goes to
    uint32_t full_par, empty_par;    // parity of the next wait on each slot's barriers
    // row state
    size_t beg;
    int nnz, nch, base_slot, issued, consumed, pass_idx, max_passes;
    bool resident;
    // look-ahead
    size_t next_beg;
    int next_nnz, pre_base;
    bool pre_issued;

    __device__ __forceinline__ StagedGather(const CgSweepParams &p_, T *rows_, T *xs_, uint64_t *full_, uint64_t *empty_,
                                            int NS_, int RS_, int warp_in_team)
        : p(p_), rows(rows_), xs(xs_), full(full_), empty(empty_), NS(NS_), RS(RS_), wt(warp_in_team), head(0), full_par(0),
          empty_par(0xffffffffu), next_beg(0), next_nnz(0), pre_base(0), pre_issued(false)
    {
        lane = threadIdx.x & 31;
        g = lane / L;
        l = lane % L;
    }

    // entry handled by group g in sub-iteration t of a chunk; for 4-lane groups the two groups that share a
    // quarter-warp take entries 4 apart, which puts their 16-byte reads on disjoint banks (entry stride/4 is odd)
    __device__ __forceinline__ int local_entry(int t) const
    {
        int gi = g;
        if (L == 4) gi = (g >> 1) + (G / 2) * (g & 1);
        return wt * EPW + t * G + gi;
    }

    // fill the slot at `head` with chunk c of the row whose entries start at rbeg
    __device__ __forceinline__ void issue_chunk(size_t rbeg, int rnnz, int c)
    {
        const int s = head;
        mbar_wait(&empty[s], (empty_par >> s) & 1u);
        empty_par ^= (1u << s);
        if (lane < EPW) {
            const int local = wt * EPW + lane;
            const int e = c * CH + local;
            if (e < rnnz) {
                const int col = p.X.idx[rbeg + e];
                T xv = p.X.val[rbeg + e];
                if (p.center_opp) xv -= __ldg(p.Gbias + col);   // the value is staged already reduced by the opposing bias
                xs[s * CH + local] = xv;
                const uint32_t bytes = (uint32_t)(p.ldG * sizeof(T));
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_copy_g2s(rows + ((size_t)s * CH + local) * RS, p.G + (size_t)col * (size_t)p.ldG, bytes, &full[s]);
            } else {
                mbar_arrive(&full[s]);
            }
        }
        head = (head + 1 == NS) ? 0 : head + 1;
    }

    __device__ __forceinline__ void issue_one()
    {
        issue_chunk(beg, nnz, resident ? issued : (issued % nch));
        issued++;
    }

    // the row that will be solved after the current one (nnz 0 = none / unknown)
    __device__ __forceinline__ void set_next(size_t nbeg, int nnnz)
    {
        next_beg = nbeg;
        next_nnz = nnnz;
    }

    // Gather the next row while the current one iterates out of shared memory: possible when the current row
    // is resident and the ring has room for all of the next row's chunks.
    __device__ __forceinline__ void prefetch_next()
    {
        if (pre_issued || next_nnz <= 0 || !resident) return;
        const int nch_next = (next_nnz + CH - 1) / CH;
        if (nch_next > NS - nch) return;
        pre_base = head;
        for (int c = 0; c < nch_next; c++) issue_chunk(next_beg, next_nnz, c);
        pre_issued = true;
    }

    __device__ __forceinline__ void begin_row(size_t beg_, int nnz_, int max_passes_)
    {
        beg = beg_;
        nnz = nnz_;
        nch = (nnz + CH - 1) / CH;
        resident = nch <= NS;
        consumed = 0;
        pass_idx = 0;
        max_passes = max_passes_;
        if (pre_issued) {
            // chunks are already on their way (prefetch_next during the previous row)
            base_slot = pre_base;
            issued = nch;
            pre_issued = false;
        } else {
            base_slot = head;
            issued = 0;
            const int first = nch < NS ? nch : NS;
            for (int q = 0; q < first; q++) issue_one();
        }
        next_nnz = 0;
    }

    // release everything this row still holds
    __device__ __forceinline__ void end_row()
    {
        if (resident) {
            // every slot of the row was waited for in pass 0; hand them back
            __syncwarp();
            if (lane == 0)
                for (int q = 0; q < nch; q++) {
                    int s = base_slot + q;
                    if (s >= NS) s -= NS;
                    mbar_arrive(&empty[s]);
                }
        } else {
            // chunks issued ahead for a pass that did not happen: wait for them to land, then free them
            while (consumed < issued) {
                int s = (base_slot + consumed) % NS;
                mbar_wait(&full[s], (full_par >> s) & 1u);
                full_par ^= (1u << s);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                consumed++;
            }
        }
    }

    template <int KIND>
    __device__ __forceinline__ void pass(const T (&vec)[C], T vecb, T (&acc)[C], T &accb)
    {
        const int kk = p.kk;
        const int total = resident ? nch : max_passes * nch;   // length of the row's chunk stream
        for (int c = 0; c < nch; c++) {
            int s;
            if (resident) {
                s = base_slot + c;
                if (s >= NS) s -= NS;
                if (pass_idx == 0) {
                    mbar_wait(&full[s], (full_par >> s) & 1u);
                    full_par ^= (1u << s);
                }
            } else {
                s = (base_slot + consumed) % NS;
                mbar_wait(&full[s], (full_par >> s) & 1u);
                full_par ^= (1u << s);
            }
            const int left = nnz - c * CH - wt * EPW;   // entries of this warp's share still in range
#pragma unroll
            for (int t = 0; t < SUB; t++) {
                if (t * G >= left) break;   // warp-uniform
                const int local = local_entry(t);
                const bool valid = c * CH + local < nnz;
                const T *srow = rows + ((size_t)s * CH + local) * RS;
                T v[C];
                if constexpr (Lay::VN > 1) {
#pragma unroll
                    for (int q = 0; q < C / Lay::VN; q++) {
                        const int cc = (q * L + l) * Lay::VN;
                        if (valid && cc < p.ldG) {
                            const typename VecOf<T>::type vv = *reinterpret_cast<const typename VecOf<T>::type *>(srow + cc);
                            const T *pv = reinterpret_cast<const T *>(&vv);
#pragma unroll
                            for (int e = 0; e < Lay::VN; e++) v[q * Lay::VN + e] = pv[e];
                        } else {
#pragma unroll
                            for (int e = 0; e < Lay::VN; e++) v[q * Lay::VN + e] = T(0);
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < C; j++) {
                        const int cc = j * L + l;
                        v[j] = (valid && cc < p.ldG) ? srow[cc] : T(0);
                    }
                }
                T x = valid ? xs[s * CH + local] : T(0);
                T d0 = T(0), d1 = T(0);
#pragma unroll
                for (int j = 0; j < C; j += 2) {
                    d0 = fma(v[j], vec[j], d0);
                    if (j + 1 < C) d1 = fma(v[j + 1], vec[j + 1], d1);
                }
                T d = group_sum<L>(d0 + d1);
                d += vecb;
                T coef = entry_coef<KIND>(d, x);
                if (!valid) coef = T(0);
#pragma unroll
                for (int j = 0; j < C; j++) acc[j] = fma(coef, v[j], acc[j]);
                accb += coef;
            }
            if (!resident) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                consumed++;
                if (issued < total) issue_one();
            }
        }
        if (pass_idx == 0) prefetch_next();
        pass_idx++;
    }
};

// ---------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------
template <typename T, int C, int L, int TW, int WPB> struct StagedSmem {
    typedef StagedGather<T, C, L, TW> Gat;
    static constexpr int TEAMS = WPB / TW;
    // bytes of one team's region for NS slots and entry stride RS
    __host__ __device__ static size_t team_bytes(int NS, int RS)
    {
        size_t b = (size_t)TeamScratch<T, C, L, TW>::elems() * sizeof(T);
        b = (b + 15) / 16 * 16;
        b += (size_t)NS * Gat::CH * sizeof(T);            // xs
        b = (b + 15) / 16 * 16;
        b += (size_t)NS * Gat::CH * RS * sizeof(T);       // rows
        b = (b + 15) / 16 * 16;
        b += (size_t)NS * 2 * sizeof(uint64_t);           // full + empty
        return (b + 127) / 128 * 128;
    }
};

template <typename T, int C, int L, int MODEL, int TW, int WPB, bool GRAM_SMEM>
__global__ void __launch_bounds__(WPB * 32) cg_sweep_staged_kernel(const CgSweepParams p, int first, int count, int NS, int RS)
{
    typedef Layout<T, C, L> Lay;
    typedef StagedSmem<T, C, L, TW, WPB> SM;
    typedef StagedGather<T, C, L, TW> Gat;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int team = warp / TW, wt = warp % TW;

    size_t gram_bytes = 0;
    const T *gram = p.gram;
    constexpr bool IMPLICIT = MODEL == kModelImplicit;
    if constexpr (IMPLICIT && GRAM_SMEM) {
        T *gram_sm = reinterpret_cast<T *>(smem_raw);
        const int kk = p.kk;
        for (int i = threadIdx.x; i < kk * Lay::KP; i += blockDim.x) {
            const int d = i / Lay::KP, c = i % Lay::KP;
            gram_sm[i] = (c < kk) ? p.gram[(size_t)d * kk + c] : T(0);
        }
        gram = gram_sm;
        gram_bytes = ((size_t)kk * Lay::KP * sizeof(T) + 127) / 128 * 128;
    }
    unsigned char *base = smem_raw + gram_bytes + (size_t)team * SM::team_bytes(NS, RS);
    T *scratch = reinterpret_cast<T *>(base);
    size_t off = ((size_t)TeamScratch<T, C, L, TW>::elems() * sizeof(T) + 15) / 16 * 16;
    T *xs = reinterpret_cast<T *>(base + off);
    off = (off + (size_t)NS * Gat::CH * sizeof(T) + 15) / 16 * 16;
    T *rows = reinterpret_cast<T *>(base + off);
    off = (off + (size_t)NS * Gat::CH * RS * sizeof(T) + 15) / 16 * 16;
    uint64_t *full = reinterpret_cast<uint64_t *>(base + off);
    uint64_t *empty = full + NS;

    if (wt == 0 && lane == 0) {
        for (int s = 0; s < NS; s++) {
            mbar_init(&full[s], TW * Gat::EPW);
            mbar_init(&empty[s], TW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    CgRow<T, C, L, MODEL, TW, GRAM_SMEM> solver(p, scratch, gram, wt, 1 + team);
    Gat gat(p, rows, xs, full, empty, NS, RS, wt);
    const int team_global = blockIdx.x * SM::TEAMS + team;
    const int total_teams = gridDim.x * SM::TEAMS;
    for (int i = team_global; i < count; i += total_teams) {
        const int row = p.plan.order[first + i];
        const size_t beg = p.X.ptr[row];
        const int nnz = (int)(p.X.ptr[row + 1] - beg);
        if (nnz <= 0) {
            solver.empty_row(row);
            continue;
        }
        gat.begin_row(beg, nnz, 1 + p.max_cg_steps);
        if (i + total_teams < count) {
            const int nrow = p.plan.order[first + i + total_teams];
            const size_t nb = p.X.ptr[nrow];
            gat.set_next(nb, (int)(p.X.ptr[nrow + 1] - nb));
        }
        solver.solve(row, nnz, gat);
        gat.end_row();
    }
}

struct DeviceInfo {
    int sms = 148;
    size_t smem_optin = 227 * 1024;
};

DeviceInfo device_info()
{
    DeviceInfo d;
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) d.sms = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess) d.smem_optin = (size_t)v;
    return d;
}

int env_int(const char *name, int dflt)
{
    const char *e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

template <typename T, int C, int L, bool IMPLICIT, int TW, int WPB>
int launch_bucket(const CgSweepParams &p, int first, int count, int blocks_per_sm, cudaStream_t stream, bool dry)
{
    if (count <= 0) return 0;
    typedef Layout<T, C, L> Lay;
    typedef StagedSmem<T, C, L, TW, WPB> SM;
    const DeviceInfo di = device_info();
    const int VN = VecOf<T>::N;
    int RS = p.ldG;
    if (((RS / VN) & 1) == 0) RS += VN;                    // odd number of 16-byte units per staged row
    const size_t budget = (di.smem_optin + 1024) / blocks_per_sm - 1024 - 256;   // 1 KB per block is reserved by the driver
    size_t gram_bytes = IMPLICIT ? ((size_t)p.kk * Lay::KP * sizeof(T) + 127) / 128 * 128 : 0;
    bool gram_smem = IMPLICIT && gram_bytes <= budget / 3;
    if (!gram_smem) gram_bytes = 0;
    int NS = 0;
    while (NS < 32 && gram_bytes + (size_t)SM::TEAMS * SM::team_bytes(NS + 1, RS) <= budget) NS++;
    if (NS < 2) return 3;   // does not fit: caller falls back to the direct-gather kernel
    const size_t smem = gram_bytes + (size_t)SM::TEAMS * SM::team_bytes(NS, RS);
    auto kern = gram_smem ? cg_sweep_staged_kernel<T, C, L, IMPLICIT ? kModelImplicit : kModelExplicit, TW, WPB, true>
                          : cg_sweep_staged_kernel<T, C, L, IMPLICIT ? kModelImplicit : kModelExplicit, TW, WPB, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 3;
    }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WPB * 32, smem);
    if (occ < 1) return 3;
    long long grid = (long long)di.sms * occ;
    const long long need = ((long long)count + SM::TEAMS - 1) / SM::TEAMS;
    if (grid > need) grid = need;
    if (dry) return 0;
    kern<<<(unsigned)grid, WPB * 32, smem, stream>>>(p, first, count, NS, RS);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

template <typename T, int C, int L, bool IMPLICIT> int launch_staged_cfg(const CgSweepParams &p, cudaStream_t stream)
{
    // buckets of the degree-sorted row list: [0, n_big) -> 16-warp teams, [n_big, n_mid) -> 4-warp teams, rest -> 1 warp
    const int n_rows = p.plan.n_rows;
    const int n_big = p.plan.n_big < n_rows ? p.plan.n_big : n_rows;
    const int n_mid = p.plan.n_mid < n_rows ? p.plan.n_mid : n_rows;
    const int bps_mid = env_int("CMFB200_STAGED_BPS_MID", 2), bps_small = env_int("CMFB200_STAGED_BPS_SMALL", 2);
    // all or nothing: first check that every bucket fits (dry run), then launch, largest rows first
    for (int dry = 1; dry >= 0; dry--) {
        int rc = launch_bucket<T, C, L, IMPLICIT, 16, 16>(p, 0, n_big, 1, stream, dry != 0);
        if (rc) return rc;
        rc = launch_bucket<T, C, L, IMPLICIT, 4, 8>(p, n_big, n_mid - n_big, bps_mid, stream, dry != 0);
        if (rc) return rc;
        rc = launch_bucket<T, C, L, IMPLICIT, 1, 8>(p, n_mid, n_rows - n_mid, bps_small, stream, dry != 0);
        if (rc) return rc;
    }
    return 0;
}

template <bool IMPLICIT> int dispatch_staged(const CgSweepParams &p, cudaStream_t stream)
{
    const int kk = p.kk;
#ifdef USE_FLOAT
    if (kk <= 16) return launch_staged_cfg<float, 4, 4, IMPLICIT>(p, stream);
    if (kk <= 32) return launch_staged_cfg<float, 8, 4, IMPLICIT>(p, stream);
    if (kk <= 64) return launch_staged_cfg<float, 16, 4, IMPLICIT>(p, stream);
    if (kk <= 128) return launch_staged_cfg<float, 16, 8, IMPLICIT>(p, stream);
    if (kk <= 256) return launch_staged_cfg<float, 16, 16, IMPLICIT>(p, stream);
#else
    if (kk <= 16) return launch_staged_cfg<double, 4, 4, IMPLICIT>(p, stream);
    if (kk <= 32) return launch_staged_cfg<double, 8, 4, IMPLICIT>(p, stream);
    if (kk <= 64) return launch_staged_cfg<double, 8, 8, IMPLICIT>(p, stream);
    if (kk <= 128) return launch_staged_cfg<double, 8, 16, IMPLICIT>(p, stream);
#endif
    return 3;
}

}  // namespace

// 0 = launched, 3 = this shape is not covered by the staged kernel (use the direct one), other = error
int launch_explicit_cg_sweep_staged(const CgSweepParams &p, cudaStream_t stream) { return dispatch_staged<false>(p, stream); }
int launch_implicit_cg_sweep_staged(const CgSweepParams &p, cudaStream_t stream) { return dispatch_staged<true>(p, stream); }

}  // namespace cmfb200
