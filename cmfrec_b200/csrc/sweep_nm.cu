// Exact (Cholesky) half-sweep with every row's NORMAL MATRIX built on the 5th-generation tensor cores.
//
//   explicit:  M = sum_e g_e g_e^T + diag(lam .. lam, lam_last),  rhs = sum_e x_e g_e
//              reference factors_closed_form, sparse branch, src/common.c:978-1013 + 1058-1070
//   implicit:  M = G^T G + lam I + sum_e x_e g_e g_e^T,            rhs = sum_e (x_e + 1) g_e
//              reference factors_implicit_chol src/common.c:2063-2126
//   collective: + Q, + q_i   (reference collective_closed_form_block, src/collective.c:1223-1847)
//
// sum_e g_e g_e^T is a [k x nnz_row] x [nnz_row x k] product: the row's gathered opposing rows are the K dimension
// of an MMA whose M and N are the k columns.  fp32 library only (kind::tf32); the fp64 library uses sweep_chol.cu.
//
// One persistent thread block per SM, four kinds of warps connected by mbarrier rings:
//   * 4 LOADER warps gather the opposing rows of the row's stored entries (16-byte read-only loads, issued one stage
//     ahead), split every value into HI = tf32(x) and LO = x - HI, and store both UNtransposed into the stage: a
//     gathered row is one K row of the MN-major SWIZZLE_128B_BASE32B operand layout (atoms of 32 columns x 4 entries,
//     a quarter-warp fills one 128-byte row with one 16-byte store each).  They also accumulate the right-hand side,
//     the column sums (the bias border of M) and sum x in registers while the values pass through.
//   * 1 MMA warp: one elected thread issues the tcgen05.mma stream.  k = 64: per 8 entries ONE instruction with
//     M = N = 128, A = B = [HI; LO]: the four 64 x 64 blocks of the accumulator are HI HI^T, HI LO^T, LO HI^T and
//     LO LO^T, i.e. all the terms of the split product.  k = 128: HI HI^T, LO HI^T and HI LO^T accumulated into one
//     128 x 128 accumulator.  It commits every stage back to the loaders and, at the end of a row, the accumulator to
//     the assemblers.  Accumulators rotate through the 512 columns of tensor memory; a row longer than 1024 entries is
//     cut into units of 1024 that get a fresh accumulator each (the tensor core adds into fp32 with truncation: 128
//     accumulations keep the drift at 1e-5 relative) and are summed with round-to-nearest adds.
//   * 4 ASSEMBLER warps (one per tensor-memory lane quarter) read the accumulator (tcgen05.ld), add its blocks into the
//     lower triangle of M in a matrix buffer in shared memory, add the regulariser / constant matrix / bias border /
//     right-hand side, and hand the buffer to a solver team.
//   * 7 SOLVER warps: a blocked left-looking Cholesky with every thread owning whole rows, the 8 x 8 diagonal blocks
//     factorised in registers with shuffles, the forward substitution riding along as one more row of the matrix,
//     then the backward substitution and the write-back.  k = 64: one warp per matrix, seven matrices in flight per
//     SM (the factorisation is a chain of dependent steps: what hides its latency is other matrices);
//     k = 128: two teams of three warps.
//
// Roofline: the gathers (L2 -> SM, nnz * k * 4 bytes per half-sweep) and the issue slots of the solver warps; the
// tensor pipe needs nnz / 8 MMAs of 64 (k = 64) / 192 (k = 128) cycles per SM.  See DESIGN.md.
#include "cg_row.cuh"
#include <cstdint>
#include <cstdlib>

namespace cmfb200 {

#ifdef USE_FLOAT

namespace {

// 16 warps per block (128 registers per thread): 4 assemblers, NLW loaders, 11 - NLW solvers, 1 MMA issuer
constexpr int kNmAsmWarps = 4;      // warps 0..3: one per tensor-memory lane quarter
constexpr int kNmFirstSolver = kNmAsmWarps;
constexpr int kNmMmaWarp = 15;
constexpr int kNmThreads = 16 * 32;
constexpr int kNmSeg = 1024;   // stored entries per accumulator unit

// NM_TIMING: per-role cycle accounting (developer builds only: make EXTRA=-DNM_TIMING), dumped by block 0 .. into
// p.debug[block][role slot]
#ifdef NM_TIMING
#define NM_T0() const long long nm_t0_ = clock64()
#define NM_ADD(var) var += clock64() - nm_t0_
#else
#define NM_T0() do {} while (0)
#define NM_ADD(var) do {} while (0)
#endif

__device__ __forceinline__ uint32_t nm_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void nm_mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nm_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void nm_mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = nm_smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void nm_mbar_arrive(uint64_t *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(nm_smem_u32(bar)) : "memory");
}
// n-th use (0-based) of a ring slot: the consumer waits for the n-th completion, the producer for the (n-1)-th release
__device__ __forceinline__ void nm_wait_filled(uint64_t *bar, uint32_t n) { nm_mbar_wait(bar, n & 1u); }
__device__ __forceinline__ void nm_wait_released(uint64_t *bar, uint32_t n)
{
    if (n > 0) nm_mbar_wait(bar, (n - 1u) & 1u);
}
// MN-major SWIZZLE_128B_BASE32B operand descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor; layout type 1):
// LBO = bytes between atoms of 32 columns, SBO = bytes between atoms of 4 entries
__device__ __forceinline__ uint64_t nm_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void nm_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void nm_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(nm_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void nm_tmem_ld32(uint32_t taddr, float (&v)[32])
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
    for (int e = 0; e < 32; e++) v[e] = __uint_as_float(r[e]);
}
__device__ __forceinline__ void nm_tmem_ld4(uint32_t taddr, float (&v)[4])
{
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int e = 0; e < 4; e++) v[e] = __uint_as_float(r[e]);
}
// barrier among the warps of the solver group (ids 1..) -- never bar 0, which __syncthreads uses
__device__ __forceinline__ void nm_bar(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// NLW loader warps (8: loader-heavy, for the CG solver and for rows with many entries; 4: solver-heavy)
template <int LD, int NLW> struct Nm {
    static_assert(LD == 64 || LD == 128, "padded row widths covered by the tensor-core sweep");
    static_assert(NLW == 4 || NLW == 8, "loader warps");
    static constexpr int NSW = 11 - NLW;                 // solver warps
    static constexpr int FIRST_LOADER = kNmFirstSolver + NSW;
    static constexpr int KT = 8 * NLW;                   // stored entries per stage: every loader warp brings 8 of them
    static constexpr int NS = LD == 64 ? 3 : 2;          // stages
    static constexpr int CH = LD / 4;                    // 16-byte chunks per gathered row
    static constexpr int EPW = 32 / CH;                  // entries a loader warp covers per step
    static constexpr int EPI = NLW * EPW;                // entries all loader warps cover per step (a multiple of 4)
    static constexpr int NIT = 8 / EPW;                  // steps per stage
    static constexpr int CVR = (KT + 31) / 32;           // (column, value) registers per lane and stage
    static constexpr uint32_t LBO = KT * 128u;           // bytes between atoms of 32 columns
    static constexpr int NATOM = LD / 32;                // atoms of the HI part (the LO part has as many)
    // atom order inside a stage: k = 64: HI HI LO LO EXT; k = 128: HI HI HI HI EXT LO LO LO LO
    // (the B operand [HI | EXT] (k = 128) / [HI | LO | EXT] (k = 64) must be contiguous)
    static constexpr uint32_t EXT_OFF = 4u * LBO;
    static constexpr uint32_t LO_OFF = (LD == 64 ? 2u : 5u) * LBO;
    static constexpr uint32_t STAGE = (2u * NATOM + 1u) * LBO;
    static_assert(STAGE % 1024u == 0, "stages keep the 1024-byte alignment of the swizzle pattern");
    static constexpr int ACC_COLS = 160;                 // tensor-memory columns per accumulator unit (144 used)
    static constexpr int NACC = 3;
    static constexpr int TW = LD == 64 ? 1 : 3;          // warps per solver team
    static constexpr int NB = NSW / TW;                  // matrix buffers = solver teams
    static constexpr int MS = LD + 4;                    // floats between rows of a matrix buffer
    static constexpr int MROWS = LD + 2;                 // the matrix (<= LD + 1 rows with a bias) and the right-hand side
    static constexpr int MBUF = MROWS * MS + 16 + LD + 8;   // + slack for whole-vector reads at the end + one vector of scratch
    static constexpr int INVD = MROWS * MS + 16;         // offset of that vector (reciprocal diagonal / CG direction)
    // instruction descriptor: D = f32, A = B = tf32, both MN-major, N >> 3 at bit 17, M >> 4 at bit 24 (M = 128)
    static constexpr uint32_t idesc(int n) {
        return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    }
    static constexpr size_t SMEM = 1024 + (size_t)NS * STAGE + (size_t)NB * MBUF * sizeof(float);
};

struct NmBars {
    uint64_t full[8], empty[8];            // stage ring               loader warp -> MMA -> loader warp
    uint64_t acc_full[4], acc_empty[4];    // accumulator ring         MMA -> assemblers -> MMA
    uint64_t mat_full[8], mat_empty[8];    // matrix buffers           assemblers -> solver team -> assemblers
    int mat_row[8], mat_nnz[8];            // which row a buffer holds (-1: no more rows)
    float red[4];                          // assemblers' reduction scratch
    uint32_t tmem_base;
};

// this thread block's rows: position i of its list is position blockIdx.x + i * gridDim.x of the degree-sorted order
struct NmRow {
    int row, nnz;
    size_t beg;
};
__device__ __forceinline__ bool nm_row_at(const CgSweepParams &p, int i, NmRow &r)
{
    const long long slot = (long long)blockIdx.x + (long long)i * gridDim.x;
    if (slot >= p.plan.n_rows) return false;
    r.row = p.plan.order[slot];
    r.beg = p.X.ptr[r.row];
    r.nnz = (int)(p.X.ptr[r.row + 1] - r.beg);
    return true;
}

// walks the stages of this thread block's rows in the order every role processes them; the extent of the next row is
// fetched while the current one is being processed (two dependent global loads that would otherwise sit on the
// critical path of every row)
template <int KT> struct NmStageIter {
    int i = -1;          // position in the block's row list
    NmRow r{0, 0, 0}, nx{0, 0, 0};
    bool have_nx = false, primed = false;
    int e0 = 0;
    uint32_t g = 0xffffffffu;   // index of the current stage in the block's stage sequence
    bool done = false;
    __device__ __forceinline__ bool next(const CgSweepParams &p)
    {
        if (done) return false;
        if (!primed) {
            have_nx = nm_row_at(p, 0, nx);
            primed = true;
        }
        if (i >= 0 && e0 + KT < r.nnz) {
            e0 += KT;
            g++;
            return true;
        }
        for (;;) {
            if (!have_nx) {
                done = true;
                return false;
            }
            r = nx;
            i++;
            have_nx = nm_row_at(p, i + 1, nx);
            if (r.nnz > 0) break;
        }
        e0 = 0;
        g++;
        return true;
    }
    __device__ __forceinline__ int cnt() const { return min(KT, r.nnz - e0); }
    __device__ __forceinline__ bool first_of_unit() const { return e0 % kNmSeg == 0; }
    __device__ __forceinline__ bool last_of_unit() const { return e0 + KT >= r.nnz || (e0 + KT) % kNmSeg == 0; }
};

// 1 / sqrt(d), one Newton step on the hardware approximation (NaN for d <= 0, like a failed factorisation)
__device__ __forceinline__ float nm_rsqrt(float d)
{
    const float r = rsqrtf(d);
    return r * fmaf(-0.5f * d, r * r, 1.5f);
}

// ---------------------------------------------------------------------------------------------------------
// Blocked left-looking Cholesky + both substitutions on one matrix in shared memory, by a team of TW warps.
// M: rows 0..n-1 hold the lower triangle of the SPD matrix (row stride MS; what sits above the diagonal is never
// read), row n holds the right-hand side; invd receives 1 / L[j][j].  On return the lanes of warp 0 hold the
// solution: sol[s] = a[lane + 32 s].
// ---------------------------------------------------------------------------------------------------------
template <int LD, int TW> struct NmChol {
    static constexpr int MS = LD + 4;
    static constexpr int T = TW * 32;
    static constexpr int SLOTS = (LD + 2 + T - 1) / T;   // rows per thread (at most LD + 1 matrix rows + the rhs row)
    static constexpr int SOL = (LD + 1 + 31) / 32;

    __device__ __forceinline__ static void team_sync(int bar_id)
    {
        if constexpr (TW == 1) __syncwarp();
        else nm_bar(bar_id, T);
    }

    __device__ static void solve(float *M, float *invd, int n, int tid, int bar_id, float (&sol)[SOL])
    {
        const int lane = tid & 31;
        for (int J = 0; J < n; J += 8) {
            const int w = min(8, n - J);
            const int nslots = (n + 1 - J + T - 1) / T;        // rows J .. n over the team (uniform)
            // ---- panel: rows i = J + tid + T s (including the right-hand-side row n), columns J .. J+7
            float acc[SLOTS][8];
#pragma unroll
            for (int s = 0; s < SLOTS; s++) {
                const int i = J + tid + T * s;
#pragma unroll
                for (int c = 0; c < 8; c++) acc[s][c] = 0.f;
                if (s < nslots && i <= n) {
                    const float4 a0 = *reinterpret_cast<const float4 *>(M + (size_t)i * MS + J);
                    const float4 a1 = *reinterpret_cast<const float4 *>(M + (size_t)i * MS + J + 4);
                    acc[s][0] = a0.x; acc[s][1] = a0.y; acc[s][2] = a0.z; acc[s][3] = a0.w;
                    acc[s][4] = a1.x; acc[s][5] = a1.y; acc[s][6] = a1.z; acc[s][7] = a1.w;
                }
            }
            for (int t4 = 0; t4 < J; t4 += 4) {
                float4 lc[8];
#pragma unroll
                for (int c = 0; c < 8; c++) lc[c] = *reinterpret_cast<const float4 *>(M + (size_t)min(J + c, n) * MS + t4);   // broadcast
#pragma unroll
                for (int s = 0; s < SLOTS; s++) {
                    if (s < nslots) {
                        const int i = min(J + tid + T * s, n);
                        const float4 li = *reinterpret_cast<const float4 *>(M + (size_t)i * MS + t4);
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            acc[s][c] = fmaf(-li.x, lc[c].x, acc[s][c]);
                            acc[s][c] = fmaf(-li.y, lc[c].y, acc[s][c]);
                            acc[s][c] = fmaf(-li.z, lc[c].z, acc[s][c]);
                            acc[s][c] = fmaf(-li.w, lc[c].w, acc[s][c]);
                        }
                    }
                }
            }
            // ---- the 8 x 8 diagonal block lives in slot 0 of threads 0..7 (rows J .. J+7): right-looking in registers
            //      with shuffles; the other rows of warp 0 follow along (triangular solve against the block)
            if (tid < 32) {
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const float d = __shfl_sync(CMF_FULL_MASK, acc[0][c], c);
                    const float inv = nm_rsqrt(d);
                    if (lane == c && c < w) invd[J + c] = inv;
#pragma unroll
                    for (int s = 0; s < SLOTS; s++)
                        if (s < nslots) acc[s][c] *= inv;
#pragma unroll
                    for (int c2 = c + 1; c2 < 8; c2++) {
                        const float l = __shfl_sync(CMF_FULL_MASK, acc[0][c], c2);   // L[J + c2][J + c]
#pragma unroll
                        for (int s = 0; s < SLOTS; s++)
                            if (s < nslots) acc[s][c2] = fmaf(-acc[s][c], l, acc[s][c2]);
                    }
                }
            }
            if constexpr (TW > 1) {
                // warp 0 publishes its rows (the diagonal block among them); the other warps solve against the block
                if (tid < 32) {
#pragma unroll
                    for (int s = 0; s < SLOTS; s++) {
                        const int i = J + tid + T * s;
                        if (s < nslots && i <= n) {
                            float *dst = M + (size_t)i * MS + J;
                            *reinterpret_cast<float4 *>(dst) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
                            if (w > 4) *reinterpret_cast<float4 *>(dst + 4) = make_float4(acc[s][4], acc[s][5], acc[s][6], acc[s][7]);
                        }
                    }
                }
                team_sync(bar_id);
                if (tid >= 32) {
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        if (c < w) {
                            const float inv = invd[J + c];
#pragma unroll
                            for (int s = 0; s < SLOTS; s++)
                                if (s < nslots) acc[s][c] *= inv;
#pragma unroll
                            for (int c2 = c + 1; c2 < 8; c2++) {
                                if (c2 < w) {
                                    const float l = M[(size_t)(J + c2) * MS + J + c];
#pragma unroll
                                    for (int s = 0; s < SLOTS; s++)
                                        if (s < nslots) acc[s][c2] = fmaf(-acc[s][c], l, acc[s][c2]);
                                }
                            }
                        }
                    }
                }
            }
            // ---- store the finished columns
            if (TW == 1 || tid >= 32) {
#pragma unroll
                for (int s = 0; s < SLOTS; s++) {
                    const int i = J + tid + T * s;
                    if (s < nslots && i <= n) {
                        float *dst = M + (size_t)i * MS + J;
                        *reinterpret_cast<float4 *>(dst) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
                        if (w > 4) *reinterpret_cast<float4 *>(dst + 4) = make_float4(acc[s][4], acc[s][5], acc[s][6], acc[s][7]);
                    }
                }
            }
            team_sync(bar_id);
        }
        // ---- row n now holds y = L^-1 rhs; backward substitution L^T a = y by warp 0 (column-oriented: once a_j is
        //      known, y_i -= L[j][i] a_j for i < j reads row j of L contiguously)
        if (tid < 32) {
#pragma unroll
            for (int s = 0; s < SOL; s++) {
                const int c = lane + 32 * s;
                sol[s] = c < n ? M[(size_t)n * MS + c] : 0.f;
            }
            for (int j = n - 1; j >= 0; j--) {
                const float *Lj = M + (size_t)j * MS;
                const float ij = invd[j];
                float lj[SOL];
#pragma unroll
                for (int s = 0; s < SOL; s++) {
                    const int c = lane + 32 * s;
                    lj[s] = c < j ? Lj[c] : 0.f;
                }
                float yj = 0.f;
#pragma unroll
                for (int s = 0; s < SOL; s++)
                    if ((j >> 5) == s) yj = sol[s];
                yj = __shfl_sync(CMF_FULL_MASK, yj, j & 31);
                const float aj = yj * ij;
#pragma unroll
                for (int s = 0; s < SOL; s++) {
                    const int c = lane + 32 * s;
                    if (c < j) sol[s] = fmaf(-lj[s], aj, sol[s]);
                    else if (c == j) sol[s] = aj;
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// The reference's truncated conjugate gradient on ONE row, run on the assembled normal matrix instead of on the
// stored entries: M (full symmetric, n x n, row stride MS) = sum_e g_e g_e^T [+ Q] WITHOUT the regulariser, row n of the
// buffer = sum_e x_e g_e [+ q].  Same iterates as factors_explicit_cg (src/common.c:1098-1188) /
// collective_block_cg (src/collective.c:2134-2902) up to summation order:  r = b - M a - lam a,  Ap = M p + lam p,
// warm start, absolute thresholds 1e-12 / 1e-8, at most max_cg_steps steps.  One warp; lane owns coordinates
// lane, lane + 32, lane + 64.
// ---------------------------------------------------------------------------------------------------------
template <int LD> struct NmCg {
    static constexpr int MS = LD + 4;
    static constexpr int SOL = (LD + 1 + 31) / 32;

    __device__ __forceinline__ static float dot(const float (&x)[SOL], const float (&y)[SOL])
    {
        float sacc = 0.f;
#pragma unroll
        for (int s = 0; s < SOL; s++) sacc = fmaf(x[s], y[s], sacc);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) sacc += __shfl_xor_sync(CMF_FULL_MASK, sacc, off);
        return sacc;
    }
    // y = M v (rows lane + 32 s); vec: n floats of scratch
    __device__ __forceinline__ static void symv(const float *M, float *vec, int n, int lane, const float (&v)[SOL], float (&y)[SOL])
    {
        __syncwarp();
#pragma unroll
        for (int s = 0; s < SOL; s++) {
            const int c = lane + 32 * s;
            if (c < n) vec[c] = v[s];
        }
        __syncwarp();
        const int n4 = n & ~3;
#pragma unroll
        for (int s = 0; s < SOL; s++) {
            const int i = lane + 32 * s;
            float a0 = 0.f, a1 = 0.f;
            if (i < n) {
                const float *Mi = M + (size_t)i * MS;
                for (int c4 = 0; c4 < n4; c4 += 4) {
                    const float4 m4 = *reinterpret_cast<const float4 *>(Mi + c4);
                    const float4 x4 = *reinterpret_cast<const float4 *>(vec + c4);
                    a0 = fmaf(m4.x, x4.x, a0); a1 = fmaf(m4.y, x4.y, a1);
                    a0 = fmaf(m4.z, x4.z, a0); a1 = fmaf(m4.w, x4.w, a1);
                }
                for (int c = n4; c < n; c++) a0 = fmaf(Mi[c], vec[c], a0);
            }
            y[s] = a0 + a1;
        }
    }

    __device__ static void solve(const CgSweepParams &p, const float *M, float *vec, int n, int kk, bool hb, int row, int nnz, int lane)
    {
        float *frow = p.F + (size_t)row * (size_t)p.ldF;
        float a[SOL], r[SOL], pv[SOL], ap[SOL], lamv[SOL];
        float lam = p.lam, lam_last = p.lam_last;
        if (p.scale_lam && nnz > 0) {
            lam *= (float)nnz;
            if (!p.scale_bias_const) lam_last *= (float)nnz;
        }
#pragma unroll
        for (int s = 0; s < SOL; s++) {
            const int c = lane + 32 * s;
            a[s] = c < kk ? frow[c] : 0.f;
            lamv[s] = c < kk ? lam : 0.f;
            if (hb && c == kk) {
                a[s] = p.bias_start_one ? 1.0f : p.Fbias[row];
                lamv[s] = lam_last;
            }
        }
        symv(M, vec, n, lane, a, ap);
#pragma unroll
        for (int s = 0; s < SOL; s++) {
            const int c = lane + 32 * s;
            const float b = c < n ? M[(size_t)n * MS + c] : 0.f;
            r[s] = c < n ? (b - ap[s]) - lamv[s] * a[s] : 0.f;
        }
        float r_old = dot(r, r);
        bool changed = false;
        if (!(r_old <= 1e-12f)) {
#pragma unroll
            for (int s = 0; s < SOL; s++) pv[s] = r[s];
            for (int step = 0; step < p.max_cg_steps; step++) {
                symv(M, vec, n, lane, pv, ap);
#pragma unroll
                for (int s = 0; s < SOL; s++) ap[s] = fmaf(lamv[s], pv[s], ap[s]);
                const float alpha = r_old / dot(pv, ap);
#pragma unroll
                for (int s = 0; s < SOL; s++) {
                    a[s] = fmaf(alpha, pv[s], a[s]);
                    r[s] = fmaf(-alpha, ap[s], r[s]);
                }
                changed = true;
                const float r_new = dot(r, r);
                if (r_new <= 1e-8f) break;
                const float beta = r_new / r_old;
#pragma unroll
                for (int s = 0; s < SOL; s++) pv[s] = fmaf(beta, pv[s], r[s]);
                r_old = r_new;
            }
        }
        // a row that exits before the first step is left exactly as it was, except that a bias coordinate restarted
        // from 1.0 is what the reference leaves in the matrix (see cg_row.cuh)
#pragma unroll
        for (int s = 0; s < SOL; s++) {
            const int c = lane + 32 * s;
            if (c < kk) {
                if (changed) frow[c] = a[s];
            } else if (hb && c == kk && (changed || p.bias_start_one)) {
                p.Fbias[row] = a[s];
            }
        }
    }
};

// MODEL as in cg_row.cuh: 0 explicit, 1 implicit, 2 collective.  SOLVER: 0 = truncated CG on M, 1 = Cholesky.
template <int LD, int MODEL, int SOLVER, int NLW>
__global__ void __launch_bounds__(kNmThreads, 1) nm_sweep_kernel(const CgSweepParams p)
{
    typedef Nm<LD, NLW> S;
    constexpr bool IMPLICIT = MODEL == kModelImplicit;
    constexpr bool CHOL = SOLVER == 1;
    static_assert(CHOL || (LD == 64 && !IMPLICIT), "the CG solver covers the explicit / collective models up to k = 64");
    extern __shared__ unsigned char nm_smem_raw[];
    __shared__ NmBars bars;
    unsigned char *stages = nm_smem_raw + ((1024u - (nm_smem_u32(nm_smem_raw) & 1023u)) & 1023u);
    float *mbuf = reinterpret_cast<float *>(stages + (size_t)S::NS * S::STAGE);        // [NB][MBUF]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kk = p.kk;
    const bool hb = !IMPLICIT && p.solve_bias;
    const int n = kk + (hb ? 1 : 0);

    if (warp == kNmMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(nm_smem_u32(&bars.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 8; i++) {
            nm_mbar_init(&bars.full[i], NLW);
            nm_mbar_init(&bars.empty[i], 1);
            nm_mbar_init(&bars.mat_full[i], 1);
            nm_mbar_init(&bars.mat_empty[i], 1);
        }
        for (int i = 0; i < 4; i++) {
            nm_mbar_init(&bars.acc_full[i], 1);
            nm_mbar_init(&bars.acc_empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = bars.tmem_base;

    if (warp >= S::FIRST_LOADER && warp < kNmMmaWarp) {
        // =====================================================================================================
        // LOADERS: every stage holds KT = 8 * NLW entries; warp lw brings 8 of them (entries t * EPI + lw * EPW + sub)
        // =====================================================================================================
        const int lw = warp - S::FIRST_LOADER;
        const int c = lane % S::CH;            // this lane's 16-byte chunk of a gathered row
        const int sub = lane / S::CH;          // which of the EPW entries of a step
        const int e_first = lw * S::EPW + sub; // this lane's entry of step 0; step t: + t * EPI (EPI % 4 == 0: same swizzle phase)
        // byte offset of this lane's chunk for step 0: atom of 32 columns, row of the entry, swizzled 32-byte chunk, half of it
        const uint32_t off0 = (uint32_t)(c >> 3) * S::LBO + (uint32_t)e_first * 128u +
                              ((((uint32_t)(c >> 1) & 3u) ^ (uint32_t)(e_first & 3)) << 5) + (uint32_t)(c & 1) * 16u;
        const uint32_t ext0 = S::EXT_OFF + (uint32_t)e_first * 128u + ((uint32_t)(e_first & 3) << 5);
        // the EXT atom: every row has one 16-byte piece that is rewritten per entry, the rest stays zero for good
        for (int t = tid - S::FIRST_LOADER * 32; t < S::NS * (int)(S::LBO / 16u); t += NLW * 32) {
            const int sl = t / (int)(S::LBO / 16u), w = t - sl * (int)(S::LBO / 16u);
            reinterpret_cast<float4 *>(stages + (size_t)sl * S::STAGE + S::EXT_OFF)[w] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        nm_bar(15, NLW * 32);                  // every loader warp writes into every stage

        struct StageRegs {
            bool valid = false;
            int cnt = 0;
            uint32_t g = 0;
            size_t base = 0;
            int col[S::CVR];                  // entries lane, lane + 32 of the stage (past the end: the last one repeated)
            float val[S::CVR], ob[S::CVR];    // ob: opposing bias of the entry (stays 0 when the model has none)
            __device__ StageRegs()
            {
#pragma unroll
                for (int q = 0; q < S::CVR; q++) {
                    col[q] = 0;
                    val[q] = 0.f;
                    ob[q] = 0.f;
                }
            }
        };
        NmStageIter<S::KT> it;
        auto fetch_desc = [&](StageRegs &d) {
            d.valid = it.next(p);
            if (!d.valid) return;
            d.cnt = it.cnt();
            d.g = it.g;
            d.base = it.r.beg + (size_t)it.e0;
        };
        // Every load below is UNCONDITIONAL (indices clamped into range; what is loaded past the end of a stage is
        // ignored when the stage is consumed): a load merged into a register under a predicate makes the merging move
        // wait for it on the spot, which serialises the gathers.
        auto load_cv = [&](StageRegs &d) {
            if (d.valid) {                                  // warp-uniform
#pragma unroll
                for (int q = 0; q < S::CVR; q++) {
                    const int e = min(q * 32 + lane, d.cnt - 1);
                    d.col[q] = p.X.idx[d.base + e];
                    d.val[q] = p.X.val[d.base + e];
                }
            }
        };
        auto issue_gathers = [&](StageRegs &d, float4 (&g)[S::NIT]) {
            if (!d.valid) return;                           // warp-uniform
            if (!IMPLICIT && p.center_opp) {
#pragma unroll
                for (int q = 0; q < S::CVR; q++) d.ob[q] = __ldg(p.Gbias + d.col[q]);
            }
            const float4 *gbase = reinterpret_cast<const float4 *>(p.G) + c;
            const size_t ld4 = (size_t)p.ldG / 4;
#pragma unroll
            for (int t = 0; t < S::NIT; t++) {
                const int col = __shfl_sync(CMF_FULL_MASK, d.col[(t * S::EPI) >> 5], (t * S::EPI + e_first) & 31);
                g[t] = __ldg(gbase + (size_t)col * ld4);
            }
        };
        long long t_wait = 0, t_total = clock64(), n_stage = 0;
        // One turn of the software pipeline: start the (column, value) loads of stage x+2 and the gathers of stage x+1,
        // then split / store stage x.  The register sets rotate by NAME (the turns are written out with the arguments
        // permuted): a register-to-register rotation would wait for the loads it moves.
        auto turn = [&](StageRegs &dc, StageRegs &dn, StageRegs &dnn, float4 (&gc)[S::NIT], float4 (&gn)[S::NIT]) -> bool {
            if (!dc.valid) return false;
            n_stage++;
            fetch_desc(dnn);
            load_cv(dnn);
            issue_gathers(dn, gn);
            const uint32_t sl = dc.g % S::NS;
            {
                NM_T0();
                nm_wait_released(&bars.empty[sl], dc.g / S::NS);
                NM_ADD(t_wait);
            }
            unsigned char *hi = stages + (size_t)sl * S::STAGE;
            const int cnt8 = (dc.cnt + 7) & ~7;
#pragma unroll
            for (int t = 0; t < S::NIT; t++) {
                const int e = t * S::EPI + e_first;
                const float val = __shfl_sync(CMF_FULL_MASK, dc.val[(t * S::EPI) >> 5], e & 31);
                const float ob = __shfl_sync(CMF_FULL_MASK, dc.ob[(t * S::EPI) >> 5], e & 31);
                if (e < cnt8) {
                    float u[4] = {gc[t].x, gc[t].y, gc[t].z, gc[t].w};
                    float coef, one = 1.0f;
                    if (IMPLICIT) {
                        // x g g^T = (sqrt(x) g)(sqrt(x) g)^T and (x + 1) g = ((x + 1) / sqrt(x)) (sqrt(x) g), x > 0
                        // (the launcher routes anything else to the FMA kernel)
                        const float sq = sqrtf(val);
                        coef = (val + 1.0f) / sq;
#pragma unroll
                        for (int q = 0; q < 4; q++) u[q] *= sq;
                    } else {
                        coef = val - ob;
                    }
                    if (e >= dc.cnt) {                       // padding up to a whole MMA step
                        u[0] = u[1] = u[2] = u[3] = 0.f;
                        coef = 0.f;
                        one = 0.f;
                    }
                    // HI = the tf32 part of x (what the tensor core reads of an fp32 word: the low 13 mantissa bits
                    // dropped), LO = x - HI exactly.  [cvt.rna.tf32 is emulated with four instructions on this part;
                    // the truncating split costs one and is covered by the LO LO^T block of the product]
                    float h[4], l[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        h[q] = __uint_as_float(__float_as_uint(u[q]) & 0xffffe000u);
                        l[q] = u[q] - h[q];
                    }
                    const uint32_t off = off0 + (uint32_t)(t * S::EPI) * 128u;
                    *reinterpret_cast<float4 *>(hi + off) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4 *>(hi + S::LO_OFF + off) = make_float4(l[0], l[1], l[2], l[3]);
                    if (c == 0) {
                        // columns 128..131 of the B operand: the right-hand-side coefficient (split like the rest) and a one,
                        // so that the tensor core also delivers sum_e coef_e g_e and sum_e g_e
                        const float ch = __uint_as_float(__float_as_uint(coef) & 0xffffe000u);
                        *reinterpret_cast<float4 *>(hi + ext0 + (uint32_t)(t * S::EPI) * 128u) = make_float4(ch, one, coef - ch, 0.f);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) nm_mbar_arrive(&bars.full[sl]);
            return true;
        };
        StageRegs d0, d1, d2;
        float4 g0[S::NIT], g1[S::NIT];
        fetch_desc(d0);
        load_cv(d0);
        fetch_desc(d1);
        load_cv(d1);
        issue_gathers(d0, g0);
        for (;;) {
            if (!turn(d0, d1, d2, g0, g1)) break;
            if (!turn(d1, d2, d0, g1, g0)) break;
            if (!turn(d2, d0, d1, g0, g1)) break;
            if (!turn(d0, d1, d2, g1, g0)) break;
            if (!turn(d1, d2, d0, g0, g1)) break;
            if (!turn(d2, d0, d1, g1, g0)) break;
        }
#ifdef NM_TIMING
        if (lane == 0 && p.debug && lw < 4) {
            long long *o = reinterpret_cast<long long *>(p.debug) + (size_t)blockIdx.x * 32 + lw * 3;
            o[0] = clock64() - t_total; o[1] = t_wait; o[2] = n_stage;
        }
#endif
    } else if (warp == kNmMmaWarp) {
        // =====================================================================================================
        // MMA ISSUER
        // =====================================================================================================
        if (lane == 0) {
            uint32_t gunit = 0, gstage = 0;
            long long t_acc = 0, t_full = 0, t_total = clock64();
            // descriptors differ from stage to stage and step to step only in the start-address field (bits 0..13, 16-byte
            // units; shared memory addresses stay below 2^18)
            const uint64_t desc0 = nm_desc(nm_smem_u32(stages), S::LBO, 512u);
            NmRow r, nx;
            bool have = nm_row_at(p, 0, r), have_nx = false;
            for (int i = 0; have; i++, r = nx, have = have_nx) {
                have_nx = nm_row_at(p, i + 1, nx);       // in flight while this row is multiplied
                for (int e0 = 0; e0 < r.nnz; e0 += S::KT, gstage++) {
                    const uint32_t a = gunit % S::NACC;
                    const bool first = e0 % kNmSeg == 0;
                    if (first) {
                        NM_T0();
                        nm_wait_released(&bars.acc_empty[a], gunit / S::NACC);
                        NM_ADD(t_acc);
                    }
                    const uint32_t sl = gstage % S::NS;
                    {
                        NM_T0();
                        nm_wait_filled(&bars.full[sl], gstage / S::NS);
                        NM_ADD(t_full);
                    }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t dcol = tmem_base + a * S::ACC_COLS;
                    const int nk8 = (min(S::KT, r.nnz - e0) + 7) >> 3;
                    uint64_t d_hi = desc0 + (uint64_t)(sl * (S::STAGE >> 4));
                    for (int k8 = 0; k8 < nk8; k8++, d_hi += 64) {
                        const uint32_t accumulate = (k8 > 0 || !first) ? 1u : 0u;
                        if (LD == 64) {
                            // A = [HI; LO] (M = 128), B = [HI | LO | EXT] (N = 144): blocks HI HI^T, HI LO^T, LO HI^T, LO LO^T
                            // and the four extra columns
                            nm_umma(dcol, d_hi, d_hi, S::idesc(144), accumulate);
                        } else {
                            // one 128 x 144 accumulator: HI [HI | EXT]^T + LO [HI | EXT]^T + HI LO^T
                            const uint64_t d_lo = d_hi + (S::LO_OFF >> 4);
                            nm_umma(dcol, d_hi, d_hi, S::idesc(144), accumulate);
                            nm_umma(dcol, d_lo, d_hi, S::idesc(144), 1u);
                            nm_umma(dcol, d_hi, d_lo, S::idesc(128), 1u);
                        }
                    }
                    nm_commit(&bars.empty[sl]);
                    if (e0 + S::KT >= r.nnz || (e0 + S::KT) % kNmSeg == 0) {
                        nm_commit(&bars.acc_full[a]);
                        gunit++;
                    }
                }
            }
#ifdef NM_TIMING
            if (p.debug) {
                long long *o = reinterpret_cast<long long *>(p.debug) + (size_t)blockIdx.x * 32 + 12;
                o[0] = clock64() - t_total; o[1] = t_acc; o[2] = t_full;
            }
#endif
        }
        __syncwarp();
    } else if (warp < kNmAsmWarps) {
        // =====================================================================================================
        // ASSEMBLERS
        // =====================================================================================================
        constexpr int GT = kNmAsmWarps * 32;
        const int q = warp;                            // tensor-memory lane quarter of this warp
        uint32_t gunit = 0, gmat = 0;
        long long t_accw = 0, t_matw = 0, t_total = clock64();
        NmRow r, nx;
        bool have = nm_row_at(p, 0, r), have_nx = false;
        for (int i = 0; have; i++, r = nx, have = have_nx) {
            have_nx = nm_row_at(p, i + 1, nx);       // in flight while this row is assembled
            const int nnz = r.nnz;
            float *frow = p.F + (size_t)r.row * (size_t)p.ldF;
            const bool solve_it = nnz > 0 || (MODEL != kModelExplicit && p.solve_all_rows);
            if (!solve_it) {
                if (IMPLICIT || MODEL == kModelCollective) {
                    // implicit: A := 0 up front (src/common.c:3334); collective without any information: zeroed too
                    for (int cc = tid; cc < kk; cc += GT) frow[cc] = 0.f;
                    if (MODEL == kModelCollective && hb && tid == 0) p.Fbias[r.row] = 0.f;
                } else if (hb && p.bias_start_one && tid == 0) {
                    p.Fbias[r.row] = 1.0f;                                   // see sweep_cg.cu
                }
                continue;
            }
            const uint32_t b = gmat % S::NB;
            {
                NM_T0();
                nm_wait_released(&bars.mat_empty[b], gmat / S::NB);
                NM_ADD(t_matw);
            }
            float *M = mbuf + (size_t)b * S::MBUF;
            const int units = (nnz + kNmSeg - 1) / kNmSeg;
            if (units == 0) {
                for (int t = tid; t < (n + 1) * S::MS; t += GT) M[t] = 0.f;
                nm_bar(1, GT);
            }
            // sum x of the row (right-hand side of the bias coordinate): the only sum the tensor core does not deliver
            float sx = 0.f;
            if (hb && nnz > 0) {
                for (int e = tid; e < nnz; e += GT) {
                    const float x = p.X.val[r.beg + e];
                    sx += p.center_opp ? x - __ldg(p.Gbias + p.X.idx[r.beg + e]) : x;
                }
            }
            for (int u = 0; u < units; u++) {
                const uint32_t a = gunit % S::NACC;
                {
                    NM_T0();
                    nm_wait_filled(&bars.acc_full[a], gunit / S::NACC);
                    NM_ADD(t_accw);
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + a * S::ACC_COLS;
                if (LD == 64) {
                    // lanes 0..63: rows of [HI HI^T | HI LO^T | HI ext], lanes 64..127: rows of [LO HI^T | LO LO^T | LO ext];
                    // M[i][j] = the sum of the four blocks.  Quarters 2, 3 go first, quarters 0, 1 add on top.
                    const int mi = 32 * (q & 1) + lane;
#pragma unroll 1
                    for (int phase = 0; phase < 2; phase++) {
                        if ((phase == 0) == (q >= 2)) {
                            const bool add = phase == 1 || u > 0;
#pragma unroll 1
                            for (int c0 = 0; c0 < 64; c0 += 32) {
                                if (c0 >= kk) break;
                                float v[32], w2[32];
                                nm_tmem_ld32(taddr + c0, v);
                                nm_tmem_ld32(taddr + 64 + c0, w2);
                                if (mi < kk) {
                                    float *dst = M + (size_t)mi * S::MS + c0;
#pragma unroll
                                    for (int e = 0; e < 32; e += 4) {
                                        float4 o = make_float4(v[e] + w2[e], v[e + 1] + w2[e + 1], v[e + 2] + w2[e + 2], v[e + 3] + w2[e + 3]);
                                        if (add) {
                                            const float4 old = *reinterpret_cast<const float4 *>(dst + e);
                                            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                                        }
                                        *reinterpret_cast<float4 *>(dst + e) = o;
                                    }
                                }
                            }
                            float x4[4];
                            nm_tmem_ld4(taddr + 128, x4);
                            if (mi < kk) {
                                float rs = x4[0] + x4[2], cs = x4[1];
                                if (add) {
                                    rs += M[(size_t)n * S::MS + mi];
                                    if (hb) cs += M[(size_t)kk * S::MS + mi];
                                }
                                M[(size_t)n * S::MS + mi] = rs;
                                if (hb) M[(size_t)kk * S::MS + mi] = cs;
                            }
                        }
                        if (phase == 0) nm_bar(1, GT);
                    }
                } else {
                    const int mi = 32 * q + lane;
#pragma unroll 1
                    for (int c0 = 0; c0 < LD; c0 += 32) {
                        if (c0 >= kk) break;
                        float v[32];
                        nm_tmem_ld32(taddr + c0, v);
                        if (mi < kk) {
                            float *dst = M + (size_t)mi * S::MS + c0;
#pragma unroll
                            for (int e = 0; e < 32; e += 4) {
                                float4 o = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                                if (u > 0) {
                                    const float4 old = *reinterpret_cast<const float4 *>(dst + e);
                                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                                }
                                *reinterpret_cast<float4 *>(dst + e) = o;
                            }
                        }
                    }
                    float x4[4];
                    nm_tmem_ld4(taddr + 128, x4);
                    if (mi < kk) {
                        float rs = x4[0] + x4[2], cs = x4[1];
                        if (u > 0) {
                            rs += M[(size_t)n * S::MS + mi];
                            if (hb) cs += M[(size_t)kk * S::MS + mi];
                        }
                        M[(size_t)n * S::MS + mi] = rs;
                        if (hb) M[(size_t)kk * S::MS + mi] = cs;
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                nm_bar(1, GT);
                if (tid == 0) nm_mbar_arrive(&bars.acc_empty[a]);
                gunit++;
            }
            // ---- regulariser (Cholesky: on the diagonal; CG keeps it apart like the reference), bias border, corner
            float lam = p.lam, lam_last = p.lam_last;
            if (!IMPLICIT && p.scale_lam && nnz > 0) {
                lam *= (float)nnz;
                if (!p.scale_bias_const) lam_last *= (float)nnz;
            }
            if (hb) {
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) sx += __shfl_xor_sync(CMF_FULL_MASK, sx, off);
                if (lane == 0) bars.red[q] = sx;
            }
            for (int cc = tid; cc < kk; cc += GT) {
                if (MODEL != kModelExplicit && p.qvec) M[(size_t)n * S::MS + cc] += p.qvec[(size_t)r.row * (size_t)p.ldq + cc];
                if (hb && !CHOL) M[(size_t)cc * S::MS + kk] = M[(size_t)kk * S::MS + cc];   // CG reads whole rows
                if (CHOL) M[(size_t)cc * S::MS + cc] += lam;
            }
            if (MODEL != kModelExplicit && p.gram) {
                // the constant matrix (row-major kk x kk in global memory, L2-resident); rows are disjoint from the above
                nm_bar(1, GT);
                for (int t = tid; t < kk * kk; t += GT) {
                    const int ii = t / kk, jj = t - ii * kk;
                    M[(size_t)ii * S::MS + jj] += __ldg(p.gram + t);
                }
            }
            nm_bar(1, GT);
            if (tid == 0) {
                if (hb) {
                    M[(size_t)n * S::MS + kk] = (bars.red[0] + bars.red[1]) + (bars.red[2] + bars.red[3]);
                    M[(size_t)kk * S::MS + kk] = (float)nnz + (CHOL ? lam_last : 0.f);
                }
                bars.mat_row[b] = r.row;
                bars.mat_nnz[b] = nnz;
                nm_mbar_arrive(&bars.mat_full[b]);
            }
            gmat++;
        }
#ifdef NM_TIMING
        if (tid == 0 && p.debug) {
            long long *o = reinterpret_cast<long long *>(p.debug) + (size_t)blockIdx.x * 32 + 16;
            o[0] = clock64() - t_total; o[1] = t_accw; o[2] = t_matw; o[3] = gmat;
        }
#endif
        // no more rows: one last hand-over of every buffer tells its team to stop
        for (uint32_t t = 0; t < (uint32_t)S::NB; t++) {
            const uint32_t b = gmat % S::NB;
            nm_wait_released(&bars.mat_empty[b], gmat / S::NB);
            if (tid == 0) {
                bars.mat_row[b] = -1;
                nm_mbar_arrive(&bars.mat_full[b]);
            }
            gmat++;
        }
    } else {
        // =====================================================================================================
        // SOLVERS: team b owns matrix buffer b
        // =====================================================================================================
        const int sw = warp - kNmFirstSolver;      // < NSW (the loaders and the MMA warp were branched off above)
        const int b = sw / S::TW;
        const int ttid = tid - (kNmFirstSolver + b * S::TW) * 32;
        float *M = mbuf + (size_t)b * S::MBUF;
        long long t_w = 0, t_total = clock64(), n_solved = 0;
        for (uint32_t use = 0; b < S::NB; use++) {
            {
                NM_T0();
                nm_wait_filled(&bars.mat_full[b], use);
                NM_ADD(t_w);
            }
            n_solved++;
            const int row = bars.mat_row[b];
            if (row < 0) break;
            if constexpr (CHOL) {
                float sol[NmChol<LD, S::TW>::SOL];
                NmChol<LD, S::TW>::solve(M, M + S::INVD, n, ttid, 2 + b, sol);
                if (ttid < 32) {
                    float *frow = p.F + (size_t)row * (size_t)p.ldF;
#pragma unroll
                    for (int s = 0; s < NmChol<LD, S::TW>::SOL; s++) {
                        const int cc = lane + 32 * s;
                        if (cc < kk) frow[cc] = sol[s];
                        else if (cc == kk && hb) p.Fbias[row] = sol[s];
                    }
                }
                NmChol<LD, S::TW>::team_sync(2 + b);
            } else {
                NmCg<LD>::solve(p, M, M + S::INVD, n, kk, hb, row, bars.mat_nnz[b], lane);
                __syncwarp();
            }
            if (ttid == 0) nm_mbar_arrive(&bars.mat_empty[b]);
        }
#ifdef NM_TIMING
        if (ttid == 0 && b == 0 && p.debug) {
            long long *o = reinterpret_cast<long long *>(p.debug) + (size_t)blockIdx.x * 32 + 20;
            o[0] = clock64() - t_total; o[1] = t_w; o[2] = n_solved;
        }
#endif
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kNmMmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int LD, int MODEL, int SOLVER, int NLW> int launch_nm_cfg(const CgSweepParams &p, cudaStream_t stream)
{
    typedef Nm<LD, NLW> S;
    auto kern = nm_sweep_kernel<LD, MODEL, SOLVER, NLW>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM) != cudaSuccess) {
        cudaGetLastError();
        return 3;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = sms;
    if (grid > p.plan.n_rows) grid = p.plan.n_rows;
    if (grid < 1) return 0;
    kern<<<(unsigned)grid, kNmThreads, S::SMEM, stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// stored entries per solved row of this half-sweep (from the degree list of the plan; 0 when unknown)
inline double nm_entries_per_row(const CgSweepParams &p)
{
    if (!p.plan.host_deg || p.plan.n_rows < 1) return 0.0;
    double tot = 0;
    for (int i = 0; i < p.plan.n_rows; i++) tot += p.plan.host_deg[i];
    return tot / p.plan.n_rows;
}

template <int MODEL> int dispatch_nm_chol(const CgSweepParams &p, cudaStream_t stream)
{
    if (p.kk < 8) return 3;
    if (((uintptr_t)p.G & 15u) != 0) return 3;
    if (p.ldG == 64) {
        // the factorisation costs about as many instructions as splitting 300 stored entries: rows longer than that
        // want the loader-heavy mix of warps, shorter ones the solver-heavy one
        static const int thr = [] { const char *e = std::getenv("CMFB200_NM_LOADER_HEAVY"); return e ? std::atoi(e) : 300; }();
        if (nm_entries_per_row(p) >= thr) return launch_nm_cfg<64, MODEL, 1, 8>(p, stream);
        return launch_nm_cfg<64, MODEL, 1, 4>(p, stream);
    }
    if (p.ldG == 128) return launch_nm_cfg<128, MODEL, 1, 4>(p, stream);
    return 3;
}

}  // namespace

// 0 = launched, 3 = shape / model not covered (nothing launched: use sweep_chol.cu / the gather kernels)
int launch_explicit_chol_sweep_nm(const CgSweepParams &p, cudaStream_t stream)
{
    return (p.gram || p.qvec || p.solve_all_rows) ? dispatch_nm_chol<kModelCollective>(p, stream)
                                                  : dispatch_nm_chol<kModelExplicit>(p, stream);
}
int launch_implicit_chol_sweep_nm(const CgSweepParams &p, cudaStream_t stream)
{
    if (!p.values_positive) return 3;   // sqrt(x) weights: zero / negative confidence values take the FMA kernel
    return dispatch_nm_chol<kModelImplicit>(p, stream);
}
int launch_explicit_cg_sweep_nm(const CgSweepParams &p, cudaStream_t stream)
{
    if (p.kk < 8 || p.ldG != 64 || ((uintptr_t)p.G & 15u) != 0) return 3;
    return (p.gram || p.qvec || p.solve_all_rows) ? launch_nm_cfg<64, kModelCollective, 0, 8>(p, stream)
                                                  : launch_nm_cfg<64, kModelExplicit, 0, 8>(p, stream);
}

#else

int launch_explicit_chol_sweep_nm(const CgSweepParams &, cudaStream_t) { return 3; }
int launch_implicit_chol_sweep_nm(const CgSweepParams &, cudaStream_t) { return 3; }
int launch_explicit_cg_sweep_nm(const CgSweepParams &, cudaStream_t) { return 3; }

#endif

}  // namespace cmfb200
