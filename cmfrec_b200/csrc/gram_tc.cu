// Gram matrix of a factor on the 5th-generation tensor cores: partial[slice] = G_slice^T G_slice with tcgen05.mma
// (kind::tf32, accumulators in tensor memory), fp32 accuracy through the 3xTF32 split.
// reference: the cblas_tsyrk call sites (src/common.c:3328 for the implicit half-sweep, src/collective.c:6276 for
// CtC / BiTBi); fp32 library only -- the fp64 library keeps the DFMA kernel of gram.cu.
//
// G is row-major [rows x LD] with LD = 64, 128 or 256 floats (the padded row width of cmf_types.h; padding columns
// are zero).  In MMA terms D[M x N] += A[M x K] B[N x K]^T with M = N = LD the columns of G and K its rows: both
// operands are the same tile of G, TRANSPOSED on its way into shared memory so that it is K-major (the layout every
// tcgen05 operand type supports; the MN-major form was measured to return zeros for kind::tf32 on this part).
//
//   * every thread block owns a contiguous slice of rows (split-K) and streams it through two shared-memory stages
//     of KT rows; a stage holds the tile twice: HI = tf32(x) (round to nearest) and LO = x - HI (exact in fp32);
//   * the tile is written in the canonical K-major SWIZZLE_128B layout the tensor core reads: atoms of 8 columns of
//     G x 128 bytes (32 rows of G), the 16-byte chunks of an atom row XOR-ed with the row index.  The lanes of a warp
//     take 32 consecutive rows of G for one 16-byte column chunk, so that each 4-byte store instruction of the
//     transposition fills one whole 128-byte atom row (conflict-free);
//   * one thread issues, per 8 rows of G, D += HI HI^T, D += HI LO^T, D += LO HI^T (the LO LO^T term is below fp32
//     rounding) and commits the stage to an mbarrier that the loaders wait on before reusing it;
//   * the epilogue reads the accumulator with tcgen05.ld (one row per thread) and writes the slice's partial;
//     gram.cu's fixed-order reduction then sums the partials, so the result stays run-to-run deterministic.
//
// The Gram is memory-bound (one pass over G): the point of the tensor cores here is that the 2 * rows * k^2 flop
// stop costing more than the read of G does (profiles/README.md: 0.5 ms -> tens of microseconds at LastFM shape).
#include "sweep.h"
#include <cstdint>
#include <cstdlib>

namespace cmfb200 {

#ifdef USE_FLOAT

namespace {

constexpr int kGramThreads = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
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
// K-major, SWIZZLE_128B shared-memory matrix descriptor (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// [0,14) start address >> 4, [16,30) leading byte offset >> 4 (not used by swizzled K-major layouts: 1), [32,46)
// stride byte offset >> 4 (between groups of 8 M / N rows), [46,48) version = 1, [61,64) layout type = 2
// `layout`: 2 = SWIZZLE_128B (K-major operands), 1 = SWIZZLE_128B_BASE32B (the only layout MN-major tf32 operands admit:
// atoms of 32 MN elements x 4 K, 128 bytes per K row, the 32-byte chunks of a row XOR-ed with the K index mod 4;
// LBO = stride between MN atoms, SBO = stride between K atoms)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
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
}

template <int LD> struct GramTc {
    static constexpr int KT = 8192 / LD;                 // rows of G per stage: 32 KB per part
    static constexpr int CHUNKS = LD / 4;                // 16-byte chunks per row of G
    static constexpr int PER_THREAD = KT * CHUNKS / kGramThreads;   // 8
    static constexpr uint32_t SBO = 1024;                // between groups of 8 columns of G (8 atom rows x 128 bytes)
    static constexpr uint32_t KATOM = (LD / 8) * SBO;    // between groups of 32 rows of G
    static constexpr uint32_t PART_BYTES = (uint32_t)KT * LD * 4;
    static constexpr uint32_t MN_LBO = (uint32_t)KT * 128u;  // MN-major staging: between atoms of 32 columns of G
    static constexpr uint32_t STAGE_BYTES = 2 * PART_BYTES;  // HI then LO
    static constexpr int M = LD >= 128 ? 128 : 64;
    static constexpr int MT = LD / M;                    // accumulator tiles stacked along M (2 at LD = 256)
    static constexpr int N = LD;
    // The tensor core adds into its fp32 accumulator with truncation, so a long chain of additions drifts low (3e-5
    // relative over the 900 additions of a LastFM-sized slice of all-positive data).  All of tensor memory is used as
    // NACC independent accumulators, one per stage in rotation, summed with round-to-nearest adds in the epilogue.
    static constexpr int TMEM_COLS = 512;
    static constexpr int NACC = TMEM_COLS / (MT * N);    // 8, 4 or 1
    static_assert(KT / 8 >= NACC, "every accumulator is first written during the first stage");
    // instruction descriptor (InstrDescriptor): D = f32 (bit 4), A = B = tf32 (2 at bits 7 and 10), both K-major
    // (bits 15, 16 clear), N >> 3 at bit 17, M >> 4 at bit 24
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    static constexpr size_t SMEM_BYTES = 2 * (size_t)STAGE_BYTES + 1024;   // + alignment slack
};

// MNMAJOR = true: the tile goes to shared memory UNtransposed (rows of G = K rows of 128-byte MN atoms, 16-byte vector
// stores, a quarter-warp fills one whole 128-byte row) and the tensor core reads it as an MN-major operand in the
// SWIZZLE_128B_BASE32B layout; false: transposed to K-major SWIZZLE_128B with 4-byte stores (round-1 path).
template <int LD, bool MNMAJOR>
__global__ void __launch_bounds__(kGramThreads, 1)
gram_tc_kernel(const float *__restrict__ G, int_t rows, int kk, int nslices, float *__restrict__ partial)
{
    typedef GramTc<LD> S;
    extern __shared__ unsigned char gram_smem_raw[];
    __shared__ uint64_t mma_done[2];
    __shared__ uint32_t tmem_base_slot;
    // 1024-byte alignment: the swizzle pattern is a function of the address bits
    unsigned char *tiles = gram_smem_raw + ((1024u - (smem_u32(gram_smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"((uint32_t)S::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(&mma_done[0], 1);
        mbar_init(&mma_done[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    const int slice = blockIdx.x;
    // slices are whole stages so that only the last one has a ragged end
    const long long tiles_total = ((long long)rows + S::KT - 1) / S::KT;
    const long long tiles_per = (tiles_total + nslices - 1) / nslices;
    const long long r_begin = (long long)slice * tiles_per * S::KT;
    long long r_end = r_begin + tiles_per * S::KT;
    if (r_end > rows) r_end = rows;
    const int ntiles = r_begin < r_end ? (int)((r_end - r_begin + S::KT - 1) / S::KT) : 0;

    // chunk c of a stage: row c % KT of the tile, 16-byte piece c / KT of that row (the lanes of a warp: 32 consecutive
    // rows, one piece); this thread owns chunks tid + i * 256
    auto load_tile = [&](int t, float4 (&buf)[S::PER_THREAD]) {
        const long long r0 = r_begin + (long long)t * S::KT;
#pragma unroll
        for (int i = 0; i < S::PER_THREAD; i++) {
            const int c = tid + i * kGramThreads;
            const long long r = r0 + (MNMAJOR ? c / S::CHUNKS : c % S::KT);
            const int j = MNMAJOR ? c % S::CHUNKS : c / S::KT;
            buf[i] = r < r_end ? __ldg(reinterpret_cast<const float4 *>(G + (size_t)r * LD) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    // element (column mn of G, row k of the tile) -> K-major swizzled position
    auto store_tile = [&](int stage, const float4 (&buf)[S::PER_THREAD]) {
        unsigned char *hi = tiles + (size_t)stage * S::STAGE_BYTES;
        unsigned char *lo = hi + S::PART_BYTES;
#pragma unroll
        for (int i = 0; i < S::PER_THREAD; i++) {
            const int c = tid + i * kGramThreads;
            const int k = MNMAJOR ? c / S::CHUNKS : c % S::KT, j = MNMAJOR ? c % S::CHUNKS : c / S::KT;
            const float x[4] = {buf[i].x, buf[i].y, buf[i].z, buf[i].w};
            if constexpr (MNMAJOR) {
                float h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    uint32_t hb;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x[e]));
                    h[e] = __uint_as_float(hb);
                    l[e] = x[e] - h[e];
                }
                const uint32_t off = (uint32_t)(j >> 3) * S::MN_LBO + (uint32_t)k * 128u + (uint32_t)((((j >> 1) & 3) ^ (k & 3)) << 5) +
                                     (uint32_t)(j & 1) * 16u;
                *reinterpret_cast<float4 *>(hi + off) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4 *>(lo + off) = make_float4(l[0], l[1], l[2], l[3]);
                continue;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int mn = 4 * j + e;
                const uint32_t off = (uint32_t)(k >> 5) * S::KATOM + (uint32_t)(mn >> 3) * S::SBO + (uint32_t)(mn & 7) * 128u +
                                     (uint32_t)((((k & 31) >> 2) ^ (mn & 7)) << 4) + (uint32_t)(k & 3) * 4u;
                uint32_t hb;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x[e]));
                const float h = __uint_as_float(hb);
                *reinterpret_cast<float *>(hi + off) = h;
                *reinterpret_cast<float *>(lo + off) = x[e] - h;
            }
        }
    };

    float4 cur[S::PER_THREAD], nxt[S::PER_THREAD];
    if (ntiles > 0) load_tile(0, cur);
    for (int t = 0; t < ntiles; t++) {
        const int stage = t & 1;
        if (t + 1 < ntiles) load_tile(t + 1, nxt);            // in flight while this tile is split and multiplied
        if (t >= 2) mbar_wait(&mma_done[stage], (uint32_t)(((t - 2) >> 1) & 1));   // the MMAs that read this stage are done
        store_tile(stage, cur);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t hi = smem_u32(tiles + (size_t)stage * S::STAGE_BYTES), lo = hi + S::PART_BYTES;
#pragma unroll 1
            for (int k8 = 0; k8 < S::KT / 8; k8++) {
                // 8 rows of G = 32 bytes along an atom row; 4 such slices per atom, then the next group of 32 rows
                const uint32_t koff = MNMAJOR ? (uint32_t)k8 * 1024u : (uint32_t)(k8 >> 2) * S::KATOM + (uint32_t)(k8 & 3) * 32u;
                const uint32_t lbo = MNMAJOR ? S::MN_LBO : 16u, sbo = MNMAJOR ? 512u : S::SBO, lay = MNMAJOR ? 1u : 2u;
                const uint32_t idesc = MNMAJOR ? (S::IDESC | (1u << 15) | (1u << 16)) : S::IDESC;
                const uint64_t b_hi = make_desc(hi + koff, lbo, sbo, lay), b_lo = make_desc(lo + koff, lbo, sbo, lay);
#pragma unroll
                for (int mt = 0; mt < S::MT; mt++) {
                    const uint32_t a_off = koff + (MNMAJOR ? (uint32_t)mt * (S::M / 32) * S::MN_LBO : (uint32_t)mt * (S::M / 8) * S::SBO);
                    const uint64_t a_hi = make_desc(hi + a_off, lbo, sbo, lay), a_lo = make_desc(lo + a_off, lbo, sbo, lay);
                    // accumulators rotate with the 8-row step (not with the stage): the shortest possible chains of
                    // truncating additions, every accumulator first written in the first stage
                    const uint32_t d = tmem_base + (uint32_t)((k8 % S::NACC) * S::MT * S::N + mt * S::N);
                    umma_tf32(d, a_hi, b_hi, idesc, (t > 0 || k8 >= S::NACC) ? 1u : 0u);
                    umma_tf32(d, a_hi, b_lo, idesc, 1u);
                    umma_tf32(d, a_lo, b_hi, idesc, 1u);
                }
            }
            umma_commit(&mma_done[stage]);   // implies tcgen05.fence::before_thread_sync
        }
#pragma unroll
        for (int i = 0; i < S::PER_THREAD; i++) cur[i] = nxt[i];
    }
    // the last commit of each stage covers everything issued before it
    if (ntiles > 0) {
        const int t = ntiles - 1;
        mbar_wait(&mma_done[t & 1], (uint32_t)((t >> 1) & 1));
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w (< 4) reads tensor-memory lanes 32w .. 32w+31, one accumulator row per thread
    float *out = partial + (size_t)slice * kk * kk;
    if (warp < 4) {
        int row_in_tile;
        bool row_ok;
        if (S::M == 128) {
            row_in_tile = warp * 32 + lane;
            row_ok = true;
        } else {   // M = 64: accumulator row 16q + i lives in lane 32q + i, i < 16
            row_in_tile = warp * 16 + lane;
            row_ok = lane < 16;
        }
#pragma unroll 1
        for (int mt = 0; mt < S::MT; mt++) {
            const int row = mt * S::M + row_in_tile;
#pragma unroll 1
            for (int c0 = 0; c0 < S::N; c0 += 32) {
                float sum[32];
#pragma unroll
                for (int e = 0; e < 32; e++) sum[e] = 0.f;
                const int nacc = ntiles > 0 ? S::NACC : 0;      // KT / 8 >= NACC: the first stage touches all of them
                for (int a = 0; a < nacc; a++) {   // warp-uniform
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * S::MT * S::N + mt * S::N + c0), v);
#pragma unroll
                    for (int e = 0; e < 32; e++) sum[e] += __uint_as_float(v[e]);
                }
                if (row_ok && row < kk) {
#pragma unroll
                    for (int e = 0; e < 32; e++)
                        if (c0 + e < kk) out[(size_t)row * kk + c0 + e] = sum[e];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)S::TMEM_COLS) : "memory");
    }
}

template <int LD> int launch_gram_tc_ld(const float *G, int_t rows, int kk, int nslices, float *partial, cudaStream_t stream)
{
    typedef GramTc<LD> S;
    static const bool mn_major = [] { const char *e = std::getenv("CMFB200_GRAM_MN"); return e ? std::atoi(e) != 0 : true; }();   // default since round 2: verified on B200 by tests/test_gpu_gram.py
    auto kern = mn_major ? gram_tc_kernel<LD, true> : gram_tc_kernel<LD, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM_BYTES) != cudaSuccess) return 1;
    kern<<<nslices, kGramThreads, S::SMEM_BYTES, stream>>>(G, rows, kk, nslices, partial);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace

// partial[slice][kk][kk] for slice < *nslices_out; returns 3 when the shape is not covered (nothing launched)
int launch_gram_partials_tc(const real_t *G, int ldG, int_t rows, int kk, int max_slices, real_t *partial, int *nslices_out,
                            cudaStream_t stream)
{
    if (ldG != 64 && ldG != 128 && ldG != 256) return 3;
    if (kk > ldG || rows < 1) return 3;
    if (((uintptr_t)G & 15u) != 0) return 3;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int kt = 8192 / ldG;
    long long ns = ((long long)rows + kt - 1) / kt;   // never more slices than stages
    if (ns > sms) ns = sms;
    if (ns > max_slices) ns = max_slices;
    if (ns < 1) ns = 1;
    *nslices_out = (int)ns;
    if (ldG == 64) return launch_gram_tc_ld<64>(G, rows, kk, (int)ns, partial, stream);
    if (ldG == 128) return launch_gram_tc_ld<128>(G, rows, kk, (int)ns, partial, stream);
    return launch_gram_tc_ld<256>(G, rows, kk, (int)ns, partial, stream);
}

#else

int launch_gram_partials_tc(const real_t *, int, int_t, int, int, real_t *, int *, cudaStream_t) { return 3; }

#endif

}  // namespace cmfb200
