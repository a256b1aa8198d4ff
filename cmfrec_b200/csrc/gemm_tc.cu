// Dense product on the 5th-generation tensor cores:  C[M x N] = A[M x K] B[N x K]^T (+ row term + column term + const),
// fp32 in / fp32 out through the 3xTF32 split (HI = tf32(x), LO = x - HI; HI HI + HI LO + LO HI), accumulators in tensor
// memory.  Both operands are K-major as they sit in memory (row-major with K contiguous), which is how the library
// stores everything it multiplies this way:
//   * scores of a batch of users against all items, A_batch B^T (reference topN / predict: cblas_tgemv at
//     src/common.c:5290-5296 once per user)                                               -> serve.cu
//   * U C and I D products of the collective model (cblas_tgemm, src/collective.c:5768-5773) -> collective.cu
// One thread block per 128 rows of A and up to 256 rows of B; K is streamed through two shared-memory stages of 32
// floats (one 128-byte swizzle atom row per operand row: canonical K-major SWIZZLE_128B, written with 16-byte stores,
// no transposition needed); one thread issues tcgen05.mma (M = 128, N = the B tile, K = 8) and commits each stage to
// an mbarrier; two accumulators alternate with the K step (the tensor core adds into fp32 with truncation: short
// chains) and are summed in the epilogue, which reads tensor memory with tcgen05.ld, one row of C per thread.
// fp32 library only; the fp64 library keeps its DFMA kernels (dense_small.cu).
#include "gemm_tc.h"
#include <cstdint>

namespace cmfb200 {

#ifdef USE_FLOAT

namespace {

constexpr int kGemmThreads = 256;
constexpr int kGemmM = 128;           // rows of A per block
constexpr int kGemmNMax = 256;        // rows of B per block
constexpr int kGemmKc = 32;           // floats of K per stage
constexpr uint32_t kAPart = kGemmM * 128u;        // bytes of one part (HI or LO) of the A tile of a stage
constexpr uint32_t kBPart = kGemmNMax * 128u;
constexpr uint32_t kStageBytes = 2 * kAPart + 2 * kBPart;   // 96 KB
constexpr size_t kGemmSmem = 2 * (size_t)kStageBytes + 1024;

__device__ __forceinline__ uint32_t g_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = g_smem_u32(bar);
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
// K-major SWIZZLE_128B operand descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address >> 4, leading
// byte offset (unused by swizzled K-major layouts: 1), stride byte offset between groups of 8 rows, version 1, layout 2
__device__ __forceinline__ uint64_t g_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void g_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void g_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(g_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void g_tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
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

// four consecutive K entries of operand row `row` starting at column k0 (zero outside the matrix)
__device__ __forceinline__ float4 g_load4(const float *__restrict__ X, int ld, long long row, long long nrows, int k0, int K, bool vec_ok)
{
    if (row >= nrows || k0 >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
    const float *src = X + (size_t)row * (size_t)ld + k0;
    if (vec_ok && k0 + 4 <= K) return __ldg(reinterpret_cast<const float4 *>(src));
    float4 v;
    v.x = __ldg(src);
    v.y = k0 + 1 < K ? __ldg(src + 1) : 0.f;
    v.z = k0 + 2 < K ? __ldg(src + 2) : 0.f;
    v.w = k0 + 3 < K ? __ldg(src + 3) : 0.f;
    return v;
}
// HI / LO parts of one 16-byte chunk into the swizzled K-major tile: row r, chunk j (of 8)
__device__ __forceinline__ void g_store_split(unsigned char *hi, unsigned char *lo, int r, int j, float4 v)
{
    const float x[4] = {v.x, v.y, v.z, v.w};
    float h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x[e]));
        h[e] = __uint_as_float(hb);
        l[e] = x[e] - h[e];
    }
    const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((j ^ (r & 7)) << 4);
    *reinterpret_cast<float4 *>(hi + off) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4 *>(lo + off) = make_float4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const float *__restrict__ A, int lda, long long M, const float *__restrict__ B, int ldb, long long N, int K,
               float *__restrict__ C, long long ldc, const float *__restrict__ row_term, const float *__restrict__ col_term,
               float add_const, int nt, bool vecA, bool vecB)
{
    extern __shared__ unsigned char gemm_smem_raw[];
    __shared__ uint64_t mma_done[2];
    __shared__ uint32_t tmem_base_slot;
    unsigned char *tiles = gemm_smem_raw + ((1024u - (g_smem_u32(gemm_smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(&tmem_base_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        g_mbar_init(&mma_done[0], 1);
        g_mbar_init(&mma_done[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    const long long m0 = (long long)blockIdx.x * kGemmM, n0 = (long long)blockIdx.y * nt;
    const int nstages = (K + kGemmKc - 1) / kGemmKc;
    // instruction descriptor: D = f32 (bit 4), A = B = tf32 (2 at bits 7 and 10), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(kGemmM >> 4) << 24);

    // chunk c of an operand tile: row c / 8, 16-byte piece c % 8 (8 consecutive threads read one 128-byte run)
    constexpr int A_PER = kGemmM * 8 / kGemmThreads;      // 4
    constexpr int B_PER = kGemmNMax * 8 / kGemmThreads;   // 8
    float4 ca[A_PER], cb[B_PER], na[A_PER], nb[B_PER];
    auto load_stage = [&](int s, float4 (&fa)[A_PER], float4 (&fb)[B_PER]) {
        const int k0 = s * kGemmKc;
#pragma unroll
        for (int i = 0; i < A_PER; i++) {
            const int c = tid + i * kGemmThreads;
            fa[i] = g_load4(A, lda, m0 + (c >> 3), M, k0 + 4 * (c & 7), K, vecA);
        }
#pragma unroll
        for (int i = 0; i < B_PER; i++) {
            const int c = tid + i * kGemmThreads;
            const int r = c >> 3;
            fb[i] = r < nt ? g_load4(B, ldb, n0 + r, N, k0 + 4 * (c & 7), K, vecB) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if (nstages > 0) load_stage(0, ca, cb);
    for (int s = 0; s < nstages; s++) {
        const int stage = s & 1;
        if (s + 1 < nstages) load_stage(s + 1, na, nb);
        if (s >= 2) g_mbar_wait(&mma_done[stage], (uint32_t)(((s - 2) >> 1) & 1));
        unsigned char *a_hi = tiles + (size_t)stage * kStageBytes, *a_lo = a_hi + kAPart, *b_hi = a_lo + kAPart, *b_lo = b_hi + kBPart;
#pragma unroll
        for (int i = 0; i < A_PER; i++) {
            const int c = tid + i * kGemmThreads;
            g_store_split(a_hi, a_lo, c >> 3, c & 7, ca[i]);
        }
#pragma unroll
        for (int i = 0; i < B_PER; i++) {
            const int c = tid + i * kGemmThreads;
            if ((c >> 3) < nt) g_store_split(b_hi, b_lo, c >> 3, c & 7, cb[i]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = g_smem_u32(a_hi), al = g_smem_u32(a_lo), bh = g_smem_u32(b_hi), bl = g_smem_u32(b_lo);
#pragma unroll
            for (int k8 = 0; k8 < kGemmKc / 8; k8++) {
                const uint32_t koff = (uint32_t)k8 * 32u;   // 8 floats along the 128-byte atom row
                const uint32_t d = tmem_base + (uint32_t)((k8 & 1) * kGemmNMax);
                const uint32_t first = (s == 0 && k8 < 2) ? 0u : 1u;
                g_umma(d, g_desc(ah + koff), g_desc(bh + koff), idesc, first);
                g_umma(d, g_desc(ah + koff), g_desc(bl + koff), idesc, 1u);
                g_umma(d, g_desc(al + koff), g_desc(bh + koff), idesc, 1u);
            }
            g_commit(&mma_done[stage]);
        }
#pragma unroll
        for (int i = 0; i < A_PER; i++) ca[i] = na[i];
#pragma unroll
        for (int i = 0; i < B_PER; i++) cb[i] = nb[i];
    }
    if (nstages > 0) {
        const int s = nstages - 1;
        g_mbar_wait(&mma_done[s & 1], (uint32_t)((s >> 1) & 1));
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w (< 4) owns tensor-memory lanes 32w .. 32w+31 = rows of the tile, one row of C per thread
    if (warp < 4) {
        const long long row = m0 + warp * 32 + lane;
        const float rterm = (row < M && row_term) ? row_term[row] : 0.f;
        for (int c0 = 0; c0 < nt; c0 += 32) {   // nt is a multiple of 16; the last step may read 16 unused columns
            uint32_t v0[32], v1[32];
            g_tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v0);
            g_tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kGemmNMax + c0), v1);
            if (row < M) {
                float *dst = C + (size_t)row * (size_t)ldc + n0 + c0;
#pragma unroll
                for (int e = 0; e < 32; e++) {
                    const long long col = n0 + c0 + e;
                    if (c0 + e < nt && col < N) {
                        float x = nstages > 0 ? __uint_as_float(v0[e]) + __uint_as_float(v1[e]) : 0.f;
                        x += rterm;
                        if (col_term) x += col_term[col];
                        dst[e] = x + add_const;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

int launch_gemm_nt_tc(const real_t *A, int lda, long long M, const real_t *B, int ldb, long long N, int K, real_t *C, long long ldc,
                      const real_t *row_term, const real_t *col_term, real_t add_const, cudaStream_t stream)
{
    if (M < 1 || N < 1) return 0;
    if (K < 1) return 3;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem) != cudaSuccess) {
            cudaGetLastError();
            return 1;
        }
        attr_set = true;
    }
    int nt = N >= kGemmNMax ? kGemmNMax : (int)((N + 15) / 16) * 16;
    const long long gx = (M + kGemmM - 1) / kGemmM, gy = (N + nt - 1) / nt;
    if (gx > 2147483647LL || gy > 65535) return 3;
    const bool vecA = (((uintptr_t)A & 15u) == 0) && (lda % 4 == 0), vecB = (((uintptr_t)B & 15u) == 0) && (ldb % 4 == 0);
    gemm_tc_kernel<<<dim3((unsigned)gx, (unsigned)gy), kGemmThreads, kGemmSmem, stream>>>(A, lda, M, B, ldb, N, K, C, ldc, row_term, col_term,
                                                                                       add_const, nt, vecA, vecB);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

#else

int launch_gemm_nt_tc(const real_t *, int, long long, const real_t *, int, long long, int, real_t *, long long, const real_t *,
                      const real_t *, real_t, cudaStream_t)
{
    return 3;
}

#endif

}  // namespace cmfb200
