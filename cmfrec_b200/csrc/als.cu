// See als.h.  Host-side orchestration only; all arithmetic on factor rows happens in the kernels.
#include "als.h"
#include "nccl_link.h"
#include "device_prep.h"
#include "collective.h"
#include "dense_small.h"
#include <chrono>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace cmfb200 {

static int env_or(const char *name, int dflt)
{
    const char *e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

WorldSetting &world_setting()
{
    static WorldSetting w;
    return w;
}

volatile int &stop_flag()
{
    static volatile int flag = 0;
    return flag;
}

bool devbuf_pool_enabled()
{
    static const bool on = [] {
        if (env_or("CMFB200_POOL", 1) == 0) return false;
        int dev = 0, supported = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return false;
        cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev);
        if (!supported) return false;
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) { cudaGetLastError(); return false; }
        unsigned long long keep = ~0ull;   // never hand freed memory back on its own
        if (cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess) { cudaGetLastError(); return false; }
        return true;
    }();
    return on;
}

void devbuf_trim_pool()
{
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        cudaDeviceSynchronize();
        cudaMemPoolTrimTo(pool, 0);
    }
    cudaGetLastError();
}

static int long_row_threshold()
{
    static int v = -1;
    if (v < 0) {
        const char *e = std::getenv("CMFB200_LONG_ROW");
        v = e ? std::atoi(e) : 1024;
        if (v < 64) v = 64;
    }
    return v;
}

// Rows are dealt to ranks round-robin in order of decreasing degree, so that every rank gets the same
// number of rows (equal-sized all-gather blocks) and nearly the same number of stored entries.
// With one rank the numbering is left untouched.
void build_renumbering(const size_t *ptr, int_t rows, int world, Renumbering &ren)
{
    ren.block = (rows + world - 1) / world;
    ren.rows_padded = ren.block * world;
    ren.to_dev.resize(rows);
    ren.to_old.assign(ren.rows_padded, -1);
    if (world == 1) {
        std::iota(ren.to_dev.begin(), ren.to_dev.end(), 0);
        std::iota(ren.to_old.begin(), ren.to_old.begin() + rows, 0);
        return;
    }
    std::vector<int_t> by_degree(rows);
    std::iota(by_degree.begin(), by_degree.end(), 0);
    std::stable_sort(by_degree.begin(), by_degree.end(), [&](int_t a, int_t b) {
        return (ptr[a + 1] - ptr[a]) > (ptr[b + 1] - ptr[b]);
    });
    for (int_t s = 0; s < rows; s++) {
        const int_t dev = (s % world) * ren.block + s / world;
        ren.to_dev[by_degree[s]] = dev;
        ren.to_old[dev] = by_degree[s];
    }
}

static int build_side(const size_t *ptr, const int_t *idx, const real_t *val, const Renumbering &rows_ren,
                      const Renumbering &cols_ren, int rank, cudaStream_t stream, DeviceSide &side)
{
    side.rows_padded = rows_ren.rows_padded;
    side.block = rows_ren.block;
    side.row_begin = rank * rows_ren.block;
    side.row_end = side.row_begin + rows_ren.block;
    std::vector<size_t> hptr((size_t)side.rows_padded + 1, 0);
    size_t total = 0;
    for (int_t r = side.row_begin; r < side.row_end; r++) {
        const int_t old = rows_ren.to_old[r];
        if (old >= 0) total += ptr[old + 1] - ptr[old];
    }
    std::vector<int_t> hidx(total);
    std::vector<real_t> hval(total);
    size_t at = 0;
    for (int_t r = 0; r < side.rows_padded; r++) {
        hptr[r] = at;
        if (r < side.row_begin || r >= side.row_end) continue;
        const int_t old = rows_ren.to_old[r];
        if (old < 0) continue;
        for (size_t e = ptr[old]; e < ptr[old + 1]; e++) {
            hidx[at] = cols_ren.to_dev[idx[e]];
            hval[at] = val[e];
            at++;
        }
    }
    hptr[side.rows_padded] = at;
    side.nnz_local = total;

    // processing order: local rows by decreasing degree; rows longer than the threshold get a whole block
    std::vector<int_t> order;
    order.reserve(side.block);
    for (int_t r = side.row_begin; r < side.row_end; r++)
        if (rows_ren.to_old[r] >= 0) order.push_back(r);
    std::stable_sort(order.begin(), order.end(), [&](int_t a, int_t b) {
        return (hptr[a + 1] - hptr[a]) > (hptr[b + 1] - hptr[b]);
    });
    const size_t thr = (size_t)long_row_threshold();
    int_t n_long = 0;
    while (n_long < (int_t)order.size() && hptr[order[n_long] + 1] - hptr[order[n_long]] >= thr) n_long++;
    side.n_order = (int_t)order.size();
    side.n_long = n_long;
    side.deg_sorted.resize(order.size());
    for (size_t i = 0; i < order.size(); i++) side.deg_sorted[i] = (int_t)(hptr[order[i] + 1] - hptr[order[i]]);
    {
        const size_t t_huge = (size_t)env_or("CMFB200_HUGE_ROW", 8192);
        int_t h = 0;
        while (h < n_long && hptr[order[h] + 1] - hptr[order[h]] >= t_huge) h++;
        side.n_huge = h;
    }

    if (!side.ptr.alloc(hptr.size()) || !side.idx.alloc(std::max<size_t>(total, 1)) ||
        !side.val.alloc(std::max<size_t>(total, 1)) || !side.order.alloc(std::max<size_t>(order.size(), 1)))
        return 1;
    if (cudaMemcpyAsync(side.ptr.p, hptr.data(), hptr.size() * sizeof(size_t), cudaMemcpyHostToDevice, stream) != cudaSuccess) return 1;
    if (total) {
        if (cudaMemcpyAsync(side.idx.p, hidx.data(), total * sizeof(int_t), cudaMemcpyHostToDevice, stream) != cudaSuccess) return 1;
        if (cudaMemcpyAsync(side.val.p, hval.data(), total * sizeof(real_t), cudaMemcpyHostToDevice, stream) != cudaSuccess) return 1;
    }
    if (!order.empty())
        if (cudaMemcpyAsync(side.order.p, order.data(), order.size() * sizeof(int_t), cudaMemcpyHostToDevice, stream) != cudaSuccess) return 1;
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : 1;
}

AlsState::~AlsState()
{
    for (auto &v : sweep_events)
        for (auto &pr : v) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    delete coll;
    delete link;
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (side_stream) cudaStreamDestroy(side_stream);
}

int AlsState::read_profile(int which, double *total_ms, long long *count)
{
    if (cudaStreamSynchronize(stream) != cudaSuccess) return 1;
    auto &v = sweep_events[which ? 1 : 0];
    double tot = 0;
    for (auto &pr : v) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, pr.first, pr.second);
        tot += ms;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    if (total_ms) *total_ms = tot;
    if (count) *count = (long long)v.size();
    v.clear();
    return 0;
}

int AlsState::setup(const AlsConfig &c, const size_t *csr_p, const int_t *csr_i, const real_t *csr_v,
                    const size_t *csc_p, const int_t *csc_i, const real_t *csc_v, cudaStream_t s, const void *nccl_id)
{
    cfg = c;
    stream = s;
    use_resident = env_or("CMFB200_RESIDENT", 1) != 0;   // read once per state, not per half-sweep
    use_nm_cg = env_or("CMFB200_NMCG", 0) != 0;
    if (cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking) != cudaSuccess) return 1;
    cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming);
    if (cfg.kk < 1 || cfg.kk > max_supported_k()) return 2;
    if (cfg.world > 1) {
        if (!nccl_id) return 2;
        link = new NcclLink();
        if (link->init(nccl_id, cfg.rank, cfg.world) != 0) return 1;
    }
    build_renumbering(csr_p, cfg.m, cfg.world, renA);
    build_renumbering(csc_p, cfg.n, cfg.world, renB);
    if (cfg.world > 1) {
        // the factor uploads / downloads renumber rows on the device
        if (!renA.d_to_dev.alloc(cfg.m) || !renB.d_to_dev.alloc(cfg.n)) return 1;
        if (cudaMemcpyAsync(renA.d_to_dev.p, renA.to_dev.data(), (size_t)cfg.m * sizeof(int_t), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
            cudaMemcpyAsync(renB.d_to_dev.p, renB.to_dev.data(), (size_t)cfg.n * sizeof(int_t), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
            cudaStreamSynchronize(stream) != cudaSuccess)
            return 1;
    }
    int rc = build_side(csr_p, csr_i, csr_v, renA, renB, cfg.rank, stream, byA);
    if (rc) return rc;
    rc = build_side(csc_p, csc_i, csc_v, renB, renA, cfg.rank, stream, byB);
    if (rc) return rc;
    if ((rc = prepare_hot())) return rc;
    ldA = cmf_ld_for(cfg.kk);
    ldB = cmf_ld_for(cfg.kk);
    if (!A.alloc((size_t)renA.rows_padded * ldA) || !B.alloc((size_t)renB.rows_padded * ldB) ||
        !biasA.alloc(renA.rows_padded) || !biasB.alloc(renB.rows_padded))
        return 1;
    if (cudaMemsetAsync(A.p, 0, A.n * sizeof(real_t), stream) != cudaSuccess) return 1;
    if (cudaMemsetAsync(B.p, 0, B.n * sizeof(real_t), stream) != cudaSuccess) return 1;
    if (cudaMemsetAsync(biasA.p, 0, biasA.n * sizeof(real_t), stream) != cudaSuccess) return 1;
    if (cudaMemsetAsync(biasB.p, 0, biasB.n * sizeof(real_t), stream) != cudaSuccess) return 1;
    if (cfg.implicit) {
        if (!gram.alloc((size_t)cfg.kk * cfg.kk) || !gram_ws.alloc(gram_workspace_elems(cfg.kk))) return 1;
        if (device_all_positive(byA.val.p, byA.nnz_local, &values_positive, stream)) return 1;
    }
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : 1;
}

// ---- hot opposing rows (sweep.h: CgSweepParams::hot_idx)
namespace {
__global__ void hot_slot_kernel(const int_t *__restrict__ hot_rows, int n_hot, int_t *__restrict__ slot1)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_hot) slot1[hot_rows[s]] = s + 1;
}
__global__ void hot_pack_kernel(const int_t *__restrict__ idx, size_t nnz, const int_t *__restrict__ slot1, int_t *__restrict__ out)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nnz) {
        const int_t col = idx[e];
        out[e] = col | (slot1[col] << 20);
    }
}
}  // namespace

// The rows of one side that most stored entries of the other side point at (its first rows in degree order) are worth a
// shared-memory table in the CG sweep over the other side when they cover a good share of all entries (item popularity is
// heavy-tailed).  One GPU, fp32, 32 < k <= 128.  Opt-in (CMFB200_RES_HOT=1 on a build with -DCMF_RES_HOT_ENABLE): at ML10M
// shape the 256 most popular items hold 19 % of the entries and the table loses (sweep_cg_resident.cu: kHotCompiled).
static int prepare_hot_side(DeviceSide &side, const DeviceSide &opp, int_t opp_rows, int kk, cudaStream_t stream)
{
    side.n_hot = 0;
    if (sizeof(real_t) != 4 || kk <= 32 || kk > 128 || env_or("CMFB200_RES_HOT", 0) == 0) return 0;   // opt-in, see sweep_cg_resident.cu
    if (opp_rows >= (1 << 20) || opp.deg_sorted.empty() || side.nnz_local == 0) return 0;
    int n_hot = env_or("CMFB200_RES_HOT_ROWS", kk <= 64 ? 256 : 128);
    n_hot = std::min<long long>(std::min(n_hot, 2047), (long long)opp.deg_sorted.size());
    if (n_hot < 1) return 0;
    size_t covered = 0;
    for (int i = 0; i < n_hot; i++) covered += (size_t)opp.deg_sorted[(size_t)i];
    if (covered * 100 < side.nnz_local * (size_t)env_or("CMFB200_RES_HOT_MINPCT", 25)) return 0;
    DevBuf<int_t> slot1;
    if (!slot1.alloc((size_t)opp_rows) || !side.hot_rows.alloc((size_t)n_hot) || !side.hot_idx.alloc(side.nnz_local)) return 1;
    if (cudaMemsetAsync(slot1.p, 0, (size_t)opp_rows * sizeof(int_t), stream) != cudaSuccess ||
        cudaMemcpyAsync(side.hot_rows.p, opp.order.p, (size_t)n_hot * sizeof(int_t), cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
        return 1;
    hot_slot_kernel<<<(n_hot + 255) / 256, 256, 0, stream>>>(side.hot_rows.p, n_hot, slot1.p);
    hot_pack_kernel<<<(unsigned)((side.nnz_local + 255) / 256), 256, 0, stream>>>(side.idx.p, side.nnz_local, slot1.p, side.hot_idx.p);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) return 1;   // slot1 is released on return
    side.n_hot = n_hot;
    return 0;
}

int AlsState::prepare_hot()
{
    if (cfg.world > 1) return 0;   // the dealt numbering of several ranks is not handled yet
    if (int rc = prepare_hot_side(byA, byB, cfg.n, cfg.kk, stream)) return rc;
    return prepare_hot_side(byB, byA, cfg.m, cfg.kk, stream);
}

// bucket boundaries of a side from its (descending) degree list
static void plan_buckets(DeviceSide &side)
{
    const std::vector<int_t> &deg = side.deg_sorted;
    const int_t rows = (int_t)deg.size();
    auto count_ge = [&](size_t thr, int_t limit) {
        int_t c = 0;
        while (c < limit && (size_t)deg[c] >= thr) c++;
        return c;
    };
    side.n_order = rows;
    side.n_long = count_ge((size_t)long_row_threshold(), rows);
    side.n_huge = count_ge((size_t)env_or("CMFB200_HUGE_ROW", 8192), side.n_long);
}

// processing order / bucket boundaries of a side whose ptr array is already on the device (single GPU, identity numbering)
static int plan_side_from_device(DeviceSide &side, int_t rows, cudaStream_t stream)
{
    side.rows_padded = rows;
    side.block = rows;
    side.row_begin = 0;
    side.row_end = rows;
    if (!side.order.alloc(std::max<size_t>((size_t)rows, 1))) return 1;
    // the degree sort runs on the device (stable radix sort: ties keep increasing row order, like std::stable_sort);
    // only the sorted counts come back, for the bucket boundaries
    size_t total = 0;
    if (int rc = device_degree_order(side.ptr.p, rows, side.order.p, side.deg_sorted, &total, stream)) return rc;
    side.nnz_local = total;
    plan_buckets(side);
    return 0;
}

template <typename T> __global__ void scale_kernel(T *x, size_t n, T s)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = x[i] * s;
}
__global__ void iota_rows_kernel(int_t *x, int_t n, int_t first)
{
    const int_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = first + i;
}

// one side of the multi-GPU dealing: full orientation -> this rank's block (see setup_from_coo)
static int deal_side(DeviceSide &full, int_t rows, int rank, int world, Renumbering &ren, cudaStream_t stream,
                     DevBuf<int_t> &order_full, std::vector<int_t> &deg_full)
{
    ren.block = (rows + world - 1) / world;
    ren.rows_padded = ren.block * world;
    if (!order_full.alloc(std::max<size_t>((size_t)rows, 1)) || !ren.d_to_dev.alloc(std::max<size_t>((size_t)rows, 1)) ||
        !ren.d_to_old.alloc((size_t)ren.rows_padded))
        return 1;
    size_t total = 0;
    if (int rc = device_degree_order(full.ptr.p, rows, order_full.p, deg_full, &total, stream)) return rc;
    (void)rank;
    return device_deal_rows(order_full.p, rows, world, ren.block, ren.d_to_dev.p, ren.d_to_old.p, stream);
}

int AlsState::setup_from_coo(const AlsConfig &c, const int_t *ixA, const int_t *ixB, const real_t *X, size_t nnz, real_t mu,
                             real_t scale, cudaStream_t s, const std::function<real_t()> *mu_later, const void *nccl_id,
                             const BiasInit *bias, bool coo_on_device)
{
    cfg = c;
    stream = s;
    use_resident = env_or("CMFB200_RESIDENT", 1) != 0;
    use_nm_cg = env_or("CMFB200_NMCG", 0) != 0;
    if (cfg.world < 1) cfg.world = 1;
    if (cfg.kk < 1 || cfg.kk > max_supported_k()) return 2;
    if (cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking) != cudaSuccess) return 1;
    cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming);
    if (cfg.world > 1) {
        if (!nccl_id) return 2;
        link = new NcclLink();
        if (link->init(nccl_id, cfg.rank, cfg.world) != 0) return 1;
    }
    // CMFB200_TIMING=1: wall-clock of every ingestion stage to stderr (synchronises after each; a measurement aid)
    const bool timing = std::getenv("CMFB200_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (!timing) return;
        cudaStreamSynchronize(stream);
        const auto t = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[cmfb200 timing]   ingest: %-34s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    // ---- the triplets on the device, centred / scaled
    DevBuf<int_t> dA, dB;
    DevBuf<real_t> dX;
    const int_t *pA = ixA, *pB = ixB;
    real_t *pX = const_cast<real_t *>(X);
    const size_t cap = std::max<size_t>(nnz, 1);
    if (!coo_on_device) {
        if (!dA.alloc(cap) || !dB.alloc(cap) || !dX.alloc(cap)) return 1;
        if (nnz) {
            if (cudaMemcpyAsync(dA.p, ixA, nnz * sizeof(int_t), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
                cudaMemcpyAsync(dB.p, ixB, nnz * sizeof(int_t), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
                cudaMemcpyAsync(dX.p, X, nnz * sizeof(real_t), cudaMemcpyHostToDevice, stream) != cudaSuccess)
                return 1;
        }
        pA = dA.p; pB = dB.p; pX = dX.p;
    }
    mark("allocate + upload COO");
    if (nnz) {
        if (mu_later) mu = (*mu_later)();   // computed on a host thread while the copies above were in flight
        if (mu != 0 && device_subtract(pX, nnz, mu, stream)) return 1;
        if (scale != 1) scale_kernel<real_t><<<(unsigned)((nnz + 255) / 256), 256, 0, stream>>>(pX, nnz, scale);
    }
    mark("wait for the mean, centre / scale");
    // ---- both orientations in full (caller numbering)
    if (!byA.ptr.alloc((size_t)cfg.m + 1) || !byA.idx.alloc(cap) || !byA.val.alloc(cap) || !byB.ptr.alloc((size_t)cfg.n + 1) ||
        !byB.idx.alloc(cap) || !byB.val.alloc(cap))
        return 1;
    int rc = device_compress(pA, pB, pX, nnz, cfg.m, byA.ptr.p, byA.idx.p, byA.val.p, stream);
    if (rc) return rc;
    rc = device_compress(pB, pA, pX, nnz, cfg.n, byB.ptr.p, byB.idx.p, byB.val.p, stream);
    if (rc) return rc;
    launches += 12;
    dA.release(); dB.release(); dX.release();
    mark("CSR + CSC");
    // ---- starting biases from the full matrices (caller numbering)
    DevBuf<real_t> bias_fullA, bias_fullB;
    if (!bias_fullA.alloc((size_t)cfg.m) || !bias_fullB.alloc((size_t)cfg.n)) return 1;
    if (cudaMemsetAsync(bias_fullA.p, 0, (size_t)cfg.m * sizeof(real_t), stream) != cudaSuccess ||
        cudaMemsetAsync(bias_fullB.p, 0, (size_t)cfg.n * sizeof(real_t), stream) != cudaSuccess)
        return 1;
    if (bias && bias->which) {
        if (bias->which == 3) {
            rc = device_init_biases_twosided(cfg.m, cfg.n, byA.ptr.p, byA.idx.p, byA.val.p, byB.ptr.p, byB.idx.p, byB.val.p,
                                             bias->lam_user, bias->lam_item, bias->scale_lam, bias_fullA.p, bias_fullB.p, stream);
            launches += 10;
        } else if (bias->which == 1) {
            rc = device_init_biases_onesided(cfg.m, byA.ptr.p, byA.val.p, bias->lam_user, bias->scale_lam, bias_fullA.p, stream);
            launches += 1;
        } else if (bias->which == 2) {
            rc = device_init_biases_onesided(cfg.n, byB.ptr.p, byB.val.p, bias->lam_item, bias->scale_lam, bias_fullB.p, stream);
            launches += 1;
        }
        if (rc) return rc;
    }
    mark("starting biases");
    ldA = cmf_ld_for(cfg.kk);
    ldB = cmf_ld_for(cfg.kk);
    if (cfg.world == 1) {
        build_renumbering(nullptr, cfg.m, 1, renA);
        build_renumbering(nullptr, cfg.n, 1, renB);
        if ((rc = plan_side_from_device(byA, cfg.m, stream))) return rc;
        if ((rc = plan_side_from_device(byB, cfg.n, stream))) return rc;
        if ((rc = prepare_hot())) return rc;
    } else {
        // ---- deal the rows to the ranks and keep this rank's blocks
        DeviceSide fullA, fullB;
        fullA.ptr.swap(byA.ptr); fullA.idx.swap(byA.idx); fullA.val.swap(byA.val);
        fullB.ptr.swap(byB.ptr); fullB.idx.swap(byB.idx); fullB.val.swap(byB.val);
        DevBuf<int_t> orderA, orderB;
        std::vector<int_t> degA, degB;
        if ((rc = deal_side(fullA, cfg.m, cfg.rank, cfg.world, renA, stream, orderA, degA))) return rc;
        if ((rc = deal_side(fullB, cfg.n, cfg.rank, cfg.world, renB, stream, orderB, degB))) return rc;
        auto extract = [&](DeviceSide &full, const std::vector<int_t> &deg, int_t rows, Renumbering &ren, Renumbering &other,
                           DeviceSide &side) -> int {
            side.rows_padded = ren.rows_padded;
            side.block = ren.block;
            side.row_begin = cfg.rank * ren.block;
            side.row_end = side.row_begin + ren.block;
            // sorted positions rank, rank + world, ... are this rank's rows, already in decreasing-degree order
            const int_t n_local = rows > cfg.rank ? (rows - cfg.rank + cfg.world - 1) / cfg.world : 0;
            side.deg_sorted.resize((size_t)n_local);
            for (int_t i = 0; i < n_local; i++) side.deg_sorted[i] = deg[(size_t)cfg.rank + (size_t)i * cfg.world];
            plan_buckets(side);
            if (!side.order.alloc(std::max<size_t>((size_t)n_local, 1))) return 1;
            if (n_local > 0) iota_rows_kernel<<<(n_local + 255) / 256, 256, 0, stream>>>(side.order.p, n_local, side.row_begin);
            return device_extract_block(full.ptr.p, full.idx.p, full.val.p, ren.d_to_old.p, other.d_to_dev.p, side.row_begin,
                                        side.row_end, n_local, ren.rows_padded, side.ptr, side.idx, side.val, &side.nnz_local, stream);
        };
        if ((rc = extract(fullA, degA, cfg.m, renA, renB, byA))) return rc;
        if ((rc = extract(fullB, degB, cfg.n, renB, renA, byB))) return rc;
        launches += 10;
        if (cudaStreamSynchronize(stream) != cudaSuccess) return 1;   // the full matrices are released on leaving this scope
    }
    mark("row plan / dealing");
    if (!A.alloc((size_t)renA.rows_padded * ldA) || !B.alloc((size_t)renB.rows_padded * ldB) ||
        !biasA.alloc(renA.rows_padded) || !biasB.alloc(renB.rows_padded))
        return 1;
    if (cudaMemsetAsync(A.p, 0, A.n * sizeof(real_t), stream) != cudaSuccess ||
        cudaMemsetAsync(B.p, 0, B.n * sizeof(real_t), stream) != cudaSuccess ||
        cudaMemsetAsync(biasA.p, 0, biasA.n * sizeof(real_t), stream) != cudaSuccess ||
        cudaMemsetAsync(biasB.p, 0, biasB.n * sizeof(real_t), stream) != cudaSuccess)
        return 1;
    if ((rc = device_scatter_rows(bias_fullA.p, 1, cfg.m, 1, renA.d_to_dev.p, biasA.p, 1, stream))) return rc;
    if ((rc = device_scatter_rows(bias_fullB.p, 1, cfg.n, 1, renB.d_to_dev.p, biasB.p, 1, stream))) return rc;
    if (cfg.implicit) {
        if (!gram.alloc((size_t)cfg.kk * cfg.kk) || !gram_ws.alloc(gram_workspace_elems(cfg.kk))) return 1;
        if (device_all_positive(byA.val.p, byA.nnz_local, &values_positive, stream)) return 1;
        if (link) {
            // every rank must take the same kernel: positive everywhere or nowhere
            int flag = values_positive ? 0 : 1;
            DevBuf<real_t> f;
            if (!f.alloc(1)) return 1;
            const real_t hv = (real_t)flag;
            cudaMemcpyAsync(f.p, &hv, sizeof(real_t), cudaMemcpyHostToDevice, stream);
            if (link->all_reduce_sum(f.p, 1, sizeof(real_t) == 8, stream)) return 1;
            real_t out = 0;
            cudaMemcpyAsync(&out, f.p, sizeof(real_t), cudaMemcpyDeviceToHost, stream);
            if (cudaStreamSynchronize(stream) != cudaSuccess) return 1;
            values_positive = out == real_t(0);
        }
    }
    mark("factor buffers");
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : 1;
}

// rows of a host matrix [rows x kk] (row stride ldh, caller numbering) -> device rows (stride ld, device numbering)
static int upload_rows(const real_t *h, int ldh, int_t rows, int kk, const Renumbering &ren, real_t *dst, int ld, cudaStream_t stream)
{
    const size_t w = (size_t)kk * sizeof(real_t);
    if (!ren.d_to_dev.p)
        return cudaMemcpy2DAsync(dst, (size_t)ld * sizeof(real_t), h, (size_t)ldh * sizeof(real_t), w, rows, cudaMemcpyHostToDevice,
                                 stream) == cudaSuccess ? 0 : 1;
    DevBuf<real_t> tmp;
    if (!tmp.alloc((size_t)rows * kk)) return 1;
    if (cudaMemcpy2DAsync(tmp.p, w, h, (size_t)ldh * sizeof(real_t), w, rows, cudaMemcpyHostToDevice, stream) != cudaSuccess) return 1;
    if (device_scatter_rows(tmp.p, kk, rows, kk, ren.d_to_dev.p, dst, ld, stream)) return 1;
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : 1;   // tmp is released on return
}
static int download_rows(const real_t *src, int ld, int_t rows, int kk, const Renumbering &ren, real_t *h, int ldh, cudaStream_t stream)
{
    const size_t w = (size_t)kk * sizeof(real_t);
    if (!ren.d_to_dev.p)
        return cudaMemcpy2DAsync(h, (size_t)ldh * sizeof(real_t), src, (size_t)ld * sizeof(real_t), w, rows, cudaMemcpyDeviceToHost,
                                 stream) == cudaSuccess ? 0 : 1;
    DevBuf<real_t> tmp;
    if (!tmp.alloc((size_t)rows * kk)) return 1;
    if (device_gather_rows_back(src, ld, rows, kk, ren.d_to_dev.p, tmp.p, kk, stream)) return 1;
    if (cudaMemcpy2DAsync(h, (size_t)ldh * sizeof(real_t), tmp.p, w, w, rows, cudaMemcpyDeviceToHost, stream) != cudaSuccess) return 1;
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : 1;
}

int AlsState::upload_factors(const real_t *hA, int lda, const real_t *hbiasA, const real_t *hB, int ldb,
                             const real_t *hbiasB)
{
    // padding columns / rows and absent biases read as zero
    if (cudaMemsetAsync(A.p, 0, A.n * sizeof(real_t), stream) != cudaSuccess ||
        cudaMemsetAsync(B.p, 0, B.n * sizeof(real_t), stream) != cudaSuccess ||
        cudaMemsetAsync(biasA.p, 0, biasA.n * sizeof(real_t), stream) != cudaSuccess ||
        cudaMemsetAsync(biasB.p, 0, biasB.n * sizeof(real_t), stream) != cudaSuccess)
        return 1;
    if (hA && upload_rows(hA, lda, cfg.m, cfg.kk, renA, A.p, ldA, stream)) return 1;
    if (hB && upload_rows(hB, ldb, cfg.n, cfg.kk, renB, B.p, ldB, stream)) return 1;
    if (hbiasA && upload_rows(hbiasA, 1, cfg.m, 1, renA, biasA.p, 1, stream)) return 1;
    if (hbiasB && upload_rows(hbiasB, 1, cfg.n, 1, renB, biasB.p, 1, stream)) return 1;
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : 1;
}

// biases of one side from a host array (caller numbering): which = 1 users, 2 items
int AlsState::upload_bias(int which, const real_t *hbias)
{
    if (!hbias) return 2;
    const int rc = which == 1 ? upload_rows(hbias, 1, cfg.m, 1, renA, biasA.p, 1, stream) : upload_rows(hbias, 1, cfg.n, 1, renB, biasB.p, 1, stream);
    if (rc) return rc;
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : 1;
}

// factor coordinates only (bias slots and padding on the device are left untouched)
int AlsState::upload_coordinates(const real_t *hA, const real_t *hB)
{
    if (hA && upload_rows(hA, cfg.kk, cfg.m, cfg.kk, renA, A.p, ldA, stream)) return 1;
    if (hB && upload_rows(hB, cfg.kk, cfg.n, cfg.kk, renB, B.p, ldB, stream)) return 1;
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : 1;
}

int AlsState::download_factors(real_t *hA, int lda, real_t *hbiasA, real_t *hB, int ldb, real_t *hbiasB)
{
    if (hA && download_rows(A.p, ldA, cfg.m, cfg.kk, renA, hA, lda, stream)) return 1;
    if (hB && download_rows(B.p, ldB, cfg.n, cfg.kk, renB, hB, ldb, stream)) return 1;
    if (hbiasA && download_rows(biasA.p, 1, cfg.m, 1, renA, hbiasA, 1, stream)) return 1;
    if (hbiasB && download_rows(biasB.p, 1, cfg.n, 1, renB, hbiasB, 1, stream)) return 1;
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : 1;
}

namespace {
__global__ void random_fill_kernel(real_t *F, size_t rows, int ld, int kk, unsigned long long seed, real_t scale, const int_t *to_old)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * (size_t)ld) return;
    const int c = (int)(i % ld);
    if (c >= kk || (to_old && to_old[i / ld] < 0)) { F[i] = real_t(0); return; }   // padding columns and padding rows stay zero
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);   // splitmix64 of the element index
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    F[i] = (real_t)(((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0)) * scale;
}
}  // namespace

int AlsState::random_factors(unsigned long long seed, real_t scale)
{
    const size_t total = (size_t)renA.rows_padded * ldA;
    if (total) random_fill_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(A.p, (size_t)renA.rows_padded, ldA, cfg.kk, seed, scale,
                                                                                                  renA.d_to_old.p);
    if (cudaMemsetAsync(B.p, 0, B.n * sizeof(real_t), stream) != cudaSuccess ||
        cudaMemsetAsync(biasA.p, 0, biasA.n * sizeof(real_t), stream) != cudaSuccess ||
        cudaMemsetAsync(biasB.p, 0, biasB.n * sizeof(real_t), stream) != cudaSuccess)
        return 1;
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : 1;
}

int AlsState::download_matrix(int which, const real_t *src, int ld, real_t *h, int ldh)
{
    if (download_rows(src, ld, which ? cfg.m : cfg.n, cfg.kk, which ? renA : renB, h, ldh, stream)) return 1;
    return cudaStreamSynchronize(stream) == cudaSuccess ? 0 : 1;
}

int AlsState::half_sweep(int which, int iter, int solver)
{
    const bool solveA = which == 1;
    CgSweepParams p{};
    p.F = solveA ? A.p : B.p;
    p.ldF = solveA ? ldA : ldB;
    p.G = solveA ? B.p : A.p;
    p.ldG = solveA ? ldB : ldA;
    p.Fbias = solveA ? biasA.p : biasB.p;
    p.Gbias = solveA ? biasB.p : biasA.p;
    p.kk = cfg.kk;
    const DeviceSide &side = solveA ? byA : byB;
    p.X = side.view();
    p.plan = side.plan();
    p.hot_idx = side.n_hot > 0 ? side.hot_idx.p : nullptr;
    p.hot_rows = side.hot_rows.p;
    p.n_hot = side.n_hot;
    p.lam = solveA ? cfg.lam_A : cfg.lam_B;
    p.lam_last = solveA ? cfg.lam_biasA : cfg.lam_biasB;
    p.scale_lam = cfg.scale_lam;
    p.scale_bias_const = cfg.scale_bias_const;
    p.last_coord_special = solveA && cfg.last_coord_special;
    p.max_cg_steps = cfg.max_cg_steps;
    const bool both = cfg.user_bias && cfg.item_bias;
    if (cfg.implicit) {
        p.solve_bias = p.center_opp = p.bias_start_one = false;
        // Gram of the opposing factor: every rank multiplies its own block of rows, the k x k partials are all-reduced
        // (the sum order over the ranks is NCCL's: multi-GPU implicit fits differ from the one-GPU fit by that noise)
        const Renumbering &oren = solveA ? renB : renA;
        const int_t opp_rows = link ? oren.block : oren.rows_padded;
        const real_t *opp = p.G + (link ? (size_t)cfg.rank * oren.block * p.ldG : 0);
        int rc = launch_gram(opp, p.ldG, opp_rows, cfg.kk, gram.p, gram_ws.p, stream);
        launches += 2;
        if (rc) return rc;
        if (link) {
            if ((rc = link->all_reduce_sum(gram.p, (size_t)cfg.kk * cfg.kk, sizeof(real_t) == 8, stream))) return rc;
            launches += 1;
        }
        p.gram = gram.p;
        p.values_positive = values_positive;
        // implicit feedback WITH side information (collective.cu): the constant matrix becomes G^T G + w C^T C, every row
        // gets the vector w C^T u_i and is solved whether or not it has entries (src/collective.c:2905-3269, 1849-2132)
        if (extraQ[which ? 1 : 0]) {
            if ((rc = launch_axpby(cfg.kk * cfg.kk, real_t(1), gram.p, real_t(1), extraQ[which ? 1 : 0], gram.p, stream))) return rc;
            launches += 1;
        }
        p.qvec = extraq[which ? 1 : 0];
        p.ldq = extra_ldq[which ? 1 : 0];
        p.solve_all_rows = extra_all_rows[which ? 1 : 0];
    } else {
        p.solve_bias = solveA ? cfg.user_bias : cfg.item_bias;
        p.center_opp = solveA ? cfg.item_bias : cfg.user_bias;
        // Where the reference keeps a 1.0 in the bias column at the moment a row is (re)started:
        // users: always when both biases are fitted; items: from the second iteration on
        // (src/collective.c:8538-8542, 8728-8732).
        p.bias_start_one = both && (solveA || iter > 0);
        p.gram = extraQ[which ? 1 : 0];
        p.qvec = extraq[which ? 1 : 0];
        p.ldq = extra_ldq[which ? 1 : 0];
        p.solve_all_rows = extra_all_rows[which ? 1 : 0];
    }
    static DevBuf<long long> nm_dbg;
    static const bool nm_dbg_on = env_or("CMFB200_NM_DEBUG", 0) != 0;
    if (nm_dbg_on) {
        if (!nm_dbg.p) nm_dbg.alloc(148 * 32);
        cudaMemsetAsync(nm_dbg.p, 0, 148 * 32 * sizeof(long long), stream);
        p.debug = nm_dbg.p;
    }
    int rc;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (profile) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, stream);
    }
    if (solver == 0) {
        rc = 3;
        // explicit / collective model up to k = 64 (fp32): every row's normal matrix from the tensor cores, the CG run on it
        // (sweep_nm.cu) -- one gather per stored entry instead of one per CG pass
        // (opt-in, CMFB200_NMCG=1: parity-green, but its producer warps do not yet keep up with the cached kernel below)
        if (use_nm_cg && !cfg.implicit) rc = launch_explicit_cg_sweep_nm(p, stream);
        // default: one warp per row (a thread block / a cluster of 8 for long rows), the first entries of every share
        // cached in shared memory, the rest streamed through a pipelined gather (sweep_cg_resident.cu)
        if (rc == 3 && use_resident) {
            int nl = 0;
            rc = cfg.implicit ? launch_implicit_cg_sweep_resident(p, stream, &nl) : launch_explicit_cg_sweep_resident(p, stream, &nl);
            if (rc == 0) launches += nl - 1;
        }
        if (rc == 3) {
            // direct gathers from L2 on every pass (sweep_cg.cu); the longest rows go to a cluster kernel on a second stream
            const bool fork = side_stream && p.plan.n_huge > 0;
            if (fork) {
                cudaEventRecord(ev_fork, stream);
                cudaStreamWaitEvent(side_stream, ev_fork, 0);
                p.side_stream = side_stream;
            }
            rc = cfg.implicit ? launch_implicit_cg_sweep(p, stream) : launch_explicit_cg_sweep(p, stream);
            if (fork) {
                cudaEventRecord(ev_join, side_stream);
                cudaStreamWaitEvent(stream, ev_join, 0);
                launches += 1;
            }
        }
    } else {
        rc = cfg.implicit ? launch_implicit_chol_sweep(p, stream) : launch_explicit_chol_sweep(p, stream);
    }
    if (profile) {
        cudaEventRecord(e1, stream);
        sweep_events[which ? 1 : 0].emplace_back(e0, e1);
    }
    launches += 1;
    if (nm_dbg_on && rc == 0) {
        std::vector<long long> h(148 * 32);
        cudaStreamSynchronize(stream);
        cudaMemcpy(h.data(), nm_dbg.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        for (int blk : {0, 1, 73, 147}) {
            const long long *o = h.data() + (size_t)blk * 32;
            std::fprintf(stderr, "[nm dbg] which=%d blk=%d | loaders total/wait/stages:", which, blk);
            for (int w = 0; w < 4; w++) std::fprintf(stderr, " %lld/%lld/%lld", o[w * 3], o[w * 3 + 1], o[w * 3 + 2]);
            std::fprintf(stderr, " | mma total/accw/fullw: %lld/%lld/%lld | asm total/accw/matw/rows: %lld/%lld/%lld/%lld | solver0 total/wait/n: %lld/%lld/%lld\n",
                         o[12], o[13], o[14], o[16], o[17], o[18], o[19], o[20], o[21], o[22]);
        }
    }
    return rc;
}

int AlsState::exchange(int which)
{
    if (cfg.world <= 1) return 0;
    const bool solveA = which == 1;
    real_t *F = solveA ? A.p : B.p;
    const int ld = solveA ? ldA : ldB;
    const int_t block = solveA ? renA.block : renB.block;
    launches += 1;
    const bool has_bias = !cfg.implicit && (solveA ? cfg.user_bias : cfg.item_bias);
    // the factor block and the bias block travel in one NCCL launch
    int rc = has_bias ? link->group_begin() : 0;
    if (rc) return rc;
    rc = link->all_gather_inplace(F, (size_t)block * ld * sizeof(real_t), stream);
    if (rc == 0 && has_bias) rc = link->all_gather_inplace(solveA ? biasA.p : biasB.p, (size_t)block * sizeof(real_t), stream);
    if (has_bias) {
        const int rc2 = link->group_end();
        if (rc == 0) rc = rc2;
    }
    return rc;
}

int AlsState::iterate(int first_iter, int n_iters, int niter_total, bool use_cg, bool finalize_chol)
{
    // returns 3 when a stop was requested (SIGINT): polled before every half-sweep like the reference does
    // (src/collective.c:8343, 8612, 8800); the state then holds the factors of the last completed half-sweep
    auto say = [&](const char *what) {
        if (verbose) {
            std::printf("%s", what);
            std::fflush(stdout);
        }
    };
    for (int it = first_iter; it < first_iter + n_iters; it++) {
        // the reference switches the last iteration to the exact solver (src/collective.c:8336-8340)
        const bool cg_now = use_cg && !(finalize_chol && it == niter_total - 1);
        const int solver = cg_now ? 0 : 1;
        if (coll) {
            if (int rcc = coll->iteration(it, solver)) return rcc;
        } else {
            if (stop_flag()) return 3;
            say("Updating B ...");
            int rc = half_sweep(0, it, solver);
            if (rc) return rc;
            if ((rc = exchange(0))) return rc;
            if (verbose && cudaStreamSynchronize(stream) != cudaSuccess) return 1;
            say(" done\n");
            if (stop_flag()) return 3;
            say("Updating A ...");
            if ((rc = half_sweep(1, it, solver))) return rc;
            if ((rc = exchange(1))) return rc;
            if (verbose && cudaStreamSynchronize(stream) != cudaSuccess) return 1;
            say(" done\n");
        }
        if (verbose) std::printf("\tCompleted ALS iteration %2d\n\n", it + 1);
    }
    return 0;
}

}  // namespace cmfb200
