// The ALS alternation on the device: state that lives in HBM for the duration of a fit (both orientations
// of X, both factor matrices, Gram workspace) and the half-sweep / iteration drivers on top of the kernels
// in sweep.h.  This is the loop body of the reference's fit_collective_explicit_als
// (src/collective.c:8334-8898) and fit_collective_implicit_als (src/collective.c:9827-10040) for models
// without side information.
#pragma once
#include <functional>
#include <utility>
#include <vector>
#include <cuda_runtime.h>
#include "cmf_types.h"
#include "sweep.h"

namespace cmfb200 {

// Device allocations come from the device's default stream-ordered memory pool with its release threshold raised, so
// that the ~20 buffers of a fit call (hundreds of MB at ML10M shape) are recycled by the next call instead of going
// through cudaMalloc / cudaFree every time (measured: 100-500 ms per call, profiles/README.md).  CMFB200_POOL=0 falls
// back to cudaMalloc / cudaFree; cmfb200_trim_pool() hands the cached memory back to the driver.
bool devbuf_pool_enabled();
void devbuf_trim_pool();

template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    bool pooled = false;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) {
            if (pooled) {
                // like cudaFree, wait for whatever may still be using the buffer (any stream), then return it to the pool
                cudaDeviceSynchronize();
                cudaFreeAsync(p, nullptr);
            } else {
                cudaFree(p);
            }
        }
        p = nullptr;
        n = 0;
    }
    void swap(DevBuf &o)
    {
        std::swap(p, o.p);
        std::swap(n, o.n);
        std::swap(pooled, o.pooled);
    }
    bool alloc(size_t count)
    {
        release();
        n = count;
        if (count == 0) return true;
        pooled = devbuf_pool_enabled();
        if (pooled) {
            if (cudaMallocAsync((void **)&p, count * sizeof(T), nullptr) == cudaSuccess) {
                // usable from every stream once the allocation itself has completed
                return cudaStreamSynchronize(nullptr) == cudaSuccess;
            }
            cudaGetLastError();
            pooled = false;
        }
        return cudaMalloc((void **)&p, count * sizeof(T)) == cudaSuccess;
    }
};

// One orientation of X (CSR = rows are users, CSC = rows are items) restricted to this rank's row block,
// with row and column ids already renumbered into the device ("dealt") numbering.
struct DeviceSide {
    int_t rows_padded = 0;        // world * block
    int_t block = 0;              // rows per rank
    int_t row_begin = 0, row_end = 0;
    size_t nnz_local = 0;
    DevBuf<size_t> ptr;
    DevBuf<int_t> idx;
    DevBuf<real_t> val;
    DevBuf<int_t> order;
    std::vector<int_t> deg_sorted;   // host: stored entries of order[i], descending
    int_t n_order = 0, n_long = 0, n_huge = 0;
    // CG sweeps over this side: the most gathered rows of the OPPOSING side kept in shared memory (AlsState::prepare_hot)
    DevBuf<int_t> hot_idx;           // idx with the table slot + 1 packed into bits 20..30
    DevBuf<int_t> hot_rows;          // the opposing rows in the table
    int n_hot = 0;
    CsrView view() const { return CsrView{ptr.p, idx.p, val.p}; }
    SweepPlan plan() const { return SweepPlan{order.p, n_order, n_long, n_huge, deg_sorted.empty() ? nullptr : deg_sorted.data()}; }
};

struct Renumbering {              // old (caller) row id <-> device row id
    std::vector<int_t> to_dev;    // [rows]          host copies: only filled by the host-side dealing of setup()
    std::vector<int_t> to_old;    // [rows_padded], -1 for padding rows
    DevBuf<int_t> d_to_dev;       // [rows] on the device; null = identity numbering (one rank)
    DevBuf<int_t> d_to_old;       // [rows_padded]
    int_t block = 0, rows_padded = 0;
};

// starting biases computed on the device while X is ingested (AlsState::setup_from_coo)
struct BiasInit {
    int which = 0;                // 0 none, 3 both sides (two-sided sweeps), 1 users only, 2 items only
    real_t lam_user = 0, lam_item = 0;
    bool scale_lam = false;
};

struct AlsConfig {
    bool implicit = false;
    int_t m = 0, n = 0, kk = 0;   // kk = k + k_main
    bool user_bias = false, item_bias = false;
    real_t lam_A = 0, lam_B = 0;          // regulariser of the factor columns
    real_t lam_biasA = 0, lam_biasB = 0;  // regulariser of the bias coordinate
    bool scale_lam = false;
    bool scale_bias_const = false;        // the bias regulariser is NOT multiplied by the row's entry count (fold-in of new rows)
    bool last_coord_special = false;      // see CgSweepParams::last_coord_special (fold-in of new rows without a user bias)
    int max_cg_steps = 3;
    int rank = 0, world = 1;
};

// deal rows to `world` ranks (round-robin over decreasing degree); identity when world == 1
void build_renumbering(const size_t *ptr, int_t rows, int world, Renumbering &ren);

class NcclLink;
class CollectiveState;

// Which rank of how many the reference-named fit entry points run as (cmfb200_set_world); one process drives one GPU.
struct WorldSetting {
    int rank = 0, world = 1;
    unsigned char nccl_id[128] = {0};
};
WorldSetting &world_setting();
// set asynchronously (SIGINT handler installed by the fit entry points); polled between half-sweeps
volatile int &stop_flag();

class AlsState {
public:
    AlsConfig cfg;
    cudaStream_t stream = nullptr;
    cudaStream_t side_stream = nullptr;   // long-row kernel runs here, fenced by the two events below
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    Renumbering renA, renB;
    DeviceSide byA, byB;          // byA: rows = users (CSR); byB: rows = items (CSC)
    int ldA = 0, ldB = 0;
    DevBuf<real_t> A, B;          // [rows_padded x ld], device numbering
    DevBuf<real_t> biasA, biasB;  // [rows_padded], device numbering (zero when the side has no bias)
    DevBuf<real_t> gram, gram_ws;
    NcclLink *link = nullptr;
    CollectiveState *coll = nullptr;   // attached side information / implicit features (collective.cu); owned
    // explicit model with side information / implicit features (collective.cu): constant matrix and per-row vector
    // added to the systems of the rows of B ([0]) / of A ([1]); rows without entries are solved too when set
    const real_t *extraQ[2] = {nullptr, nullptr};
    const real_t *extraq[2] = {nullptr, nullptr};
    int extra_ldq[2] = {0, 0};
    bool extra_all_rows[2] = {false, false};
    bool values_positive = false;   // implicit model: every stored value is > 0 (checked once when the state is set up)
    bool verbose = false;         // progress lines on stdout like the reference's ("Updating B ... done")
    bool use_nm_cg = false;       // CMFB200_NMCG=1: run the explicit model's CG on the tensor-core-built normal matrix (sweep_nm.cu)
    bool use_resident = true;     // CMFB200_RESIDENT=0 selects the direct-gather CG kernel (read when the state is set up)
    long long launches = 0;       // kernels launched so far (for bench.py's gpu_launches)
    // optional per-launch timing of the row-solve kernel (CUDA events on `stream`, resolved on demand)
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> sweep_events[2];   // [which]
    int read_profile(int which, double *total_ms, long long *count);    // synchronises the stream

    ~AlsState();
    // host CSR/CSC in caller numbering (values already centred / scaled as the model requires)
    int setup(const AlsConfig &c, const size_t *csr_p, const int_t *csr_i, const real_t *csr_v, const size_t *csc_p,
              const int_t *csc_i, const real_t *csc_v, cudaStream_t s, const void *nccl_id);
    // single-GPU ingestion straight from host COO triplets: upload, subtract `mu`, multiply by `scale`, build both
    // orientations on the device (device_prep.cu).  Values are transformed as  (x - mu) * scale  in real_t.
    // With cfg.world > 1 (nccl_id: the communicator's unique id) every rank passes the SAME triplets: both orientations are
    // built in full on every device, the rows are dealt to the ranks there (decreasing degree, round-robin) and only this
    // rank's blocks are kept.  coo_on_device: ixA / ixB / X are device pointers already.  bias: starting biases computed
    // on the device from the full matrices (before the dealing) and stored in device numbering.
    int prepare_hot();   // after both sides are planned (one GPU): see DeviceSide::hot_idx
    int setup_from_coo(const AlsConfig &c, const int_t *ixA, const int_t *ixB, const real_t *X, size_t nnz, real_t mu,
                       real_t scale, cudaStream_t s, const std::function<real_t()> *mu_later = nullptr,   // mu_later: the mean is still being computed on the host; asked for once the uploads are in flight
                       const void *nccl_id = nullptr, const BiasInit *bias = nullptr, bool coo_on_device = false);
    // factors in caller numbering: A [m x kk] (ld = lda), biasA [m] or null; same for B
    int upload_factors(const real_t *hA, int lda, const real_t *hbiasA, const real_t *hB, int ldb, const real_t *hbiasB);
    int upload_coordinates(const real_t *hA, const real_t *hB);   // keeps the bias slots already on the device
    int upload_bias(int which, const real_t *hbias);
    int random_factors(unsigned long long seed, real_t scale);   // A ~ U(0, scale) from a counter hash, B = 0, biases = 0 (device)
    int download_factors(real_t *hA, int lda, real_t *hbiasA, real_t *hB, int ldb, real_t *hbiasB);
    // any device matrix with one row per user (which = 1) / item (0) in device numbering -> host, caller numbering, kk columns
    int download_matrix(int which, const real_t *src, int ld, real_t *h, int ldh);
    // which: 0 = update B (items) from A, 1 = update A (users) from B.  `solver`: 0 = CG, 1 = Cholesky
    int half_sweep(int which, int iter, int solver);
    int exchange(int which);      // all-gather the freshly solved block (no-op on one GPU)
    int iterate(int first_iter, int n_iters, int niter_total, bool use_cg, bool finalize_chol);
};

}  // namespace cmfb200
