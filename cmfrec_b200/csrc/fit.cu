// fit_collective_explicit_als / fit_collective_implicit_als on the GPU, behind the reference's own
// argument lists (reference src/cmfrec.h:1851-1921; drivers src/collective.c:7263-9370, 9375-10207).
//
// What is covered here is the models the BASELINE configurations fit: sparse COO input, missing = unknown,
// no observation weights, no L1 / non-negativity constraints, no side information (U, I), with or without
// user/item biases, centring, scale_lam, CG and/or Cholesky per-row solvers, finalize_chol.
// Argument combinations outside of that are refused loudly (return code 2, message on stderr): there is
// deliberately no CPU fallback.
#include "fit.h"
#include "als.h"
#include "host_prep.h"
#include "postfit.h"
#include "collective.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include <chrono>
#include <csignal>
#include <cstdlib>

namespace cmfb200 {

// CMFB200_TIMING=1 prints a wall-clock breakdown of a fit call to stderr
struct StageTimer {
    bool on;
    std::chrono::steady_clock::time_point t0, born;
    StageTimer() : on(std::getenv("CMFB200_TIMING") != nullptr), t0(std::chrono::steady_clock::now()), born(t0) {}
    ~StageTimer()
    {
        if (on)
            std::fprintf(stderr, "[cmfb200 timing] %-28s %8.2f ms\n", "total inside the call",
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - born).count());
    }
    void lap(const char *what)
    {
        if (!on) return;
        cudaDeviceSynchronize();
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[cmfb200 timing] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// SIGINT between half-sweeps stops the alternation (reference: set_interrup_global_variable / should_stop_procedure,
// src/helpers.c:1493, polled at src/collective.c:8343, 8612, 8800): the handler is installed for the duration of a fit
// and the previous one is put back afterwards.
struct InterruptScope {
    struct sigaction old_action;
    bool installed = false;
    InterruptScope()
    {
        stop_flag() = 0;
        struct sigaction sa;
        std::memset(&sa, 0, sizeof(sa));
        sa.sa_handler = [](int) { stop_flag() = 1; };
        sigemptyset(&sa.sa_mask);
        installed = sigaction(SIGINT, &sa, &old_action) == 0;
    }
    ~InterruptScope()
    {
        if (installed) sigaction(SIGINT, &old_action, nullptr);
    }
};

// host threads for the starting factors: what the caller allows, within this process's share of the cores when several
// ranks drive several GPUs of one box
static int rng_threads(int nthreads)
{
    const int hw = (int)std::thread::hardware_concurrency();
    const int world = std::max(1, world_setting().world);
    const int share = hw > 0 ? std::max(1, hw / world) : nthreads;
    return std::max(1, std::min(nthreads, share));
}

static int refuse(const char *what)
{
    std::fprintf(stderr, "cmfrec_b200: %s is not supported by the GPU path (no CPU fallback).\n", what);
    return 2;
}

static int cuda_ready()
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        std::fprintf(stderr, "cmfrec_b200: no CUDA device available; this library has no CPU path.\n");
        return 1;
    }
    return 0;
}

int fit_explicit(const ExplicitArgs &a)
{
    if (a.k_user && a.U == nullptr && a.nnz_U == 0) return 2;
    if (a.k_item && a.II == nullptr && a.nnz_I == 0) return 2;
    if (a.k_main && a.Xfull == nullptr && a.nnz == 0) return 2;
    if (a.Xfull) return refuse("dense X (Xfull)");
    if (a.weight) return refuse("observation weights");
    if (a.NA_as_zero_X) return refuse("NA_as_zero_X");
    if (a.nnz_U || a.nnz_I) return refuse("sparse side information (U_sp / I_sp)");
    if (a.NA_as_zero_U || a.NA_as_zero_I) return refuse("NA_as_zero_U / NA_as_zero_I");
    if (a.U && a.m_u != a.m) return refuse("side information U with a different number of rows than X");
    if (a.II && a.n_i != a.n) return refuse("side information I with a different number of rows than X has columns");
    if (a.nonneg || a.nonneg_C || a.nonneg_D) return refuse("non-negativity constraints");
    if (a.l1_lam != 0 || a.l1_lam_unique) return refuse("L1 regularisation");
    if (a.precondition_cg) return refuse("precondition_cg");
    if (a.k_user || a.k_item) return refuse("k_user / k_item");
    if (a.scale_lam_sideinfo) return refuse("scale_lam_sideinfo");
    if (a.scale_bias_const && (a.scale_lam || a.scale_lam_sideinfo) && (a.user_bias || a.item_bias))
        return refuse("scale_bias_const");
    const bool collective = a.U || a.II || a.add_implicit_features;
    if (collective && a.k_main) return refuse("k_main together with side information / implicit features");
    if (a.add_implicit_features && (!a.Ai || !a.Bi)) return 2;
    if ((a.U && !a.C) || (a.II && !a.D)) return 2;
    if (!a.reset_values) return refuse("reset_values = false");
    if (a.m < 1 || a.n < 1 || a.k + a.k_main < 1) return 2;
    if (int rc = cuda_ready()) return rc;

    const int_t m = a.m, n = a.n;
    const int kk = a.k + a.k_main;
    const size_t nnz = a.nnz;
    const bool has_bias = a.user_bias || a.item_bias;
    const bool scale_lam = a.scale_lam || a.scale_lam_sideinfo;
    bool use_cg = a.use_cg;
    bool finalize_chol = a.finalize_chol && use_cg;

    // regularisation, with w_main folded away (src/collective.c:7497-7521)
    real_t lam = a.lam;
    real_t lam_u[6];
    const bool has_unique = a.lam_unique != nullptr;
    for (int i = 0; i < 6; i++) lam_u[i] = has_unique ? a.lam_unique[i] : a.lam;
    if (a.w_main != 1) {
        lam /= a.w_main;
        for (int i = 0; i < 6; i++) lam_u[i] /= a.w_main;
    }

    StageTimer tm;
    // starting factors on a host thread while the GPU ingests X: A random, B zero with CG
    // (src/collective.c:8241-8274; reference helpers.c:930 semantics live in host_prep.cpp)
    const size_t sizeA = (size_t)m * kk, sizeB = (size_t)n * kk;
    const bool fill_B = a.II || a.add_implicit_features;   // src/collective.c:8243
    std::thread rng([&]() {
        random_init(a.A, sizeA, fill_B ? a.B : nullptr, fill_B ? sizeB : 0, a.seed, true, rng_threads(a.nthreads));
        if (use_cg && !fill_B) std::memset(a.B, 0, sizeB * sizeof(real_t));
    });
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{rng};

    // centring (src/collective.c:7555-7568 -> src/common.c:3423): mean on the host (parity-critical summation
    // order), subtraction on the device
    real_t glob_mean = 0;
    std::thread mean_thread;
    if (a.center) mean_thread = std::thread([&]() { glob_mean = global_mean(a.X, nnz, a.nthreads); });
    struct MeanJoiner { std::thread &t; ~MeanJoiner() { if (t.joinable()) t.join(); } } mean_joiner{mean_thread};
    const std::function<real_t()> mean_later = [&]() {
        if (mean_thread.joinable()) mean_thread.join();
        return glob_mean;
    };

    AlsConfig cfg;
    cfg.implicit = false;
    cfg.m = m; cfg.n = n; cfg.kk = kk;
    cfg.user_bias = a.user_bias; cfg.item_bias = a.item_bias;
    cfg.lam_A = lam_u[2]; cfg.lam_B = lam_u[3];
    cfg.lam_biasA = a.user_bias ? lam_u[0] : lam_u[2];
    cfg.lam_biasB = a.item_bias ? lam_u[1] : lam_u[3];
    cfg.scale_lam = scale_lam;
    cfg.max_cg_steps = a.max_cg_steps;
    (void)lam;

    // upload COO, centre, build CSR + CSC on the device (src/collective.c:7593 -> src/helpers.c:1375), starting biases
    // (src/collective.c:8164-8226) from the full matrices, then (world > 1) deal the rows to the ranks
    const WorldSetting &ws = world_setting();
    cfg.rank = ws.rank;
    cfg.world = ws.world;
    if (cfg.world > 1 && (a.U || a.II)) return refuse("dense side information on more than one GPU");
    BiasInit bi;
    if (has_bias) {
        if (a.user_bias && a.item_bias) bi.which = 3;
        else if (a.user_bias) bi.which = 1;
        else if (use_cg) bi.which = 2;
        bi.lam_user = lam_u[0];
        bi.lam_item = lam_u[1];
        bi.scale_lam = scale_lam;
    }
    AlsState st;
    int rc = st.setup_from_coo(cfg, a.ixA, a.ixB, a.X, nnz, real_t(0), real_t(1), nullptr, &mean_later,
                               cfg.world > 1 ? ws.nccl_id : nullptr, &bi);
    mean_later();
    if (a.glob_mean) *a.glob_mean = glob_mean;
    if (rc) return rc == 2 ? refuse("this value of k") : rc;
    tm.lap("mean (host thread) + upload COO + CSR/CSC + biases (GPU)");

    rng.join();
    tm.lap("factor initialisation (host)");
    // Cholesky-only fits never initialise B: pass the caller's buffer through like the reference does
    rc = st.upload_coordinates(a.A, (use_cg && !fill_B) ? nullptr : a.B);
    if (rc) return rc;
    if (a.item_bias && !a.user_bias && !use_cg) {
        if ((rc = st.upload_bias(2, a.biasB))) return rc;   // left as the caller passed it (src/collective.c:8184)
    }
    tm.lap("upload factors");
    InterruptScope interrupt_scope;
    st.verbose = a.verbose;
    bool interrupted = false;
    if (a.verbose) { std::printf("Starting ALS optimization routine\n\n"); std::fflush(stdout); }
    if (!collective) {
        rc = st.iterate(0, a.niter, a.niter, use_cg, finalize_chol);
        if (rc == 3) interrupted = true;
        else if (rc) return rc == 2 ? refuse("this solver / k combination") : rc;
    } else {
        // side information: column-centre (U_colmeans / I_colmeans are outputs), then the C, D, Bi, Ai, B, A loop
        real_t w_user = a.w_user, w_item = a.w_item, w_imp = a.w_implicit;
        if (a.w_main != 1) { w_user /= a.w_main; w_item /= a.w_main; w_imp /= a.w_main; }
        std::vector<real_t> Uc, Ic;
        CollectiveConfig cc;
        if (a.U) {
            cc.p = a.p;
            if (a.U_colmeans) { if (center_side_info(a.U, m, a.p, a.U_colmeans, Uc)) return refuse("missing values in U"); }
            else Uc.assign(a.U, a.U + (size_t)m * a.p);
        }
        if (a.II) {
            cc.q = a.q;
            if (a.I_colmeans) { if (center_side_info(a.II, n, a.q, a.I_colmeans, Ic)) return refuse("missing values in I"); }
            else Ic.assign(a.II, a.II + (size_t)n * a.q);
        }
        for (real_t v : Uc) if (std::isnan(v)) return refuse("missing values in U");
        for (real_t v : Ic) if (std::isnan(v)) return refuse("missing values in I");
        cc.implicit_features = a.add_implicit_features;
        cc.w_user = w_user; cc.w_item = w_item; cc.w_implicit = w_imp;
        cc.lam_C = lam_u[4] / w_user * (scale_lam ? (real_t)m : real_t(1));
        cc.lam_D = lam_u[5] / w_item * (scale_lam ? (real_t)n : real_t(1));
        cc.lam_Bi = lam_u[3] / w_imp * (scale_lam ? (real_t)m : real_t(1));
        cc.lam_Ai = lam_u[2] / w_imp * (scale_lam ? (real_t)n : real_t(1));
        CollectiveState cs;
        if ((rc = cs.setup(&st, cc, Uc.data(), Ic.data()))) return rc;
        rc = cs.iterate(a.niter, use_cg, finalize_chol);
        if (rc == 3) interrupted = true;
        else if (rc) return rc == 2 ? refuse("this solver / k combination") : rc;
        if ((rc = cs.download(a.C, a.D, a.Ai, a.Bi))) return rc;
    }
    tm.lap("ALS iterations");
    rc = st.download_factors(a.A, kk, a.user_bias ? a.biasA : nullptr, a.B, kk, a.item_bias ? a.biasB : nullptr);
    if (rc) return rc;
    if (cudaStreamSynchronize(nullptr) != cudaSuccess) return 1;
    tm.lap("download factors");
    if (a.verbose && !interrupted) {
        std::printf(std::isnan((double)a.A[0]) ? "ALS procedure failed\n" : "ALS procedure terminated successfully\n");
        std::fflush(stdout);
    }
    // an interrupted fit still finishes the precomputed matrices when asked to handle the interrupt
    // (src/collective.c:8890-8897), and reports code 3 either way
    if (a.precompute_for_predictions && (!interrupted || a.handle_interrupt)) {
        PostfitExplicit pf;
        pf.B = a.B; pf.biasB = a.item_bias ? a.biasB : nullptr; pf.n = n; pf.kk = kk;
        pf.user_bias = a.user_bias; pf.item_bias = a.item_bias;
        pf.lam = lam_u[2]; pf.lam_bias = lam_u[0]; pf.scale_lam = scale_lam;
        pf.B_plus_bias = a.B_plus_bias; pf.BtB = a.precomputedBtB; pf.TransBtBinvBt = a.precomputedTransBtBinvBt;
        if (collective) {
            real_t w_user = a.w_user, w_imp = a.w_implicit;
            if (a.w_main != 1) { w_user /= a.w_main; w_imp /= a.w_main; }
            pf.C = a.U ? a.C : nullptr; pf.p = a.U ? a.p : 0; pf.w_user = w_user;
            pf.Bi = a.add_implicit_features ? a.Bi : nullptr; pf.implicit_features = a.add_implicit_features; pf.w_implicit = w_imp;
            pf.scale_lam_sideinfo = a.scale_lam_sideinfo;
            pf.BiTBi = a.precomputedBiTBi; pf.TransCtCinvCt = a.precomputedTransCtCinvCt; pf.CtCw = a.precomputedCtCw;
            pf.BeTBeChol = a.precomputedBeTBeChol;
        }
        rc = postfit_explicit(pf);
        if (rc) return rc;
    }
    return interrupted ? 3 : 0;
}

int fit_implicit(const ImplicitArgs &a)
{
    if (a.k_user && a.U == nullptr && a.nnz_U == 0) return 2;
    if (a.k_item && a.II == nullptr && a.nnz_I == 0) return 2;
    if (a.nnz_U || a.nnz_I) return refuse("sparse side information (U_sp / I_sp)");
    if (a.NA_as_zero_U || a.NA_as_zero_I) return refuse("NA_as_zero_U / NA_as_zero_I");
    if (a.U && a.m_u != a.m) return refuse("side information U with a different number of rows than X");
    if (a.II && a.n_i != a.n) return refuse("side information I with a different number of rows than X has columns");
    if ((a.U && !a.C) || (a.II && !a.D)) return 2;
    const bool collective = a.U || a.II;
    if (collective && a.k_main) return refuse("k_main together with side information");
    if (a.nonneg || a.nonneg_C || a.nonneg_D) return refuse("non-negativity constraints");
    if (a.l1_lam != 0 || a.l1_lam_unique) return refuse("L1 regularisation");
    if (a.precondition_cg) return refuse("precondition_cg");
    if (a.k_user || a.k_item) return refuse("k_user / k_item without side information");
    if (!a.reset_values) return refuse("reset_values = false");
    if (a.m < 1 || a.n < 1 || a.k + a.k_main < 1) return 2;
    if (int rc = cuda_ready()) return rc;

    const int_t m = a.m, n = a.n;
    const int kk = a.k + a.k_main;
    const size_t nnz = a.nnz;
    bool use_cg = a.use_cg;
    const bool finalize_chol = a.finalize_chol && use_cg;

    StageTimer tm;
    // starting point on a host thread: A uniform (normal for tiny problems), B zero with CG (src/collective.c:9750-9774)
    const bool fill_B = a.II != nullptr;   // src/collective.c:9752
    std::thread rng([&]() {
        random_init(a.A, (size_t)m * kk, fill_B ? a.B : nullptr, fill_B ? (size_t)n * kk : 0, a.seed, false, rng_threads(a.nthreads));
        if (use_cg && !fill_B) std::memset(a.B, 0, (size_t)n * kk * sizeof(real_t));
    });
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{rng};

    // value transform (src/collective.c:9578-9599): log on the host (libm, as the reference), alpha on the device
    std::vector<real_t> Xlog;
    const real_t *Xsrc = a.X;
    if (a.apply_log_transf) {
        Xlog.assign(a.X, a.X + nnz);
        for (size_t e = 0; e < nnz; e++) Xlog[e] = std::log(Xlog[e]);
        Xsrc = Xlog.data();
    }

    // weight bookkeeping (src/collective.c:9776-9811)
    real_t w_main = a.w_main;
    real_t mult = 1;
    if (a.adjust_weight) {
        mult = (real_t)((long double)nnz / (long double)((size_t)m * (size_t)n));
        w_main *= mult;
    }
    if (a.w_main_multiplier) *a.w_main_multiplier = mult;
    real_t lamA = a.lam_unique ? a.lam_unique[2] : a.lam;
    real_t lamB = a.lam_unique ? a.lam_unique[3] : a.lam;
    real_t lamC = a.lam_unique ? a.lam_unique[4] : a.lam;
    real_t lamD = a.lam_unique ? a.lam_unique[5] : a.lam;
    real_t lam_plain = a.lam;
    real_t w_user = a.w_user, w_item = a.w_item;
    if (w_main != 1) {
        lamA /= w_main; lamB /= w_main; lamC /= w_main; lamD /= w_main; lam_plain /= w_main;
        w_user /= w_main; w_item /= w_main;
    }

    AlsConfig cfg;
    cfg.implicit = true;
    cfg.m = m; cfg.n = n; cfg.kk = kk;
    cfg.lam_A = lamA; cfg.lam_B = lamB;
    cfg.max_cg_steps = a.max_cg_steps;

    const WorldSetting &ws = world_setting();
    cfg.rank = ws.rank;
    cfg.world = ws.world;
    AlsState st;
    int rc = st.setup_from_coo(cfg, a.ixA, a.ixB, Xsrc, nnz, real_t(0), a.alpha, nullptr, nullptr,
                               cfg.world > 1 ? ws.nccl_id : nullptr);
    if (rc) return rc == 2 ? refuse("this value of k") : rc;
    tm.lap("upload COO + CSR/CSC (GPU)");
    rng.join();
    tm.lap("factor initialisation (host)");
    rc = st.upload_coordinates(a.A, (use_cg && !fill_B) ? nullptr : a.B);
    if (rc) return rc;
    InterruptScope interrupt_scope;
    st.verbose = a.verbose;
    bool interrupted = false;
    if (a.verbose) { std::printf("Starting ALS optimization routine\n\n"); std::fflush(stdout); }
    if (!collective) {
        rc = st.iterate(0, a.niter, a.niter, use_cg, finalize_chol);
        if (rc == 3) interrupted = true;
        else if (rc) return rc == 2 ? refuse("this solver / k combination") : rc;
    } else {
        // implicit feedback + dense side information (src/collective.c:9832-10022: C, D, then B and A through
        // optimizeA_collective_implicit): column-centre U / I, then the same machinery as the explicit collective model
        if (cfg.world > 1) return refuse("side information on more than one GPU");
        std::vector<real_t> Uc, Ic;
        CollectiveConfig cc;
        if (a.U) {
            cc.p = a.p;
            if (a.U_colmeans) { if (center_side_info(a.U, m, a.p, a.U_colmeans, Uc)) return refuse("missing values in U"); }
            else Uc.assign(a.U, a.U + (size_t)m * a.p);
        }
        if (a.II) {
            cc.q = a.q;
            if (a.I_colmeans) { if (center_side_info(a.II, n, a.q, a.I_colmeans, Ic)) return refuse("missing values in I"); }
            else Ic.assign(a.II, a.II + (size_t)n * a.q);
        }
        for (real_t v : Uc) if (std::isnan(v)) return refuse("missing values in U");
        for (real_t v : Ic) if (std::isnan(v)) return refuse("missing values in I");
        cc.w_user = w_user; cc.w_item = w_item;
        cc.lam_C = lamC / w_user;
        cc.lam_D = lamD / w_item;
        CollectiveState cs;
        if ((rc = cs.setup(&st, cc, Uc.data(), Ic.data()))) return rc;
        rc = cs.iterate(a.niter, use_cg, finalize_chol);
        if (rc == 3) interrupted = true;
        else if (rc) return rc == 2 ? refuse("this solver / k combination") : rc;
        if ((rc = cs.download(a.C, a.D, nullptr, nullptr))) return rc;
    }
    tm.lap("ALS iterations");
    rc = st.download_factors(a.A, kk, nullptr, a.B, kk, nullptr);
    if (rc) return rc;
    tm.lap("download factors");

    if (a.verbose && !interrupted) {
        std::printf(std::isnan((double)a.A[0]) ? "ALS procedure failed\n" : "ALS procedure terminated successfully\n");
        std::fflush(stdout);
    }
    if (a.precompute_for_predictions && (!interrupted || a.handle_interrupt)) {
        // BtB + lam*I (src/collective.c:10057-10074); with user side information also BeTBe and its Cholesky factor
        if (a.verbose) { std::printf("Finishing precomputed matrices..."); std::fflush(stdout); }
        postfit_implicit(a.B, n, kk, lam_plain, a.precomputedBtB, a.U ? a.C : nullptr, a.U ? a.p : 0, w_user, a.precomputedBeTBe,
                         a.precomputedBeTBeChol, use_cg && !finalize_chol);
        if (a.verbose) { std::printf("  done\n"); std::fflush(stdout); }
    }
    return interrupted ? 3 : 0;
}

}  // namespace cmfb200
