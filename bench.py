#!/usr/bin/env python
"""Benchmark of the ALS hot path (BASELINE.json metric: rows solved / s and s per ALS iteration at k=64 on an
ML10M-shaped CSR).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one full ALS iteration (B half-sweep + A half-sweep) over the whole synthetic matrix.
  value        rows solved per second, whole job, inputs resident in HBM (CUDA-event time of the K steps,
               max over ranks, L2 flushed between steps)
  e2e          the same metric through the host-pointer C ABI (fit_collective_*_als with numpy buffers):
               centring, COO->CSR/CSC, bias / factor initialisation, H2D, K iterations, D2H all inside the timed region
  roofline     achieved algorithmic GB/s of the row-solve kernel (CUDA events around every launch of it inside the
               timed region) against the measured HBM copy bandwidth
  cpu_baseline the reference's own OpenMP+BLAS implementation (oracle/_ref, built from /root/reference) on the box's
               host cores, same workload
`--impl reference` times only that CPU implementation and prints the line with "impl": "reference".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly one JSON line: whatever a library prints on file descriptor 1 during the run (NCCL's version
# banner, for one) is sent to stderr, and the line itself is written to the original descriptor at the end
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: shape, implicit, k, dtype, solver, extra
    "ml10m_explicit_cg_k64_f32": dict(shape="ml10m", implicit=False, k=64, dtype="f32", use_cg=True),
    "lastfm_implicit_cg_k64_f32": dict(shape="lastfm", implicit=True, k=64, dtype="f32", use_cg=True),
    "ml10m_explicit_chol_k64_f32": dict(shape="ml10m", implicit=False, k=64, dtype="f32", use_cg=False),
    "ml10m_explicit_chol_k128_f32": dict(shape="ml10m", implicit=False, k=128, dtype="f32", use_cg=False),
    "lastfm_implicit_chol_k64_f32": dict(shape="lastfm", implicit=True, k=64, dtype="f32", use_cg=False),
    "ml10m_explicit_cg_k128_f32": dict(shape="ml10m", implicit=False, k=128, dtype="f32", use_cg=True),
    "lastfm_implicit_cg_k128_f32": dict(shape="lastfm", implicit=True, k=128, dtype="f32", use_cg=True),
    "lastfm_implicit_cg_k256_f32": dict(shape="lastfm", implicit=True, k=256, dtype="f32", use_cg=True),
    "cfg1_explicit_cg_k16_f64": dict(shape="cfg1", implicit=False, k=16, dtype="f64", use_cg=True),
    # BASELINE.json config 3: Cholesky, fp64, k=128, dense user and item side information (p = q = 32, N(0,1), seed 3)
    "ml10m_explicit_chol_k128_f64_sideinfo": dict(shape="ml10m", implicit=False, k=128, dtype="f64", use_cg=False, side=32),
    # BASELINE.json config 4: CG, fp32, k=64, add_implicit_features (w_implicit = 0.5)
    "ml10m_explicit_cg_k64_f32_implicit_features": dict(shape="ml10m", implicit=False, k=64, dtype="f32", use_cg=True,
                                                        implicit_features=True, w_implicit=0.5),
    # BASELINE.json config 5: CMF_implicit ALS-CG k=256 fp32, 10 M x 1 M, ~1e9 entries, generated ON THE DEVICE;
    # --scale shrinks m, n (and so nnz) for smaller boxes: 1.0 is the full configuration, meant for 8 GPUs
    "cfg5_implicit_cg_k256_f32": dict(shape="cfg5", implicit=True, k=256, dtype="f32", use_cg=True, device_generated=True),
}
HYPER = dict(explicit=dict(lam=0.05, scale_lam=True, user_bias=True, item_bias=True, center=True, max_cg_steps=3),
             implicit=dict(lam=5.0, alpha=1.0, max_cg_steps=3))


def load_data(w):
    from cmfrec_b200 import synth
    dt = np.dtype(np.float32 if w["dtype"] == "f32" else np.float64)
    cache = os.path.join("/tmp", "cmfb200_%s_%s.npz" % (w["shape"], w["dtype"]))
    if os.path.exists(cache):
        z = np.load(cache)
        return z["a"], z["b"], z["x"], int(z["m"]), int(z["n"]), dt
    a, b, x, m, n = synth.make(w["shape"], dt)
    try:
        np.savez(cache + ".tmp.npz", a=a, b=b, x=x, m=m, n=n)
        os.replace(cache + ".tmp.npz", cache)
    except OSError:
        pass
    return a, b, x, m, n, dt


def side_info(w, m, n, dt):
    if not w.get("side"):
        return None, None
    rng = np.random.default_rng(3)
    return rng.normal(size=(m, w["side"])).astype(dt), rng.normal(size=(n, w["side"])).astype(dt)


def algorithmic_bytes(w, m, n, nnz, dt):
    """every stored entry gathers its opposing row once per half-sweep, plus CSR index+value,
    plus indptr and read+write of the solved row; implicit adds one pass over the opposing factor for the Gram."""
    wd = dt.itemsize
    k1 = w["k"] + (0 if w["implicit"] else 1)
    extra_gather = w["k"] * wd if w.get("implicit_features") else 0      # the q-vector gathers one Bi row per entry
    half = lambda rows, opp: (nnz * (k1 * wd + 4 + wd + extra_gather) + rows * (8 + 2 * k1 * wd)
                              + (opp * w["k"] * wd if w["implicit"] else 0) + rows * w.get("side", 0) * wd)
    return half(n, m), half(m, n)       # (B sweep, A sweep)


def measured_dram_traffic(w):
    """DRAM bytes per launch of the row-solve kernel from the committed `ncu --set full` capture of this workload's
    shape (profiles/, condensed by tools/summarize_ncu.py); None when no capture of that shape / model is committed."""
    import csv
    if not w["use_cg"] or w["k"] != 64 or w.get("side") or w.get("implicit_features"):
        return None, None
    path = os.path.join(ROOT, "profiles", "r1_ncu_full_cg_sweep_%s.csv" % w["shape"])
    if not os.path.exists(path):
        return None, None
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
    rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    per = [float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]] for r in rows[2:] if "cg_resident_kernel" in r[1]]
    if not per:
        return None, None
    return sum(per) / len(per), os.path.relpath(path, ROOT)


class ClockSampler:
    """SM clock and throttle reasons polled through NVML (every ~2 ms) while the timed region runs."""

    def __init__(self, dev):
        self.dev, self.sm, self.reasons, self.max_mhz, self._stop, self.t, self.h = dev, [], set(), None, False, None, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(visible.split(",")[dev]) if visible and visible.split(",")[dev].isdigit() else dev
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.h = None

    def _poll(self):
        nv = self.nv
        names = {getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                 getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                 getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                 getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap"}
        while not self._stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.002)

    def start(self):
        if self.h is None:
            return
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def stop(self):
        if self.h is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvml unavailable"], samples=0)
        self._stop = True
        self.t.join(timeout=1)
        return dict(sm_mhz=float(np.median(self.sm)) if self.sm else None, sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons), samples=len(self.sm))


def time_reference(w, data, steps, warmup, nthreads):
    """The reference's own CPU implementation of the same fit: sec / iteration = (t(K iters) - t(0 iters)) / K."""
    import refload
    from cmfrec_b200.calls import fit_explicit, fit_implicit
    a, b, x, m, n, dt = data
    R = refload.ref(dt)
    kind = "reference"
    if R is None:
        raise RuntimeError("oracle/_ref missing: build it where /root/reference exists (make -C oracle ref)")
    if w["implicit"]:
        h = HYPER["implicit"]
        run = lambda it: fit_implicit(R, dt, a, b, x, m, n, w["k"], lam=h["lam"], alpha=h["alpha"], niter=it,
                                      use_cg=w["use_cg"], max_cg_steps=h["max_cg_steps"], nthreads=nthreads)
    else:
        h = HYPER["explicit"]
        U, I = side_info(w, m, n, dt)
        extra = dict(U=U, I=I, add_implicit_features=bool(w.get("implicit_features")), w_implicit=w.get("w_implicit", 1.0))
        run = lambda it: fit_explicit(R, dt, a, b, x, m, n, w["k"], lam=h["lam"], scale_lam=h["scale_lam"], niter=it,
                                      use_cg=w["use_cg"], max_cg_steps=h["max_cg_steps"], nthreads=nthreads, **extra)
    run(max(1, min(warmup, 1)))
    t0 = time.perf_counter(); run(0); t_prep = time.perf_counter() - t0
    t0 = time.perf_counter(); out = run(steps); t_full = time.perf_counter() - t0
    assert out["rc"] == 0
    if t_full - t_prep < 0.2 * t_full:      # too short to time by difference: repeat with many more iterations
        big = steps * 200
        t0 = time.perf_counter(); run(big); t_big = time.perf_counter() - t0
        sec_iter = max(t_big - t_prep, 1e-9) / big
    else:
        sec_iter = max(t_full - t_prep, 1e-9) / steps
    return dict(sec_iter=sec_iter, rows_per_s=(m + n) / sec_iter, kind=kind, prep_s=t_prep, call_s=t_full,
                call_rows_per_s=(m + n) * steps / t_full,
                sample="full workload, one fit call of %d ALS iterations, nthreads=%d" % (steps, nthreads))


def generate_cfg5_on_device(torch, m, n, seed):
    """COO triplets of BASELINE config 5 on the current CUDA device (every rank generates the same arrays): row degrees
    lognormal with mean 100, columns Zipf-like (shifted, exponent 1; ranks scattered over the ids by an affine map),
    values ceil(lognormal(1, 1.5)) -- SURVEY.md 8(d).  Returns int32 rows, int32 cols, float32 values."""
    import math
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    deg = torch.exp(torch.randn(m, generator=g, device=dev) + (math.log(100.0) - 0.5)).round_().clamp_(1, max(1, n // 2)).to(torch.int64)
    nnz = int(deg.sum().item())
    assert nnz < 2 ** 31, "the device-side compression indexes entries with 32 bits"
    rows = torch.repeat_interleave(torch.arange(m, dtype=torch.int32, device=dev), deg, output_size=nnz)
    del deg
    cols = torch.empty(nnz, dtype=torch.int32, device=dev)
    vals = torch.empty(nnz, dtype=torch.float32, device=dev)
    step = 1 << 26
    r0 = max(1.0, n / 10000.0)
    mult = 2654435761 % n
    while math.gcd(mult, n) != 1:
        mult += 1
    for s in range(0, nnz, step):
        e = min(nnz, s + step)
        u = torch.rand(e - s, generator=g, device=dev, dtype=torch.float64)
        # shifted Zipf, p(rank) ~ 1 / (rank + r0): the most popular item holds ~0.1 % of the entries (LastFM-360K's top artist
        # holds 0.45 %), not the 7 % an unshifted exponent-1 law over 10^6 items would give it
        c = (torch.exp(u * math.log((n + r0) / r0)) * r0 - r0).to(torch.int64).clamp_(0, n - 1)
        cols[s:e] = ((c * mult + 12345) % n).to(torch.int32)
        z = torch.randn(e - s, generator=g, device=dev)
        vals[s:e] = torch.ceil(torch.exp(1.0 + 1.5 * z)).clamp_(1.0, 1e6)
        del u, c, z
    return rows, cols, vals, nnz


def run_device_generated(args, w, config, rank, world, local_rank):
    """Workloads whose data never exists on the host (config 5): generate on the device, build / deal on the device
    (cmfb200_als_create_from_device_coo), device-side starting factors, timed iterations as in main()."""
    import torch
    import torch.distributed as dist
    from cmfrec_b200 import _lib
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dt = np.dtype(np.float32)
    L = _lib.load(dt)
    if L.cmfb200_device_count() < 1:
        raise RuntimeError("no CUDA device: cmfrec_b200 has no CPU path")
    m, n = max(64, int(10_000_000 * args.scale)), max(64, int(1_000_000 * args.scale))
    t0 = time.perf_counter()
    rows, cols, vals, nnz = generate_cfg5_on_device(torch, m, n, 20260105)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    h = HYPER["implicit"]
    config.update(m=m, n=n, nnz=nnz, scale=args.scale, generated="on the device, every rank the same arrays (seed 20260105)")
    stream = torch.cuda.current_stream().cuda_stream
    opt = L.AlsOptions()
    opt.implicit = 1
    opt.m, opt.n, opt.k = m, n, w["k"]
    opt.lam_A = opt.lam_B = opt.lam_biasA = opt.lam_biasB = h["lam"]
    opt.max_cg_steps = h["max_cg_steps"]
    opt.rank, opt.world = rank, world
    idbuf = None
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_ubyte * 128)()
            assert L.cmfb200_nccl_unique_id(raw) == 0
            idt = torch.tensor(list(raw), dtype=torch.uint8)
        idt = idt.cuda()
        dist.broadcast(idt, 0)
        idbuf = (C.c_ubyte * 128)(*idt.cpu().tolist())
        opt.nccl_id = C.cast(idbuf, C.c_void_p)
    opt.stream = stream
    hnd = C.c_void_p()
    t0 = time.perf_counter()
    rc = L.cmfb200_als_create_from_device_coo(C.byref(hnd), C.byref(opt), C.c_void_p(rows.data_ptr()), C.c_void_p(cols.data_ptr()),
                                              C.c_void_p(vals.data_ptr()), nnz, 0.0, h["alpha"])
    assert rc == 0, "cmfb200_als_create_from_device_coo -> %d" % rc
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    del rows, cols, vals
    torch.cuda.empty_cache()
    assert L.cmfb200_als_random_factors(hnd, 1, 2.0 ** -7) == 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    ms = C.c_float(0)
    it = 0
    for _ in range(max(args.warmup, 3)):
        assert L.cmfb200_als_timed_iterate(hnd, it, 1, 1 << 30, 1, 0, C.byref(ms)) == 0
        it += 1
    L.cmfb200_als_set_profile(hnd, 1)
    launches0 = L.cmfb200_als_launch_count(hnd)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    total_ms = 0.0
    for _ in range(args.steps):
        flush.zero_()
        assert L.cmfb200_als_timed_iterate(hnd, it, 1, 1 << 30, 1, 0, C.byref(ms)) == 0
        total_ms += ms.value
        it += 1
    barrier()
    clocks = sampler.stop()
    launches = L.cmfb200_als_launch_count(hnd) - launches0
    kt = [C.c_double(0), C.c_double(0)]
    kc = [C.c_longlong(0), C.c_longlong(0)]
    for which in (0, 1):
        L.cmfb200_als_read_profile(hnd, which, C.byref(kt[which]), C.byref(kc[which]))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    bytes_B, bytes_A = algorithmic_bytes(w, m, n, nnz, dt)
    kernel_ms = kt[0].value + kt[1].value
    kernel_launches = kc[0].value + kc[1].value
    bytes_per_launch = (bytes_B + bytes_A) / 2.0 / world
    avg_ms = kernel_ms / max(kernel_launches, 1)
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=None,
                    kernel="cg_resident_kernel", avg_launch_ms=avg_ms, launches_timed=int(kernel_launches),
                    algorithmic_bytes_per_launch=bytes_per_launch, kernel_share_of_step=kernel_ms / total_ms if total_ms else None,
                    peak_source=peak_src, B_sweep_ms=kt[0].value / max(kc[0].value, 1), A_sweep_ms=kt[1].value / max(kc[1].value, 1))
    L.cmfb200_als_destroy(hnd)
    if rank == 0:
        emit(dict(metric="rows_solved_per_sec", value=(m + n) / (ms_per_step * 1e-3), unit="rows/s", n_gpus=world, steps=args.steps,
                  warmup=max(args.warmup, 3), ms_per_step=ms_per_step, higher_is_better=True, scaling="strong", vs_baseline=None,
                  dtype=w["dtype"], data="synthetic", config=config, roofline=roofline, clocks=clocks, gpu_launches=int(launches),
                  e2e=None, e2e_note="the triplets never exist on the host at this size: generated on the device in %.1f s, "
                                     "CSR / CSC built and dealt on the device in %.1f s (both outside the timed region)" % (t_gen, t_setup),
                  cpu_baseline=None, sec_per_iter=ms_per_step * 1e-3))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ml10m_explicit_cg_k64_f32", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scale", type=float, default=1.0, help="size factor of the device-generated workload (cfg5)")
    args = ap.parse_args()
    claim_stdout()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncpu = os.cpu_count() or 1
    config = dict(workload=args.workload, shape=w["shape"], k=w["k"], solver="cg" if w["use_cg"] else "cholesky",
                  feedback="implicit" if w["implicit"] else "explicit",
                  hyper=HYPER["implicit" if w["implicit"] else "explicit"],
                  l2="flushed between steps (256 MiB memset); CSR+CSC streamed per step exceed L2",
                  parallelism="rows of A and of B dealt over %d GPU(s); all-gather after each half-sweep" % world)

    if args.impl == "reference":
        if rank != 0:
            return
        data = load_data(w)
        r = time_reference(w, data, args.steps, args.warmup, ncpu)
        m, n = data[3], data[4]
        config.update(m=m, n=n, nnz=int(data[2].size))
        # value = the whole fit call (preprocessing + K iterations, host buffers in and out): the same thing our
        # arm reports as e2e.  The iteration-only rate (fit time minus a 0-iteration fit) is given beside it.
        v = r["call_rows_per_s"]
        line = dict(impl="reference", metric="rows_solved_per_sec", value=v, unit="rows/s",
                    n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=r["call_s"] / args.steps * 1e3,
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype=w["dtype"], data="synthetic",
                    config=config,
                    cpu_baseline=dict(value=v, unit="rows/s", cores=ncpu, kind=r["kind"], sample=r["sample"],
                                      iteration_only_rows_per_s=r["rows_per_s"], sec_per_iter=r["sec_iter"],
                                      preprocessing_s=r["prep_s"]),
                    e2e=dict(value=v, unit="rows/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0, sec_per_iter=r["sec_iter"])
        emit(line)
        return

    if w.get("device_generated"):
        return run_device_generated(args, w, config, rank, world, local_rank)

    import torch
    import torch.distributed as dist
    from cmfrec_b200 import _lib
    from cmfrec_b200.calls import csr_csc, fit_explicit, fit_implicit, ptr

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    a, b, x, m, n, dt = data = load_data(w)
    nnz = int(x.size)
    config.update(m=m, n=n, nnz=nnz)
    L = _lib.load(dt)
    if L.cmfb200_device_count() < 1:
        raise RuntimeError("no CUDA device: cmfrec_b200 has no CPU path")

    # ---- prepare exactly what the fit entry point prepares
    if w["implicit"]:
        h = HYPER["implicit"]
        xs = (x * dt.type(h["alpha"])).astype(dt)
        csr = csr_csc(L, dt, a, b, xs, m, n)
        A0 = np.zeros((m, w["k"]), dt); B0 = np.zeros((n, w["k"]), dt)
        L.cmfb200_random_init(ptr(A0), A0.size, None, 0, 1, False)
        bA = bB = None
        lamA = lamB = lbA = lbB = h["lam"]
    else:
        h = HYPER["explicit"]
        mu = L.cmfb200_global_mean(ptr(x), nnz, ncpu)
        xc = (x - dt.type(mu)).astype(dt)
        csr = csr_csc(L, dt, a, b, xc, m, n)
        bA = np.zeros(m, dt); bB = np.zeros(n, dt)
        L.cmfb200_init_biases_twosided(m, n, *[ptr(t) for t in csr], h["lam"], h["lam"], h["scale_lam"], False, ptr(bA), ptr(bB), ncpu)
        A0 = np.zeros((m, w["k"]), dt); B0 = np.zeros((n, w["k"]), dt)
        fill_B = bool(w.get("side") or w.get("implicit_features"))
        L.cmfb200_random_init(ptr(A0), A0.size, ptr(B0) if fill_B else None, B0.size if fill_B else 0, 1, True)
        lamA = lamB = lbA = lbB = h["lam"]

    stream = torch.cuda.current_stream().cuda_stream
    opt = L.AlsOptions()
    opt.implicit = int(w["implicit"])
    opt.m, opt.n, opt.k = m, n, w["k"]
    opt.user_bias = opt.item_bias = int(not w["implicit"])
    opt.lam_A, opt.lam_B, opt.lam_biasA, opt.lam_biasB = lamA, lamB, lbA, lbB
    opt.scale_lam = int((not w["implicit"]) and h["scale_lam"])
    opt.max_cg_steps = h["max_cg_steps"]
    opt.rank, opt.world = rank, world
    idbufs = []

    def fresh_nccl_id():
        """a new communicator id (single use) made on rank 0 and broadcast through torch.distributed"""
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_ubyte * 128)()
            assert L.cmfb200_nccl_unique_id(raw) == 0
            idt = torch.tensor(list(raw), dtype=torch.uint8)
        idt = idt.cuda()
        dist.broadcast(idt, 0)
        idbufs.append((C.c_ubyte * 128)(*idt.cpu().tolist()))
        opt.nccl_id = C.cast(idbufs[-1], C.c_void_p)

    if world > 1:
        fresh_nccl_id()
    opt.stream = stream
    hnd = C.c_void_p()
    rc = L.cmfb200_als_create(C.byref(hnd), C.byref(opt), *[ptr(t) for t in csr])
    assert rc == 0, "cmfb200_als_create -> %d" % rc
    assert L.cmfb200_als_set_factors(hnd, ptr(A0), ptr(bA), ptr(B0), ptr(bB)) == 0
    if w.get("side") or w.get("implicit_features"):
        assert world == 1 or not w.get("side"), "dense side information is single-GPU; implicit features shard like A / B"
        U, I = side_info(w, m, n, dt)
        if U is not None:
            U = (U - U.mean(axis=0, dtype=dt)).astype(dt); I = (I - I.mean(axis=0, dtype=dt)).astype(dt)
        wi = w.get("w_implicit", 1.0)
        sm, sn = (float(m), float(n)) if h["scale_lam"] else (1.0, 1.0)
        rc = L.cmfb200_als_attach_collective(hnd, ptr(U), w.get("side", 0), ptr(I), w.get("side", 0),
                                             int(bool(w.get("implicit_features"))), 1.0, 1.0, wi, h["lam"] * sm, h["lam"] * sn,
                                             h["lam"] / wi * sm, h["lam"] / wi * sn)
        assert rc == 0, "attach_collective -> %d" % rc
    use_cg = int(w["use_cg"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    ms = C.c_float(0)
    it = 0
    for _ in range(max(args.warmup, 3)):
        assert L.cmfb200_als_timed_iterate(hnd, it, 1, 1 << 30, use_cg, 0, C.byref(ms)) == 0
        it += 1
    L.cmfb200_als_set_profile(hnd, 1)
    launches0 = L.cmfb200_als_launch_count(hnd)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    total_ms = 0.0
    for _ in range(args.steps):
        flush.zero_()                                           # evict L2 (not timed)
        assert L.cmfb200_als_timed_iterate(hnd, it, 1, 1 << 30, use_cg, 0, C.byref(ms)) == 0
        total_ms += ms.value
        it += 1
    barrier()
    clocks = sampler.stop()
    launches = L.cmfb200_als_launch_count(hnd) - launches0
    kt = [C.c_double(0), C.c_double(0)]
    kc = [C.c_longlong(0), C.c_longlong(0)]
    for which in (0, 1):
        L.cmfb200_als_read_profile(hnd, which, C.byref(kt[which]), C.byref(kc[which]))
    L.cmfb200_als_set_profile(hnd, 0)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = (m + n) / (ms_per_step * 1e-3)

    # roofline of the row-solve kernel (both launches of a step are the same kernel on the two orientations)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    bytes_B, bytes_A = algorithmic_bytes(w, m, n, nnz, dt)
    kernel_ms = kt[0].value + kt[1].value
    kernel_launches = kc[0].value + kc[1].value
    bytes_per_launch = (bytes_B + bytes_A) / 2.0 / world
    avg_ms = kernel_ms / max(kernel_launches, 1)
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    traffic, traffic_src = measured_dram_traffic(w)
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                    traffic_unit="GB per launch (dram__bytes_read.sum + dram__bytes_write.sum)", traffic_source=traffic_src,
                    kernel="cg_resident_kernel" if w["use_cg"] else "chol_sweep_kernel", avg_launch_ms=avg_ms,
                    launches_timed=int(kernel_launches), algorithmic_bytes_per_launch=bytes_per_launch,
                    kernel_share_of_step=kernel_ms / total_ms if total_ms else None, peak_source=peak_src,
                    B_sweep_ms=kt[0].value / max(kc[0].value, 1), A_sweep_ms=kt[1].value / max(kc[1].value, 1))
    if not w["use_cg"]:
        # exact (Cholesky) sweeps are bound by the tensor pipe, not by HBM: algorithmic flop of the normal matrices and their
        # factorisations (nnz k'(k'+1) + rows (k'^3 / 3 + 2 k'^2) per half-sweep) against the measured GEMM peak
        k1 = w["k"] + (0 if w["implicit"] else 1)
        flop = 2 * nnz * k1 * (k1 + 1) + (m + n) * (k1 ** 3 / 3.0 + 2 * k1 * k1)
        fp = json.load(open(os.path.join(ROOT, "profiles", "r2_fp_peaks.json")))
        tpeak = fp["fp64_tflops"] if w["dtype"] == "f64" else fp["tf32_tflops"]
        ach = flop / 2.0 / world / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        roofline.update(bound="tensor", achieved=ach, peak=tpeak, unit="TFLOP/s", frac=ach / tpeak,
                        kernel="chol_dmma_build_kernel + chol_dmma_factor_kernel" if w["dtype"] == "f64" else "nm_sweep_kernel",
                        algorithmic_flop_per_launch=flop / 2.0 / world, hbm_model_frac=achieved / peak,
                        peak_source="profiles/r2_fp_peaks.json: measured %s GEMM (torch.matmul 8192^3); the fp32 kernel issues 3 TF32 "
                                    "products per algorithmic one" % ("FP64" if w["dtype"] == "f64" else "TF32"))
    L.cmfb200_als_destroy(hnd)
    del flush

    # ---- end to end through the host-pointer C ABI
    e2e = None
    if not args.no_e2e:
        K = args.steps
        if world == 1:
            # the caller's buffers: COO triplets and the output factors in pinned host memory, handed to the entry point as
            # they are (the library never writes to its inputs, so no defensive copies)
            pin = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).pin_memory().numpy()
            pa, pb, px = pin(a), pin(b), pin(x)
            outbuf = dict(A=pin(np.zeros((m, w["k"]), dt)), B=pin(np.zeros((n, w["k"]), dt)))
            if w["implicit"]:
                run = lambda nit: fit_implicit(L, dt, pa, pb, px, m, n, w["k"], lam=h["lam"], alpha=h["alpha"], niter=nit,
                                               use_cg=w["use_cg"], max_cg_steps=h["max_cg_steps"], nthreads=ncpu,
                                               copy_inputs=False, out=outbuf)
            else:
                outbuf.update(biasA=pin(np.zeros(m, dt)), biasB=pin(np.zeros(n, dt)))
                U, I = side_info(w, m, n, dt)
                extra = dict(U=U, I=I, add_implicit_features=bool(w.get("implicit_features")), w_implicit=w.get("w_implicit", 1.0))
                run = lambda nit: fit_explicit(L, dt, pa, pb, px, m, n, w["k"], lam=h["lam"], scale_lam=h["scale_lam"],
                                               niter=nit, use_cg=w["use_cg"], max_cg_steps=h["max_cg_steps"], nthreads=ncpu,
                                               copy_inputs=False, out=outbuf, **extra)
            run(1)
            run(2)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = run(K)
            t_e2e = time.perf_counter() - t0
            assert out["rc"] == 0
            wd = dt.itemsize
            ld = ((w["k"] + 1) * wd + 15) // 16 * 16
            h2d = 2 * (nnz * (4 + wd)) + (m + n + 2) * 8 + (m + n) * ld + (m + n) * 4
            d2h = (m + n) * ld
            e2e = dict(value=(m + n) * K / t_e2e, unit="rows/s", h2d_bytes_per_step=h2d / K, d2h_bytes_per_step=d2h / K,
                       seconds=t_e2e, iterations=K,
                       what="one fit_collective_%s_als call, pinned host buffers in / out" % ("implicit" if w["implicit"] else "explicit"))
        else:
            # N > 1: the SAME reference-named entry point on every rank after cmfb200_set_world (rank, world, NCCL id):
            # every rank passes the full COO triplets from pinned host memory and receives the full factors; ingestion
            # (CSR / CSC, centring, biases), the dealing of the rows to the ranks, K iterations with their all-gathers and
            # the download are all inside the timed region
            assert not w.get("side")
            # host threads per rank: the ranks share the box's cores; at least 8, the reference's threshold for its parallel mean
            nthr = max(8, ncpu // world)
            pin = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).pin_memory().numpy()
            pa, pb, px = pin(a), pin(b), pin(x)
            outbuf = dict(A=pin(np.zeros((m, w["k"]), dt)), B=pin(np.zeros((n, w["k"]), dt)))
            if w["implicit"]:
                run = lambda nit: fit_implicit(L, dt, pa, pb, px, m, n, w["k"], lam=h["lam"], alpha=h["alpha"], niter=nit,
                                               use_cg=w["use_cg"], max_cg_steps=h["max_cg_steps"], nthreads=nthr,
                                               copy_inputs=False, out=outbuf)
            else:
                outbuf.update(biasA=pin(np.zeros(m, dt)), biasB=pin(np.zeros(n, dt)))
                extra = dict(add_implicit_features=bool(w.get("implicit_features")), w_implicit=w.get("w_implicit", 1.0))
                run = lambda nit: fit_explicit(L, dt, pa, pb, px, m, n, w["k"], lam=h["lam"], scale_lam=h["scale_lam"],
                                               niter=nit, use_cg=w["use_cg"], max_cg_steps=h["max_cg_steps"], nthreads=nthr,
                                               copy_inputs=False, out=outbuf, **extra)
            fresh_nccl_id()
            assert L.cmfb200_set_world(rank, world, opt.nccl_id) == 0
            assert run(1)["rc"] == 0
            assert run(2)["rc"] == 0
            barrier()
            t0 = time.perf_counter()
            out = run(K)
            barrier()
            t_e2e = time.perf_counter() - t0
            assert out["rc"] == 0
            L.cmfb200_set_world(0, 1, None)
            tt = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_e2e = float(tt.item())
            wd = dt.itemsize
            h2d = nnz * (8 + wd) + (m + n) * w["k"] * wd
            d2h = (m + n) * (w["k"] + 1) * wd
            e2e = dict(value=(m + n) * K / t_e2e, unit="rows/s", h2d_bytes_per_step=h2d / K, d2h_bytes_per_step=d2h / K,
                       seconds=t_e2e, iterations=K,
                       what="one fit_collective_%s_als call per rank after cmfb200_set_world, pinned host buffers in / out, max "
                            "over ranks (h2d / d2h bytes are per rank)" % ("implicit" if w["implicit"] else "explicit"))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = time_reference(w, data, 2, 1, ncpu)
            cpu = dict(value=r["rows_per_s"], unit="rows/s", cores=ncpu, kind=r["kind"],
                       sample=r["sample"] + "; value = iteration-only rate (fit time minus a 0-iteration fit)",
                       sec_per_iter=r["sec_iter"], whole_call_rows_per_s=r["call_rows_per_s"], preprocessing_s=r["prep_s"])
        except Exception as e:  # noqa: BLE001
            cpu = dict(value=None, unit="rows/s", cores=ncpu, kind="reference", sample="unavailable: %s" % e)

    if rank == 0:
        line = dict(metric="rows_solved_per_sec", value=value, unit="rows/s", n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms_per_step, higher_is_better=True, scaling="strong",
                    vs_baseline=None, dtype=w["dtype"], data="synthetic", config=config, sec_per_iter=ms_per_step * 1e-3,
                    roofline=roofline, cpu_baseline=cpu, e2e=e2e, gpu_launches=int(launches), clocks=clocks)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
