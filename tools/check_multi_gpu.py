#!/usr/bin/env python
"""Run under torchrun with N ranks (one GPU each).  Two checks, each against the 1-GPU result computed on rank 0:
  * the building-block API (cmfb200_als_create with world > 1: host-side dealing),
  * the reference-named fit entry points after cmfb200_set_world (device-side dealing, biases, Gram all-reduce).
Explicit-feedback fits must be bit-identical (a row's arithmetic does not depend on where it is solved); implicit ones
differ by the summation order of the all-reduced Gram matrix only."""
import ctypes as C
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
from cmfrec_b200 import _lib
from cmfrec_b200.multi import ShardedAls, nccl_id_for_all_ranks
from support import csr_csc, fit_explicit, fit_implicit, synth_coo

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True


def report(what, got, want, implicit, dt):
    global ok
    same = all(np.array_equal(a, b) for a, b in zip(got, want))
    scale = max(float(np.abs(b).max()) for b in want)
    err = max(float(np.abs(a.astype(np.float64) - b).max()) for a, b in zip(got, want)) / max(scale, 1e-30)
    good = same if not implicit else err < (5e-3 if dt == np.float32 else 1e-8)
    print("%-46s %s world=%d identical=%s max rel diff=%.3e %s" % (what, dt.name, world, same, err, "ok" if good else "FAIL"), flush=True)
    ok = ok and good


for implicit in (False, True):
    for dt in (np.dtype(np.float32), np.dtype(np.float64)):
        L = _lib.load(dt)
        m, n, k = 5003, 3001, 32
        ixA, ixB, X = synth_coo(m, n, 200000, dt, seed=4, kind="counts" if implicit else "ratings")
        # ---- building blocks
        Xc = X if implicit else (X - X.mean()).astype(dt)
        csr = csr_csc(L, dt, ixA, ixB, Xc, m, n)
        rng = np.random.default_rng(1)
        A0 = (rng.random((m, k)) * 0.1).astype(dt); B0 = np.zeros((n, k), dt)
        bA = (rng.normal(size=m) * 0.1).astype(dt); bB = (rng.normal(size=n) * 0.1).astype(dt)
        kw = dict(implicit=implicit, lam_A=2.0, lam_B=2.0, user_bias=not implicit, item_bias=not implicit)
        nid = nccl_id_for_all_ranks(L, rank, world)
        with ShardedAls(dt, csr[:3], csr[3:], m, n, k, rank=rank, world=world, nccl_id=nid, **kw) as s:
            s.set_factors(A0, bA, B0, bB)
            s.iterate(0, 3, 3, use_cg=True, finalize_chol=True)
            got = s.get_factors()
        if rank == 0:
            with ShardedAls(dt, csr[:3], csr[3:], m, n, k, **kw) as s:
                s.set_factors(A0, bA, B0, bB)
                s.iterate(0, 3, 3, use_cg=True, finalize_chol=True)
                want = s.get_factors()
            report("building blocks implicit=%d" % implicit, got, want, implicit, dt)
        dist.barrier()
        # ---- the reference-named entry point on every rank
        nid = nccl_id_for_all_ranks(L, rank, world)
        assert L.cmfb200_set_world(rank, world, C.cast(nid, C.c_void_p)) == 0
        fit = (lambda: fit_implicit(L, dt, ixA, ixB, X, m, n, k, niter=3, finalize_chol=True)) if implicit else \
              (lambda: fit_explicit(L, dt, ixA, ixB, X, m, n, k, lam=0.05, scale_lam=True, niter=3, finalize_chol=True))
        a = fit()
        assert a["rc"] == 0, a["rc"]
        assert L.cmfb200_set_world(0, 1, None) == 0
        if rank == 0:
            b = fit()
            keys = ("A", "B") if implicit else ("A", "B", "biasA", "biasB")
            report("fit_collective_%s_als" % ("implicit" if implicit else "explicit"), [a[key] for key in keys], [b[key] for key in keys], implicit, dt)
        dist.barrier()
# ---- BASELINE config 4 style: explicit feedback with implicit features (Ai, Bi sharded like A, B; collective.cu)
for dt in (np.dtype(np.float32), np.dtype(np.float64)):
    L = _lib.load(dt)
    m, n, k = 5003, 3001, 32
    ixA, ixB, X = synth_coo(m, n, 200000, dt, seed=6)
    nid = nccl_id_for_all_ranks(L, rank, world)
    assert L.cmfb200_set_world(rank, world, C.cast(nid, C.c_void_p)) == 0
    fit = lambda: fit_explicit(L, dt, ixA, ixB, X, m, n, k, lam=0.05, scale_lam=True, niter=3, finalize_chol=True,
                               add_implicit_features=True, w_implicit=0.5)
    a = fit()
    assert a["rc"] == 0, a["rc"]
    assert L.cmfb200_set_world(0, 1, None) == 0
    if rank == 0:
        b = fit()
        keys = ("A", "B", "biasA", "biasB", "Ai", "Bi")
        # the k x k Grams of A / B / Ai / Bi are summed over the replica in DEVICE row order, which depends on the dealing:
        # rounding-level differences against the 1-GPU fit, like the implicit model's Gram
        report("fit_collective_explicit_als implicit features", [a[key] for key in keys], [b[key] for key in keys], True, dt)
    dist.barrier()
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
