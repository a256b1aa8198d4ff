#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference itself (oracle/_ref, built from /root/reference/src by
oracle/Makefile).  Run in the authoring container, where the reference exists; the fixtures are committed so
that the oracle restatement and the GPU path can be checked against reference outputs on machines that have
neither /root/reference nor oracle/_ref.

Fixtures (small on purpose, a few hundred KB in total):
  init_{f32,f64}.npz        random_parallel outputs (normal + uniform, incl. element > 2^18), seeds fixed
  prep_{f32,f64}.npz        COO -> CSR/CSC, global mean (nthreads 1 and 8), two-sided bias initialisation
  sweep_{f32,f64}.npz       one optimizeA / optimizeA_implicit call each for CG and Cholesky
  fit_{f32,f64}.npz         whole fits: explicit (biases, centre, CG + final Cholesky) and implicit (CG), 2 iterations
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refload import ArraysToFill, ptr, ref  # noqa: E402
from support import fit_explicit, fit_implicit, ref_optimizeA, ref_optimizeA_implicit, synth_coo  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def main():
    for dt in (np.dtype(np.float32), np.dtype(np.float64)):
        tag = "f32" if dt == np.float32 else "f64"
        R = ref(dt)
        assert R is not None, "build oracle/_ref first (make -C oracle ref)"
        # ---- init
        d = {}
        for name, (sa, sb, normal, seed) in dict(small_normal=(3000, 0, True, 1), small_uniform_req=(3000, 0, False, 7),
                                                 big_normal=(300001, 1003, True, 123),
                                                 big_uniform=(300001, 1003, False, 123)).items():
            A = np.zeros(sa, dt); B = np.zeros(max(sb, 1), dt)
            R.random_parallel(ArraysToFill(ptr(A), sa, ptr(B) if sb else None, sb), seed, normal, 4)
            # keep heads, tails and a strided sample: enough to pin the stream without committing megabytes
            sel = np.unique(np.concatenate([np.arange(min(sa, 256)), np.arange(max(sa - 256, 0), sa), np.arange(0, sa, 997)]))
            d[name + "_args"] = np.array([sa, sb, int(normal), seed])
            d[name + "_selA"] = sel; d[name + "_A"] = A[sel]; d[name + "_B"] = B[:sb][:256]
            d[name + "_sumA"] = np.array([A.astype(np.float64).sum(), np.abs(A.astype(np.float64)).sum()])
        np.savez_compressed(os.path.join(OUT, "init_%s.npz" % tag), **d)
        # ---- prep
        m, n, nnz = 500, 320, 6000
        ixA, ixB, X = synth_coo(m, n, nnz, dt, seed=5)
        nz = X.size
        csr = (np.zeros(m + 1, np.uint64), np.zeros(nz, np.int32), np.zeros(nz, dt), np.zeros(n + 1, np.uint64),
               np.zeros(nz, np.int32), np.zeros(nz, dt))
        # unsorted copy of the COO so that "order of appearance" is exercised
        perm = np.random.default_rng(0).permutation(nz)
        ia, ib, xx = ixA[perm].copy(), ixB[perm].copy(), X[perm].copy()
        R.coo_to_csr_and_csc(ptr(ia), ptr(ib), ptr(xx), None, m, n, nz, *[ptr(t) for t in csr], None, None, 2)
        d = dict(m=m, n=n, ixA=ia, ixB=ib, X=xx, csr_p=csr[0], csr_i=csr[1], csr_v=csr[2], csc_p=csr[3], csc_i=csr[4],
                 csc_v=csr[5])
        for nt in (1, 8):
            g = np.zeros(1, dt); f1 = C.c_bool(False); f2 = C.c_bool(False); xp = C.c_void_p(xx.ctypes.data)
            R.calc_mean_and_center(ptr(ia), ptr(ib), C.byref(xp), nz, None, None, m, n, None, None, None, None, None, None,
                                   None, False, False, True, nt, ptr(g), C.byref(f1), C.byref(f2), False)
            d["mean_nt%d" % nt] = g
        xc = (xx - d["mean_nt1"][0]).astype(dt)
        csrc = tuple(np.zeros_like(t) for t in csr)
        R.coo_to_csr_and_csc(ptr(ia), ptr(ib), ptr(xc), None, m, n, nz, *[ptr(t) for t in csrc], None, None, 2)
        for scale in (0, 1):
            bA = np.zeros(m, dt); bB = np.zeros(n, dt)
            R.initialize_biases_twosided(None, None, None, None, m, n, False, False, float(d["mean_nt1"][0]),
                                         *[ptr(t) for t in csrc], None, None, None, None, 0.05, 0.07, bool(scale), None,
                                         None, ptr(bA), ptr(bB), 2)
            d["biasA_scale%d" % scale] = bA; d["biasB_scale%d" % scale] = bB
        np.savez_compressed(os.path.join(OUT, "prep_%s.npz" % tag), **d)
        # ---- single half-sweeps
        k = 12
        rng = np.random.default_rng(3)
        A0 = (rng.normal(size=(m, k)) * 0.1).astype(dt); B0 = (rng.normal(size=(n, k)) * 0.1).astype(dt)
        d = dict(A0=A0, B0=B0, lam=0.7, lam_last=1.3)
        for solver in ("cg", "chol"):
            for scale in (0, 1):
                A1 = A0.copy()
                ref_optimizeA(R, dt, A1, B0.copy(), csrc[0], csrc[1], csrc[2], lam=0.7 if not scale else 0.05,
                              lam_last=1.3 if not scale else 0.09, scale_lam=bool(scale), use_cg=solver == "cg", max_cg_steps=3)
                d["explicit_%s_scale%d" % (solver, scale)] = A1
            A1 = np.abs(A0).copy()
            ref_optimizeA_implicit(R, dt, A1, np.abs(B0).copy(), csr[0], csr[1], csr[2], lam=2.0, use_cg=solver == "cg",
                                   max_cg_steps=3)
            d["implicit_%s" % solver] = A1
        np.savez_compressed(os.path.join(OUT, "sweep_%s.npz" % tag), **d)
        # ---- whole fits
        m2, n2, k2 = 800, 450, 10
        ea, eb, ex = synth_coo(m2, n2, 12000, dt, seed=21)
        fe = fit_explicit(R, dt, ea, eb, ex, m2, n2, k2, lam=0.8, niter=2, finalize_chol=True, nthreads=2)
        m3, n3 = 20000, 9000
        ia3, ib3, x3 = synth_coo(m3, n3, 60000, dt, seed=22, kind="counts")
        x3 = np.minimum(x3, 20).astype(dt)
        fi = fit_implicit(R, dt, ia3, ib3, x3, m3, n3, 16, lam=4.0, niter=2, use_cg=False, nthreads=2)
        sel = np.arange(0, m3, 97)
        np.savez_compressed(os.path.join(OUT, "fit_%s.npz" % tag), e_ixA=ea, e_ixB=eb, e_X=ex, e_A=fe["A"], e_B=fe["B"],
                            e_biasA=fe["biasA"], e_biasB=fe["biasB"], e_glob_mean=fe["glob_mean"],
                            i_args=np.array([m3, n3, 60000, 22]), i_selA=sel, i_A=fi["A"][sel], i_B=fi["B"][::47])
    sizes = {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))}
    print(sizes, "total %.1f KB" % (sum(sizes.values()) / 1024))


if __name__ == "__main__":
    main()
