"""One-process-per-GPU driver for the device-resident ALS state (include/cmfrec_b200.h PART 2).

torch.distributed is used only as plumbing: to agree on the NCCL unique id the library's own communicator is
created from, and for barriers / max-over-ranks timing.  The data path collective (all-gather of the freshly solved
factor block after every half-sweep) is issued by the library itself on its stream.
"""
import ctypes as C

import numpy as np

from . import _lib


def nccl_id_for_all_ranks(lib, rank, world, device=None):
    """128-byte NCCL unique id created on rank 0 and broadcast with torch.distributed (any backend)."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = (C.c_ubyte * 128)()
        if lib.cmfb200_nccl_unique_id(raw) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
        buf = torch.tensor(list(raw), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        buf = buf.cuda() if device is None else buf.to(device)
    dist.broadcast(buf, 0)
    return (C.c_ubyte * 128)(*buf.cpu().tolist())


class ShardedAls:
    """ALS state of one rank.  `csr` / `csc` are the full matrices (every rank passes the same ones); the state
    keeps this rank's block of rows of each orientation and a full replica of both factor matrices."""

    def __init__(self, dtype, csr, csc, m, n, k, *, implicit, rank=0, world=1, nccl_id=None, stream=None, user_bias=False,
                 item_bias=False, lam_A=0.0, lam_B=0.0, lam_biasA=None, lam_biasB=None, scale_lam=False, max_cg_steps=3):
        self.dt = np.dtype(dtype)
        self.lib = lib = _lib.load(self.dt)
        self.m, self.n, self.k = m, n, k
        opt = lib.AlsOptions()
        opt.implicit = int(implicit)
        opt.m, opt.n, opt.k = m, n, k
        opt.user_bias, opt.item_bias = int(user_bias), int(item_bias)
        opt.lam_A, opt.lam_B = lam_A, lam_B
        opt.lam_biasA = lam_A if lam_biasA is None else lam_biasA
        opt.lam_biasB = lam_B if lam_biasB is None else lam_biasB
        opt.scale_lam, opt.max_cg_steps = int(scale_lam), max_cg_steps
        opt.rank, opt.world = rank, world
        self._id = nccl_id
        opt.nccl_id = C.cast(nccl_id, C.c_void_p) if nccl_id is not None else None
        opt.stream = stream
        self.h = C.c_void_p()
        rc = lib.cmfb200_als_create(C.byref(self.h), C.byref(opt), *[_lib.ptr(t) for t in csr], *[_lib.ptr(t) for t in csc])
        if rc:
            raise RuntimeError("cmfb200_als_create failed with code %d" % rc)

    def close(self):
        if self.h:
            self.lib.cmfb200_als_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_factors(self, A, biasA, B, biasB):
        rc = self.lib.cmfb200_als_set_factors(self.h, _lib.ptr(A), _lib.ptr(biasA), _lib.ptr(B), _lib.ptr(biasB))
        if rc:
            raise RuntimeError("set_factors -> %d" % rc)

    def get_factors(self):
        A = np.zeros((self.m, self.k), self.dt); B = np.zeros((self.n, self.k), self.dt)
        bA = np.zeros(self.m, self.dt); bB = np.zeros(self.n, self.dt)
        rc = self.lib.cmfb200_als_get_factors(self.h, _lib.ptr(A), _lib.ptr(bA), _lib.ptr(B), _lib.ptr(bB))
        if rc:
            raise RuntimeError("get_factors -> %d" % rc)
        return A, bA, B, bB

    def iterate(self, first_iter, n_iters, niter_total, use_cg=True, finalize_chol=False):
        ms = C.c_float(0)
        rc = self.lib.cmfb200_als_timed_iterate(self.h, first_iter, n_iters, niter_total, int(use_cg), int(finalize_chol),
                                                C.byref(ms))
        if rc:
            raise RuntimeError("iterate -> %d" % rc)
        return ms.value
