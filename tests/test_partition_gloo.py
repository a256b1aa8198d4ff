"""Host-side logic of the multi-GPU path, on CPU: how rows are dealt to ranks (cmfb200_partition_rows) and that
all-gathering the per-rank blocks reassembles the full factor matrix.  world_size = 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cmfrec_b200 import _lib
from support import csr_csc, ptr, synth_coo


def partition(lib, indptr, rows, world):
    to_dev = np.zeros(rows, np.int32)
    import ctypes as C
    block = C.c_int(0)
    assert lib.cmfb200_partition_rows(ptr(indptr), rows, world, ptr(to_dev), C.byref(block)) == 0
    return to_dev, block.value


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_properties(world):
    dt = np.dtype(np.float32)
    L = _lib.load(dt)
    m, n = 1001, 333
    ixA, ixB, X = synth_coo(m, n, 30000, dt, seed=2)
    csr = csr_csc(L, dt, ixA, ixB, X, m, n)
    deg = np.diff(csr[0].astype(np.int64))
    to_dev, block = partition(L, csr[0], m, world)
    assert block == (m + world - 1) // world
    assert len(set(to_dev.tolist())) == m and to_dev.min() >= 0 and to_dev.max() < block * world   # injective
    if world == 1:
        assert np.array_equal(to_dev, np.arange(m))
        return
    owner = to_dev // block
    per_rank_rows = np.bincount(owner, minlength=world)
    per_rank_nnz = np.bincount(owner, weights=deg, minlength=world)
    assert per_rank_rows.max() - per_rank_rows.min() <= 1
    assert per_rank_nnz.max() <= 1.05 * per_rank_nnz.mean() + deg.max()
    # empty matrix and fewer rows than ranks
    z = np.zeros(3, np.uint64)
    td, b = partition(L, z, 2, world)
    assert b == 1 and len(set(td.tolist())) == 2


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dt = np.dtype(np.float64)
    L = _lib.load(dt)
    m, n, k = 257, 101, 5
    ixA, ixB, X = synth_coo(m, n, 4000, dt, seed=3)
    csr = csr_csc(L, dt, ixA, ixB, X, m, n)
    to_dev, block = partition(L, csr[0], m, world)
    # every rank "solves" only the rows it owns: row r -> f(r) (stand-in for the kernel), in device numbering
    full = torch.zeros(block * world, k, dtype=torch.float64)
    mine = [r for r in range(m) if to_dev[r] // block == rank]
    local = torch.zeros(block, k, dtype=torch.float64)
    for r in mine:
        local[to_dev[r] - rank * block] = torch.arange(k, dtype=torch.float64) + 10.0 * r
    gathered = [torch.zeros(block, k, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, local)                       # what ncclAllGather does in place on the device
    full = torch.cat(gathered)
    back = full[torch.from_numpy(to_dev.astype(np.int64))]  # un-permute to the caller's numbering
    want = torch.arange(k, dtype=torch.float64)[None, :] + 10.0 * torch.arange(m, dtype=torch.float64)[:, None]
    ok = bool(torch.equal(back, want))
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


def test_blocks_reassemble_over_gloo():
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1
