"""CPU tests (gloo, world_size 2) of the source-parallel sharding layer.

The solver behind the sharding layer is replaced by a thin adapter over the CPU oracle (tests may use
oracle/), so the whole multi-rank flow -- broadcast of the model, block partition of the sources, local
solves, all-gather of receiver traveltimes -- is exercised without a GPU and checked against a
single-process run.
"""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT


class OracleGrid:
    """Stand-in with the slice of the Grid3d surface that ttcr_b200.distributed uses."""

    def __init__(self, x, y, z, dtype=np.float64, weno=False):
        import oracle as O
        self.O, self.x, self.y, self.z, self.dtype, self.weno = O, x, y, z, np.dtype(dtype), weno
        self.shape = (x.size, y.size, z.size)
        self._s = None
        self._it = (0, 0)

    def set_slowness(self, s):
        self._s = self.O.to_cxx(np.asarray(s, dtype=self.dtype).reshape(self.shape))

    def raytrace(self, source, rcv, thread_no=0):
        n = self.shape
        dx = float(self.x[1] - self.x[0])
        tt, ni, nw = self.O.solve(n[0] - 1, n[1] - 1, n[2] - 1, dx, self._s, source[:, 1:4], source[:, 0],
                                  float(self.x[0]), float(self.y[0]), float(self.z[0]), weno=self.weno, dtype=self.dtype)
        self._it = (ni, nw)
        return self.O.interp(n[0] - 1, n[1] - 1, n[2] - 1, dx, tt, rcv, float(self.x[0]), float(self.y[0]),
                             float(self.z[0]), dtype=self.dtype)

    def get_niter(self, thread_no=0):
        return self._it


def _case():
    rng = np.random.default_rng(5)
    x = np.arange(13) * 0.5
    y = np.arange(11) * 0.5
    z = np.arange(15) * 0.5
    s = rng.uniform(0.3, 1.0, (13, 11, 15))
    src = np.column_stack([rng.uniform(0, 6, 5), rng.uniform(0, 5, 5), rng.uniform(0, 7, 5)])
    rcv = np.column_stack([rng.uniform(0, 6, 7), rng.uniform(0, 5, 7), rng.uniform(0, 7, 7)])
    return x, y, z, s, src, rcv


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from ttcr_b200.distributed import raytrace_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, y, z, s, src, rcv = _case()
    g = OracleGrid(x, y, z)
    tt, it = raytrace_sharded(g, src, rcv, slowness=s if rank == 0 else None)
    np.save(os.path.join(out, f"tt{rank}.npy"), tt)
    np.save(os.path.join(out, f"it{rank}.npy"), it)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_sources_matches_reference_block_partition():
    from ttcr_b200.distributed import shard_sources
    # Grid3D::get_blk_size (ttcr/Grid3D.h:451-465): min(nThreads, nTx) blocks, sizes dealt round-robin
    for n, w in ((5, 2), (64, 8), (3, 8), (8, 8), (9, 4), (1, 2)):
        seen = np.concatenate([shard_sources(n, w, r) for r in range(w)])
        assert np.array_equal(seen, np.arange(n))
        sizes = [len(shard_sources(n, w, r)) for r in range(w)]
        n_blk = min(w, n)
        ref = [0] * n_blk
        k = n
        while k > 0:
            for b in range(n_blk):
                ref[b] += 1
                k -= 1
                if k == 0:
                    break
        assert sizes[:n_blk] == ref and all(v == 0 for v in sizes[n_blk:])


def test_two_rank_gloo_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from ttcr_b200.distributed import raytrace_sharded
    x, y, z, s, src, rcv = _case()
    tt1, it1 = raytrace_sharded(OracleGrid(x, y, z), src, rcv, slowness=s)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"tt{r}.npy"), tt1)
        assert np.array_equal(np.load(tmp_path / f"it{r}.npy"), it1)
    assert np.all(tt1 > 0) and np.all(it1[:, 0] >= 2)
