"""Source-parallel sharding of the FSM path over several GPUs (one process per GPU).

The reference's only parallelism is one source per thread with one traveltime slot per thread
(ttcr/Grid3D.h:810-853, block partition ``get_blk_size`` at :451-465).  The multi-GPU analogue is one
rank per GPU, one full grid replica per rank, sources dealt to ranks, and exactly two collectives:

* one broadcast of the slowness model from rank 0 (NCCL: device to device over NVLink, straight into
  ``ttcr_b200_set_slowness_device``; gloo: host tensors, used by the CPU tests), and
* one all-gather of the per-source receiver traveltimes (tiny).

There is no halo exchange and no per-iteration communication: sources are independent.
"""
from __future__ import annotations

import numpy as np


def shard_sources(n_sources: int, world_size: int, rank: int) -> np.ndarray:
    """Indices of the sources rank ``rank`` solves: the reference's block partition
    (Grid3D::get_blk_size, ttcr/Grid3D.h:451-465: blocks of size ceil/floor, larger blocks first)."""
    n_blk = min(world_size, n_sources)
    if rank >= n_blk:
        return np.zeros(0, dtype=np.int64)
    base, extra = divmod(n_sources, n_blk)
    sizes = [base + (1 if r < extra else 0) for r in range(n_blk)]
    start = int(np.sum(sizes[:rank]))
    return np.arange(start, start + sizes[rank], dtype=np.int64)


def broadcast_slowness(grid, slowness, src_rank: int = 0, chunks: int = 8):
    """Give every rank's ``grid`` the slowness model held by ``src_rank`` (``slowness`` may be None elsewhere).

    With the NCCL backend the model is uploaded ONCE (on ``src_rank``; asynchronously when it sits in pinned host memory),
    travels GPU to GPU and is handed to the solver as a device pointer.  Upload and broadcast are pipelined chunk by chunk
    (the broadcast of chunk c runs on NCCL's stream while chunk c+1 is copied), so N ranks pay one host-to-device copy,
    not N concurrent ones.  With gloo the model travels as a host tensor."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        grid.set_slowness(slowness)
        return
    nx, ny, nz = grid.shape
    tdtype = torch.float32 if grid.dtype == np.float32 else torch.float64
    use_cuda = dist.get_backend() == "nccl"
    if not use_cuda:
        if dist.get_rank() == src_rank:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(slowness, dtype=grid.dtype).reshape(nx, ny, nz)))
        else:
            t = torch.empty((nx, ny, nz), dtype=tdtype)
        dist.broadcast(t, src=src_rank)
        grid.set_slowness(t.numpy())
        return
    dev = torch.device("cuda", torch.cuda.current_device())
    t = getattr(grid, "_bcast_buf", None)
    if t is None or t.numel() != nx * ny * nz or t.dtype != tdtype or t.device != dev:
        t = torch.empty(nx * ny * nz, dtype=tdtype, device=dev)
        grid._bcast_buf = t          # landing buffer, kept for the next model
    src = None
    if dist.get_rank() == src_rank:
        src = torch.from_numpy(np.ascontiguousarray(np.asarray(slowness, dtype=grid.dtype)).reshape(-1))
    n = t.numel()
    # chunks of whole x planes: copies (rank `src_rank`) and broadcasts of all chunks are queued on a side stream up front;
    # the host then follows the broadcasts chunk by chunk and hands each chunk to the solver's import while the next
    # ones are still on the bus / the links.  (Cell models are averaged as a whole: one import at the end.)
    plane = ny * nz
    piecewise = not getattr(grid, "cell_slowness", False) and nx >= 4 * chunks
    bounds = [nx * c // chunks for c in range(chunks + 1)] if piecewise else [0, nx]
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    works = []
    with torch.cuda.stream(side):
        for a, b in zip(bounds[:-1], bounds[1:]):
            if src is not None:
                t[a * plane:b * plane].copy_(src[a * plane:b * plane], non_blocking=True)
            works.append(dist.broadcast(t[a * plane:b * plane], src=src_rank, async_op=True))
    for (a, b), w in zip(zip(bounds[:-1], bounds[1:]), works):
        w.wait()                                    # the current stream waits for this chunk's broadcast ...
        ev = torch.cuda.Event()
        ev.record()
        ev.synchronize()                            # ... and so does the host
        if piecewise:
            grid.set_slowness_device_planes(t.data_ptr(), n, a, b - a)
    if not piecewise:
        grid.set_slowness_device(t.data_ptr(), n)


def raytrace_sharded(grid, sources, rcv, slowness=None, t0=None):
    """Solve ``sources`` (n x 3) against the same receivers ``rcv`` (m x 3), sharded over the ranks.

    Every rank returns the full ``(n, m)`` array of receiver traveltimes and the ``(n, 2)`` array of
    (niter, niterw).  ``slowness`` is read on rank 0 only."""
    import torch
    import torch.distributed as dist

    sources = np.asarray(sources, dtype=np.float64).reshape(-1, 3)
    rcv = np.asarray(rcv, dtype=np.float64).reshape(-1, 3)
    n, m = sources.shape[0], rcv.shape[0]
    t0 = np.zeros(n) if t0 is None else np.asarray(t0, dtype=np.float64)
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    world = dist.get_world_size() if distributed else 1
    rank = dist.get_rank() if distributed else 0
    if distributed:
        # rank 0 decides whether a model travels (None there = the model of an earlier call is still resident everywhere)
        use_cuda = dist.get_backend() == "nccl"
        flag = torch.tensor([0 if slowness is None else 1], dtype=torch.int32,
                            device=torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu"))
        dist.broadcast(flag, src=0)
        if int(flag.item()):
            broadcast_slowness(grid, slowness)
    elif slowness is not None:
        broadcast_slowness(grid, slowness)
    mine = shard_sources(n, world, rank)
    tt_local = np.zeros((len(mine), m))
    it_local = np.zeros((len(mine), 2), dtype=np.int64)
    if hasattr(grid, "raytrace_sources") and len(mine):
        # the rank's sources go to the slots of its grid (n_threads of them, one CUDA stream each): with two slots the
        # solves overlap on the device (measured 1.34x aggregate at 512^3)
        tt_local, it_local = grid.raytrace_sources(sources[mine], rcv, t0[mine])
        tt_local = np.asarray(tt_local, dtype=np.float64)
    else:
        for a, s in enumerate(mine):
            src = np.concatenate([[t0[s]], sources[s]]).reshape(1, 4)
            tt_local[a] = grid.raytrace(src, rcv, thread_no=0)
            it_local[a] = grid.get_niter(0)
    if not distributed:
        return tt_local, it_local
    # all-gather with padding to the largest shard (shards differ by at most one source)
    cap = int(np.ceil(n / min(world, n)))
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    buf = torch.zeros((cap, m + 2), dtype=torch.float64, device=dev)
    if len(mine):
        buf[:len(mine), :m] = torch.from_numpy(tt_local).to(dev)
        buf[:len(mine), m:] = torch.from_numpy(it_local.astype(np.float64)).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    tt = np.zeros((n, m))
    it = np.zeros((n, 2), dtype=np.int64)
    for r in range(world):
        idx = shard_sources(n, world, r)
        if len(idx):
            o = out[r].cpu().numpy()
            tt[idx] = o[:len(idx), :m]
            it[idx] = o[:len(idx), m:].astype(np.int64)
    return tt, it
