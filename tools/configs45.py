"""Measurement of BASELINE.json configs[3] and configs[4] as SURVEY section 8(d) states them.  Run plainly (1 GPU) or under
torchrun (one rank per GPU, NCCL).

config 4: 512^3 nodes / 511^3 cells (Grid3Drcfs), cell slowness 1/(1+0.1 z_c) with 5 % lognormal noise
          (default_rng(12345)), 64 sources uniform in [0.5,19.5]^3 from the same generator snapped to the nearest node
          (and the same 64 left off-node), sharded source-parallel over the ranks, two slots per GPU.
config 5: 1024^3 nodes, node slowness 1/(1+0.1 z), fp32, 8 sources at (+-1/4 L) around the centre (block-dealt to the
          ranks), eps in {1e-4, 1e-5, 1e-6}: iterations and Mnodes/s.

usage: configs45.py [4] [5] [--nsrc N] [--n5 N]
"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from ttcr_b200 import Grid3d
from ttcr_b200.distributed import raytrace_sharded, shard_sources

rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
args = sys.argv[1:]
which = [a for a in args if a in ("4", "5")] or ["4", "5"]
nsrc = int(args[args.index("--nsrc") + 1]) if "--nsrc" in args else 64
n5 = int(args[args.index("--n5") + 1]) if "--n5" in args else 1024


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def say(msg):
    if rank == 0:
        print(msg, flush=True)


if "4" in which:
    n = 512
    x = np.linspace(0.0, 20.0, n)
    rng = np.random.default_rng(12345)
    sc = None
    if rank == 0:
        zc = 0.5 * (x[1:] + x[:-1])
        sc = ((1.0 / (1.0 + 0.1 * zc))[None, None, :] * np.exp(0.05 * rng.standard_normal((n - 1, n - 1, n - 1), dtype=np.float32))).astype(np.float32)
    else:
        rng.standard_normal((n - 1, n - 1, n - 1), dtype=np.float32)    # keep the generator in step
    src_off = rng.uniform(0.5, 19.5, (64, 3))[:nsrc]
    dx = x[1] - x[0]
    src_on = np.round(src_off / dx) * dx
    rcv = np.array([[1.0, 1.0, 1.0], [19.0, 19.0, 19.0], [10.0, 10.0, 0.0], [3.3, 16.2, 8.7]])
    g = Grid3d(x, x, x, n_threads=2, cell_slowness=1, tt_from_rp=False, weno=0, dtype=np.float32, device=local)
    # cell model: every rank averages it to nodes itself; rank 0 hands the cells out through the host here (one-off)
    if world > 1:
        t = torch.from_numpy(sc).cuda() if rank == 0 else torch.empty((n - 1,) * 3, dtype=torch.float32, device="cuda")
        dist.broadcast(t, src=0)
        sc = t.cpu().numpy()
        del t
    g.set_slowness(sc)
    for name, src in (("snapped to nodes", src_on), ("off-node", src_off)):
        mine = shard_sources(len(src), world, rank)
        g.raytrace_sources(src[mine[:2]], rcv)      # warm-up
        sync()
        t0 = time.perf_counter()
        tt, its = raytrace_sharded(g, src, rcv)
        sync()
        dt = time.perf_counter() - t0
        sweeps = 8 * int(its[:, 0].sum())
        say(f"config4 [{name}] {world} GPU(s) x 2 slots: {len(src)} sources at 512^3 cells->nodes in {dt:.3f} s, "
            f"{float(n) ** 3 * sweeps / dt / 1e6:.0f} Mnodes/s aggregate, niter min/median/max "
            f"{its[:, 0].min()}/{int(np.median(its[:, 0]))}/{its[:, 0].max()}, tt range {tt.min():.3f}..{tt.max():.3f}")
    g.close()
    del sc

if "5" in which:
    n = n5
    x = np.linspace(0.0, 20.0, n)
    s = None
    if rank == 0:
        s = np.ascontiguousarray(np.broadcast_to((1.0 / (1.0 + 0.1 * x)).astype(np.float32)[None, None, :], (n, n, n)))
    c, q = 10.0, 5.0
    src = np.array([[c + q * (1 if b & 1 else -1), c + q * (1 if b & 2 else -1), c + q * (1 if b & 4 else -1)] for b in range(8)])
    rcv = np.array([[20.0, 20.0, 20.0], [0.0, 0.0, 0.0], [10.0, 3.0, 17.0]])
    for eps in (1e-4, 1e-5, 1e-6):
        g = Grid3d(x, x, x, n_threads=1, cell_slowness=0, tt_from_rp=False, weno=0, eps=eps, dtype=np.float32, device=local)
        sync()
        t0 = time.perf_counter()
        tt, its = raytrace_sharded(g, src, rcv, s)       # includes the broadcast / upload of the model
        sync()
        t1 = time.perf_counter()
        tt, its = raytrace_sharded(g, src, rcv)          # model resident
        sync()
        dt = time.perf_counter() - t1
        sweeps = 8 * int(its[:, 0].sum())
        say(f"config5 eps={eps:g} {world} GPU(s): 8 sources at {n}^3 in {dt:.3f} s (first call with model upload {t1 - t0:.3f} s), "
            f"{float(n) ** 3 * sweeps / dt / 1e6:.0f} Mnodes/s aggregate, niter per source {its[:, 0].tolist()}")
        g.close()

if world > 1:
    dist.barrier()
    dist.destroy_process_group()
