#!/usr/bin/env python
"""bench.py -- the headline benchmark of the B200 FSM path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n NODES]

Workload (config.workload): BASELINE.json configs[2] -- ttcrpy.rgrid.Grid3d 512^3 nodes, fp32, node
slowness s(z) = 1/(1+0.1 z) on a 20^3 domain (the reference's gradient model, tests/files/mk_models3d.py),
one source at the corner node, first-order FSM (weno=0), eps=1e-5, maxit=50.  One STEP = one complete
solve of one source (reinit + initFSM + 8-direction sweeps until converged + convergence reductions).

metric  Mnodes/s = grid nodes x directional sweeps / solve time / 1e6          
value   inputs (slowness, source) resident in HBM; device time from CUDA events (ttcr_b200_solve)
e2e     same metric through the public API Grid3d.raytrace(src, rcv, slowness): slowness from PINNED
        host memory every step (H2D), receiver traveltimes back to the host (D2H), wall clock
roofline  dominant kernel = the directional-sweep kernel; 12 algorithmic bytes per node per launch
cpu_baseline  the reference's own CPU FSM (oracle/_ref, built from /root/reference) on a bounded sample

With N > 1 (torchrun, one rank per GPU) every rank solves its own source of the same model (weak scaling,
source-parallel).  The device-resident arm has no collective in its timed region (the model is resident); the e2e arm
is raytrace_sharded(): the model sits in pinned memory on rank 0 only, is uploaded once and broadcast over NCCL, the
receiver times are all-gathered -- every step, inside the timed region.
detail.default_arguments (rank 0): the reference's default arguments (fp64, weno=1) at 256^3, weno=1 at 512^3 fp32, and BASELINE.json
configs[4]'s size (1024^3 fp32, one source) with its own roofline fraction.
detail.config4 (every N): BASELINE.json configs[3], 511^3 cells -> Grid3Drcfs averaging, 64 sources sharded over the
ranks through raytrace_sharded (strong scaling; model on rank 0, broadcast + all-gather inside its wall-clock time).
--impl reference times the reference's own CPU implementation on rank 0: 1 thread for 1 source; for --gpus N > 1 the N
sources of the N-GPU run through the reference's own threaded Grid3D::raytrace on min(cores, N) threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mnodes/s (grid nodes x sweep-iters / s) on 512^3 FSM"
UNIT = "Mnodes/s"
BYTES_PER_NODE_SWEEP = 12.0   # tt read + tt write + slowness read, fp32
CPU_SAMPLE_N = 192            # nodes per side of the bounded CPU sample
KERNEL_NAMES = {1: "k_sweep_plane", 2: "k_sweep_tile", 6: "k_sweep_planes_coop", 7: "k_sweep_march"}


def gradient_model(n, dtype=np.float32):
    x = np.linspace(0.0, 20.0, n)
    s = np.broadcast_to((1.0 / (1.0 + 0.1 * x))[None, None, :], (n, n, n))
    return x, np.ascontiguousarray(s, dtype=dtype)


def receivers(x):
    # the reference's rcv.dat pattern: a 21 x 21 lattice on the z = 0 face
    g = np.linspace(x[0], x[-1], 21)
    X, Y = np.meshgrid(g, g, indexing="ij")
    return np.column_stack([X.ravel(), Y.ravel(), np.zeros(X.size)])


def source_for_rank(x, rank):
    if rank == 0:
        return np.array([[0.0, 0.0, 0.0]])
    # config 5 pattern: points +-1/4 L around the centre
    c, q = 0.5 * (x[0] + x[-1]), 0.25 * (x[-1] - x[0])
    b = rank - 1
    return np.array([[c + q * (1 if b & 1 else -1), c + q * (1 if b & 2 else -1), c + q * (1 if b & 4 else -1)]])


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def _lines(self):
        try:
            with open(self.f.name) as f:
                return f.read().splitlines()
        except OSError:
            return []

    def wait_first_sample(self, timeout=8.0):
        """nvidia-smi needs about a second to start; block until it has written a line"""
        t0 = time.perf_counter()
        while self.p is not None and not self._lines() and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def mark(self):
        """the timed region starts here: earlier samples are dropped"""
        self.skip = len(self._lines())

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        lines = self.f.read().splitlines()
        if len(lines) > getattr(self, "skip", 0):
            lines = lines[getattr(self, "skip", 0):]
        for line in lines:
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------------
def cpu_reference(n, steps, warmup, dtype=np.float32):
    """The reference's own CPU FSM on an n^3 sample of the workload.  Returns (Mnodes/s, s/step, kind, niter)."""
    import oracle as O
    x, s = gradient_model(n, dtype)
    xt = x.astype(dtype)
    dx = float(xt[1] - xt[0])
    src = np.array([[0.0, 0.0, 0.0]])
    times, niter = [], 0
    if O.have_ref():
        kind = "reference"
        g = O.RefGrid(n - 1, n - 1, n - 1, dx, weno=False, dtype=dtype, eps=1e-5, maxit=50)
        g.set_slowness(O.to_cxx(s))
        for k in range(warmup + steps):
            _, sec = g.raytrace(src, 0.0, np.zeros((0, 3)))
            if k >= warmup:
                times.append(sec)
        niter = g.niter()[0]
        g.close()
    else:
        kind = "port"
        sf = O.to_cxx(s)
        for k in range(warmup + steps):
            t = time.perf_counter()
            _, niter, _ = O.solve(n - 1, n - 1, n - 1, dx, sf, src, 0.0, weno=False, dtype=dtype)
            if k >= warmup:
                times.append(time.perf_counter() - t)
    sec = float(np.mean(times))
    return n ** 3 * 8 * niter / sec / 1e6, sec, kind, niter


def cpu_reference_multi(n, n_src, steps, warmup, dtype=np.float32):
    """The reference's own threaded multi-source fan-out (Grid3D::raytrace over a vector of sources, ttcr/Grid3D.h:810-853:
    nt = min(n_threads, n_sources) worker threads, one traveltime slot each) on an n^3 sample: the sources bench.py gives
    to ranks 0 .. n_src-1.  Returns (aggregate Mnodes/s, s/step, threads, iterations of source 0)."""
    import oracle as O
    x, s = gradient_model(n, dtype)
    xt = x.astype(dtype)
    dx = float(xt[1] - xt[0])
    srcs = np.vstack([source_for_rank(x, r) for r in range(n_src)])
    nt = max(1, min(os.cpu_count() or 1, n_src))
    g = O.RefGrid(n - 1, n - 1, n - 1, dx, weno=False, dtype=dtype, eps=1e-5, maxit=50, n_threads=nt)
    g.set_slowness(O.to_cxx(s))
    times = []
    for k in range(warmup + steps):
        _, sec = g.raytrace_multi(srcs, 0.0, np.zeros((1, 3)))
        if k >= warmup:
            times.append(sec)
    niter = g.niter()[0]
    g.close()
    sec = float(np.mean(times))
    # (sources away from the corner may need another iteration; slot 0's count is used for all: the same approximation the
    # GPU arm does not need, since it counts its sweeps)
    return n ** 3 * 8 * niter * n_src / sec / 1e6, sec, nt, niter


def run_reference(args, rank):
    if rank != 0:
        return
    import oracle as O
    n = CPU_SAMPLE_N
    t0 = time.perf_counter()
    n_src = max(1, args.gpus)
    if n_src > 1 and O.have_ref():
        v, sec, cores, niter = cpu_reference_multi(n, n_src, args.steps, args.warmup)
        kind = "reference"
        what = (f"{n}^3-node sample of the workload (same model, eps), the {n_src} sources of the {n_src}-GPU run through the "
                f"reference's own threaded Grid3D::raytrace (ttcr/Grid3D.h:810-853) on {cores} threads, Grid3Drnfs<float>, "
                f"{niter} iterations")
    else:
        v, sec, kind, niter = cpu_reference(n, args.steps, args.warmup)
        cores = 1
        what = (f"{n}^3-node sample of the workload (same model, source, eps), Grid3Drnfs<float>, {niter} iterations, "
                f"1 thread: the reference has no intra-source parallelism")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.n, extra={"sample": f"{n}^3", "sources": n_src}),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": what},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n, extra=None):
    c = {"workload": f"ttcrpy.rgrid.Grid3d {n}^3 nodes, linear-gradient node slowness s=1/(1+0.1z), 1 source per GPU "
                     f"(rank 0: corner node), first-order FSM (weno=0), eps=1e-5, fp32 (BASELINE.json configs[2])",
         "nodes": n ** 3, "l2": "inputs larger than L2 (tt + slowness = 1 GiB+ per sweep at 512^3 vs 126 MB L2)",
         "parallelism": "source-parallel, one grid replica per GPU"}
    if extra:
        c.update(extra)
    return c


def config4_side(torch, dist, rank, world, local_rank, n_src):
    """BASELINE.json configs[3] as SURVEY section 8(d) words it: 512^3 nodes / 511^3 cells (Grid3Drcfs), cell slowness
    1/(1+0.1 z_c) with 5 % lognormal noise (default_rng(12345)), `n_src` of the 64 sources uniform in [0.5,19.5]^3 from the
    same generator, sharded source-parallel over the ranks, two slots per GPU.  Wall clock around raytrace_sharded()."""
    from ttcr_b200 import Grid3d
    from ttcr_b200.distributed import raytrace_sharded
    n = 512
    x = np.linspace(0.0, 20.0, n)
    sc = None
    srcs = torch.zeros((64, 3), dtype=torch.float64, device="cuda")
    if rank == 0:
        rng = np.random.default_rng(12345)
        zc = 0.5 * (x[1:] + x[:-1])
        sc = ((1.0 / (1.0 + 0.1 * zc))[None, None, :] * np.exp(0.05 * rng.standard_normal((n - 1, n - 1, n - 1), dtype=np.float32))).astype(np.float32)
        sc = torch.from_numpy(sc).pin_memory().numpy()
        srcs = torch.from_numpy(rng.uniform(0.5, 19.5, (64, 3))).cuda()
    if world > 1:
        dist.broadcast(srcs, src=0)
    src = srcs.cpu().numpy()[:n_src]
    rcv = np.array([[1.0, 1.0, 1.0], [19.0, 19.0, 19.0], [10.0, 10.0, 0.0], [3.3, 16.2, 8.7]])
    g4 = Grid3d(x, x, x, n_threads=2, cell_slowness=1, method="FSM", tt_from_rp=0, eps=1e-5, maxit=50, weno=0,
                dtype=np.float32, device=local_rank)
    raytrace_sharded(g4, src[:2 * world], rcv, slowness=sc)          # warm-up: allocations, first launches
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    tt, it = raytrace_sharded(g4, src, rcv, slowness=sc)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    g4.close()
    sweeps = 8 * int(it.sum())
    return {"workload": f"512^3 nodes / 511^3 cells (Grid3Drcfs), {len(src)} sources sharded over {world} GPU(s), 2 slots per GPU",
            "scaling": "strong", "seconds": dt, "value": float(n) ** 3 * sweeps / dt / 1e6, "unit": UNIT,
            "niter_min_max": [int(it[:, 0].min()), int(it[:, 0].max())],
            "includes": "pinned upload of the 511^3 cells on rank 0, NCCL broadcast, cell-to-node averaging, solves, all-gather"}


def grid2d_side():
    """SURVEY section 8 row f4 side line: the reference's published 2-D GPU table (docs/performance.rst:125-217: homogeneous
    square grid, source at the centre, fp32, default weno=1, minimum of three runs) at 1000 x 1000 cells on ttcr_b200.Grid2d."""
    from ttcr_b200 import Grid2d
    n = 1000
    x = np.arange(n + 1, dtype=np.float64)
    g = Grid2d(x, x, cell_slowness=0, method="FSM", weno=1, dtype=np.float32)
    g.set_slowness(np.ones((n + 1, n + 1), dtype=np.float32))
    src = np.array([[n / 2.0, n / 2.0]])
    rcv = np.array([[1.0, 1.0], [n - 1.0, n - 2.0]])
    g.raytrace(src, rcv)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        g.raytrace(src, rcv)
        best = min(best, time.perf_counter() - t0)
    ni = g.get_niter()
    g.close()
    return {"workload": "Grid2d 1000 x 1000 cells, homogeneous, centre source, fp32, weno=1 (docs/performance.rst:125-217)",
            "seconds": best, "niter": list(ni), "published_cpu_s": 5.105, "published_opencl_gpu_s": 1.381,
            "note": "one CTA per source; independent sources run on different SMs"}


def default_path_side(local_rank):
    """Side lines for the reference's DEFAULT arguments (dtype float64, weno=1), which do not run the headline kernel:
    fp64 first-order (k_sweep_tile) and fp32 / fp64 with the WENO stage (plane kernels) at 256^3, same model and source."""
    from ttcr_b200 import Grid3d
    n = 256
    x, s = gradient_model(n, np.float64)
    src = np.array([[0.0, 0.0, 0.0]])
    out = {"workload": "256^3 gradient model, corner source", "unit": UNIT}
    for name, dtype, weno in (("fp64_first_order", np.float64, 0), ("fp32_weno", np.float32, 1), ("fp64_weno", np.float64, 1)):
        g = Grid3d(x, x, x, cell_slowness=0, tt_from_rp=0, eps=1e-5, maxit=50, weno=weno, dtype=dtype, device=local_rank)
        g.set_slowness(s.astype(dtype))
        g.solve(src)
        st = g.solve(src)
        out[name] = {"solve_ms": st["solve_ms"], "niter": st["niter"], "niterw": st["niterw"],
                     "value": float(n) ** 3 * st["sweeps"] / (st["solve_ms"] * 1e-3) / 1e6}
        g.close()
    # the reference's default path (weno=1) at the HEADLINE size, fp32: first-order stage (k_sweep_march) + WENO stage
    # (k_sweep_march_weno) of the same 512^3 workload; a WENO sweep moves the same 12 algorithmic bytes per node
    n = 512
    x, s = gradient_model(n, np.float32)
    g = Grid3d(x, x, x, cell_slowness=0, tt_from_rp=0, eps=1e-5, maxit=50, weno=1, dtype=np.float32, device=local_rank)
    g.set_slowness(s)
    g.solve(src)
    st = g.solve(src)
    peak, _ = peaks()
    per_sweep_ms = st["sweep_ms"] / max(st["sweeps"], 1)
    out["weno"] = {"workload": "512^3 gradient model, corner source, fp32, weno=1 (first-order stage + WENO stage)",
                   "solve_ms": st["solve_ms"], "niter": st["niter"], "niterw": st["niterw"], "sweeps": st["sweeps"],
                   "value": float(n) ** 3 * st["sweeps"] / (st["solve_ms"] * 1e-3) / 1e6,
                   "avg_sweep_ms": per_sweep_ms,
                   "roofline_frac": BYTES_PER_NODE_SWEEP * float(n) ** 3 / (per_sweep_ms * 1e-3) / 1e9 / peak,
                   "kernel": "k_sweep_march_weno (WENO stage), k_sweep_march (first-order stage)"}
    g.close()
    # BASELINE.json configs[4] size: 1024^3 nodes fp32, one source at -1/4 L of the centre, first order (k_sweep_march4: the
    # marching kernel with four nodes per thread, which the library picks for grids of ~700^3 and more)
    n = 1024
    x, s = gradient_model(n, np.float32)
    g = Grid3d(x, x, x, cell_slowness=0, tt_from_rp=0, eps=1e-5, maxit=50, weno=0, dtype=np.float32, device=local_rank)
    g.set_slowness(s)
    del s
    dxn = float(x[1] - x[0])
    src5 = np.round(np.array([[5.0, 5.0, 5.0]]) / dxn) * dxn
    g.solve(src5)
    st = g.solve(src5)
    per_sweep_ms = st["sweep_ms"] / max(st["sweeps"], 1)
    out["config5_1024"] = {"workload": "1024^3 gradient model, one source on a node near (5, 5, 5), fp32, first order (BASELINE.json configs[4] size)",
                           "solve_ms": st["solve_ms"], "niter": st["niter"], "sweeps": st["sweeps"],
                           "value": float(n) ** 3 * st["sweeps"] / (st["solve_ms"] * 1e-3) / 1e6, "avg_sweep_ms": per_sweep_ms,
                           "roofline_frac": BYTES_PER_NODE_SWEEP * float(n) ** 3 / (per_sweep_ms * 1e-3) / 1e9 / peak,
                           "device_bytes": g.device_bytes(), "kernel": "k_sweep_march4 (four nodes per thread)"}
    g.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=512, help="nodes per side (512 = the headline workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c4-sources", type=int, default=64, help="sources of the configs[3] side measurement (0 = skip)")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (development)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from ttcr_b200 import Grid3d
    from ttcr_b200.distributed import broadcast_slowness

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device (ttcr_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.n
    warm = max(args.warmup, 3) if args.warmup > 0 else 0
    x, s_host = gradient_model(n) if rank == 0 else (np.linspace(0.0, 20.0, n), None)
    g = Grid3d(x, x, x, n_threads=1, cell_slowness=0, method="FSM", tt_from_rp=0, eps=1e-5, maxit=50, weno=0,
               dtype=np.float32, device=local_rank)
    for kv in args.opt:
        k, v = kv.split("=")
        g.set_option(k, float(v))
    # the model: generated on rank 0, broadcast once over NCCL (device to device), outside the timed region
    broadcast_slowness(g, s_host)
    src = source_for_rank(x, rank)
    rcv = receivers(x)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -----------------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    for _ in range(warm):
        g.solve(src)
    clocks.wait_first_sample()
    barrier()
    clocks.mark()
    t_wall = time.perf_counter()
    dev_ms = sweep_ms = 0.0
    sweeps = launches = sweep_launches = 0
    st = None
    for _ in range(args.steps):
        st = g.solve(src)
        dev_ms += st["solve_ms"]; sweep_ms += st["sweep_ms"]
        sweeps += st["sweeps"]; launches += st["launches"]; sweep_launches += st["sweep_launches"]
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    clk = clocks.stop()

    # ---- end-to-end arm: public API, pinned host slowness in, receiver traveltimes out ---------------
    # 1 GPU:  Grid3d.raytrace(src, rcv, slowness) with the model in pinned host memory.
    # N GPUs: raytrace_sharded(grid, sources, rcv, slowness): the model lives on rank 0 only and is uploaded ONCE per step,
    #         broadcast over NCCL (chunk-pipelined with the upload), every rank solves its source, the receiver times are
    #         all-gathered -- all inside the timed region.
    from ttcr_b200.distributed import raytrace_sharded
    s_np = torch.from_numpy(s_host).pin_memory().numpy() if rank == 0 else None
    src4 = np.concatenate([[0.0], src[0]]).reshape(1, 4)
    all_src = np.vstack([source_for_rank(x, r) for r in range(world)])

    def e2e_step():
        if world == 1:
            out = g.raytrace(src4, rcv, s_np)
            return out, g.get_stats()["sweeps"]
        out, it = raytrace_sharded(g, all_src, rcv, slowness=s_np)
        return out, 8 * int(it[rank].sum())

    for _ in range(min(warm, 2)):
        e2e_step()
    barrier()
    t_e2e = time.perf_counter()
    e2e_sweeps = 0
    for _ in range(args.steps):
        tt, sw = e2e_step()
        e2e_sweeps += sw
    barrier()
    e2e_ms = (time.perf_counter() - t_e2e) * 1e3
    s_bytes = n ** 3 * 4
    h2d = int(s_bytes + src4.astype(np.float32).nbytes + rcv.astype(np.float32).nbytes) if rank == 0 else int(
        src4.astype(np.float32).nbytes + rcv.astype(np.float32).nbytes)
    d2h = int(np.asarray(tt).astype(np.float32).nbytes)
    if s_np is None:
        s_np = g.get_slowness()

    # ---- side measurement (1 GPU, not the headline): two independent sources solved concurrently on two slots of one
    # grid (configs[3]'s situation: many sources per GPU); aggregate node-sweeps per second of the pair
    pair = None
    if world == 1 and n <= 512:
        try:
            g2 = Grid3d(x, x, x, n_threads=2, cell_slowness=0, method="FSM", tt_from_rp=0, eps=1e-5, maxit=50, weno=0,
                        dtype=np.float32, device=local_rank)
            g2.set_slowness(s_np)
            two = np.vstack([src[0], source_for_rank(x, 1)[0]])
            g2.raytrace_sources(two, rcv)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            reps = max(2, min(args.steps, 5))
            sw2 = 0
            for _ in range(reps):
                g2.raytrace_sources(two, rcv)
                sw2 += g2.get_stats(0)["sweeps"] + g2.get_stats(1)["sweeps"]
            torch.cuda.synchronize()
            pair = {"sources": 2, "value": float(n) ** 3 * sw2 / (time.perf_counter() - t2) / 1e6, "unit": UNIT,
                    "note": "two slots / CUDA streams of one grid, wall clock around raytrace_sources()"}
            g2.close()
        except Exception as e:
            pair = {"error": str(e)}

    # ---- side measurement: configs[3], strong scaling of 64 sources over the ranks
    c4 = None
    if args.c4_sources > 0 and n == 512:
        try:
            c4 = config4_side(torch, dist, rank, world, local_rank, min(64, args.c4_sources))
        except Exception as e:   # a side line never costs the headline
            c4 = {"error": str(e)}

    g2 = dflt = None
    if rank == 0 and n == 512:
        try:
            g2 = grid2d_side()
        except Exception as e:
            g2 = {"error": str(e)}
        try:
            dflt = default_path_side(local_rank)
        except Exception as e:
            dflt = {"error": str(e)}

    # ---- reduce over ranks: whole-job node-sweeps, max time ------------------------------------------
    nodes = float(n) ** 3
    mine = torch.tensor([nodes * sweeps, dev_ms, wall_ms, nodes * e2e_sweeps, e2e_ms, sweep_ms, float(sweeps),
                         float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tot = mine.clone(); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        mx = mine.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    else:
        tot, mx = mine, mine
    tot, mx = tot.cpu().numpy(), mx.cpu().numpy()
    value = tot[0] / (mx[1] * 1e-3) / 1e6
    e2e_value = tot[3] / (mx[4] * 1e-3) / 1e6

    if rank == 0:
        peak, peak_src = peaks()
        per_launch_ms = sweep_ms / max(sweeps, 1)
        achieved = BYTES_PER_NODE_SWEEP * nodes / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        tj = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tj):
            with open(tj) as f:
                traffic = json.load(f).get(f"{n}")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": mx[1] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(n),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": mx[4] / args.steps,
                    "call": ("Grid3d.raytrace(src, rcv, slowness): slowness from pinned host memory, 441 receiver times back" if world == 1 else
                             "raytrace_sharded(grid, sources, rcv, slowness): one upload on rank 0 + NCCL broadcast + one solve per rank + "
                             "all-gather of the receiver times; h2d/d2h bytes are rank 0's")},
            "gpu_launches": int(tot[7]),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": KERNEL_NAMES.get(st["kernel"], str(st["kernel"])) + " (one launch = one directional sweep)",
                         "algorithmic_bytes_per_launch": BYTES_PER_NODE_SWEEP * nodes,
                         "avg_launch_ms": per_launch_ms, "peak_source": peak_src},
            "clocks": clk,
            "detail": {"niter": st["niter"], "sweeps_per_step": st["sweeps"], "solve_ms_last": st["solve_ms"],
                       "sweep_ms_per_step": sweep_ms / args.steps, "wall_ms_per_step": mx[2] / args.steps,
                       "mnode_iters_per_s": value / 8.0, "kernel": st["kernel"],
                       "device_bytes": g.device_bytes(), "launches_per_step": launches / args.steps,
                       "concurrent_sources": pair, "config4": c4, "grid2d": g2, "default_arguments": dflt},
        }
        if not args.no_cpu_baseline:
            try:
                v, sec, kind, niter = cpu_reference(CPU_SAMPLE_N, 2, 0)
                line["cpu_baseline"] = {
                    "value": v, "unit": UNIT, "cores": 1, "kind": kind,
                    "sample": f"{CPU_SAMPLE_N}^3-node sample of the workload (same model, source, eps), "
                              f"Grid3Drnfs<float>, {niter} iterations, {sec:.1f} s per solve, 1 thread: the reference "
                              f"has no intra-source parallelism"}
            except Exception as e:   # the baseline is a reported number, never a reason to lose the bench line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": str(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
