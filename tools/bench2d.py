"""The reference's published 2-D GPU table (docs/performance.rst:125-217: homogeneous square grids, source at the centre,
single precision, default weno=1, minimum of three runs) on ttcr_b200.Grid2d.  usage: bench2d.py [N ...]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from ttcr_b200 import Grid2d

PUBLISHED = {500: (1.265, 0.650), 1000: (5.105, 1.381), 2000: (20.629, 2.759)}   # N: (CPU Drnfs s, OpenCL GPU Drnfs s)
for n in [int(a) for a in sys.argv[1:]] or [500, 1000, 2000]:
    x = np.arange(n + 1, dtype=np.float64)
    s = np.ones((n + 1, n + 1), dtype=np.float32)
    g = Grid2d(x, x, cell_slowness=0, method="FSM", weno=1, dtype=np.float32)
    g.set_slowness(s)
    src = np.array([[n / 2.0, n / 2.0]])
    rcv = np.array([[1.0, 1.0], [n - 1.0, n - 2.0]])
    g.raytrace(src, rcv)
    best, best_dev = 1e9, 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        tt = g.raytrace(src, rcv)
        best = min(best, time.perf_counter() - t0)
        best_dev = min(best_dev, g.last_solve_ms() * 1e-3)
    pub = PUBLISHED.get(n)
    print(f"{n} x {n} cells: wall {best:.3f} s, device {best_dev:.3f} s, niter {g.get_niter()}, tt {tt.tolist()}"
          + (f"  | published: CPU {pub[0]} s, OpenCL GPU {pub[1]} s -> {pub[1] / best:.1f}x the published GPU time, {pub[0] / best:.0f}x the published CPU time" if pub else ""), flush=True)
