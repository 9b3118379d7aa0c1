"""Development probe: time of a default (weno=1) solve, whose WENO stage runs plane-per-launch, with and without the
captured CUDA graph / programmatic dependent launch.  usage: weno_time.py N [N ...]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from ttcr_b200 import Grid3d

fields = {}
for n in [int(a) for a in sys.argv[1:]] or [256]:
    x = np.linspace(0.0, 20.0, n)
    s = np.ascontiguousarray(np.broadcast_to((1.0 / (1.0 + 0.1 * x))[None, None, :], (n, n, n)), dtype=np.float32)
    src = np.array([[0.0, 0.0, 0.0]])
    variants = ((1, 1, 7, 0), (1, 1, 1, 0), (0, 0, 6, 4)) if n <= 256 else ((1, 1, 7, 0), (0, 0, 6, 4))
    for graph, pdl, wk, cc in variants:
        g = Grid3d(x, x, x, cell_slowness=0, tt_from_rp=False, weno=1, dtype=np.float32)
        g.set_option("plane_graph", graph)
        g.set_option("plane_pdl", pdl)
        g.set_option("weno_kernel", wk)
        if cc:
            g.set_option("coop_ctas", cc)
        g.set_slowness(s)
        g.raytrace(src, src)            # warm-up (captures the graphs)
        t = time.perf_counter()
        g.raytrace(src, src)
        dt = time.perf_counter() - t
        st = g.get_stats()
        f = g.get_grid_traveltimes()
        same = fields.setdefault(n, f) is f or np.array_equal(fields[n], f)
        print(f"n={n} graph={graph} pdl={pdl} weno_kernel={wk} coop_ctas={cc}: wall {dt*1e3:.1f} ms, solve {st['solve_ms']:.1f} ms, sweeps {st['sweep_ms']:.1f} ms, "
              f"niter {g.get_niter()}, launches {st['launches']}, {st['sweep_ms']*1e3/max(1, st['sweep_launches']):.2f} us/launch, "
              f"field identical to first variant: {same}", flush=True)
        g.close()
