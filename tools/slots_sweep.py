"""configs[3]-like workload (512^3 gradient model, 16 random sources) with 1 .. 4 slots (CUDA streams) of one grid: aggregate
Mnodes/s of raytrace_sources (development aid)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from ttcr_b200 import Grid3d

n = 512
x = np.linspace(0.0, 20.0, n)
s = np.ascontiguousarray(np.broadcast_to((1.0 / (1.0 + 0.1 * x))[None, None, :], (n, n, n)), dtype=np.float32)
rng = np.random.default_rng(1)
src = rng.uniform(0.5, 19.5, (16, 3))
rcv = np.array([[1.0, 1.0, 1.0], [19.0, 19.0, 19.0]])
for slots in (1, 2, 3, 4):
    g = Grid3d(x, x, x, n_threads=slots, cell_slowness=0, tt_from_rp=False, weno=0, dtype=np.float32)
    g.set_slowness(s)
    g.raytrace_sources(src[:slots], rcv)
    t0 = time.perf_counter()
    tt, it = g.raytrace_sources(src, rcv)
    dt = time.perf_counter() - t0
    print(f"slots {slots}: {dt*1e3:.1f} ms for 16 sources, {n**3 * 8 * it[:, 0].sum() / dt / 1e6:.0f} Mnodes/s aggregate, niter {sorted(set(it[:, 0].tolist()))}", flush=True)
    g.close()
