"""Development probe for profilers: one default (weno=1) fp32 solve of the n^3 gradient model with a small maxit.
    TTCR_B200_WENO_KERNEL=7 ncu ... python tools/weno_probe.py 256 2"""
import sys
import numpy as np
sys.path.insert(0, ".")
from ttcr_b200 import Grid3d

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x = np.linspace(0.0, 20.0, n)
s = np.ascontiguousarray(np.broadcast_to((1.0 / (1.0 + 0.1 * x))[None, None, :], (n, n, n)), dtype=np.float32)
src = np.array([[0.0, 0.0, 0.0]])
g = Grid3d(x, x, x, cell_slowness=0, tt_from_rp=False, weno=1, maxit=maxit, dtype=np.float32)
g.set_slowness(s)
g.raytrace(src, src)
st = g.get_stats()
print(f"n={n} niter {g.get_niter()} solve {st['solve_ms']:.1f} ms sweeps {st['sweep_ms']:.1f} ms kernel {st['kernel']} launches {st['launches']}")
