"""Development probe: per-step globaltimer stamps of one tile of the first 512^3 sweep, for tools/hop_report.py.
Needs a library built with -DTTCR_T5_STEP_TRACE=1 -DTTCR_T5_CLOCK=gtime (TTCR_B200_LIB) and the environment
TTCR_B200_TRACE, TTCR_B200_TRACE_STEPS, TTCR_B200_TRACE_TILE, TTCR_B200_TRACE_A0 (see sweep_tile5.cuh)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from ttcr_b200 import Grid3d
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x = np.linspace(0.0, 20.0, n)
s = np.ascontiguousarray(np.broadcast_to((1.0 / (1.0 + 0.1 * x))[None, None, :], (n, n, n)), dtype=np.float32)
g = Grid3d(x, x, x, cell_slowness=0, tt_from_rp=False, weno=0, dtype=np.float32)
g.set_slowness(s)
st = g.solve(np.array([[0.0, 0.0, 0.0]]))
print("solve", st["solve_ms"], "ms, niter", st["niter"])
