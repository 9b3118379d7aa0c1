"""Debug helper: repeat full solves with the patch kernel on one shape, compare with the plane kernel."""
import sys
import numpy as np
sys.path.insert(0, ".")
from ttcr_b200 import Grid3d  # noqa: E402
ni, nj, nk = (int(a) for a in sys.argv[1:4])
opts = dict(kv.split("=") for kv in sys.argv[4:])
reps = int(opts.pop("reps", 5))
rng = np.random.default_rng(0)
x, y, z = (np.arange(m) * 0.25 for m in (ni, nj, nk))
s = rng.uniform(0.3, 1.0, (ni, nj, nk))
src = np.array([[12.25, 4.0, 64.75]]) if (ni, nj, nk) == (50, 17, 260) else np.array([[x[ni // 3] + 0.1, y[nj // 2], z[nk // 4] + 0.05]])
g = Grid3d(x, y, z, cell_slowness=0, tt_from_rp=False, weno=0, dtype=np.float32)
g.set_option("kernel", 1)
g.raytrace(src, src, s)
ref = g.get_grid_traveltimes()
for rep in range(reps):
    g = Grid3d(x, y, z, cell_slowness=0, tt_from_rp=False, weno=0, dtype=np.float32)
    g.set_option("kernel", 5)
    for k, v in opts.items():
        g.set_option(k, float(v))
    try:
        g.raytrace(src, src, s)
        print(rep, "niter", g.get_niter(), "equal", np.array_equal(g.get_grid_traveltimes(), ref), flush=True)
    except Exception as e:  # noqa: BLE001
        print(rep, "FAILED", str(e)[:200], flush=True)
        break
