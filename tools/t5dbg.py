"""Debug helper: one small solve with the patch kernel (kernel 5) compared bitwise with the plane kernel.
    python tools/t5dbg.py NI NJ NK [tile_warps] [dirs]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from ttcr_b200 import Grid3d  # noqa: E402

ni, nj, nk = (int(a) for a in sys.argv[1:4])
warps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
rng = np.random.default_rng(0)
x, y, z = (np.arange(m) * 0.25 for m in (ni, nj, nk))
s = rng.uniform(0.3, 1.0, (ni, nj, nk))
src = np.array([[x[ni // 3] + 0.1, y[nj // 2], z[nk // 4] + 0.05]])
res = []
for kernel in (1, 5):
    g = Grid3d(x, y, z, cell_slowness=0, tt_from_rp=False, weno=0, dtype=np.float32)
    g.set_option("kernel", kernel)
    g.set_option("tile_warps", warps)
    if len(sys.argv) > 5:
        g.set_option("maxit", int(sys.argv[5]))
    g.raytrace(src, src, s)
    res.append((g.get_grid_traveltimes(), g.get_niter()))
    print("kernel", kernel, "niter", g.get_niter(), g.get_stats(), flush=True)
a, b = res[0][0], res[1][0]
d = np.abs(a.astype(np.float64) - b)
print("max diff", d.max(), "count", np.count_nonzero(d), "of", d.size)
if np.count_nonzero(d):
    idx = np.argwhere(d > 0)
    print("first mismatches (i,j,k):", idx[:10].tolist())
    print("bbox", idx.min(0), idx.max(0))
    for i, j, k in idx[:5]:
        print((i, j, k), a[i, j, k], b[i, j, k])
