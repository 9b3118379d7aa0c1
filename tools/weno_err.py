"""fp32 WENO stage against the float CPU oracle (development aid): error statistics and niterw on seeded models.
    python tools/weno_err.py [n ...]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import oracle as O  # noqa: E402  (a tool, not the product)
from ttcr_b200 import Grid3d  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [64, 128]:
    rng = np.random.default_rng(100 + n)
    x = np.linspace(0.0, 20.0, n)
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    for name, s in (("smooth", (1 + 0.3 * np.sin(0.7 * X) * np.cos(0.9 * Y)) / (1 + 0.1 * Z)), ("gradient", np.broadcast_to(1 / (1 + 0.1 * Z), X.shape))):
        s = np.ascontiguousarray(s, dtype=np.float32)
        src = np.array([[x[n // 3], x[n // 2], x[n // 5]]])
        g = Grid3d(x, x, x, cell_slowness=0, tt_from_rp=False, weno=1, dtype=np.float32)
        g.raytrace(src, src, s)
        f = g.get_grid_traveltimes()
        dx = float(np.float32(x[1]) - np.float32(x[0]))
        ref, ni, nw = O.solve(n - 1, n - 1, n - 1, dx, O.to_cxx(s), src.astype(np.float32), 0.0, weno=True, dtype=np.float32)
        ref = O.from_cxx(ref, (n, n, n))
        # the reference in DOUBLE on the same (float-valued) model: how far is the reference's own float build from it?
        xd = x.astype(np.float32).astype(np.float64)
        refd, nid, nwd = O.solve(n - 1, n - 1, n - 1, float(xd[1] - xd[0]), O.to_cxx(s.astype(np.float64)), src.astype(np.float32).astype(np.float64), 0.0,
                                 weno=True, dtype=np.float64)
        refd = O.from_cxx(refd, (n, n, n))
        floor = dx * float(s.min())
        def q(a, b):
            e = np.abs(a.astype(np.float64) - b) / np.maximum(b, floor)
            return f"max {e.max():.3g} mean {e.mean():.3g} q99 {np.quantile(e, 0.99):.3g} q99.9 {np.quantile(e, 0.999):.3g}"
        print(f"n={n} {name}: gpu niter {g.get_niter()} float oracle ({ni},{nw}) double oracle ({nid},{nwd})", flush=True)
        print(f"    gpu fp32     vs float  oracle: {q(f, ref)}", flush=True)
        print(f"    gpu fp32     vs double oracle: {q(f, refd)}", flush=True)
        print(f"    float oracle vs double oracle: {q(ref, refd)}", flush=True)
