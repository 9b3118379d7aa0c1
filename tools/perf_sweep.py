"""GPU development harness (not a bench line): correctness of the tile kernel against the plane
kernel, then solve timings for a list of sizes / option sets.  Usage:
    python tools/perf_sweep.py check            # bitwise TILE vs PLANE on a few shapes
    python tools/perf_sweep.py time 256 512     # timings
"""
import itertools
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from ttcr_b200 import Grid3d  # noqa: E402

HBM = 6452.2e9


def gradient(n):
    x = np.linspace(0.0, 20.0, n)
    z = x[None, None, :]
    s = np.broadcast_to(1.0 / (1.0 + 0.1 * z), (n, n, n))
    return x, np.ascontiguousarray(s, dtype=np.float32)


def check():
    rng = np.random.default_rng(0)
    ok = True
    for shape, dtype, src in (((40, 33, 70), np.float32, [3.3, 2.2, 9.1]), ((65, 64, 31), np.float32, [0, 0, 0]),
                              ((33, 100, 45), np.float64, [8.0, 20.0, 11.0]), ((129, 128, 130), np.float32, [16.0, 16.0, 16.0]),
                              ((21, 30, 200), np.float32, [2.6, 3.1, 30.2]), ((50, 17, 260), np.float32, [12.25, 4.0, 64.75])):
        x, y, z = (np.arange(m) * 0.25 for m in shape)
        s = rng.uniform(0.3, 1.0, shape)
        res = []
        for kernel, opts in ((1, {}), (2, {}), (2, {"tile_warps": 4, "tile_urows": 2, "tile_rows": 2, "tile_depth": 4}),
                             (3, {}), (3, {"tile_warps": 4, "tile_rows": 2}), (3, {"tile_rows": 16, "ctas_per_sm": 1}),
                             (4, {}), (4, {"tile_warps": 16}), (4, {"tile_depth": 4, "ctas_per_sm": 1}),
                             (5, {"tile_warps": 4}), (5, {"tile_warps": 8}), (5, {"tile_warps": 8, "tile_depth": 16}),
                             (5, {"tile_urows": 2, "ctas_per_sm": 1})):
            g = Grid3d(x, y, z, cell_slowness=0, tt_from_rp=False, weno=0, dtype=dtype)
            g.set_option("kernel", kernel)
            for k, v in opts.items():
                g.set_option(k, v)
            t0 = time.time()
            g.raytrace(np.array([src]), np.array([src]), s)
            st = g.get_stats()
            res.append((g.get_grid_traveltimes(), g.get_niter(), st))
            print(shape, np.dtype(dtype).name, "kernel", kernel, opts, "niter", g.get_niter(), f"solve {st['solve_ms']:.2f} ms",
                  "launches", st["launches"], f"wall {time.time() - t0:.2f}s", flush=True)
        for r in res[1:]:
            same = np.array_equal(r[0], res[0][0]) and r[1] == res[0][1]
            if not same:
                d = np.abs(r[0].astype(np.float64) - res[0][0])
                print("   MISMATCH max", d.max(), "count", np.count_nonzero(d), "niter", r[1], res[0][1])
            ok &= same
    print("CHECK", "OK" if ok else "FAILED")
    return ok


def timing(sizes):
    for n in sizes:
        x, s = gradient(n)
        g = Grid3d(x, x, x, cell_slowness=0, tt_from_rp=False, weno=0, dtype=np.float32)
        g.set_slowness(s)
        print(f"--- {n}^3 device bytes {g.device_bytes() / 2**30:.2f} GiB", flush=True)
        combos = [dict(kernel=5, tile_warps=w, tile_urows=r, tile_depth=dp) for w, r, dp in ((8, 1, 8), (8, 1, 16), (4, 1, 8), (4, 2, 8))]
        combos += [dict(kernel=4, tile_warps=8, tile_depth=4, ctas_per_sm=0)]
        for src in ([0.0, 0.0, 0.0],):
            for c in combos:
                for k, v in c.items():
                    g.set_option(k, v)
                best = None
                for rep in range(3):
                    st = g.solve(np.array([src]))
                    if best is None or st["solve_ms"] < best["solve_ms"]:
                        best = st
                nsw = 8 * (best["niter"] + best["niterw"])
                mn = n ** 3 * nsw / (best["solve_ms"] * 1e-3) / 1e6
                frac = 12.0 * n ** 3 * nsw / (best["solve_ms"] * 1e-3) / HBM
                print(json.dumps(dict(n=n, src=src[0], **c, niter=best["niter"], solve_ms=round(best["solve_ms"], 3),
                                      ms_per_sweep=round(best["solve_ms"] / nsw, 4), mnodes_s=round(mn), hbm_frac=round(frac, 4))), flush=True)


def one(n, opts):
    x, s = gradient(n)
    g = Grid3d(x, x, x, cell_slowness=0, tt_from_rp=False, weno=0, dtype=np.float32)
    g.set_slowness(s)
    g.set_option("kernel", 4)
    for kv in opts:
        k, v = kv.split("=")
        g.set_option(k, float(v))
    import os
    for _ in range(int(os.environ.get("REPS", "1"))):
        st = g.solve(np.array([[0.0, 0.0, 0.0]]))
        print(st)


if __name__ == "__main__":
    if sys.argv[1] == "one":
        one(int(sys.argv[2]), sys.argv[3:])
        sys.exit(0)
    if sys.argv[1] == "check":
        sys.exit(0 if check() else 1)
    timing([int(a) for a in sys.argv[2:]])
