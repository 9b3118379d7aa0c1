"""GPU development harness for k_sweep_march (not a bench line).
    python tools/march_check.py check           # bitwise MARCH vs PLANE on a few shapes (fp32), both tile shapes
    python tools/march_check.py time 256 512    # sweep timings, MARCH next to TILE5
Every wait in the kernel is bounded (spin_limit); the script sets a small limit so that a protocol bug costs seconds.
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from ttcr_b200 import Grid3d  # noqa: E402

HBM = 6542.1e9
MARCH, TILE5, PLANE = 7, 5, 1


def gradient(n):
    x = np.linspace(0.0, 20.0, n)
    z = x[None, None, :]
    s = np.broadcast_to(1.0 / (1.0 + 0.1 * z), (n, n, n))
    return x, np.ascontiguousarray(s, dtype=np.float32)


def check(quick=False):
    rng = np.random.default_rng(0)
    ok = True
    shapes = [((40, 33, 70), [3.3, 2.2, 9.1]), ((65, 64, 31), [0, 0, 0]), ((33, 100, 45), [8.0, 20.0, 11.0]),
              ((129, 128, 130), [16.0, 16.0, 16.0]), ((21, 30, 200), [2.6, 3.1, 30.2]), ((50, 17, 260), [12.25, 4.0, 64.75]),
              ((16, 16, 16), [3.75, 3.75, 3.75]), ((5, 3, 2), [0.5, 0.25, 0.1])]
    if quick:
        shapes = shapes[:2]
    for shape, src in shapes:
        x, y, z = (np.arange(m) * 0.25 for m in shape)
        s = rng.uniform(0.3, 1.0, shape)
        res = []
        for kernel, opts in ((PLANE, {}), (MARCH, {"march_nodes": 4}), (MARCH, {"march_nodes": 4, "max_ctas": 3}), (MARCH, {"march_nodes": 2}), (MARCH, {"march_nodes": 2, "tile_warps": 12}), (MARCH, {"march_nodes": 2, "tile_depth": 3}), (MARCH, {"march_nodes": 2, "max_ctas": 3})):
            g = Grid3d(x, y, z, cell_slowness=0, tt_from_rp=False, weno=0, dtype=np.float32)
            g.set_option("kernel", kernel)
            g.set_option("spin_limit", 1 << 16)   # x 512 cycles = 17 ms per wait
            for k, v in opts.items():
                g.set_option(k, v)
            t0 = time.time()
            try:
                g.raytrace(np.array([src]), np.array([src]), s)
            except Exception as e:  # noqa: BLE001
                print(shape, "kernel", kernel, opts, "FAILED:", e, flush=True)
                ok = False
                res.append(None)
                continue
            st = g.get_stats()
            res.append((g.get_grid_traveltimes(), g.get_niter(), st))
            print(shape, "kernel", kernel, opts, "niter", g.get_niter(), f"solve {st['solve_ms']:.2f} ms", "launches", st["launches"],
                  f"wall {time.time() - t0:.2f}s", flush=True)
        for r in res[1:]:
            if r is None or res[0] is None:
                continue
            same = np.array_equal(r[0], res[0][0]) and r[1] == res[0][1]
            if not same:
                d = np.abs(r[0].astype(np.float64) - res[0][0])
                bad = np.argwhere(d > 0)
                print("   MISMATCH max", d.max(), "count", np.count_nonzero(d), "niter", r[1], res[0][1], "first", bad[:4].tolist(), flush=True)
            ok &= same
    print("CHECK", "OK" if ok else "FAILED", flush=True)
    return ok


def wcheck(quick=False):
    """WENO stage: the marching kernel's WENO variant against the plane kernels, bit for bit (fp32, weno=1)"""
    rng = np.random.default_rng(0)
    ok = True
    shapes = [((40, 33, 70), [3.3, 2.2, 9.1]), ((65, 64, 31), [0, 0, 0]), ((33, 100, 45), [8.0, 20.0, 11.0]),
              ((129, 128, 130), [16.0, 16.0, 16.0]), ((21, 30, 200), [2.6, 3.1, 30.2]), ((16, 16, 16), [3.75, 3.75, 3.75]), ((7, 6, 5), [0.5, 0.25, 0.1])]
    if quick:
        shapes = shapes[:2]
    for shape, src in shapes:
        x, y, z = (np.arange(m) * 0.25 for m in shape)
        X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
        s = (1 + 0.3 * np.sin(0.7 * X) * np.cos(0.9 * Y)) / (1 + 0.1 * Z) + 0.02 * rng.uniform(0, 1, shape)
        res = []
        for opts in ({"weno_kernel": PLANE}, {"weno_kernel": MARCH}, {"weno_kernel": MARCH, "max_ctas": 3}):
            g = Grid3d(x, y, z, cell_slowness=0, tt_from_rp=False, weno=1, maxit=8, dtype=np.float32)
            g.set_option("spin_limit", 1 << 16)
            for k, v in opts.items():
                g.set_option(k, v)
            try:
                g.raytrace(np.array([src]), np.array([src]), s)
            except Exception as e:  # noqa: BLE001
                print(shape, opts, "FAILED:", e, flush=True)
                ok = False
                res.append(None)
                continue
            st = g.get_stats()
            res.append((g.get_grid_traveltimes(), g.get_niter()))
            print(shape, opts, "niter", g.get_niter(), f"solve {st['solve_ms']:.2f} ms", flush=True)
        for r in res[1:]:
            if r is None or res[0] is None:
                continue
            same = np.array_equal(r[0], res[0][0]) and r[1] == res[0][1]
            if not same:
                d = np.abs(r[0].astype(np.float64) - res[0][0])
                bad = np.argwhere(d > 0)
                print("   MISMATCH max", d.max(), "count", np.count_nonzero(d), "of", d.size, "niter", r[1], res[0][1], "first", bad[:6].tolist(), flush=True)
            ok &= same
    print("WCHECK", "OK" if ok else "FAILED", flush=True)
    return ok


def timing(sizes, combos=None):
    for n in sizes:
        x, s = gradient(n)
        g = Grid3d(x, x, x, cell_slowness=0, tt_from_rp=False, weno=0, dtype=np.float32)
        g.set_slowness(s)
        print(f"--- {n}^3 device bytes {g.device_bytes() / 2**30:.2f} GiB", flush=True)
        combos_n = combos or [dict(kernel=MARCH), dict(kernel=MARCH, march_nodes=4), dict(kernel=MARCH, march_nodes=2), dict(kernel=MARCH, march_nodes=2, tile_warps=12)]   # (the first: the library's own choice)
        for src in ([0.0, 0.0, 0.0],):
            for c in combos_n:
                g.set_option("tile_warps", 0); g.set_option("tile_urows", 1); g.set_option("ctas_per_sm", 0); g.set_option("tile_depth", 8); g.set_option("march_nodes", 0)
                for k, v in c.items():
                    g.set_option(k, v)
                best = None
                try:
                    for rep in range(3):
                        st = g.solve(np.array([src]))
                        if best is None or st["solve_ms"] < best["solve_ms"]:
                            best = st
                except Exception as e:  # noqa: BLE001
                    print(n, c, "FAILED:", e, flush=True)
                    continue
                nsw = 8 * (best["niter"] + best["niterw"])
                mn = n ** 3 * nsw / (best["solve_ms"] * 1e-3) / 1e6
                print(json.dumps(dict(n=n, **c, niter=best["niter"], solve_ms=round(best["solve_ms"], 3), sweep_ms=round(best["sweep_ms"] / nsw, 4),
                                      mnodes_s=round(mn), hbm_frac_sweep=round(12.0 * n ** 3 / (best["sweep_ms"] / nsw * 1e-3) / HBM, 4))), flush=True)


def one(shape, opts):
    """solve a gradient model of the given shape with kernel MARCH; REPS env = repetitions"""
    import os
    ni, nj, nk = shape
    x, y, z = (np.arange(m) * 0.25 for m in shape)
    s = np.ascontiguousarray(np.broadcast_to(1.0 / (1.0 + 0.1 * z[None, None, :]), shape), dtype=np.float32)
    g = Grid3d(x, y, z, cell_slowness=0, tt_from_rp=False, weno=0, dtype=np.float32)
    g.set_slowness(s)
    g.set_option("kernel", MARCH)
    for kv in opts:
        k, v = kv.split("=")
        g.set_option(k, float(v))
    for _ in range(int(os.environ.get("REPS", "2"))):
        st = g.solve(np.array([[0.0, 0.0, 0.0]]))
        nsw = st["sweeps"]
        print(shape, opts, "niter", st["niter"], f"solve {st['solve_ms']:.3f} ms, per sweep {st['sweep_ms'] / nsw * 1e3:.1f} us", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "one":
        one(tuple(int(a) for a in sys.argv[2].split("x")), sys.argv[3:])
        sys.exit(0)
    if sys.argv[1] == "wcheck":
        sys.exit(0 if wcheck(len(sys.argv) > 2 and sys.argv[2] == "quick") else 1)
    if sys.argv[1] == "check":
        sys.exit(0 if check(len(sys.argv) > 2 and sys.argv[2] == "quick") else 1)
    timing([int(a) for a in sys.argv[2:]])
