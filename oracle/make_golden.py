"""TEST INFRASTRUCTURE: generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the container where /root/reference is mounted:

    make -C oracle && python oracle/make_golden.py

Every file holds the inputs (grid, slowness, sources, receivers) and what the reference's own
Grid3Drnfs / Grid3Drcfs (through oracle/_ref/libttcr_ref.so) produced for them: the full
traveltime field, receiver traveltimes, niter / niterw; tests/golden/rays/ holds the reference's raypaths and the
traveltimes integrated along them for some of the cases (`--rays-only` regenerates just those).  The two `ref_*` cases use the
reference's own test fixtures (tests/files/gradient_medium.vtr, layers_medium.vtr, src.dat,
rcv.dat) and also store the analytic solution at the receivers
(sol_analytique_*_tt.vtr), so the reference's acceptance criterion (mean relative error
< 0.01, tests/test_grid3d.cpp:68-96,181,199) can be re-checked anywhere.
Arrays are stored in numpy (nx,ny,nz) C order.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from ttcr_b200.vtr import read_vtr  # noqa: E402

REF = "/root/reference/tests/files"
OUT = os.path.join(ROOT, "tests", "golden")


def run_ref(x, y, z, slowness, cell, src, t0, rcv, weno, dtype, eps=1e-5, maxit=50, translate=False):
    ncx, ncy, ncz = x.size - 1, y.size - 1, z.size - 1
    dx = float(np.asarray(x, dtype=dtype)[1] - np.asarray(x, dtype=dtype)[0])
    g = O.RefGrid(ncx, ncy, ncz, dx, float(x[0]), float(y[0]), float(z[0]), eps=eps, maxit=maxit, weno=weno,
                  cell_slowness=cell, dtype=dtype, translate_grid=translate)
    g.set_slowness(O.to_cxx(slowness))
    tt_rcv, _ = g.raytrace(src, t0, rcv)
    field = O.from_cxx(g.get_tt(), (x.size, y.size, z.size))
    node_s = O.from_cxx(g.get_slowness(), (x.size, y.size, z.size))
    ni, nw = g.niter()
    g.close()
    out = dict(tt_grid=np.ascontiguousarray(field), tt_rcv=tt_rcv.astype(dtype), niter=ni, niterw=nw)
    if cell:
        out["node_slowness"] = np.ascontiguousarray(node_s.astype(dtype))
    return out


def save(name, **kw):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print(f"{name}: niter={kw.get('niter')} niterw={kw.get('niterw')} {os.path.getsize(path) / 1024:.0f} KiB")


def reference_fixtures():
    src = np.loadtxt(os.path.join(REF, "src.dat"), skiprows=1).reshape(1, 4)
    rcv = np.loadtxt(os.path.join(REF, "rcv.dat"), skiprows=1)
    for tag, model, key, cell, sol in (("gradient", "gradient_medium.vtr", "point_data", False, "sol_analytique_gradient_tt.vtr"),
                                       ("layers", "layers_medium.vtr", "cell_data", True, "sol_analytique_couches_tt.vtr")):
        m = read_vtr(os.path.join(REF, model))
        x, y, z = m["x"], m["y"], m["z"]
        dim = (x.size - 1, y.size - 1, z.size - 1) if cell else (x.size, y.size, z.size)
        slowness = np.ascontiguousarray(m[key]["Slowness"].reshape(dim, order="F"))
        a = read_vtr(os.path.join(REF, sol))
        at = a["point_data"]["Travel Time"].reshape((a["x"].size, a["y"].size, a["z"].size), order="F")
        # rcv.dat points are integer lattice points of the analytic grid (z = 0 face)
        ia = np.rint((rcv[:, 0] - a["x"][0]) / (a["x"][1] - a["x"][0])).astype(int)
        ja = np.rint((rcv[:, 1] - a["y"][0]) / (a["y"][1] - a["y"][0])).astype(int)
        ka = np.rint((rcv[:, 2] - a["z"][0]) / (a["z"][1] - a["z"][0])).astype(int)
        analytic = at[ia, ja, ka]
        for dtype in (np.float64, np.float32):
            for weno in (1, 0):
                r = run_ref(x, y, z, slowness, cell, src[:, 1:4], src[:, 0], rcv, weno, dtype)
                err = np.mean(np.abs(r["tt_rcv"][1:] - analytic[1:]) / analytic[1:])
                assert (not weno) or err < 0.01, err   # the reference's criterion is on weno3=1 runs
                if dtype == np.float32:
                    r["tt_grid"] = r["tt_grid"].astype(np.float32)
                save(f"ref_{tag}_medium_{'w' if weno else 'f'}_{np.dtype(dtype).name}", x=x, y=y, z=z,
                     slowness=slowness, cell_slowness=cell, src=src, rcv=rcv, weno=weno, eps=1e-5, maxit=50,
                     analytic_rcv=analytic, ref_mean_rel_err=err, **r)


def synthetic():
    rng = np.random.default_rng(12345)
    cases = []
    # (name, n nodes per axis (nx,ny,nz), cell, sources (t0,x,y,z), weno, translate, origin)
    cases.append(("syn_het21_corner", (21, 21, 21), False, [[0, 0, 0, 0]], 0.0))
    cases.append(("syn_het21_offnode", (21, 21, 21), False, [[0.5, 3.3, 7.1, 12.9]], 0.0))
    cases.append(("syn_het_ragged", (15, 21, 34), False, [[0, 6.0, 10.0, 4.0]], 0.0))
    cases.append(("syn_het_multitx", (19, 17, 33), False, [[0.1, 2.0, 2.0, 2.0], [0.0, 15.5, 11.2, 3.9], [0.3, 18.0, 16.0, 32.0]], 0.0))
    cases.append(("syn_cells_ragged", (18, 14, 35), True, [[0, 10.0, 8.0, 17.0]], 0.0))
    cases.append(("syn_translate", (17, 17, 17), False, [[0, 1005.0, 2010.0, -495.0]], 1000.0))
    for name, (nx, ny, nz), cell, srcs, org in cases:
        h = 1.0
        x = org + h * np.arange(nx)
        y = 2 * org + h * np.arange(ny)
        z = -0.5 * org + h * np.arange(nz)
        dim = (nx - 1, ny - 1, nz - 1) if cell else (nx, ny, nz)
        X, Y, Z = np.meshgrid(np.arange(dim[0]) * h, np.arange(dim[1]) * h, np.arange(dim[2]) * h, indexing="ij")
        slowness = (1 + 0.3 * np.sin(0.7 * X) * np.cos(0.9 * Y)) / (1 + 0.1 * Z) * np.exp(0.05 * rng.standard_normal(dim))
        src = np.array(srcs, dtype=np.float64)
        rcv = np.column_stack([rng.uniform(x[0], x[-1], 40), rng.uniform(y[0], y[-1], 40), rng.uniform(z[0], z[-1], 40)])
        rcv[:5] = np.column_stack([x[[0, -1, 3, 5, 7]], y[[0, -1, 3, 6, 2]], z[[0, -1, 9, 1, 4]]])   # on-node receivers
        rcv[5, 0] = x[4]; rcv[6, 1] = y[2]; rcv[7, 2] = z[8]                                      # on-face receivers
        rcv[8, :2] = (x[3], y[3]); rcv[9, 1:] = (y[1], z[1]); rcv[10, ::2] = (x[2], z[2])          # on-edge receivers
        for dtype in (np.float64, np.float32):
            for weno in (0, 1):
                r = run_ref(x, y, z, slowness, cell, src[:, 1:4], src[:, 0], rcv, weno, dtype, translate=(org != 0.0))
                save(f"{name}_{'w' if weno else 'f'}_{np.dtype(dtype).name}", x=x, y=y, z=z, slowness=slowness,
                     cell_slowness=cell, src=src, rcv=rcv, weno=weno, eps=1e-5, maxit=50, translate=(org != 0.0), **r)


def mt19937_64_uniform(seed, n, lo, hi):
    """std::mt19937_64(seed) driving std::uniform_real_distribution<double>(lo, hi) as libstdc++ implements it (one
    64-bit draw per number: generate_canonical = draw * 2^-64 in double, clamped below 1): the source positions of the
    reference's accuracy study (tests/accuracy_grid3d.cpp:352-360)."""
    NN, MM, MASK = 312, 156, (1 << 64) - 1
    mt = [0] * NN
    mt[0] = seed & MASK
    for i in range(1, NN):
        mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & MASK
    idx = NN
    out = np.empty(n)
    for k in range(n):
        if idx >= NN:
            for i in range(NN):
                x = (mt[i] & 0xFFFFFFFF80000000) | (mt[(i + 1) % NN] & 0x7FFFFFFF)
                xa = x >> 1
                if x & 1:
                    xa ^= 0xB5026F5AA96619E9
                mt[i] = mt[(i + MM) % NN] ^ xa
            idx = 0
        y = mt[idx]
        idx += 1
        y ^= (y >> 29) & 0x5555555555555555
        y ^= (y << 17) & 0x71D67FFFEDA60000
        y ^= (y << 37) & 0xFFF7EEE000000000
        y ^= y >> 43
        u = float(y) * (1.0 / 18446744073709551616.0)
        if u >= 1.0:
            u = np.nextafter(1.0, 0.0)
        out[k] = u * (hi - lo) + lo
    return out


def accuracy_kat():
    """tests/golden/kat/kat_constant_medium.npz: study 2 of the reference's tests/accuracy_grid3d.cpp (constant_medium.vtr,
    100 random sources from mt19937_64(12345), receivers rcv.dat, FSM with weno3 = 1, double, eps 1e-5, nitermax 50,
    receiver times by interpolation), run through the UNMODIFIED reference; its published result for this case is a mean
    relative error of 0.00152022 (tests/accuracy_grid3d.csv, line "double,constant,FAST_SWEEPING,Grid3Drnfs,medium")."""
    m = read_vtr(os.path.join(REF, "constant_medium.vtr"))
    x, y, z = m["x"], m["y"], m["z"]
    slowness = np.ascontiguousarray(m["point_data"]["Slowness"].reshape((x.size, y.size, z.size), order="F"))
    rcv = np.loadtxt(os.path.join(REF, "rcv.dat"), skiprows=1)
    src = mt19937_64_uniform(12345, 300, 0.5, 19.5).reshape(100, 3)
    dx = float(x[1] - x[0])
    g = O.RefGrid(x.size - 1, y.size - 1, z.size - 1, dx, float(x[0]), float(y[0]), float(z[0]), eps=1e-5, maxit=50, weno=True,
                  cell_slowness=False, dtype=np.float64)
    g.set_slowness(O.to_cxx(slowness))
    tt = np.empty((100, rcv.shape[0]))
    it = np.empty((100, 2), dtype=np.int64)
    for n in range(100):
        tt[n], _ = g.raytrace(src[n:n + 1], 0.0, rcv)
        it[n] = g.niter()
    g.close()
    s0 = float(slowness.ravel()[0])
    ref = s0 * np.sqrt(((rcv[None, :, :] - src[:, None, :]) ** 2).sum(axis=2))
    err = float(np.mean(np.abs((ref - tt) / ref)[ref != 0.0]))
    os.makedirs(os.path.join(OUT, "kat"), exist_ok=True)
    save(os.path.join("kat", "kat_constant_medium"), x=x, y=y, z=z, slowness=np.float64(s0), src=src, rcv=rcv, tt_rcv=tt, iters=it, error=err,
         published_error=0.00152022)
    print(f"kat_constant_medium: mean relative error {err:.8f} (published 0.00152022), niterw {it[:, 1].min()}..{it[:, 1].max()}")


RAY_CASES = ("syn_het21_offnode_w_float64", "syn_het21_offnode_f_float32", "syn_het_ragged_f_float64", "syn_het_ragged_w_float32",
             "syn_cells_ragged_w_float64", "syn_cells_ragged_f_float32", "syn_translate_w_float64")


def raypath_case(base):
    g = dict(np.load(os.path.join(OUT, base + ".npz")))
    x, y, z = g["x"], g["y"], g["z"]
    dtype = g["tt_grid"].dtype
    rcv = g["rcv"]
    # receivers on the faces of the grid are left out, and so are cases with a source point on a face (the reference's own
    # gradient_medium / src.dat fixture: source on the corner; syn_het_multitx: one Tx on z max): the reference's walk reads
    # out of bounds there and dies with SIGSEGV (the CUDA path and the restatement clamp their indices instead)
    inside = np.all((rcv > [x[0], y[0], z[0]]) & (rcv < [x[-1], y[-1], z[-1]]), axis=1)
    rcv = rcv[inside]
    dx = float(np.asarray(x, dtype=dtype)[1] - np.asarray(x, dtype=dtype)[0])
    ref = O.RefGrid(x.size - 1, y.size - 1, z.size - 1, dx, float(x[0]), float(y[0]), float(z[0]), eps=float(g["eps"]),
                    maxit=int(g["maxit"]), weno=bool(g["weno"]), cell_slowness=bool(g["cell_slowness"]), dtype=dtype,
                    translate_grid=bool(g.get("translate", False)))
    ref.set_slowness(O.to_cxx(g["slowness"]))
    tt, rays = ref.raytrace_rays(g["src"][:, 1:4], g["src"][:, 0], rcv)
    ref.close()
    npts = np.array([len(r) for r in rays], dtype=np.int64)
    out = os.path.join(OUT, "rays")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, base + ".npz")
    np.savez_compressed(path, base=base, rcv=rcv, rp_tt=tt.astype(dtype), rp_npts=npts, rp_xyz=np.vstack(rays))
    print(f"rays/{base}: {len(rays)} rays, {int(npts.sum())} points, {os.path.getsize(path) / 1024:.0f} KiB", flush=True)


def raypaths():
    """tests/golden/rays/<case>.npz: what the reference's rays overload (Grid3D::raytrace(Tx,t0,Rx,tt,r_data,threadNo),
    i.e. Grid3Drn::getRaypath per receiver) returns for some of the cases above: traveltimes along the rays, points per
    ray and the points.  One subprocess per case: a case the reference cannot walk kills only its own process."""
    import subprocess
    for base in RAY_CASES:
        rc = subprocess.run([sys.executable, os.path.abspath(__file__), "--ray-case", base]).returncode
        if rc:
            print(f"rays/{base}: the reference died (rc {rc}); no fixture written", flush=True)


if __name__ == "__main__":
    if not O.have_ref():
        sys.exit("oracle/_ref/libttcr_ref.so missing: run `make -C oracle` where /root/reference exists")
    if "--ray-case" in sys.argv:
        raypath_case(sys.argv[sys.argv.index("--ray-case") + 1])
        sys.exit(0)
    if "--kat-only" in sys.argv:
        accuracy_kat()
        sys.exit(0)
    if "--rays-only" not in sys.argv:
        reference_fixtures()
        synthetic()
        accuracy_kat()
    raypaths()
