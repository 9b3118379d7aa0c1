"""CPU tests of the oracle (tests/ may use oracle/; the product may not).

The plain-C restatement (oracle/fsm_oracle.c) must reproduce, bit for bit, the fields the
UNMODIFIED reference produced (tests/golden/*.npz, written by oracle/make_golden.py through
oracle/_ref) -- and the live reference when oracle/_ref is present.
"""
import os

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden, golden_ray_names, load_golden_rays


def _run_oracle(O, g, order=0):
    dt = g["dtype"]
    x, y, z = g["x"], g["y"], g["z"]
    ncx, ncy, ncz = x.size - 1, y.size - 1, z.size - 1
    xt = np.asarray(x, dtype=dt)
    dx = float(xt[1] - xt[0])
    org = np.array([x[0], y[0], z[0]], dtype=dt)
    s = O.to_cxx(g["slowness"].astype(dt))
    if g["cell_slowness"]:
        s = O.cell_to_node(s, ncx, ncy, ncz, dtype=dt)
    src = g["src"]
    tx = src[:, 1:4].astype(dt)
    rx = g["rcv"].astype(dt)
    mins = org.copy()
    if g["translate"]:   # Grid3D.h:478-485, Grid3Drn.h:362-369
        tx = tx - org
        rx = rx - org
        mins = np.zeros(3, dtype=dt)
    tt, ni, nw = O.solve(ncx, ncy, ncz, dx, s, tx, src[:, 0].astype(dt), float(mins[0]), float(mins[1]), float(mins[2]),
                         eps=g["eps"], maxit=g["maxit"], weno=g["weno"], dtype=dt, order=order)
    trx = O.interp(ncx, ncy, ncz, dx, tt, rx, float(mins[0]), float(mins[1]), float(mins[2]), dtype=dt)
    return O.from_cxx(tt, (x.size, y.size, z.size)), trx, ni, nw, s


@pytest.mark.parametrize("name", golden_names())
def test_restatement_matches_reference_golden_bitwise(oracle, name):
    g = load_golden(name)
    tt, trx, ni, nw, s = _run_oracle(oracle, g)
    assert (ni, nw) == (g["niter"], g["niterw"])
    assert tt.dtype == g["tt_grid"].dtype
    assert np.array_equal(tt, g["tt_grid"]), f"max |d| = {np.abs(tt - g['tt_grid']).max()}"
    assert np.array_equal(trx, g["tt_rcv"])
    if g["cell_slowness"]:
        assert np.array_equal(oracle.from_cxx(s, tt.shape), g["node_slowness"])


@pytest.mark.parametrize("name", golden_names("syn_het_ragged*") + golden_names("syn_cells*"))
def test_plane_order_is_bit_identical_to_lexicographic(oracle, name):
    """any topological order of the Gauss-Seidel DAG gives the same field (what the GPU kernels rely on)"""
    g = load_golden(name)
    tt, _, ni, nw, _ = _run_oracle(oracle, g, order=1)
    assert (ni, nw) == (g["niter"], g["niterw"])
    assert np.array_equal(tt, g["tt_grid"])


@pytest.mark.parametrize("name", golden_names("ref_*_w_*"))
def test_reference_acceptance_criterion(oracle, name):
    """tests/test_grid3d.cpp:68-96,181,199: mean relative error vs analytic < 0.01 over rcv.dat[1:]"""
    g = load_golden(name)
    _, trx, _, _, _ = _run_oracle(oracle, g)
    err = np.mean(np.abs(trx[1:] - g["analytic_rcv"][1:]) / g["analytic_rcv"][1:])
    assert err < 0.01
    assert abs(err - float(g["ref_mean_rel_err"])) < 1e-12


def test_live_reference_if_present(oracle):
    """where /root/reference was compiled (oracle/_ref), re-pin the restatement on a fresh random case"""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built on this machine")
    rng = np.random.default_rng(7)
    n = (17, 23, 20)
    s = rng.uniform(0.2, 1.0, n)
    for dt in (np.float64, np.float32):
        for weno in (0, 1):
            g = oracle.RefGrid(n[0] - 1, n[1] - 1, n[2] - 1, 0.5, 1.0, -2.0, 3.0, weno=weno, dtype=dt)
            g.set_slowness(oracle.to_cxx(s))
            tx = np.array([[3.25, 1.0, 7.5], [1.0, -2.0, 3.0]])
            g.raytrace(tx, [0.0, 0.25], np.zeros((0, 3)))
            ref = g.get_tt()
            tt, ni, nw = oracle.solve(n[0] - 1, n[1] - 1, n[2] - 1, 0.5, g.get_slowness().astype(dt), tx, [0.0, 0.25],
                                      1.0, -2.0, 3.0, weno=weno, dtype=dt)
            assert (ni, nw) == g.niter()
            assert np.array_equal(tt, ref)


def test_oracle_errors(oracle):
    s = np.ones(27)
    with pytest.raises(RuntimeError):
        oracle.solve(2, 2, 2, 1.0, s, [[5.0, 0, 0]])
    with pytest.raises(ValueError):
        oracle.solve(2, 2, 2, 1.0, np.ones(26), [[0.0, 0, 0]])
    with pytest.raises(ValueError):
        oracle.cell_to_node(np.ones(7), 2, 2, 2)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("weno", [False, True])
def test_tt_from_raypath_restatement_is_bit_identical_to_reference(oracle, dtype, weno):
    """Grid3Drn::getTraveltimeFromRaypath (+ grad, computeSlowness): the restatement walks the reference's own field and
    must return the reference's own receiver times, bit for bit, in double and in float"""
    O = oracle
    if not O.have_ref():
        pytest.skip("needs oracle/_ref (built from /root/reference); the golden raypath times cover the GPU box")
    n = 33
    x = np.linspace(0.0, 20.0, n)
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    s = ((1 + 0.3 * np.sin(0.7 * X) * np.cos(0.9 * Y)) / (1 + 0.1 * Z)).astype(dtype)
    xt = x.astype(dtype)
    dx = float(xt[1] - xt[0])
    rng = np.random.default_rng(1)
    for src, t0 in ((np.array([[3.3, 7.1, 12.9]]), 0.0), (np.array([[x[4], x[9], x[20]], [15.2, 3.3, 8.8]]), np.array([0.1, 0.3]))):
        rcv = np.vstack([rng.uniform(1.5, 18.5, (60, 3)), [[x[5], x[7], 3.3], [x[5], 2.2, x[9]], [1.1, x[3], x[4]],
                                                           [x[10], x[11], x[12]], src[0], [x[2], 5.5, 7.7]]])
        g = O.RefGrid(n - 1, n - 1, n - 1, dx, weno=weno, dtype=dtype, tt_from_rp=True)
        g.set_slowness(O.to_cxx(s))
        tref, _ = g.raytrace(src, t0, rcv)
        t = O.tt_from_rp(n - 1, n - 1, n - 1, dx, g.get_tt(), O.to_cxx(s), src, t0, rcv, dtype=dtype)
        g.close()
        assert np.array_equal(t.astype(np.float64), tref)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_raypath_restatement_is_bit_identical_to_reference(oracle, dtype):
    """Grid3Drn::getRaypath (the rays overload of Grid3D::raytrace): same points, same count, same traveltimes"""
    O = oracle
    if not O.have_ref():
        pytest.skip("needs oracle/_ref (built from /root/reference)")
    n = 33
    x = np.linspace(0.0, 20.0, n)
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    s = ((1 + 0.3 * np.sin(0.7 * X) * np.cos(0.9 * Y)) / (1 + 0.1 * Z)).astype(dtype)
    dx = float(x.astype(dtype)[1] - x.astype(dtype)[0])
    rng = np.random.default_rng(2)
    for src, t0 in ((np.array([[3.3, 7.1, 12.9]]), 0.0), (np.array([[x[4], x[9], x[20]], [15.2, 3.3, 8.8]]), np.array([0.1, 0.3]))):
        rcv = np.vstack([rng.uniform(1.5, 18.5, (40, 3)), [[x[5], x[7], 3.3], [x[10], x[11], x[12]], src[0]]])
        g = O.RefGrid(n - 1, n - 1, n - 1, dx, weno=True, dtype=dtype, tt_from_rp=False)
        g.set_slowness(O.to_cxx(s))
        tref, rref = g.raytrace_rays(src, t0, rcv)
        t, r = O.raypaths(n - 1, n - 1, n - 1, dx, g.get_tt(), O.to_cxx(s), src, t0, rcv, dtype=dtype)
        g.close()
        assert np.array_equal(t.astype(np.float64), tref)
        assert len(r) == len(rref)
        for a, b in zip(r, rref):
            assert a.shape == b.shape and np.array_equal(a, b)
        assert r[-1].shape == (1, 3)      # a receiver on the source: the ray is that point alone
        assert min(len(a) for a in r[:-1]) >= 2


@pytest.mark.parametrize("name", golden_ray_names())
def test_raypath_restatement_matches_golden_rays(oracle, name):
    """the committed raypaths of the reference (tests/golden/rays, Grid3D::raytrace with r_data): the restatement walks the
    reference's golden field and returns the same traveltimes and the same points, bit for bit, in double and in float"""
    g = load_golden(name)
    r = load_golden_rays(name)
    if g["translate"]:
        pytest.skip("origin translation is the library's business (GPU test), not the restatement's")
    dtype = g["dtype"]
    x, y, z = g["x"], g["y"], g["z"]
    dx = float(np.asarray(x, dtype=dtype)[1] - np.asarray(x, dtype=dtype)[0])
    s_node = g["node_slowness"] if g["cell_slowness"] else g["slowness"]
    tt, rays = oracle.raypaths(x.size - 1, y.size - 1, z.size - 1, dx, oracle.to_cxx(g["tt_grid"]), oracle.to_cxx(s_node),
                               g["src"][:, 1:4], g["src"][:, 0], r["rcv"], float(x[0]), float(y[0]), float(z[0]), dtype=dtype)
    assert np.array_equal(tt, r["rp_tt"])
    assert [len(a) for a in rays] == r["rp_npts"].tolist()
    for a, b in zip(rays, r["rays"]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_raypath_restatement_interp_vel_bit_identical_to_reference(oracle, dtype):
    """interp_vel = 1 (processVel, Grid3Drn.h:2489-2669): node VELOCITIES are interpolated along the raypath and inverted"""
    O = oracle
    if not O.have_ref():
        pytest.skip("needs oracle/_ref (built from /root/reference)")
    n = 29
    x = np.linspace(0.0, 14.0, n)
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    s = ((1 + 0.3 * np.sin(0.7 * X) * np.cos(0.9 * Y)) / (1 + 0.1 * Z)).astype(dtype)
    dx = float(x.astype(dtype)[1] - x.astype(dtype)[0])
    rng = np.random.default_rng(4)
    src = np.array([[3.3, 7.1, 9.9]])
    rcv = np.vstack([rng.uniform(1.5, 12.5, (40, 3)), [[x[5], x[7], 3.3], [x[10], x[11], x[12]], [x[3], 5.5, x[9]]]])
    out = {}
    for iv in (False, True):
        g = O.RefGrid(n - 1, n - 1, n - 1, dx, weno=True, dtype=dtype, tt_from_rp=True, interp_vel=iv)
        g.set_slowness(O.to_cxx(s))
        tref, rref = g.raytrace_rays(src, 0.0, rcv)
        t, r = O.raypaths(n - 1, n - 1, n - 1, dx, g.get_tt(), O.to_cxx(s), src, 0.0, rcv, dtype=dtype, interp_vel=iv)
        g.close()
        assert np.array_equal(t.astype(np.float64), tref)
        for a, b in zip(r, rref):
            assert np.array_equal(a, b)
        out[iv] = t
    assert not np.array_equal(out[False], out[True])     # the option does change the traveltimes


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", ["square", "square_rotated", "xz", "weno", "weno_xz"])
def test_2d_restatement_is_bit_identical_to_reference(oracle, dtype, case):
    """Grid2Drnfs (oracle only so far): all five node updates, on- and off-node and multi-point
    sources, first-order and WENO, dx == dz and dx != dz: field and iteration counts equal to the unmodified reference's"""
    O = oracle
    if not O.have_ref():
        pytest.skip("needs oracle/_ref (built from /root/reference)")
    ncx, ncz = 47, 61
    dx = 0.5
    dz = 0.5 if "xz" not in case else 0.3
    rng = np.random.default_rng(21)
    X, Z = np.meshgrid(np.arange(ncx + 1) * dx, np.arange(ncz + 1) * dz, indexing="ij")
    s = ((1 + 0.3 * np.sin(0.5 * X) * np.cos(0.4 * Z)) / (1 + 0.05 * Z) * np.exp(0.03 * rng.standard_normal(X.shape))).astype(dtype)
    weno = case.startswith("weno")
    rot = case == "square_rotated"
    for tx, t0 in (([[0.0, 0.0]], 0.0), ([[7.3, 4.9]], 0.25), ([[3 * dx, 5 * dz], [20.1, 11.7]], [0.0, 0.4]),
                   ([[ncx * dx, ncz * dz]], 0.0)):
        a, ni, nw = O.solve2d(ncx, ncz, dx, dz, s, tx, t0, weno=weno, rotated=rot, dtype=dtype)
        b, ri, rw = O.ref_solve2d(ncx, ncz, dx, dz, s, tx, t0, weno=weno, rotated=rot, dtype=dtype)
        assert (ni, nw) == (ri, rw)
        assert np.array_equal(a, b)
        assert np.all(np.isfinite(a)) and a.max() < 1e3


def test_2d_restatement_homogeneous_analytic(oracle):
    """constant slowness: the 2-D solution is s * distance; first-order FSM error a few percent of a cell, WENO smaller"""
    n, h, sl = 101, 0.2, 0.5
    s = np.full((n, n), sl)
    src = [[10.0, 10.0]]
    X, Z = np.meshgrid(np.arange(n) * h, np.arange(n) * h, indexing="ij")
    exact = sl * np.hypot(X - 10.0, Z - 10.0)
    m = exact > 3 * h * sl
    errs = {}
    for weno in (False, True):
        t, ni, nw = oracle.solve2d(n - 1, n - 1, h, h, s, src, weno=weno)
        errs[weno] = float(np.mean(np.abs(t[m] - exact[m]) / exact[m]))
        assert ni >= 2 and (nw >= 2) == weno
    assert errs[False] < 2e-2 and errs[True] < errs[False]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("weno", [False, True])
def test_2d_cell_slowness_and_receivers_bit_identical_to_reference(oracle, dtype, weno):
    """Grid2Drcfs: cell -> node averaging (checked through the solve: the reference does not hand the node values out) and
    Grid2Drn::getTraveltime at receivers inside cells, on edges and on nodes"""
    O = oracle
    if not O.have_ref():
        pytest.skip("needs oracle/_ref (built from /root/reference)")
    ncx, ncz, h = 40, 33, 0.5
    rng = np.random.default_rng(8)
    sc = rng.uniform(0.3, 1.0, (ncx, ncz)).astype(dtype)
    tx, t0 = [[6.2, 9.9]], 0.1
    rx = np.vstack([np.column_stack([rng.uniform(0, ncx * h, 30), rng.uniform(0, ncz * h, 30)]),
                    [[3 * h, 7.77], [4.44, 5 * h], [8 * h, 9 * h], [0.0, 0.0], [ncx * h - 0.2, ncz * h - 0.3]]])
    sn = O.cell_to_node2d(sc, ncx, ncz, dtype=dtype)
    a, ni, nw = O.solve2d(ncx, ncz, h, h, sn, tx, t0, weno=weno, dtype=dtype)
    b, ri, rw, tr = O.ref_solve2d(ncx, ncz, h, h, sc, tx, t0, weno=weno, dtype=dtype, cell_slowness=True, rx=rx)
    assert (ni, nw) == (ri, rw) and np.array_equal(a, b)
    assert np.array_equal(O.interp2d(ncx, ncz, h, h, a, rx, dtype=dtype), tr)


# the reference's own accuracy table (tests/accuracy_grid3d.csv, study 1: FAST_SWEEPING on the medium models, mean relative
# error of the receiver times against the analytic solution); make_golden.py stored what the unmodified reference gives here
PUBLISHED = {("gradient", "float64"): 0.00228619, ("gradient", "float32"): 0.00228538,
             ("layers", "float64"): 0.00669965, ("layers", "float32"): 0.00670199}


@pytest.mark.parametrize("model,dt", sorted(PUBLISHED))
def test_published_accuracy_numbers_are_reproduced(oracle, model, dt):
    """the reference run here, and the restatement, reproduce the reference's published accuracy figures to all six digits"""
    g = load_golden(f"ref_{model}_medium_w_{dt}")
    assert float(f"{float(g['ref_mean_rel_err']):.6g}") == PUBLISHED[(model, dt)]
    dtype = g["dtype"]
    x, y, z = g["x"], g["y"], g["z"]
    dx = float(np.asarray(x, dtype=dtype)[1] - np.asarray(x, dtype=dtype)[0])
    s = g["slowness"]
    if g["cell_slowness"]:
        s_node = oracle.cell_to_node(oracle.to_cxx(s.astype(dtype)), x.size - 1, y.size - 1, z.size - 1, dtype=dtype)
    else:
        s_node = oracle.to_cxx(s.astype(dtype))
    tt, _, _ = oracle.solve(x.size - 1, y.size - 1, z.size - 1, dx, s_node, g["src"][:, 1:4].astype(dtype), g["src"][:, 0], weno=True,
                            dtype=dtype)
    t = oracle.interp(x.size - 1, y.size - 1, z.size - 1, dx, tt, g["rcv"], dtype=dtype).astype(np.float64)
    a = g["analytic_rcv"]
    err = np.mean(np.abs(t[1:] - a[1:]) / a[1:])
    assert float(f"{err:.6g}") == PUBLISHED[(model, dt)]


def test_accuracy_study_constant_model_random_sources(oracle):
    """study 2 of tests/accuracy_grid3d.cpp (constant_medium.vtr, sources from mt19937_64(12345), receivers rcv.dat, weno,
    double) as the unmodified reference computes it here (tests/golden/kat/kat_constant_medium.npz): the restatement gives the
    same receiver times bit for bit (every fourth source), and the fixture's mean relative error is the one of its times.
    (The reference's csv lists 0.00152022 for this case; the reference built here from the same sources gives 0.00118933,
    with the very source positions libstdc++ generates -- the csv predates the tree or comes from another build.)"""
    with np.load(os.path.join(ROOT, "tests", "golden", "kat", "kat_constant_medium.npz")) as f:
        k = {n: f[n] for n in f.files}
    x = k["x"]
    n = x.size
    s0 = float(k["slowness"])
    s_node = np.full(n ** 3, s0)
    ref = s0 * np.sqrt(((k["rcv"][None, :, :] - k["src"][:, None, :]) ** 2).sum(axis=2))
    assert abs(float(np.mean(np.abs((ref - k["tt_rcv"]) / ref)[ref != 0.0])) - float(k["error"])) < 1e-15
    assert abs(float(k["error"]) - 0.00118933) < 5e-9
    for i in range(0, 100, 4):
        tt, ni, nw = oracle.solve(n - 1, n - 1, n - 1, float(x[1] - x[0]), s_node, k["src"][i:i + 1], 0.0, weno=True)
        assert (ni, nw) == tuple(k["iters"][i])
        assert np.array_equal(oracle.interp(n - 1, n - 1, n - 1, float(x[1] - x[0]), tt, k["rcv"]), k["tt_rcv"][i])


def test_headline_size_digests_float_restatement_vs_reference_double():
    """tests/golden/digest (oracle/make_digest.py): at 512^3 the C restatement's fp32 field and the UNMODIFIED reference's
    Grid3Drnfs<double> field agree to 1e-5 on three full planes and at the 441 receivers, same iteration count -- the
    fp32-vs-reference-double distance the GPU tests' 1e-4 has to cover"""
    import os
    base = os.path.join(os.path.dirname(__file__), "golden", "digest")
    a = np.load(os.path.join(base, "c3_512_f32.npz"))
    r = np.load(os.path.join(base, "c3_512_ref_f64.npz"))
    assert int(a["niter"]) == int(r["niter"]) == 2
    assert np.array_equal(a["planes_i"], r["planes_i"]) and np.array_equal(a["rcv"], r["rcv"])
    floor = (20.0 / 511.0) / 3.0
    e = np.abs(a["planes"].astype(np.float64) - r["planes"]) / np.maximum(r["planes"], floor)
    assert e.max() <= 1e-5
    e = np.abs(a["tt_rcv"].astype(np.float64) - r["tt_rcv"]) / np.maximum(r["tt_rcv"], floor)
    assert e.max() <= 1e-5
    for name in ("c4_511c_f32", "c5_1024_f32"):
        d = np.load(os.path.join(base, name + ".npz"))
        assert d["planes_i"].size == 1 and np.all(np.isfinite(d["planes" if "planes" in d else "planes_0"]))
