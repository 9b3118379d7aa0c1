"""CPU tests of the host-side sparse operators (ttcr_b200/matrices.py) mirrored from ttcrpy.rgrid.Grid3d.compute_D /
compute_K (src/ttcrpy/rgrid.pyx:610-756): checked against a per-point / per-parameter loop written from the reference's
description of the operators, and through exactness properties of the interpolation and of the difference stencils."""
import numpy as np
import pytest

from ttcr_b200.matrices import compute_D, compute_K, slowness_at


def _grid():
    x = 1.0 + 0.5 * np.arange(7)
    y = -2.0 + 0.25 * np.arange(6)
    z = 0.4 * np.arange(9)
    return x, y, z


def test_compute_D_nodes_matches_per_point_loop_and_interpolates_linear_fields_exactly():
    x, y, z = _grid()
    rng = np.random.default_rng(0)
    pts = np.column_stack([rng.uniform(x[0], x[-1] - 1e-3, 50), rng.uniform(y[0], y[-1] - 1e-3, 50), rng.uniform(z[0], z[-1] - 1e-3, 50)])
    pts[0] = (x[2], y[3], z[4])          # on a node: weight 1 there, explicit zeros elsewhere
    pts[1] = (x[2], y[3] + 0.1, z[4])    # on an edge
    D = compute_D(x, y, z, pts, cell_slowness=False)
    assert D.shape == (50, 7 * 6 * 9)
    dx, dy, dz = x[1] - x[0], y[1] - y[0], z[1] - z[0]
    ref = np.zeros(D.shape)
    for n, p in enumerate(pts):          # rgrid.pyx:655-672
        i1, j1, k1 = int(1e-6 + (p[0] - x[0]) / dx), int(1e-6 + (p[1] - y[0]) / dy), int(1e-6 + (p[2] - z[0]) / dz)
        for i in (i1, i1 + 1):
            for j in (j1, j1 + 1):
                for k in (k1, k1 + 1):
                    ref[n, (i * y.size + j) * z.size + k] += ((1 - abs(p[0] - x[i]) / dx) * (1 - abs(p[1] - y[j]) / dy) *
                                                              (1 - abs(p[2] - z[k]) / dz))
    assert np.array_equal(D.toarray(), ref)
    assert np.allclose(np.asarray(D.sum(axis=1)).ravel(), 1.0, atol=1e-14)
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    f = 3.0 + 0.7 * X - 1.3 * Y + 0.2 * Z
    assert np.allclose(D @ f.ravel(), 3.0 + 0.7 * pts[:, 0] - 1.3 * pts[:, 1] + 0.2 * pts[:, 2], atol=1e-12)
    assert D[0, (2 * y.size + 3) * z.size + 4] == 1.0


def test_compute_D_cells_and_errors():
    x, y, z = _grid()
    pts = np.array([[x[0] + 0.1, y[0] + 0.1, z[0] + 0.1], [x[3] + 0.2, y[4] + 0.01, z[7] + 0.39]])
    D = compute_D(x, y, z, pts, cell_slowness=True)
    assert D.shape == (2, 6 * 5 * 8) and D.nnz == 2
    assert D[0, 0] == 1.0 and D[1, (3 * 5 + 4) * 8 + 7] == 1.0
    with pytest.raises(ValueError, match="outside grid"):
        compute_D(x, y, z, np.array([[0.0, 0.0, 0.0]]), cell_slowness=False)


def test_compute_K_matches_per_parameter_loop_and_is_exact_for_quadratics():
    shape, (dx, dy, dz) = (5, 4, 6), (0.5, 0.25, 2.0)
    K = compute_K(shape, dx, dy, dz)
    n = int(np.prod(shape))
    idx = np.arange(n).reshape(shape)
    for axis, (Kq, h) in enumerate(zip(K, (dx, dy, dz))):
        assert Kq.shape == (n, n) and Kq.nnz == 3 * n
        ref = np.zeros((n, n))
        for p in np.ndindex(*shape):      # central operator inside, forward / backward on the first / last index
            c = min(max(p[axis], 1), shape[axis] - 2)
            for off, v in ((-1, 1.0), (0, -2.0), (1, 1.0)):
                q = list(p)
                q[axis] = c + off
                ref[idx[p], idx[tuple(q)]] += v / (h * h)
        assert np.array_equal(Kq.toarray(), ref)
    I, J, Kk = np.meshgrid(np.arange(5) * dx, np.arange(4) * dy, np.arange(6) * dz, indexing="ij")
    f = (1.5 * I ** 2 - 0.5 * J ** 2 + 0.25 * Kk ** 2 + I * J).ravel()
    assert np.allclose(K[0] @ f, 3.0) and np.allclose(K[1] @ f, -1.0) and np.allclose(K[2] @ f, 0.5)
    with pytest.raises(ValueError):
        compute_K((2, 4, 4), 1.0, 1.0, 1.0)


@pytest.mark.parametrize("interp_vel", [False, True])
def test_slowness_at_agrees_with_the_restatement_of_computeSlowness(oracle, interp_vel):
    """Grid3d.get_s0's interpolation against the oracle's Grid3Drn::computeSlowness (bit-identical to the reference inside
    the raypath tests): inside cells, on faces, edges and nodes"""
    n = 17
    x = np.linspace(0.0, 8.0, n)
    rng = np.random.default_rng(3)
    s = rng.uniform(0.3, 1.2, (n, n, n))
    pts = np.vstack([rng.uniform(0.0, 7.99, (200, 3)), [[x[3], x[5], x[7]], [x[3], 2.2, x[7]], [x[3], 2.2, 6.1], [1.1, x[5], 6.1],
                                                         [x[0], x[0], x[0]], [x[-1], x[-1], x[-1]], [x[-1], 3.3, 4.4]]])
    ref = oracle.slowness_at(n - 1, n - 1, n - 1, float(x[1] - x[0]), oracle.to_cxx(s), pts, interp_vel=interp_vel)
    got = slowness_at(x, x, x, s, pts, interp_vel=interp_vel)
    assert np.allclose(got, ref, rtol=1e-12, atol=0)
