"""Host-side sparse operators of ``ttcrpy.rgrid.Grid3d`` that inversion codes build next to the solver:
``compute_D`` (interpolation weights of velocity data points, rgrid.pyx:610-677) and ``compute_K`` (second-derivative
smoothing operators, rgrid.pyx:679-756).  Pure numpy / scipy, vectorised; parameters are indexed like the reference's
flattened (nx, ny, nz) C-order arrays: node (i, j, k) -> (i * ny + j) * nz + k, cell likewise with (ny - 1), (nz - 1).
"""
from __future__ import annotations

import numpy as np


def compute_D(x, y, z, coord, cell_slowness):
    """csr matrix (npts, nparams) of interpolation weights at the points ``coord`` (npts, 3).

    Cells: one entry of 1 per point, in the cell that holds it (index truncated, as rgrid.pyx:645-648).
    Nodes: the 8 trilinear weights of the surrounding nodes (index int(1e-6 + (p - min) / d), rgrid.pyx:658-672); like the
    reference, no special case for points on nodes, edges or faces (their zero weights are stored)."""
    import scipy.sparse as sp
    x, y, z = (np.asarray(a, dtype=np.float64) for a in (x, y, z))
    coord = np.asarray(coord, dtype=np.float64).reshape(-1, 3)
    if (coord[:, 0].min(initial=x[0]) < x[0] or coord[:, 0].max(initial=x[0]) > x[-1] or
            coord[:, 1].min(initial=y[0]) < y[0] or coord[:, 1].max(initial=y[0]) > y[-1] or
            coord[:, 2].min(initial=z[0]) < z[0] or coord[:, 2].max(initial=z[0]) > z[-1]):
        raise ValueError("Velocity data point outside grid")
    dx, dy, dz = x[1] - x[0], y[1] - y[0], z[1] - z[0]
    npts = coord.shape[0]
    if cell_slowness:
        i = ((coord[:, 0] - x[0]) / dx).astype(np.int64)
        j = ((coord[:, 1] - y[0]) / dy).astype(np.int64)
        k = ((coord[:, 2] - z[0]) / dz).astype(np.int64)
        col = (i * (y.size - 1) + j) * (z.size - 1) + k
        ncell = (x.size - 1) * (y.size - 1) * (z.size - 1)
        return sp.csr_matrix((np.ones(npts), (np.arange(npts), col)), shape=(npts, ncell))
    i1 = (1.e-6 + (coord[:, 0] - x[0]) / dx).astype(np.int64)
    j1 = (1.e-6 + (coord[:, 1] - y[0]) / dy).astype(np.int64)
    k1 = (1.e-6 + (coord[:, 2] - z[0]) / dz).astype(np.int64)
    rows, cols, vals = [], [], []
    for di in (0, 1):
        for dj in (0, 1):
            for dk in (0, 1):
                i, j, k = i1 + di, j1 + dj, k1 + dk
                if i.max(initial=0) >= x.size or j.max(initial=0) >= y.size or k.max(initial=0) >= z.size:
                    raise IndexError("data point on the upper face of the grid (the reference indexes past the axis there too)")
                rows.append(np.arange(npts))
                cols.append((i * y.size + j) * z.size + k)
                vals.append((1. - np.abs(coord[:, 0] - x[i]) / dx) * (1. - np.abs(coord[:, 1] - y[j]) / dy) *
                            (1. - np.abs(coord[:, 2] - z[k]) / dz))
    # entry order within a row as in the reference: i outer, j, k inner
    rows = np.stack(rows, axis=1).ravel()
    cols = np.stack(cols, axis=1).ravel()
    vals = np.stack(vals, axis=1).ravel()
    return sp.csr_matrix((vals, (rows, cols)), shape=(npts, x.size * y.size * z.size))


def compute_K(shape, dx, dy, dz):
    """(Kx, Ky, Kz): second-derivative operators on the (nx, ny, nz) parameter grid, rows = parameters.
    Central stencil (1, -2, 1) / h^2 inside, the same stencil shifted to a forward / backward one on the first / last index
    of the axis (rgrid.pyx:688-690)."""
    import scipy.sparse as sp
    nx, ny, nz = shape
    n = nx * ny * nz
    idx = np.arange(n, dtype=np.int64).reshape(nx, ny, nz)
    out = []
    for axis, (m, h) in enumerate(((nx, dx), (ny, dy), (nz, dz))):
        if m < 3:
            raise ValueError("compute_K needs at least 3 parameters along every axis")
        centre = np.clip(np.arange(m), 1, m - 2)          # the stencil's middle index: shifted inwards at both ends
        sl = [slice(None)] * 3
        cols = []
        for off in (-1, 0, 1):
            sl[axis] = centre + off
            cols.append(idx[tuple(sl)].ravel())
        rows = np.repeat(np.arange(n, dtype=np.int64), 3)
        cols = np.stack(cols, axis=1).ravel()
        vals = np.tile(np.array([1., -2., 1.]), n) / (h * h)
        out.append(sp.csr_matrix((vals, (rows, cols)), shape=(n, n)))
    return tuple(out)


def slowness_at(x, y, z, s_node, pts, interp_vel=False):
    """Node slowness interpolated at ``pts`` (n, 3) as ``Grid3Drn::computeSlowness`` does it (ttcr/Grid3Drn.h:2451-2676):
    trilinear between the 8 nodes of the cell int(1e-4 + (p - min) / d), of the slowness or -- ``interp_vel`` -- of the
    velocity (then inverted).  The reference returns the node / edge / face values through separate branches; their
    results agree with the trilinear formula to rounding, which is what ``Grid3d.get_s0`` needs."""
    x, y, z = (np.asarray(a, dtype=np.float64) for a in (x, y, z))
    s = np.asarray(s_node, dtype=np.float64).reshape(x.size, y.size, z.size)
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
    f = 1.0 / s if interp_vel else s
    dx, dy, dz = x[1] - x[0], y[1] - y[0], z[1] - z[0]
    i = np.minimum((1.e-4 + (pts[:, 0] - x[0]) / dx).astype(np.int64), x.size - 2)
    j = np.minimum((1.e-4 + (pts[:, 1] - y[0]) / dy).astype(np.int64), y.size - 2)
    k = np.minimum((1.e-4 + (pts[:, 2] - z[0]) / dz).astype(np.int64), z.size - 2)
    wx, wy, wz = (pts[:, 0] - x[i]) / dx, (pts[:, 1] - y[j]) / dy, (pts[:, 2] - z[k]) / dz
    v = 0.0
    for di, ax in ((0, 1 - wx), (1, wx)):
        for dj, ay in ((0, 1 - wy), (1, wy)):
            for dk, az in ((0, 1 - wz), (1, wz)):
                v = v + f[i + di, j + dj, k + dk] * ax * ay * az
    return 1.0 / v if interp_vel else v
