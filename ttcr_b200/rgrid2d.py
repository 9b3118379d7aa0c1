"""``ttcrpy.rgrid.Grid2d`` (fast-sweeping branch) over the B200 library: the 2-D twins of the FSM path.

Mirrors ``src/ttcrpy/rgrid.pyx`` ``Grid2d`` for ``method='FSM'``: same constructor arguments, the same ``raytrace`` /
``set_slowness`` conventions and error strings; arrays are numpy ``(nx, nz)`` in C order, which IS the reference's node /
cell order in 2-D.  Everything outside the FSM branch (SPM / DSPM, anisotropy, raypaths, L matrices) is out of scope
and refused.  The solver is ``ttcr_b200/csrc/grid2d.cuh`` behind the C ABI of ``include/ttcr_b200.h``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_DT = {np.dtype(np.float64): _lib.F64, np.dtype(np.float32): _lib.F32}


class Grid2d:
    """Grid2d(x, z, n_threads=1, cell_slowness=1, method='FSM', aniso='iso', eps=1.e-5, maxit=50, weno=1, rotated_template=0,
    nsnx=10, nsnz=10, n_secondary=3, n_tertiary=3, radius_factor_tertiary=3.0, tt_from_rp=0, fsm_gpu=True, dtype=np.float64,
    device=-1)"""

    def __init__(self, x, z, n_threads=1, cell_slowness=1, method="FSM", aniso="iso", eps=1.e-5, maxit=50, weno=1,
                 rotated_template=0, nsnx=10, nsnz=10, n_secondary=3, n_tertiary=3, radius_factor_tertiary=3.0, tt_from_rp=0,
                 fsm_gpu=True, dtype=np.float64, device=-1):
        self.dtype = np.dtype(dtype)
        if self.dtype not in _DT:
            raise ValueError("dtype must be np.float32 or np.float64, got {}".format(dtype))
        self._h = None
        if method != "FSM":
            raise ValueError("ttcr_b200 implements the fast-sweeping method only (method='FSM')")   # SPM / DSPM: out of scope
        if aniso != "iso":
            raise ValueError("Anisotropy is implemented only for the SPM method")                   # as rgrid.pyx says
        if tt_from_rp:
            raise NotImplementedError("tt_from_rp: raypaths are not part of the 2-D FSM path of ttcr_b200")
        self._x = np.ascontiguousarray(x, dtype=self.dtype)
        self._z = np.ascontiguousarray(z, dtype=self.dtype)
        if self._x.ndim != 1 or self._z.ndim != 1 or self._x.size < 2 or self._z.size < 2:
            raise ValueError("x and z should be 1D arrays of at least two node coordinates")
        self._dx = float(self._x[1] - self._x[0])
        self._dz = float(self._z[1] - self._z[0])
        for v, d in ((self._x, self._dx), (self._z, self._dz)):
            if np.any(np.abs(np.diff(v.astype(np.float64)) - d) > 1.e-4 * abs(d)):
                raise ValueError("FSM: grid spacing must be constant along each axis")
        if weno and rotated_template:
            rotated_template = 0   # (the reference ignores the rotated template in the WENO scheme, Grid2Drnfs.h:232-262)
        self.cell_slowness = bool(cell_slowness)
        self._n_threads = int(n_threads)
        self.eps, self.maxit, self.weno, self.rotated_template = float(eps), int(maxit), bool(weno), bool(rotated_template)
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.ttcr_b200_create2d(C.byref(h), self._x.size - 1, self._z.size - 1, self._dx, self._dz, float(self._x[0]),
                                                float(self._z[0]), self.eps, self.maxit, int(self.weno), int(self.rotated_template),
                                                self._n_threads, int(self.cell_slowness), _DT[self.dtype], int(device)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ttcr_b200_destroy2d(self._h)
            self._h = None

    __del__ = close

    def _chk(self, rc):
        _lib.check(rc, None)

    # ---- properties of the reference class ----------------------------------------------------------
    @property
    def x(self):
        return self._x.copy()

    @property
    def z(self):
        return self._z.copy()

    @property
    def dx(self):
        return self._dx

    @property
    def dz(self):
        return self._dz

    @property
    def n_threads(self):
        return self._n_threads

    @property
    def shape(self):
        if self.cell_slowness:
            return (self._x.size - 1, self._z.size - 1)
        return (self._x.size, self._z.size)

    @property
    def nparams(self):
        return int(np.prod(self.shape))

    def get_number_of_nodes(self):
        return self._x.size * self._z.size

    def get_number_of_cells(self):
        return (self._x.size - 1) * (self._z.size - 1)

    def is_outside(self, pts):
        pts = np.asarray(pts)
        return bool(np.any(pts[:, 0] < self._x[0]) or np.any(pts[:, 0] > self._x[-1]) or np.any(pts[:, 1] < self._z[0]) or
                    np.any(pts[:, 1] > self._z[-1]))

    # ---- model ---------------------------------------------------------------------------------------
    def set_slowness(self, slowness):
        """Assign slowness: ndarray of shape (nx, nz) nodes or cells, or flattened in 'C' order."""
        nx, nz = self.shape
        slowness = np.asarray(slowness)
        if slowness.size != nx * nz:
            raise ValueError("Slowness vector has wrong size")
        if slowness.ndim == 2:
            if slowness.shape != (nx, nz):
                raise ValueError("Slowness has wrong shape")
        elif slowness.ndim != 1:
            raise ValueError("Slowness must be 1D or 2D ndarray")
        s = np.ascontiguousarray(slowness, dtype=self.dtype).reshape(-1)
        self._chk(self._lib.ttcr_b200_set_slowness2d(self._h, s.ctypes.data, s.size))

    def set_velocity(self, velocity):
        self.set_slowness(1.0 / np.asarray(velocity, dtype=np.float64))

    def get_slowness(self):
        """node slowness, shape (nx+1, nz+1) (cell models: after the reference's cell -> node averaging)"""
        out = np.empty((self._x.size, self._z.size), dtype=self.dtype)
        self._chk(self._lib.ttcr_b200_get_slowness2d(self._h, out.ctypes.data))
        return out

    def get_grid_traveltimes(self, thread_no=0):
        if thread_no >= self._n_threads:
            raise ValueError("Thread number is larger than number of threads")
        out = np.empty((self._x.size, self._z.size), dtype=self.dtype)
        self._chk(self._lib.ttcr_b200_get_tt2d(self._h, out.ctypes.data, int(thread_no)))
        return out

    def get_niter(self, thread_no=0):
        a, b = C.c_int(), C.c_int()
        self._chk(self._lib.ttcr_b200_get_niter2d(self._h, int(thread_no), C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_solve_ms(self):
        return float(self._lib.ttcr_b200_last_solve_ms2d(self._h))

    # ---- raytrace --------------------------------------------------------------------------------------
    def raytrace(self, source, rcv, slowness=None, thread_no=None, aggregate_src=False, compute_L=False, return_rays=False):
        """Perform raytracing; arguments and return value as ``ttcrpy.rgrid.Grid2d.raytrace`` (FSM branch).

        source: 2D array with 2 (x,z), 3 (t0,x,z) or 4 (evID,t0,x,z) columns; rcv: 2D array (x,z)."""
        source = np.asarray(source)
        rcv = np.asarray(rcv)
        if source.ndim != 2 or rcv.ndim != 2:
            raise ValueError("source and rcv should be 2D arrays")
        if compute_L or return_rays:
            raise NotImplementedError("L matrices and raypaths are not part of the 2-D FSM path of ttcr_b200")
        evID = None
        if source.shape[1] == 4:
            src, t0, evID = source[:, 2:4], source[:, 1], source[:, 0]
            eid = np.sort(np.unique(evID))
            nTx = len(eid)
        elif source.shape[1] == 2:
            src = source
            _, ind = np.unique(source, axis=0, return_index=True)
            Tx = source[np.sort(ind), :]
            t0 = np.zeros((Tx.shape[0],))
            nTx = Tx.shape[0]
        elif source.shape[1] == 3:
            src = source[:, 1:3]
            _, ind = np.unique(source, axis=0, return_index=True)
            tmp = source[np.sort(ind), :]
            nTx = tmp.shape[0]
            Tx, t0 = tmp[:, 1:3], tmp[:, 0]
        else:
            raise ValueError("source should be either nsrc x 2, 3 or 4")
        if src.shape[1] != 2 or rcv.shape[1] != 2:
            raise ValueError("src and rcv should be ndata x 2")
        if self.is_outside(src):
            raise ValueError("Source point outside grid")
        if self.is_outside(rcv):
            raise ValueError("Receiver outside grid")
        if slowness is not None:
            self.set_slowness(slowness)
        vTx, vt0, vRx, iRx = [], [], [], []
        if evID is None:
            if nTx == 1:
                vTx.append(src[0:1, :]); vt0.append(t0[0:1]); vRx.append(rcv); iRx.append(np.arange(rcv.shape[0]))
            elif aggregate_src:
                vTx.append(Tx); vt0.append(t0); vRx.append(rcv); iRx.append(np.arange(rcv.shape[0]))
                nTx = 1
            else:
                if src.shape != rcv.shape:
                    raise ValueError("src and rcv should be of equal size")
                for n in range(nTx):
                    ind = np.sum(Tx[n, :] == src, axis=1) == 2
                    iRx.append(np.nonzero(ind)[0])
                    vTx.append(Tx[n:n + 1, :]); vt0.append(t0[n:n + 1]); vRx.append(rcv[ind, :])
        else:
            if src.shape != rcv.shape:
                raise ValueError("src and rcv should be of equal size")
            for n in range(nTx):
                i0 = int(np.nonzero(evID == eid[n])[0][0])
                vTx.append(src[i0:i0 + 1, :]); vt0.append(t0[i0:i0 + 1])
                ii = np.nonzero(evID == eid[n])[0]
                iRx.append(ii); vRx.append(rcv[ii, :])
        tt = np.zeros((rcv.shape[0],), dtype=self.dtype)
        if thread_no is not None:
            assert nTx == 1
            txa = np.ascontiguousarray(vTx[0], dtype=self.dtype)
            t0a = np.ascontiguousarray(vt0[0], dtype=self.dtype)
            rxa = np.ascontiguousarray(vRx[0], dtype=self.dtype)
            out = np.empty(rxa.shape[0], dtype=self.dtype)
            self._chk(self._lib.ttcr_b200_raytrace2d(self._h, txa.ctypes.data, t0a.ctypes.data, txa.shape[0], rxa.ctypes.data, rxa.shape[0],
                                                     out.ctypes.data, int(thread_no)))
            tt[iRx[0]] = out
            return tt
        tx_off = np.zeros(nTx + 1, dtype=np.uintp)
        rx_off = np.zeros(nTx + 1, dtype=np.uintp)
        tx_off[1:] = np.cumsum([v.shape[0] for v in vTx])
        rx_off[1:] = np.cumsum([v.shape[0] for v in vRx])
        txa = np.ascontiguousarray(np.vstack(vTx), dtype=self.dtype)
        t0a = np.ascontiguousarray(np.concatenate(vt0), dtype=self.dtype)
        rxa = np.ascontiguousarray(np.vstack(vRx), dtype=self.dtype)
        out = np.empty(rxa.shape[0], dtype=self.dtype)
        self._niter_all = np.zeros((nTx, 2), dtype=np.int32)
        self._chk(self._lib.ttcr_b200_raytrace2d_multi(self._h, nTx, tx_off.ctypes.data, txa.ctypes.data, t0a.ctypes.data, rx_off.ctypes.data,
                                                       rxa.ctypes.data, out.ctypes.data, self._niter_all.ctypes.data))
        for n in range(nTx):
            tt[iRx[n]] = out[rx_off[n]:rx_off[n + 1]]
        return tt
