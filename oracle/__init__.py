"""TEST INFRASTRUCTURE ONLY -- ctypes access to the CPU oracle.

* ``solve`` / ``cell_to_node`` / ``interp`` : the plain-C restatement
  (``oracle/fsm_oracle.c``), always available once ``make -C oracle`` ran.
* ``RefGrid`` : the UNMODIFIED reference (``ttcr/Grid3Drnfs.h``, ``Grid3Drcfs.h``)
  behind ``oracle/ref_shim.cpp``; available when ``oracle/_ref/libttcr_ref.so``
  exists (built in the container where ``/root/reference`` is mounted; the
  prebuilt ``.so`` travels to the GPU box).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``ttcr_b200`` never does.

Array conventions are the reference's C++ ones: flat, x fastest,
``n = (k*ny1 + j)*nx1 + i`` (``ttcr/Grid3Drn.h:2823``).  Helpers ``to_cxx`` /
``from_cxx`` convert from / to numpy ``(nx, ny, nz)`` C-order arrays the way
``src/ttcrpy/rgrid.pyx:559-566`` and ``:435`` do (flatten / reshape ``order='F'``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libfsm_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libttcr_ref.so")


def build(quiet: bool = True) -> None:
    """(Re)build the oracle libraries with ``make`` (compiling the checker, not using it)."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


_lib = None
_ref = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
    return _lib


def have_ref() -> bool:
    return os.path.exists(_REF)


def _load_ref():
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libttcr_ref.so not built (needs /root/reference)")
        lib = C.CDLL(_REF)
        lib.ttcr_ref_last_error.restype = C.c_char_p
        lib.ttcr_ref_create.restype = C.c_void_p
        lib.ttcr_ref_create.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double,
                                        C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int]
        lib.ttcr_ref_destroy.argtypes = [C.c_void_p]
        lib.ttcr_ref_nnodes.restype = C.c_size_t
        lib.ttcr_ref_nnodes.argtypes = [C.c_void_p]
        dp = C.POINTER(C.c_double)
        lib.ttcr_ref_set_slowness.argtypes = [C.c_void_p, dp, C.c_size_t]
        lib.ttcr_ref_get_slowness.argtypes = [C.c_void_p, dp]
        lib.ttcr_ref_raytrace.argtypes = [C.c_void_p, dp, dp, C.c_size_t, dp, C.c_size_t, dp, C.c_size_t, dp]
        lib.ttcr_ref_raytrace_multi.argtypes = [C.c_void_p, C.c_size_t, dp, dp, dp, C.c_size_t, dp, dp]
        lib.ttcr_ref_get_tt.argtypes = [C.c_void_p, dp, C.c_size_t]
        if hasattr(lib, "ttcr_ref_raytrace_rays"):
            lib.ttcr_ref_raytrace_rays.argtypes = [C.c_void_p, dp, dp, C.c_size_t, dp, C.c_size_t, dp, C.c_size_t, C.c_void_p, dp,
                                                   C.c_size_t]
        if hasattr(lib, "ttcr_ref_raytrace_m"):
            lib.ttcr_ref_raytrace_m.argtypes = [C.c_void_p, dp, dp, C.c_size_t, dp, C.c_size_t, dp, C.c_size_t, C.c_void_p, C.c_void_p, dp,
                                                C.c_size_t, C.c_int]
        lib.ttcr_ref_get_niter.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _ref = lib
    return _ref


def to_cxx(a: np.ndarray) -> np.ndarray:
    """numpy (nx,ny,nz) C-order -> flat x-fastest (rgrid.pyx:559)."""
    return np.ascontiguousarray(np.asarray(a).flatten(order="F"))


def from_cxx(v: np.ndarray, shape) -> np.ndarray:
    """flat x-fastest -> numpy (nx,ny,nz) (rgrid.pyx:435)."""
    return np.asarray(v).reshape(shape, order="F")


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class RefGrid:
    """The reference's own Grid3Drnfs / Grid3Drcfs (double or float)."""

    def __init__(self, ncx, ncy, ncz, dx, xmin=0.0, ymin=0.0, zmin=0.0, eps=1e-5, maxit=50, weno=True,
                 cell_slowness=False, dtype=np.float64, tt_from_rp=False, n_threads=1, translate_grid=False, interp_vel=False):
        lib = _load_ref()
        self._lib = lib
        self.dtype = np.dtype(dtype)
        self.shape = (ncx + 1, ncy + 1, ncz + 1)
        self._h = lib.ttcr_ref_create(0 if self.dtype == np.float64 else 1, int(bool(cell_slowness)), ncx, ncy,
                                      ncz, dx, xmin, ymin, zmin, eps, maxit, int(bool(weno)),
                                      int(bool(tt_from_rp)) | (2 if interp_vel else 0), n_threads, int(bool(translate_grid)))
        if not self._h:
            raise RuntimeError(lib.ttcr_ref_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ttcr_ref_destroy(self._h)
            self._h = None

    __del__ = close

    def _chk(self, rc):
        if rc:
            msg = self._lib.ttcr_ref_last_error().decode()
            raise {1: RuntimeError, 2: ValueError, 3: ArithmeticError}.get(rc, RuntimeError)(msg)

    def set_slowness(self, s_flat):
        s = np.ascontiguousarray(s_flat, dtype=np.float64).ravel()
        self._chk(self._lib.ttcr_ref_set_slowness(self._h, _dp(s), s.size))

    def get_slowness(self):
        out = np.empty(self._lib.ttcr_ref_nnodes(self._h))
        self._chk(self._lib.ttcr_ref_get_slowness(self._h, _dp(out)))
        return out

    def raytrace(self, tx, t0, rx, thread_no=0):
        """returns (tt at rx, seconds)"""
        tx = np.ascontiguousarray(tx, dtype=np.float64).reshape(-1, 3)
        t0 = np.ascontiguousarray(np.broadcast_to(np.asarray(t0, dtype=np.float64), (tx.shape[0],)))
        rx = np.ascontiguousarray(rx, dtype=np.float64).reshape(-1, 3)
        tt = np.empty(rx.shape[0])
        sec = C.c_double()
        self._chk(self._lib.ttcr_ref_raytrace(self._h, _dp(tx), _dp(t0), tx.shape[0], _dp(rx), rx.shape[0],
                                              _dp(tt), thread_no, C.byref(sec)))
        return tt, sec.value

    def raytrace_rays(self, tx, t0, rx, thread_no=0, cap=4096):
        """Grid3D::raytrace(..., r_data, threadNo): returns (tt at rx, list of (npts, 3) arrays)"""
        tx = np.ascontiguousarray(tx, dtype=np.float64).reshape(-1, 3)
        t0 = np.ascontiguousarray(np.broadcast_to(np.asarray(t0, dtype=np.float64), (tx.shape[0],)))
        rx = np.ascontiguousarray(rx, dtype=np.float64).reshape(-1, 3)
        tt = np.empty(rx.shape[0])
        npts = np.zeros(rx.shape[0], dtype=np.uintp)
        xyz = np.zeros((rx.shape[0], cap, 3))
        self._chk(self._lib.ttcr_ref_raytrace_rays(self._h, _dp(tx), _dp(t0), tx.shape[0], _dp(rx), rx.shape[0], _dp(tt),
                                                   thread_no, npts.ctypes.data, _dp(xyz), cap))
        assert npts.max(initial=0) <= cap
        return tt, [xyz[i, :int(n)].copy() for i, n in enumerate(npts)]

    def raytrace_m(self, tx, t0, rx, thread_no=0, cap=65536, with_rays=False):
        """Grid3D::raytrace(..., m_data, threadNo) (Grid3D.h:743-780) or, with_rays, (..., r_data, m_data, threadNo) (:646-690):
        returns (tt at rx, list of (columns uint64, values) per receiver, in the order the reference leaves them in m_data)"""
        tx = np.ascontiguousarray(tx, dtype=np.float64).reshape(-1, 3)
        t0 = np.ascontiguousarray(np.broadcast_to(np.asarray(t0, dtype=np.float64), (tx.shape[0],)))
        rx = np.ascontiguousarray(rx, dtype=np.float64).reshape(-1, 3)
        tt = np.empty(rx.shape[0])
        nnz = np.zeros(rx.shape[0], dtype=np.uintp)
        col = np.zeros((rx.shape[0], cap), dtype=np.uint64)
        val = np.zeros((rx.shape[0], cap))
        self._chk(self._lib.ttcr_ref_raytrace_m(self._h, _dp(tx), _dp(t0), tx.shape[0], _dp(rx), rx.shape[0], _dp(tt), thread_no,
                                                nnz.ctypes.data, col.ctypes.data, _dp(val), cap, int(bool(with_rays))))
        assert nnz.max(initial=0) <= cap
        return tt, [(col[i, :int(n)].copy(), val[i, :int(n)].copy()) for i, n in enumerate(nnz)]

    def raytrace_multi(self, tx, t0, rx):
        """one Tx point per source; same receivers for all; the reference's own thread fan-out."""
        tx = np.ascontiguousarray(tx, dtype=np.float64).reshape(-1, 3)
        t0 = np.ascontiguousarray(np.broadcast_to(np.asarray(t0, dtype=np.float64), (tx.shape[0],)))
        rx = np.ascontiguousarray(rx, dtype=np.float64).reshape(-1, 3)
        tt = np.empty((tx.shape[0], rx.shape[0]))
        sec = C.c_double()
        self._chk(self._lib.ttcr_ref_raytrace_multi(self._h, tx.shape[0], _dp(tx), _dp(t0), _dp(rx),
                                                    rx.shape[0], _dp(tt), C.byref(sec)))
        return tt, sec.value

    def get_tt(self, thread_no=0):
        out = np.empty(self._lib.ttcr_ref_nnodes(self._h))
        self._chk(self._lib.ttcr_ref_get_tt(self._h, _dp(out), thread_no))
        return out.astype(self.dtype)

    def niter(self):
        a, b = C.c_int(), C.c_int()
        self._chk(self._lib.ttcr_ref_get_niter(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "_d", C.c_double
    if dtype == np.float32:
        return "_f", C.c_float
    raise ValueError(dtype)


def cell_to_node(s_cell_flat, ncx, ncy, ncz, dtype=np.float64):
    """Grid3Drcfs::setSlowness (Grid3Drcfs.h:88-171) on flat x-fastest arrays."""
    sfx, ct = _sfx(dtype)
    lib = _load()
    sc = np.ascontiguousarray(s_cell_flat, dtype=dtype).ravel()
    if sc.size != ncx * ncy * ncz:
        raise ValueError("Error: slowness vectors of incompatible size.")
    sn = np.empty((ncx + 1) * (ncy + 1) * (ncz + 1), dtype=dtype)
    f = getattr(lib, "fsmo_cell_to_node" + sfx)
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]
    f(sc.ctypes.data, ncx, ncy, ncz, sn.ctypes.data)
    return sn


def solve(ncx, ncy, ncz, dx, s_node_flat, tx, t0=0.0, xmin=0.0, ymin=0.0, zmin=0.0, eps=1e-5, maxit=50,
          weno=False, dtype=np.float64, order=0):
    """Grid3Drnfs::raytrace (Grid3Drnfs.h:84-155) on flat arrays.

    returns (tt flat x-fastest, niter, niterw).  ``order=1`` visits nodes by diagonal level.
    """
    sfx, ct = _sfx(dtype)
    lib = _load()
    s = np.ascontiguousarray(s_node_flat, dtype=dtype).ravel()
    n = (ncx + 1) * (ncy + 1) * (ncz + 1)
    if s.size != n:
        raise ValueError("Error: slowness vectors of incompatible size.")
    tx = np.ascontiguousarray(np.asarray(tx, dtype=dtype).reshape(-1, 3))
    t0 = np.ascontiguousarray(np.broadcast_to(np.asarray(t0, dtype=dtype), (tx.shape[0],)))
    tt = np.empty(n, dtype=dtype)
    ni, nw = C.c_int(), C.c_int()
    f = getattr(lib, "fsmo_solve" + sfx)
    f.restype = C.c_int
    f.argtypes = [C.c_size_t] * 3 + [ct] * 5 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_size_t, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                             C.c_int]
    rc = f(ncx, ncy, ncz, dx, xmin, ymin, zmin, eps, maxit, int(bool(weno)), s.ctypes.data, tx.ctypes.data,
           t0.ctypes.data, tx.shape[0], tt.ctypes.data, C.byref(ni), C.byref(nw), order)
    if rc:
        raise RuntimeError("Error: Point outside grid.")
    return tt, ni.value, nw.value


def interp(ncx, ncy, ncz, dx, tt_flat, rx, xmin=0.0, ymin=0.0, zmin=0.0, dtype=np.float64):
    """Grid3Drn::getTraveltime (Grid3Drn.h:794-930) at receiver points."""
    sfx, ct = _sfx(dtype)
    lib = _load()
    tt = np.ascontiguousarray(tt_flat, dtype=dtype).ravel()
    rx = np.ascontiguousarray(np.asarray(rx, dtype=dtype).reshape(-1, 3))
    out = np.empty(rx.shape[0], dtype=dtype)
    f = getattr(lib, "fsmo_interp" + sfx)
    f.restype = None
    f.argtypes = [C.c_size_t] * 3 + [ct] * 4 + [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    f(ncx, ncy, ncz, dx, xmin, ymin, zmin, tt.ctypes.data, rx.ctypes.data, rx.shape[0], out.ctypes.data)
    return out


def tt_from_rp(ncx, ncy, ncz, dx, tt_flat, s_node_flat, tx, t0, rx, xmin=0.0, ymin=0.0, zmin=0.0, dtype=np.float64):
    """Grid3Drn::getTraveltimeFromRaypath (Grid3Drn.h:1103-1243) at receiver points, from a solved field."""
    sfx, ct = _sfx(dtype)
    lib = _load()
    tt = np.ascontiguousarray(tt_flat, dtype=dtype).ravel()
    sl = np.ascontiguousarray(s_node_flat, dtype=dtype).ravel()
    tx = np.ascontiguousarray(np.asarray(tx, dtype=dtype).reshape(-1, 3))
    t0 = np.ascontiguousarray(np.broadcast_to(np.asarray(t0, dtype=dtype), (tx.shape[0],)))
    rx = np.ascontiguousarray(np.asarray(rx, dtype=dtype).reshape(-1, 3))
    out = np.empty(rx.shape[0], dtype=dtype)
    f = getattr(lib, "fsmo_tt_from_rp" + sfx)
    f.restype = C.c_int
    f.argtypes = [C.c_size_t] * 3 + [ct] * 4 + [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    rc = f(ncx, ncy, ncz, dx, xmin, ymin, zmin, tt.ctypes.data, sl.ctypes.data, tx.ctypes.data, t0.ctypes.data, tx.shape[0],
           rx.ctypes.data, rx.shape[0], out.ctypes.data)
    if rc == 1:
        raise RuntimeError("Error while computing raypaths: going outside grid")
    if rc:
        raise RuntimeError("raypath did not reach a source")
    return out


def raypaths(ncx, ncy, ncz, dx, tt_flat, s_node_flat, tx, t0, rx, xmin=0.0, ymin=0.0, zmin=0.0, dtype=np.float64, cap=4096,
             interp_vel=False):
    """Grid3Drn::getRaypath (Grid3Drn.h:1339-1500) from a solved field: (traveltimes, list of (npts, 3) float64 arrays)."""
    sfx, ct = _sfx(dtype)
    lib = _load()
    tt = np.ascontiguousarray(tt_flat, dtype=dtype).ravel()
    sl = np.ascontiguousarray(s_node_flat, dtype=dtype).ravel()
    tx = np.ascontiguousarray(np.asarray(tx, dtype=dtype).reshape(-1, 3))
    t0 = np.ascontiguousarray(np.broadcast_to(np.asarray(t0, dtype=dtype), (tx.shape[0],)))
    rx = np.ascontiguousarray(np.asarray(rx, dtype=dtype).reshape(-1, 3))
    out = np.empty(rx.shape[0], dtype=dtype)
    npts = np.zeros(rx.shape[0], dtype=np.uintp)
    xyz = np.zeros((rx.shape[0], cap, 3), dtype=dtype)
    f = getattr(lib, "fsmo_raypaths" + sfx)
    f.restype = C.c_int
    f.argtypes = [C.c_size_t] * 3 + [ct] * 4 + [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                                                  C.c_void_p, C.c_size_t, C.c_int]
    rc = f(ncx, ncy, ncz, dx, xmin, ymin, zmin, tt.ctypes.data, sl.ctypes.data, tx.ctypes.data, t0.ctypes.data, tx.shape[0],
           rx.ctypes.data, rx.shape[0], out.ctypes.data, xyz.ctypes.data, npts.ctypes.data, cap, int(bool(interp_vel)))
    if rc == 1:
        raise RuntimeError("Error while computing raypaths: going outside grid")
    if rc:
        raise RuntimeError("raypath did not reach a source")
    assert npts.max(initial=0) <= cap
    return out, [xyz[i, :int(n)].astype(np.float64) for i, n in enumerate(npts)]


def slowness_at(ncx, ncy, ncz, dx, s_node_flat, pts, xmin=0.0, ymin=0.0, zmin=0.0, dtype=np.float64, interp_vel=False):
    """Grid3Drn::computeSlowness (Grid3Drn.h:2451-2676) at points (n, 3); s_node_flat in the reference's x-fastest order"""
    sfx, ct = _sfx(dtype)
    lib = _load()
    sl = np.ascontiguousarray(s_node_flat, dtype=dtype).ravel()
    pts = np.ascontiguousarray(np.asarray(pts, dtype=dtype).reshape(-1, 3))
    out = np.empty(pts.shape[0], dtype=dtype)
    f = getattr(lib, "fsmo_slowness_at" + sfx)
    f.restype = None
    f.argtypes = [C.c_size_t] * 3 + [ct] * 4 + [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    f(ncx, ncy, ncz, dx, xmin, ymin, zmin, sl.ctypes.data, pts.ctypes.data, pts.shape[0], int(bool(interp_vel)), out.ctypes.data)
    return out


# ---- 2-D twin (Grid2Drnfs): oracle for SURVEY section 8 row f4, not yet built in the product ------------------------
def solve2d(ncx, ncz, dx, dz, s_node, tx, t0=0.0, xmin=0.0, zmin=0.0, eps=1e-5, maxit=20, weno=False, rotated=False,
            dtype=np.float64):
    """Grid2Drnfs::raytrace restated (oracle/fsm2d_oracle.c): s_node (ncx+1, ncz+1) C order (z fastest), tx (n, 2) = (x, z).
    Returns (traveltime field (ncx+1, ncz+1), niter, niterw)."""
    sfx, ct = _sfx(dtype)
    lib = _load()
    s = np.ascontiguousarray(s_node, dtype=dtype).ravel()
    if s.size != (ncx + 1) * (ncz + 1):
        raise ValueError("Error: slowness vectors of incompatible size.")
    tx = np.ascontiguousarray(np.asarray(tx, dtype=dtype).reshape(-1, 2))
    t0 = np.ascontiguousarray(np.broadcast_to(np.asarray(t0, dtype=dtype), (tx.shape[0],)))
    tt = np.empty(s.size, dtype=dtype)
    ni, nw = C.c_int(), C.c_int()
    f = getattr(lib, "fsmo2d_solve" + sfx)
    f.restype = C.c_int
    f.argtypes = [C.c_size_t, C.c_size_t] + [ct] * 5 + [C.c_int] * 3 + [C.c_void_p] * 3 + [C.c_size_t, C.c_void_p,
                                                                                            C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rc = f(ncx, ncz, dx, dz, xmin, zmin, eps, maxit, int(bool(weno)), int(bool(rotated)), s.ctypes.data, tx.ctypes.data,
           t0.ctypes.data, tx.shape[0], tt.ctypes.data, C.byref(ni), C.byref(nw))
    if rc:
        raise RuntimeError("Error: Point outside grid.")
    return tt.reshape(ncx + 1, ncz + 1), ni.value, nw.value


def cell_to_node2d(s_cell, ncx, ncz, dtype=np.float64):
    """Grid2Drcfs::setSlowness (Grid2Drcfs.h:99-138): (ncx, ncz) cell slowness -> (ncx+1, ncz+1) node slowness"""
    sfx, _ = _sfx(dtype)
    lib = _load()
    s = np.ascontiguousarray(s_cell, dtype=dtype).ravel()
    if s.size != ncx * ncz:
        raise ValueError("Error: slowness vectors of incompatible size.")
    out = np.empty((ncx + 1) * (ncz + 1), dtype=dtype)
    f = getattr(lib, "fsmo2d_cell_to_node" + sfx)
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    f(s.ctypes.data, ncx, ncz, out.ctypes.data)
    return out.reshape(ncx + 1, ncz + 1)


def interp2d(ncx, ncz, dx, dz, tt, rx, xmin=0.0, zmin=0.0, dtype=np.float64):
    """Grid2Drn::getTraveltime (Grid2Drn.h:359-415) at receivers rx (n, 2)"""
    sfx, ct = _sfx(dtype)
    lib = _load()
    t = np.ascontiguousarray(tt, dtype=dtype).ravel()
    rx = np.ascontiguousarray(np.asarray(rx, dtype=dtype).reshape(-1, 2))
    out = np.empty(rx.shape[0], dtype=dtype)
    f = getattr(lib, "fsmo2d_interp" + sfx)
    f.restype = None
    f.argtypes = [C.c_size_t, C.c_size_t] + [ct] * 4 + [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    f(ncx, ncz, dx, dz, xmin, zmin, t.ctypes.data, rx.ctypes.data, rx.shape[0], out.ctypes.data)
    return out


def ref_solve2d(ncx, ncz, dx, dz, slowness, tx, t0=0.0, xmin=0.0, zmin=0.0, eps=1e-5, maxit=20, weno=False, rotated=False,
                dtype=np.float64, cell_slowness=False, rx=None):
    """the same solve through the UNMODIFIED reference (Grid2Drnfs / Grid2Drcfs <T, uint32_t, sxz<T>>, oracle/ref_shim.cpp);
    returns (field, niter, niterw) or, with receivers, (field, niter, niterw, receiver times)"""
    lib = _load_ref()
    dp = C.POINTER(C.c_double)
    lib.ttcr_ref2d_solve.argtypes = ([C.c_int, C.c_int, C.c_uint32, C.c_uint32] + [C.c_double] * 5 + [C.c_int] * 3 +
                                     [dp, C.c_size_t, dp, dp, C.c_size_t, dp, C.c_size_t, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)])
    s = np.ascontiguousarray(np.asarray(slowness, dtype=dtype), dtype=np.float64).ravel()
    tx = np.ascontiguousarray(np.asarray(tx, dtype=np.float64).reshape(-1, 2))
    t0 = np.ascontiguousarray(np.broadcast_to(np.asarray(t0, dtype=np.float64), (tx.shape[0],)))
    r = np.zeros((0, 2)) if rx is None else np.ascontiguousarray(np.asarray(rx, dtype=np.float64).reshape(-1, 2))
    tt = np.empty((ncx + 1) * (ncz + 1))
    ttr = np.empty(max(1, r.shape[0]))
    ni, nw = C.c_int(), C.c_int()
    rc = lib.ttcr_ref2d_solve(0 if np.dtype(dtype) == np.float64 else 1, int(bool(cell_slowness)), ncx, ncz, dx, dz, xmin, zmin, eps,
                              maxit, int(bool(weno)), int(bool(rotated)), _dp(s), s.size, _dp(tx), _dp(t0), tx.shape[0], _dp(r),
                              r.shape[0], _dp(tt), _dp(ttr), C.byref(ni), C.byref(nw))
    if rc:
        raise RuntimeError(lib.ttcr_ref_last_error().decode())
    field = tt.astype(dtype).reshape(ncx + 1, ncz + 1)
    if rx is None:
        return field, ni.value, nw.value
    return field, ni.value, nw.value, ttr[:r.shape[0]].astype(dtype)
