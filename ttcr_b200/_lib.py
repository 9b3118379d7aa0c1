"""ctypes binding of ``libttcr_b200.so`` (the C ABI declared in ``include/ttcr_b200.h``).

There is no CPU implementation behind this module: if the shared library is missing the
import fails loudly, and if no CUDA device is present ``ttcr_b200_create`` returns
``TTCR_B200_ERR_CUDA``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# development only: TTCR_B200_LIB points at another build of the same library (e.g. one compiled with step tracing)
LIB_PATH = os.environ.get("TTCR_B200_LIB") or os.path.join(_HERE, "libttcr_b200.so")

OK, ERR_RUNTIME, ERR_LENGTH, ERR_LOGIC, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = range(7)
F64, F32 = 0, 1
ORDER_X_FASTEST, ORDER_Z_FASTEST = 0, 1
KERNEL_AUTO, KERNEL_PLANE, KERNEL_TILE, KERNEL_TILE3 = 0, 1, 2, 3

# every symbol include/ttcr_b200.h declares (tests check that the library exports all of them)
SYMBOLS = (
    "ttcr_b200_create", "ttcr_b200_destroy", "ttcr_b200_last_error", "ttcr_b200_set_slowness",
    "ttcr_b200_set_slowness_device", "ttcr_b200_set_slowness_device_planes", "ttcr_b200_get_tt_device", "ttcr_b200_get_slowness", "ttcr_b200_raytrace", "ttcr_b200_raytrace_multi", "ttcr_b200_get_tt",
    "ttcr_b200_get_niter", "ttcr_b200_set_option", "ttcr_b200_n_slots", "ttcr_b200_solve",
    "ttcr_b200_get_stats", "ttcr_b200_device_bytes", "ttcr_b200_version", "ttcr_b200_raytrace_rays", "ttcr_b200_get_rays", "ttcr_b200_get_m_terms",
    "ttcr_b200_create2d", "ttcr_b200_destroy2d", "ttcr_b200_set_slowness2d", "ttcr_b200_get_slowness2d", "ttcr_b200_raytrace2d",
    "ttcr_b200_raytrace2d_multi", "ttcr_b200_get_tt2d", "ttcr_b200_get_niter2d", "ttcr_b200_last_solve_ms2d",
)


class Stats(C.Structure):
    _fields_ = [("niter", C.c_int), ("niterw", C.c_int), ("solve_ms", C.c_double), ("sweep_ms", C.c_double),
                ("launches", C.c_longlong), ("sweep_launches", C.c_longlong), ("sweeps", C.c_int), ("last_change", C.c_double),
                ("kernel", C.c_int)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CudaError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library (raises if it has not been built: run ``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(make -C ttcr_b200/csrc).  ttcr_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, sz, dbl, i32 = C.c_void_p, C.c_size_t, C.c_double, C.c_int
    lib.ttcr_b200_version.restype = C.c_char_p
    lib.ttcr_b200_last_error.restype = C.c_char_p
    lib.ttcr_b200_last_error.argtypes = [vp]
    lib.ttcr_b200_create.argtypes = [C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_uint32, dbl, dbl, dbl, dbl, dbl, i32,
                                     i32, i32, i32, sz, i32, i32, i32, i32]
    lib.ttcr_b200_destroy.argtypes = [vp]
    lib.ttcr_b200_destroy.restype = None
    lib.ttcr_b200_set_slowness.argtypes = [vp, vp, sz, i32]
    lib.ttcr_b200_set_slowness_device.argtypes = [vp, vp, sz, i32]
    lib.ttcr_b200_set_slowness_device_planes.argtypes = [vp, vp, sz, i32, i32]
    lib.ttcr_b200_get_tt_device.argtypes = [vp, vp, sz, i32]
    lib.ttcr_b200_get_slowness.argtypes = [vp, vp, i32]
    lib.ttcr_b200_raytrace.argtypes = [vp, vp, vp, sz, vp, sz, vp, sz]
    lib.ttcr_b200_raytrace_rays.argtypes = [vp, vp, vp, sz, vp, sz, vp, vp, sz]
    lib.ttcr_b200_get_rays.argtypes = [vp, sz, vp]
    lib.ttcr_b200_get_m_terms.argtypes = [vp, sz, vp, vp]
    lib.ttcr_b200_raytrace_multi.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.ttcr_b200_get_tt.argtypes = [vp, vp, sz, i32]
    lib.ttcr_b200_get_niter.argtypes = [vp, sz, C.POINTER(i32), C.POINTER(i32)]
    lib.ttcr_b200_set_option.argtypes = [vp, C.c_char_p, dbl]
    lib.ttcr_b200_n_slots.argtypes = [vp]
    lib.ttcr_b200_n_slots.restype = sz
    lib.ttcr_b200_solve.argtypes = [vp, vp, vp, sz, sz]
    lib.ttcr_b200_get_stats.argtypes = [vp, sz, C.POINTER(Stats)]
    lib.ttcr_b200_device_bytes.argtypes = [vp]
    lib.ttcr_b200_device_bytes.restype = sz
    lib.ttcr_b200_create2d.argtypes = [C.POINTER(vp), C.c_uint32, C.c_uint32, dbl, dbl, dbl, dbl, dbl, i32, i32, i32, sz, i32, i32, i32]
    lib.ttcr_b200_destroy2d.argtypes = [vp]
    lib.ttcr_b200_destroy2d.restype = None
    lib.ttcr_b200_set_slowness2d.argtypes = [vp, vp, sz]
    lib.ttcr_b200_get_slowness2d.argtypes = [vp, vp]
    lib.ttcr_b200_raytrace2d.argtypes = [vp, vp, vp, sz, vp, sz, vp, sz]
    lib.ttcr_b200_raytrace2d_multi.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, vp]
    lib.ttcr_b200_get_tt2d.argtypes = [vp, vp, sz]
    lib.ttcr_b200_get_niter2d.argtypes = [vp, sz, C.POINTER(i32), C.POINTER(i32)]
    lib.ttcr_b200_last_solve_ms2d.argtypes = [vp]
    lib.ttcr_b200_last_solve_ms2d.restype = dbl
    _lib = lib
    return lib


def check(rc: int, handle=None) -> None:
    """Turn a status code into the exception the reference would raise through Cython's ``except +``
    (std::runtime_error -> RuntimeError, std::length_error -> ValueError? no: Cython maps
    length_error to ... ``ValueError`` is what rgrid.pyx's own pre-check raises, so use that)."""
    if rc == OK:
        return
    msg = load().ttcr_b200_last_error(handle).decode(errors="replace")
    if rc == ERR_RUNTIME:
        raise RuntimeError(msg)
    if rc == ERR_LENGTH:
        raise ValueError(msg)
    if rc == ERR_LOGIC:
        raise ArithmeticError(msg) if "WENO" in msg else RuntimeError(msg)
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise CudaError(msg)
