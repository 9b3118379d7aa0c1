"""ttcr_b200 -- B200-native 3D rectilinear fast-sweeping eikonal solver.

A from-scratch CUDA (sm_100a) implementation of ONE hot path of groupeLIAMG/ttcr -- the
classes Grid3Drnfs / Grid3Drcfs behind ``ttcrpy.rgrid.Grid3d`` -- exposed through a C ABI
(``include/ttcr_b200.h``, ``libttcr_b200.so``) and this thin Python mirror of the reference's
interface, plus the 2-D twins Grid2Drnfs / Grid2Drcfs behind ``ttcrpy.rgrid.Grid2d`` (FSM branch).  See DESIGN.md
and INTEGRATION.md.
"""
from . import _lib  # noqa: F401
from .rgrid import Grid3d, Grid3d_d, Grid3d_f  # noqa: F401
from .rgrid2d import Grid2d  # noqa: F401
from .vtr import read_vtr, write_vtr  # noqa: F401

__all__ = ["Grid3d", "Grid3d_d", "Grid3d_f", "Grid2d", "read_vtr", "write_vtr"]
