"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/ttcr_b200.h
declares, and refuses to run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_cuda


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "ttcr_b200.h")) as f:
        txt = f.read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ttcr_b200_[a-z_0-9]+)\s*\(", txt)))


def test_header_and_binding_agree():
    from ttcr_b200 import _lib
    assert _declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from ttcr_b200 import _lib
    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.ttcr_b200_version()


def test_product_does_not_touch_the_oracle():
    """the product path must not import, link or call anything under oracle/"""
    pkg = os.path.join(ROOT, "ttcr_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or fn == "Makefile":
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                assert "oracle" not in txt.lower().replace("no oracle", ""), os.path.join(dirpath, fn)


@pytest.mark.skipif(has_cuda(), reason="only meaningful on a machine without a GPU")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    from ttcr_b200 import Grid3d, _lib
    x = np.arange(5.0)
    with pytest.raises(_lib.CudaError):
        Grid3d(x, x, x, cell_slowness=0, tt_from_rp=0)


def test_null_handle_is_rejected():
    from ttcr_b200 import _lib
    lib = _lib.load()
    assert lib.ttcr_b200_set_slowness(None, None, 0, 0) == _lib.ERR_INVALID
    assert lib.ttcr_b200_n_slots(None) == 0
    assert lib.ttcr_b200_create(None, 1, 1, 1, 1.0, 0, 0, 0, 1e-5, 1, 0, 0, 0, 1, 0, 0, 0, -1) == _lib.ERR_INVALID


def test_argument_checks_before_the_device():
    from ttcr_b200 import Grid3d
    x = np.arange(5.0)
    with pytest.raises(ValueError, match="cubic"):
        Grid3d(x, 2 * x, x)
    with pytest.raises(ValueError, match="undefined"):
        Grid3d(x, x, x, method="XYZ")
    with pytest.raises(NotImplementedError):
        Grid3d(x, x, x, method="SPM")
    with pytest.raises(ValueError, match="dtype"):
        Grid3d(x, x, x, dtype=np.int32)


def test_vtr_roundtrip(tmp_path):
    from ttcr_b200 import read_vtr, write_vtr
    x, y, z = np.linspace(0, 1, 4), np.linspace(0, 2, 5), np.linspace(-1, 1, 6)
    pd = {"Slowness": np.random.default_rng(0).random(4 * 5 * 6), "f": np.arange(120, dtype=np.float32)}
    cd = {"c": np.arange(3 * 4 * 5, dtype=np.float64)}
    for compress in (True, False):
        fn = str(tmp_path / f"t{int(compress)}.vtr")
        write_vtr(fn, x, y, z, pd, cd, compress=compress)
        d = read_vtr(fn)
        assert np.array_equal(d["x"], x) and np.array_equal(d["y"], y) and np.array_equal(d["z"], z)
        for k in pd:
            assert np.array_equal(d["point_data"][k], pd[k]) and d["point_data"][k].dtype == pd[k].dtype
        assert np.array_equal(d["cell_data"]["c"], cd["c"])
