import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names(pattern="*"):
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, pattern + ".npz")))


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as f:
        d = {k: f[k] for k in f.files}
    for k in ("niter", "niterw", "weno", "maxit"):
        d[k] = int(d[k])
    d["cell_slowness"] = bool(d["cell_slowness"])
    d["translate"] = bool(d.get("translate", False))
    d["eps"] = float(d["eps"])
    d["dtype"] = np.dtype(np.float32 if name.endswith("float32") else np.float64)
    return d


def golden_ray_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "rays", "*.npz")))


def load_golden_rays(name):
    """the reference's raypaths for golden case `name` (oracle/make_golden.py::raypath_case): receivers, traveltimes along
    the rays, and the rays as a list of (npts, 3) arrays"""
    with np.load(os.path.join(GOLDEN_DIR, "rays", name + ".npz")) as f:
        d = {k: f[k] for k in f.files}
    ends = np.cumsum(d["rp_npts"])
    d["rays"] = [d["rp_xyz"][e - n:e] for n, e in zip(d["rp_npts"], ends)]
    return d


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O
