"""TEST INFRASTRUCTURE: digests of the CPU oracle at the headline sizes (BASELINE.json configs[2..4]).

The full fields are 0.5 - 4 GiB, so what is committed under tests/golden/digest/ is, per case: sha256 of the fp32 field
(numpy (nx, ny, nz) C order), niter, three full i-planes, the traveltimes at the receivers, and -- where the reference's
own Grid3Drnfs<double> fits in this container's memory -- the same planes and receiver times from the UNMODIFIED reference
in double (the "reference CPU Grid3Drnfs" the north star's 1e-4 is stated against).

    python oracle/make_digest.py c3        # 512^3 gradient model, corner source: C restatement fp32 (~8 min of CPU)
    python oracle/make_digest.py c3ref     # the same through oracle/_ref, Grid3Drnfs<double> (~35 GB of RAM)
    python oracle/make_digest.py c4        # 511^3 cells (Grid3Drcfs averaging), 4 of the 64 sources, restatement fp32
    python oracle/make_digest.py c5        # 1024^3 gradient model, one source, restatement fp32 (~1 h of CPU, 12 GB)

The models and sources are the ones tests/test_gpu_fullsize.py and tools/configs45.py build.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "digest")


def gradient(n, dtype=np.float32):
    x = np.linspace(0.0, 20.0, n)
    s = np.ascontiguousarray(np.broadcast_to((1.0 / (1.0 + 0.1 * x))[None, None, :], (n, n, n)), dtype=dtype)
    return x, s


def rcv_pattern():
    """the 21 x 21 receiver pattern of the reference's tests/files/rcv.dat (x = 0..20, y = 0..20, z = 0)"""
    g = np.arange(21.0)
    X, Y = np.meshgrid(g, g, indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), np.zeros(X.size)], axis=1)


def config4_model(n=512):
    x = np.linspace(0.0, 20.0, n)
    rng = np.random.default_rng(12345)
    zc = 0.5 * (x[1:] + x[:-1])
    sc = ((1.0 / (1.0 + 0.1 * zc))[None, None, :] * np.exp(0.05 * rng.standard_normal((n - 1, n - 1, n - 1), dtype=np.float32))).astype(np.float32)
    src_off = rng.uniform(0.5, 19.5, (64, 3))
    dx = x[1] - x[0]
    src_on = np.round(src_off / dx) * dx
    return x, sc, src_on, src_off


def digest(field, planes):
    f = np.ascontiguousarray(field)
    return dict(sha256=np.frombuffer(hashlib.sha256(f.tobytes()).digest(), dtype=np.uint8), planes_i=np.asarray(planes),
                planes=np.ascontiguousarray(f[planes]))


def save(name, **kw):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print(f"{name}: niter={kw.get('niter')} {os.path.getsize(path) / 2**20:.1f} MiB", flush=True)


def restated(x, s_node, src, rcv, eps=1e-5, maxit=50):
    n = x.size
    dx = float(np.float32(x[1]) - np.float32(x[0]))
    t0 = time.time()
    tt, ni, nw = O.solve(n - 1, n - 1, n - 1, dx, O.to_cxx(s_node), src, eps=eps, maxit=maxit, weno=False, dtype=np.float32)
    sec = time.time() - t0
    tt_rcv = O.interp(n - 1, n - 1, n - 1, dx, tt, rcv, dtype=np.float32)
    return O.from_cxx(tt, (n, n, n)), ni, tt_rcv, sec


def c3():
    n = 512
    x, s = gradient(n)
    src = np.array([[0.0, 0.0, 0.0]])
    rcv = rcv_pattern()
    f, ni, tt_rcv, sec = restated(x, s, src, rcv)
    save("c3_512_f32", niter=ni, src=src, rcv=rcv, tt_rcv=tt_rcv, seconds=sec, **digest(f, [0, 255, 511]))


def c3ref():
    n = 512
    x, s = gradient(n, np.float64)
    src = np.array([[0.0, 0.0, 0.0]])
    rcv = rcv_pattern()
    dx = float(x[1] - x[0])
    g = O.RefGrid(n - 1, n - 1, n - 1, dx, eps=1e-5, maxit=50, weno=False, cell_slowness=False, dtype=np.float64)
    g.set_slowness(O.to_cxx(s))
    del s
    tt_rcv, sec = g.raytrace(src, 0.0, rcv)
    f = O.from_cxx(g.get_tt(), (n, n, n))
    ni, _ = g.niter()
    g.close()
    planes = [0, 255, 511]
    save("c3_512_ref_f64", niter=ni, src=src, rcv=rcv, tt_rcv=tt_rcv, seconds=sec, planes_i=np.asarray(planes), planes=np.ascontiguousarray(f[planes]))


def c4():
    n = 512
    x, sc, src_on, src_off = config4_model(n)
    rcv = np.array([[1.0, 1.0, 1.0], [19.0, 19.0, 19.0], [10.0, 10.0, 0.0], [3.3, 16.2, 8.7]])
    sn = O.from_cxx(O.cell_to_node(O.to_cxx(sc), n - 1, n - 1, n - 1, dtype=np.float32), (n, n, n))
    out = dict(node_slowness_sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(sn).tobytes()).digest(), dtype=np.uint8))
    srcs = np.stack([src_on[0], src_on[37], src_off[0], src_off[37]])
    out["src"] = srcs
    out["rcv"] = rcv
    for k, sxyz in enumerate(srcs):
        f, ni, tt_rcv, sec = restated(x, sn, sxyz.reshape(1, 3), rcv)
        d = digest(f, [255])
        out[f"niter_{k}"] = ni
        out[f"tt_rcv_{k}"] = tt_rcv
        out[f"sha256_{k}"] = d["sha256"]
        out[f"planes_{k}"] = d["planes"]
        print("c4 source", k, "niter", ni, f"{sec:.0f} s", flush=True)
    out["planes_i"] = np.asarray([255])
    save("c4_511c_f32", **out)


def c5():
    n = 1024
    x, s = gradient(n)
    dxn = x[1] - x[0]
    src = np.round(np.array([[5.0, 5.0, 5.0]]) / dxn) * dxn
    rcv = np.array([[20.0, 20.0, 20.0], [0.0, 0.0, 0.0], [10.0, 3.0, 17.0]])
    f, ni, tt_rcv, sec = restated(x, s, src, rcv)
    save("c5_1024_f32", niter=ni, src=src, rcv=rcv, tt_rcv=tt_rcv, seconds=sec, **digest(f, [511]))


if __name__ == "__main__":
    for a in sys.argv[1:]:
        {"c3": c3, "c3ref": c3ref, "c4": c4, "c5": c5}[a]()
