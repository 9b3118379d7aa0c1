"""Development probe: aggregate throughput with S sources solved concurrently (one slot / CUDA stream each)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from ttcr_b200 import Grid3d

n = int(sys.argv[1]); kernel = int(sys.argv[2]) if len(sys.argv) > 2 else 3
opts = dict(kv.split("=") for kv in sys.argv[3:])
x = np.linspace(0.0, 20.0, n)
s = np.ascontiguousarray(np.broadcast_to((1.0 / (1.0 + 0.1 * x))[None, None, :], (n, n, n)), dtype=np.float32)
rng = np.random.default_rng(12345)
for S in (1, 2, 3, 4, 6):
    g = Grid3d(x, x, x, n_threads=S, cell_slowness=0, tt_from_rp=False, weno=0, dtype=np.float32)
    g.set_option("kernel", kernel)
    for k_, v_ in opts.items():
        g.set_option(k_, float(v_))
    g.set_slowness(s)
    src = rng.uniform(0.5, 19.5, (S, 3)); rcv = src.copy()
    g.raytrace(src, rcv)                      # warm-up
    t = time.perf_counter()
    reps = 3
    for _ in range(reps):
        g.raytrace(src, rcv)
    dt = (time.perf_counter() - t) / reps
    sweeps = sum(g.get_stats(i)["sweeps"] for i in range(S))
    print(f"n={n} kernel={kernel} S={S}: {dt*1e3:.1f} ms per batch, {n**3 * sweeps / dt / 1e6:.0f} Mnodes/s aggregate, "
          f"niter {[g.get_niter(i)[0] for i in range(S)]}", flush=True)
    g.close()
