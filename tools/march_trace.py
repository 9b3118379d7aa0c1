"""Summarise a TTCR_B200_TRACE dump of k_sweep_march (debug aid): per-tile ticket / start / end times.
    TTCR_B200_TRACE=gpurun_out/x.trace python tools/march_check.py one 256   ->   python tools/march_trace.py gpurun_out/x.trace [sweep]
Record per tile: [0] ticket taken, [1] first chunk landed (march starts), [5] march done, [7] SM id (globaltimer, ns)."""
import sys

import numpy as np

raw = open(sys.argv[1], "rb").read()
pos, k = 0, 0
while pos < len(raw):
    ntiles, nU, nV, PUT = np.frombuffer(raw, dtype=np.int32, count=4, offset=pos)
    pos += 16
    RL, PUT = max(8, PUT // 1000), PUT % 1000
    t = np.frombuffer(raw, dtype=np.int64, count=ntiles * RL, offset=pos).reshape(ntiles, RL).astype(np.float64)
    pos += ntiles * RL * 8
    k += 1
    if len(sys.argv) > 2 and k != int(sys.argv[2]):
        continue
    t0 = t[:, 0].min()
    T = (t.reshape(nU, nV, RL) - t0) / 1e3
    tk, st, en = T[:, :, 0], T[:, :, 1], T[:, :, 5]
    print(f"sweep {k}: tiles {ntiles} (nU {nU} x nV {nV}, PUT {PUT}); span {en.max():.1f} us; SMs used {len(np.unique(t[:, 7]))}")
    run = en - st
    print(f"  tile (0,0): ticket {tk[0,0]:.1f} start {st[0,0]:.1f} end {en[0,0]:.1f} run {run[0,0]:.1f} us")
    print(f"  run per tile: mean {run.mean():.1f} min {run.min():.1f} max {run.max():.1f}; wait before start mean {(st - tk).mean():.1f} max {(st - tk).max():.1f}")
    print("  run time, V = 0 column:", run[:, 0].round(0)[:: max(1, nU // 8)])
    print("  run time, U = 0 row   :", run[0].round(0))
    du = np.diff(en, axis=0)
    dv = np.diff(en, axis=1)
    print(f"  end-to-end lag per U-hop: mean {du.mean():.2f} us (V=0: {du[:, 0].mean():.2f}); per V-hop: mean {dv.mean():.2f} (U=0: {dv[0].mean():.2f})")
    print(f"  last tile: ticket {tk[-1,-1]:.1f} start {st[-1,-1]:.1f} end {en[-1,-1]:.1f}")
    print("  end times, V = 0 column:", en[:, 0].round(0)[:: max(1, nU // 8)])
    print("  end times, last U row  :", en[-1].round(0))
    if T[:, :, 2].any():   # quarter times (step-trace builds)
        for name, i in (("1/4", 2), ("1/2", 3), ("3/4", 4), ("end", 5)):
            x = T[:, :, i]
            print(f"  {name}: V=0 column {x[:, 0].round(0)[:6]} ... U-hop lag mean {np.diff(x[:, 0]).mean():.2f}; U=0 row {x[0].round(0)[:6]} ... V-hop lag mean {np.diff(x[0]).mean():.2f}")
