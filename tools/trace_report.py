"""Summarise a TTCR_B200_TRACE dump (debug aid): per-tile start / wait / run times of the tile kernel."""
import sys

import numpy as np

raw = open(sys.argv[1], "rb").read()
pos, k = 0, 0
while pos < len(raw):
    ntiles, nU, nV, NW = np.frombuffer(raw, dtype=np.int32, count=4, offset=pos)
    pos += 16
    t = np.frombuffer(raw, dtype=np.int64, count=ntiles * 8, offset=pos).reshape(ntiles, 8).astype(np.float64)
    pos += ntiles * 64
    k += 1
    if len(sys.argv) > 2 and k != int(sys.argv[2]):
        continue
    t0 = t[:, 0].min()
    tk, tp, q1, q2, q3, te = [(t[:, i] - t0) / 1e3 for i in range(6)]   # us
    print(f"sweep {k}: tiles {ntiles} (nU {nU} x nV {nV}, NW {NW}); span {te.max():.1f} us")
    print(f"  wait for deps at start (prologue): mean {np.mean(tp - tk):.1f} us, max {np.max(tp - tk):.1f}")
    print(f"  run (prologue done -> end): mean {np.mean(te - tp):.1f} us; quarter times mean {np.mean(q2 - q1):.1f} {np.mean(q3 - q2):.1f} us")
    T = t.reshape(nU, nV, 8)
    for V in (0, nV // 2, nV - 1):
        st = (T[:, V, 1] - t0) / 1e3
        en = (T[:, V, 5] - t0) / 1e3
        tkk = (T[:, V, 0] - t0) / 1e3
        d = np.diff(st)
        print(f"  V={V}: ticket[0..3] {tkk[:4].round(1)} start[0..5] {st[:6].round(1)} ... start lag per U-hop mean {d.mean():.2f} us (min {d.min():.2f} max {d.max():.2f}); end last {en[-1]:.1f}")
    for U in (0, nU // 2):
        st = (T[U, :, 1] - t0) / 1e3
        print(f"  U={U}: start over V {st.round(1)}")
    # progress-based lags: time at which each tile reached 1/4, 1/2, 3/4 of its rows
    for name, i in (("q1", 2), ("q2", 3), ("q3", 4), ("end", 5)):
        x = (T[:, :, i] - t0) / 1e3
        du = np.diff(x, axis=0)
        dv = np.diff(x, axis=1)
        print(f"  {name}: U-hop lag mean {du.mean():.2f} (V=0: {du[:, 0].mean():.2f}, p10 {np.percentile(du, 10):.2f}, p90 {np.percentile(du, 90):.2f});"
              f" V-hop lag mean {dv.mean():.2f} (U=0: {dv[0].mean():.2f})")
    run = (T[:, :, 5] - T[:, :, 1]) / 1e3
    print(f"  run time per tile: U=0 row {run[0].round(0)}; V=0 col [::8] {run[::8, 0].round(0)}")
    print(f"  tile (0,0): start {(T[0,0,1]-t0)/1e3:.1f} q1 {(T[0,0,2]-t0)/1e3:.1f} q2 {(T[0,0,3]-t0)/1e3:.1f} q3 {(T[0,0,4]-t0)/1e3:.1f} end {(T[0,0,5]-t0)/1e3:.1f}")
    print(f"  last tile: ticket {(T[-1,-1,0]-t0)/1e3:.1f} start {(T[-1,-1,1]-t0)/1e3:.1f} end {(T[-1,-1,5]-t0)/1e3:.1f}")
