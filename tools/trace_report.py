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
