"""Start / end / run time of every tile of one sweep from a TTCR_B200_TRACE dump of k_sweep_march (development aid)."""
import sys

import numpy as np

raw = open(sys.argv[1], "rb").read()
want = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pos, k = 0, 0
while pos < len(raw):
    ntiles, nU, nV, PUT = np.frombuffer(raw, dtype=np.int32, count=4, offset=pos)
    pos += 16
    RL, PUT = max(8, PUT // 1000), PUT % 1000
    t = np.frombuffer(raw, dtype=np.int64, count=ntiles * RL, offset=pos).reshape(ntiles, RL).astype(np.float64)
    pos += ntiles * RL * 8
    k += 1
    if k != want:
        continue
    T = (t.reshape(nU, nV, RL) - t[:, 0].min()) / 1e3
    np.set_printoptions(linewidth=250, precision=0, suppress=True)
    print("start (us), rows U, columns V"); print(T[:, :, 1].round(0))
    print("end"); print(T[:, :, 5].round(0))
    print("run"); print((T[:, :, 5] - T[:, :, 1]).round(0))
    sm = t.reshape(nU, nV, RL)[:, :, 7].astype(int)
    busy = {}
    for U in range(nU):
        for V in range(nV):
            busy.setdefault(sm[U, V], []).append((T[U, V, 0], T[U, V, 5], U, V))
    gaps = []
    for s, lst in busy.items():
        lst.sort()
        gaps.append(sum(b[0] - a[1] for a, b in zip(lst[:-1], lst[1:])))
    print("tiles per SM: min", min(len(v) for v in busy.values()), "max", max(len(v) for v in busy.values()), "; idle between tiles per SM, mean us:", np.mean(gaps).round(1))
    print("SM 0 sequence:", [(int(a), int(b), U, V) for a, b, U, V in sorted(busy[sm[0, 0]])])
    if RL >= 16:
        R = t.reshape(nU, nV, RL)
        print("warp 0: us waiting for ring words (at 1.9 GHz)"); print((R[:, :, 2] / 1900).round(0))
        print("warp 0: us waiting for chunks"); print((R[:, :, 3] / 1900).round(0))
        print("warp 0: ring waits (count)"); print(R[:, :, 4].round(0))
        print("last warp: us waiting for ring words"); print((R[:, :, 6] / 1900).round(0))
        print("importer: loop iterations"); print(R[:, :, 8].round(0))
        print("importer: deliveries by lane 0"); print(R[:, :, 9].round(0))
        print("importer: us"); print((R[:, :, 10] / 1900).round(0))
        print("V importer: rounds"); print(R[:, :, 11].round(0))
