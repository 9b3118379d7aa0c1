"""In-CTA lag growth from a TTCR_B200_TRACE_STEPS dump: start of step a of warp w minus start of step a of warp w-1 (cycles)."""
import sys
import numpy as np
raw = np.fromfile(sys.argv[1], dtype=np.int64)
t = raw[:8192].reshape(16, 128, 4)[:8].astype(np.float64)
print("step:   " + "  ".join(f"{a:6d}" for a in (0, 4, 8, 16, 32, 48, 64, 96, 127)))
for w in range(1, 8):
    if t[w][:, 0].max() == 0:
        continue
    print(f"w{w}-w{w-1}: " + "  ".join(f"{t[w][a,0]-t[w-1][a,0]:6.0f}" for a in (0, 4, 8, 16, 32, 48, 64, 96, 127)))
print("w7-w0:  " + "  ".join(f"{t[7][a,0]-t[0][a,0]:6.0f}" for a in (0, 4, 8, 16, 32, 48, 64, 96, 127)))
print("warp 0 step time (start to start): " + " ".join(f"{v:.0f}" for v in np.diff(t[0][:40, 0])))
print("warp 0 waits (stamp1-stamp0):      " + " ".join(f"{v:.0f}" for v in (t[0][:40, 1] - t[0][:40, 0])))
print("warp 3 step time:                  " + " ".join(f"{v:.0f}" for v in np.diff(t[3][:40, 0])))
print("warp 3 waits:                      " + " ".join(f"{v:.0f}" for v in (t[3][:40, 1] - t[3][:40, 0])))
