"""Per-step clocks of tile 0 from a step-trace build of k_sweep_march (TTCR_B200_TRACE_STEPS dump; development aid).
Stamps per step: 0 top, 1 inputs there (after the wait branch), 2 update done, 3 end."""
import sys

import numpy as np

raw = np.fromfile(sys.argv[1], dtype=np.int64)
rec = raw.reshape(-1, 32, 64, 4)
k = int(sys.argv[2]) if len(sys.argv) > 2 else rec.shape[0] - 1
r = rec[k].astype(np.float64)
for w in range(32):
    t = r[w]
    if not t[:, 0].any():
        continue
    ok = (t[:, 0] > 0) & (t[:, 3] > 0)
    t = t[ok]
    per = np.diff(t[:, 0])
    print(f"warp {w:2d}: step period mean {per.mean():7.1f} (p10 {np.percentile(per, 10):.0f} p90 {np.percentile(per, 90):.0f}) | wait {np.mean(t[:, 1] - t[:, 0]):6.1f}"
          f" update {np.mean(t[:, 2] - t[:, 1]):6.1f} tail {np.mean(t[:, 3] - t[:, 2]):6.1f} next-top {np.mean(t[1:, 0] - t[:-1, 3]):6.1f}")
