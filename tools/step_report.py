"""Per-step clocks of one tile of k_sweep_patch (TTCR_B200_TRACE_STEPS dump): [warp][step 200..327][start, waits done, math start, end]."""
import sys
import numpy as np
raw = np.fromfile(sys.argv[1], dtype=np.int64)
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t = raw[k * 8192:(k + 1) * 8192].reshape(16, 128, 4)[:8].astype(np.float64)
for w in range(8):
    x = t[w]
    if x[:, 0].max() == 0:
        continue
    st = np.diff(x[:, 0])
    print(f"warp {w}: step period mean {st.mean():.0f} cyc (even->odd {st[0::2].mean():.0f}, odd->even {st[1::2].mean():.0f}); "
          f"loads+waits {np.mean(x[:,1]-x[:,0]):.0f} (even {np.mean(x[0::2,1]-x[0::2,0]):.0f} odd {np.mean(x[1::2,1]-x[1::2,0]):.0f}), "
          f"shfl/backpressure {np.mean(x[:,2]-x[:,1]):.0f}, math+stores {np.mean(x[:,3]-x[:,2]):.0f}, "
          f"gap to next step {np.mean(x[1:,0]-x[:-1,3]):.0f} (after even {np.mean(x[1::2,0]-x[0:-1:2,3]):.0f}, after odd {np.mean(x[2::2,0]-x[1:-1:2,3]):.0f})")
w0 = t[0]
for w in range(1, 8):
    if t[w][:, 0].max() == 0:
        continue
    print(f"warp {w} start of step a minus warp {w-1} end of step a+1: mean {np.mean(t[w][:-1,0]-t[w-1][1:,3]):.0f} cyc")
print("start of steps 0..7 of the window per warp, relative to warp 0 step 0:")
for w in range(8):
    if t[w][:, 0].max() == 0:
        continue
    print(w, [int(v - w0[0, 0]) for v in t[w][:8, 0]], " end of step 0:", int(t[w][0, 3] - w0[0, 0]), "waits done:", [int(v - w0[0, 0]) for v in t[w][:4, 1]])
