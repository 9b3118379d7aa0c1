"""Cross-tile hand-off latency from two TTCR_B200_TRACE_STEPS dumps taken with globaltimer stamps:
   file A = producer tile (stamp 3 of its last warp = row published), file B = consumer tile (stamp 1 of warp 0 = row seen)."""
import sys
import numpy as np
def load(f):
    return np.fromfile(f, dtype=np.int64)[:8192].reshape(16, 128, 4)[:8].astype(np.float64)
A, B = load(sys.argv[1]), load(sys.argv[2])
nw = int(sys.argv[3])
pub = A[nw - 1][:, 3]      # producer's last warp: end of step a (row a published)
seen = B[0][:, 1]          # consumer warp 0: waits of step a done
start = B[0][:, 0]
print("row published -> consumer warp 0 has it (ns): mean %.0f  p10 %.0f  p90 %.0f" % (np.mean(seen - pub), *np.percentile(seen - pub, [10, 90])))
print("consumer warp 0 arrives at step a (start) relative to publication (ns): mean %.0f" % np.mean(start - pub))
print("producer last warp period (ns): %.0f   consumer warp 0 period: %.0f" % (np.mean(np.diff(pub)), np.mean(np.diff(seen))))
for w in range(1, nw):
    print("within producer CTA: warp %d end of step a - warp %d end of step a (ns): %.0f" % (w, w - 1, np.mean(A[w][:, 3] - A[w - 1][:, 3])))
