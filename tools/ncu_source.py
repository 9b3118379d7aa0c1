"""Summarise `ncu --page source --csv` of a sweep kernel (development aid): executed warp instructions and stall samples by
SASS address range, plus the hottest instructions.  usage: ncu_source.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
col = {n: i for i, n in enumerate(H)}
data = [r for r in rows[hdr + 1:] if len(r) == len(H)]
stall_cols = [n for n in H if n.startswith("stall_") and "Not Issued" not in n]
tot_inst = sum(float(r[col["Instructions Executed"]] or 0) for r in data)
tot_samp = sum(float(r[col["# Samples"]] or 0) for r in data)
print(f"instructions {len(data)}, executed warp instructions {tot_inst:.4g}, samples {tot_samp:.0f}")
tot = {n: sum(float(r[col[n]] or 0) for r in data) for n in stall_cols}
print("stall reasons, share of samples:", {n[6:]: round(v / tot_samp, 3) for n, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v / tot_samp > 0.01})
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("hottest instructions (samples share, executed share, main stall):")
for r in sorted(data, key=lambda r: -float(r[col["# Samples"]] or 0))[:top]:
    st = max(stall_cols, key=lambda n: float(r[col[n]] or 0))
    print(f"  {r[col['Address']][-5:]} {float(r[col['# Samples']]) / tot_samp:6.3f} {float(r[col['Instructions Executed']]) / tot_inst:6.3f} {st[6:]:14s} {r[col['Source']][:90]}")
# cumulative profile in 64 equal address bins
n = len(data)
print("address profile (bin: share of executed instructions / share of samples):")
B = 32
for b in range(B):
    seg = data[b * n // B:(b + 1) * n // B]
    e = sum(float(r[col["Instructions Executed"]] or 0) for r in seg) / tot_inst
    s = sum(float(r[col["# Samples"]] or 0) for r in seg) / tot_samp
    print(f"  {seg[0][col['Address']][-5:]}..{seg[-1][col['Address']][-5:]}  exec {e:6.3f}  samples {s:6.3f}")
