#!/usr/bin/env python
"""Per-role view of a warp-specialised kernel from ncu's source page:  ncu -i REP --page source --csv > src.csv;  tools/ncu_roles.py src.csv [steps]
The SASS is cut at EXIT instructions (every role of k_backward_cols leaves through its own); for each segment: instructions per step in
its loop, stall samples by reason, hottest lines."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 999
h = next(k for k, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
data = [r for r in rows[h + 1:] if len(r) >= len(hdr) and r[0] != "Address"]
half = len(data) // 2
if half and [r[1] for r in data[:half]] == [r[1] for r in data[half:2 * half]]:
    data = data[:half]
col = {n: i for i, n in enumerate(hdr)}
num = lambda s: int(s) if s.strip().lstrip("-").isdigit() else 0
reasons = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
segs, cur = [], []
for k, r in enumerate(data):
    cur.append(k)
    if "EXIT" in r[col["Source"]]:
        segs.append(cur); cur = []
if cur:
    segs.append(cur)
tot = sum(num(r[col["Warp Stall Sampling (All Samples)"]]) for r in data)
print("total samples", tot)
for si, seg in enumerate(segs):
    smp = sum(num(data[k][col["Warp Stall Sampling (All Samples)"]]) for k in seg)
    ex = sum(num(data[k][col["Instructions Executed"]]) for k in seg)
    if smp < 0.01 * tot and ex < steps * 10:
        continue
    by = {n: sum(num(data[k][col[n]]) for k in seg) for n in reasons}
    by = {n[6:]: v for n, v in sorted(by.items(), key=lambda kv: -kv[1]) if v > 0}
    print("\nsegment %d: SASS lines %d..%d, %.1f warp-instructions per step, %d samples (%.1f%%)" % (si, seg[0], seg[-1], ex / steps, smp, 100.0 * smp / max(tot, 1)))
    print("   stalls:", " ".join("%s=%d" % kv for kv in list(by.items())[:8]))
    top = sorted(seg, key=lambda k: -num(data[k][col["Warp Stall Sampling (All Samples)"]]))[:8]
    for k in sorted(top):
        r = data[k]
        s = num(r[col["Warp Stall Sampling (All Samples)"]])
        if s:
            why = max(reasons, key=lambda n: num(r[col[n]]))[6:]
            print("   %5d %5.1f%% x%-9s %-60s %s" % (k, 100.0 * s / max(tot, 1), r[col["Instructions Executed"]], r[col["Source"]].strip()[:60], why))
