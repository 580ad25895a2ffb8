#!/usr/bin/env python
"""Hottest SASS instructions (warp stall samples) of one profiled launch:  tools/ncu_hot.py REPORT.ncu-rep [launch-index] [top-n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(k for k, r in enumerate(rows) if r and r[0] == "Address")
print(rows[0][1][:120] if rows[0] else "")
hdr = rows[h]; data = [r for r in rows[h + 1:] if len(r) >= len(hdr) and r[0] != "Address"]
isrc, iw, ie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
num = lambda s: int(s) if s.strip().isdigit() else 0
tot, totx = sum(num(r[iw]) for r in data), sum(num(r[ie]) for r in data)
print("total samples", tot, "warp instructions", totx, "SASS lines", len(data))
top = sorted(range(len(data)), key=lambda k: -num(data[k][iw]))[:topn]
for k in sorted(top):
    r = data[k]
    print("%5d %6.2f%% x%11s  %s" % (k, 100.0 * num(r[iw]) / max(tot, 1), r[ie], r[isrc][:120]))
