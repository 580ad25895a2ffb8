#!/usr/bin/env python
"""Executed instruction mix (per opcode) of one profiled launch:  tools/ncu_mix.py REPORT.ncu-rep launch-index warp_steps"""
import csv, io, subprocess, sys, collections
rep, skip, wsteps = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(k for k, r in enumerate(rows) if r and r[0] == "Address")
print(rows[0][1][:120])
hdr = rows[h]; data = [r for r in rows[h + 1:] if len(r) >= len(hdr) and r[0] != "Address"]
half = len(data) // 2
if [r[1] for r in data[:half]] == [r[1] for r in data[half:2 * half]]:
    data = data[:half]   # ncu lists the function twice when it has two source views
isrc, iw, ie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
cnt, smp = collections.Counter(), collections.Counter()
for r in data:
    n = int(r[ie]) if r[ie].isdigit() else 0
    s = r[isrc].strip()
    if s.startswith("@"):
        s = s.split(None, 1)[1]
    op = s.split()[0].split(".")[0] if s else "?"
    cnt[op] += n; smp[op] += int(r[iw]) if r[iw].isdigit() else 0
tot, ts = sum(cnt.values()), sum(smp.values())
print("total warp instructions %d = %.1f per warp-step" % (tot, tot / wsteps))
for op, n in cnt.most_common(28):
    print("%-10s %6.2f%% of instr (%6.1f per warp-step)  %5.1f%% of samples" % (op, 100 * n / tot, n / wsteps, 100 * smp[op] / max(ts, 1)))
