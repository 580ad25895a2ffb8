#!/usr/bin/env python
"""Per-launch counters of the sweep kernels from `ncu --set full --import-source on` captures of the CURRENT build, for bench.py's roofline
block:   python profiles/kernel_counters.py REPORT.ncu-rep [...] > profiles/r02_kernel_counters.json

For every profiled launch of a sweep kernel: duration, DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), executed FP64-pipe warp
instructions (DFMA + DMUL + DADD + DSETP + 64-bit F2F, summed from the source page) and the number of warp-steps they belong to
(TMA sweeps: warps of the grid x (TT-1); candidate kernel: executions of MUFU.RCP64H, one per rollout step).  bench.py scales the
per-lane-and-step figures by the lanes and steps its own run processed.  The file records the hash of csrc/ it was captured from;
bench.py ignores it for any other build."""
import collections
import csv
import hashlib
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TT = 1000
KEYS = [("k_backward_tma", "k_backward_tma"), ("k_forward_cand0_tma", "k_forward_cand0_tma"), ("k_candidates_list", "k_candidates"),
        ("k_rollout_write_tma", "k_rollout_write_tma<.,1>"), ("k_backward_split", "k_backward_split"), ("k_backward_cols", "k_backward_cols"), ("k_search_fused", "k_search_fused")]


def source_sha16():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "aircraftoptimalcontrol_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def raw_rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return rows[0], rows[2:]


def source_mix(rep, idx):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = next(k for k, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    data = [r for r in rows[h + 1:] if len(r) >= len(hdr) and r[0] != "Address"]
    half = len(data) // 2
    if half and [r[1] for r in data[:half]] == [r[1] for r in data[half:2 * half]]:
        data = data[:half]
    isrc, ie = hdr.index("Source"), hdr.index("Instructions Executed")
    cnt = collections.Counter()
    for r in data:
        n = int(r[ie]) if r[ie].isdigit() else 0
        s = r[isrc].strip()
        if s.startswith("@"):
            s = s.split(None, 1)[1]
        op = s.split()[0] if s else "?"
        base = op.split(".")[0]
        if base in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX") or (base == "F2F" and "64" in op):
            cnt["fp64"] += n
        if op.startswith("MUFU.RCP64H"):
            cnt["rcp64"] += n
        cnt["all"] += n
    return cnt


def main(reps):
    out = {"source_sha16": source_sha16(), "instances": 65536, "TT": TT, "how": "ncu --set full --import-source on --clock-control none, bench.py --no-split "
           "(single stream), launches of Newton iteration 16 (float32-noise phase); profiles/kernel_counters.py", "kernels": {}}
    for rep in reps:
        hdr, data = raw_rows(rep)
        col = {n: hdr.index(n) for n in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
                                         "launch__block_size", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                                         "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread")}
        for idx, r in enumerate(data):
            name = r[col["Kernel Name"]]
            key = next((k for pat, k in KEYS if pat in name), None)
            if key is None:
                continue
            if key == "k_rollout_write_tma<.,1>":   # MODE is the 4th template argument; MODE 0 is candidate 0 of the unfused search
                m = re.search(r"k_rollout_write_tma<([^>]*)>", name)
                if not m or not m.group(1).split(",")[3].strip().endswith("1"):
                    continue
            mix = source_mix(rep, idx)
            grid, block = int(float(r[col["launch__grid_size"]])), int(float(r[col["launch__block_size"]]))
            if key == "k_candidates":
                wsteps = mix["rcp64"]
            else:
                wsteps = grid * (block // 32) * (TT - 1)
            dram = (float(r[col["dram__bytes_read.sum"]]) + float(r[col["dram__bytes_write.sum"]])) * 1e9   # (the raw page prints Gbyte)
            e = {"kernel_name": name[:160], "duration_ms": float(r[col["gpu__time_duration.sum"]]), "dram_bytes": dram, "grid": grid, "block": block,
                 "registers": int(float(r[col["launch__registers_per_thread"]])), "fp64_warp_instructions": mix["fp64"], "warp_instructions": mix["all"],
                 "warp_steps": wsteps, "fp64_inst_per_lane_step": mix["fp64"] / max(wsteps, 1), "inst_per_lane_step": mix["all"] / max(wsteps, 1),
                 "dram_bytes_per_lane_step": dram / max(wsteps * 32, 1),
                 "fp64_pipe_pct": float(r[col["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]]),
                 "issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]])}
            if key not in out["kernels"] or e["duration_ms"] > out["kernels"][key]["duration_ms"]:
                out["kernels"][key] = e
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1:])
