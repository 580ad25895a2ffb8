#!/usr/bin/env python
"""Static count of the FP64-pipe instructions inside the time loop of every sweep kernel of libacoc.so.

    python tools/sass_fp64_count.py [path/to/libacoc.so] > profiles/r02_sass_fp64_counts.json

For each kernel the time loop is taken to be the LARGEST region closed by a backward branch (the sweeps have one hot loop over the
horizon; everything per time step sits inside it, the out-of-line slow paths of division / sincos sit behind it and are not counted).
Counted per loop body: DFMA, DMUL, DADD, DSETP/DMNMX, F2F with a 64-bit side, MUFU.*64H -- the instructions that occupy the FP64
pipe (one issue slot each).  bench.py multiplies `fp64_pipe` by the lanes a launch processes to report the EXECUTED FP64 instruction
rate against the measured DFMA peak (a static count: both arms of data-dependent branches inside the loop are included, which
overstates the executed number by a few instructions, e.g. the +0.5*I gain branch of the backward sweep).
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "aircraftoptimalcontrol_b200", "libacoc.so")
KEEP = re.compile(r"k_(backward|forward|rollout_write|candidates|gradient|search_fused|update|candidate0|track|init_guess|traj_cost|lq_dense)")


def source_hash():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "aircraftoptimalcontrol_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = {}
    out = {}
    cur, ins = None, []

    def flush():
        if cur is None or not ins:
            return
        addr = [a for a, _, _ in ins]
        best = None
        for a, op, rest in ins:
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", rest)
                if m:
                    tgt = int(m.group(1), 16)
                    if tgt < a and (best is None or a - tgt > best[1] - best[0]):
                        best = (tgt, a)
        if best is None:
            return
        body = [(op, rest) for a, op, rest in ins if best[0] <= a <= best[1]]
        c = dict(dfma=0, dmul=0, dadd=0, dcmp=0, f2f64=0, mufu64=0, ffma=0, fp32_other=0, lds=0, ldg=0, stg=0, total=len(body))
        for op, rest in body:
            base = op.split(".")[0]
            if base == "DFMA": c["dfma"] += 1
            elif base == "DMUL": c["dmul"] += 1
            elif base == "DADD": c["dadd"] += 1
            elif base in ("DSETP", "DMNMX"): c["dcmp"] += 1
            elif base == "F2F" and "64" in op: c["f2f64"] += 1
            elif base == "MUFU" and "64" in op: c["mufu64"] += 1
            elif base == "FFMA": c["ffma"] += 1
            elif base in ("FMUL", "FADD"): c["fp32_other"] += 1
            elif base in ("LDS", "LDSM"): c["lds"] += 1
            elif base in ("LDG", "LD"): c["ldg"] += 1
            elif base in ("STG", "ST"): c["stg"] += 1
        c["fp64_pipe"] = c["dfma"] + c["dmul"] + c["dadd"] + c["dcmp"] + c["f2f64"]
        out[cur] = c

    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            flush()
            cur, ins = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);", line)
        if m and cur:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
    flush()
    dem = subprocess.run(["c++filt"], input="\n".join(out.keys()), capture_output=True, text=True).stdout.splitlines()
    res = {}
    for mangled, d in zip(out.keys(), dem):
        if not KEEP.search(d):
            continue
        short = re.sub(r"\(.*", "", d).replace("void ", "").replace("acoc::", "")
        res[short] = out[mangled]
    json.dump({"source_sha16": source_hash(), "how": "cuobjdump -sass, largest backward-branch region per kernel (tools/sass_fp64_count.py)",
               "kernels": res}, sys.stdout, indent=1, sort_keys=True)
    print()


if __name__ == "__main__":
    main()
