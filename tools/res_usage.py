#!/usr/bin/env python
"""Per-kernel resources of the built libacoc.so (registers, static shared memory, local-memory stack = spills) from
`cuobjdump -res-usage`, and a census of the SASS mnemonics that identify the Blackwell data path (`cuobjdump -sass`): UBLKCP = bulk
asynchronous copy (TMA), SYNCS = mbarrier operations, LDGSTS = cp.async, DFMA / DMUL / DADD = the FP64 pipe, HMMA / IMMA / UTCMMA =
tensor cores (expected: none).  Runs without a GPU:  python tools/res_usage.py > profiles/r02_final_build_resources.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (source_sha16)

so = os.path.join(ROOT, "aircraftoptimalcontrol_b200", "libacoc.so")
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function (\S+?):", res)), capture_output=True, text=True).stdout.split("\n")
rows = []
for (mangled, usage), name in zip(re.findall(r"Function (\S+?):\s*\n\s*(.*)", res), names):
    f = dict(kv.split(":") for kv in usage.split())
    short = re.sub(r"\(.*", "", name.replace("acoc::", "").replace("(bool)", "").replace("(int)", "")).replace("void ", "")
    rows.append((short, int(f.get("REG", 0)), int(f.get("SHARED", 0)), int(f.get("STACK", 0)), int(f.get("LOCAL", 0))))
print("libacoc.so built from csrc/ hash %s (sm_100a only); %d kernel instantiations" % (bench.source_sha16(), len(rows)))
print("\nkernel template: instantiations, registers (min..max), static shared bytes (max; 1024 = the reserved kilobyte), stack frame bytes (max; spills and the out-of-line\nslow path of sincos for huge arguments)")
by = collections.defaultdict(list)
for r in rows:
    by[re.sub(r"<.*", "", r[0])].append(r)
for k in sorted(by):
    v = by[k]
    print("  %-28s %3d  regs %3d..%-3d  smem %6d  stack %4d" % (k, len(v), min(r[1] for r in v), max(r[1] for r in v), max(r[2] for r in v), max(r[3] for r in v)))
print("\nhot instantiations of the default configuration (float64 arithmetic, float state slots, diagonal weights, float32 quantisation):")
hot = ("k_backward_tma<1, double, float, 1>", "k_backward_tma<0, double, float, 1>", "k_forward_cand0_tma<1, double, float, 1>",
       "k_rollout_write_tma<1, double, float, 1, 1>", "k_candidates_list<1, double, 9, 1, 2>", "k_backward_cols<1, double, float, 1>",
       "k_backward_cols<0, double, float, 1>", "k_backward_split<1, double, float, 1>", "k_search_fused<1, double, float>", "k_gradient_tma<double, float>")
for r in rows:
    if r[0] in hot:
        print("  %-48s regs %3d  smem %6d  stack %4d" % r[:4])
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
ops = collections.Counter(m.group(1).split(".")[0] for m in re.finditer(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", sass, re.M))
print("\nSASS mnemonic census of the whole library (static instruction counts):")
for k in ("UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DMUL", "DADD", "F2F", "MUFU", "BAR", "HMMA", "IMMA", "DMMA", "UTCMMA", "UTMALDG", "LDL", "STL"):
    print("  %-8s %7d" % (k, ops.get(k, 0)))
print("  total    %7d" % sum(ops.values()))
