#!/usr/bin/env python
"""Diagnostic: per-iteration device time and active-instance count of a full batched solve, plus the wall-clock
breakdown of the end-to-end path (H2D / init / solve / D2H).  python tools/solve_trace.py [--instances N] [--workload step|acro]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import aircraftoptimalcontrol_b200 as pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--instances", type=int, default=65536)
ap.add_argument("--workload", default="step")
ap.add_argument("--armijo", default="lazy")
ap.add_argument("--state", default="f32")
a = ap.parse_args()
n = a.instances
xr, ur, dx0, (Q, R, QT), _ = bench.make_problem(a.workload, n, (0, 1))
xr_p, ur_p = bench.pinned_like(xr), bench.pinned_like(ur)
import torch
xs_t = torch.empty((n, 6, bench.TT), dtype=torch.float64, pin_memory=True)
us_t = torch.empty((n, 2, bench.TT), dtype=torch.float64, pin_memory=True)
bn = pkg.BatchedNewton(n, TT=bench.TT, state=a.state, armijo=a.armijo)
bn.set_weights(Q, R, QT)
out = {}
for rep in range(2):
    t0 = time.perf_counter(); bn.set_refs(xr_p.numpy(), ur_p.numpy()); t1 = time.perf_counter()
    bn.init_guess(dx0=dx0); t2 = time.perf_counter()
    tot = bn.solve(); t3 = time.perf_counter()
    bn.result(out=(xs_t.numpy(), us_t.numpy())); t4 = time.perf_counter()
    st = bn.stats(); t5 = time.perf_counter()
    out["e2e_rep%d" % rep] = dict(h2d_s=t1 - t0, init_s=t2 - t1, solve_s=t3 - t2, d2h_s=t4 - t3, stats_s=t5 - t4, total_s=t5 - t0, iters=int(tot),
                                  solve_device_ms=bn.timing()["total_ms"])
# per-iteration trace
bn.init_guess(dx0=dx0)
bn.set_profiling(True)
trace, phases = [], []
active = n
while active > 0 and len(trace) < 199:
    active = bn.iterate(1)
    tm = bn.timing()
    trace.append((round(tm["total_ms"], 3), active))
    phases.append([round(tm["phases"][k], 2) for k in ("backward", "forward", "candidates", "update")])
out["trace_ms_active"] = trace
out["phases_bwd_fwd_cand_upd"] = phases
out["n_armijo_mean_per_iter"] = [round(float(x), 2) for x in bn.history()["n_armijo"][:, :len(trace)].mean(axis=0)]
st = bn.stats()
out["iters_hist"] = np.bincount(st["iters"]).tolist()
print(json.dumps(out))
