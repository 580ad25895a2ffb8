#!/usr/bin/env python
"""CPU baselines (i) and (ii) of SURVEY.md 8(d): the UNMODIFIED Python reference timed on this machine's cores.

    python oracle/time_python_reference.py [--iters 3] [--out profiles/r02_python_reference_cpu.json]

TEST / BENCH INFRASTRUCTURE ONLY (imports the live reference through oracle/pyref.py; needs /root/reference or
ACOC_REFERENCE_ROOT).  For every batched configuration of BASELINE.json it runs `--iters` Newton iterations
(NewtonMethod.optimize, optcon.py:341-529) of instance 0 of the batch
  (i)  in one process on one core, and
  (ii) in one process per core on independent instances (instances 0..cores-1),
and records trajectory-Newton-iterations per second.  The reference cannot travel to the GPU box, so bench.py reports the
numbers recorded here (labelled "recorded", with the machine they were taken on) when the reference tree is absent, and
re-measures (i) live when it is present.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def instance_problem(workload, i, n_total=None):
    """References, weights and initial guess of instance i of the batched configuration (same generators / seeds as bench.py)."""
    from aircraftoptimalcontrol_b200 import refgen
    from oracle import corcl
    if workload == "step":
        zf, xf = refgen.config4_params(n_total or 65536, 2024)
        xr, ur = refgen.step_problem(float(xf[i]), float(zf[i]))
        Q, R, QT = refgen.weights("step")
        xi, ui = corcl.initial_trajectory(xr)
    else:
        dx0, zf = refgen.config5_params(n_total or 1048576, 7)
        xr, ur = refgen.acrobatic_problem(float(zf[i]))
        Q, R, QT = refgen.weights("acro")
        x = xr.copy()
        x[:, 0] += dx0[i]
        xi, ui = corcl.initial_trajectory(x)
    return xr, ur, xi, ui, Q, R, QT


def time_one(args):
    workload, i, iters = args
    from oracle import pyref
    mods = pyref.load()
    xr, ur, xi, ui, Q, R, QT = instance_problem(workload, i)
    t0 = time.perf_counter()
    h = pyref.run_newton(mods, xr, ur, xi, ui, Q, R, QT, max_iters=iters + 1)   # range(max_iters - 1) loop bodies, optcon.py:415
    dt = time.perf_counter() - t0
    return int(h["iters"]), dt


def measure(workload, iters, cores):
    done, dt = time_one((workload, 0, iters))
    one = {"value": done / dt, "unit": "traj-Newton-it/s", "cores": 1, "iterations": done, "wall_s": dt}
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(time_one, [(workload, i, iters) for i in range(cores)])
    wall = time.perf_counter() - t0
    tot = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    allc = {"value": tot / busy, "unit": "traj-Newton-it/s", "cores": cores, "iterations": tot, "wall_s": busy,
            "wall_with_process_start_s": wall, "how": "one process per core, independent instances 0..%d" % (cores - 1)}
    return {"one_core": one, "all_cores": allc}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_python_reference_cpu.json"))
    a = ap.parse_args()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    out = {"what": "unmodified Python reference (NewtonMethod.optimize, optcon.py:341-529), %d Newton iterations per instance" % a.iters,
           "machine": {"cpu": cpu, "cores": cores, "platform": platform.platform(), "numpy": np.__version__, "where": "build container"},
           "workloads": {w: measure(w, a.iters, cores) for w in ("step", "acro")}}
    # whole solves of configs 1 and 2 recorded when the golden fixtures were generated (oracle/gen_golden.py), 1 core
    g = {}
    for name in ("newton_step_f32", "newton_acro_f32"):
        d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        g[name] = {"iterations": int(d["iters"]), "wall_s": float(d["ref_wall_s"]), "value": int(d["iters"]) / float(d["ref_wall_s"]),
                   "unit": "traj-Newton-it/s", "cores": 1}
    out["whole_solves_configs_1_2"] = g
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
