#!/usr/bin/env python
"""Diagnostic: GradientMethod.optimize as a batch -- device time per steepest-descent iteration and the costate sweep's share.
python tools/time_gradient.py [N]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import aircraftoptimalcontrol_b200 as pkg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
xr, ur, dx0, (Q, R, QT) = bench.make_problem("step", n, (0, 1))[:4]
out = {"instances": n}
for tma in (True, False):
    with pkg.BatchedNewton(n, TT=bench.TT, armijo="lazy", method="gradient", max_iters=40, tma=tma) as bn:
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess(dx0=dx0)
        bn.iterate(3)
        bn.iterate(8)
        per_iter = bn.timing()["total_ms"] / 8
        bn.set_profiling(True)
        bn.iterate(4)
        ph = bn.timing()["phases"]
        nc = float(bn.history()["n_armijo"][:, 3:15].mean())
    # costate sweep: 128 B read (x, u, refs as float64) + 16 B written per instance-step (algorithmic)
    sweep_ms = ph["backward"] / 4
    out["tma" if tma else "plain"] = dict(ms_per_iteration=round(per_iter, 3), costate_sweep_ms=round(sweep_ms, 3),
                                          costate_sweep_algorithmic_GBs=round(n * (bench.TT - 1) * 144 / sweep_ms / 1e6, 1),
                                          candidates_ms=round(ph["candidates"] / 4, 3), update_ms=round(ph["update"] / 4, 3),
                                          mean_candidates=round(nc, 2))
print(json.dumps(out))
