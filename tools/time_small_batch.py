#!/usr/bin/env python
"""Diagnostic: Newton iteration time of small batches (single trajectories, late survivor generations) with and without the fused
line search.  python tools/time_small_batch.py [N ...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import aircraftoptimalcontrol_b200 as pkg  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [1, 256, 2048, 4096]
out = {}
for n in sizes:
    xr, ur, dx0, (Q, R, QT), _ = bench.make_problem("step", n, (0, 1))
    for fused in (True, False):
        with pkg.BatchedNewton(n, TT=bench.TT, armijo="lazy", fused=fused, generations=False) as bn:
            bn.set_weights(Q, R, QT)
            bn.set_refs(xr, ur)
            bn.init_guess(dx0=dx0)
            bn.iterate(2)
            bn.iterate(6)
            clean = bn.timing()["total_ms"] / 6    # Gauss-Newton iterations 2..7: candidate 0 accepted everywhere
            bn.iterate(6)
            bn.iterate(6)
            noise = bn.timing()["total_ms"] / 6    # iterations 14..19: float32-noise phase, searches run to exhaustion
            bn.init_guess(dx0=dx0)
            tot = bn.solve()
            whole = bn.timing()["total_ms"]
        out["n%d_%s" % (n, "fused" if fused else "separate")] = dict(clean_ms_per_iter=round(clean, 3), noise_ms_per_iter=round(noise, 3),
                                                                    whole_solve_ms=round(whole, 2), iters=int(tot))
print(json.dumps(out))
