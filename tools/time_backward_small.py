#!/usr/bin/env python
"""Diagnostic: per-phase device time of Newton iterations of small batches (profiling events on): backward sweep vs line search.
python tools/time_backward_small.py [N ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import aircraftoptimalcontrol_b200 as pkg  # noqa: E402
from aircraftoptimalcontrol_b200 import _lib as L  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [1, 4096]
out = {}
for n in sizes:
    xr, ur, dx0, (Q, R, QT), _ = bench.make_problem("step", n, (0, 1))
    with pkg.BatchedNewton(n, TT=bench.TT, armijo="lazy", generations=False) as bn:
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess(dx0=dx0)
        bn.iterate(2)
        L.check(L.lib().acoc_set_profiling(bn._h, 1))
        res = {}
        for name, k in (("gauss_newton_2_7", 6), ("exact_8_13", 6), ("noise_14_19", 6)):
            bn.iterate(k)
            t = bn.timing()
            res[name] = dict(total=round(t["total_ms"] / k, 3), **{p: round(v / k, 3) for p, v in t["phases"].items() if v > 0})
        out["n%d" % n] = res
print(json.dumps(out))
