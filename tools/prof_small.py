#!/usr/bin/env python
"""Small workloads for ncu / compute-sanitizer: a 4096-instance Newton batch through the fused line search and a 65,536-instance
steepest-descent iteration.  python tools/prof_small.py newton|gradient|sanitize"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import aircraftoptimalcontrol_b200 as pkg  # noqa: E402

what = sys.argv[1]
if what == "sanitize":   # tiny: every kernel family once, short horizon
    import numpy as np
    from aircraftoptimalcontrol_b200 import refgen
    n, TT = 70, 40
    rng = np.random.default_rng(0)
    xr, ur = refgen.step_problem(rng.uniform(14, 18, n), rng.uniform(1.5, 3.5, n), tf=TT * 1e-3, TT=TT)
    Q, R, QT = refgen.weights("step")
    for kw in (dict(armijo="lazy"), dict(armijo="lazy", fused=False), dict(armijo="speculative", method="gradient"), dict(armijo="lazy", tma=False, fused=False)):
        with pkg.BatchedNewton(n, TT=TT, max_iters=14, **kw) as bn:
            bn.set_weights(Q, R, QT)
            bn.set_refs(xr, ur)
            bn.init_guess()
            bn.solve()
            bn.result()
    print("sanitize workload done")
else:
    n = 4096 if what == "newton" else 65536
    xr, ur, dx0, (Q, R, QT) = bench.make_problem("step", n, (0, 1))[:4]
    with pkg.BatchedNewton(n, TT=bench.TT, armijo="lazy", method=what, generations=False) as bn:
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess(dx0=dx0)
        bn.iterate(3)
    print("done")
