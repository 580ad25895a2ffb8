#!/usr/bin/env python
"""Diagnostic: run the same ragged batch several times with and without the tile-range split and report whether totals / per-instance
iteration counts are reproducible.  python tools/det_check.py [n] [TT] [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aircraftoptimalcontrol_b200 as pkg
from aircraftoptimalcontrol_b200 import refgen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40001
TT = int(sys.argv[2]) if len(sys.argv) > 2 else 48
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rng = np.random.default_rng(23)
zf, xf = rng.uniform(1.5, 3.5, n), rng.uniform(14, 18, n)
xr, ur = refgen.step_problem(xf, zf, tf=TT * 1e-3, TT=TT)
Q, R, QT = refgen.weights("step")
ref = None
for split in (False, True):
    for rep in range(reps):
        with pkg.BatchedNewton(n, TT=TT, armijo="lazy", split=split) as bn:
            bn.set_weights(Q, R, QT); bn.set_refs(xr, ur); bn.init_guess()
            bn.iterate(3)
            tot = bn.solve()
            it = bn.stats()["iters"].copy()
            h = bn.history()
        if ref is None:
            ref = (it, h)
        dif = np.zeros(n, dtype=bool)
        for key in ("JJ", "descent", "stepsize", "n_armijo"):
            dif |= np.any(h[key] != ref[1][key], axis=1)
        d = np.nonzero(dif)[0]
        print("split=%s rep %d total %d instances with a different history: %d" % (split, rep, tot, len(d)), flush=True)
        for i in d[:4]:
            ks = [int(np.nonzero(h[key][i] != ref[1][key][i])[0][0]) if np.any(h[key][i] != ref[1][key][i]) else 999 for key in ("JJ", "descent", "stepsize", "n_armijo")]
            k = min(ks)
            print("   instance %d (tile %d lane %d): first difference at iteration %d in %s: JJ %.17g/%.17g descent %.17g/%.17g step %g/%g ncand %d/%d" % (
                i, i // 32, i % 32, k, [key for key, kk in zip(("JJ", "descent", "stepsize", "n_armijo"), ks) if kk == k],
                h["JJ"][i, k], ref[1]["JJ"][i, k], h["descent"][i, k], ref[1]["descent"][i, k], h["stepsize"][i, k], ref[1]["stepsize"][i, k],
                h["n_armijo"][i, k], ref[1]["n_armijo"][i, k]), flush=True)
