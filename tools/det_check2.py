#!/usr/bin/env python
"""Diagnostic: iterate a ragged batch one Newton iteration at a time with and without the tile-range split and compare the gains
(backward sweep output), deltau (forward pass output) and the new iterate after every iteration."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aircraftoptimalcontrol_b200 as pkg
from aircraftoptimalcontrol_b200 import refgen
n, TT = 40001, 48
rng = np.random.default_rng(23)
zf, xf = rng.uniform(1.5, 3.5, n), rng.uniform(14, 18, n)
xr, ur = refgen.step_problem(xf, zf, tf=TT * 1e-3, TT=TT)
Q, R, QT = refgen.weights("step")
ctx = []
for split in (False, True):
    bn = pkg.BatchedNewton(n, TT=TT, armijo="lazy", split=split)
    bn.set_weights(Q, R, QT); bn.set_refs(xr, ur); bn.init_guess()
    ctx.append(bn)
for kk in range(12):
    outs = []
    for bn in ctx:
        bn.iterate(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
        K, s = bn.gains()
        outs.append((K, s, bn.deltau(), bn.iterate_at(0), bn.stats()["descent"].copy(), bn.timing()["launches"]))
    a, b = outs
    dK = np.nonzero(np.any(a[0] != b[0], axis=(1, 2, 3)))[0]
    ds = np.nonzero(np.any(a[1] != b[1], axis=(1, 2)))[0]
    dd = np.nonzero(np.any(a[2] != b[2], axis=(1, 2)))[0]
    dx = np.nonzero(np.any(a[3][0] != b[3][0], axis=(1, 2)))[0]
    dde = np.nonzero(a[4] != b[4])[0]
    print("after call %d (launches %d / %d): instances with different K %d, sigma %d, deltau %d, x %d, descent %d; tiles K %s deltau %s" % (
        kk, a[5], b[5], len(dK), len(ds), len(dd), len(dx), len(dde), sorted(set((dK // 32).tolist()))[:6], sorted(set((dd // 32).tolist()))[:6]), flush=True)
    if len(dK):
        i = dK[0]
        tdiff = np.nonzero(np.any(a[0][i] != b[0][i], axis=(0, 1)))[0]
        print("   instance %d: K differs at time steps %s" % (i, tdiff[:12]), flush=True)
    if len(ds) and not len(dK):
        for i in ds[:3]:
            tdiff = np.nonzero(np.any(a[1][i] != b[1][i], axis=0))[0]
            print("   instance %d: sigma differs at %d time steps, first %d last %d (K identical); sigma[last] %s vs %s" % (
                i, len(tdiff), tdiff[0], tdiff[-1], a[1][i][:, tdiff[-1]], b[1][i][:, tdiff[-1]]), flush=True)
