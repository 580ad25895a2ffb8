#!/usr/bin/env python
"""Diagnostic: per-call latency of the pointwise drop-in entry points (Dynamics.step, Cost.stagecost, lqr_tracking_batch) through the C ABI."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aircraftoptimalcontrol_b200.aircraft_simplified import Cost, Dynamics
from aircraftoptimalcontrol_b200.lqr_tracking import lqr_tracking_batch
from aircraftoptimalcontrol_b200 import refgen

def best(f, reps=200):
    f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts))

dyn = Dynamics()
x, u, lam = np.array([0, 0, 16.0, 0.01, 0, 0.01]), np.array([46.0, 0.1]), np.ones(6)
print("Dynamics.step(xx, uu)            %.1f us per call" % best(lambda: dyn.step(x, u)))
print("Dynamics.step(xx, uu, lmbd)      %.1f us per call" % best(lambda: dyn.step(x, u, lam)))
Q, R, QT = refgen.weights("step")
c = Cost(Q, R, QT)
print("Cost.stagecost                   %.1f us per call" % best(lambda: c.stagecost(x, u, x * 0.9, u * 0.9)))
X, U = np.tile(x, (10000, 1)), np.tile(u, (10000, 1))
print("Dynamics.step_batch (10^4)       %.1f us per call" % best(lambda: dyn.step_batch(X, U), 50))
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "lqr_tracking.npz"))
delta = refgen.config3_deltas(4096)
print("lqr_tracking_batch (4096)        %.1f ms per call" % (best(lambda: lqr_tracking_batch(d["xx_opt"], d["uu_opt"], delta), 5) / 1e3))
