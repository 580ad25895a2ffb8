#!/usr/bin/env python
"""Headless drop-in for the reference's acrobatic_newton.py (bump maneuver, config 2) on the GPU.

Constants of acrobatic_newton.py:34-44, :55-65, :72-76; calls of :144-196; outputs
Data/xx_star_acrobatic.npy, Data/uu_star_acrobatic.npy (:201-203).
    python scripts/acrobatic_newton.py [--state f32|f64] [--out Data]
"""
import argparse

import numpy as np

import _common  # noqa: F401
from aircraftoptimalcontrol_b200 import refgen
from aircraftoptimalcontrol_b200.aircraft_simplified import Cost, Dynamics
from aircraftoptimalcontrol_b200.optcon import NewtonMethod

ap = argparse.ArgumentParser()
ap.add_argument("--state", default="f32", choices=["f32", "f64"])
ap.add_argument("--out", default="Data")
ap.add_argument("--ref-init", default=None)
args = ap.parse_args()

max_iters, stepsize_0, cc, beta, armijo_maxiters, term_cond = int(2e2), 1, 0.5, 0.7, 10, 1e-6
dyn = Dynamics(state=args.state)
QQt, RRt, QQT = refgen.weights("acro")
tf, dt = 1, 1e-3
dyn.dt = dt
TT = int(tf / dt)
tt = np.linspace(0, tf, TT)
xx_ref, uu_ref = refgen.acrobatic_problem(zf=2.71, xf=18, tf=tf, TT=TT)
cst = Cost(QQt, RRt, QQT)
NM = NewtonMethod(dyn, cst, xx_ref, uu_ref, max_iters=max_iters, stepsize_0=stepsize_0, cc=cc, beta=beta,
                  armijo_maxiters=armijo_maxiters, term_cond=term_cond)
if args.ref_init:
    z = np.load(args.ref_init)
    xx_init, uu_init = z["xx_init"], z["uu_init"]
else:
    xx_init, uu_init = dyn.get_initial_trajectory(xx_ref, tt)
xx_star, uu_star = NM.optimize(xx_init, uu_init, tf, dt)
_common.save("xx_star_acrobatic.npy", xx_star, args.out)
_common.save("uu_star_acrobatic.npy", uu_star, args.out)
print("Newton iterations: %d, final cost %.6f" % (NM.history["iters"], NM.history["JJ"][-1]))
