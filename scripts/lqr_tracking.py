#!/usr/bin/env python
"""Headless drop-in for the reference's lqr_tracking.py __main__ block (:321-342): LQR tracking of a saved optimum
from perturbed initial states (config 3 with --instances 4096).
    python scripts/lqr_tracking.py --xx Data/xx_star.npy --uu Data/uu_star.npy [--instances 4096] [--out Data]
"""
import argparse

import numpy as np

import _common  # noqa: F401
from aircraftoptimalcontrol_b200 import refgen
from aircraftoptimalcontrol_b200.aircraft_simplified import Dynamics
from aircraftoptimalcontrol_b200.lqr_tracking import lqr_tracking_batch

ap = argparse.ArgumentParser()
ap.add_argument("--xx", default="Data/xx_star.npy")
ap.add_argument("--uu", default="Data/uu_star.npy")
ap.add_argument("--instances", type=int, default=1)
ap.add_argument("--state", default="f32", choices=["f32", "f64"])
ap.add_argument("--out", default="Data")
args = ap.parse_args()

xx_opt, uu_opt = np.load(args.xx), np.load(args.uu)
QQt, RRt, QQT = refgen.weights("track")
deltas = refgen.config3_deltas(args.instances)  # instance 0 = the shipped 0.1*ones(6) (lqr_tracking.py:259)
xx_lqr, uu_lqr = lqr_tracking_batch(xx_opt, uu_opt, deltas, Dynamics(state=args.state), QQt, RRt, QQT)
_common.save("xx_lqr.npy", xx_lqr[0] if args.instances == 1 else xx_lqr, args.out)
_common.save("uu_lqr.npy", uu_lqr[0] if args.instances == 1 else uu_lqr, args.out)
print("tracked %d perturbed initial states; final tracking error of instance 0: %s" % (args.instances, np.abs(xx_lqr[0, :, -1] - xx_opt[:, -1])))
