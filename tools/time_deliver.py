#!/usr/bin/env python
"""Diagnostic: device->host bandwidth of the result read-back: staged download (result_f32) vs direct delivery into page-locked memory."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aircraftoptimalcontrol_b200 as pkg
from aircraftoptimalcontrol_b200 import refgen
n, TT = 65536, 1000
zf, xf = refgen.config4_params(n, 2024)
Q, R, QT = refgen.weights("step")
xs_t = torch.empty((n, 6, TT), dtype=torch.float32, pin_memory=True); us_t = torch.empty((n, 2, TT), dtype=torch.float64, pin_memory=True)
xs, us = xs_t.numpy(), us_t.numpy()
gb = (xs.nbytes + us.nbytes) / 1e9
with pkg.BatchedNewton(n, TT=TT, armijo="lazy") as bn:
    bn.set_weights(Q, R, QT); bn.set_refs_step(zf, xf); bn.init_guess()
    bn.solve()
    for rep in range(3):
        t0 = time.perf_counter(); bn.result_f32(out=(xs, us)); t1 = time.perf_counter()
        print("staged download  %.1f ms  %.1f GB/s" % ((t1 - t0) * 1e3, gb / (t1 - t0)), flush=True)
    for rep in range(3):
        t0 = time.perf_counter(); bn.solve_deliver((xs, us)); t1 = time.perf_counter()
        print("direct delivery (+ ~1 ms of empty iterations)  %.1f ms  %.1f GB/s" % ((t1 - t0) * 1e3, gb / (t1 - t0)), flush=True)
