#!/usr/bin/env python
"""Diagnostic: wall-clock breakdown of one end-to-end solve of the headline configuration on one context."""
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
with pkg.BatchedNewton(n, TT=TT, armijo="lazy") as bn:
    bn.set_weights(Q, R, QT)
    for rep in range(3):
        direct = rep != 1
        t0 = time.perf_counter(); bn.set_refs_step(zf, xf); t1 = time.perf_counter()
        bn.init_guess(); t2 = time.perf_counter()
        if direct:
            bn.solve_deliver((xs, us)); t3 = time.perf_counter(); t4 = t3
        else:
            bn.solve(); t3 = time.perf_counter(); bn.result_f32(out=(xs, us)); t4 = time.perf_counter()
        st = bn.stats(); t5 = time.perf_counter()
        print("rep %d %s: refs %.1f ms, init_guess %.1f ms, solve%s %.1f ms (device %.1f), result %.1f ms, stats %.1f ms, total %.1f ms" % (
            rep, "direct" if direct else "staged", (t1 - t0) * 1e3, (t2 - t1) * 1e3, "+deliver" if direct else "", (t3 - t2) * 1e3, bn.timing()["total_ms"],
            (t4 - t3) * 1e3, (t5 - t4) * 1e3, (t5 - t0) * 1e3), file=sys.stderr, flush=True)
