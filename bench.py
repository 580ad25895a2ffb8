#!/usr/bin/env python
"""bench.py -- trajectory-Newton-iterations / second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--instances M] [--workload step|acro]
                    [--armijo lazy|speculative] [--state f32|f64] [--no-e2e] [--no-cpu]
    torchrun --nproc-per-node N bench.py --gpus N ...          (one rank per GPU, launched by the driver)

Workload (config.workload): BASELINE.json configs[3] -- "batched step-maneuver Newton, 65,536 randomised step
references", per GPU (weak scaling: rank r of world w holds instances r::w of a 65,536*w batch).  A "step" is ONE
Newton iteration (loop body of optcon.py:415-501: fused backward Riccati/costate sweep, LQ forward pass + descent,
Armijo candidate rollouts, update rollout) over the whole batch.  W warm-up iterations, then K timed ones -- these
are the iterations W..W+K-1 of the real solve from the device-generated initial guess; every instance is still
active there.  Time = CUDA events on the context's stream around the K iterations, max over ranks.

The JSON line carries: value (device-resident), e2e (full solve through the Python API from pinned host buffers:
H2D of the references, device initial guess, solve to the reference's criterion, D2H of results and stats),
roofline for the dominant kernel (HBM GB/s against MEASURED_PEAKS.json, FP64 TFLOP/s against a DFMA microbenchmark
run here), cpu_baseline (the C port of the reference in oracle/, OpenMP over instances, bounded sample), clocks.

--impl reference times that CPU port alone (the reference itself is Python and is not present on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "trajectory_newton_iterations_per_second"
UNIT = "traj-Newton-it/s"
TT = 1000

# algorithmic FP64 flops / HBM bytes per instance per time step (dense ns=6, ni=2 accounting)
FLOPS = dict(backward=2106, forward=126 + 30, cost=107, candidate=159, update=52 + 107)
BYTES = dict(backward=128 + 128, forward=128 + 64 + 16, cost=128, candidate=32 + 64, candidate_write=32 + 64 + 64, update=32 + 64 + 64)


def moved_bytes(args):
    """Bytes per instance per time step the kernels actually move in the selected mode.  The parity path stores the float32-quantised
    states as float (24 instead of 48 B per state read/write, bit-identical results); the FP32 mode halves everything."""
    if args.precision == "f32":
        return {k: v // 2 for k, v in BYTES.items()}
    if args.state == "f32" and args.x_storage == "auto":
        return dict(backward=BYTES["backward"] - 24, forward=BYTES["forward"] - 24, cost=BYTES["cost"] - 24, candidate=BYTES["candidate"],
                    candidate_write=BYTES["candidate_write"] - 24, update=BYTES["update"] - 24)
    return dict(BYTES)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU with nvidia-smi while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def make_problem(workload, n_total, lo_hi_stride, seed=None):
    """Per-rank shard of the batched configuration: instances rank::world of the n_total-instance batch."""
    from aircraftoptimalcontrol_b200 import refgen
    r, w = lo_hi_stride
    if workload == "step":
        zf, xf = refgen.config4_params(n_total, 2024 if seed is None else seed)
        xr, ur = refgen.step_problem(xf[r::w], zf[r::w], TT=TT)
        return xr, ur, None, refgen.weights("step")
    dx0, zf = refgen.config5_params(n_total, 7 if seed is None else seed)
    xr, ur = refgen.acrobatic_problem(zf[r::w], TT=TT)
    return xr, ur, np.ascontiguousarray(dx0[r::w]), refgen.weights("acro")


def pinned_like(a):
    """Copy a numpy array into pinned host memory (torch is only the allocator here)."""
    import torch
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    t.numpy()[...] = a
    return t  # keep the tensor alive; .numpy() is the view


def cpu_port_rate(workload, sample, W, K, state):
    """Newton iterations W..W+K-1 of `sample` instances with the C port of the reference on all host threads."""
    from oracle import corcl
    corcl.build()
    xr, ur, dx0, (Q, R, QT) = make_problem(workload, sample, (0, 1))
    xi = np.zeros((sample, 6, TT))
    ui = np.zeros((sample, 2, TT))
    for i in range(sample):  # the initial guess is an input of the hot path (float64 P-law rollout, same as the GPU's)
        xref_i = xr[i].copy()
        if dx0 is not None:
            xref_i[:, 0] += dx0[i]
        xi[i], ui[i] = corcl.initial_trajectory(xref_i if dx0 is not None else xr[i], quant_f32=(state == "f32"))
    nt = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    kw = dict(quant_f32=(state == "f32"), n_threads=nt)
    m = min(sample, nt)   # untimed: starts the OpenMP thread pool and faults in the library
    corcl.newton_batch(xr[:m], ur[:m], xi[:m], ui[:m], Q, R, QT, n_iters_cap=1, **kw)
    t0 = time.perf_counter()
    a = corcl.newton_batch(xr, ur, xi, ui, Q, R, QT, n_iters_cap=W, **kw) if W > 0 else None
    t1 = time.perf_counter()
    b = corcl.newton_batch(xr, ur, xi, ui, Q, R, QT, n_iters_cap=W + K, **kw)
    t2 = time.perf_counter()
    its = int(b["iters"].sum()) - (int(a["iters"].sum()) if a is not None else 0)
    dt = (t2 - t1) - (t1 - t0 if a is not None else 0.0)
    if dt <= 0 or its <= 0:  # (only with samples so small that timer noise exceeds K iterations: time the W+K run as a whole)
        dt, its = t2 - t1, int(b["iters"].sum())
    return its / dt, nt, dt, its


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    sample = args.cpu_sample
    rate, nt, dt, its = cpu_port_rate(args.workload, sample, args.warmup, args.steps, args.state)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(args, sample, world=1),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": nt, "kind": "port",
                         "sample": "%d instances x Newton iterations %d..%d of the workload, C port of the reference (oracle/acoc_oracle.c), "
                                   "OpenMP over instances; time(W+K iterations) - time(W iterations)" % (sample, args.warmup, args.warmup + args.steps - 1)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def config_dict(args, n_per_gpu, world):
    return {"workload": ("BASELINE.json configs[3]: batched step-maneuver Newton, randomised step references (zf~U(1.5,3.5), xf~U(14,18), seed 2024)"
                         if args.workload == "step" else
                         "BASELINE.json configs[4]: acrobatic Newton OCP batch (x0 perturbed, bump height zf~U(2.0,3.4), seed 7)"),
            "instances_per_gpu": n_per_gpu, "instances_total": n_per_gpu * world, "TT": TT, "ns": 6, "ni": 2,
            "state_quant": args.state, "precision": args.precision,
            "state_storage": "float32 in HBM (lossless: quantised states are float32 values)" if moved_bytes(args)["backward"] == 232 else
                             ("float32" if args.precision == "f32" else "float64"),
            "armijo": args.armijo, "armijo_maxiters": 10, "max_iters": 200,
            "step": "one Newton iteration over the whole batch (iterations W..W+K-1 of the solve)",
            "l2": "working set per GPU (%.1f GB) >> 126 MB L2, no flush needed" % (n_per_gpu * 400e3 / 1e9),
            "parallelism": "instances sharded round-robin, %d per GPU, no hot-path collective" % n_per_gpu}


def emit(line: dict):
    """Print the ONE JSON line on the process's original stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)


def main():
    os.dup2(2, 1)  # anything native code prints to fd 1 (e.g. the NCCL version banner) lands on stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=65536, help="instances per GPU")
    ap.add_argument("--workload", default="step", choices=["step", "acro"])
    ap.add_argument("--armijo", default="lazy", choices=["lazy", "speculative"])
    ap.add_argument("--state", default="f32", choices=["f32", "f64"])
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"], help="f64 = the parity path (default); f32 = the optional FP32 mode")
    ap.add_argument("--x-storage", default="auto", choices=["auto", "f64"], help="f64: keep float32-valued states in float64 buffers (A/B)")
    ap.add_argument("--no-tma", action="store_true", help="plain-load sweeps instead of the TMA rings (A/B)")
    ap.add_argument("--no-split", action="store_true", help="one stream for the whole batch instead of the two-range sweep (A/B)")
    ap.add_argument("--no-fused", action="store_true", help="separate LQ forward pass / candidate sweeps instead of the fused ones (A/B)")
    ap.add_argument("--cpu-sample", type=int, default=16384, help="instances of the bounded CPU sample (about 10 s of work on 16 host threads)")
    ap.add_argument("--chunks", type=int, default=4, help="sub-batches of the pipelined end-to-end solve (4: 0.42 s, 8: 0.44-0.46 s, 12: 0.52 s)")
    ap.add_argument("--no-stagger", action="store_true", help="end-to-end leg: same stream priority for every sub-batch (A/B)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import aircraftoptimalcontrol_b200 as pkg
    from aircraftoptimalcontrol_b200 import _lib, dist as D

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.instances
    n_total = n * world
    K, W = args.steps, args.warmup

    xr, ur, dx0, (Q, R, QT) = make_problem(args.workload, n_total, (rank, world))
    bn = pkg.BatchedNewton(n, TT=TT, device=local, state=args.state, armijo=args.armijo, precision=args.precision, x_storage=args.x_storage, tma=not args.no_tma, split=not args.no_split,
                           fused=not args.no_fused)
    bn.set_weights(Q, R, QT)
    bn.set_refs(xr, ur)
    bn.init_guess(dx0=dx0)
    if W:
        bn.iterate(W, count_active=False)
    bn.sync()

    state = {"bn_open": True}

    def barrier():
        if dist is not None:
            dist.barrier()
        if state["bn_open"]:
            bn.sync()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()  # keeps sampling through the timed region, the per-phase re-run and the end-to-end solve
    its_before = int(bn.stats()["iters"].astype(np.int64).sum())
    barrier()
    bn.iterate(K, count_active=False)  # K Newton iterations, timed by CUDA events inside the library
    bn.sync()
    barrier()
    tm = bn.timing()
    ms = tm["total_ms"]
    launches = tm["launches"]
    st = bn.stats()
    its_done = int(st["iters"].astype(np.int64).sum()) - its_before  # instance-iterations actually executed in the timed region
    if dist is not None:
        import torch
        t = torch.tensor([ms, float(launches), float(its_done)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, launches, its_done = float(tmax[0]), int(t[1]), int(t[2])
    n_active = int((st["status"] == 0).sum())
    value = its_done / (ms * 1e-3)   # = instances x K while every instance is still iterating (the default W, K)

    # ---- roofline of the dominant kernel (rank 0): re-run iterations with per-phase events --------------------
    roofline = fp64 = phases = None
    peaks, peak_src = load_peaks()
    if rank == 0 and not args.no_roofline:
        bn.set_profiling(True)
        bn.init_guess(dx0=dx0)
        bn.iterate(W, count_active=False)
        bn.iterate(K, count_active=False)
        tp = bn.timing()
        bn.set_profiling(False)
        h = bn.history()
        ncand = h["n_armijo"][:, W:W + K].astype(np.float64)
        steps_per = float(TT - 1)
        # algorithmic bytes / flops of every phase over the K timed iterations (SURVEY.md 8(d) per-unit figures)
        # lazy search on the TMA path: the LQ forward pass and candidate 0 are ONE sweep (k_forward_cand0_tma); its time is the
        # "forward" phase and it is accounted with the algorithmic bytes / flops of both (SURVEY.md 8(d) counts du written and read
        # back and u read twice: 368 B; the fused kernel moves 276 B of them with float state slots)
        fused_fc = args.armijo == "lazy" and not args.no_tma and not args.no_fused
        MB = moved_bytes(args)
        fwd_bytes, fwd_flops, fwd_moved = BYTES["forward"], FLOPS["forward"], MB["forward"]
        if fused_fc:
            fwd_bytes += BYTES["candidate_write"]
            fwd_flops += FLOPS["candidate"]
            # minus the du read-back and the second read of u, minus the state components the ring does not fetch (only V, theta, gamma)
            x_float = args.precision == "f32" or MB["forward"] != BYTES["forward"]
            fwd_moved += MB["candidate_write"] - (16 if args.precision == "f32" else 32) - (12 if x_float else 24)
        if args.armijo == "lazy":
            # candidate 0 for everyone, written tentatively (it is the update when accepted); the other 9 candidates and a
            # separate update rollout only for instances whose candidate 0 failed
            n_fail = float(np.sum(ncand > 1))
            n_c0 = 0.0 if fused_fc else float(ncand.size)
            cand_bytes = steps_per * (n_c0 * BYTES["candidate_write"] + 9 * n_fail * BYTES["candidate"])
            cand_flops = steps_per * (n_c0 + 9 * n_fail) * FLOPS["candidate"]
            upd_units = n_fail
        else:
            cand_bytes = steps_per * ncand.size * 10 * BYTES["candidate"]
            cand_flops = steps_per * ncand.size * 10 * FLOPS["candidate"]
            upd_units = float(ncand.size)
        phases = tp["phases"]
        per_launch_ms = {k: v / K for k, v in phases.items()}
        dom = max(("backward", "forward", "candidates", "update"), key=lambda k: phases[k])
        tot_bytes = {"backward": steps_per * n * K * BYTES["backward"], "forward": steps_per * n * K * fwd_bytes,
                     "candidates": cand_bytes, "update": steps_per * upd_units * BYTES["update"]}
        tot_flops = {"backward": steps_per * n * K * FLOPS["backward"], "forward": steps_per * n * K * fwd_flops,
                     "candidates": cand_flops, "update": steps_per * upd_units * FLOPS["update"]}
        fp64_peak = _lib.measure_fp64_peak(local)
        moved_ratio = {"backward": MB["backward"] / BYTES["backward"], "forward": fwd_moved / fwd_bytes,
                       "update": MB["update"] / BYTES["update"],
                       "candidates": ((n_c0 * MB["candidate_write"] + 9 * n_fail * MB["candidate"]) / max(n_c0 * BYTES["candidate_write"] + 9 * n_fail * BYTES["candidate"], 1.0)
                                      if args.armijo == "lazy" else MB["candidate"] / BYTES["candidate"])}
        tbl = {}
        for k in ("backward", "forward", "candidates", "update"):
            if phases[k] <= 0 or tot_bytes[k] <= 0:
                continue
            gbs = tot_bytes[k] / (phases[k] * 1e-3) / 1e9
            tfs = tot_flops[k] / (phases[k] * 1e-3) / 1e12
            tbl[k] = {"ms_per_iteration": per_launch_ms[k], "hbm_gbs": gbs, "hbm_frac": gbs / peaks["hbm_gbs"], "fp64_tflops": tfs, "fp64_frac": tfs / fp64_peak,
                      "hbm_gbs_moved": gbs * moved_ratio[k], "hbm_frac_moved": gbs * moved_ratio[k] / peaks["hbm_gbs"]}
        whole = sum(tot_bytes.values()) / (sum(phases.values()) * 1e-3) / 1e9
        tbl["whole_iteration"] = {"ms_per_iteration": sum(phases.values()) / K, "hbm_gbs": whole, "hbm_frac": whole / peaks["hbm_gbs"]}
        d = tbl[dom]
        # HBM is the binding resource: the bytes are irreducible, while the kernels execute far fewer flops than the dense
        # accounting of SURVEY.md 8(d) (sparsity of A, B and symmetry of P), so the FP64 figure is reported beside it.
        traffic = None
        try:  # DRAM bytes per launch of the dominant kernel from the committed ncu capture, scaled to this instance count
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            traffic = (tj.get("float_state", {}) if MB["backward"] == 232 else tj).get("k_" + dom) if args.precision == "f64" else None
            if traffic is not None:
                traffic = traffic * n / tj["instances"]
        except Exception:
            pass
        roofline = {"kernel": "k_" + dom, "bound": "hbm", "achieved": d["hbm_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": d["hbm_frac"], "traffic": traffic, "traffic_source": "profiles/r01_traffic.json (ncu --set full, per launch)",
                    "algorithmic_bytes_per_launch": tot_bytes[dom] / K, "peak_source": peak_src,
                    "fp64": {"achieved_algorithmic_tflops": d["fp64_tflops"], "peak_tflops": fp64_peak, "frac_algorithmic": d["fp64_frac"],
                             "peak_source": "DFMA microbenchmark run in this process (acoc_measure_fp64_peak)",
                             "note": "algorithmic = dense ns=6/ni=2 flop count of SURVEY.md 8(d); executed flops are lower (structure-exploiting sweep)"},
                    "per_kernel": tbl, "forward_phase": ("k_forward_cand0_tma: LQ forward pass + candidate 0 in one sweep" if fused_fc else "k_forward"),
                    "share_of_step": phases[dom] / max(sum(phases.values()), 1e-9),
                    "algorithmic": {"bytes_per_instance_step": BYTES, "flops_per_instance_step": FLOPS,
                                    "note": "SURVEY.md 8(d) per-unit figures x (TT-1) x instances per launch"},
                    "moved": {"bytes_per_instance_step": MB, "achieved": d["hbm_gbs_moved"], "frac": d["hbm_frac_moved"],
                              "note": "bytes the kernels move in this mode (float32-valued states stored as float: -24 B per state access); "
                                      "'achieved'/'frac' above use the SURVEY.md 8(d) float64 accounting, so they can exceed the moved figure"}}
        fp64 = {"peak_tflops_measured": fp64_peak, "dependent_dfma_latency_cycles": _lib.measure_fp64_latency(local)}

    # ---- whole solve, device-resident (every instance to the reference's criterion; includes the float32-noise phase
    #      with its full Armijo searches and the thinning tail) ------------------------------------------------------
    bn.init_guess(dx0=dx0)
    bn.solve()  # untimed: the first solve allocates the survivor-generation contexts (cudaMalloc inside the timed span otherwise)
    bn.init_guess(dx0=dx0)
    tot_solve = bn.solve()
    tsolve = bn.timing()
    whole = {"value": tot_solve / (tsolve["total_ms"] * 1e-3), "unit": UNIT, "device_ms": tsolve["total_ms"], "total_newton_iterations": int(tot_solve),
             "lockstep_iterations": int(bn.stats()["iters"].max()), "gpu_launches": tsolve["launches"],
             "survivor_generation_moves_ms": tsolve["phases"]["select"], "scope": "this rank"}

    device_bytes = bn.device_bytes
    bn.close()  # the end-to-end leg below builds its own contexts: release this one (and its survivor generations) first
    state["bn_open"] = False

    # ---- end to end through the public API, host buffers, full solve -----------------------------------------
    e2e = None
    if not args.no_e2e:
        xr_p, ur_p = pinned_like(xr), pinned_like(ur)
        import torch
        xs_t = torch.empty((n, 6, TT), dtype=torch.float64, pin_memory=True)
        us_t = torch.empty((n, 2, TT), dtype=torch.float64, pin_memory=True)
        pn = pkg.PipelinedNewton(n, n_chunks=args.chunks, TT=TT, device=local, state=args.state, armijo=args.armijo, precision=args.precision,
                                 x_storage=args.x_storage, tma=not args.no_tma, split=not args.no_split, fused=not args.no_fused,
                                 stagger=not args.no_stagger)
        pn.set_weights(Q, R, QT)
        pn.solve(xr_p.numpy(), ur_p.numpy(), dx0=dx0, out=(xs_t.numpy(), us_t.numpy()))   # untimed warm-up of the whole path
        barrier()
        t0 = time.perf_counter()
        # H2D references -> device initial guess (N1) -> solve every instance to descent >= -1e-6 -> D2H trajectories + stats,
        # pipelined over independent sub-batches
        _, _, st2 = pn.solve(xr_p.numpy(), ur_p.numpy(), dx0=dx0, out=(xs_t.numpy(), us_t.numpy()))
        barrier()
        t1 = time.perf_counter()
        pn.close()
        wall = t1 - t0
        g = D.gather_stats(st2, n_total)                 # NCCL all_gather of the per-instance statistics (off the hot path)
        t2 = time.perf_counter()
        tot = int(g["iters"].sum())
        if dist is not None:
            import torch
            tw = torch.tensor([wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            wall = float(tw[0])
        steps_e2e = int(g["iters"].max())
        h2d = (xr.nbytes + ur.nbytes) * world
        d2h = (xs_t.numel() + us_t.numel()) * 8 * world + n_total * 28
        e2e = {"value": tot / wall, "unit": UNIT, "h2d_bytes_per_step": h2d / max(steps_e2e, 1), "d2h_bytes_per_step": d2h / max(steps_e2e, 1),
               "wall_s": wall, "total_newton_iterations": tot, "solver_steps": steps_e2e,
               "converged": int((g["status"] == 1).sum()), "instances": n_total, "mean_iters": float(g["iters"].mean()),
               "stats_gather_s": t2 - t1,
               "chunks": args.chunks, "staggered_priorities": not args.no_stagger,
               "what": "PipelinedNewton.solve(pinned host refs): per sub-batch set_refs (H2D) -> init_guess (device) -> solve() to descent >= -1e-6 "
                       "-> result()/stats() (D2H to pinned host); sub-batches overlap copies with compute"}

    clocks = sampler.summary() if sampler else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, nt, dt, its = cpu_port_rate(args.workload, args.cpu_sample, W, K, args.state)
        cpu = {"value": rate, "unit": UNIT, "cores": nt, "kind": "port",
               "sample": "%d instances x Newton iterations %d..%d of the same workload, C port of the reference (oracle/acoc_oracle.c), OpenMP over "
                         "instances, %.1f s" % (args.cpu_sample, W, W + K - 1, dt)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / max(K, 1),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": config_dict(args, n, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "active_after_timed_region": n_active, "instance_iterations_timed": its_done, "whole_solve": whole, "roofline": roofline, "fp64": fp64, "phase_ms": phases, "cpu_baseline": cpu,
                "device": pkg.device_info(local)["name"], "device_bytes": device_bytes}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
