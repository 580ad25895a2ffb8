"""Live import of the UNMODIFIED reference as a parity checker (TEST INFRASTRUCTURE ONLY).

Nothing in the product package may import this module.  It is used by
``oracle/gen_golden.py`` (run in the build container, where ``/root/reference``
is mounted) to produce the committed fixtures under ``tests/golden/`` and by
``-m "not gpu"`` tests that pin the C restatement (``oracle/acoc_oracle.c``)
against the reference when the reference tree is present.

The reference (``/root/reference``) is six flat Python modules; ``optcon.py:2``
imports matplotlib and the two Newton scripts import ``cvxpy`` and ``animate``
(``main_newton_method.py:10-13``), none of which exist in this image, so they are
stubbed with ``MagicMock`` before the import.

Two flavours are offered:

* ``load(f64_state=False)`` -- the reference exactly as shipped: every ``step``
  rounds the next state to float32 (``aircraft_simplified.py:300``).
* ``load(f64_state=True)``  -- a temp copy whose ONLY edit is ``np.float32`` ->
  ``np.float64`` on that line; this is the "clean" mode used for the 1e-9 bar.
"""
from __future__ import annotations

import importlib.util
import io
import os
import shutil
import sys
import tempfile
import types
from contextlib import redirect_stdout
from unittest.mock import MagicMock

import numpy as np

REFERENCE_ROOT = os.environ.get("ACOC_REFERENCE_ROOT", "/root/reference")

_STUBS = ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.ticker", "cvxpy")
_cache: dict = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "optcon.py"))


def _install_stubs() -> None:
    for m in _STUBS:
        sys.modules.setdefault(m, MagicMock())
    if "animate" not in sys.modules:
        fake = types.ModuleType("animate")
        fake.Airfoil = MagicMock()
        sys.modules["animate"] = fake


def _import_from(path: str, name: str, alias: str):
    spec = importlib.util.spec_from_file_location(alias, os.path.join(path, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class RefModules:
    """Handle on one flavour of the reference: ``.aircraft`` and ``.optcon`` modules."""

    def __init__(self, aircraft, optcon, root, f64_state):
        self.aircraft = aircraft
        self.optcon = optcon
        self.root = root
        self.f64_state = f64_state


def load(f64_state: bool = False) -> RefModules:
    if f64_state in _cache:
        return _cache[f64_state]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    import warnings

    root = REFERENCE_ROOT
    if f64_state:
        root = tempfile.mkdtemp(prefix="acoc_ref_f64_")
        for f in ("aircraft_simplified.py", "optcon.py"):
            shutil.copy(os.path.join(REFERENCE_ROOT, f), root)
        p = os.path.join(root, "aircraft_simplified.py")
        src = open(p).read().split("\n")
        assert "np.float32" in src[299], src[299]  # aircraft_simplified.py:300
        src[299] = src[299].replace("np.float32", "np.float64")
        open(p, "w").write("\n".join(src))
    tag = "f64" if f64_state else "f32"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # SyntaxWarning from '\i' in the docstrings
        ac = _import_from(root, "aircraft_simplified", "_acoc_ref_aircraft_" + tag)
        oc = _import_from(root, "optcon", "_acoc_ref_optcon_" + tag)
    mods = RefModules(ac, oc, root, f64_state)
    _cache[f64_state] = mods
    return mods


# ----------------------------------------------------------------------------------------
# Full Newton solve with history capture (no stdout parsing: the wrappers see the exact
# float64 values that optcon.py:482 / :488 pass around).
# ----------------------------------------------------------------------------------------
def run_newton(mods: RefModules, xx_ref, uu_ref, xx_init, uu_init, QQt, RRt, QQT,
               tf=1, dt=1e-3, max_iters=200, stepsize_0=1, cc=0.5, beta=0.7,
               armijo_maxiters=10, params=None, keep_iterates=(), capture_lq_at=()):
    """Run ``NewtonMethod.optimize`` (optcon.py:341) and return its observable history.

    Returns a dict with
      JJ[k], descent[k]   : cost / descent seen by the Armijo call of iteration k (optcon.py:482)
      stepsize[k]         : value returned by ``armijo_stepsize`` (optcon.py:327)
      n_armijo[k]         : number of candidates the sequential search rolled out
      iters               : number of loop bodies executed (k = 0..iters-1)
      xx_star, uu_star    : what ``optimize`` returned (iterate kk-1, optcon.py:503-505)
      xx_last, uu_last    : the last iterate written by ``get_update``
      iterates[k]         : (xx, uu) produced by get_update in iteration k, for k in keep_iterates
      lq[k]               : dict(KK, deltax, deltau) of the ltv_LQR call in iteration k (capture_lq_at)
    """
    ac, oc = mods.aircraft, mods.optcon
    dyn = ac.Dynamics()
    dyn.dt = dt
    if params is not None:
        for k, v in params.items():
            setattr(dyn, k, v)
    cst = ac.Cost(QQt, RRt, QQT)
    NM = oc.NewtonMethod(dyn, cst, xx_ref, uu_ref, max_iters=max_iters, stepsize_0=stepsize_0,
                         cc=cc, beta=beta, armijo_maxiters=armijo_maxiters, term_cond=1e-6)
    hist = dict(JJ=[], descent=[], stepsize=[], n_armijo=[], iterates={}, lq={})
    orig_armijo, orig_update, orig_step = NM.armijo_stepsize, NM.get_update, dyn.step
    orig_lq = oc.ltv_LQR
    state = dict(in_armijo=False, steps=0, k=0)

    def step_counting(*a):
        if state["in_armijo"]:
            state["steps"] += 1
        return orig_step(*a)

    def armijo(uu, deltau, xr, ur, x0, TT, JJ, descent, JP):
        state["in_armijo"], state["steps"] = True, 0
        s = orig_armijo(uu, deltau, xr, ur, x0, TT, JJ, descent, JP)
        state["in_armijo"] = False
        hist["JJ"].append(float(JP))
        hist["descent"].append(float(descent))
        hist["stepsize"].append(float(s))
        hist["n_armijo"].append(state["steps"] // (TT - 1))
        return s

    def update(stepsize, uu, deltau, x0):
        xx_t, uu_t = orig_update(stepsize, uu, deltau, x0)
        k = state["k"]
        if k in keep_iterates:
            hist["iterates"][k] = (xx_t.copy(), uu_t.copy())
        hist["xx_last"], hist["uu_last"] = xx_t.copy(), uu_t.copy()
        state["k"] = k + 1
        return xx_t, uu_t

    def lq(*a, **kw):
        out = orig_lq(*a, **kw)
        k = state["k"]
        if k in capture_lq_at:
            hist["lq"][k] = dict(KK=out[0].copy(), deltax=out[2].copy(), deltau=out[3].copy())
        return out

    dyn.step = step_counting
    NM.armijo_stepsize = armijo
    NM.get_update = update
    oc.ltv_LQR = lq
    try:
        with redirect_stdout(io.StringIO()):
            xs, us = NM.optimize(np.array(xx_init, dtype=np.float64), np.array(uu_init, dtype=np.float64), tf, dt)
    finally:
        oc.ltv_LQR = orig_lq
    hist["xx_star"], hist["uu_star"] = np.array(xs), np.array(us)
    hist["iters"] = len(hist["JJ"])
    for k in ("JJ", "descent", "stepsize"):
        hist[k] = np.array(hist[k], dtype=np.float64)
    hist["n_armijo"] = np.array(hist["n_armijo"], dtype=np.int32)
    return hist


# ----------------------------------------------------------------------------------------
# GradientMethod.optimize (optcon.py:27-174).  The reference's own call of its line search
# (optcon.py:125) passes 8 of armijo_stepsize's 9 arguments and raises TypeError, so the
# method cannot run as shipped.  The reference SOURCE stays untouched here: the instance's
# armijo_stepsize is replaced by a call adapter that accepts the 8 arguments :125 passes and
# forwards them to the original method with the missing JP = JJ[kk] and with the directional
# derivative -descent[kk] (descent[kk] = sum |deltau|^2 > 0, :118) as the slope of the
# sufficient-decrease test (:268).  This is the repair include/acoc.h documents for
# ACOC_METHOD_GRADIENT and oracle/acoc_oracle.c::orc_gradient restates.
# ----------------------------------------------------------------------------------------
def run_gradient(mods: RefModules, xx_ref, uu_ref, xx_init, uu_init, QQt, RRt, QQT,
                 tf=1, dt=1e-3, max_iters=200, stepsize_0=1e-2, cc=0.5, beta=0.7, armijo_maxiters=20, keep_iterates=()):
    """History dict like run_newton (descent = the reference's positive descent[kk]); deltau_first = deltau of iteration 0."""
    ac, oc = mods.aircraft, mods.optcon
    dyn = ac.Dynamics()
    dyn.dt = dt
    cst = ac.Cost(QQt, RRt, QQT)
    GM = oc.GradientMethod(dyn, cst, xx_ref, uu_ref, max_iters=max_iters, stepsize_0=stepsize_0, cc=cc, beta=beta,
                           armijo_maxiters=armijo_maxiters, term_cond=1e-6)
    hist = dict(JJ=[], descent=[], stepsize=[], n_armijo=[], iterates={})
    orig_armijo, orig_update, orig_step = GM.armijo_stepsize, GM.get_update, dyn.step
    state = dict(in_armijo=False, steps=0, k=0)

    def step_counting(*a):
        if state["in_armijo"]:
            state["steps"] += 1
        return orig_step(*a)

    def armijo_adapter(uu, deltau, xr, ur, x0, TT, JJk, descentk):   # the 8 arguments of optcon.py:125
        if state["k"] == 0:
            hist["deltau_first"] = np.array(deltau)
        state["in_armijo"], state["steps"] = True, 0
        s = orig_armijo(uu, deltau, xr, ur, x0, TT, JJk, -descentk, JJk)
        state["in_armijo"] = False
        hist["JJ"].append(float(JJk))
        hist["descent"].append(float(descentk))
        hist["stepsize"].append(float(s))
        hist["n_armijo"].append(state["steps"] // (TT - 1))
        return s

    def update(stepsize, uu, deltau, x0):
        xx_t, uu_t = orig_update(stepsize, uu, deltau, x0)
        k = state["k"]
        if k in keep_iterates:
            hist["iterates"][k] = (xx_t.copy(), uu_t.copy())
        hist["xx_last"], hist["uu_last"] = xx_t.copy(), uu_t.copy()
        state["k"] = k + 1
        return xx_t, uu_t

    dyn.step = step_counting
    GM.armijo_stepsize = armijo_adapter
    GM.get_update = update
    with redirect_stdout(io.StringIO()):
        xs, us = GM.optimize(np.array(xx_init, dtype=np.float64), np.array(uu_init, dtype=np.float64), tf, dt)
    hist["xx_star"], hist["uu_star"] = np.array(xs), np.array(us)
    hist["iters"] = len(hist["JJ"])
    for k in ("JJ", "descent", "stepsize"):
        hist[k] = np.array(hist[k], dtype=np.float64)
    hist["n_armijo"] = np.array(hist["n_armijo"], dtype=np.int32)
    return hist


def run_script(name: str, workdir: str | None = None, skip_optimize: bool = False):
    """Run one of the reference's scripts unmodified (runpy) from a writable copy; returns its globals.

    ``skip_optimize`` replaces ``NewtonMethod.optimize`` by a stub that hands back the initial guess, so
    that only the script's problem set-up (references, equilibrium, initial trajectory) is executed.
    """
    import runpy
    import warnings

    _install_stubs()
    wd = workdir or tempfile.mkdtemp(prefix="acoc_ref_run_")
    for f in os.listdir(REFERENCE_ROOT):
        if f.endswith(".py"):
            shutil.copy(os.path.join(REFERENCE_ROOT, f), wd)
    os.makedirs(os.path.join(wd, "Data"), exist_ok=True)
    os.makedirs(os.path.join(wd, "Figures"), exist_ok=True)
    for f in os.listdir(os.path.join(REFERENCE_ROOT, "Data")):
        shutil.copy(os.path.join(REFERENCE_ROOT, "Data", f), os.path.join(wd, "Data"))
    cwd = os.getcwd()
    sys.path.insert(0, wd)
    for m in ("optcon", "aircraft_simplified"):
        sys.modules.pop(m, None)
    try:
        os.chdir(wd)
        with warnings.catch_warnings(), redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            if skip_optimize:
                import optcon as _oc  # the copy in wd (sys.path[0])
                _oc.NewtonMethod.optimize = lambda self, xi, ui, tf, dt: (xi.copy(), ui.copy())
            g = runpy.run_path(os.path.join(wd, name), run_name="__main__")
    finally:
        os.chdir(cwd)
        sys.path.remove(wd)
        for m in ("optcon", "aircraft_simplified"):
            sys.modules.pop(m, None)
    return g
