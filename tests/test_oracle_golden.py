"""Pins oracle/acoc_oracle.c (the CPU restatement) against fixtures produced by the LIVE unmodified reference
(oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from tests.util import golden, relerr


def test_step_kat_bit_exact_next_state(oracle):
    d = golden("step_kat.npz")
    for i in range(len(d["x"])):
        x, u, lam = d["x"][i], d["u"][i], d["lam"][i]
        o32 = oracle.step(x, u, None, quant_f32=True)
        o64 = oracle.step(x, u, None, quant_f32=False)
        ol = oracle.step(x, u, lam, quant_f32=False)
        # aircraft_simplified.py:303-310 incl. the float32 rounding of :300 -- bit for bit
        assert np.array_equal(o32[0], d["xxp32"][i])
        assert np.array_equal(o64[0], d["xxp64"][i])
        # derivatives: same math, different association -> 1e-12 of the largest entry
        for got, ref in ((o64[1], d["fx"][i]), (o64[2], d["fu"][i]), (o64[3], d["fxx"][i]), (o64[5], d["fux"][i]),
                         (ol[3], d["fxxc"][i]), (ol[5], d["fuxc"][i])):
            assert got.shape == ref.shape
            assert relerr(ref, got) < 1e-12
    assert d["fuu_zero"].all()


def test_cost_kat(oracle):
    d = golden("cost_kat.npz")
    for i in range(len(d["x"])):
        w = int(d["which"][i])
        ll, lx, lu = oracle.stagecost(d["Q"][w], d["R"][w], d["x"][i], d["u"][i], d["xr"][i], d["ur"][i])
        lt, ltx = oracle.termcost(d["QT"][w], d["x"][i], d["xr"][i])
        assert abs(ll - d["ll"][i]) <= 1e-13 * abs(d["ll"][i])
        assert abs(lt - d["llT"][i]) <= 1e-13 * abs(d["llT"][i])
        assert relerr(d["lx"][i], lx) < 1e-14 and relerr(d["lu"][i], lu) < 1e-14 and relerr(d["lTx"][i], ltx) < 1e-14


def test_ltv_lqr_forced_regularisation(oracle):
    """optcon.py:745-749: the +0.5*I branch, augmented and plain."""
    d = golden("lq_forced_reg.npz")
    TT = d["A"].shape[2]
    K, P, x, u, n = oracle.ltv_lqr(d["A"], d["B"], d["Q"], d["R"], d["S"], d["Qf"], TT, np.zeros(6), d["q"], d["r"], d["qf"], return_nreg=True)
    assert n == int(d["n_reg_aug"]) > 0
    assert K.shape == (2, 7, TT) and P.shape == (7, 7, TT)
    for got, ref in ((K, d["K_aug"]), (P, d["P_aug"]), (x, d["x_aug"]), (u, d["u_aug"])):
        assert relerr(ref, got) < 1e-10
    K, P, x, u, n = oracle.ltv_lqr(d["A"], d["B"], d["Q"], d["R"], d["S"], d["Qf"], TT, d["x0"], return_nreg=True)
    assert n == int(d["n_reg_non"]) > 0
    assert K.shape == (2, 6, TT)
    for got, ref in ((K, d["K_non"]), (P, d["P_non"]), (x, d["x_non"]), (u, d["u_non"])):
        assert relerr(ref, got) < 1e-10


@pytest.mark.parametrize("name", ["newton_step_f32", "newton_step_f64", "newton_acro_f32", "newton_acro_f64"])
def test_newton_history(oracle, name):
    """Configs 1 and 2, both state quantisations: same iteration count, same Armijo step at EVERY iteration,
    cost/descent history and trajectories to 1e-9 (bit-identical states in float32 mode)."""
    d = golden(name + ".npz")
    f64 = name.endswith("f64")
    h = oracle.newton(d["xx_ref"], d["uu_ref"], d["xx_init"], d["uu_init"], d["Q"], d["R"], d["QT"], quant_f32=not f64)
    k = int(d["iters"])
    assert h["iters"] == k
    assert np.array_equal(h["stepsize"], d["stepsize"])
    assert np.array_equal(h["n_armijo"], d["n_armijo"])
    assert np.max(np.abs(h["JJ"] - d["JJ"]) / np.abs(d["JJ"])) < 1e-12
    assert np.max(np.abs(h["descent"] - d["descent"]) / np.abs(d["descent"])) < 1e-9
    assert relerr(d["xx_star"], h["xx_star"]) < 1e-9 and relerr(d["uu_star"], h["uu_star"]) < 1e-9
    assert relerr(d["xx_last"], h["xx_last"]) < 1e-9 and relerr(d["uu_last"], h["uu_last"]) < 1e-9
    if not f64:
        assert np.array_equal(h["xx_star"], d["xx_star"])  # float32-quantised states: bit-identical


def test_lq_inside_newton(oracle):
    """ltv_LQR on the sub-problems captured from live Newton iterations kk in {0,5,9,15} (Gauss-Newton and exact
    Hessian phases): rebuild A,B,Q,S,q,r from the stored iterate with the oracle's own step/cost and compare
    deltau / deltax / gains with the reference's."""
    d = golden("newton_acro_f32.npz")
    xr, ur, Q, R, QT = d["xx_ref"], d["uu_ref"], d["Q"], d["R"], d["QT"]
    TT = xr.shape[1]
    for kk in d["lq_at"]:
        xx, uu = d["it%d_xx" % kk], d["it%d_uu" % kk]
        A, B = np.zeros((6, 6, TT)), np.zeros((6, 2, TT))
        Qs, Rs, Ss = np.zeros((6, 6, TT)), np.zeros((2, 2, TT)), np.zeros((2, 6, TT))
        q, r = np.zeros((6, TT)), np.zeros((2, TT))
        lam = oracle.termcost(QT, xx[:, -1], xr[:, -1])[1]
        Qs[:, :, -1] = QT
        q[:, -1] = lam
        for t in reversed(range(TT - 1)):
            _, lx, lu = oracle.stagecost(Q, R, xx[:, t], uu[:, t], xr[:, t], ur[:, t])
            _, fx, fu, fxx, _, fux = oracle.step(xx[:, t], uu[:, t], lam, quant_f32=True)
            A[:, :, t], B[:, :, t] = fx.T, fu.T
            Qs[:, :, t] = Q + (fxx if kk > 8 else 0)
            Rs[:, :, t] = R
            Ss[:, :, t] = fux if kk > 8 else 0
            q[:, t], r[:, t] = lx, lu
            lam = fx @ lam + lx
        K, _, dx, du = oracle.ltv_lqr(A, B, Qs, Rs, Ss, QT, TT, np.zeros(6), q, r, q[:, -1])
        assert relerr(d["it%d_deltau" % kk], du) < 1e-9
        assert relerr(d["it%d_deltax" % kk], dx) < 1e-9
        assert relerr(d["it%d_KK" % kk], K) < 1e-9


def test_lqr_tracking(oracle):
    d = golden("lqr_tracking.npz")
    xr, ur, K = oracle.lqr_tracking(d["xx_opt"], d["uu_opt"], d["Q"], d["R"], d["QT"], d["delta"])
    assert relerr(d["KK"], K) < 1e-11
    assert np.array_equal(xr, d["xx_reg"])  # float32-quantised closed loop: bit-identical
    assert relerr(d["uu_reg"], ur) < 1e-11
    assert not ur[:, :, -1].any()


def test_armijo_exhaustion_returns_untested_step(oracle):
    """optcon.py:268-273, :327: when every candidate fails the search returns stepsize_0*beta**maxiters."""
    d = golden("newton_step_f32.npz")
    TT = d["xx_ref"].shape[1]
    du = np.zeros((2, TT))
    du[0] = 1e3  # an ascent direction: every candidate is worse than JP with a (claimed) negative descent
    JP = oracle.traj_cost(d["Q"], d["R"], d["QT"], d["xx_init"], d["uu_init"], d["xx_ref"], d["uu_ref"])
    s, costs, ntried, accepted = oracle.armijo(d["xx_init"][:, 0], d["uu_init"], du, d["Q"], d["R"], d["QT"], d["xx_ref"], d["uu_ref"],
                                                JP, -1.0)
    sref = 1.0
    for _ in range(10):
        sref = 0.7 * sref
    assert not accepted and ntried == 10 and s == sref and np.all(costs > JP)


def test_newton_return_slot_quirks(oracle):
    """optcon.py:499-505: (a) max_iters exhausted -> last iterate written; (b) converged at kk = 0 -> the all-zero slot -1."""
    d = golden("newton_quirks.npz")
    a = oracle.newton(d["xx_ref"], d["uu_ref"], d["a_xx_init"], d["a_uu_init"], d["Q"], d["R"], d["QT"], quant_f32=False, max_iters=4)
    assert a["iters"] == int(d["a_iters"]) == 3 and np.array_equal(a["stepsize"], d["a_stepsize"])
    assert relerr(d["a_xx_star"], a["xx_star"]) < 1e-9 and relerr(d["a_uu_star"], a["uu_star"]) < 1e-9
    b = oracle.newton(d["xx_ref"], d["uu_ref"], d["b_xx_init"], d["b_uu_init"], d["Q"], d["R"], d["QT"], quant_f32=False)
    assert b["iters"] == int(d["b_iters"]) == 1 and np.array_equal(b["stepsize"], d["b_stepsize"])
    assert not b["xx_star"].any() and not b["uu_star"].any() and not d["b_xx_star"].any()


@pytest.mark.parametrize("name", ["step_f32", "step_f64", "acro_f32"])
def test_gradient_method_pinned(oracle, name):
    """orc_gradient (GradientMethod.optimize, optcon.py:27-174, with the repaired line-search call) against the live reference run
    through the call adapter of oracle/pyref.py::run_gradient: identical Armijo steps and candidate counts, histories, iterates."""
    g = golden("gradient_%s.npz" % name)
    d = golden(str(g["base"]))
    o = oracle.gradient(d["xx_ref"], d["uu_ref"], d["xx_init"], d["uu_init"], d["Q"], d["R"], d["QT"], quant_f32=name.endswith("f32"),
                        max_iters=int(g["max_iters"]), stepsize_0=float(g["stepsize_0"]), cc=float(g["cc"]), beta=float(g["beta"]),
                        armijo_maxiters=int(g["armijo_maxiters"]))
    assert o["iters"] == int(g["iters"])
    assert np.array_equal(o["stepsize"], g["stepsize"]) and np.array_equal(o["n_armijo"], g["n_armijo"])
    assert np.max(np.abs(o["JJ"] - g["JJ"]) / g["JJ"]) < 1e-12 and np.max(np.abs(o["descent"] - g["descent"]) / g["descent"]) < 1e-9
    assert relerr(g["deltau_first"], o["deltau_first"]) < 1e-12
    if name.endswith("f32"):
        assert np.array_equal(o["xx_last"], g["xx_last"]) and np.array_equal(o["xx_star"], g["xx_star"])
    assert relerr(g["xx_last"], o["xx_last"]) < 1e-9 and relerr(g["uu_last"], o["uu_last"]) < 1e-9
    assert relerr(g["uu_star"], o["uu_star"]) < 1e-9


@pytest.mark.parametrize("cfg,idx", [("config4", 0), ("config4", 1), ("config5", 0), ("config5", 1)])
def test_batched_config_instances_pinned_to_the_reference(oracle, cfg, idx):
    """the oracle the GPU's batched parity tests compare with is itself validated against the live Python reference on
    instances of the BATCHED configurations -- instances 0, 1 of config 4 (randomised step references, seed 2024) and of config 5
    (acrobatic, x0 + delta, bump height, seed 7), TT = 1000, float32-quantised state: same iteration count, every Armijo step and
    candidate count, histories to 1e-12 / 1e-9, result bit-identical (oracle/gen_golden.py::gen_batched_instances)."""
    from aircraftoptimalcontrol_b200 import refgen
    d = golden("newton_batched_instances.npz")
    t = "%s_%d_" % (cfg, idx)
    if cfg == "config4":
        zf, xf = refgen.config4_params()
        assert np.array_equal(d[t + "par"], [zf[idx], xf[idx]])
        xr, ur = refgen.step_problem(xf[idx:idx + 1], zf[idx:idx + 1])
        Q, R, QT = refgen.weights("step")
    else:
        dx0, zf = refgen.config5_params()
        assert np.array_equal(d[t + "par"], [zf[idx]]) and np.array_equal(d[t + "dx0"], dx0[idx])
        xr, ur = refgen.acrobatic_problem(zf[idx:idx + 1])
        Q, R, QT = refgen.weights("acro")
    start = xr[0].copy()
    start[:, 0] += d[t + "dx0"]
    xi, ui = oracle.initial_trajectory(start, quant_f32=True)
    assert np.array_equal(xi, d[t + "xx_init"]) and np.array_equal(ui, d[t + "uu_init"])   # the input both sides were given
    h = oracle.newton(xr[0], ur[0], xi, ui, Q, R, QT, quant_f32=True)
    assert h["iters"] == int(d[t + "iters"])
    assert np.array_equal(h["stepsize"], d[t + "stepsize"]) and np.array_equal(h["n_armijo"], d[t + "n_armijo"])
    assert np.max(np.abs(h["JJ"] - d[t + "JJ"]) / np.abs(d[t + "JJ"])) < 1e-12
    assert np.max(np.abs(h["descent"] - d[t + "descent"]) / np.abs(d[t + "descent"])) < 1e-9
    assert np.array_equal(h["xx_star"], d[t + "xx_star"]) and relerr(d[t + "uu_star"], h["uu_star"]) < 1e-9
