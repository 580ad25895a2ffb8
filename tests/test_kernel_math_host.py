"""Checks the per-instance code of the CUDA kernels (csrc/acoc_kernels.cuh), replayed on the host by
tests/host_emul, against the golden fixtures of the live reference and against the oracle.  CPU only: this is
what lets kernel arithmetic be validated in a container without a GPU; the same assertions run against the real
GPU in test_gpu_parity.py."""
import numpy as np
import pytest

from tests.util import golden, relerr


def test_step_sample(emul):
    d = golden("step_kat.npz")
    xxp32, A, B, fxx, fux = emul.step_batch(d["x"], d["u"], None, state_f64=False)
    xxp64, _, _, fxxc, fuxc = emul.step_batch(d["x"], d["u"], d["lam"], state_f64=True)
    # host replay uses glibc sin/cos like numpy; the square is V*V instead of pow(V,2): <= 1 ulp in D, L, then
    # scaled by dt/m -- the float32-rounded next state is identical, the float64 one to 1e-15
    assert np.array_equal(xxp32, d["xxp32"])
    assert relerr(d["xxp64"], xxp64) < 1e-15
    assert relerr(d["fx"], np.swapaxes(A, 1, 2)) < 1e-12 and relerr(d["fu"], np.swapaxes(B, 1, 2)) < 1e-12
    assert relerr(d["fxx"], fxx) < 1e-12 and relerr(d["fux"], fux) < 1e-12
    assert relerr(d["fxxc"], fxxc) < 1e-12 and relerr(d["fuxc"], fuxc) < 1e-12


def test_cost_sample(emul):
    d = golden("cost_kat.npz")
    for w in range(len(d["Q"])):
        m = d["which"] == w
        ll, lx, lu, llT, lTx = emul.cost_batch(d["Q"][w], d["R"][w], d["QT"][w], d["x"][m], d["u"][m], d["xr"][m], d["ur"][m])
        assert np.max(np.abs(ll - d["ll"][m]) / np.abs(d["ll"][m])) < 1e-13
        assert np.max(np.abs(llT - d["llT"][m]) / np.abs(d["llT"][m])) < 1e-13
        assert relerr(d["lx"][m], lx) < 1e-14 and relerr(d["lu"][m], lu) < 1e-14 and relerr(d["lTx"][m], lTx) < 1e-14


def test_dense_lq_forced_regularisation(emul):
    d = golden("lq_forced_reg.npz")
    tm = lambda M: np.ascontiguousarray(np.moveaxis(M, 2, 0))
    K, P, x, u, n = emul.ltv_lqr(tm(d["A"]), tm(d["B"]), tm(d["Q"]), tm(d["R"]), tm(d["S"]), d["Qf"], np.zeros(6),
                                 np.ascontiguousarray(d["q"].T), np.ascontiguousarray(d["r"].T), d["qf"])
    assert n == int(d["n_reg_aug"])
    assert relerr(d["K_aug"], np.moveaxis(K, 0, 2)) < 1e-10 and relerr(d["P_aug"], np.moveaxis(P, 0, 2)) < 1e-10
    assert relerr(d["x_aug"], x.T) < 1e-10 and relerr(d["u_aug"], u.T) < 1e-10
    K, P, x, u, n = emul.ltv_lqr(tm(d["A"]), tm(d["B"]), tm(d["Q"]), tm(d["R"]), tm(d["S"]), d["Qf"], d["x0"])
    assert n == int(d["n_reg_non"])
    assert relerr(d["K_non"], np.moveaxis(K, 0, 2)) < 1e-10 and relerr(d["x_non"], x.T) < 1e-10 and relerr(d["u_non"], u.T) < 1e-10


@pytest.mark.parametrize("name", ["newton_step_f32", "newton_step_f64", "newton_acro_f32", "newton_acro_f64"])
@pytest.mark.parametrize("lazy", [False, True])
def test_newton_kernels_reproduce_reference(emul, name, lazy):
    """The fused sparse backward sweep + LQ forward + candidate rollouts + select + update, driven like
    acoc_newton_iterate, reproduce the live reference: iteration count, every Armijo step, history, result."""
    d = golden(name + ".npz")
    f64 = name.endswith("f64")
    h = emul.newton_batch(d["xx_ref"], d["uu_ref"], d["xx_init"][None], d["uu_init"][None], d["Q"], d["R"], d["QT"], state_f64=f64, lazy=lazy)
    k = int(d["iters"])
    assert h["iters"][0] == k and h["status"][0] == 1
    assert np.array_equal(h["stepsize"][0, :k], d["stepsize"])
    assert np.array_equal(h["n_armijo"][0, :k], d["n_armijo"])
    assert np.max(np.abs(h["JJ"][0, :k] - d["JJ"]) / np.abs(d["JJ"])) < 1e-12
    assert np.max(np.abs(h["descent"][0, :k] - d["descent"]) / np.abs(d["descent"])) < 1e-9
    assert relerr(d["xx_star"], h["xx_star"][0]) < 1e-9 and relerr(d["uu_star"], h["uu_star"][0]) < 1e-9
    assert relerr(d["xx_last"], h["xx_last"][0]) < 1e-9 and relerr(d["uu_last"], h["uu_last"][0]) < 1e-9
    if not f64:
        assert np.array_equal(h["xx_star"][0], d["xx_star"])
    assert h["n_reg"][0] == 0  # the +0.5 I branch never fires on the shipped configs


def test_backward_forward_match_reference_lq(emul):
    """K, sigma, deltau of the fused sweeps vs the reference's ltv_LQR captured inside live Newton iterations
    (Gauss-Newton kk=0,5 and exact-Hessian kk=9,15 phases)."""
    d = golden("newton_acro_f32.npz")
    for kk in d["lq_at"]:
        kk = int(kk)
        # one iteration from the stored iterate with the Hessian mode of iteration kk
        h = emul.newton_batch(d["xx_ref"], d["uu_ref"], d["it%d_xx" % kk][None], d["it%d_uu" % kk][None], d["Q"], d["R"], d["QT"],
                              state_f64=False, n_iters_cap=1, exact_after=(-1 if kk > 8 else 8))
        KK = d["it%d_KK" % kk]  # (2,7,TT): column 0 = sigma, 1: = K
        assert relerr(KK[:, 1:, :], h["K"][0]) < 1e-9
        assert relerr(KK[:, 0, :], h["sigma"][0]) < 1e-9
        assert relerr(d["it%d_deltau" % kk], h["deltau"][0]) < 1e-9


def test_batch_matches_oracle_ragged(emul, oracle):
    """33 instances (not a multiple of the warp size), per-instance references and initial states, 6 iterations
    crossing the Gauss-Newton -> exact-Hessian switch at kk = 2: emulated kernels vs oracle, instance by instance."""
    from aircraftoptimalcontrol_b200 import refgen
    rng = np.random.default_rng(3)
    n, TT = 33, 200
    zf, xf = rng.uniform(1.5, 3.5, n), rng.uniform(14, 18, n)
    xr, ur = refgen.step_problem(xf, zf, tf=0.2, TT=TT)
    Q, R, QT = refgen.weights("step")
    xi, ui = emul.init_guess(xr)
    for i in (0, 7, 32):
        xo, uo = oracle.initial_trajectory(xr[i])
        assert np.array_equal(xo, xi[i]) and np.array_equal(uo, ui[i])
    h = emul.newton_batch(xr, ur, xi, ui, Q, R, QT, n_iters_cap=6, exact_after=1)
    o = oracle.newton_batch(xr, ur, xi, ui, Q, R, QT, n_iters_cap=6, exact_after=1)
    assert np.array_equal(h["iters"], o["iters"])
    for i in range(n):
        k = o["iters"][i]
        assert np.array_equal(h["stepsize"][i, :k], o["stepsize"][i, :k])
        assert np.max(np.abs(h["JJ"][i, :k] - o["JJ"][i, :k]) / np.abs(o["JJ"][i, :k])) < 1e-12
        assert np.max(np.abs(h["descent"][i, :k] - o["descent"][i, :k]) / np.abs(o["descent"][i, :k])) < 1e-9
    assert relerr(o["xx_star"], h["xx_star"]) < 1e-9 and relerr(o["uu_star"], h["uu_star"]) < 1e-9


def test_tracking(emul):
    d = golden("lqr_tracking.npz")
    xr, ur, K = emul.lqr_tracking(d["xx_opt"], d["uu_opt"], d["Q"], d["R"], d["QT"], d["delta"])
    assert relerr(d["KK"], K) < 1e-11
    assert np.array_equal(xr, d["xx_reg"])
    assert relerr(d["uu_reg"], ur) < 1e-11


def test_dense_weights_path(emul, oracle):
    """Dense (non-diagonal, symmetric) Q/R/QT exercise the general cost path of the kernels."""
    d = golden("newton_step_f32.npz")
    rng = np.random.default_rng(11)
    TT = 120
    sl = slice(0, TT)
    xr, ur, xi, ui = d["xx_ref"][:, sl], d["uu_ref"][:, sl], d["xx_init"][:, sl], d["uu_init"][:, sl]
    E = rng.normal(size=(6, 6)) * 1e-4
    Q = d["Q"] + E @ E.T
    QT = d["QT"] + 3 * (E @ E.T)
    R = d["R"] + 1e-7 * np.array([[1.0, 0.3], [0.3, 2.0]])
    h = emul.newton_batch(xr, ur, xi[None], ui[None], Q, R, QT, n_iters_cap=4)
    o = oracle.newton(xr, ur, xi, ui, Q, R, QT, n_iters_cap=4)
    assert np.array_equal(h["stepsize"][0, :4], o["stepsize"])
    assert np.max(np.abs(h["JJ"][0, :4] - o["JJ"]) / np.abs(o["JJ"])) < 1e-12
    assert relerr(o["xx_last"], h["xx_last"][0]) < 1e-9 and relerr(o["uu_last"], h["uu_last"][0]) < 1e-9


def test_return_slot_quirks(emul):
    d = golden("newton_quirks.npz")
    a = emul.newton_batch(d["xx_ref"], d["uu_ref"], d["a_xx_init"][None], d["a_uu_init"][None], d["Q"], d["R"], d["QT"], state_f64=True, max_iters=4)
    assert a["iters"][0] == 3 and a["status"][0] == 2 and np.array_equal(a["stepsize"][0, :3], d["a_stepsize"])
    assert relerr(d["a_xx_star"], a["xx_star"][0]) < 1e-9 and relerr(d["a_uu_star"], a["uu_star"][0]) < 1e-9
    b = emul.newton_batch(d["xx_ref"], d["uu_ref"], d["b_xx_init"][None], d["b_uu_init"][None], d["Q"], d["R"], d["QT"], state_f64=True)
    assert b["iters"][0] == 1 and b["status"][0] == 1 and not b["xx_star"][0].any() and not b["uu_star"][0].any()


def test_float_state_slots_lossless_on_host(emul):
    """Host replay of the <F = double, XT = float> instantiation (float state slots, exact x0 kept aside) against the
    <double, double> one: bit-identical histories and results on config 1 (the GPU test repeats this at scale)."""
    d = golden("newton_step_f32.npz")
    xi = d["xx_init"].copy()
    xi[:, 0] += np.array([1e-9, 2e-9, 3e-9, 1e-10, 0.0, 7e-11])  # an x0 that is not a float32 value; later states stay float32
    a, b = (emul.newton_batch(d["xx_ref"], d["uu_ref"], xi[None], d["uu_init"][None], d["Q"], d["R"], d["QT"], lazy=True, mode=m, n_iters_cap=12)
            for m in (0, 1))
    for k in ("JJ", "descent", "stepsize", "n_armijo", "xx_star", "uu_star", "xx_last", "uu_last", "deltau", "K"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["xx_last"][0][:, 0], xi[:, 0])


def test_fp32_mode_on_host(emul):
    """Host replay of the FP32 instantiation on config 1: same convergence behaviour as the reference with its float32 state,
    same optimum within the tolerance stated for the mode (tests/test_gpu_parity.py FP32_TOL)."""
    d = golden("newton_step_f32.npz")
    h = emul.newton_batch(d["xx_ref"], d["uu_ref"], d["xx_init"][None], d["uu_init"][None], d["Q"], d["R"], d["QT"], lazy=True, mode=2)
    k = int(h["iters"][0])
    assert h["status"][0] == 1 and 16 <= k <= 35
    assert abs(h["JJ"][0, k - 1] - d["JJ"][-1]) / d["JJ"][-1] < 2e-6
    assert np.max(np.abs(h["xx_last"][0] - d["xx_last"])) < 2e-3
    assert np.max(np.abs(h["uu_last"][0] - d["uu_last"])) < 2e-4 * np.max(np.abs(d["uu_last"]))


@pytest.mark.parametrize("name", ["step_f32", "step_f64", "acro_f32"])
@pytest.mark.parametrize("lazy", [False, True])
def test_gradient_kernels_reproduce_reference(emul, name, lazy):
    """GradientMethod.optimize (optcon.py:27-174, line-search call repaired as include/acoc.h describes): the costate sweep of
    k_gradient_tma + candidate rollouts + select + update, driven like acoc_newton_iterate, reproduce the live reference run through
    oracle/pyref.py::run_gradient -- every Armijo step and candidate count, cost/descent history, iterates."""
    g = golden("gradient_%s.npz" % name)
    d = golden(str(g["base"]))
    h = emul.newton_batch(d["xx_ref"], d["uu_ref"], d["xx_init"][None], d["uu_init"][None], d["Q"], d["R"], d["QT"], state_f64=name.endswith("f64"),
                          lazy=lazy, method=1, max_iters=int(g["max_iters"]), stepsize_0=float(g["stepsize_0"]), armijo_maxiters=int(g["armijo_maxiters"]))
    k = int(g["iters"])
    assert h["iters"][0] == k and h["status"][0] == 2  # ran max_iters-1 bodies (optcon.py:85), like the reference did
    assert np.array_equal(h["stepsize"][0, :k], g["stepsize"])
    assert np.array_equal(h["n_armijo"][0, :k], g["n_armijo"])
    assert np.max(np.abs(h["JJ"][0, :k] - g["JJ"]) / np.abs(g["JJ"])) < 1e-12
    assert np.max(np.abs(-h["descent"][0, :k] - g["descent"]) / np.abs(g["descent"])) < 1e-9
    assert relerr(g["xx_last"], h["xx_last"][0]) < 1e-9 and relerr(g["uu_last"], h["uu_last"][0]) < 1e-9
    assert relerr(g["xx_star"], h["xx_star"][0]) < 1e-9 and relerr(g["uu_star"], h["uu_star"][0]) < 1e-9


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("weights", ["diag", "dense", "tiny_R"])
def test_riccati_by_columns_bit_identical(emul, exact, weights):
    """k_backward_cols runs the matrix half of the Riccati step as one warp per column; the column pieces must reproduce
    riccati_matrix() bit for bit (K, sigma, P, p of every step of a real backward sweep), +0.5 I branch included."""
    d = golden("newton_acro_f32.npz")
    Q, R, QT = d["Q"].copy(), d["R"].copy(), d["QT"].copy()
    if weights == "dense":
        E = np.random.default_rng(5).normal(size=(6, 6)) * 1e-4
        Q, QT, R = Q + E @ E.T, QT + 3 * (E @ E.T), R + 1e-7 * np.array([[1.0, 0.3], [0.3, 2.0]])
    if weights == "tiny_R":   # G = R + B'PB loses positive definiteness somewhere along the sweep -> the regularised gain is exercised
        R = -np.abs(R) * 50.0
    for xx, uu in ((d["xx_init"], d["uu_init"]), (d["xx_star"], d["uu_star"])):
        bad, nreg = emul.riccati_cols_check(xx, uu, d["xx_ref"], d["uu_ref"], Q, R, QT, exact=exact)
        assert bad == 0
        if weights == "tiny_R":
            assert nreg > 0


@pytest.mark.parametrize("lazy", [False, True])
def test_kernels_reproduce_reference_on_batched_config_instances(emul, lazy):
    """The kernels' per-instance code (host replay, driven like acoc_newton_iterate) on the four instances of BASELINE configs[3] / [4]
    that the live Python reference solved (tests/golden/newton_batched_instances.npz): the two step-maneuver instances in ONE batch
    with per-instance references, the two acrobatic ones with their perturbed x0 -- iteration counts, every Armijo step and candidate
    count identical, histories 1e-12 / 1e-9, float32-quantised results bit-identical."""
    from aircraftoptimalcontrol_b200 import refgen
    d = golden("newton_batched_instances.npz")
    for cfg in ("config4", "config5"):
        if cfg == "config4":
            zf, xf = refgen.config4_params()
            xr, ur = refgen.step_problem(xf[:2], zf[:2])
            Q, R, QT = refgen.weights("step")
        else:
            _, zf = refgen.config5_params()
            xr, ur = refgen.acrobatic_problem(zf[:2])
            Q, R, QT = refgen.weights("acro")
        xi = np.stack([d["%s_%d_xx_init" % (cfg, i)] for i in range(2)])
        ui = np.stack([d["%s_%d_uu_init" % (cfg, i)] for i in range(2)])
        h = emul.newton_batch(xr, ur, xi, ui, Q, R, QT, lazy=lazy)
        for i in range(2):
            t = "%s_%d_" % (cfg, i)
            k = int(d[t + "iters"])
            assert h["iters"][i] == k and h["status"][i] == 1
            assert np.array_equal(h["stepsize"][i, :k], d[t + "stepsize"]) and np.array_equal(h["n_armijo"][i, :k], d[t + "n_armijo"])
            assert np.max(np.abs(h["JJ"][i, :k] - d[t + "JJ"]) / np.abs(d[t + "JJ"])) < 1e-12
            assert np.max(np.abs(h["descent"][i, :k] - d[t + "descent"]) / np.abs(d[t + "descent"])) < 1e-9
            assert np.array_equal(h["xx_star"][i], d[t + "xx_star"]) and relerr(d[t + "uu_star"], h["uu_star"][i]) < 1e-9
