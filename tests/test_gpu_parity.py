"""Parity of the CUDA path (through the C ABI, libacoc.so) with the live reference's golden fixtures and with
the CPU oracle.  Runs on the B200 box: python -m pytest tests -m gpu.

Tolerances: next-state / rollout values are compared bit for bit (the float32 rounding of
aircraft_simplified.py:300 absorbs the <= 2 ulp difference between CUDA's and glibc's sin/cos); everything that
goes through the Riccati recursion is compared at the north-star tolerance 1e-9 relative; Armijo steps and
iteration counts must be IDENTICAL."""
import numpy as np
import pytest

from tests.util import golden, relerr

pytestmark = pytest.mark.gpu


def test_step_batch_kat(gpu):
    from aircraftoptimalcontrol_b200.aircraft_simplified import Dynamics
    d = golden("step_kat.npz")
    r32 = Dynamics(state="f32").step_batch(d["x"], d["u"])
    r64 = Dynamics(state="f64").step_batch(d["x"], d["u"], d["lam"])
    assert np.array_equal(r32["xxp"], d["xxp32"])
    assert relerr(d["xxp64"], r64["xxp"]) < 1e-15
    assert relerr(d["fx"], np.swapaxes(r32["A"], 1, 2)) < 1e-12 and relerr(d["fu"], np.swapaxes(r32["B"], 1, 2)) < 1e-12
    assert relerr(d["fxx"], r32["fxx"]) < 1e-12 and relerr(d["fux"], r32["fux"]) < 1e-12
    assert relerr(d["fxxc"], r64["fxx"]) < 1e-12 and relerr(d["fuxc"], r64["fux"]) < 1e-12


def test_cost_batch_kat(gpu):
    from aircraftoptimalcontrol_b200.aircraft_simplified import Cost
    d = golden("cost_kat.npz")
    for w in range(len(d["Q"])):
        m = d["which"] == w
        c = Cost(d["Q"][w], d["R"][w], d["QT"][w])
        ll, lx, lu = c.stagecost_batch(d["x"][m], d["u"][m], d["xr"][m], d["ur"][m])
        llT, lTx = c.termcost_batch(d["x"][m], d["xr"][m])
        assert np.max(np.abs(ll - d["ll"][m]) / np.abs(d["ll"][m])) < 1e-13
        assert np.max(np.abs(llT - d["llT"][m]) / np.abs(d["llT"][m])) < 1e-13
        assert relerr(d["lx"][m], lx) < 1e-14 and relerr(d["lu"][m], lu) < 1e-14 and relerr(d["lTx"][m], lTx) < 1e-14


def test_step_and_cost_10k_random_samples(gpu, oracle):
    """SURVEY 8(c) parity protocol: Dynamics.step / Cost on 10^4 random (x, u, lambda) with V in [5, 30] against the C oracle (itself
    pinned to the live reference on the 256-sample KAT): float32-rounded next state bit-exact, float64 next state 1e-15, Jacobians and
    costate-contracted Hessians 1e-12, costs and gradients 1e-13."""
    from aircraftoptimalcontrol_b200.aircraft_simplified import Cost, Dynamics
    rng = np.random.default_rng(424242)
    n = 10000
    X = np.stack([rng.uniform(-50, 50, n), rng.uniform(-20, 20, n), rng.uniform(5, 30, n), rng.uniform(-1.2, 1.2, n), rng.uniform(-3, 3, n),
                  rng.uniform(-1.2, 1.2, n)], axis=1)
    U = np.stack([rng.uniform(-200, 600, n), rng.uniform(-150, 150, n)], axis=1)
    LAM = rng.normal(size=(n, 6)) * np.array([1, 10, 0.1, 1, 0.01, 1])
    r32 = Dynamics(state="f32").step_batch(X, U)
    r64 = Dynamics(state="f64").step_batch(X, U, LAM)
    bad32 = 0
    e64 = eA = eB = eH = eS = 0.0
    for k in range(n):
        o32 = oracle.step(X[k], U[k], quant_f32=True)
        o64 = oracle.step(X[k], U[k], LAM[k], quant_f32=False)
        bad32 += not np.array_equal(o32[0], r32["xxp"][k])
        e64 = max(e64, relerr(o64[0], r64["xxp"][k]))
        eA, eB = max(eA, relerr(o32[1], r32["A"][k].T)), max(eB, relerr(o32[2], r32["B"][k].T))
        eH, eS = max(eH, relerr(o64[3], r64["fxx"][k])), max(eS, relerr(o64[5], r64["fux"][k]))
    assert bad32 == 0 and e64 < 1e-15 and eA < 1e-12 and eB < 1e-12 and eH < 1e-12 and eS < 1e-12, (bad32, e64, eA, eB, eH, eS)
    Q, R, QT = (rng.normal(size=(6, 6)), rng.normal(size=(2, 2)), rng.normal(size=(6, 6)))
    Q, R, QT = Q @ Q.T, R @ R.T, QT @ QT.T
    XR, UR = X + rng.normal(size=(n, 6)), U + rng.normal(size=(n, 2)) * 10
    c = Cost(Q, R, QT)
    ll, lx, lu = c.stagecost_batch(X, U, XR, UR)
    llT, lTx = c.termcost_batch(X, XR)
    for k in range(0, n, 7):
        ol, olx, olu = oracle.stagecost(Q, R, X[k], U[k], XR[k], UR[k])
        oT, oTx = oracle.termcost(QT, X[k], XR[k])
        assert abs(ll[k] - ol) < 1e-13 * abs(ol) and abs(llT[k] - oT) < 1e-13 * abs(oT), k
        assert relerr(olx, lx[k]) < 1e-13 and relerr(olu, lu[k]) < 1e-13 and relerr(oTx, lTx[k]) < 1e-13, k


def test_ltv_lqr_forced_regularisation(gpu):
    from aircraftoptimalcontrol_b200.optcon import ltv_LQR
    d = golden("lq_forced_reg.npz")
    TT = d["A"].shape[2]
    K, P, x, u, n = ltv_LQR(d["A"], d["B"], d["Q"], d["R"], d["S"], d["Qf"], TT, np.zeros(6), d["q"], d["r"], d["qf"], return_nreg=True)
    assert n == int(d["n_reg_aug"]) > 0 and K.shape == (2, 7, TT) and P.shape == (7, 7, TT)
    for got, ref in ((K, d["K_aug"]), (P, d["P_aug"]), (x, d["x_aug"]), (u, d["u_aug"])):
        assert relerr(ref, got) < 1e-9
    K, P, x, u, n = ltv_LQR(d["A"], d["B"], d["Q"], d["R"], d["S"], d["Qf"], TT, d["x0"], return_nreg=True)
    assert n == int(d["n_reg_non"]) > 0 and K.shape == (2, 6, TT)
    for got, ref in ((K, d["K_non"]), (P, d["P_non"]), (x, d["x_non"]), (u, d["u_non"])):
        assert relerr(ref, got) < 1e-9
    # constant (6,6)/(2,2) weights are broadcast along t like optcon.py:603-606
    K2 = ltv_LQR(d["A"], d["B"], d["Q"][:, :, 0], np.eye(2), d["S"], d["Qf"], TT, d["x0"])[0]
    K3 = ltv_LQR(d["A"], d["B"], np.repeat(d["Q"][:, :, :1], TT, 2), np.repeat(np.eye(2)[:, :, None], TT, 2), d["S"], d["Qf"], TT, d["x0"])[0]
    assert np.array_equal(K2, K3)
    with pytest.raises(ValueError):
        ltv_LQR(d["A"], d["B"], np.eye(5), d["R"], d["S"], d["Qf"], TT, d["x0"])


def test_ltv_lqr_batch_equals_single_calls(gpu, oracle):
    """acoc_ltv_lqr with nb > 1 (optcon.ltv_LQR_batch): several independent problems in one launch, each equal to its own ltv_LQR call
    bit for bit and to the oracle's ltv_LQR (optcon.py:533-771) to 1e-9 -- augmented and plain branches, one problem that takes the
    +0.5*I branch and two that do not."""
    from aircraftoptimalcontrol_b200.optcon import ltv_LQR, ltv_LQR_batch
    d = golden("lq_forced_reg.npz")
    TT = d["A"].shape[2]
    rng = np.random.default_rng(5)
    Qpd = d["Q"] + 5.0 * np.eye(6)[:, :, None]   # positive definite: no regularised step
    base = [(d["A"], d["B"], d["Q"], d["R"], d["S"], d["Qf"], d["x0"]),
            (d["A"], d["B"], Qpd, d["R"] + np.eye(2)[:, :, None], 0 * d["S"], np.eye(6), rng.normal(size=6)),
            (d["A"] * 0.9, d["B"], Qpd[:, :, 0], np.eye(2), 0 * d["S"], 2 * np.eye(6), rng.normal(size=6))]
    for aug in (False, True):
        probs = [p + ((d["q"] * (k + 1), d["r"], d["qf"] * (1 - k)) if aug else ()) for k, p in enumerate(base)]
        K, P, x, u, nreg = ltv_LQR_batch(probs, TT)
        n = 7 if aug else 6
        assert K.shape == (3, 2, n, TT) and P.shape == (3, n, n, TT) and x.shape == (3, 6, TT) and u.shape == (3, 2, TT)
        assert nreg[0] > 0 and nreg[1] == 0 and nreg[2] == 0
        for b, p in enumerate(probs):
            args = p[:6] + (TT, p[6]) + tuple(p[7:])
            K1, P1, x1, u1, n1 = ltv_LQR(*args, return_nreg=True)
            assert np.array_equal(K[b], K1) and np.array_equal(P[b], P1) and np.array_equal(x[b], x1) and np.array_equal(u[b], u1) and nreg[b] == n1
            Q3 = p[2] if np.ndim(p[2]) == 3 else np.repeat(p[2][:, :, None], TT, 2)
            R3 = p[3] if np.ndim(p[3]) == 3 else np.repeat(p[3][:, :, None], TT, 2)
            Ko, Po, xo, uo = oracle.ltv_lqr(p[0], p[1], Q3, R3, p[4], p[5], TT, p[6], *p[7:])[:4]
            for got, ref in ((K[b], Ko), (P[b], Po), (x[b], xo), (u[b], uo)):
                assert relerr(ref, got) < 1e-9
    with pytest.raises(ValueError):
        ltv_LQR_batch([base[0], base[1] + (d["q"], d["r"], d["qf"])], TT)   # mixed branches


@pytest.mark.parametrize("name", ["newton_step_f32", "newton_step_f64", "newton_acro_f32", "newton_acro_f64"])
@pytest.mark.parametrize("armijo", ["speculative", "lazy"])
def test_newton_configs_1_2(gpu, name, armijo):
    """BASELINE.json configs[0] and configs[1] (single trajectory, GPU vs the reference's numpy float64):
    identical iteration count and Armijo step at every iteration, histories and trajectories within 1e-9."""
    d = golden(name + ".npz")
    f64 = name.endswith("f64")
    TT = d["xx_ref"].shape[1]
    with gpu.BatchedNewton(1, TT=TT, state="f64" if f64 else "f32", refs_shared=True, armijo=armijo) as bn:
        bn.set_weights(d["Q"], d["R"], d["QT"])
        bn.set_refs(d["xx_ref"], d["uu_ref"])
        bn.set_init(d["xx_init"][None], d["uu_init"][None])
        total = bn.solve()
        xs, us = bn.result()
        xl, ul = bn.iterate_at(0)
        h, st = bn.history(), bn.stats()
    k = int(d["iters"])
    assert total == k and st["iters"][0] == k and st["status"][0] == 1 and st["n_reg"][0] == 0
    assert np.array_equal(h["stepsize"][0, :k], d["stepsize"])
    assert np.array_equal(h["n_armijo"][0, :k], d["n_armijo"])
    assert np.max(np.abs(h["JJ"][0, :k] - d["JJ"]) / np.abs(d["JJ"])) < 1e-9
    assert np.max(np.abs(h["descent"][0, :k] - d["descent"]) / np.abs(d["descent"])) < 1e-9
    assert relerr(d["xx_star"], xs[0]) < 1e-9 and relerr(d["uu_star"], us[0]) < 1e-9
    assert relerr(d["xx_last"], xl[0]) < 1e-9 and relerr(d["uu_last"], ul[0]) < 1e-9


def test_fused_sweeps_match_reference_lq(gpu):
    """K, sigma, deltau of the fused backward/forward kernels vs the reference's ltv_LQR inside live iterations."""
    d = golden("newton_acro_f32.npz")
    TT = d["xx_ref"].shape[1]
    for kk in d["lq_at"]:
        kk = int(kk)
        with gpu.BatchedNewton(1, TT=TT, refs_shared=True) as bn:
            bn.set_weights(d["Q"], d["R"], d["QT"])
            bn.set_refs(d["xx_ref"], d["uu_ref"])
            bn.set_init(d["it%d_xx" % kk][None], d["it%d_uu" % kk][None])
            bn.backward(exact=kk > 8)
            bn.forward()
            K, sig = bn.gains()
            du = bn.deltau()
        KK = d["it%d_KK" % kk]
        assert relerr(KK[:, 1:, :], K[0]) < 1e-9 and relerr(KK[:, 0, :], sig[0]) < 1e-9
        assert relerr(d["it%d_deltau" % kk], du[0]) < 1e-9


def _random_batch(n, TT, seed, tf):
    from aircraftoptimalcontrol_b200 import refgen
    rng = np.random.default_rng(seed)
    zf, xf = rng.uniform(1.5, 3.5, n), rng.uniform(14, 18, n)
    xr, ur = refgen.step_problem(xf, zf, tf=tf, TT=TT)
    return (xr, ur) + refgen.weights("step")


def test_rollouts_bit_exact_at_scale(gpu, oracle):
    """4096 instances x 999 steps of get_update (optcon.py:176-200) with random inputs and steps: the float32-quantised
    state trajectories must be bit-identical to the oracle's (glibc trig) for every instance."""
    n, TT = 4096, 1000
    rng = np.random.default_rng(5)
    xr, ur, Q, R, QT = _random_batch(n, TT, 5, 1.0)
    with gpu.BatchedNewton(n, TT=TT) as bn:
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess()
        xi, ui = bn.iterate_at(0)
        du = rng.normal(size=(n, 2, TT)) * np.array([30.0, 5.0])[None, :, None]
        s = rng.uniform(0.02, 1.0, n)
        bn.set_deltau(du)
        bn.update(s)
        xn, un = bn.iterate_at(0)
        Jn = bn.stats()["J"]
    bad_init = bad_roll = 0
    for i in range(0, n, 16):  # 256 sampled instances against the scalar CPU oracle
        xo, uo = oracle.initial_trajectory(xr[i])
        bad_init += not (np.array_equal(xo, xi[i]) and np.array_equal(uo, ui[i]))
        xo, uo, Jo = oracle.rollout(xi[i, :, 0], ui[i], du[i], s[i], cost=(Q, R, QT), xr=xr[i], ur=ur[i])
        bad_roll += not (np.array_equal(xo, xn[i]) and np.array_equal(uo, un[i]))
        assert abs(Jo - Jn[i]) <= 1e-12 * abs(Jo)
    assert bad_init == 0 and bad_roll == 0


def test_batch_solve_matches_oracle(gpu, oracle):
    """Config-4 style batch (randomised step references, device-side initial guess), 96 instances to convergence:
    per-instance iteration counts, Armijo steps and results vs the CPU oracle."""
    n, TT = 96, 1000
    xr, ur, Q, R, QT = _random_batch(n, TT, 2024, 1.0)
    with gpu.BatchedNewton(n, TT=TT, armijo="lazy") as bn:
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess()
        xi, ui = bn.iterate_at(0)
        total = bn.solve()
        xs, us = bn.result()
        h, st = bn.history(), bn.stats()
    o = oracle.newton_batch(xr, ur, xi, ui, Q, R, QT)
    assert np.array_equal(st["iters"], o["iters"]) and total == int(o["iters"].sum())
    assert np.all(st["status"] == 1)
    for i in range(n):
        k = o["iters"][i]
        assert np.array_equal(h["stepsize"][i, :k], o["stepsize"][i, :k]), i
        assert np.max(np.abs(h["descent"][i, :k] - o["descent"][i, :k]) / np.abs(o["descent"][i, :k])) < 1e-9
    assert relerr(o["xx_star"], xs) < 1e-9 and relerr(o["uu_star"], us) < 1e-9


@pytest.mark.parametrize("n", [1, 33, 100])
def test_ragged_and_modes_agree(gpu, n):
    """Instance counts that are not multiples of the warp/CTA size; speculative and lazy Armijo, shared and
    per-instance reference storage must give bit-identical results."""
    d = golden("newton_step_f32.npz")
    TT = 1000
    rng = np.random.default_rng(n)
    xi = np.repeat(d["xx_init"][None], n, 0)
    ui = np.repeat(d["uu_init"][None], n, 0) + rng.normal(size=(n, 2, TT)) * 0.5
    ui[0] = d["uu_init"]
    res = []
    for armijo, shared in (("speculative", True), ("lazy", True), ("speculative", False)):
        with gpu.BatchedNewton(n, TT=TT, refs_shared=shared, armijo=armijo) as bn:
            bn.set_weights(d["Q"], d["R"], d["QT"])
            if shared:
                bn.set_refs(d["xx_ref"], d["uu_ref"])
            else:
                bn.set_refs(np.repeat(d["xx_ref"][None], n, 0), np.repeat(d["uu_ref"][None], n, 0))
            bn.set_init(xi, ui)
            bn.iterate(12)
            res.append((bn.iterate_at(0), bn.history(), bn.stats()))
    (x0, u0), h0, s0 = res[0]
    for (x1, u1), h1, s1 in res[1:]:
        assert np.array_equal(x0, x1) and np.array_equal(u0, u1)
        assert np.array_equal(h0["stepsize"], h1["stepsize"]) and np.array_equal(h0["JJ"], h1["JJ"])
        assert np.array_equal(s0["iters"], s1["iters"])
    k = 12
    assert np.array_equal(h0["stepsize"][0, :k], d["stepsize"][:k])  # instance 0 is config 1 itself


def test_lqr_tracking_config3(gpu, oracle):
    """BASELINE.json configs[2]: LQR tracking of Data/xx_star.npy from 4096 perturbed initial states."""
    from aircraftoptimalcontrol_b200 import refgen
    from aircraftoptimalcontrol_b200.lqr_tracking import lqr_tracking, lqr_tracking_batch
    d = golden("lqr_tracking.npz")
    xr, ur, K = lqr_tracking_batch(d["xx_opt"], d["uu_opt"], d["delta"], return_gains=True)
    assert relerr(d["KK"], K) < 1e-9
    assert np.array_equal(xr, d["xx_reg"]) and relerr(d["uu_reg"], ur) < 1e-9
    x1, u1 = lqr_tracking(d["xx_opt"], d["uu_opt"], np.linspace(0, 1, 1000))
    assert np.array_equal(x1, d["xx_reg"][0])
    deltas = refgen.config3_deltas(4096)
    xg, ug = lqr_tracking_batch(d["xx_opt"], d["uu_opt"], deltas)
    xo, uo, _ = oracle.lqr_tracking(d["xx_opt"], d["uu_opt"], d["Q"], d["R"], d["QT"], deltas)
    assert np.mean(np.all(xg == xo, axis=(1, 2))) == 1.0
    assert relerr(uo, ug) < 1e-9


def test_full_size_determinism(gpu):
    """Size-independent property at BASELINE scale: 16384 copies of config 1 solved in one batch all reproduce the
    single-instance golden history exactly (no cross-instance interference, no dependence on the lane/CTA)."""
    d = golden("newton_step_f32.npz")
    n, TT = 16384, 1000
    with gpu.BatchedNewton(n, TT=TT, refs_shared=True, armijo="lazy") as bn:
        bn.set_weights(d["Q"], d["R"], d["QT"])
        bn.set_refs(d["xx_ref"], d["uu_ref"])
        bn.set_init(np.ascontiguousarray(np.broadcast_to(d["xx_init"], (n, 6, TT))), np.ascontiguousarray(np.broadcast_to(d["uu_init"], (n, 2, TT))))
        total = bn.solve()
        h, st = bn.history(), bn.stats()
        xs, _ = bn.result()
    k = int(d["iters"])
    assert total == n * k and np.all(st["iters"] == k)
    assert np.all(h["stepsize"][:, :k] == d["stepsize"][None]) and np.all(h["JJ"] == h["JJ"][0])
    assert np.all(xs == xs[0]) and np.array_equal(xs[0], d["xx_star"])


def test_error_paths(gpu):
    from aircraftoptimalcontrol_b200 import _lib
    with pytest.raises(_lib.AcocError):
        gpu.BatchedNewton(0)
    with gpu.BatchedNewton(2, TT=50) as bn:
        with pytest.raises(_lib.AcocError, match="not ready"):
            bn.iterate(1)
        with pytest.raises(ValueError):
            bn.set_refs(np.zeros((2, 6, 49)), np.zeros((2, 2, 49)))
        with pytest.raises(_lib.AcocError, match="symmetric"):
            bn.set_weights(np.triu(np.ones((6, 6))), np.eye(2), np.eye(6))


def test_pipelined_chunks_equal_single_batch(gpu):
    """PipelinedNewton (independent sub-batches on their own streams/threads) returns exactly what one big batch does."""
    n, TT = 200, 300
    xr, ur, Q, R, QT = _random_batch(n, TT, 77, 0.3)
    with gpu.BatchedNewton(n, TT=TT, armijo="lazy") as bn:
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess()
        bn.solve()
        xs, us = bn.result()
        st = bn.stats()
    with gpu.PipelinedNewton(n, n_chunks=3, TT=TT, armijo="lazy") as pn:
        pn.set_weights(Q, R, QT)
        xp, up, sp = pn.solve(xr, ur)
    assert np.array_equal(xs, xp) and np.array_equal(us, up)
    assert np.array_equal(st["iters"], sp["iters"]) and np.array_equal(st["status"], sp["status"]) and np.array_equal(st["J"], sp["J"])


def test_return_slot_quirks(gpu):
    """optimize()'s return slot (optcon.py:499-505) in one mixed batch: instance 0 exhausts max_iters = 4 (-> last iterate),
    instance 1 starts converged and stops at kk = 0 (-> all zeros)."""
    d = golden("newton_quirks.npz")
    xi = np.stack([d["a_xx_init"], d["b_xx_init"]])
    ui = np.stack([d["a_uu_init"], d["b_uu_init"]])
    with gpu.BatchedNewton(2, TT=1000, state="f64", refs_shared=True, max_iters=4) as bn:
        bn.set_weights(d["Q"], d["R"], d["QT"])
        bn.set_refs(d["xx_ref"], d["uu_ref"])
        bn.set_init(xi, ui)
        total = bn.solve()
        xs, us = bn.result()
        st, h = bn.stats(), bn.history()
    assert total == 4 and list(st["iters"]) == [3, 1] and list(st["status"]) == [2, 1]
    assert np.array_equal(h["stepsize"][0, :3], d["a_stepsize"]) and np.array_equal(h["stepsize"][1, :1], d["b_stepsize"])
    assert relerr(d["a_xx_star"], xs[0]) < 1e-9 and relerr(d["a_uu_star"], us[0]) < 1e-9
    assert not xs[1].any() and not us[1].any()


def test_minimal_horizon_and_options(gpu, oracle):
    """TT = 3 (two steps), non-default Armijo parameters and exact Hessian from the first iteration, vs the oracle."""
    rng = np.random.default_rng(4)
    n, TT = 5, 3
    from aircraftoptimalcontrol_b200 import refgen
    xr, ur = refgen.step_problem(rng.uniform(14, 18, n), rng.uniform(1.5, 3.5, n), tf=0.003, TT=TT)
    Q, R, QT = refgen.weights("step")
    kw = dict(max_iters=6, stepsize_0=0.8, cc=0.3, beta=0.5, armijo_maxiters=4, exact_after=-1)
    with gpu.BatchedNewton(n, TT=TT, **kw) as bn:
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess()
        xi, ui = bn.iterate_at(0)
        bn.solve()
        xs, us = bn.result()
        h, st = bn.history(), bn.stats()
    o = oracle.newton_batch(xr, ur, xi, ui, Q, R, QT, **kw)
    assert np.array_equal(st["iters"], o["iters"])
    for i in range(n):
        k = o["iters"][i]
        assert np.array_equal(h["stepsize"][i, :k], o["stepsize"][i, :k]) and np.array_equal(h["n_armijo"][i, :k], o["n_armijo"][i, :k])
    assert relerr(o["xx_star"], xs) < 1e-9 and relerr(o["uu_star"], us) < 1e-9


def test_survivor_generations_equal_in_place(gpu, oracle):
    """solve() gathers the still-iterating instances into smaller internal batches once half of a batch has finished
    (and folds them back); results, histories and statistics must be bit-identical to iterating in place, and match
    the oracle on a sample."""
    n, TT = 8192, 200
    xr, ur, Q, R, QT = _random_batch(n, TT, 99, 0.2)
    out = []
    for gen in (True, False):
        with gpu.BatchedNewton(n, TT=TT, armijo="lazy", generations=gen) as bn:
            bn.set_weights(Q, R, QT)
            bn.set_refs(xr, ur)
            bn.init_guess()
            xi, ui = bn.iterate_at(0)
            total = bn.solve()
            out.append((total, bn.result(), bn.iterate_at(0), bn.history(), bn.stats()))
    (t0, (x0, u0), (xl0, ul0), h0, s0), (t1, (x1, u1), (xl1, ul1), h1, s1) = out
    assert t0 == t1 and np.array_equal(s0["iters"], s1["iters"]) and np.array_equal(s0["status"], s1["status"])
    assert len(np.unique(s0["iters"])) > 3  # instances finish at different iterations, so generations were spawned
    assert np.array_equal(x0, x1) and np.array_equal(u0, u1) and np.array_equal(xl0, xl1) and np.array_equal(ul0, ul1)
    for k in ("JJ", "descent", "stepsize", "n_armijo"):
        assert np.array_equal(h0[k], h1[k]), k
    assert np.array_equal(s0["J"], s1["J"]) and np.array_equal(s0["n_reg"], s1["n_reg"])
    idx = np.arange(0, n, 64)
    o = oracle.newton_batch(xr[idx], ur[idx], xi[idx], ui[idx], Q, R, QT)
    assert np.array_equal(s0["iters"][idx], o["iters"])
    assert relerr(o["xx_star"], x0[idx]) < 1e-9 and relerr(o["uu_star"], u0[idx]) < 1e-9


def test_float_state_slots_are_lossless(gpu):
    """With the float32 state quantisation every stored state is a float32 value, so the library keeps the state iterates as
    float in HBM.  Forcing float64 buffers (x_storage="f64") must give bit-identical results, histories and statistics --
    including through survivor generations and for an exact-but-arbitrary float64 x0."""
    n, TT = 8192, 200
    xr, ur, Q, R, QT = _random_batch(n, TT, 5, 0.2)
    dx0 = np.random.default_rng(3).normal(size=(n, 6)) * np.array([.05, .05, .2, .02, .05, .02])  # x0 is not a float32 value
    out = []
    for xs_mode in ("auto", "f64"):
        with gpu.BatchedNewton(n, TT=TT, armijo="lazy", x_storage=xs_mode) as bn:
            bn.set_weights(Q, R, QT)
            bn.set_refs(xr, ur)
            bn.init_guess(dx0=dx0)
            first = bn.iterate_at(0)
            total = bn.solve()
            out.append((total, first, bn.result(), bn.iterate_at(0), bn.iterate_at(1), bn.history(), bn.stats()))
    a, b = out
    assert a[0] == b[0]
    for k in (1, 2, 3, 4):
        assert np.array_equal(a[k][0], b[k][0]) and np.array_equal(a[k][1], b[k][1]), k
    assert np.array_equal(a[1][0][:, :, 0], xr[:, :, 0] + dx0)  # the exact float64 x0 comes back, not its float32 rounding
    for k in ("JJ", "descent", "stepsize", "n_armijo"):
        assert np.array_equal(a[5][k], b[5][k]), k
    for k in ("iters", "status", "J", "descent", "n_reg"):
        assert np.array_equal(a[6][k], b[6][k]), k


def test_inexact_initial_states_keep_float64_slots(gpu, oracle):
    """acoc_set_init with states that are NOT float32 values (nothing the reference would produce, but legal input): the
    context falls back to float64 state buffers, the first iterate comes back unchanged and the solve matches the oracle."""
    d = golden("newton_step_f32.npz")
    n, TT = 8, 1000
    rng = np.random.default_rng(11)
    xi = np.repeat(d["xx_init"][None], n, 0)
    xi[:, :, 1:] += rng.normal(size=(n, 6, TT - 1)) * 1e-7
    ui = np.repeat(d["uu_init"][None], n, 0) + rng.normal(size=(n, 2, TT)) * 0.1
    with gpu.BatchedNewton(n, TT=TT, refs_shared=True, armijo="lazy") as bn:
        bn.set_weights(d["Q"], d["R"], d["QT"])
        bn.set_refs(d["xx_ref"], d["uu_ref"])
        bn.set_init(xi, ui)
        x_back, u_back = bn.iterate_at(0)
        bn.iterate(6)
        xs, us = bn.iterate_at(0)
        h = bn.history()
    assert np.array_equal(x_back, xi) and np.array_equal(u_back, ui)
    o = oracle.newton_batch(np.repeat(d["xx_ref"][None], n, 0), np.repeat(d["uu_ref"][None], n, 0), xi, ui, d["Q"], d["R"], d["QT"], n_iters_cap=6)
    assert np.array_equal(h["stepsize"][:, :6], o["stepsize"][:, :6])
    assert np.max(np.abs(h["JJ"][:, :6] - o["JJ"][:, :6]) / np.abs(o["JJ"][:, :6])) < 1e-9
    # (the oracle's capped result is the newest iterate with optimize()'s uu[:, -1] = uu[:, -2] fix-up applied)
    assert relerr(o["xx_star"], xs) < 1e-9 and relerr(o["uu_star"][:, :, :-1], us[:, :, :-1]) < 1e-9


# FP32 mode (ACOC_FP32): stated tolerance against the float64 parity path / the reference
# (the optimum is flat: where the float32-noise phase of the line search happens to stop moves the states by a few 1e-4; the
# reference's own float32-state and float64-state variants differ from each other by the same amount)
FP32_TOL = dict(J_rel=2e-6, x_abs=2e-3, u_rel_range=2e-4)


@pytest.mark.parametrize("name", ["newton_step_f32", "newton_acro_f32"])
def test_fp32_mode_configs_1_2(gpu, name):
    """Optional FP32 mode on configs 1 and 2: converges by the reference's criterion in about as many iterations as the
    reference needs with its float32 state (the float64-state variant of the same problem needs 17 / 20), to the same optimum
    within FP32_TOL."""
    d = golden(name + ".npz")
    TT = d["xx_ref"].shape[1]
    with gpu.BatchedNewton(1, TT=TT, refs_shared=True, armijo="lazy", precision="f32") as bn:
        bn.set_weights(d["Q"], d["R"], d["QT"])
        bn.set_refs(d["xx_ref"], d["uu_ref"])
        bn.set_init(d["xx_init"][None], d["uu_init"][None])
        bn.solve()
        xl, ul = bn.iterate_at(0)
        h, st = bn.history(), bn.stats()
    k = int(st["iters"][0])
    assert st["status"][0] == 1 and 0.7 * int(d["iters"]) <= k <= 1.5 * int(d["iters"])
    assert abs(h["JJ"][0, k - 1] - d["JJ"][-1]) / d["JJ"][-1] < FP32_TOL["J_rel"]
    assert np.max(np.abs(h["JJ"][0, :8] - d["JJ"][:8]) / d["JJ"][:8]) < 1e-4  # the clean phase follows the same path
    assert np.max(np.abs(xl[0] - d["xx_last"])) < FP32_TOL["x_abs"]
    assert np.max(np.abs(ul[0] - d["uu_last"])) < FP32_TOL["u_rel_range"] * np.max(np.abs(d["uu_last"]))


def test_fp32_mode_batch_vs_float64_path(gpu):
    """FP32 mode on a config-4 style batch against the float64 path on the same inputs: every instance converges, final costs
    and trajectories within FP32_TOL, iteration counts of the same order."""
    n, TT = 256, 1000
    xr, ur, Q, R, QT = _random_batch(n, TT, 2024, 1.0)
    res = {}
    for prec in ("f64", "f32"):
        with gpu.BatchedNewton(n, TT=TT, armijo="lazy", precision=prec) as bn:
            bn.set_weights(Q, R, QT)
            bn.set_refs(xr, ur)
            bn.init_guess()
            bn.solve()
            res[prec] = (bn.iterate_at(0), bn.stats())
    (x64, u64), s64 = res["f64"]
    (x32, u32), s32 = res["f32"]
    assert np.all(s32["status"] == 1) and np.all(s64["status"] == 1)
    assert np.max(np.abs(s32["J"] - s64["J"]) / s64["J"]) < FP32_TOL["J_rel"]
    assert np.max(np.abs(x32 - x64)) < FP32_TOL["x_abs"]
    assert np.max(np.abs(u32 - u64)) < FP32_TOL["u_rel_range"] * np.max(np.abs(u64))
    assert 0.6 < s32["iters"].mean() / s64["iters"].mean() < 1.6


def test_tma_rings_and_plain_loads_identical(gpu):
    """The warp-private TMA rings (default) and the plain-load sweeps (ACOC_NO_TMA) run the same per-step arithmetic: whole solves
    must agree bit for bit -- ragged batch (padding lanes in the last tile), per-instance references, lazy Armijo with its Gauss-Newton
    candidate split, survivor generations."""
    n, TT = 4099, 200
    xr, ur, Q, R, QT = _random_batch(n, TT, 17, 0.2)
    out = []
    for tma in (True, False):
        with gpu.BatchedNewton(n, TT=TT, armijo="lazy", tma=tma) as bn:
            bn.set_weights(Q, R, QT)
            bn.set_refs(xr, ur)
            bn.init_guess()
            total = bn.solve()
            out.append((total, bn.result(), bn.iterate_at(0), bn.history(), bn.stats(), bn.deltau(), bn.gains()))
    a, b = out
    assert a[0] == b[0]
    for k in (1, 2):
        assert np.array_equal(a[k][0], b[k][0]) and np.array_equal(a[k][1], b[k][1]), k
    for k in ("JJ", "descent", "stepsize", "n_armijo"):
        assert np.array_equal(a[3][k], b[3][k]), k
    for k in ("iters", "status", "J", "descent", "n_reg"):
        assert np.array_equal(a[4][k], b[4][k]), k
    act = a[4]["iters"] == a[4]["iters"].max()  # deltau / gains are scratch of the last iteration an instance took part in
    assert np.array_equal(a[5][act], b[5][act]) and np.array_equal(a[6][0][act], b[6][0][act])


@pytest.mark.parametrize("n", [40001, 90001])
def test_two_range_sweep_identical(gpu, n):
    """A fully active batch with more backward CTAs than fit on the device at once (4 per SM x 148 SMs = 592 CTAs of 64 instances) is
    swept as two independent tile ranges on two streams.  Instances are independent, so results, histories and statistics must be
    bit-identical to the single-stream sweep -- checked on ragged 40,001- and 90,001-instance batches (626 / 1407 CTAs; the survivor
    generations of the larger one are themselves big enough to be split while holding fewer instances than their capacity) through a
    whole solve."""
    TT = 48
    xr, ur, Q, R, QT = _random_batch(n, TT, 23, 0.048)
    out = []
    for split in (True, False):
        with gpu.BatchedNewton(n, TT=TT, armijo="lazy", split=split) as bn:
            bn.set_weights(Q, R, QT)
            bn.set_refs(xr, ur)
            bn.init_guess()
            bn.iterate(3)
            launches = bn.timing()["launches"]
            mid = bn.iterate_at(0)
            total = bn.solve()
            out.append((total, mid, bn.result(), bn.iterate_at(0), bn.history(), bn.stats(), launches))
    a, b = out
    assert a[6] > b[6]  # the split path really ran: iterations 1 and 2 launched their kernels once per range
    assert a[0] == b[0] and a[0] > 3 * n
    for k in (1, 2, 3):
        assert np.array_equal(a[k][0], b[k][0]) and np.array_equal(a[k][1], b[k][1]), k
    for k in ("JJ", "descent", "stepsize", "n_armijo"):
        assert np.array_equal(a[4][k], b[4][k]), k
    for k in ("iters", "status", "J", "descent", "n_reg"):
        assert np.array_equal(a[5][k], b[5][k]), k


@pytest.mark.parametrize("state,armijo,n,TT", [("f32", "lazy", 1000, 1000), ("f32", "speculative", 37, 300), ("f64", "lazy", 300, 300),
                                                ("f32", "lazy", 9001, 250), ("f64", "lazy", 4200, 120)])
def test_fused_search_identical(gpu, state, armijo, n, TT):
    """Small batches run the LQ forward pass, every Armijo candidate and the exhausted step as ONE sweep (k_search_fused) and take
    get_update as a copy of the chosen row (k_pick).  It is the same per-step arithmetic as the separate sweeps, so whole solves must
    agree bit for bit with fused=False -- ragged batches (padding lanes, finished lanes inside live tiles), both state modes, both
    Armijo modes, searches that run to exhaustion (the float32-noise phase), max_iters small enough that some instances stop at the
    iteration limit (their result IS the last copied slot).
    Batches of more than 4096 instances fuse only the LQ forward pass with candidate 0 of the lazy search (k_forward_cand0_tma); the
    9001- and 4200-instance cases run through that kernel first and, once their survivor generations are small, through k_search_fused."""
    xr, ur, Q, R, QT = _random_batch(n, TT, 31, TT * 1e-3)
    out = []
    for fused in (True, False):
        with gpu.BatchedNewton(n, TT=TT, armijo=armijo, state=state, fused=fused, max_iters=26) as bn:
            bn.set_weights(Q, R, QT)
            bn.set_refs(xr, ur)
            bn.init_guess()
            bn.iterate(2)
            launches = bn.timing()["launches"]
            mid = bn.iterate_at(0)
            total = bn.solve()
            out.append((total, mid, bn.result(), bn.iterate_at(0), bn.history(), bn.stats(), launches, bn.deltau()))
    a, b = out
    if armijo == "lazy":
        assert a[6] != b[6]  # the fused path really ran (different number of launches per iteration; speculative: six either way)
    assert a[0] == b[0]
    for k in (1, 2, 3):
        assert np.array_equal(a[k][0], b[k][0]) and np.array_equal(a[k][1], b[k][1]), k
    for k in ("JJ", "descent", "stepsize", "n_armijo"):
        assert np.array_equal(a[4][k], b[4][k]), k
    for k in ("iters", "status", "J", "descent", "n_reg"):
        assert np.array_equal(a[5][k], b[5][k]), k
    assert np.any(a[5]["status"] == 1)
    if state == "f32" and TT == 1000:  # float32-noise phase: searches that run to exhaustion, instances stopped by the iteration limit
        assert np.any(a[4]["n_armijo"] == 10) and np.any(a[5]["status"] == 2)
    act = a[5]["iters"] == a[5]["iters"].max()
    assert np.array_equal(a[7][act], b[7][act])


@pytest.mark.parametrize("tma", [True, False])
def test_gradient_method_batch_matches_oracle(gpu, oracle, tma):
    """GradientMethod.optimize as a batch (ACOC_METHOD_GRADIENT): 70 randomised step references (ragged: padding lanes), lazy Armijo,
    12 iterations, against the C oracle (itself pinned to the live reference) -- Armijo steps and candidate counts identical,
    histories and iterates within 1e-9; TMA-ring and plain-load costate sweeps."""
    n, TT = 70, 1000
    xr, ur, Q, R, QT = _random_batch(n, TT, 77, 1.0)
    with gpu.BatchedNewton(n, TT=TT, armijo="lazy", method="gradient", max_iters=13, tma=tma) as bn:
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess()
        xi, ui = bn.iterate_at(0)
        total = bn.solve()
        xs, us = bn.result()
        xl, ul = bn.iterate_at(0)
        h, st = bn.history(), bn.stats()
    o = oracle.gradient_batch(xr, ur, xi, ui, Q, R, QT, max_iters=13, stepsize_0=1.0, armijo_maxiters=10)
    assert total == 12 * n and np.all(st["iters"] == 12) and np.all(o["iters"] == 12) and np.all(st["status"] == 2)
    assert np.array_equal(h["stepsize"][:, :12], o["stepsize"][:, :12]) and np.array_equal(h["n_armijo"][:, :12], o["n_armijo"][:, :12])
    assert len(np.unique(h["stepsize"][:, :12])) >= 4   # the searches really backtrack
    assert np.max(np.abs(h["JJ"][:, :12] - o["JJ"][:, :12]) / o["JJ"][:, :12]) < 1e-9
    assert np.max(np.abs(-h["descent"][:, :12] - o["descent"][:, :12]) / o["descent"][:, :12]) < 1e-9
    assert relerr(o["xx_star"], xs) < 1e-9 and relerr(o["uu_star"], us) < 1e-9


def test_gradient_method_fp32_mode(gpu):
    """The optional FP32 mode also covers the gradient method: 12 steepest-descent iterations of a 64-instance batch stay within the
    FP32 tolerances of the float64 path (final cost 2e-6 relative, inputs 2e-4 of their range)."""
    n, TT = 64, 1000
    xr, ur, Q, R, QT = _random_batch(n, TT, 91, 1.0)
    res = {}
    for prec in ("f64", "f32"):
        with gpu.BatchedNewton(n, TT=TT, armijo="lazy", method="gradient", max_iters=13, precision=prec) as bn:
            bn.set_weights(Q, R, QT)
            bn.set_refs(xr, ur)
            bn.init_guess()
            bn.solve()
            res[prec] = (bn.iterate_at(0), bn.stats(), bn.history())
    (x64, u64), s64, h64 = res["f64"]
    (x32, u32), s32, h32 = res["f32"]
    assert np.all(s32["iters"] == 12) and np.all(s64["iters"] == 12)
    assert np.max(np.abs(s32["J"] - s64["J"]) / s64["J"]) < 1e-4      # 12 line searches on float32 noise: same descent path, not the same floats
    assert np.max(np.abs(h32["JJ"][:, 0] - h64["JJ"][:, 0]) / h64["JJ"][:, 0]) < FP32_TOL["J_rel"]
    assert np.max(np.abs(u32 - u64)) < 50 * FP32_TOL["u_rel_range"] * np.max(np.abs(u64))


def test_gradient_method_survivor_generations(gpu):
    """The gradient method shares the driver's survivor generations: with a loose termination threshold (chosen from a probing run so
    that about half of the instances cross it within 11 iterations, at different iterations) the survivors of an 8192-instance batch move
    into smaller contexts, and everything must equal iterating in place bit for bit."""
    n, TT = 8192, 100
    xr, ur, Q, R, QT = _random_batch(n, TT, 55, 0.1)

    def solver(**kw):
        bn = gpu.BatchedNewton(n, TT=TT, armijo="lazy", method="gradient", **kw)
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess()
        return bn

    with solver(max_iters=13, generations=False) as bn:
        bn.solve()
        thr = float(np.median(np.min(-bn.history()["descent"][:, :12], axis=1)))   # history holds the slope -|deltau|^2
    out = []
    for gen in (True, False):
        with solver(max_iters=40, generations=gen, term_cond=-thr) as bn:
            total = bn.solve()
            out.append((total, bn.result(), bn.history(), bn.stats()))
    a, b = out
    assert a[0] == b[0]
    assert np.array_equal(a[1][0], b[1][0]) and np.array_equal(a[1][1], b[1][1])
    for k in ("JJ", "descent", "stepsize", "n_armijo"):
        assert np.array_equal(a[2][k], b[2][k]), k
    for k in ("iters", "status", "J", "descent"):
        assert np.array_equal(a[3][k], b[3][k]), k
    it = a[3]["iters"]
    assert np.sum(a[3]["status"] == 1) >= n // 2 and len(np.unique(it[a[3]["status"] == 1])) >= 2   # they stopped, at different iterations


def test_dense_weights_and_model_parameters(gpu, oracle):
    """Dense (non-diagonal, symmetric) Q/R/QT take the general cost path of the sweeps, and non-default Dynamics attributes (heavier
    aircraft, other drag/lift coefficients, dt = 2e-3) go through acoc_set_model: 7 Newton iterations with the exact Hessian from
    iteration 3 on (exact_after = 2), a 5-instance batch, against the oracle with the same weights and parameters."""
    d = golden("newton_step_f32.npz")
    rng = np.random.default_rng(11)
    n, TT = 5, 160
    sl = slice(0, TT)
    xr = np.repeat(d["xx_ref"][None, :, sl], n, 0) * (1.0 + 0.02 * rng.normal(size=(n, 6, 1)))
    ur = np.repeat(d["uu_ref"][None, :, sl], n, 0)
    E = rng.normal(size=(6, 6)) * 1e-4
    Q = d["Q"] + E @ E.T
    QT = d["QT"] + 3 * (E @ E.T)
    R = d["R"] + 1e-7 * np.array([[1.0, 0.3], [0.3, 2.0]])
    params = np.array([0.19, 2.1, 3.4, 14.0, 9.81, 0.61, 1.2, 0.3, 2e-3])   # cd0, cda, cla, m, g, S, rho, J, dt
    xi = np.zeros((n, 6, TT))
    ui = np.zeros((n, 2, TT))
    for i in range(n):
        xi[i], ui[i] = oracle.initial_trajectory(xr[i], params=params)
    o = oracle.newton_batch(xr, ur, xi, ui, Q, R, QT, params=params, exact_after=2, max_iters=8)
    assert o["iters"].min() >= 4 and len(np.unique(o["stepsize"][:, :4])) >= 4   # exact-Hessian iterations that backtrack
    for armijo in ("lazy", "speculative"):
        with gpu.BatchedNewton(n, TT=TT, armijo=armijo, params=params, exact_after=2, max_iters=8) as bn:
            bn.set_weights(Q, R, QT)
            bn.set_refs(xr, ur)
            bn.set_init(xi, ui)
            bn.solve()
            xs, us = bn.result()
            h, st = bn.history(), bn.stats()
        assert np.array_equal(st["iters"], o["iters"])
        for i in range(n):
            k = int(o["iters"][i])
            assert np.array_equal(h["stepsize"][i, :k], o["stepsize"][i, :k]) and np.array_equal(h["n_armijo"][i, :k], o["n_armijo"][i, :k])
            assert np.max(np.abs(h["JJ"][i, :k] - o["JJ"][i, :k]) / np.abs(o["JJ"][i, :k])) < 1e-9
            assert np.max(np.abs(h["descent"][i, :k] - o["descent"][i, :k]) / np.abs(o["descent"][i, :k])) < 1e-9
        assert relerr(o["xx_star"], xs) < 1e-9 and relerr(o["uu_star"], us) < 1e-9
