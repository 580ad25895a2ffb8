"""Round-2 parity cases (through the C ABI, libacoc.so): BASELINE.json configs[4] as a batch, a sample of the real
65,536-instance pipelined solve, the non-finite freeze, the weight-symmetry contract, armijo_maxiters > 10 and survivor
generations with non-default Armijo parameters.  Same bar as tests/test_gpu_parity.py: identical iteration counts and
Armijo steps, histories / trajectories within 1e-9 relative of the CPU oracle (itself pinned to the live reference)."""
import numpy as np
import pytest

from tests.util import golden, relerr

pytestmark = pytest.mark.gpu


def _check_against_oracle(o, h, st, xs, us, idx=None, tol=1e-9):
    """o: oracle.newton_batch result of the sampled instances; h/st/xs/us: GPU history, stats, results; idx: their indices."""
    idx = np.arange(len(o["iters"])) if idx is None else np.asarray(idx)
    assert np.array_equal(st["iters"][idx], o["iters"]), (st["iters"][idx], o["iters"])
    for j, i in enumerate(idx):
        k = int(o["iters"][j])
        assert np.array_equal(h["stepsize"][i, :k], o["stepsize"][j, :k]), (i, h["stepsize"][i, :k], o["stepsize"][j, :k])
        assert np.array_equal(h["n_armijo"][i, :k], o["n_armijo"][j, :k]), i
        assert np.max(np.abs(h["JJ"][i, :k] - o["JJ"][j, :k]) / np.abs(o["JJ"][j, :k])) < tol, i
        assert np.max(np.abs(h["descent"][i, :k] - o["descent"][j, :k]) / np.abs(o["descent"][j, :k])) < tol, i
    assert relerr(o["xx_star"], xs[idx]) < tol and relerr(o["uu_star"], us[idx]) < tol


@pytest.mark.parametrize("armijo", ["lazy", "speculative"])
def test_config5_batch_matches_oracle(gpu, oracle, armijo):
    """BASELINE.json configs[4] (acrobatic_newton.py:133-196 with x0 + delta, bump height zf ~ U(2.0, 3.4), seed 7; optcon.py:398):
    the first 64 instances of the 1,048,576-instance batch at TT = 1000, device-side initial guess from the perturbed x0, solved to
    the reference's criterion -- iteration counts, every Armijo step / candidate count, histories and results vs the oracle."""
    from aircraftoptimalcontrol_b200 import refgen
    n, TT = 64, 1000
    xr, ur, dx0, Q, R, QT = refgen.config5(1048576, seed=7, lo=0, hi=n, TT=TT)
    with gpu.BatchedNewton(n, TT=TT, armijo=armijo) as bn:
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess(dx0=dx0)
        xi, ui = bn.iterate_at(0)
        total = bn.solve()
        xs, us = bn.result()
        h, st = bn.history(), bn.stats()
    assert np.array_equal(xi[:, :, 0], xr[:, :, 0] + dx0)  # x0 = xx_init[:,0] is the perturbed state, exactly
    o = oracle.newton_batch(xr, ur, xi, ui, Q, R, QT)
    assert total == int(o["iters"].sum()) and np.all(st["status"] == 1)
    assert len(np.unique(o["iters"])) > 2  # the perturbations / bump heights give different iteration counts
    _check_against_oracle(o, h, st, xs, us)


def test_pipelined_65536_sample_matches_oracle(gpu, oracle):
    """The benchmarked configuration itself: BASELINE.json configs[3], 65,536 instances at TT = 1000 through PipelinedNewton (4
    sub-batches, tile ranges, survivor generations, lazy search, fused sweeps) -- 64 instances sampled out of that solve are
    compared with the oracle run on the same inputs: iteration counts, every Armijo step, histories, results."""
    from aircraftoptimalcontrol_b200 import refgen
    n, TT = 65536, 1000
    xr, ur, Q, R, QT = refgen.config4(n, 2024, TT=TT)
    with gpu.PipelinedNewton(n, n_chunks=4, TT=TT, armijo="lazy") as pn:
        pn.set_weights(Q, R, QT)
        xs, us, st = pn.solve(xr, ur)
        idx = np.concatenate([np.arange(0, n, 1040), [n - 1]])[:64]   # every sub-batch, different tiles and lanes
        rows = {k: [] for k in ("stepsize", "n_armijo", "JJ", "descent")}
        for k, p in enumerate(pn.parts):
            lo, hi = pn.bounds[k], pn.bounds[k + 1]
            sel = idx[(idx >= lo) & (idx < hi)] - lo
            hk = p.history()
            for key in rows:
                rows[key].append(hk[key][sel])
    h = {key: np.concatenate(v) for key, v in rows.items()}
    assert np.all(st["status"] == 1) and st["iters"].min() >= 10
    xi = np.zeros((len(idx), 6, TT))
    ui = np.zeros((len(idx), 2, TT))
    for j, i in enumerate(idx):  # the device initial guess is bit-identical to this one (test_rollouts_bit_exact_at_scale)
        xi[j], ui[j] = oracle.initial_trajectory(xr[i])
    o = oracle.newton_batch(xr[idx], ur[idx], xi, ui, Q, R, QT)
    sub = dict(iters=st["iters"][idx])
    _check_against_oracle(o, h, sub, xs[idx], us[idx])


def test_nonfinite_instance_freezes_alone(gpu):
    """aircraft_simplified.py:310,321,363 divide by V: an instance with V = 0 in x0 (and one with a NaN input) must be frozen with
    ACOC_INST_NONFINITE after its first loop body while every other instance of its tile and batch is bit-identical to the same batch
    without the poisoned inputs."""
    n, TT = 96, 300
    from aircraftoptimalcontrol_b200 import refgen
    rng = np.random.default_rng(12)
    xr, ur = refgen.step_problem(rng.uniform(14, 18, n), rng.uniform(1.5, 3.5, n), tf=0.3, TT=TT)
    Q, R, QT = refgen.weights("step")
    with gpu.BatchedNewton(n, TT=TT) as bn:
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess()
        xi, ui = bn.iterate_at(0)
    bad = [5, 40]
    xb, ub = xi.copy(), ui.copy()
    xb[5, 2, 0] = 0.0          # V = 0 at t = 0: 1/V in the linearisation and in every rollout
    ub[40, 0, 150] = np.nan    # NaN thrust in the middle of the horizon
    out = []
    for armijo in ("lazy", "speculative"):
        for x_, u_ in ((xi, ui), (xb, ub)):
            with gpu.BatchedNewton(n, TT=TT, armijo=armijo) as bn:
                bn.set_weights(Q, R, QT)
                bn.set_refs(xr, ur)
                bn.set_init(x_, u_)
                bn.solve()
                out.append((bn.result(), bn.history(), bn.stats()))
    good = np.setdiff1d(np.arange(n), bad)
    for a, b in ((out[0], out[1]), (out[2], out[3]), (out[0], out[2])):
        (xa, ua), ha, sa = a
        (xb_, ub_), hb, sb = b
        assert np.array_equal(xa[good], xb_[good]) and np.array_equal(ua[good], ub_[good])
        for k in ("JJ", "descent", "stepsize", "n_armijo"):
            assert np.array_equal(ha[k][good], hb[k][good]), k
        for k in ("iters", "status", "J", "descent", "n_reg"):
            assert np.array_equal(sa[k][good], sb[k][good]), k
    for (_, _), _, s in (out[1], out[3]):
        assert list(s["status"][bad]) == [3, 3] and list(s["iters"][bad]) == [1, 1]
        assert np.all(s["status"][good] == 1)
    assert np.all(out[0][2]["status"] == 1)


def test_weight_symmetry_contract(gpu):
    """Documented deviation (include/acoc.h, acoc_set_weights): the batched Newton context rejects non-symmetric Q / R / QT, the
    pointwise Cost entry points accept any matrix like aircraft_simplified.py:61-68."""
    from aircraftoptimalcontrol_b200 import _lib
    from aircraftoptimalcontrol_b200.aircraft_simplified import Cost
    rng = np.random.default_rng(8)
    Qn, Rn, QTn = rng.normal(size=(6, 6)), rng.normal(size=(2, 2)), rng.normal(size=(6, 6))
    Qs, Rs, QTs = Qn + Qn.T, Rn + Rn.T, QTn + QTn.T
    with gpu.BatchedNewton(4, TT=20) as bn:
        for args in ((Qn, Rs, QTs), (Qs, Rn, QTs), (Qs, Rs, QTn)):
            with pytest.raises(_lib.AcocError, match="symmetric"):
                bn.set_weights(*args)
        bn.set_weights(Qs, Rs, QTs)  # symmetric dense weights are fine
    x, u, xr, ur = rng.normal(size=(50, 6)), rng.normal(size=(50, 2)), rng.normal(size=(50, 6)), rng.normal(size=(50, 2))
    c = Cost(Qn, Rn, QTn)
    ll, lx, lu = c.stagecost_batch(x, u, xr, ur)
    llT, lTx = c.termcost_batch(x, xr)
    dx, du = x - xr, u - ur
    ref_ll = 0.5 * np.einsum("ni,ij,nj->n", dx, Qn, dx) + 0.5 * np.einsum("ni,ij,nj->n", du, Rn, du)
    assert np.max(np.abs(ll - ref_ll)) < 1e-12 * np.max(np.abs(ref_ll))
    assert relerr(dx @ Qn.T, lx) < 1e-13 and relerr(du @ Rn.T, lu) < 1e-13           # lx = Q dx, lu = R du (:63-64)
    assert np.max(np.abs(llT - 0.5 * np.einsum("ni,ij,nj->n", dx, QTn, dx))) < 1e-12 * np.max(np.abs(llT))
    assert relerr(dx @ QTn.T, lTx) < 1e-13                                           # lTx = QT dx (:94)


def test_armijo_maxiters_20_single_trajectory(gpu, oracle):
    """NewtonMethod's constructor default armijo_maxiters = 20 (optcon.py:335) on config 1: more candidates than one candidate CTA
    holds (the kernel strides inside the thread) and no fused small-batch search -- vs the oracle, lazy and speculative."""
    d = golden("newton_step_f32.npz")
    TT = d["xx_ref"].shape[1]
    o = oracle.newton(d["xx_ref"], d["uu_ref"], d["xx_init"], d["uu_init"], d["Q"], d["R"], d["QT"], armijo_maxiters=20)
    k = int(o["iters"])
    assert o["n_armijo"].max() > 10  # the longer search is actually used
    for armijo in ("lazy", "speculative"):
        with gpu.BatchedNewton(1, TT=TT, refs_shared=True, armijo=armijo, armijo_maxiters=20) as bn:
            bn.set_weights(d["Q"], d["R"], d["QT"])
            bn.set_refs(d["xx_ref"], d["uu_ref"])
            bn.set_init(d["xx_init"][None], d["uu_init"][None])
            total = bn.solve()
            xs, us = bn.result()
            h = bn.history()
        assert total == k
        assert np.array_equal(h["stepsize"][0, :k], o["stepsize"]) and np.array_equal(h["n_armijo"][0, :k], o["n_armijo"])
        assert np.max(np.abs(h["JJ"][0, :k] - o["JJ"]) / np.abs(o["JJ"])) < 1e-9
        assert relerr(o["xx_star"], xs[0]) < 1e-9 and relerr(o["uu_star"], us[0]) < 1e-9


def _short_batch(n, TT, seed):
    from aircraftoptimalcontrol_b200 import refgen
    rng = np.random.default_rng(seed)
    xr, ur = refgen.step_problem(rng.uniform(14, 18, n), rng.uniform(1.5, 3.5, n), tf=TT * 1e-3, TT=TT)
    return (xr, ur) + refgen.weights("step")


def test_armijo_maxiters_20_batch(gpu, oracle):
    """armijo_maxiters = 20 on a batch of more than 4096 instances (lazy search over the need lists, candidates 1..19 in one launch,
    survivor generations): a 64-instance sample vs the oracle."""
    n, TT = 4608, 200
    xr, ur, Q, R, QT = _short_batch(n, TT, 21)
    with gpu.BatchedNewton(n, TT=TT, armijo="lazy", armijo_maxiters=20) as bn:
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess()
        xi, ui = bn.iterate_at(0)
        bn.solve()
        xs, us = bn.result()
        h, st = bn.history(), bn.stats()
    idx = np.arange(0, n, 72)
    o = oracle.newton_batch(xr[idx], ur[idx], xi[idx], ui[idx], Q, R, QT, armijo_maxiters=20)
    _check_against_oracle(o, h, st, xs, us, idx)


@pytest.mark.parametrize("kw", [dict(stepsize_0=0.5), dict(beta=0.5)])
def test_generations_with_nondefault_armijo_table(gpu, oracle, kw):
    """Survivor generations must search with the PARENT's step table: only stepsize_0 or beta changed (max_iters / armijo_maxiters at
    their defaults, so the child context is not re-created) -- bit-identical to iterating in place, sample vs the oracle."""
    n, TT = 8192, 200
    xr, ur, Q, R, QT = _short_batch(n, TT, 31)
    out = []
    for gen in (True, False):
        with gpu.BatchedNewton(n, TT=TT, armijo="lazy", generations=gen, **kw) as bn:
            bn.set_weights(Q, R, QT)
            bn.set_refs(xr, ur)
            bn.init_guess()
            xi, ui = bn.iterate_at(0)
            bn.solve()
            out.append((bn.result(), bn.history(), bn.stats()))
    ((x0, u0), h0, s0), ((x1, u1), h1, s1) = out
    assert len(np.unique(s0["iters"])) > 3
    assert np.array_equal(x0, x1) and np.array_equal(u0, u1)
    for k in ("JJ", "descent", "stepsize", "n_armijo"):
        assert np.array_equal(h0[k], h1[k]), k
    for k in ("iters", "status", "J"):
        assert np.array_equal(s0[k], s1[k]), k
    idx = np.arange(0, n, 128)
    o = oracle.newton_batch(xr[idx], ur[idx], xi[idx], ui[idx], Q, R, QT, **kw)
    _check_against_oracle(o, h0, s0, x0, u0, idx)


def test_options_can_be_changed_without_growing_the_context(gpu):
    """acoc_set_options resizes / reuses the history buffers instead of stacking new ones."""
    with gpu.BatchedNewton(256, TT=50) as bn:
        b0 = bn.device_bytes
        import ctypes as C
        from aircraftoptimalcontrol_b200 import _lib as L
        for _ in range(3):
            L.check(L.lib().acoc_set_options(bn._h, C.addressof(bn.opts)))
        assert bn.device_bytes == b0


def test_device_reference_generators_bit_identical(gpu):
    """acoc_set_refs_generated (main_newton_method.py:96-142, acrobatic_newton.py:99-154 on the device) vs the host arrays of
    refgen (pinned bit for bit to the scripts' globals in tests/test_refgen.py): identical references, hence identical solves."""
    from aircraftoptimalcontrol_b200 import refgen
    n, TT = 200, 1000
    zf, xf = refgen.config4_params(n, 2024)
    xr, ur = refgen.step_problem(xf, zf, TT=TT)
    with gpu.BatchedNewton(n, TT=TT, armijo="lazy") as bn:
        bn.set_refs_step(zf, xf)
        gx, gu = bn.refs()
        assert np.array_equal(gx, xr) and np.array_equal(gu, ur)
    dx0, zf5 = refgen.config5_params(n, 7)
    xr5, ur5 = refgen.acrobatic_problem(zf5, TT=TT)
    with gpu.BatchedNewton(n, TT=TT, armijo="lazy") as bn:
        bn.set_refs_acrobatic(zf5)
        gx, gu = bn.refs()
        assert np.array_equal(gx, xr5) and np.array_equal(gu, ur5)
    # other horizons / durations
    for TT2, tf in ((37, 0.037), (300, 0.3)):
        a, b = refgen.step_problem(xf[:40], zf[:40], tf=tf, TT=TT2)
        with gpu.BatchedNewton(40, TT=TT2) as bn:
            bn.set_refs_step(zf[:40], xf[:40], tf=tf)
            gx, gu = bn.refs()
        assert np.array_equal(gx, a) and np.array_equal(gu, b)
    with gpu.BatchedNewton(4, TT=50, refs_shared=True) as bn:
        with pytest.raises(ValueError):
            bn.set_refs_step(zf[:4], xf[:4])


def test_float32_state_download_is_lossless(gpu):
    """acoc_get_result_f32: the states of a float32-quantised solve downloaded as float32 equal the float64 download exactly (t >= 1),
    column 0 is float32(x0) with the exact x0 returned beside it; float64-state contexts refuse."""
    from aircraftoptimalcontrol_b200 import _lib, refgen
    n, TT = 300, 200
    xr, ur, Q, R, QT = _short_batch(n, TT, 41)
    dx0 = np.random.default_rng(2).normal(size=(n, 6)) * np.array([.05, .05, .2, .02, .05, .02])
    with gpu.BatchedNewton(n, TT=TT, armijo="lazy", max_iters=9) as bn:   # mixed result slots: some converge, some hit max_iters
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess(dx0=dx0)
        bn.solve()
        x64, u64 = bn.result()
        x32, u32, x0 = bn.result_f32()
    assert x32.dtype == np.float32 and np.array_equal(x32[:, :, 1:].astype(np.float64), x64[:, :, 1:]) and np.array_equal(u32, u64)
    assert np.array_equal(x0, xr[:, :, 0] + dx0) and np.array_equal(x0, x64[:, :, 0])
    assert np.array_equal(x32[:, :, 0], x0.astype(np.float32))
    with gpu.BatchedNewton(8, TT=TT, state="f64") as bn:
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr[:8], ur[:8])
        bn.init_guess()
        bn.iterate(1)
        with pytest.raises(_lib.AcocError, match="float32"):
            bn.result_f32()


def test_pipelined_generated_refs_and_f32_download_equal_host_path(gpu):
    """PipelinedNewton with refs=("step"|"acrobatic", ...) and x_dtype=float32 (16 B per instance up, 40 B per step down) returns
    exactly what the host-buffer path returns."""
    from aircraftoptimalcontrol_b200 import refgen
    n, TT = 400, 300
    rng = np.random.default_rng(5)
    zf, xf = rng.uniform(1.5, 3.5, n), rng.uniform(14, 18, n)
    xr, ur = refgen.step_problem(xf, zf, tf=0.3, TT=TT)
    Q, R, QT = refgen.weights("step")
    with gpu.PipelinedNewton(n, n_chunks=3, TT=TT, armijo="lazy") as pn:
        pn.set_weights(Q, R, QT)
        xa, ua, sa = pn.solve(xr, ur)
        xb, ub, sb = pn.solve(refs=("step", zf, xf, 0.3), x_dtype=np.float32)
    assert np.array_equal(xa[:, :, 1:], xb[:, :, 1:].astype(np.float64)) and np.array_equal(ua, ub)
    assert np.array_equal(sb["x0"], xa[:, :, 0])
    for k in ("iters", "status", "J", "descent"):
        assert np.array_equal(sa[k], sb[k]), k
    dx0 = rng.normal(size=(n, 6)) * np.array([.05, .05, .2, .02, .05, .02])
    zf5 = rng.uniform(2.0, 3.4, n)
    xr5, ur5 = refgen.acrobatic_problem(zf5, tf=0.3, TT=TT)
    Q, R, QT = refgen.weights("acro")
    with gpu.PipelinedNewton(n, n_chunks=2, TT=TT, armijo="lazy", max_iters=12) as pn:
        pn.set_weights(Q, R, QT)
        xa, ua, sa = pn.solve(xr5, ur5, dx0=dx0)
        xb, ub, sb = pn.solve(refs=("acrobatic", zf5, 18, 0.3), dx0=dx0, x_dtype=np.float32)
    assert np.array_equal(xa[:, :, 1:], xb[:, :, 1:].astype(np.float64)) and np.array_equal(ua, ub)
    assert np.array_equal(sa["iters"], sb["iters"]) and np.array_equal(sa["J"], sb["J"])


@pytest.mark.parametrize("kind,n,TT,kw", [("step", 4500, 200, dict(armijo="lazy")), ("step", 300, 300, dict(armijo="speculative")),
                                          ("acrobatic", 4500, 200, dict(armijo="lazy", max_iters=14)), ("step", 700, 150, dict(armijo="lazy", tma=False)),
                                          ("step", 40001, 48, dict(armijo="lazy"))])
def test_compact_references_identical_solves(gpu, kind, n, TT, kw):
    """References built on the device are kept in their parametric form (X, Z formed in the sweeps from shared tables, only V stored:
    8 or 0 instead of 64 bytes per instance and step).  Whole solves must be bit-identical to the same references written out as
    per-instance arrays (refs_compact=False) and to the host arrays uploaded with set_refs -- through the TMA sweeps, the candidate
    ring, the fused small-batch search, the plain-load kernels, tile ranges and survivor generations."""
    from aircraftoptimalcontrol_b200 import refgen
    rng = np.random.default_rng(n + TT)
    tf = TT * 1e-3
    dx0 = None
    if kind == "step":
        zf, xf = rng.uniform(1.5, 3.5, n), rng.uniform(14, 18, n)
        xr, ur = refgen.step_problem(xf, zf, tf=tf, TT=TT)
        Q, R, QT = refgen.weights("step")
    else:
        zf = rng.uniform(2.0, 3.4, n)
        xr, ur = refgen.acrobatic_problem(zf, tf=tf, TT=TT)
        dx0 = rng.normal(size=(n, 6)) * np.array([.05, .05, .2, .02, .05, .02])
        Q, R, QT = refgen.weights("acro")
    out = []
    for mode in ("host", "expanded", "compact"):
        with gpu.BatchedNewton(n, TT=TT, refs_compact=(mode != "expanded"), **kw) as bn:
            bn.set_weights(Q, R, QT)
            if mode == "host":
                bn.set_refs(xr, ur)
            elif kind == "step":
                bn.set_refs_step(zf, xf, tf=tf)
            else:
                bn.set_refs_acrobatic(zf, tf=tf)
            bn.init_guess(dx0=dx0)
            first = bn.iterate_at(0)
            total = bn.solve()
            out.append((total, first, bn.result(), bn.iterate_at(0), bn.history(), bn.stats()))
            if mode == "compact":
                gx, gu = bn.refs()
                assert np.array_equal(gx, xr) and np.array_equal(gu, ur)
    a = out[0]
    assert a[0] > n
    for b in out[1:]:
        assert a[0] == b[0]
        for k in (1, 2, 3):
            assert np.array_equal(a[k][0], b[k][0]) and np.array_equal(a[k][1], b[k][1]), k
        for k in ("JJ", "descent", "stepsize", "n_armijo"):
            assert np.array_equal(a[4][k], b[4][k]), k
        for k in ("iters", "status", "J", "descent", "n_reg"):
            assert np.array_equal(a[5][k], b[5][k]), k


def _pinned(shape, dtype):
    import torch
    t = torch.empty(shape, dtype=torch.float32 if dtype == np.float32 else torch.float64, pin_memory=True)
    return t, t.numpy()


@pytest.mark.parametrize("n,TT,kw", [(8192, 200, dict()), (300, 200, dict(max_iters=9)), (9000, 120, dict(state="f64")), (5000, 150, dict(max_iters=22))])
def test_solve_deliver_equals_solve_and_result(gpu, n, TT, kw):
    """acoc_newton_solve_deliver writes optimize()'s results straight into page-locked host arrays, the finished instances of a batch
    at the moment the rest moves on to a survivor generation.  Must equal solve() + result() / result_f32() exactly -- float32 and
    float64 states, batches with and without generations, mixed result slots (converged / max_iters), pageable arrays (staged path)."""
    xr, ur, Q, R, QT = _short_batch(n, TT, n)
    dx0 = np.random.default_rng(n).normal(size=(n, 6)) * np.array([.05, .05, .2, .02, .05, .02])
    f64_state = kw.get("state") == "f64"

    def fresh():
        bn = gpu.BatchedNewton(n, TT=TT, armijo="lazy", **kw)
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess(dx0=dx0)
        return bn

    with fresh() as bn:
        total = bn.solve()
        x64, u64 = bn.result()
        st = bn.stats()
        h = bn.history()
    for x_dtype, pinned in ((np.float64, True), (np.float32, True), (np.float64, False)):
        if x_dtype == np.float32 and f64_state:
            continue
        if pinned:
            keep_x, xs = _pinned((n, 6, TT), x_dtype)
            keep_u, us = _pinned((n, 2, TT), np.float64)
        else:
            xs, us = np.empty((n, 6, TT), dtype=x_dtype), np.empty((n, 2, TT))
        xs[...] = -7.0
        us[...] = -7.0
        with fresh() as bn:
            tot2, x0 = bn.solve_deliver((xs, us))
            st2, h2 = bn.stats(), bn.history()
            xr2, ur2 = bn.result()     # the context itself is in the same state as after solve()
        assert tot2 == total and np.array_equal(x0, xr[:, :, 0] + dx0)
        assert np.array_equal(us, u64)
        if x_dtype == np.float64:
            assert np.array_equal(xs, x64)
        else:
            assert np.array_equal(xs[:, :, 1:].astype(np.float64), x64[:, :, 1:])
            conv0 = st["iters"] > 1    # (an instance that stops at kk = 0 returns the all-zero slot, column 0 included)
            assert np.array_equal(xs[conv0, :, 0], x0[conv0].astype(np.float32))
        assert np.array_equal(xr2, x64) and np.array_equal(ur2, u64)
        for k in ("iters", "status", "J"):
            assert np.array_equal(st[k], st2[k]), k
        assert np.array_equal(h["stepsize"], h2["stepsize"])


def test_solve_deliver_zero_slot_and_pipelined(gpu):
    """The all-zero result of an instance that stops at kk = 0 (optcon.py:503 with kk = 0) through the direct delivery, and
    PipelinedNewton with page-locked outputs (direct delivery per sub-batch) against pageable outputs."""
    d = golden("newton_quirks.npz")
    xi = np.stack([d["a_xx_init"], d["b_xx_init"]])
    ui = np.stack([d["a_uu_init"], d["b_uu_init"]])
    keep_x, xs = _pinned((2, 6, 1000), np.float64)
    keep_u, us = _pinned((2, 2, 1000), np.float64)
    xs[...] = 5.0
    us[...] = 5.0
    with gpu.BatchedNewton(2, TT=1000, state="f64", refs_shared=True, max_iters=4) as bn:
        bn.set_weights(d["Q"], d["R"], d["QT"])
        bn.set_refs(d["xx_ref"], d["uu_ref"])
        bn.set_init(xi, ui)
        tot, _ = bn.solve_deliver((xs, us))
    assert tot == 4 and relerr(d["a_xx_star"], xs[0]) < 1e-9 and relerr(d["a_uu_star"], us[0]) < 1e-9
    assert not xs[1].any() and not us[1].any()
    n, TT = 9000, 150
    xr, ur, Q, R, QT = _short_batch(n, TT, 77)
    keep_x2, xp = _pinned((n, 6, TT), np.float32)
    keep_u2, up = _pinned((n, 2, TT), np.float64)
    with gpu.PipelinedNewton(n, n_chunks=2, TT=TT, armijo="lazy") as pn:
        pn.set_weights(Q, R, QT)
        xa, ua, sa = pn.solve(xr, ur, x_dtype=np.float32)                       # pageable outputs: staged download
        xb, ub, sb = pn.solve(xr, ur, x_dtype=np.float32, out=(xp, up))         # page-locked outputs: direct delivery
    assert np.array_equal(xa, xb) and np.array_equal(ua, ub) and np.array_equal(sa["x0"], sb["x0"])
    assert np.array_equal(sa["iters"], sb["iters"]) and np.array_equal(sa["J"], sb["J"])


def test_compact_references_gradient_method_and_fp32_delivery(gpu):
    """Two remaining combinations of the round-2 paths: the steepest-descent sweep (k_gradient_tma) with parametric references against
    host arrays, and the direct delivery from an FP32-mode context against its own staged download."""
    from aircraftoptimalcontrol_b200 import refgen
    n, TT = 4500, 150
    rng = np.random.default_rng(91)
    zf, xf = rng.uniform(1.5, 3.5, n), rng.uniform(14, 18, n)
    xr, ur = refgen.step_problem(xf, zf, tf=TT * 1e-3, TT=TT)
    Q, R, QT = refgen.weights("step")
    out = []
    for host in (True, False):
        with gpu.BatchedNewton(n, TT=TT, armijo="lazy", method="gradient", max_iters=8) as bn:
            bn.set_weights(Q, R, QT)
            if host:
                bn.set_refs(xr, ur)
            else:
                bn.set_refs_step(zf, xf, tf=TT * 1e-3)
            bn.init_guess()
            bn.solve()
            out.append((bn.result(), bn.history(), bn.stats()))
    ((xa, ua), ha, sa), ((xb, ub), hb, sb) = out
    assert np.array_equal(xa, xb) and np.array_equal(ua, ub)
    for k in ("JJ", "descent", "stepsize", "n_armijo"):
        assert np.array_equal(ha[k], hb[k]), k
    assert np.array_equal(sa["iters"], sb["iters"]) and sa["iters"].max() == 7
    keep_x, xs = _pinned((n, 6, TT), np.float32)
    keep_u, us = _pinned((n, 2, TT), np.float64)
    res = []
    for direct in (False, True):
        with gpu.BatchedNewton(n, TT=TT, armijo="lazy", precision="f32", max_iters=16) as bn:
            bn.set_weights(Q, R, QT)
            bn.set_refs_step(zf, xf, tf=TT * 1e-3)     # (FP32 mode keeps expanded references)
            bn.init_guess()
            if direct:
                bn.solve_deliver((xs, us))
                res.append((xs.copy(), us.copy()))
            else:
                bn.solve()
                res.append(bn.result_f32()[:2])
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


_BWD_KERNELS_SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import aircraftoptimalcontrol_b200 as pkg
from aircraftoptimalcontrol_b200 import refgen
out = {}
for tag, n, TT, dense, state in (("a", 333, 160, False, "f32"), ("b", 70, 90, True, "f64")):
    rng = np.random.default_rng(31)
    xr, ur = refgen.step_problem(rng.uniform(14, 18, n), rng.uniform(1.5, 3.5, n), tf=TT * 1e-3, TT=TT)
    Q, R, QT = refgen.weights("step")
    if dense:
        E = rng.normal(size=(6, 6)) * 1e-4
        Q, QT, R = Q + E @ E.T, QT + 3 * (E @ E.T), R + 1e-7 * np.array([[1.0, 0.3], [0.3, 2.0]])
    with pkg.BatchedNewton(n, TT=TT, armijo="lazy", state=state, exact_after=2, max_iters=12) as bn:
        bn.set_weights(Q, R, QT)
        bn.set_refs(xr, ur)
        bn.init_guess()
        bn.solve()
        (xs, us), h, st, (K, sig) = bn.result(), bn.history(), bn.stats(), bn.gains()[:2]
    out.update({tag + "_x": xs, tag + "_u": us, tag + "_K": K, tag + "_sig": sig, tag + "_JJ": h["JJ"], tag + "_descent": h["descent"],
                tag + "_step": h["stepsize"], tag + "_iters": st["iters"], tag + "_nreg": st["n_reg"]})
np.savez(sys.argv[2], **out)
"""


def test_backward_kernels_identical(gpu, tmp_path):
    """The three backward sweeps of the TMA path -- k_backward_cols (pipeline of 12 warp roles, the default up to one tile per SM),
    k_backward_split (two warps per tile) and k_backward_tma (one thread per instance) -- run the same expression for every number
    (optcon.py:429-464, :716-751): whole solves must agree bit for bit, gains of the last sweep included.  The choice is made per
    process (environment), hence the subprocesses.  Ragged batches (padding lanes), Gauss-Newton and exact-Hessian iterations,
    diagonal and dense weights, float and float64 state slots."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "bwd.py"
    script.write_text(_BWD_KERNELS_SCRIPT)
    res = {}
    for name, env in (("cols", {}), ("split", {"ACOC_NO_BWD_COLS": "1"}), ("tma", {"ACOC_NO_BWD_COLS": "1", "ACOC_NO_BWD_SPLIT": "1"})):
        f = tmp_path / (name + ".npz")
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, str(script), root, str(f)], check=True, env=e, timeout=300)
        res[name] = dict(np.load(f))
    assert res["cols"]["a_iters"].max() >= 6   # exact-Hessian iterations were run
    for other in ("split", "tma"):
        for k, v in res["cols"].items():
            assert np.array_equal(v, res[other][k]), (other, k)
