"""Host reference generators (N2) reproduce the set-up of the reference's scripts bit for bit."""
import numpy as np

from aircraftoptimalcontrol_b200 import refgen
from tests.util import golden


def test_step_problem_matches_script():
    d = golden("newton_step_f32.npz")
    xr, ur = refgen.step_problem()
    assert np.array_equal(xr, d["xx_ref"]) and np.array_equal(ur, d["uu_ref"])
    Q, R, QT = refgen.weights("step")
    assert np.array_equal(Q, d["Q"]) and np.array_equal(R, d["R"]) and np.array_equal(QT, d["QT"])
    assert xr[1, 0] != 0 and abs(xr[1, 0]) < 1e-200  # the 1.93e-217 of SURVEY.md 8(d)


def test_acrobatic_problem_matches_script():
    d = golden("newton_acro_f32.npz")
    xr, ur = refgen.acrobatic_problem()
    assert np.array_equal(xr, d["xx_ref"]) and np.array_equal(ur, d["uu_ref"])
    Q, R, QT = refgen.weights("acro")
    assert np.array_equal(Q, d["Q"]) and np.array_equal(QT, d["QT"])
    assert np.array_equal(d["xxe"][[2, 3, 5]], [refgen.TRIM_V, refgen.TRIM_THETA, refgen.TRIM_GAMMA])


def test_batched_generators_are_consistent():
    zf, xf = refgen.config4_params(64)
    xr, ur, Q, R, QT = refgen.config4(64, lo=8, hi=12)
    for j, i in enumerate(range(8, 12)):
        a, b = refgen.step_problem(float(xf[i]), float(zf[i]))
        assert np.array_equal(a, xr[j]) and np.array_equal(b, ur[j])
    xr5, ur5, dx0, *_ = refgen.config5(32, lo=0, hi=4)
    _, zf5 = refgen.config5_params(32)
    a, b = refgen.acrobatic_problem(float(zf5[2]))
    assert np.array_equal(a, xr5[2]) and np.array_equal(b, ur5[2]) and dx0.shape == (4, 6)
    d = refgen.config3_deltas(16)
    assert np.all(d[0] == 0.1) and np.all(np.abs(d) <= 0.1)


def test_tracking_weights():
    d = golden("lqr_tracking.npz")
    Q, R, QT = refgen.weights("track")
    assert np.array_equal(Q, d["Q"]) and np.array_equal(R, d["R"]) and np.array_equal(QT, d["QT"])


def test_generator_formula_of_the_device_kernel_matches_the_scripts():
    """acoc_set_refs_generated builds X = 0 + vx*tt, Z = 0 + zshape*(zf - 0), V = sqrt((vshape*zf)^2 + vx^2) from refgen's shared time
    bases; with this numpy those expressions are bit-identical to the scripts' arrays (`**2` is a multiply, `**0.5` a sqrt)."""
    rng = np.random.default_rng(0)
    n = 512
    zf, xf = rng.uniform(1.5, 3.5, n), rng.uniform(14, 18, n)
    xr, ur = refgen.step_problem(xf, zf)
    tt, s, ds = refgen.step_bases()
    vx = (xf - 0) / 1
    zd = ds * (zf[:, None] - 0.0)
    assert np.array_equal(0.0 + vx[:, None] * tt, xr[:, 0]) and np.array_equal(0.0 + s * (zf[:, None] - 0.0), xr[:, 1])
    assert np.array_equal(np.sqrt(zd * zd + (vx * vx)[:, None]), xr[:, 2]) and not xr[:, 3:].any()
    assert np.all(ur[:, 0] == refgen.STEP_CONST[1][0]) and not ur[:, 1].any()
    zf = rng.uniform(2.0, 3.4, n)
    xr, ur = refgen.acrobatic_problem(zf)
    tt, bump = refgen.acrobatic_bases()
    xc, uc = refgen.ACRO_CONST
    assert np.array_equal(0.0 + bump * (zf[:, None] - 0.0), xr[:, 1]) and np.array_equal(np.broadcast_to(0.0 + 18.0 * tt, (n, 1000)), xr[:, 0])
    for c in (2, 3, 4, 5):
        assert np.all(xr[:, c] == xc[c])
    assert np.all(ur[:, 0] == uc[0]) and np.all(ur[:, 1] == uc[1])


def test_aero_helpers_and_equilibrium_match_the_reference():
    """Host-side helpers of the drop-in Dynamics that the reference exposes next to `step`: dragForce / liftForce
    (aircraft_simplified.py:212-261), get_equilibrium with its int-truncated thrust (:152-178), round_theta (:6-14), against values of
    the live reference (tests/golden/aero_kat.npz, oracle/gen_golden.py::gen_aero_kat)."""
    import os
    from aircraftoptimalcontrol_b200.aircraft_simplified import Dynamics, round_theta
    from tests.util import GOLDEN
    g = np.load(os.path.join(GOLDEN, "aero_kat.npz"))
    d = Dynamics()
    for k, x in enumerate(g["x"]):
        D, dD = d.dragForce(x)
        Lf, dL = d.liftForce(x)
        assert dD.shape == (6, 1) and dL.shape == (6, 1)
        assert abs(D - g["D"][k]) <= 1e-13 * abs(g["D"][k]) and abs(Lf - g["L"][k]) <= 1e-13 * max(abs(g["L"][k]), 1e-300)
        assert np.allclose(dD, g["dD"][k], rtol=1e-13, atol=0) and np.allclose(dL, g["dL"][k], rtol=1e-13, atol=0)
    assert np.allclose([round_theta(t) for t in g["th"]], g["th_rounded"], rtol=0, atol=1e-12)
    xe, ue = d.get_equilibrium(np.array([0.0, 0.0, 16.0, 0.0, 0.0, 0.0]), np.linspace(0, 1, 1000))
    assert np.allclose(xe, g["xe"], rtol=1e-9, atol=1e-12) and np.array_equal(np.asarray(ue, dtype=np.float64), g["ue"]) and ue[0] == 46
    assert (d.Temp, d.eps_init, d.eps_end, d.speedLimit, d.epsilon) == (None, 1.5, 0.1, 480, 1.5)   # constructor attributes (:120-124)
