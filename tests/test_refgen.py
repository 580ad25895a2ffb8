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
