"""Reference-trajectory generators and the batched problem configurations.

Vectorised host numpy restatements of the set-up code of the reference's scripts -- these produce the INPUTS
of the hot path (xx_ref, uu_ref, weights); they are checked bit-for-bit against the scripts' own globals in
tests/test_refgen.py.

  sigmoid_fcn                      main_newton_method.py:80-93   (= acrobatic_newton.py:83-96)
  reference_position_step          main_newton_method.py:96-114
  reference_position_acrobatic     acrobatic_newton.py:99-126
  weights("step"|"acro"|"track")   main_newton_method.py:52-63, acrobatic_newton.py:55-65, lqr_tracking.py:324-328
  step_problem / acrobatic_problem main_newton_method.py:120-142, acrobatic_newton.py:133-154
  config3 / config4 / config5      BASELINE.json configs
"""
from __future__ import annotations

import numpy as np

# Trim point returned by Dynamics.get_equilibrium(zeros(6), tt) in the reference (scipy least_squares,
# aircraft_simplified.py:152-178), captured from a live run (tests/golden/newton_*.npz: xxe, uue).
TRIM_V = float.fromhex("0x1.3731c819643f6p+3")       # 9.724826860039666
TRIM_THETA = float.fromhex("0x1.2d7f6f96a7f18p-1")   # 0.5888628837019239
TRIM_GAMMA = float.fromhex("-0x1.4cf0744667cd2p-3")  # -0.1625680049882932
TRIM_THRUST_INT = 46.0  # the reference truncates the thrust to an integer (aircraft_simplified.py:170-174)


def sigmoid_fcn(tt, slope):
    ss = 1 / (1 + np.exp((-tt) * slope))
    return ss, ss * (1 - ss)


def reference_position_step(tt, p0, pT):
    """Smooth step (main_newton_method.py:96-114); pT may be an array (N,) -> (N,TT)."""
    slope = tt.shape[0] * 1
    s, ds = sigmoid_fcn(tt - tt[-1] / 2, slope)
    pT = np.asarray(pT, dtype=np.float64)[..., None]
    return p0 + s * (pT - p0), ds * (pT - p0)


def reference_position_acrobatic(tt, p0, pT):
    """Bump (acrobatic_newton.py:99-126); pT may be an array (N,) -> (N,TT)."""
    TT = tt.shape[0]
    slope = TT * 0.1
    pT = np.asarray(pT, dtype=np.float64)[..., None]
    h = TT // 2
    up, dup = sigmoid_fcn(tt[:h] - tt[h] / 2, slope)
    dn, ddn = sigmoid_fcn(-tt[:h] + tt[h] / 2, slope)
    shape = pT.shape[:-1] + (TT,)
    temp, vv = np.zeros(shape), np.zeros(shape)
    temp[..., :h] = p0 + up * (pT - p0)
    vv[..., :h] = dup * (pT - p0)
    temp[..., h:] = p0 + dn * (pT - p0)
    vv[..., h:] = ddn * (pT - p0)
    pp = np.zeros(shape)
    N = TT
    pp[..., int(0.05 * N):int(0.50 * N)] = temp[..., :int(0.45 * N)]
    pp[..., int(0.50 * N):int(0.95 * N)] = temp[..., -int(0.45 * N):]
    return pp, vv


def weights(kind):
    if kind == "track":
        Q = np.eye(6) * 0.01
        Q[1, 1] = 10
        Q[0, 0] = 10
        return Q, np.eye(2) * 1e-5, Q.copy()
    m, g, J = 12, 9.81, 0.24
    Q = np.eye(6) * 1e-6
    Q[1, 1] = m * g * 0.01
    Q[2, 2] = 0.5 * m * 0.001
    Q[3, 3] = 0.01
    Q[4, 4] = 0.5 * J * 0.001
    R = 1e-6 * np.eye(2)
    QT = Q.copy()
    QT[1, 1] = QT[1, 1] * (20 if kind == "step" else 100)
    QT[3, 3] = QT[1, 1]
    QT[0, 0] = QT[1, 1]
    return Q, R, QT


def step_problem(xf=16, zf=2.71, tf=1, TT=1000):
    """xx_ref, uu_ref of main_newton_method.py:120-142; xf, zf scalars -> (6,TT)/(2,TT), arrays (N,) -> (N,6,TT)/(N,2,TT)."""
    tt = np.linspace(0, tf, TT)
    xf_a, zf_a = np.asarray(xf, dtype=np.float64), np.asarray(zf, dtype=np.float64)
    batched = xf_a.ndim > 0 or zf_a.ndim > 0
    xf_a, zf_a = np.broadcast_arrays(np.atleast_1d(xf_a), np.atleast_1d(zf_a))
    N = xf_a.shape[0]
    x0, z0 = 0, 0
    zz, zzd = reference_position_step(tt, z0, zf_a)
    xx_ref = np.zeros((N, 6, TT))
    uu_ref = np.zeros((N, 2, TT))
    vx = (xf_a - x0) / tf
    xx_ref[:, 0, :] = x0 + vx[:, None] * tt
    xx_ref[:, 1, :] = zz
    xx_ref[:, 2, :] = (zzd ** 2 + vx[:, None] ** 2) ** 0.5
    uu_ref[:, 0, :] = TRIM_THRUST_INT
    return (xx_ref, uu_ref) if batched else (xx_ref[0], uu_ref[0])


def acrobatic_problem(zf=2.71, xf=18, tf=1, TT=1000):
    """xx_ref, uu_ref of acrobatic_newton.py:133-154 (bump height zf may be an array)."""
    tt = np.linspace(0, tf, TT)
    zf_a = np.asarray(zf, dtype=np.float64)
    batched = zf_a.ndim > 0
    zf_a = np.atleast_1d(zf_a)
    N = zf_a.shape[0]
    zz, _ = reference_position_acrobatic(tt, 0, zf_a)
    xx_ref = np.zeros((N, 6, TT))
    uu_ref = np.zeros((N, 2, TT))
    xx_ref[:, 0, :] = 0 + ((xf - 0) / tf) * tt
    xx_ref[:, 1, :] = zz
    xx_ref[:, 2, :] = TRIM_V
    xx_ref[:, 4, :] = 0.0
    xx_ref[:, 5, :] = TRIM_GAMMA
    xx_ref[:, 3, :] = 0.0
    uu_ref[:, 0, :] = TRIM_THRUST_INT * 10
    uu_ref[:, 1, :] = -60
    return (xx_ref, uu_ref) if batched else (xx_ref[0], uu_ref[0])


# ---------------------------------------------------------------------------------------------------------
# shared time bases for the device-side generators (acoc_set_refs_generated): the per-instance references above are
#   X = 0 + vx*tt,  Z = 0 + zshape*(zf - 0),  V = ((vshape*zf)**2 + vx**2)**0.5   (step)   /   V = const (acrobatic)
# with these tables, which are computed by the very functions above (so np.exp is numpy's, as in the scripts)
# ---------------------------------------------------------------------------------------------------------
def step_bases(tf=1, TT=1000):
    """(tt, zshape, vshape) of step_problem: zshape = sigmoid, vshape = its derivative term (main_newton_method.py:96-114)."""
    tt = np.linspace(0, tf, TT)
    s, ds = sigmoid_fcn(tt - tt[-1] / 2, tt.shape[0] * 1)
    return tt, s, ds


def acrobatic_bases(tf=1, TT=1000):
    """(tt, zshape) of acrobatic_problem: the unit-height bump (acrobatic_newton.py:99-126 with pT = 1)."""
    tt = np.linspace(0, tf, TT)
    return tt, reference_position_acrobatic(tt, 0, 1.0)[0]


STEP_CONST = (np.zeros(6), np.array([TRIM_THRUST_INT, 0.0]))
ACRO_CONST = (np.array([0.0, 0.0, TRIM_V, 0.0, 0.0, TRIM_GAMMA]), np.array([TRIM_THRUST_INT * 10, -60.0]))


# ---------------------------------------------------------------------------------------------------------
# batched configurations of BASELINE.json
# ---------------------------------------------------------------------------------------------------------
def config3_deltas(n=4096, seed=1234):
    """LQR tracking: instance 0 is the shipped perturbation 0.1*ones(6), the rest U(-0.1,0.1)^6."""
    rng = np.random.default_rng(seed)
    d = rng.uniform(-0.1, 0.1, size=(n, 6))
    d[0] = 0.1
    return d


def config4_params(n=65536, seed=2024):
    """Randomised step references: zf ~ U(1.5,3.5), xf ~ U(14,18)."""
    rng = np.random.default_rng(seed)
    return rng.uniform(1.5, 3.5, n), rng.uniform(14.0, 18.0, n)


def config4(n=65536, seed=2024, lo=0, hi=None, TT=1000):
    """(xx_ref, uu_ref, Q, R, QT) for instances lo..hi of the n-instance config 4; the initial guess is the P-law
    rollout (aircraft_simplified.py:126-148) done on the device (BatchedNewton.init_guess)."""
    zf, xf = config4_params(n, seed)
    hi = n if hi is None else hi
    xr, ur = step_problem(xf[lo:hi], zf[lo:hi], TT=TT)
    return (xr, ur) + weights("step")


def config5_params(n=1048576, seed=7):
    """Acrobatic batch: x0_i = x0 + N(0, diag(.05,.05,.2,.02,.05,.02)^2), bump height zf_i ~ U(2.0,3.4)."""
    rng = np.random.default_rng(seed)
    dx0 = rng.normal(size=(n, 6)) * np.array([0.05, 0.05, 0.2, 0.02, 0.05, 0.02])
    zf = rng.uniform(2.0, 3.4, n)
    return dx0, zf


def config5(n=1048576, seed=7, lo=0, hi=None, TT=1000):
    """(xx_ref, uu_ref, dx0, Q, R, QT) for instances lo..hi.  The initial guess is the P-law rollout against each
    instance's reference started from xx_ref[:,0] + dx0 (BatchedNewton.init_guess(dx0=...)); x0 = xx_init[:,0]
    (optcon.py:398) is therefore the perturbed initial state."""
    dx0, zf = config5_params(n, seed)
    hi = n if hi is None else hi
    xr, ur = acrobatic_problem(zf[lo:hi], TT=TT)
    return (xr, ur, dx0[lo:hi]) + weights("acro")
