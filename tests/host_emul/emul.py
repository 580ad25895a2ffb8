"""ctypes binding of tests/host_emul/libacoc_emul.so (TEST-ONLY host replay of the kernel code)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libacoc_emul.so")
_lib = None
DEFAULT_PARAMS = np.array([0.1716, 2.395, 3.256, 12.0, 9.81, 0.61, 1.2, 0.24, 1e-3])


def build(force=False):
    srcs = [os.path.join(_HERE, "emul.cpp")] + [os.path.join(_HERE, "..", "..", "aircraftoptimalcontrol_b200", "csrc", f)
                                               for f in ("acoc_math.cuh", "acoc_kernels.cuh")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-o", _SO, srcs[0]])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def step_batch(x, u, lam=None, params=DEFAULT_PARAMS, state_f64=False):
    x, u = _c(x), _c(u)
    n = x.shape[0]
    xxp, A, B = np.zeros((n, 6)), np.zeros((n, 6, 6)), np.zeros((n, 6, 2))
    if lam is None:
        fxx, fux, l = np.zeros((n, 6, 6, 6)), np.zeros((n, 2, 6, 6)), None
    else:
        fxx, fux, l = np.zeros((n, 6, 6)), np.zeros((n, 2, 6)), _c(lam)
    lib().emul_step_batch(C.c_int(n), _p(_c(params)), C.c_int(int(state_f64)), _p(x), _p(u), _p(l), _p(xxp), _p(A), _p(B), _p(fxx), _p(fux))
    return xxp, A, B, fxx, fux


def cost_batch(Q, R, QT, x, u, xr, ur):
    x = _c(x)
    n = x.shape[0]
    ll, lx, lu, llT, lTx = np.zeros(n), np.zeros((n, 6)), np.zeros((n, 2)), np.zeros(n), np.zeros((n, 6))
    lib().emul_cost_batch(C.c_int(n), _p(_c(Q)), _p(_c(R)), _p(_c(QT)), _p(x), _p(_c(u)), _p(_c(xr)), _p(_c(ur)), _p(ll), _p(lx), _p(lu), _p(llT), _p(lTx))
    return ll, lx, lu, llT, lTx


def ltv_lqr(A, B, Q, R, S, Qf, x0, q=None, r=None, qf=None):
    """time-major inputs (TT,6,6) ...; returns K (TT,2,n), P (TT,n,n), xout (TT,6), uout (TT,2), n_reg"""
    A = _c(A)
    TT = A.shape[0]
    n = 7 if q is not None else 6
    K, P, xo, uo = np.zeros((TT, 2, n)), np.zeros((TT, n, n)), np.zeros((TT, 6)), np.zeros((TT, 2))
    nreg = C.c_int(0)
    lib().emul_ltv_lqr(C.c_int(TT), _p(A), _p(_c(B)), _p(_c(Q)), _p(_c(R)), _p(_c(S)), _p(_c(Qf)), _p(_c(x0)),
                       _p(None if q is None else _c(q)), _p(None if r is None else _c(r)), _p(None if qf is None else _c(qf)),
                       _p(K), _p(P), _p(xo), _p(uo), C.byref(nreg))
    return K, P, xo, uo, nreg.value


def newton_batch(xx_ref, uu_ref, xx_init, uu_init, Q, R, QT, params=DEFAULT_PARAMS, state_f64=False, max_iters=200,
                 stepsize_0=1.0, cc=0.5, beta=0.7, armijo_maxiters=10, exact_after=8, term_cond=-1e-6, n_iters_cap=0, lazy=False, mode=0, method=0):
    """method 0: NewtonMethod.optimize, 1: GradientMethod.optimize (ACOC_METHOD_GRADIENT; `descent` is then the slope -sum|deltau|^2).
    mode 0: float64 arithmetic, float state slots when exact (the library's default); 1: float64 state slots; 2: FP32 mode"""
    xx_init, uu_init = _c(xx_init), _c(uu_init)
    N, _, TT = xx_init.shape
    xr, ur = _c(xx_ref), _c(uu_ref)
    shared = int(xr.ndim == 2)
    hJ, hD, hS = (np.zeros((N, max_iters)) for _ in range(3))
    hN = np.zeros((N, max_iters), dtype=np.int32)
    iters, status, nreg = (np.zeros(N, dtype=np.int32) for _ in range(3))
    xs, us, xl, ul, dul = np.zeros((N, 6, TT)), np.zeros((N, 2, TT)), np.zeros((N, 6, TT)), np.zeros((N, 2, TT)), np.zeros((N, 2, TT))
    Kl, sl = np.zeros((N, 12, TT)), np.zeros((N, 2, TT))
    kk = lib().emul_newton_batch(C.c_int(N), C.c_int(TT), _p(_c(params)), C.c_int(int(state_f64)), _p(_c(Q)), _p(_c(R)), _p(_c(QT)),
                                 _p(xr), _p(ur), C.c_int(shared), _p(xx_init), _p(uu_init), C.c_int(max_iters), C.c_double(stepsize_0),
                                 C.c_double(cc), C.c_double(beta), C.c_int(armijo_maxiters), C.c_int(exact_after), C.c_double(term_cond),
                                 C.c_int(n_iters_cap), C.c_int(int(lazy)), _p(hJ), _p(hD), _p(hS), _p(hN), _p(iters), _p(status),
                                 _p(xs), _p(us), _p(xl), _p(ul), _p(dul), _p(Kl), _p(sl), _p(nreg), C.c_int(int(mode) | (int(method) << 8)))
    return dict(JJ=hJ, descent=hD, stepsize=hS, n_armijo=hN, iters=iters, status=status, xx_star=xs, uu_star=us, xx_last=xl,
                uu_last=ul, deltau=dul, K=Kl.reshape(N, 2, 6, TT), sigma=sl, n_reg=nreg, kk=kk)


def rollout_batch(x0, uu, du, s, Q, R, QT, xx_ref, uu_ref, params=DEFAULT_PARAMS, state_f64=False):
    uu = _c(uu)
    N, _, TT = uu.shape
    xo, uo, J = np.zeros((N, 6, TT)), np.zeros((N, 2, TT)), np.zeros(N)
    lib().emul_rollout_batch(C.c_int(N), C.c_int(TT), _p(_c(params)), C.c_int(int(state_f64)), _p(_c(Q)), _p(_c(R)), _p(_c(QT)),
                             _p(_c(xx_ref)), _p(_c(uu_ref)), _p(_c(x0)), _p(uu), _p(_c(du)), _p(_c(s)), _p(xo), _p(uo), _p(J))
    return xo, uo, J


def lqr_tracking(xx_opt, uu_opt, Q, R, QT, delta, params=DEFAULT_PARAMS, state_f64=False):
    TT = xx_opt.shape[1]
    delta = _c(np.atleast_2d(delta))
    N = delta.shape[0]
    xr, ur, K = np.zeros((N, 6, TT)), np.zeros((N, 2, TT)), np.zeros((TT, 2, 6))
    lib().emul_lqr_tracking(C.c_int(N), C.c_int(TT), _p(_c(params)), C.c_int(int(state_f64)), _p(_c(Q)), _p(_c(R)), _p(_c(QT)),
                            _p(_c(xx_opt)), _p(_c(uu_opt)), _p(delta), _p(xr), _p(ur), _p(K))
    return xr, ur, np.moveaxis(K, 0, 2).copy()


def init_guess(xx_ref, params=DEFAULT_PARAMS, state_f64=False, kp=5.0, kt=2.5):
    xr = _c(xx_ref)
    N, _, TT = xr.shape
    xx, uu = np.zeros((N, 6, TT)), np.zeros((N, 2, TT))
    lib().emul_init_guess(C.c_int(N), C.c_int(TT), _p(_c(params)), C.c_int(int(state_f64)), _p(xr), C.c_double(kp), C.c_double(kt), _p(xx), _p(uu))
    return xx, uu


def riccati_cols_check(xx, uu, xx_ref, uu_ref, Q, R, QT, exact=True, params=DEFAULT_PARAMS):
    """riccati_matrix() vs its decomposition by columns along the backward sweep of one trajectory (6,TT)/(2,TT).
    Returns (number of differing results, steps that took the +0.5 I branch)."""
    xx = _c(xx)
    TT = xx.shape[1]
    nreg = C.c_int(0)
    f = lib().emul_riccati_cols_check
    f.restype = C.c_long
    bad = f(C.c_int(TT), _p(_c(params)), _p(_c(Q)), _p(_c(R)), _p(_c(QT)), _p(xx), _p(_c(uu)), _p(_c(xx_ref)), _p(_c(uu_ref)),
            C.c_int(int(exact)), C.byref(nreg))
    return int(bad), nreg.value
