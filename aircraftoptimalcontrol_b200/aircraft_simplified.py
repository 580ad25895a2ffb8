"""Drop-in for the reference's `aircraft_simplified` module: same names, same positional signatures, same
return shapes/dtypes, with the arithmetic done by libacoc's CUDA kernels.

    from aircraftoptimalcontrol_b200.aircraft_simplified import Dynamics, Cost

Reference interface mirrored (paths into MohamedAtwan/AirCraftOptimalControl):
  Cost(QQt, RRt, QQT)                  aircraft_simplified.py:16-23
  Cost.stagecost(xx,uu,xx_ref,uu_ref)  :25-69   -> (ll (1,1), lx (6,1), lu (2,1), lxx, lxu, lux, luu)
  Cost.termcost(xx,xx_ref)             :71-97   -> (llT (1,1), lTx (6,1), lTxx)
  Dynamics()                           :101-124 (mutable attributes cd0,cda,cla,m,g,S,rho,J,ns,ni,dt)
  Dynamics.step(xx,uu,*args)           :263-393 -> (xxp float32 (6,), fx (6,6), fu (2,6), fxx, fuu, fux)
  Dynamics.get_initial_trajectory      :126-148
  Dynamics.get_equilibrium             :152-178
  Dynamics.dragForce / liftForce       :212-261 -> (value, gradient (6,1))   (host-side helpers)
  round_theta(th)                      :6-14
  tensorCont(P, a)                     :397-404

Every call is one kernel launch on a batch of one; the `*_batch` variants take a leading sample axis and
are what a performance-minded caller uses.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L


def round_theta(th):
    """An angle brought back into [-2*pi, 2*pi] by whole turns (aircraft_simplified.py:6-14; never called by the reference itself)."""
    two_pi = 2 * np.pi
    while abs(th) > two_pi:
        th = th + two_pi if th < -two_pi else th - two_pi
    return th


def tensorCont(P, a):
    """sum_i P[:,:,i]*a[i] (aircraft_simplified.py:397-404); host-side, the device path contracts in-kernel."""
    a = np.asarray(a, dtype=np.float64).squeeze()
    T = np.zeros(P.shape[:-1])
    for i in range(P.shape[-1]):
        T += P[:, :, i] * a[i]
    return T


class Cost:
    def __init__(self, QQt, RRt, QQT, device=0):
        self.QQt, self.RRt, self.QQT = QQt, RRt, QQT
        self.device = device

    def _w(self):
        return L.f64(self.QQt, (6, 6), "QQt"), L.f64(self.RRt, (2, 2), "RRt"), L.f64(self.QQT, (6, 6), "QQT")

    def stagecost_batch(self, xx, uu, xx_ref, uu_ref):
        """n samples: xx (n,6), uu (n,2), refs alike -> ll (n,), lx (n,6), lu (n,2)."""
        Q, R, QT = self._w()
        x, u, xr, ur = (L.f64(np.atleast_2d(a)) for a in (xx, uu, xx_ref, uu_ref))
        n = x.shape[0]
        ll, lx, lu = np.zeros(n), np.zeros((n, 6)), np.zeros((n, 2))
        L.check(L.lib().acoc_cost_batch(self.device, n, L.ptr(Q), L.ptr(R), L.ptr(QT), L.ptr(x), L.ptr(u), L.ptr(xr), L.ptr(ur),
                                        L.ptr(ll), L.ptr(lx), L.ptr(lu), None, None))
        return ll, lx, lu

    def termcost_batch(self, xx, xx_ref):
        Q, R, QT = self._w()
        x, xr = L.f64(np.atleast_2d(xx)), L.f64(np.atleast_2d(xx_ref))
        n = x.shape[0]
        llT, lTx = np.zeros(n), np.zeros((n, 6))
        L.check(L.lib().acoc_cost_batch(self.device, n, L.ptr(Q), L.ptr(R), L.ptr(QT), L.ptr(x), None, L.ptr(xr), None,
                                        None, None, None, L.ptr(llT), L.ptr(lTx)))
        return llT, lTx

    def stagecost(self, xx, uu, xx_ref, uu_ref):
        ll, lx, lu = self.stagecost_batch(np.reshape(xx, (1, 6)), np.reshape(uu, (1, 2)), np.reshape(xx_ref, (1, 6)), np.reshape(uu_ref, (1, 2)))
        ns, ni = 6, 2
        return (ll.reshape(1, 1), lx.reshape(ns, 1), lu.reshape(ni, 1), np.array(self.QQt, dtype=np.float64).copy(),
                np.zeros((ns, ni)), np.zeros((ni, ns)), np.array(self.RRt, dtype=np.float64).copy())

    def termcost(self, xx, xx_ref):
        llT, lTx = self.termcost_batch(np.reshape(xx, (1, 6)), np.reshape(xx_ref, (1, 6)))
        return llT.reshape(1, 1), lTx.reshape(6, 1), self.QQT


class Dynamics:
    def __init__(self, device=0, state="f32"):
        # aircraft_simplified.py:108-118
        self.cd0, self.cda, self.cla = 0.1716, 2.395, 3.256
        self.m, self.g, self.S, self.rho, self.J = 12, 9.81, 0.61, 1.2, 0.24
        self.ns, self.ni = 6, 2
        self.dt = 1e-3
        # attributes the reference's constructor also sets (:120-124); nothing on the path reads them
        self.Temp = None
        self.eps_init, self.eps_end, self.speedLimit = 1.5, 0.1, 480
        self.epsilon = self.eps_init
        self.device = device
        self.state = state  # "f32": next state rounded to float32 like the reference (:300); "f64": not

    @property
    def params(self):
        return np.array([self.cd0, self.cda, self.cla, self.m, self.g, self.S, self.rho, self.J, self.dt], dtype=np.float64)

    def _aero(self, xx, coeff_of_alpha):
        """0.5*rho*V^2*S*c(alpha) and its gradient w.r.t. the state (nonzero for V, theta, gamma only), c given as (c, dc/dalpha)."""
        x = np.asarray(xx, dtype=np.float64).reshape(-1)
        V, alpha = x[2], x[3] - x[5]
        c, dc = coeff_of_alpha(alpha)
        qS = 0.5 * self.rho * V ** 2 * self.S
        grad = np.zeros((self.ns, 1))
        grad[2, 0] = self.rho * V * self.S * c
        grad[3, 0] = qS * dc
        grad[5, 0] = -qS * dc
        return qS * c, grad

    def dragForce(self, xx):
        """(D, dD_x (6,1)): drag and its state gradient, aircraft_simplified.py:212-236.  Host-side helper of the reference's public
        interface (the kernels evaluate the same formula inside `step`, csrc/acoc_math.cuh)."""
        return self._aero(xx, lambda a: (self.cd0 + self.cda * a ** 2, 2 * self.cda * a))

    def liftForce(self, xx):
        """(L, dL_x (6,1)): lift and its state gradient, aircraft_simplified.py:238-261 (host-side helper, see dragForce)."""
        return self._aero(xx, lambda a: (self.cla * a, self.cla))

    def step_batch(self, xx, uu, lmbd=None):
        """n samples: xx (n,6), uu (n,2), lmbd (n,6) or None.  Returns dict(xxp (n,6) float64, A (n,6,6) = fx.T,
        B (n,6,2) = fu.T, fxx, fux) with full tensors (n,6,6,6)/(n,2,6,6) or costate-contracted (n,6,6)/(n,2,6)."""
        x, u = L.f64(np.atleast_2d(xx)), L.f64(np.atleast_2d(uu))
        n = x.shape[0]
        if x.shape != (n, 6) or u.shape != (n, 2):
            raise ValueError("xx must be (n,6) and uu (n,2)")
        lam = None if lmbd is None else L.f64(np.atleast_2d(lmbd), (n, 6), "lmbd")
        xxp, A, B = np.zeros((n, 6)), np.zeros((n, 6, 6)), np.zeros((n, 6, 2))
        fxx = np.zeros((n, 6, 6, 6) if lam is None else (n, 6, 6))
        fux = np.zeros((n, 2, 6, 6) if lam is None else (n, 2, 6))
        p = self.params
        L.check(L.lib().acoc_step_batch(self.device, n, L.ptr(p), int(self.state == "f64"), L.ptr(x), L.ptr(u), L.ptr(lam),
                                        L.ptr(xxp), L.ptr(A), L.ptr(B), L.ptr(fxx), L.ptr(fux)))
        return dict(xxp=xxp, A=A, B=B, fxx=fxx, fux=fux)

    def step(self, xx, uu, *args):
        """(xxp, fx, fu, fxx, fuu, fux) exactly as aircraft_simplified.py:393 returns them; fx/fu are the
        gradients, i.e. the TRANSPOSED Jacobians (:322, :325)."""
        lam = None
        if args:
            lam = np.asarray(args[0], dtype=np.float64).reshape(-1)[:6]
        r = self.step_batch(np.reshape(xx, (1, 6)), np.reshape(uu, (1, 2)), None if lam is None else lam.reshape(1, 6))
        xxp = r["xxp"][0].astype(np.float32) if self.state == "f32" else r["xxp"][0]
        fuu = np.zeros((2, 2, 6)) if lam is None else np.zeros((2, 2))
        return xxp, r["A"][0].T.copy(), r["B"][0].T.copy(), r["fxx"][0], fuu, r["fux"][0]

    def get_initial_trajectory(self, xx_ref, tt):
        """P-law rollout of aircraft_simplified.py:126-148 on the device, in float64 arithmetic.

        Under NumPy >= 2 the reference runs part of this loop in float32 (the float32 xxp is fed back into
        step, :145), so agreement with it is ~1e-5, not bit-level; see DESIGN.md."""
        from .batch import BatchedNewton
        xr = L.f64(xx_ref)
        TT = np.asarray(tt).shape[0]
        with BatchedNewton(1, TT=TT, device=self.device, state=self.state, refs_shared=True, params=self.params) as bn:
            bn.set_weights(np.eye(6), np.eye(2), np.eye(6))
            bn.set_refs(xr, np.zeros((2, TT)))
            bn.init_guess(5.0, 2.5)
            xx, uu = bn.iterate_at(0)
        return xx[0], uu[0]

    def get_equilibrium(self, x0, tt):
        """Trim point, aircraft_simplified.py:152-178 (host scipy, called once per problem set-up).  The reference
        aliases an INTEGER array as `uu` (:170-171), so the thrust is truncated to an int (46); kept."""
        from scipy.optimize import least_squares
        m, g, rho, Cla, Cda, Cd0, S = self.m, self.g, self.rho, self.cla, self.cda, self.cd0, self.S

        def cost(z):
            V, T, th, gam = z
            alpha = th - gam
            D = 0.5 * rho * V ** 2 * S * (Cd0 + Cda * alpha ** 2)
            Lf = 0.5 * rho * V ** 2 * S * Cla * alpha
            return [-D - m * g * np.sin(gam) + T * np.cos(alpha), Lf - m * g * np.cos(gam) + T * np.sin(alpha)]

        x_init = np.array([10, 0, 0, 0])
        sol = least_squares(cost, x_init, bounds=[(-50, 0, -np.pi, -np.pi), (50, 1000, np.pi, np.pi)])
        xx, uu = np.array(x0, dtype=np.float64).copy(), x_init
        uu[0] = sol.x[1]
        xx[2], xx[3], xx[5] = sol.x[0], sol.x[2], sol.x[3]
        return xx, uu
