"""Drop-in for the reference's `optcon` module: ltv_LQR, GradientMethod (armijo_stepsize, get_update) and
NewtonMethod.optimize, plus batched variants, all running on libacoc's CUDA kernels.

Reference interface mirrored (paths into MohamedAtwan/AirCraftOptimalControl):
  ltv_LQR(AAin,BBin,QQin,RRin,SSin,QQfin,TT,x0,qq,rr,qqf) -> (KK,PP,xxout,uuout)      optcon.py:533-771
  GradientMethod / NewtonMethod (Dynamics,cost,xx_ref,uu_ref,max_iters=200,stepsize_0=1e-2,cc=0.5,beta=0.7,
                                 armijo_maxiters=20,term_cond=1e-6,visu_armijo=False)    optcon.py:11-13, :335-337
  .get_update(stepsize,uu,deltau,x0) -> (xx_temp,uu_temp)                               optcon.py:176-200
  .armijo_stepsize(uu,deltau,xx_ref,uu_ref,x0,TT,JJ,descent,JP) -> stepsize            optcon.py:204-327
  NewtonMethod.optimize(xx_init,uu_init,tf,dt) -> (xx_star,uu_star)                     optcon.py:341-529
  GradientMethod.optimize(xx_init,uu_init,tf,dt) -> (xx_star,uu_star)                   optcon.py:27-174

Differences that are deliberate and documented in DESIGN.md: shape errors raise ValueError instead of
print()+exit() (optcon.py:585-596); the matplotlib figures (optcon.py:167-174, :280-325, :513-528) are drawn only
when matplotlib is importable (the data behind the visu_armijo figure is always available: `last_armijo_sweep`);
GradientMethod.optimize is broken in the reference (optcon.py:125 passes 8 arguments to the 9-argument armijo_stepsize
and raises TypeError) -- here it runs, with the missing JP = JJ[kk] supplied and the slope -descent[kk] handed to the
line search (include/acoc.h, ACOC_METHOD_GRADIENT; the oracle applies the same repair to the live reference).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .batch import BatchedNewton


def _stack(M, TT, rows, cols, name):
    """Reference stacks are (rows, cols[, TT]); a constant matrix is repeated along t (optcon.py:552-608)."""
    M = np.asarray(M, dtype=np.float64)
    if M.ndim == 2:
        M = M[:, :, None]
    if M.ndim != 3 or M.shape[0] != rows or M.shape[1] != cols:
        raise ValueError("%s has shape %s, expected (%d,%d[,TT])" % (name, M.shape, rows, cols))
    if M.shape[2] == 1:
        M = M.repeat(TT, axis=2)
    if M.shape[2] < TT:
        raise ValueError("%s has %d time samples, need 1 or >= %d" % (name, M.shape[2], TT))
    return np.ascontiguousarray(np.moveaxis(M[:, :, :TT], 2, 0))


def _lq_problem(AAin, BBin, QQin, RRin, SSin, QQfin, TT, x0, qq, rr, qqf):
    """One problem of ltv_LQR in the library's time-major layout: (A, B, Q, R, S, Qf, x0, q, r, qf), the last three None in the
    non-augmented branch (optcon.py:614)."""
    ns, ni = 6, 2
    A = _stack(AAin, TT, ns, ns, "AAin")
    B = _stack(BBin, TT, ns, ni, "BBin")
    Q = _stack(QQin, TT, ns, ns, "QQin")   # "Matrix Q does not match number of states" (optcon.py:585)
    R = _stack(RRin, TT, ni, ni, "RRin")
    S = _stack(SSin, TT, ni, ns, "SSin")
    Qf = L.f64(QQfin, (ns, ns), "QQfin")
    x0 = L.f64(np.asarray(x0, dtype=np.float64).reshape(-1), (ns,), "x0")
    aug = (qq is not None) or (rr is not None) or (qqf is not None)   # optcon.py:614
    q = r = qf = None
    if aug:
        def aff(v, n, name):
            if v is None:
                return np.zeros((TT, n))
            v = np.asarray(v, dtype=np.float64)
            if v.ndim == 1:
                v = v[:, None]
            if v.shape[0] != n:
                raise ValueError("%s does not match dimension %d" % (name, n))   # optcon.py:642-647
            if v.shape[1] == 1:
                v = v.repeat(TT, axis=1)
            return np.ascontiguousarray(v[:, :TT].T)
        q, r = aff(qq, ns, "qq"), aff(rr, ni, "rr")
        qf = np.zeros(ns) if qqf is None else L.f64(np.asarray(qqf, dtype=np.float64).reshape(-1), (ns,), "qqf")
    return A, B, Q, R, S, Qf, x0, q, r, qf


def _lq_solve(problems, TT, device):
    """acoc_ltv_lqr on nb problems that are all augmented or all plain -> K (nb,TT,2,n), P (nb,TT,n,n), x (nb,TT,6), u (nb,TT,2), n_reg (nb,)."""
    nb, ns, ni = len(problems), 6, 2
    aug = problems[0][7] is not None
    if any((p[7] is not None) != aug for p in problems):
        raise ValueError("the problems of one batch must all have affine terms (qq/rr/qqf) or none")
    n = ns + 1 if aug else ns
    cat = [np.ascontiguousarray(np.stack([p[k] for p in problems])) if (k < 7 or aug) else None for k in range(10)]
    K, P = np.zeros((nb, TT, ni, n)), np.zeros((nb, TT, n, n))
    xo, uo = np.zeros((nb, TT, ns)), np.zeros((nb, TT, ni))
    nreg = np.zeros(nb, dtype=np.int32)
    L.check(L.lib().acoc_ltv_lqr(device, nb, TT, *[L.ptr(a) for a in cat], L.ptr(K), L.ptr(P), L.ptr(xo), L.ptr(uo), L.ptr(nreg)))
    return K, P, xo, uo, nreg


def ltv_LQR(AAin, BBin, QQin, RRin, SSin, QQfin, TT, x0, qq=None, rr=None, qqf=None, device=0, return_nreg=False):
    """LQR for an LTV system with (time-varying) affine cost, optcon.py:533-771, for ns = 6, ni = 2."""
    K, P, xo, uo, nreg = _lq_solve([_lq_problem(AAin, BBin, QQin, RRin, SSin, QQfin, TT, x0, qq, rr, qqf)], TT, device)
    out = (np.moveaxis(K[0], 0, 2).copy(), np.moveaxis(P[0], 0, 2).copy(), xo[0].T.copy(), uo[0].T.copy())
    return out + (int(nreg[0]),) if return_nreg else out


def ltv_LQR_batch(problems, TT, device=0):
    """nb independent ltv_LQR problems in ONE launch (one thread per problem, acoc_ltv_lqr with nb > 1).  `problems` is a sequence of
    argument tuples (AAin, BBin, QQin, RRin, SSin, QQfin, x0[, qq, rr, qqf]) with ltv_LQR's shapes; all of them with affine terms or
    none.  Returns (KK (nb,2,n,TT), PP (nb,n,n,TT), xxout (nb,6,TT), uuout (nb,2,TT), n_reg (nb,)), problem b equal to what
    ltv_LQR(*problems[b]) returns."""
    if len(problems) == 0:
        raise ValueError("no problem given")
    prob = []
    for pr in problems:
        pr = tuple(pr)
        if len(pr) not in (7, 10):
            raise ValueError("a problem is (AAin, BBin, QQin, RRin, SSin, QQfin, x0[, qq, rr, qqf])")
        prob.append(_lq_problem(*pr[:6], TT, pr[6], *(pr[7:] if len(pr) == 10 else (None, None, None))))
    K, P, xo, uo, nreg = _lq_solve(prob, TT, device)
    return (np.ascontiguousarray(np.moveaxis(K, 1, 3)), np.ascontiguousarray(np.moveaxis(P, 1, 3)), np.ascontiguousarray(np.moveaxis(xo, 1, 2)),
            np.ascontiguousarray(np.moveaxis(uo, 1, 2)), nreg)


class GradientMethod:
    """Steepest-descent method and base class of NewtonMethod: the problem, the two line-search helpers (optcon.py:7-25, :176-327),
    optimize (optcon.py:27-174) and the batched optimize_batch."""

    def __init__(self, Dynamics, cost, xx_ref, uu_ref, max_iters=200, stepsize_0=1e-2, cc=0.5, beta=0.7,
                 armijo_maxiters=20, term_cond=1e-6, visu_armijo=False):
        self.dyn, self.cst = Dynamics, cost
        self.ns, self.ni = self.dyn.ns, self.dyn.ni
        self.xx_ref, self.uu_ref = xx_ref, uu_ref
        self.max_iters = max_iters
        self.stepsize_0 = stepsize_0
        self.cc, self.beta = cc, beta
        self.term_cond = term_cond
        self.armijo_maxiters = armijo_maxiters
        self.visu_armijo = visu_armijo  # optcon.py:280-325: armijo_stepsize then also evaluates the cost sweep behind the figure

    # -- helpers ----------------------------------------------------------------------------------------
    def _state(self):
        return getattr(self.dyn, "state", "f32")

    _method = "gradient"   # which optimize() loop the batched driver runs (NewtonMethod overrides)

    def _solver(self, n, TT, xx_ref, uu_ref, armijo="speculative"):
        xr, ur = np.asarray(xx_ref, dtype=np.float64), np.asarray(uu_ref, dtype=np.float64)
        bn = BatchedNewton(n, TT=TT, device=getattr(self.dyn, "device", 0), state=self._state(), refs_shared=(xr.ndim == 2),
                           armijo=armijo, params=self.dyn.params, max_iters=self.max_iters, stepsize_0=self.stepsize_0,
                           cc=self.cc, beta=self.beta, armijo_maxiters=self.armijo_maxiters, method=self._method)
        bn.set_weights(self.cst.QQt, self.cst.RRt, self.cst.QQT)
        bn.set_refs(xr, ur)
        return bn

    def _exhausted_step(self):
        """stepsize_0*beta**armijo_maxiters built like optcon.py:270: what the search returns, untested, when every
        candidate fails (:327).  It is smaller than every tested candidate."""
        s = self.stepsize_0
        for _ in range(self.armijo_maxiters):
            s = self.beta * s
        return s

    def get_update(self, stepsize, uu, deltau, x0):
        """Roll out u + stepsize*deltau from x0 (optcon.py:176-200) -> (xx_temp (6,TT), uu_temp (2,TT))."""
        uu, deltau = np.asarray(uu, dtype=np.float64), np.asarray(deltau, dtype=np.float64)
        TT = uu.shape[1]
        xi = np.zeros((1, 6, TT))
        xi[0, :, 0] = np.asarray(x0, dtype=np.float64).reshape(-1)
        with self._solver(1, TT, self.xx_ref, self.uu_ref) as bn:
            bn.set_init(xi, uu[None])
            bn.set_deltau(deltau[None])
            bn.update(float(stepsize))
            xx_t, uu_t = bn.iterate_at(0)
        return xx_t[0], uu_t[0]

    def armijo_stepsize(self, uu, deltau, xx_ref, uu_ref, x0, TT, JJ, descent, JP):
        """Backtracking line search of optcon.py:204-327.  All armijo_maxiters candidates are rolled out
        concurrently; the returned step is the one the sequential search returns (including the untested
        stepsize_0*beta**armijo_maxiters when every candidate fails, :327)."""
        uu, deltau = np.asarray(uu, dtype=np.float64), np.asarray(deltau, dtype=np.float64)
        xi = np.zeros((1, 6, TT))
        xi[0, :, 0] = np.asarray(x0, dtype=np.float64).reshape(-1)
        with self._solver(1, TT, xx_ref, uu_ref) as bn:
            bn.set_init(xi, uu[None])
            bn.set_deltau(deltau[None])
            bn.set_scalars(J=np.array([float(np.asarray(JP).squeeze())]), descent=np.array([float(np.asarray(descent).squeeze())]))
            s, costs = bn.armijo()
            if self.visu_armijo:   # the data of the reference's figure (optcon.py:282-296): cost along deltau on linspace(0, stepsize_0, armijo_maxiters)
                steps = np.linspace(0, self.stepsize_0, int(self.armijo_maxiters))
                self.last_armijo_sweep = dict(steps=steps, costs=bn.armijo_sweep(steps)[0], JP=float(np.asarray(JP).squeeze()),
                                              descent=float(np.asarray(descent).squeeze()))
        self.last_armijo_costs = costs[0]
        if self.visu_armijo:
            self._plot_armijo(s[0])
        if s[0] != self._exhausted_step():
            print('Armijo stepsize = {}'.format(s[0]))   # optcon.py:272 prints only when a candidate is accepted
        return float(s[0])

    def _plot_armijo(self, stepsize):
        """The visu_armijo figure (optcon.py:298-325) from last_armijo_sweep; silently skipped without matplotlib."""
        try:
            import matplotlib.pyplot as plt
        except Exception:
            return
        sw = self.last_armijo_sweep
        plt.figure(1)
        plt.clf()
        plt.plot(sw["steps"], sw["costs"], color='g', label='$J(\\mathbf{u}^k - stepsize*d^k)$')
        plt.plot(sw["steps"], sw["JP"] + sw["descent"] * sw["steps"], color='r', label='$J(\\mathbf{u}^k) - stepsize*\\nabla J^{\\top} d^k$')
        plt.plot(sw["steps"], sw["JP"] + self.cc * sw["descent"] * sw["steps"], color='g', linestyle='dashed',
                 label='$J(\\mathbf{u}^k) - stepsize*c*\\nabla J^{\\top} d^k$')
        plt.scatter(stepsize, np.interp(stepsize, sw["steps"], sw["costs"]), marker='*')
        plt.grid()
        plt.xlabel('stepsize')
        plt.legend()
        plt.draw()
        plt.show(block=False)

    def _plot_history(self, descent, JJ):
        """The two closing figures of optimize (optcon.py:167-174 / :513-528); silently skipped without matplotlib."""
        try:
            import matplotlib.pyplot as plt
        except Exception:
            return
        k = np.arange(len(JJ))
        plt.figure('descent direction')
        plt.plot(k, np.abs(descent))
        plt.xlabel('$k$')
        plt.ylabel('||$\\nabla J(\\mathbf{u}^k)||$')
        plt.yscale('log')
        plt.grid()
        plt.show(block=False)
        plt.figure('cost')
        plt.plot(k, JJ)
        plt.xlabel('$k$')
        plt.ylabel('$J(\\mathbf{u}^k)$')
        plt.yscale('log')
        plt.grid()
        plt.show(block=False)

    def optimize_batch(self, xx_init, uu_init, tf, dt, xx_ref=None, uu_ref=None, armijo="speculative", return_solver=False):
        """N instances at once: xx_init (N,6,TT), uu_init (N,2,TT); references default to the constructor's
        (shared (6,TT) or per-instance (N,6,TT)).  Returns (xx_star (N,6,TT), uu_star (N,2,TT), info) where info has
        per-instance histories JJ/descent/stepsize/n_armijo (N,max_iters) and iters/status/J/last_descent/n_reg (N,).
        `descent` follows the reference's sign convention of the method: sum g'deltau (negative) for NewtonMethod
        (optcon.py:474-477), sum |deltau|^2 (positive) for GradientMethod (optcon.py:118)."""
        xx_init, uu_init = np.asarray(xx_init, dtype=np.float64), np.asarray(uu_init, dtype=np.float64)
        if xx_init.ndim != 3 or xx_init.shape[1] != 6:
            raise ValueError("xx_init must be (N,6,TT)")
        N, _, TT = xx_init.shape
        if TT != int(tf / dt):
            raise ValueError("xx_init has %d samples but int(tf/dt) = %d" % (TT, int(tf / dt)))
        if abs(float(dt) - float(self.dyn.dt)) > 0:
            raise ValueError("dt (%g) differs from Dynamics.dt (%g); the reference uses dyn.dt inside step" % (dt, self.dyn.dt))
        bn = self._solver(N, TT, self.xx_ref if xx_ref is None else xx_ref, self.uu_ref if uu_ref is None else uu_ref, armijo)
        try:
            bn.set_init(xx_init, uu_init)
            total = bn.solve()
            xs, us = bn.result()
            info = bn.history()
            st = bn.stats()
            sign = -1.0 if self._method == "gradient" else 1.0   # the driver keeps the slope -descent in gradient mode
            info["descent"] = sign * info["descent"]
            info.update(iters=st["iters"], status=st["status"], J=st["J"], last_descent=sign * st["descent"], n_reg=st["n_reg"], total_iters=total)
        except Exception:
            bn.close()
            raise
        if return_solver:
            return xs, us, info, bn
        bn.close()
        return xs, us, info

    def optimize(self, xx_init, uu_init, tf, dt):
        """Steepest descent of optcon.py:27-174: (xx_star, uu_star) = iterate kk-1 with uu_star[:,-1] = uu_star[:,-2] (:163-165).
        Prints the lines of :78, :272, :155.  See the module docstring for the repaired line-search call."""
        xs, us, info = self.optimize_batch(np.asarray(xx_init, dtype=np.float64)[None], np.asarray(uu_init, dtype=np.float64)[None], tf, dt)
        print('-*-*-*-*-*-')
        k = int(info["iters"][0])
        self.history = {key: info[key][0, :k].copy() for key in ("JJ", "descent", "stepsize", "n_armijo")}
        self.history["iters"] = k
        for kk in range(k):
            if info["stepsize"][0, kk] != self._exhausted_step():
                print('Armijo stepsize = {}'.format(info["stepsize"][0, kk]))
            print('Iter = {}\t Descent = {}\t Cost = {}'.format(kk, info["descent"][0, kk], info["JJ"][0, kk]))
        self._plot_history(self.history["descent"], self.history["JJ"])
        return xs[0], us[0]


class NewtonMethod(GradientMethod):
    """Regularized Newton method of optcon.py:329-529 on the GPU."""

    _method = "newton"

    def optimize(self, xx_init, uu_init, tf, dt):
        """(xx_star, uu_star) = iterate kk-1 with uu_star[:,-1] = uu_star[:,-2], optcon.py:499-505.  Prints the
        per-iteration lines of optcon.py:410, :272, :497-498 so the cost-descent history is observable the same way."""
        xs, us, info = self.optimize_batch(np.asarray(xx_init, dtype=np.float64)[None], np.asarray(uu_init, dtype=np.float64)[None], tf, dt)
        print('-*-*-*-*-*-')
        k = int(info["iters"][0])
        self.history = {key: info[key][0, :k].copy() for key in ("JJ", "descent", "stepsize", "n_armijo")}
        self.history["iters"] = k
        for kk in range(k):
            print("Augmented term!")                                    # optcon.py:616
            if info["stepsize"][0, kk] != self._exhausted_step():
                print('Armijo stepsize = {}'.format(info["stepsize"][0, kk]))
            print('Iter = {}\t Descent = {}\t Cost = {}'.format(kk, info["descent"][0, kk], info["JJ"][0, kk]))
            print('term = {}'.format(-1e-6))
        self._plot_history(self.history["descent"], self.history["JJ"])
        return xs[0], us[0]
