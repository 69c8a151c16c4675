"""Dense "reference-literal" restatement of the integrator constraints and the IPOPT
callback glue.  TEST INFRASTRUCTURE ONLY (also the timed CPU baseline, kind "port").

Follows the reference's *algorithm and memory shape* -- dense ``(m, n)`` Jacobian, dense
``(m, n, n)`` Hessian tensor, O(H^3) -- so that timing it is timing what the reference
does.  Written from the behaviour of (citations into ``/root/reference/pyNeuralEMPC``):

* ``integrator/discret.py:13-81``  (``kind='discrete'``)
* ``integrator/unity.py:15-81``    (``kind='unity'``)
* ``integrator/rk4.py:57-285``     (``kind='rk4'``)  -- with ``np.eye(d)`` where the
  reference hard-codes ``np.eye(3,3)`` (``rk4.py:246,255,261``), so it equals the
  reference for ``x_dim+u_dim == 3`` and generalises it otherwise.
* ``integrator/base.py:83-115``    numeric Hessian-structure probing.
* ``optimizer/ipopt.py:20-108``    ``IpoptProblem`` callbacks.

``model`` is anything with the reference ``Model`` call signature (``forward(x,u)``,
``jacobian(x,u)`` dense ``(N*x, N*d)``, ``hessian(x,u)`` dense ``(N, x, N*d, N*d)``).
Validated against the unmodified reference in ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import numpy as np

KINDS = ("discrete", "unity", "rk4")


def _prev_states(x, x0):
    # rows x_{t-1}: x0 first, then x_1..x_{H-1}            (discret.py:22, rk4.py:66)
    return np.vstack([np.reshape(x0, (1, -1)), x[:-1]])


def _interleaved_columns(N, xd, ud):
    """column order ``x_0,u_0,x_1,u_1,...`` expressed in the model's ``[all x | all u]`` layout
    (rk4.py:87-88)."""
    cols = []
    for i in range(N):
        cols.extend(range(i * xd, (i + 1) * xd))
        cols.extend(range(N * xd + i * ud, N * xd + (i + 1) * ud))
    return np.asarray(cols)


class DenseIntegrator:
    def __init__(self, model, H, kind="discrete", DT=None):
        if kind not in KINDS:
            raise ValueError(kind)
        if kind == "rk4" and DT is None:
            raise ValueError("rk4 needs DT")
        self.model, self.H, self.kind, self.DT = model, int(H), kind, DT
        self.nb_contraints = model.x_dim * self.H
        self._structure = None

    # -- per-step diagonal blocks of the dense model outputs (rk4.py:85-110) ---------
    def _model_jac_blocks(self, xp, u):
        N, xd = xp.shape
        ud = u.shape[1]
        dense = self.model.jacobian(xp, u)
        dense = dense[:, _interleaved_columns(N, xd, ud)].reshape(N, xd, N, xd + ud)
        return np.stack([dense[i, :, i, :] for i in range(N)])

    def _model_hes_blocks(self, xp, u):
        N, xd = xp.shape
        ud = u.shape[1]
        d = xd + ud
        dense = self.model.hessian(xp, u).reshape(N, xd, N * d, N * d)
        order = _interleaved_columns(N, xd, ud)
        dense = dense[:, :, :, order][:, :, order, :]
        return np.stack([dense[i, :, i * d:(i + 1) * d, i * d:(i + 1) * d] for i in range(N)])

    # -- residual ------------------------------------------------------------------
    def _rk4_stages(self, xp, u):
        f, DT = self.model.forward, self.DT
        k1 = f(xp, u)
        k2 = f(xp + k1 * DT / 2.0, u)
        k3 = f(xp + k2 * DT / 2.0, u)
        k4 = f(xp + k3 * DT, u)
        return k1, k2, k3, k4

    def forward(self, x, u, x0, p=None, tvp=None):
        assert x.ndim == 2 and u.ndim == 2, "x and u tensor must have dim 2"
        assert np.ndim(x0) == 1, "x0 shape must have dim 1"
        xp = _prev_states(x, x0)
        if self.kind == "unity":                               # unity.py:29
            pred = self.model.forward(xp, u)
        elif self.kind == "discrete":                          # discret.py:27
            pred = xp + self.model.forward(xp, u)
        else:                                                  # rk4.py:69-80
            k1, k2, k3, k4 = self._rk4_stages(xp, u)
            pred = xp + (k1 + 2 * k2 + 2 * k3 + k4) * self.DT / 6.0
        return (pred - x).reshape(-1)

    # -- Jacobian --------------------------------------------------------------------
    def _rk4_jac_chain(self, xp, u):
        """local stage Jacobians and chained ``dk_s`` (rk4.py:143-157)."""
        N, xd = xp.shape
        ud = u.shape[1]
        d = xd + ud
        DT = self.DT
        k1, k2, k3, _ = self._rk4_stages(xp, u)
        pad = lambda a: np.concatenate([a, np.zeros((N, ud, d))], axis=1)      # extend_dim, rk4.py:5-10
        eye = np.eye(d)
        j1 = self._model_jac_blocks(xp, u)
        j2 = self._model_jac_blocks(xp + k1 * DT / 2.0, u)
        j3 = self._model_jac_blocks(xp + k2 * DT / 2.0, u)
        j4 = self._model_jac_blocks(xp + k3 * DT, u)
        dk1 = j1
        dk2 = j2 @ (eye + pad(dk1) * DT / 2.0)
        dk3 = j3 @ (eye + pad(dk2) * DT / 2.0)
        dk4 = j4 @ (eye + pad(dk3) * DT)
        return (k1, k2, k3), (j1, j2, j3, j4), (dk1, dk2, dk3, dk4), pad

    def jacobian(self, x, u, x0, p=None, tvp=None):
        H, xd, ud = self.H, self.model.x_dim, self.model.u_dim
        xp = _prev_states(x, x0)
        Jd = np.zeros((xd * H, (xd + ud) * H))
        Jd[:, :xd * H] -= np.eye(xd * H)                       # d(-x_t)/dx_t   (discret.py:41)
        if self.kind in ("discrete", "unity"):
            mj = self.model.jacobian(xp, u)                    # dense, [all x | all u]
            if self.kind == "discrete":
                Jd[xd:, :xd * (H - 1)] += np.eye(xd * (H - 1))            # discret.py:52
            Jd[xd:, :xd * (H - 1)] += mj[xd:, xd:xd * H]       # discret.py:53 / unity.py:53
            Jd[:, xd * H:] += mj[:, xd * H:]                   # discret.py:56
            return Jd
        _, _, dks, _ = self._rk4_jac_chain(xp, u)
        blocks = (self.DT / 6.0) * (dks[0] + 2 * dks[1] + 2 * dks[2] + dks[3])   # rk4.py:159
        Jd[xd:, :xd * (H - 1)] += np.eye(xd * (H - 1))         # rk4.py:168
        for t in range(H):                                     # rk4.py:170-176
            rows = slice(xd * t, xd * (t + 1))
            if t > 0:
                Jd[rows, xd * (t - 1):xd * t] += blocks[t, :, :xd]
            Jd[rows, xd * H + ud * t:xd * H + ud * (t + 1)] += blocks[t, :, xd:]
        return Jd

    # -- Hessian ---------------------------------------------------------------------
    def hessian(self, x, u, x0, p=None, tvp=None):
        H, xd, ud = self.H, self.model.x_dim, self.model.u_dim
        d = xd + ud
        xp = _prev_states(x, x0)
        sx = xd * H                       # first control column
        if self.kind in ("discrete", "unity"):                 # discret.py:61-81 / unity.py:61-81
            mh = self.model.hessian(xp, u)
            mh = mh.reshape(-1, *mh.shape[2:])
            out = np.zeros_like(mh)
            lo, hi = xd, xd * H                                # model columns of x_1..x_{H-1}
            out[:, :xd * (H - 1), :xd * (H - 1)] += mh[:, lo:hi, lo:hi]
            out[:, sx:, sx:] += mh[:, sx:, sx:]
            out[:, :xd * (H - 1), sx:] += mh[:, lo:hi, sx:]
            out[:, sx:, :xd * (H - 1)] += mh[:, sx:, lo:hi]
            return out
        # rk4.py:181-285
        DT = self.DT
        (k1, k2, k3), (j1, j2, j3, j4), (dk1, dk2, dk3, _), pad = self._rk4_jac_chain(xp, u)
        eye = np.eye(d)

        def congruence(Rm, hs):            # dot_right(dot_left(T(R), h), R), rk4.py:185-189
            return np.einsum("ikj,ipkl,ilm->ipjm", Rm, hs, Rm)

        def mix(jloc, hprev):              # cross_sum, rk4.py:191-197 (only the x_dim state columns)
            return np.einsum("ipk,ikab->ipab", jloc[:, :, :xd], hprev)

        h1 = self._model_hes_blocks(xp, u)
        R1 = eye + (DT / 2.0) * pad(dk1)
        h2 = congruence(R1, self._model_hes_blocks(xp + k1 * DT / 2.0, u)) + (DT / 2.0) * mix(j2, h1)
        R2 = eye + (DT / 2.0) * pad(dk2)
        h3 = congruence(R2, self._model_hes_blocks(xp + k2 * DT / 2.0, u)) + (DT / 2.0) * mix(j3, h2)
        R3 = eye + DT * pad(dk3)
        h4 = congruence(R3, self._model_hes_blocks(xp + k3 * DT, u)) + DT * mix(j4, h3)
        blk = (h1 + 2 * h2 + 2 * h3 + h4) * (DT / 6.0)         # (H, x, d, d), rk4.py:266
        out = np.zeros((H, xd, d * H, d * H))
        for t in range(H):                                     # rk4.py:270-283
            cu = slice(sx + t * ud, sx + (t + 1) * ud)
            out[t, :, cu, cu] += blk[t, :, xd:, xd:]
            if t > 0:
                cx = slice(xd * (t - 1), xd * t)
                out[t, :, cx, cx] += blk[t, :, :xd, :xd]
                out[t, :, cx, cu] += blk[t, :, :xd, xd:]
                out[t, :, cu, cx] += blk[t, :, xd:, :xd]
        return out.reshape(-1, d * H, d * H)

    # -- numeric structure probing (integrator/base.py:83-115) ---------------------------
    def hessianstructure(self, rng=None):
        if self._structure is None:
            rng = np.random.default_rng(0) if rng is None else rng
            seen = None
            for _ in range(3):
                xs = rng.uniform(size=(self.H, self.model.x_dim))
                us = rng.uniform(size=(self.H, self.model.u_dim))
                nz = self.hessian(xs, us, xs[0]) != 0.0
                seen = nz if seen is None else (seen | nz)
            self._structure = seen.any(axis=0).astype(np.float64)
        return self._structure

    def get_lower_bounds(self, _=None):
        return [0.0] * self.nb_contraints

    def get_upper_bounds(self, _=None):
        return [0.0] * self.nb_contraints


class DenseIpoptProblem:
    """``IpoptProblem`` (optimizer/ipopt.py:7-108) over a :class:`DenseIntegrator` and an
    objective exposing ``forward/gradient/hessian/hessianstructure`` (objective/jax.py:28-65)."""

    def __init__(self, x0, objective_func, integrator):
        self.x0, self.objective_func, self.integrator = x0, objective_func, integrator
        self.x_dim, self.u_dim = integrator.model.x_dim, integrator.model.u_dim
        self.H = integrator.H

    def _split(self, z):                                       # ipopt.py:20-28
        nx = self.x_dim * self.H
        return z[:nx].reshape(self.H, self.x_dim), z[nx:nx + self.u_dim * self.H].reshape(self.H, self.u_dim)

    def objective(self, z):
        return self.objective_func.forward(*self._split(z))

    def gradient(self, z):
        return self.objective_func.gradient(*self._split(z))

    def constraints(self, z):
        s, u = self._split(z)
        return self.integrator.forward(s, u, self.x0)

    def jacobian(self, z):                                     # dense (m, n), ipopt.py:88-96
        s, u = self._split(z)
        return self.integrator.jacobian(s, u, self.x0)

    def hessianstructure(self):                                # ipopt.py:55-62
        both = (self.objective_func.hessianstructure(self.H, self.integrator.model)
                + self.integrator.hessianstructure()).astype(bool)
        return np.nonzero(np.tril(both))

    def hessian(self, z, lagrange, obj_factor):                # ipopt.py:66-86
        s, u = self._split(z)
        total = obj_factor * self.objective_func.hessian(s, u)
        per_constraint = self.integrator.hessian(s, u, self.x0)
        for lam, Hc in zip(lagrange, per_constraint):          # m dense n x n AXPYs, ipopt.py:79-80
            total = total + lam * Hc
        r, c = self.hessianstructure()
        return total[r, c]
