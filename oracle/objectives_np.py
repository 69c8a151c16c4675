"""Objective oracle.  TEST INFRASTRUCTURE ONLY.

The reference cost is an arbitrary JAX scalar function differentiated by ``jax.grad`` /
``jax.hessian`` (``/root/reference/pyNeuralEMPC/objective/jax.py:28-57``); JAX is absent
here, so the cost families that actually appear with it are restated in closed form:

* linear in u, ``sum(u * c)``                       -- ``examples/lotka_volterra/run.py:83-84``
* ``sum((u - r)**2)``                               -- ``test.py:59-60``
* diagonal quadratic tracking on states and controls (the usual MPC stage cost)

All three are instances of the separable form

    f(z) = sum_i  lin_i * z_i + quad_i * (z_i - ref_i)**2 ,   z = [states.ravel(), u.ravel()]

whose gradient is ``lin + 2 quad (z - ref)`` and whose Hessian is ``diag(2 quad)``.
Layouts follow ``objective/jax.py``: gradient ``(n,)`` = ``[d/dstates | d/du]`` (:32-41),
Hessian dense ``(n, n)`` (:43-57), structure = ``hessian != 0`` (:67-90; the Hessian is
constant so the three random probes of the reference see the same pattern).
"""
from __future__ import annotations

import numpy as np


class SeparableQuadraticObjective:
    def __init__(self, lin, quad, ref):
        self.lin = np.asarray(lin, np.float64).ravel()
        self.quad = np.asarray(quad, np.float64).ravel()
        self.ref = np.asarray(ref, np.float64).ravel()
        assert self.lin.shape == self.quad.shape == self.ref.shape

    # -- the three named families ---------------------------------------------------
    @classmethod
    def linear_in_u(cls, H, x_dim, u_dim, cost_vec):
        n = H * (x_dim + u_dim)
        lin = np.zeros(n)
        lin[H * x_dim:] = np.broadcast_to(np.asarray(cost_vec, np.float64).ravel(), (H * u_dim,))
        return cls(lin, np.zeros(n), np.zeros(n))

    @classmethod
    def control_setpoint(cls, H, x_dim, u_dim, target):
        n = H * (x_dim + u_dim)
        quad = np.zeros(n)
        quad[H * x_dim:] = 1.0
        ref = np.zeros(n)
        ref[H * x_dim:] = target
        return cls(np.zeros(n), quad, ref)

    @classmethod
    def tracking(cls, H, x_dim, u_dim, q_diag, r_diag, x_ref=None, u_ref=None):
        q = np.tile(np.asarray(q_diag, np.float64), H)
        r = np.tile(np.asarray(r_diag, np.float64), H)
        xr = np.zeros((H, x_dim)) if x_ref is None else np.broadcast_to(x_ref, (H, x_dim))
        ur = np.zeros((H, u_dim)) if u_ref is None else np.broadcast_to(u_ref, (H, u_dim))
        return cls(np.zeros(H * (x_dim + u_dim)), np.concatenate([q, r]),
                   np.concatenate([np.ravel(xr), np.ravel(ur)]))

    # -- ObjectiveFunc interface ------------------------------------------------------
    @staticmethod
    def _z(states, u):
        return np.concatenate([np.ravel(states), np.ravel(u)])

    def forward(self, states, u, p=None, tvp=None):
        z = self._z(states, u)
        return float(np.sum(self.lin * z + self.quad * (z - self.ref) ** 2))

    def gradient(self, states, u, p=None, tvp=None):
        z = self._z(states, u)
        return self.lin + 2.0 * self.quad * (z - self.ref)

    def hessian(self, states, u, p=None, tvp=None):
        return np.diag(2.0 * self.quad)

    def hessianstructure(self, H=None, model=None):
        return (np.diag(2.0 * self.quad) != 0.0).astype(np.float64)


class QuadraticFormObjective:
    """Non-separable quadratic cost, the kind of function a user hands to ``JAXObjectifFunc`` (objective/jax.py:7-57) beyond the shipped
    ones: full stage weights, a terminal weight, a control-rate penalty and a state-control cross term.  ``forward`` is the LITERAL
    definition (a loop over the horizon); ``gradient`` / ``hessian`` come from the dense matrices assembled separately, and the tests
    check the two against each other, so the matrix assembly is pinned to the definition.  Dense layouts as objective/jax.py:32-57:
    gradient (n,), hessian (n, n) over z = [states.ravel() | u.ravel()]."""

    def __init__(self, H, x_dim, u_dim, Q=None, R=None, Qf=None, S=None, x_ref=None, u_ref=None, N=None, lin=None):
        self.H, self.xd, self.ud = H, x_dim, u_dim
        z = lambda r, c: np.zeros((r, c))
        self.Q = z(x_dim, x_dim) if Q is None else np.asarray(Q, np.float64)
        self.Qf = self.Q if Qf is None else np.asarray(Qf, np.float64)
        self.R = z(u_dim, u_dim) if R is None else np.asarray(R, np.float64)
        self.S = z(u_dim, u_dim) if S is None else np.asarray(S, np.float64)
        self.N = z(x_dim, u_dim) if N is None else np.asarray(N, np.float64)
        self.xr = np.zeros((H, x_dim)) if x_ref is None else np.broadcast_to(np.asarray(x_ref, np.float64), (H, x_dim))
        self.ur = np.zeros((H, u_dim)) if u_ref is None else np.broadcast_to(np.asarray(u_ref, np.float64), (H, u_dim))
        self.lin = np.zeros(H * (x_dim + u_dim)) if lin is None else np.asarray(lin, np.float64).ravel()

    def forward(self, states, u, p=None, tvp=None):
        H = self.H
        dx, du = np.asarray(states, np.float64) - self.xr, np.asarray(u, np.float64) - self.ur
        f = float(self.lin @ np.concatenate([np.ravel(states), np.ravel(u)]))
        for t in range(H):                                    # states[t] is x_{t+1}
            W = self.Qf if t == H - 1 else self.Q
            f += dx[t] @ W @ dx[t] + du[t] @ self.R @ du[t]
            if t >= 1:
                dd = np.asarray(u, np.float64)[t] - np.asarray(u, np.float64)[t - 1]
                f += dd @ self.S @ dd
                f += 2.0 * dx[t - 1] @ self.N @ du[t]         # x_t with u_t, t = 1..H-1
        return float(f)

    def _numeric(self, states, u):
        """gradient and Hessian of ``forward`` by exact differencing of a quadratic (central differences are exact up to rounding)"""
        z0 = np.concatenate([np.ravel(states), np.ravel(u)]).astype(np.float64)
        n, H, xd, ud = z0.size, self.H, self.xd, self.ud
        f = lambda z: self.forward(z[:H * xd].reshape(H, xd), z[H * xd:].reshape(H, ud))
        E = np.eye(n)
        g = np.array([(f(z0 + E[i]) - f(z0 - E[i])) / 2.0 for i in range(n)])
        f0 = f(z0)
        Hd = np.zeros((n, n))
        for i in range(n):
            for j in range(i + 1):
                if i == j:
                    Hd[i, i] = f(z0 + E[i]) - 2.0 * f0 + f(z0 - E[i])
                else:
                    Hd[i, j] = Hd[j, i] = (f(z0 + E[i] + E[j]) - f(z0 + E[i] - E[j]) - f(z0 - E[i] + E[j]) + f(z0 - E[i] - E[j])) / 4.0
        return g, Hd

    def matrices(self):
        """dense P (n, n), q (n,), c with f(z) = 1/2 z'Pz + q'z + c, assembled term by term"""
        H, xd, ud = self.H, self.xd, self.ud
        n = H * (xd + ud)
        P = np.zeros((n, n))
        ix = lambda t: np.arange(t * xd, (t + 1) * xd)                       # states[t]
        iu = lambda t: H * xd + np.arange(t * ud, (t + 1) * ud)
        for t in range(H):
            W = self.Qf if t == H - 1 else self.Q
            P[np.ix_(ix(t), ix(t))] += W + W.T
            P[np.ix_(iu(t), iu(t))] += self.R + self.R.T
            if t >= 1:
                Ss = self.S + self.S.T
                P[np.ix_(iu(t), iu(t))] += Ss; P[np.ix_(iu(t - 1), iu(t - 1))] += Ss
                P[np.ix_(iu(t), iu(t - 1))] -= Ss; P[np.ix_(iu(t - 1), iu(t))] -= Ss
                P[np.ix_(ix(t - 1), iu(t))] += 2.0 * self.N; P[np.ix_(iu(t), ix(t - 1))] += 2.0 * self.N.T
        zr = np.concatenate([self.xr.ravel(), self.ur.ravel()])
        return P, self.lin - P @ zr, 0.5 * float(zr @ P @ zr)

    def gradient(self, states, u, p=None, tvp=None):
        P, q, _ = self.matrices()
        return P @ np.concatenate([np.ravel(states), np.ravel(u)]) + q

    def hessian(self, states, u, p=None, tvp=None):
        return self.matrices()[0]

    def hessianstructure(self, H=None, model=None):
        return (self.matrices()[0] != 0.0).astype(np.float64)
