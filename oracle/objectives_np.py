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
