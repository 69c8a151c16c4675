"""``pyNeuralEMPC.objective.jax`` under its reference name: ``JAXObjectifFunc(func)`` (``objective/jax.py:7-90``).

The reference differentiates an arbitrary scalar function ``func(states, u, p, tvp)`` with JAX on every callback.  The CUDA path
evaluates the separable quadratic family ``f(z) = c + sum_i lin_i z_i + quad_i z_i^2`` (objective/__init__.py), which holds every cost
the reference ships (``examples/lotka_volterra/run.py:79-87``: ``sum(u * cost_vec)``; ``test.py:55-60``: ``sum((u - 2)^2)``).  So this
class IDENTIFIES the user's function instead of tracing it: once the problem dimensions are known it calls ``func`` on ``2n + 1`` probe
points (0, +-e_i), reads ``c``, ``lin`` and ``quad`` off the values, and then CHECKS the fit on random points -- a cost with cross
terms or higher-order terms fails that check and is refused (``NotImplementedError``) rather than silently approximated.  ``func`` only
has to accept numpy arrays (``jax.numpy`` functions do; JAX itself is not needed by this package)."""
from __future__ import annotations

import numpy as np

from . import CudaSeparableObjective


class JAXObjectifFunc(CudaSeparableObjective):
    def __init__(self, func, device=0, check_points=4, rtol=1e-4):
        self.func = func
        self.device = device
        self._dev = None
        self.n = None
        self.offset = 0.0
        self._check_points, self._rtol = check_points, rtol
        self.cached_hessian_structure = dict()                 # objective/jax.py:12

    def _call(self, z, H, x_dim, u_dim, p, tvp):
        return float(np.asarray(self.func(z[:H * x_dim].reshape(H, x_dim), z[H * x_dim:].reshape(H, u_dim), p, tvp)))

    def prepare(self, H, x_dim, u_dim, p=None, tvp=None):
        """identify ``func`` for horizon ``H`` (idempotent per shape); called by the problem classes before the first evaluation"""
        n = H * (x_dim + u_dim)
        if self.n == n and getattr(self, "_shape", None) == (H, x_dim, u_dim):
            return
        f = lambda z: self._call(z, H, x_dim, u_dim, p, tvp)
        c = f(np.zeros(n))
        lin, quad = np.zeros(n), np.zeros(n)
        e = np.zeros(n)
        for i in range(n):
            e[i] = 1.0
            fp = f(e)
            e[i] = -1.0
            fm = f(e)
            e[i] = 0.0
            quad[i] = 0.5 * (fp + fm) - c
            lin[i] = 0.5 * (fp - fm)
        scale = max(1.0, abs(c), float(np.abs(lin).max()), float(np.abs(quad).max()))
        quad[np.abs(quad) < 1e-6 * scale] = 0.0                 # float32 noise of a JAX function must not create Hessian structure
        lin[np.abs(lin) < 1e-6 * scale] = 0.0
        rng = np.random.default_rng(0)
        for _ in range(self._check_points):
            z = rng.uniform(-1.5, 1.5, n)
            want, got = f(z), c + float(lin @ z + quad @ (z * z))
            if abs(want - got) > self._rtol * max(1.0, abs(want), float(np.abs(lin) @ np.abs(z) + np.abs(quad) @ (z * z))):
                raise NotImplementedError(
                    "JAXObjectifFunc: the cost is not of the separable form c + sum_i lin_i z_i + quad_i z_i^2 that the CUDA objective kernel "
                    f"evaluates (probe fit {got!r} vs function value {want!r}); costs with cross terms go through CudaQuadraticFormObjective")
        self.lin, self.quad, self.ref, self.offset = lin, quad, np.zeros(n), c
        self.n, self._shape, self._dev = n, (H, x_dim, u_dim), None

    def _ensure(self, states, u, p, tvp):
        states, u = np.asarray(states), np.asarray(u)
        self.prepare(states.shape[0], states.shape[1], u.shape[1], p, tvp)

    def forward(self, states, u, p=None, tvp=None):
        self._ensure(states, u, p, tvp)
        return super().forward(states, u) + self.offset

    def gradient(self, states, u, p=None, tvp=None):
        self._ensure(states, u, p, tvp)
        return super().gradient(states, u)

    def hessian(self, states, u, p=None, tvp=None):
        self._ensure(states, u, p, tvp)
        return super().hessian(states, u)

    def hessianstructure(self, H=None, model=None):
        if self.n is None or (H is not None and model is not None and self._shape != (H, model.x_dim, model.u_dim)):
            if H is None or model is None:
                raise ValueError("JAXObjectifFunc.hessianstructure needs (H, model) before the first evaluation")
            self.prepare(H, model.x_dim, model.u_dim)
        return super().hessianstructure()
