"""Objective layer: cost value / gradient / Hessian with the value and gradient evaluated on the GPU.

Mirrors ``/root/reference/pyNeuralEMPC/objective/base.py:4-18`` (``ObjectiveFunc``) and the call contract of
``objective/jax.py:28-65`` (``JAXObjectifFunc``): ``forward(states, u, p, tvp) -> scalar``, ``gradient -> (n,)``
ordered ``[d/dstates | d/du]``, ``hessian -> dense (n, n)``, ``hessianstructure(H, model) -> (n, n)`` 0/1.

The reference differentiates an arbitrary JAX function; the CUDA path implements the separable family

    f(z) = sum_i lin_i z_i + quad_i (z_i - ref_i)^2,       z = [states.ravel() | u.ravel()]

which covers every cost that ships with the reference (run.py:83-84 linear in u; test.py:59-60 squared set-point)
and diagonal tracking costs.  Its Hessian ``diag(2 quad)`` is constant, so the reference's numeric structure probing
(objective/jax.py:67-90) reduces to ``quad != 0``.
"""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _lib


class ObjectiveFunc:
    """Abstract cost (reference objective/base.py:4-18)."""

    def __init__(self):
        pass

    def forward(self, states, u, p=None, tvp=None):
        raise NotImplementedError("")

    def gradient(self, states, u, p=None, tvp=None):
        raise NotImplementedError("")

    def hessian(self, states, u, p=None, tvp=None):
        raise NotImplementedError("")

    def hessianstructure(self, H, model):
        raise NotImplementedError("")


class CudaSeparableObjective(ObjectiveFunc):
    def __init__(self, lin, quad, ref, device=0):
        super().__init__()
        self.lin = np.ascontiguousarray(lin, np.float64).ravel()
        self.quad = np.ascontiguousarray(quad, np.float64).ravel()
        self.ref = np.ascontiguousarray(ref, np.float64).ravel()
        if not (self.lin.shape == self.quad.shape == self.ref.shape):
            raise ValueError("lin, quad and ref must have the same length n = H*(x_dim+u_dim)")
        self.n = self.lin.shape[0]
        self.device = device
        self.offset = 0.0                          # constant term of the cost (objective/jax.py identifies one); not on the device
        self._dev = None

    def prepare(self, H, x_dim, u_dim, p=None, tvp=None):
        """hook for costs that are only known once the problem dimensions are (objective/jax.py); nothing to do here"""

    def _device_params(self):
        if self._dev is None:
            import torch
            dev = torch.device("cuda", self.device)
            self._dev = tuple(torch.as_tensor(a, dtype=torch.float64, device=dev) for a in (self.lin, self.quad, self.ref))
        return self._dev

    def _eval(self, states, u, want_grad):
        import torch
        z = np.concatenate([np.asarray(states, np.float64).reshape(-1), np.asarray(u, np.float64).reshape(-1)])
        if z.shape[0] != self.n:
            raise ValueError(f"objective built for n={self.n}, got {z.shape[0]} variables")
        lin, quad, ref = self._device_params()
        zd = torch.as_tensor(z, dtype=torch.float64, device=lin.device).reshape(1, -1)
        obj = torch.empty(1, dtype=torch.float64, device=lin.device)
        grad = torch.empty((1, self.n), dtype=torch.float64, device=lin.device) if want_grad else None
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        s = torch.cuda.current_stream(lin.device).cuda_stream
        _lib.check(_lib.load().nempc_objective_eval(_lib.F64, 1, self.n, p(zd), p(lin), p(quad), p(ref), p(obj), p(grad),
                                                    ctypes.c_void_p(s)), None, "nempc_objective_eval")
        return obj, grad

    def forward(self, states, u, p=None, tvp=None):
        return float(self._eval(states, u, False)[0].item())

    def gradient(self, states, u, p=None, tvp=None):
        g = self._eval(states, u, True)[1][0].cpu().numpy()
        return np.nan_to_num(g, nan=0.0)                      # objective/jax.py:40

    def hessian(self, states, u, p=None, tvp=None):
        return np.diag(2.0 * self.quad)

    def hessianstructure(self, H=None, model=None):
        return (np.diag(2.0 * self.quad) != 0.0).astype(np.float64)


def CudaLinearObjective(H, x_dim, u_dim, cost_vec, device=0):
    """``sum(u.ravel() * cost_vec)`` -- the shipped Lotka-Volterra cost (examples/lotka_volterra/run.py:79-87)."""
    n = H * (x_dim + u_dim)
    lin = np.zeros(n)
    lin[H * x_dim:] = np.broadcast_to(np.asarray(cost_vec, np.float64).ravel(), (H * u_dim,))
    return CudaSeparableObjective(lin, np.zeros(n), np.zeros(n), device)


def CudaSetpointObjective(H, x_dim, u_dim, target, device=0):
    """``sum((u - target)^2)`` -- the cost of the reference's test.py:55-60."""
    n = H * (x_dim + u_dim)
    quad, ref = np.zeros(n), np.zeros(n)
    quad[H * x_dim:] = 1.0
    ref[H * x_dim:] = target
    return CudaSeparableObjective(np.zeros(n), quad, ref, device)


def CudaQuadraticObjective(H, x_dim, u_dim, q_diag, r_diag, x_ref=None, u_ref=None, device=0):
    """diagonal tracking cost ``sum_t (x_t-xr_t)' Q (x_t-xr_t) + (u_t-ur_t)' R (u_t-ur_t)``."""
    q = np.tile(np.asarray(q_diag, np.float64), H)
    r = np.tile(np.asarray(r_diag, np.float64), H)
    xr = np.zeros((H, x_dim)) if x_ref is None else np.broadcast_to(x_ref, (H, x_dim))
    ur = np.zeros((H, u_dim)) if u_ref is None else np.broadcast_to(u_ref, (H, u_dim))
    return CudaSeparableObjective(np.zeros(H * (x_dim + u_dim)), np.concatenate([q, r]),
                                  np.concatenate([np.ravel(xr), np.ravel(ur)]), device)
