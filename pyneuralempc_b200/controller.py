"""MPC controller entry point -- the contract of ``/root/reference/pyNeuralEMPC/controller.py:7-113``:
``NMPC(integrator, objective_func, constraint_list, H, DT, optimizer, use_hessian)`` and
``next(x0, p, tvp, init_x, init_u) -> (x_pred (H, x_dim), u (H, u_dim))`` or ``(None, None)`` on solver failure.
Unlike the reference (controller.py:86-105 never calls ``set_use_hessian``), ``use_hessian`` is forwarded to the
problem factory so the Lagrangian-Hessian path is actually reachable."""
from __future__ import annotations

import numpy as np

from .constraints import DomainConstraint
from .optimizer import Optimizer
from .optimizer.ipm import CudaIpm
from .optimizer.slsqp import Slsqp


class NMPC:
    def __init__(self, integrator, objective_func, constraint_list, H, DT, optimizer=None, use_hessian=True):
        self.integrator = integrator
        self.objective_func = objective_func
        self.constraint_list = list(constraint_list)
        domain = [c for c in self.constraint_list if isinstance(c, DomainConstraint)]
        if not domain:
            raise ValueError("constraint_list must contain a DomainConstraint")
        self.domain_constraint = domain[0]
        self.constraint_list.remove(self.domain_constraint)
        self.H = H
        self.DT = DT
        self.optimizer = optimizer if optimizer is not None else Slsqp(verbose=0)
        self.use_hessian = use_hessian

    def _check(self, x0, p, tvp, init_x, init_u):
        m = self.integrator.model
        assert len(x0.shape) == 1, "x0 must be a vector"
        assert x0.shape[0] == m.x_dim, "x0 dim must set according to your model !"
        if p is not None:
            assert len(p.shape) == 1, "p must be a vector"
            assert p.shape[0] == m.p_dim, "p dim must set according to your model !"
        if tvp is not None:
            assert len(tvp.shape) == 2, "tvp must be a vector"
            assert tvp.shape[1] == m.tvp_dim, "tvp dim must set according to your model !"
            assert tvp.shape[0] == self.H, "tvp first dim must set according to the horizon size !"
        assert (init_x is None) == (init_u is None), "you must give both init values"
        if init_x is not None:
            assert init_x.shape[1] == m.x_dim, f"init_x dim must have the good feature size (expected {m.x_dim})"
            assert init_u.shape[1] == m.u_dim, f"init_u dim mist have the good feature size (expected {m.u_dim})"

    def get_pb(self, x0, p=None, tvp=None, init_x=None, init_u=None):
        self._check(x0, p, tvp, init_x, init_u)
        f = self.optimizer.get_factory()
        f.set_x0(x0)
        f.set_objective(self.objective_func)
        f.set_integrator(self.integrator)
        f.set_constraints(self.constraint_list)
        if hasattr(f, "set_use_hessian") and type(f).__name__.startswith("Ipopt"):
            f.set_use_hessian(self.use_hessian)
        if init_x is not None:
            f.set_init_values(init_x, init_u)
        if tvp is not None:
            f.set_tvp(tvp)
        if p is not None:
            f.set_p(p)
        return f.getProblemInterface()

    def next(self, x0, p=None, tvp=None, init_x=None, init_u=None):
        pb = self.get_pb(x0, p, tvp, init_x, init_u)
        res = self.optimizer.solve(pb, self.domain_constraint)
        if res == Optimizer.SUCCESS:
            nx = self.integrator.model.x_dim * self.integrator.H
            sol = self.optimizer.prev_result
            return sol[:nx].reshape(self.integrator.H, -1), sol[nx:].reshape(self.integrator.H, -1)
        return None, None


class BatchedNMPC:
    """``NMPC`` for B independent problems (scenarios or closed-loop replicas) solved together on the device.

    The reference has no batch axis (one ``NMPC`` = one problem); this is its batched sibling:
    ``next(X0 (B, x_dim)) -> (x_pred (B, H, x_dim), u (B, H, u_dim), ok (B,))`` with ``ok[b] = False`` where the
    single-problem controller would have returned ``(None, None)`` (controller.py:109-113)."""

    def __init__(self, integrator, objective_func, constraint_list, H, DT, optimizer=None, warm_start=True):
        from .objective import CudaSeparableObjective
        if not isinstance(objective_func, CudaSeparableObjective):
            raise ValueError("BatchedNMPC needs a CudaSeparableObjective")
        domain = [c for c in constraint_list if isinstance(c, DomainConstraint)]
        if not domain:
            raise ValueError("constraint_list must contain a DomainConstraint")
        if len(domain) != len(constraint_list):
            raise NotImplementedError("extra constraints are not supported by the on-device solver")
        self.integrator, self.objective_func, self.domain_constraint = integrator, objective_func, domain[0]
        self.H, self.DT = H, DT
        self.optimizer = optimizer if optimizer is not None else CudaIpm()
        self.warm_start = warm_start
        self.ev = integrator.evaluator
        objective_func.prepare(H, self.ev.x_dim, self.ev.u_dim)
        self.ev.set_objective(objective_func.lin, objective_func.quad, objective_func.ref)
        self._prev = None
        self.last_info = None

    def _shifted(self, Z):
        """previous solution shifted by one step (the batch version of optimizer/ipopt.py:141-147)."""
        import torch
        H, xd, ud = self.H, self.ev.x_dim, self.ev.u_dim
        nx = xd * H
        return torch.cat([Z[:, xd:nx], Z[:, nx - xd:nx], Z[:, nx + ud:], Z[:, nx + ud * (H - 1):nx + ud * H]], dim=1)

    def next(self, X0, p=None, tvp=None):
        """``p``: (p_dim,) shared or (B, p_dim); ``tvp``: (H, tvp_dim) shared or (B, H, tvp_dim) -- when the model declares them."""
        X0 = np.asarray(X0, np.float64) if not hasattr(X0, "device") else X0
        assert X0.ndim == 2 and X0.shape[1] == self.ev.x_dim, "X0 must be (B, x_dim)"
        if self.ev.tvp_dim or self.ev.p_dim:
            self.ev.set_exogenous(tvp, p)
        else:
            assert p is None and tvp is None, "the model declares no p / tvp input"
        if self.ev.bound_to is not self:               # another problem / controller used the shared evaluator since: put our cost back
            self.ev.set_objective(self.objective_func.lin, self.objective_func.quad, self.objective_func.ref)
            self.ev.bound_to = self
        z0 = None
        if self.warm_start and self._prev is not None and self._prev.shape[0] == X0.shape[0]:
            z0 = self._shifted(self._prev)
        out = self.optimizer.solve_batch(self.ev, X0, self.domain_constraint, Z_init=z0)
        ok = out["status"] == 0
        self._prev = out["z"] if bool(ok.all()) else None
        self.last_info = out
        B, H, xd = X0.shape[0], self.H, self.ev.x_dim
        return out["z"][:, :H * xd].reshape(B, H, xd), out["z"][:, H * xd:].reshape(B, H, -1), ok
