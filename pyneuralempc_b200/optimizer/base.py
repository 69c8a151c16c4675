"""Callback protocol, problem factory and optimizer base -- the contract of
``/root/reference/pyNeuralEMPC/optimizer/base.py`` (ProblemInterfaceHessianFree :7-32, ProblemInterface :34-67,
ProblemFactory :70-126, Optimizer :128-149), unchanged so that ``NMPC`` and user code keep working."""
from __future__ import annotations

import numpy as np  # noqa: F401


class ProblemInterfaceHessianFree:
    """Wrapper exposing everything but ``hessian*`` -- how IPOPT is told to use L-BFGS (ipopt.py:159-160)."""

    def __init__(self, core):
        self.core = core

    def objective(self, x):
        return self.core.objective(x)

    def gradient(self, x):
        return self.core.gradient(x)

    def constraints(self, x):
        return self.core.constraints(x)

    def jacobian(self, x):
        return self.core.jacobian(x)

    def get_constraint_lower_bounds(self):
        return self.core.get_constraint_lower_bounds()

    def get_constraint_upper_bounds(self):
        return self.core.get_constraint_upper_bounds()

    def get_init_value(self):
        return self.core.get_init_value()


class ProblemInterface:
    def __init__(self, use_hessian: bool):
        self.use_hessian = use_hessian

    def objective(self, x):
        raise NotImplementedError("")

    def gradient(self, x):
        raise NotImplementedError("")

    def constraints(self, x):
        raise NotImplementedError("")

    def hessianstructure(self):
        raise NotImplementedError("")

    def hessian(self, x, lagrange, obj_factor):
        raise NotImplementedError("")

    def jacobian(self, x):
        raise NotImplementedError("")

    def get_constraint_lower_bounds(self):
        raise NotImplementedError("")

    def get_constraint_upper_bounds(self):
        raise NotImplementedError("")

    def get_init_value(self):
        raise NotImplementedError("")

    def get_init_variables(self):
        raise NotImplementedError("")


class ProblemFactory:
    def __init__(self):
        self.x0 = None
        self.p = None
        self.tvp = None
        self.objective = None
        self.constraints = None
        self.use_hessian = False
        self.integrator = None
        self.init_u, self.init_x = None, None

    def getProblemInterface(self) -> ProblemInterface:
        for name, val in (("x0", self.x0), ("objective", self.objective), ("constraints", self.constraints),
                          ("integrator", self.integrator)):
            if val is None:
                raise RuntimeError(f"Not ready yet ! {name} is missing")
        return self._process()

    def set_integrator(self, integrator):
        self.integrator = integrator

    def set_x0(self, x0):
        self.x0 = x0

    def set_init_values(self, init_x, init_u):
        self.init_x = init_x
        self.init_u = init_u

    def set_p(self, p):
        self.p = p

    def set_tvp(self, tvp):
        self.tvp = tvp

    def set_objective(self, obj):
        self.objective = obj

    def set_constraints(self, ctrs: list):
        self.constraints = ctrs

    def set_use_hessian(self, hessian: bool):
        self.use_hessian = hessian

    def _process(self):
        raise NotImplementedError("")


class Optimizer:
    FAIL = 1
    SUCCESS = 0

    def __init__(self):
        pass

    def get_factory(self) -> ProblemFactory:
        raise NotImplementedError("")

    def solve(self, problem: ProblemInterface, domain_constraint) -> int:
        raise NotImplementedError("")


def initial_guess(problem, optimizer):
    """Initial decision vector shared by the solvers (reference ipopt.py:141-149, slsqp.py:151-164):
    explicit ``init_x/init_u`` if the problem carries them, else the previous solution shifted by one step
    (``init_with_last_result``), else ``[x0 tiled H times | zeros]``."""
    H = problem.integrator.H
    xd, ud = problem.integrator.model.x_dim, problem.integrator.model.u_dim
    x0 = np.asarray(problem.get_init_value(), np.float64)
    init_x, init_u = problem.get_init_variables()
    if init_x is not None and init_u is not None:
        assert init_u.shape[0] == H, f"The init u values is not compliant with the MPC horizon size (receive ={init_u.shape[0]}, expected={H})"
        assert init_x.shape[0] == H, f"The init x values is not compliant with the MPC horizon size (receive ={init_x.shape[0]}, expected={H})"
        return np.concatenate([np.asarray(init_x, np.float64).reshape(-1), np.asarray(init_u, np.float64).reshape(-1)])
    prev = getattr(optimizer, "prev_result", None)
    if getattr(optimizer, "init_with_last_result", False) and prev is not None:
        nx = xd * H
        return np.concatenate([prev[xd:nx], prev[nx - xd:nx], prev[nx + ud:nx + ud * H], prev[nx + ud * (H - 1):nx + ud * H]])
    return np.concatenate([np.tile(x0, H), np.zeros(ud * H)])
