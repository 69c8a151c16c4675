"""On-device NMPC solver as an ``Optimizer``.

The reference hands every problem to a sequential host solver: IPOPT through cyipopt (``optimizer/ipopt.py:162-189``)
or SciPy SLSQP (``optimizer/slsqp.py:172-173``), re-entering Python for every callback.  ``CudaIpm`` keeps the same
``Optimizer`` contract (``get_factory`` / ``solve(problem, domain_constraint)`` -> ``SUCCESS`` / ``FAIL``, result in
``prev_result``, ``init_with_last_result`` warm start) but runs the whole interior-point loop on the GPU
(``nempc_solve``), for one problem or -- through ``controller.BatchedNMPC`` -- for thousands at once."""
from __future__ import annotations

import numpy as np

from .base import Optimizer, ProblemFactory, initial_guess
from .ipopt import CudaIpoptProblem


class CudaIpmProblemFactory(ProblemFactory):
    def _process(self):
        return CudaIpoptProblem(self.x0, self.objective, self.constraints, self.integrator, p=self.p, tvp=self.tvp,
                                use_hessian=True, init_x=self.init_x, init_u=self.init_u, sparse_jacobian=True)


class CudaIpm(Optimizer):
    def __init__(self, max_iteration=60, tolerance=1e-6, init_with_last_result=False, **solver_options):
        super().__init__()
        self.max_iteration = max_iteration
        self.tolerance = tolerance
        self.init_with_last_result = init_with_last_result
        self.solver_options = solver_options
        self.prev_result = None
        self.last_info = None

    def get_factory(self):
        return CudaIpmProblemFactory()

    def _options(self):
        return dict(max_iter=int(self.max_iteration), tol=float(self.tolerance), **self.solver_options)

    def solve(self, problem, domain_constraint):
        if problem.constraints_list:
            raise NotImplementedError("the on-device solver handles the integrator constraints and the DomainConstraint box only; "
                                      "use optimizer.TrustConstr / Slsqp / Ipopt with extra constraints")
        if not hasattr(problem.ev, "solve"):
            raise NotImplementedError("the on-device solver's Riccati sweep assumes the one-step band; rolling-window models go through "
                                      "optimizer.TrustConstr / Slsqp / Ipopt")
        H = problem.integrator.H
        lb, ub = domain_constraint.get_lower_bounds(H), domain_constraint.get_upper_bounds(H)
        warm = (self.init_with_last_result and self.prev_result is not None) or problem.get_init_variables()[0] is not None
        z0 = initial_guess(problem, self)[None] if warm else None
        out = problem.ev.solve(np.asarray(problem.get_init_value(), np.float64)[None], lb, ub, z_init=z0, **self._options())
        status = int(out["status"][0].item())
        self.last_info = dict(status=status, iterations=int(out["iterations"][0].item()), kkt_error=float(out["kkt_error"][0].item()),
                              device_side_loop=out["used_graph"],            # the iteration loop ran as one CUDA graph (nempc_solve_stats)
                              unaccepted_steps=out["unaccepted_steps"])      # steps taken although the line search ran out of halvings
        if status != 0:
            return Optimizer.FAIL
        self.prev_result = out["z"][0].cpu().numpy()
        return Optimizer.SUCCESS

    def solve_batch(self, evaluator, X0, domain_constraint, Z_init=None):
        """B problems at once: returns the dict of ``NlpEvaluator.solve`` (CUDA tensors)."""
        H = evaluator.H
        return evaluator.solve(X0, domain_constraint.get_lower_bounds(H), domain_constraint.get_upper_bounds(H), z_init=Z_init,
                               **self._options())
