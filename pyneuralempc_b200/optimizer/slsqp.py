"""SciPy solvers driven by the CUDA callbacks.

``Slsqp`` / ``SlsqpProblem`` mirror ``/root/reference/pyNeuralEMPC/optimizer/slsqp.py`` (:10-198; Hessian-free,
dense Jacobian, retry loop doubling ``ftol``).  ``TrustConstr`` is an addition: SciPy's ``trust-constr`` consumes
the sparse Jacobian AND the Lagrangian Hessian, so it exercises the full hot path in a solver loop where IPOPT /
cyipopt are not installed."""
from __future__ import annotations

import warnings

import numpy as np
from scipy.optimize import Bounds, NonlinearConstraint, minimize
from scipy.sparse import coo_matrix

from ..constraints import Constraint
from .base import Optimizer, ProblemFactory, initial_guess
from .ipopt import CudaIpoptProblem


class SlsqpProblem(CudaIpoptProblem):
    def __init__(self, x0, objective_func, constraints, integrator, p=None, tvp=None, init_x=None, init_u=None):
        super().__init__(x0, objective_func, constraints, integrator, p=p, tvp=tvp, use_hessian=False,
                         init_x=init_x, init_u=init_u, sparse_jacobian=False)
        self.debug_mode = False
        self.debug_x, self.debug_u = [], []

    def set_debug(self, debug_mode):
        self.debug_mode = debug_mode

    def objective(self, x):
        if self.debug_mode:
            s, u, _, _ = self._split(np.asarray(x))
            self.debug_x.append(s.copy())
            self.debug_u.append(u.copy())
        return super().objective(x)

    def _integrator_rows_dense(self, x):
        pt = self._at(x)
        J = np.zeros((self.ev.m, self.ev.n))
        J[self.ev.jac_rows, self.ev.jac_cols] = pt["jac"]
        return J

    def constraints(self, x, eq=True):                          # slsqp.py:54-72
        # eq block = integrator residual + the user's EQ_TYPE constraints ONLY; INEQ / INTER rows go to the 'ineq' block
        s, u, tvp, p = self._split(np.asarray(x))
        rows = [self._at(x)["resid"]] if eq else []
        for c in self.constraints_list:
            t = c.get_type(self.H)
            if eq and t == Constraint.EQ_TYPE:
                rows.append(np.asarray(c.forward(s, u, p=p, tvp=tvp), np.float64))
            elif not eq and t == Constraint.INEQ_TYPE:
                rows.append(np.asarray(c.forward(s, u, p=p, tvp=tvp), np.float64))
            elif not eq and t == Constraint.INTER_TYPE:
                rows.append(c.forward(s, u, p=p, tvp=tvp) - np.asarray(c.get_lower_bounds(self.H), np.float64))
                rows.append(-c.forward(s, u, p=p, tvp=tvp) + np.asarray(c.get_upper_bounds(self.H), np.float64))
        return np.concatenate(rows, axis=0)

    def jacobian(self, x, eq=True):                             # slsqp.py:82-100
        s, u, tvp, p = self._split(np.asarray(x))
        rows = [self._integrator_rows_dense(x)] if eq else []
        for c in self.constraints_list:
            t = c.get_type(self.H)
            if eq and t == Constraint.EQ_TYPE:
                rows.append(np.asarray(c.jacobian(s, u, p=p, tvp=tvp), np.float64))
            elif not eq and t == Constraint.INEQ_TYPE:
                rows.append(np.asarray(c.jacobian(s, u, p=p, tvp=tvp), np.float64))
            elif not eq and t == Constraint.INTER_TYPE:
                rows.append(np.asarray(c.jacobian(s, u, p=p, tvp=tvp), np.float64))
                rows.append(-np.asarray(c.jacobian(s, u, p=p, tvp=tvp), np.float64))
        return np.concatenate(rows, axis=0)

    def get_constraints_dict(self):                             # slsqp.py:102-110
        result = [{"type": "eq", "fun": lambda x: self.constraints(x, eq=True), "jac": lambda x: self.jacobian(x, eq=True)}]
        if any(c.get_type(self.H) in (Constraint.INEQ_TYPE, Constraint.INTER_TYPE) for c in self.constraints_list):
            result.append({"type": "ineq", "fun": lambda x: self.constraints(x, eq=False), "jac": lambda x: self.jacobian(x, eq=False)})
        return result


class SlsqpProblemFactory(ProblemFactory):
    def _process(self):
        return SlsqpProblem(self.x0, self.objective, self.constraints, self.integrator, p=self.p, tvp=self.tvp,
                            init_x=self.init_x, init_u=self.init_u)


class Slsqp(Optimizer):
    def __init__(self, max_iteration=200, tolerance=0.5e-6, verbose=1, init_with_last_result=False, nb_max_try=15, debug=False):
        super().__init__()
        self.max_iteration = max_iteration
        self.verbose = verbose
        self.tolerance = tolerance
        self.init_with_last_result = init_with_last_result
        self.prev_result = None
        self.nb_max_try = nb_max_try
        self.debug = debug
        self.last_result = None

    def get_factory(self):
        return SlsqpProblemFactory()

    def _run(self, problem, x_init, bounds, ftol):
        return minimize(problem.objective, x_init, method="SLSQP", jac=problem.gradient, constraints=problem.get_constraints_dict(),
                        options={"maxiter": self.max_iteration, "ftol": ftol, "disp": bool(self.verbose), "iprint": self.verbose},
                        bounds=bounds)

    def solve(self, problem, domain_constraint):
        problem.set_debug(self.debug)
        x_init = initial_guess(problem, self)
        H = problem.integrator.H
        bounds = Bounds(domain_constraint.get_lower_bounds(H), domain_constraint.get_upper_bounds(H))
        res = self._run(problem, x_init, bounds, self.tolerance)
        if self.debug:
            self.constraints_val = problem.constraints(res.x)
            self.debug_x, self.debug_u = problem.debug_x, problem.debug_u
        if not res.success:
            warnings.warn("Process do not converge ! ")
            if self.debug:
                return Optimizer.FAIL
            if np.max(problem.constraints(res.x, eq=True)) > 1e-5:               # slsqp.py:184-194 (equality rows only)
                x0 = np.asarray(problem.get_init_value(), np.float64)
                cold = np.concatenate([np.tile(x0, H), np.zeros(problem.integrator.model.u_dim * H)])
                for i in range(self.nb_max_try):
                    res = self._run(problem, cold, bounds, self.tolerance * (2.0 ** i))
                    if np.max(problem.constraints(res.x, eq=True)) < 1e-5 or res.success:
                        break
            if not res.success and np.max(problem.constraints(res.x, eq=True)) > 1e-5:
                return Optimizer.FAIL
        self.prev_result = res.x
        self.last_result = res
        return Optimizer.SUCCESS


class TrustConstrProblemFactory(ProblemFactory):
    def _process(self):
        return CudaIpoptProblem(self.x0, self.objective, self.constraints, self.integrator, p=self.p, tvp=self.tvp,
                                use_hessian=True, init_x=self.init_x, init_u=self.init_u, sparse_jacobian=True)


class TrustConstr(Optimizer):
    """SciPy ``trust-constr`` on the sparse Jacobian and the sparse Lagrangian Hessian of the CUDA path."""

    def __init__(self, max_iteration=200, gtol=1e-8, xtol=1e-10, verbose=0, init_with_last_result=False):
        super().__init__()
        self.max_iteration, self.gtol, self.xtol, self.verbose = max_iteration, gtol, xtol, verbose
        self.init_with_last_result = init_with_last_result
        self.prev_result = None
        self.last_result = None

    def get_factory(self):
        return TrustConstrProblemFactory()

    def solve(self, problem, domain_constraint):
        x_init = initial_guess(problem, self)
        H = problem.integrator.H
        n = problem.ev.n
        cl, cu = problem.get_constraint_lower_bounds(), problem.get_constraint_upper_bounds()
        m = len(cl)                         # integrator rows + the rows of the extra constraints
        jr, jc = problem.jacobianstructure()
        hr, hc = problem.hessianstructure()
        off = hr != hc

        def jac(x):
            return coo_matrix((problem.jacobian(x), (jr, jc)), shape=(m, n)).tocsr()

        def lag_hess(x, v):                 # sum_i v_i * hess c_i  (objective part handled separately)
            vals = problem.hessian(x, v, 0.0)
            return coo_matrix((np.concatenate([vals, vals[off]]), (np.concatenate([hr, hc[off]]), np.concatenate([hc, hr[off]]))),
                              shape=(n, n)).tocsr()

        def obj_hess(x):
            vals = problem.hessian(x, np.zeros(m), 1.0)
            return coo_matrix((np.concatenate([vals, vals[off]]), (np.concatenate([hr, hc[off]]), np.concatenate([hc, hr[off]]))),
                              shape=(n, n)).tocsr()

        con = NonlinearConstraint(problem.constraints, cl, cu, jac=jac, hess=lag_hess)
        res = minimize(problem.objective, x_init, method="trust-constr", jac=problem.gradient, hess=obj_hess, constraints=[con],
                       bounds=Bounds(domain_constraint.get_lower_bounds(H), domain_constraint.get_upper_bounds(H)),
                       options={"maxiter": self.max_iteration, "gtol": self.gtol, "xtol": self.xtol, "verbose": self.verbose})
        self.last_result = res
        self.prev_result = res.x
        ok = res.success or (res.constr_violation < 1e-6 and res.optimality < 1e-4)
        return Optimizer.SUCCESS if ok else Optimizer.FAIL
