"""IPOPT glue with every callback served by one batched device evaluation.

Mirrors ``/root/reference/pyNeuralEMPC/optimizer/ipopt.py``: ``IpoptProblem`` (:7-108) -> ``CudaIpoptProblem``
(same constructor and callbacks, plus ``jacobianstructure()`` so the solver receives the block-banded Jacobian
instead of a dense ``m x n`` matrix), ``IpoptProblemFactory`` (:111-113), ``Ipopt`` (:116-195, same kwargs and the
same three options; needs ``cyipopt``, which is an optional third-party dependency exactly as in the reference).

IPOPT calls ``objective, gradient, constraints, jacobian`` at the same iterate one after the other
(SURVEY 3.1): the first of them triggers ONE kernel launch producing all four (pinned host buffers), the others
are served from that result -- keyed on the exact bytes of ``x`` (the reference's RK4 cache keys on ``str(x)``,
rk4.py:27, which collides for large arrays).  ``hessian(x, lagrange, obj_factor)`` is a second launch.
"""
from __future__ import annotations

import numpy as np

from ..objective import CudaQuadraticFormObjective, CudaSeparableObjective
from .base import Optimizer, ProblemFactory, ProblemInterface, ProblemInterfaceHessianFree, initial_guess


class CudaIpoptProblem(ProblemInterface):
    def __init__(self, x0, objective_func, constraints, integrator, p=None, tvp=None, use_hessian=True,
                 init_x=None, init_u=None, sparse_jacobian=True):
        super().__init__(use_hessian)
        if not hasattr(integrator, "evaluator"):
            raise ValueError("CudaIpoptProblem needs a CUDA integrator (pyneuralempc_b200.integrator)")
        self.x0 = np.asarray(x0, np.float64)
        self.objective_func = objective_func
        self.constraints_list = list(constraints)
        self.integrator = integrator
        m = integrator.model
        self.x_dim, self.u_dim, self.p_dim, self.tvp_dim = m.x_dim, m.u_dim, m.p_dim, m.tvp_dim
        self.H = integrator.H
        self.p, self.tvp = p, tvp
        self.init_x, self.init_u = init_x, init_u
        self.sparse_jacobian = sparse_jacobian
        self.ev = integrator.evaluator
        self._key, self._point = None, None
        self._device_objective = isinstance(objective_func, CudaSeparableObjective)
        self._quadform = isinstance(objective_func, CudaQuadraticFormObjective)       # non-separable quadratic: own kernels + Hessian merge
        if self._device_objective or self._quadform:
            objective_func.prepare(self.H, self.x_dim, self.u_dim, p, tvp)
        if not (self._device_objective or self._quadform) and use_hessian:
            raise NotImplementedError("the Lagrangian-Hessian path needs a CudaSeparableObjective or a CudaQuadraticFormObjective")
        self._merge = None
        self._bind()
        # extra (user, host-side) constraints: rows appended after the integrator's, ipopt.py:49-50, 93-94; their Jacobian rows are dense
        self._extra_dims = [int(np.size(c.get_lower_bounds(self.H))) for c in self.constraints_list]
        self._key, self._point = None, None

    # ---- one launch per iterate -----------------------------------------------------------------------------
    def _split(self, x):                                         # ipopt.py:20-28
        nx = self.x_dim * self.H
        return x[:nx].reshape(self.H, self.x_dim), x[nx:nx + self.u_dim * self.H].reshape(self.H, self.u_dim), self.tvp, self.p

    def _bind(self):
        """The cost and the p / tvp rows are state of the integrator's ONE shared evaluator: a problem built earlier and evaluated after
        another problem (or a BatchedNMPC) used the evaluator puts its own back before it evaluates."""
        if getattr(self.ev, "bound_to", None) is self:
            return
        self.integrator._set_exogenous(self.p, self.tvp)       # model inputs that stay fixed during this solve (controller.py:65-113)
        if self._device_objective:
            self.ev.set_objective(self.objective_func.lin, self.objective_func.quad, self.objective_func.ref)
        elif self._quadform:
            self.ev.set_objective(None, None, None)            # constraint pattern only: the cost's Hessian is merged in afterwards
            self._merge = None
        self.ev.bound_to = self
        self._key, self._point = None, None

    def _at(self, x):
        self._bind()
        x = np.ascontiguousarray(x, np.float64)
        key = x.tobytes()
        if key != self._key:
            want = ("resid", "jac", "obj", "grad") if self._device_objective else ("resid", "jac")
            out = self.ev.eval_host(x, self.x0, want=want)
            self._point = {k: v[0].copy() for k, v in out.items()}
            self._key = key
        return self._point

    def objective(self, x):
        if self._device_objective:
            return float(self._at(x)["obj"]) + self.objective_func.offset
        s, u, tvp, p = self._split(np.asarray(x))
        return self.objective_func.forward(s, u, p=p, tvp=tvp)

    def gradient(self, x):
        if self._device_objective:
            return self._at(x)["grad"]
        s, u, tvp, p = self._split(np.asarray(x))
        return self.objective_func.gradient(s, u, p=p, tvp=tvp)

    def constraints(self, x):
        res = self._at(x)["resid"]
        if not self.constraints_list:
            return res
        s, u, tvp, p = self._split(np.asarray(x))
        return np.concatenate([res] + [c.forward(s, u, p=p, tvp=tvp) for c in self.constraints_list])

    def jacobianstructure(self):
        rows, cols = self.ev.jac_rows.astype(np.int64), self.ev.jac_cols.astype(np.int64)
        extra = sum(self._extra_dims)
        if extra:                                                 # user constraints may touch any variable: dense rows
            rows = np.concatenate([rows, np.repeat(np.arange(self.ev.m, self.ev.m + extra), self.ev.n)])
            cols = np.concatenate([cols, np.tile(np.arange(self.ev.n), extra)])
        return rows, cols

    def jacobian(self, x):
        vals = self._at(x)["jac"]
        if self.sparse_jacobian:
            if not self.constraints_list:
                return vals
            s, u, tvp, p = self._split(np.asarray(x))
            return np.concatenate([vals] + [np.asarray(c.jacobian(s, u, p=p, tvp=tvp), np.float64).reshape(-1)
                                            for c in self.constraints_list])
        J = np.zeros((self.ev.m, self.ev.n))                      # dense (m, n) like ipopt.py:88-96
        J[self.ev.jac_rows, self.ev.jac_cols] = vals
        if self.constraints_list:
            s, u, tvp, p = self._split(np.asarray(x))
            J = np.concatenate([J] + [c.jacobian(s, u, p=p, tvp=tvp) for c in self.constraints_list], axis=0)
        return J

    def _merge_tables(self):
        if self._merge is None:
            import torch
            self._bind()
            r, c, src, pval = self.objective_func.merge_tables(self.ev.hes_rows, self.ev.hes_cols)
            dev = self.ev.tdevice
            self._merge = (r.astype(np.int64), c.astype(np.int64), torch.as_tensor(src, device=dev), torch.as_tensor(pval, device=dev))
        return self._merge

    def hessianstructure(self):                                   # ipopt.py:55-62, analytic, same order
        if self._quadform:
            return self._merge_tables()[:2]
        return self.ev.hes_rows.astype(np.int64), self.ev.hes_cols.astype(np.int64)

    def hessian(self, x, lagrange, obj_factor):                   # ipopt.py:66-86
        self._bind()
        x = np.ascontiguousarray(x, np.float64)
        if self._quadform:
            import torch
            _, _, src, pval = self._merge_tables()
            dev = self.ev.tdevice
            kern = self.ev.eval(torch.as_tensor(x[None], device=dev), torch.as_tensor(self.x0[None], device=dev),
                                torch.as_tensor(np.asarray(lagrange, np.float64)[None, : self.ev.m], device=dev), 0.0, want=("hes",))["hes"]
            vals = self.objective_func.merge_hessian(kern, src, pval, float(obj_factor))[0].cpu().numpy()
        else:
            out = self.ev.eval_host(x, self.x0, lam=np.asarray(lagrange, np.float64)[: self.ev.m], obj_factor=float(obj_factor),
                                    want=("hes",))
            vals = out["hes"][0].copy()
        if self.constraints_list:
            # ipopt.py:75-80: sum_i lagrange_i * ctr.hessian[i], gathered on the objective + integrator pattern (ipopt.py:55-62, 84-86:
            # entries of a constraint Hessian outside that pattern are dropped by the reference as well).  A constraint without a
            # ``hessian`` method counts as linear; the reference raises AttributeError there (constraints.py:36-96 defines none).
            s, u, tvp, p = self._split(x)
            lam = np.asarray(lagrange, np.float64)
            off = self.ev.m
            for c, dim in zip(self.constraints_list, self._extra_dims):
                if hasattr(c, "hessian"):
                    Hc = np.asarray(c.hessian(s, u, p=p, tvp=tvp), np.float64).reshape(dim, self.ev.n, self.ev.n)
                    hr, hc = self.hessianstructure()
                    vals += np.einsum("i,ik->k", lam[off:off + dim], Hc[:, hr, hc])
                off += dim
        return vals

    def get_init_value(self):
        return self.x0

    def get_init_variables(self):
        return self.init_x, self.init_u

    def get_constraint_lower_bounds(self):
        return np.concatenate([np.asarray(c.get_lower_bounds(self.H), np.float64) for c in [self.integrator] + self.constraints_list])

    def get_constraint_upper_bounds(self):
        return np.concatenate([np.asarray(c.get_upper_bounds(self.H), np.float64) for c in [self.integrator] + self.constraints_list])


IpoptProblem = CudaIpoptProblem


class _DenseJacobianView(ProblemInterfaceHessianFree):
    pass


class IpoptProblemFactory(ProblemFactory):
    def _process(self):
        return CudaIpoptProblem(self.x0, self.objective, self.constraints, self.integrator, p=self.p, tvp=self.tvp,
                                use_hessian=self.use_hessian, init_x=self.init_x, init_u=self.init_u)


class Ipopt(Optimizer):
    """Same constructor as the reference (ipopt.py:117-131).  ``use_hessian`` is forwarded by ``NMPC`` here (the
    reference never forwards it, controller.py:86-105, so its default path is Hessian-free)."""

    def __init__(self, max_iteration=500, init_with_last_result=False, mu_strategy="monotone", mu_target=0,
                 mu_linear_decrease_factor=0.2, alpha_for_y="primal", obj_scaling_factor=1, nlp_scaling_max_gradient=100.0):
        super().__init__()
        self.max_iteration = max_iteration
        self.mu_strategy = mu_strategy
        self.mu_target = mu_target
        self.mu_linear_decrease_factor = mu_linear_decrease_factor
        self.alpha_for_y = alpha_for_y
        self.obj_scaling_factor = obj_scaling_factor
        self.nlp_scaling_max_gradient = nlp_scaling_max_gradient
        self.init_with_last_result = init_with_last_result
        self.prev_result = None
        self.last_info = None

    def get_factory(self):
        return IpoptProblemFactory()

    def solve(self, problem, domain_constraint):
        try:
            import cyipopt
        except ImportError as e:   # the reference fails the same way at import time (ipopt.py:4)
            raise ImportError("Ipopt.solve needs the optional third-party package 'cyipopt' (+ IPOPT); "
                              "use optimizer.Slsqp / optimizer.TrustConstr where it is not installed") from e
        x_init = initial_guess(problem, self)
        H = problem.integrator.H
        lb, ub = domain_constraint.get_lower_bounds(H), domain_constraint.get_upper_bounds(H)
        cl, cu = problem.get_constraint_lower_bounds(), problem.get_constraint_upper_bounds()
        if not problem.use_hessian:
            core = problem
            problem = ProblemInterfaceHessianFree(core)           # ipopt.py:159-160
            if getattr(core, "sparse_jacobian", False):
                problem.jacobianstructure = core.jacobianstructure
        nlp = cyipopt.Problem(n=len(x_init), m=len(cl), problem_obj=problem, lb=lb, ub=ub, cl=cl, cu=cu)
        add = getattr(nlp, "add_option", None) or nlp.addOption    # cyipopt renamed addOption -> add_option
        add("max_iter", self.max_iteration)                       # ipopt.py:172
        add("tol", 1e-1)                                          # ipopt.py:184
        add("acceptable_tol", 1e-4)                               # ipopt.py:185
        add("print_level", 0)                                     # ipopt.py:186
        x, info = nlp.solve(x_init)
        self.prev_result = x
        self.last_info = info
        return Optimizer.SUCCESS if info["status"] in (0, 1) else Optimizer.FAIL
