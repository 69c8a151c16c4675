"""Box bounds and the abstract extra-constraint API (host side; static vectors, no kernels).

Mirrors ``/root/reference/pyNeuralEMPC/constraints.py``: ``DomainConstraint`` (:3-33) replicates per-variable
bounds over the horizon in the order of the decision vector ``[states | controls]``; ``Constraint`` /
``EqualityConstraint`` / ``InequalityConstraint`` (:36-96) are the user-extensible hooks."""
from __future__ import annotations

import numpy as np


class DomainConstraint:
    def __init__(self, states_constraint: list, control_constraint: list):
        if len(states_constraint) == 0:
            raise ValueError("States constraint empty !")
        if len(control_constraint) == 0:
            raise ValueError("Control constraint empty !")
        for name, group in (("states", states_constraint), ("control", control_constraint)):
            if any(len(pair) != 2 for pair in group):
                raise ValueError(f"Your {name} constraint must be a list of bound couple  ! [(lower_bound, upper_bound), ...]")
        self.states_constraint = states_constraint
        self.control_constraint = control_constraint

    def get_dim(self, H):
        return len(self.states_constraint), len(self.control_constraint)

    def _side(self, H, k):
        return [pair[k] for pair in self.states_constraint] * H + [pair[k] for pair in self.control_constraint] * H

    def get_lower_bounds(self, H):
        return self._side(H, 0)

    def get_upper_bounds(self, H):
        return self._side(H, 1)

    def get_type(self):
        return Constraint.EQ_TYPE


class Constraint:
    EQ_TYPE = 0
    INEQ_TYPE = 1
    INTER_TYPE = 2

    def forward(self, x, u, p=None, tvp=None):
        pass

    def jacobian(self, x, u, p=None, tvp=None):
        pass

    def get_lower_bounds(self, H):
        raise NotImplementedError()

    def get_upper_bounds(self, H):
        raise NotImplementedError()

    def get_type(self, H=None):
        lo, up = np.asarray(self.get_lower_bounds(H)), np.asarray(self.get_upper_bounds(H))
        if (up == lo).all() and (lo == 0).all():
            return Constraint.EQ_TYPE
        if (up == np.inf).all() and (lo == 0).all():
            return Constraint.INEQ_TYPE
        return Constraint.INTER_TYPE


class EqualityConstraint(Constraint):
    def forward(self, x, u, p=None, tvp=None):
        raise NotImplementedError()

    def jacobian(self, x, u, p=None, tvp=None):
        raise NotImplementedError()

    def get_dim(self, H=None):
        raise NotImplementedError()

    def get_lower_bounds(self, H):
        return np.zeros(int(self.get_dim(H)))

    def get_upper_bounds(self, H):
        return np.zeros(int(self.get_dim(H)))


class InequalityConstraint(Constraint):
    def forward(self, x, u, p=None, tvp=None):
        raise NotImplementedError()

    def jacobian(self, x, u, p=None, tvp=None):
        raise NotImplementedError()

    def get_dim(self, H=None):
        raise NotImplementedError()

    def get_lower_bounds(self, H):
        return np.zeros(int(self.get_dim(H)))

    def get_upper_bounds(self, H):
        return np.ones(int(self.get_dim(H))) * np.inf
