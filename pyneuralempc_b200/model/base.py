"""``Model`` base class (reference ``model/base.py:3-18``).

The reference's ``Integrator.__init__`` insists on ``isinstance(model, pyNeuralEMPC.model.base.Model)`` (``integrator/base.py:16-17``).
So that a ``CudaMLPModel`` can be handed to a reference integrator when both packages live in one process, ``adopt_reference_base()``
re-parents this ``Model`` onto the reference's class as soon as ``pyNeuralEMPC.model.base`` is loaded (it is never imported from here:
that would pull in TensorFlow and JAX)."""
from __future__ import annotations

import sys


class _Root:
    """plain Python base so that ``Model.__bases__`` stays assignable (a class deriving from ``object`` directly is not)"""


class Model(_Root):
    """Abstract dynamics model (reference model/base.py:3-18)."""

    def __init__(self, x_dim: int, u_dim: int, p_dim=None, tvp_dim=None):
        self.x_dim = x_dim
        self.u_dim = u_dim
        self.p_dim = p_dim
        self.tvp_dim = tvp_dim

    def forward(self, x, u, p=None, tvp=None):
        raise NotImplementedError("")

    def jacobian(self, x, u, p=None, tvp=None):
        raise NotImplementedError("")

    def hessian(self, x, u, p=None, tvp=None):
        raise NotImplementedError("")


def adopt_reference_base(ref_model_cls=None):
    """make ``Model`` a subclass of the reference's ``Model`` (given, or found in ``sys.modules``); returns True when adopted"""
    if ref_model_cls is None:
        mod = sys.modules.get("pyNeuralEMPC.model.base")
        ref_model_cls = getattr(mod, "Model", None)
    if ref_model_cls is None or ref_model_cls is Model:
        return False
    if ref_model_cls in Model.__mro__:
        return True
    Model.__bases__ = (ref_model_cls,)
    return True


adopt_reference_base()
