"""``pyNeuralEMPC.integrator.base`` under its reference name (``integrator/base.py:6-123``)."""
from . import Integrator  # noqa: F401
