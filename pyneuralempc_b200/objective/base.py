"""``pyNeuralEMPC.objective.base`` under its reference name (``objective/base.py:4-18``)."""
from . import ObjectiveFunc  # noqa: F401
