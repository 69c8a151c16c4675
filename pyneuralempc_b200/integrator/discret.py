"""``pyNeuralEMPC.integrator.discret`` under its reference name (``integrator/discret.py:8-81``): x_{t+1} = x_t + f(x_t, u_t)."""
from . import CudaDiscretIntegrator, DiscretIntegrator  # noqa: F401
