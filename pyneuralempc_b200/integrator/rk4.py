"""``pyNeuralEMPC.integrator.rk4`` under its reference name (``integrator/rk4.py:46-285``): classic Runge-Kutta 4 over f."""
from . import CudaRK4Integrator, RK4Integrator  # noqa: F401
