"""``pyNeuralEMPC.integrator.unity`` under its reference name (``integrator/unity.py:9-81``): x_{t+1} = f(x_t, u_t)."""
from . import CudaUnityIntegrator, UnityIntegrator  # noqa: F401
