"""``pyNeuralEMPC.model.jax`` under its reference name.  ``DiffDiscretJaxModel`` (``model/jax.py:10-88``) differentiates an ARBITRARY
Python function with JAX; the CUDA path evaluates feed-forward networks, so the class exists to fail with directions instead of an
``AttributeError`` deep inside a user script (SURVEY 8a lists it as the same three outputs in the same layouts as ``KerasTFModel``)."""
from __future__ import annotations

from .base import Model


class DiffDiscretJaxModel(Model):
    def __init__(self, forward_func, x_dim: int, u_dim: int, p_dim=None, tvp_dim=None, vector_mode=False, safe_mode=True):
        raise NotImplementedError(
            "pyneuralempc_b200 evaluates neural-network dynamics (Dense / tanh, sigmoid, softplus, relu) with CUDA kernels; an arbitrary JAX "
            "function cannot be compiled to them.  Wrap the network with pyneuralempc_b200.model.tensorflow.KerasTFModel (or CudaMLPModel) "
            "instead, or keep the reference's DiffDiscretJaxModel together with the reference's own integrators.")
