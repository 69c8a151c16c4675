"""Import the UNMODIFIED reference package from ``/root/reference`` with its absent
third-party imports stubbed.  TEST INFRASTRUCTURE ONLY; only usable in the build
container (``/root/reference`` does not exist on the GPU box), so nothing that runs under
``-m gpu``, ``smoke()`` or ``bench.py`` may call it.  Used by
``tests/golden/make_golden.py`` to record what the real reference computes and by the
container-only tests in ``tests/test_reference_live.py``.

What gets stubbed and why:
  tensorflow  ``model/tensorflow.py:1,5,77,112`` needs ``tf.function`` (identity here) and
              ``tf.compat.v1.logging``; no TensorFlow arithmetic is executed.
  jax, jax.numpy  imported by ``model/jax.py:4-5`` and ``objective/jax.py:2-3``.
  cyipopt     imported by ``optimizer/ipopt.py:4``; ``Ipopt.solve`` is never called.
Each stub carries a real ``ModuleSpec`` (torch's ``find_spec`` probes raise otherwise).
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NEMPC_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyNeuralEMPC"))


def _stub(name):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__nempc_stub__ = True
    sys.modules[name] = m
    return m


def load_reference():
    """returns the imported ``pyNeuralEMPC`` module of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if "pyNeuralEMPC" in sys.modules:
        return sys.modules["pyNeuralEMPC"]
    tf = _stub("tensorflow")
    if getattr(tf, "__nempc_stub__", False):
        tf.function = lambda f=None, **kw: f if f is not None else (lambda g: g)
        tf.compat = types.SimpleNamespace(v1=types.SimpleNamespace(
            logging=types.SimpleNamespace(set_verbosity=lambda *_: None, ERROR=40)))
    jax = _stub("jax")
    jnp = _stub("jax.numpy")
    if getattr(jax, "__nempc_stub__", False):
        jax.numpy = jnp
    _stub("cyipopt")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("pyNeuralEMPC")


def make_reference_model(mlp):
    """a subclass of the reference's own ``Model`` (``integrator/base.py:16-17`` checks
    ``isinstance``) whose forward / jacobian / hessian come from ``oracle.mlp_np.MLP`` in the
    dense ``KerasTFModel`` layouts."""
    ref = load_reference()

    class OracleBackedModel(ref.model.base.Model):
        def __init__(self):
            super().__init__(mlp.x_dim, mlp.u_dim, 0, 0)

        def forward(self, x, u, p=None, tvp=None):
            return mlp.forward(x, u)

        def jacobian(self, x, u, p=None, tvp=None):
            return mlp.dense_jacobian(x, u)

        def hessian(self, x, u, p=None, tvp=None):
            return mlp.dense_hessian(x, u)

    return OracleBackedModel()
