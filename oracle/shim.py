"""Import the UNMODIFIED reference package from ``/root/reference`` with its absent
third-party imports stubbed.  TEST / BASELINE INFRASTRUCTURE ONLY.  In the build container the
package is imported from its sources; on the GPU box (no ``/root/reference``) from the bytecode
``oracle/build_ref.py`` compiled into ``oracle/_ref/`` -- nothing under ``-m gpu`` or ``smoke()``
reads either, ``bench.py --impl reference`` and the ``cpu_baseline`` leg use the bytecode to time
the reference's own integrators and ``IpoptProblem``.  Used by ``tests/golden/make_golden.py`` to
record what the real reference computes and by the container-only tests in
``tests/test_reference_live.py``.

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
# the same package compiled to bytecode by oracle/build_ref.py (git-ignored build output that travels to the GPU box)
BYTECODE_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "pyNeuralEMPC_bytecode.zip")


def reference_available():
    """the source tree (build container) or its compiled bytecode (GPU box) is importable"""
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyNeuralEMPC")) or os.path.isfile(BYTECODE_ROOT)


def source_tree_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyNeuralEMPC"))


def import_root():
    return REFERENCE_ROOT if source_tree_available() else BYTECODE_ROOT


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
    root = import_root()
    if root not in sys.path:
        sys.path.insert(0, root)
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
