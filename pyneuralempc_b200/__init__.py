"""B200-native NLP-evaluation hot path of pyNeuralEMPC behind the reference's own interfaces.

Sub-modules mirror the reference package (``/root/reference/pyNeuralEMPC``): ``model``, ``integrator``,
``objective``, ``constraints``, ``optimizer``, ``controller``; ``engine`` is the batched device evaluator they
share and ``_lib`` the ctypes binding of the C ABI (``include/nempc.h``).  All arithmetic runs in hand-written
sm_100a CUDA kernels (``csrc/``); there is no CPU fallback."""
__version__ = "0.1"

from . import structure  # noqa: F401
from .engine import NlpEvaluator  # noqa: F401
from . import model, integrator, objective, constraints, optimizer, controller  # noqa: F401
