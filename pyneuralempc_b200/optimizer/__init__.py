"""Solver interfaces (mirrors /root/reference/pyNeuralEMPC/optimizer/__init__.py)."""
from .base import Optimizer, ProblemFactory, ProblemInterface, ProblemInterfaceHessianFree  # noqa: F401
from .ipopt import CudaIpoptProblem, Ipopt, IpoptProblem, IpoptProblemFactory  # noqa: F401
from .slsqp import Slsqp, SlsqpProblem, TrustConstr  # noqa: F401
from .ipm import CudaIpm  # noqa: F401
