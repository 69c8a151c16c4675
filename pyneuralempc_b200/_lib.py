"""ctypes binding of libnempc.so (C ABI in include/nempc.h).  There is no CPU fallback: if the library is missing
or no CUDA device is usable, every compute call raises."""
from __future__ import annotations

import ctypes
import os

from . import build
from .build import LIB_PATH

MAX_LAYERS = 8
OK = 0
F32, F64 = 0, 1
INTEGRATORS = {"discrete": 0, "unity": 1, "rk4": 2}
ACTIVATIONS = {"tanh": 0, "sigmoid": 1, "softplus": 2, "relu": 3}
KERNELS = {"auto": 0, "generic": 1, "fast": 2, "tc": 3}
ABI_VERSION = 4                             # NEMPC_ABI_VERSION of include/nempc.h this binding was written against

EXPORTS = ("nempc_version", "nempc_last_error", "nempc_create", "nempc_destroy", "nempc_set_weights",
           "nempc_set_objective", "nempc_set_exogenous", "nempc_structure_counts", "nempc_structure_fill", "nempc_dims", "nempc_structure",
           "nempc_eval", "nempc_eval_host", "nempc_eval_blocks", "nempc_model_eval", "nempc_launch_count",
           "nempc_kernel_name", "nempc_flops_per_step", "nempc_measure_fma_peak", "nempc_objective_eval", "nempc_solver_defaults", "nempc_solve", "nempc_solve_stats",
           "nempc_abi_info", "nempc_source_hash", "nempc_rolling_gather", "nempc_rolling_assemble",
           "nempc_quadform_eval", "nempc_hessian_merge")


class NempcDesc(ctypes.Structure):
    _fields_ = [("x_dim", ctypes.c_int32), ("u_dim", ctypes.c_int32), ("horizon", ctypes.c_int32),
                ("n_layers", ctypes.c_int32), ("widths", ctypes.c_int32 * MAX_LAYERS),
                ("activation", ctypes.c_int32), ("integrator", ctypes.c_int32), ("dt", ctypes.c_double),
                ("compute_dtype", ctypes.c_int32), ("io_dtype", ctypes.c_int32), ("device", ctypes.c_int32),
                ("kernel", ctypes.c_int32), ("tvp_dim", ctypes.c_int32), ("p_dim", ctypes.c_int32)]


class SolverOpts(ctypes.Structure):
    _fields_ = [("max_iter", ctypes.c_int32), ("max_backtrack", ctypes.c_int32)] + \
               [(k, ctypes.c_double) for k in ("tol", "mu_init", "mu_min", "kappa_eps", "kappa_mu", "theta_mu", "tau_min",
                                               "bound_push", "eta", "reg_init", "reg_max")]


class NempcError(RuntimeError):
    pass


_lib = None


def load():
    """dlopen libnempc.so (built by ``__graft_entry__.build()`` / ``pyneuralempc_b200.build.build_library()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NempcError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(pyneuralempc_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    if not hasattr(lib, "nempc_abi_info") or not hasattr(lib, "nempc_source_hash"):
        raise NempcError(f"{LIB_PATH} predates nempc_abi_info: rebuild it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib.nempc_source_hash.restype = ctypes.c_char_p
    built_from = lib.nempc_source_hash().decode()
    if not os.environ.get("NEMPC_LIB_PATH") and built_from != "unknown" and built_from != build.source_hash():
        raise NempcError(f"{LIB_PATH} was built from other sources (hash {built_from}) than the ones in pyneuralempc_b200/csrc "
                         f"({build.source_hash()}): rebuild it with `python -c 'import __graft_entry__ as g; g.build()'`")
    ds, os_ = ctypes.c_int32(), ctypes.c_int32()
    ver = lib.nempc_abi_info(ctypes.byref(ds), ctypes.byref(os_))
    if ver != ABI_VERSION or ds.value != ctypes.sizeof(NempcDesc) or os_.value != ctypes.sizeof(SolverOpts):
        raise NempcError(f"{LIB_PATH}: ABI {ver} with nempc_desc {ds.value} B / nempc_solver_opts {os_.value} B, this binding expects ABI "
                         f"{ABI_VERSION} with {ctypes.sizeof(NempcDesc)} B / {ctypes.sizeof(SolverOpts)} B: rebuild the library")
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    lib.nempc_version.restype = ctypes.c_char_p
    lib.nempc_last_error.restype = ctypes.c_char_p
    lib.nempc_last_error.argtypes = [vp]
    lib.nempc_create.argtypes = [ctypes.POINTER(NempcDesc), ctypes.POINTER(vp)]
    lib.nempc_destroy.argtypes = [vp]
    lib.nempc_set_weights.argtypes = [vp, i32, vp, vp]
    lib.nempc_set_objective.argtypes = [vp, vp, vp, vp]
    lib.nempc_set_exogenous.argtypes = [vp, i64, vp, i64, vp]
    lib.nempc_structure_counts.argtypes = [i32, i32, i32, vp, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.nempc_structure_fill.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp]
    lib.nempc_dims.argtypes = [vp] + [ctypes.POINTER(i64)] * 4
    lib.nempc_structure.argtypes = [vp, vp, vp, vp, vp]
    lib.nempc_eval.argtypes = [vp, i64, vp, vp, vp, vp, dbl, vp, vp, vp, vp, vp, vp]
    lib.nempc_eval_host.argtypes = [vp, i64, vp, vp, vp, vp, dbl, vp, vp, vp, vp, vp]
    lib.nempc_eval_blocks.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp]
    lib.nempc_model_eval.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    lib.nempc_launch_count.argtypes = [vp]
    lib.nempc_launch_count.restype = i64
    lib.nempc_kernel_name.argtypes = [vp]
    lib.nempc_kernel_name.restype = ctypes.c_char_p
    lib.nempc_flops_per_step.argtypes = [vp]
    lib.nempc_flops_per_step.restype = dbl
    lib.nempc_measure_fma_peak.argtypes = [i32, i32, i32, ctypes.POINTER(dbl)]
    lib.nempc_objective_eval.argtypes = [i32, i64, i64, vp, vp, vp, vp, vp, vp, vp]
    lib.nempc_solver_defaults.argtypes = [ctypes.POINTER(SolverOpts)]
    lib.nempc_solve.argtypes = [vp, i64, vp, vp, vp, vp, i32, vp, vp, vp, vp, ctypes.POINTER(SolverOpts), ctypes.POINTER(i32), vp]
    lib.nempc_solve_stats.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i64)]
    lib.nempc_rolling_gather.argtypes = [i32, i64, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.nempc_rolling_assemble.argtypes = [i32, i64, i32, i32, i32, i32, i32, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl,
                                           vp, vp, vp, vp]
    lib.nempc_quadform_eval.argtypes = [i32, i64, i64, vp, vp, vp, vp, vp, dbl, vp, vp, vp]
    lib.nempc_hessian_merge.argtypes = [i32, i64, i64, i64, vp, vp, vp, vp, dbl, vp, vp]
    for name in EXPORTS:
        getattr(lib, name)                  # AttributeError here = the .so does not match include/nempc.h
    _lib = lib
    return lib


def check(rc, handle=None, what=""):
    if rc != OK:
        msg = load().nempc_last_error(handle)
        raise NempcError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
