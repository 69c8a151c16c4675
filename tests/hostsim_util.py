"""ctypes wrapper around tests/_hostsim/libnempc_hostsim.so -- the TEST-ONLY host emulation of the device kernel
bodies (tests/hostsim/hostsim.cpp).  Never imported by the product package."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SRC = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "_hostsim")
LIB = os.path.join(OUT_DIR, "libnempc_hostsim.so")
_DEPS = [SRC] + [os.path.join(ROOT, "pyneuralempc_b200", "csrc", f)
                 for f in ("nempc_generic.cuh", "nempc_fast.cuh", "nempc_layout.h", "nempc_solver.cuh")]
INTEG = {"discrete": 0, "unity": 1, "rk4": 2}
ACT = {"tanh": 0, "sigmoid": 1, "softplus": 2, "relu": 3}


def build():
    os.makedirs(OUT_DIR, exist_ok=True)
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in _DEPS):
        return LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", LIB, SRC])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.hostsim_run.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def run(mlp, kind, H, DT, Z, X0, lam=None, sigma=1.0, quad=None, compute_f64=True, kernel="generic",
        what="eval", want_jac=True, want_hes=True, tvp=None, p=None):
    """what='eval' -> dict(resid, jac_vals, hes_vals); 'blocks' -> (pred, AB, Hblk); 'model' -> (f, J, Hs) with
    Z holding the stacked network inputs (N, d)."""
    from oracle import structure as S
    xd, ud = mlp.x_dim, mlp.u_dim
    d = xd + ud
    widths = np.asarray([W.shape[1] for W, _ in mlp.weights], np.int32)
    wflat = np.concatenate([np.concatenate([np.asarray(W, np.float64).ravel(), np.asarray(b, np.float64).ravel()])
                            for W, b in mlp.weights])
    Z = np.ascontiguousarray(Z, np.float64)
    X0 = None if X0 is None else np.ascontiguousarray(X0, np.float64)
    lam = None if lam is None else np.ascontiguousarray(lam, np.float64)
    quad = None if quad is None else np.ascontiguousarray(quad, np.float64)
    sig_arr, sig_s = None, 1.0
    if np.ndim(sigma) == 0:
        sig_s = float(sigma)
    else:
        sig_arr = np.ascontiguousarray(sigma, np.float64)
    B = Z.shape[0]
    w = {"eval": 0, "blocks": 1, "model": 2}[what]
    if what == "eval":
        nj = S.nnz_jacobian(H, xd, ud)
        r, _ = S.hessian_structure(H, xd, ud, None if quad is None else quad != 0)
        o0 = np.full((B, H * xd), np.nan)
        o1 = np.full((B, nj), np.nan) if (want_jac or want_hes) and True else None
        if not want_jac and not want_hes:
            o1 = None
        o2 = np.full((B, len(r)), np.nan) if (want_hes and lam is not None) else None
    else:
        N = B if what == "model" else B * H
        o0 = np.full((N, xd), np.nan)
        o1 = np.full((N, xd, d), np.nan) if (want_jac or want_hes) else None
        o2 = np.full((N, xd, d, d), np.nan) if want_hes else None
    if tvp is not None or p is not None:                      # ExoMLP: rows in C-ABI layout (include/nempc.h nempc_set_exogenous)
        tvp = None if tvp is None else np.ascontiguousarray(tvp, np.float64)
        p = None if p is None else np.ascontiguousarray(p, np.float64)
        td, pd = (0 if tvp is None else tvp.shape[-1]), (0 if p is None else p.shape[-1])
        tvp_b = 0 if (tvp is None or tvp.ndim == 2 and what != "model") else H * td
        p_b = 0 if (p is None or p.ndim == 1) else pd
        lib().hostsim_set_exo(td, pd, _p(tvp), ctypes.c_longlong(tvp_b), _p(p), ctypes.c_longlong(p_b))
    rc = lib().hostsim_run(xd, ud, int(H), len(widths), _p(widths), ACT[mlp.activation], INTEG.get(kind, 0),
                           ctypes.c_double(0.0 if DT is None else DT), int(compute_f64), 1 if kernel == "fast" else 0, w,
                           _p(wflat), _p(quad), ctypes.c_longlong(B), _p(Z), _p(X0), _p(lam), _p(sig_arr),
                           ctypes.c_double(sig_s), _p(o0), _p(o1), _p(o2))
    if rc != 0:
        raise RuntimeError(f"hostsim_run: unsupported combination (rc={rc})")
    if what == "eval":
        return {"resid": o0, "jac_vals": o1, "hes_vals": o2}
    return o0, o1, o2


def solve(mlp, kind, H, DT, obj, X0, lb, ub, Z_init=None, max_iter=60, tol=1e-6):
    """host emulation of nempc_solve (float64 arithmetic).  Returns Z, lam, dict(status, iterations, kkt_error, outer)."""
    xd, ud = mlp.x_dim, mlp.u_dim
    n, m = H * (xd + ud), H * xd
    widths = np.asarray([W.shape[1] for W, _ in mlp.weights], np.int32)
    wflat = np.concatenate([np.concatenate([np.asarray(W, np.float64).ravel(), np.asarray(b, np.float64).ravel()]) for W, b in mlp.weights])
    X0 = np.ascontiguousarray(np.atleast_2d(X0), np.float64)
    B = X0.shape[0]
    Z = np.zeros((B, n)) if Z_init is None else np.ascontiguousarray(Z_init, np.float64).copy()
    lam = np.zeros((B, m)); status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32); kkt = np.zeros(B)
    lin, quad, ref = (np.ascontiguousarray(a, np.float64) for a in (obj.lin, obj.quad, obj.ref))
    lb, ub = np.ascontiguousarray(lb, np.float64), np.ascontiguousarray(ub, np.float64)
    fn = lib().hostsim_solve
    fn.restype = ctypes.c_int
    outer = fn(xd, ud, int(H), len(widths), _p(widths), ACT[mlp.activation], INTEG[kind], ctypes.c_double(0.0 if DT is None else DT),
               _p(wflat), _p(lin), _p(quad), _p(ref), ctypes.c_longlong(B), _p(X0), _p(lb), _p(ub), _p(Z), int(Z_init is not None),
               _p(lam), _p(status), _p(iters), _p(kkt), int(max_iter), ctypes.c_double(tol))
    return Z, lam, dict(status=status, iterations=iters, kkt_error=kkt, outer=outer)
