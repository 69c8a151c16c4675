"""Batched device evaluator: one ``nempc_handle`` (C ABI, include/nempc.h) wrapped for Python.

This is the object every drop-in class shares.  One ``eval`` = residual + sparse Jacobian values + sparse
Lagrangian-Hessian values (+ objective value / gradient) of ``B`` independent transcription NLPs at one iterate,
i.e. what IPOPT asks of ``IpoptProblem.constraints / jacobian / hessian / objective / gradient``
(reference optimizer/ipopt.py:30-96) -- for a whole batch, in one or two kernel launches.

torch is used for device memory, streams and pinned host buffers only; all arithmetic is in libnempc.so.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

_NP = {"float32": np.float32, "float64": np.float64}


def _vp(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class NlpEvaluator:
    def __init__(self, weights, x_dim, u_dim, H, integrator="discrete", DT=None, activation="tanh",
                 compute_dtype="float32", io_dtype="float64", device=0, kernel="auto", tvp_dim=0, p_dim=0):
        import torch
        self._torch = torch
        self.lib = _lib.load()
        self.x_dim, self.u_dim, self.H = int(x_dim), int(u_dim), int(H)
        self.d = self.x_dim + self.u_dim
        self.tvp_dim, self.p_dim = int(tvp_dim or 0), int(p_dim or 0)
        self.integrator, self.DT, self.activation = integrator, DT, activation
        self.compute_dtype, self.io_dtype = str(compute_dtype), str(io_dtype)
        self.device = int(device)
        if integrator not in _lib.INTEGRATORS:
            raise ValueError(f"integrator must be one of {sorted(_lib.INTEGRATORS)}")
        if integrator == "rk4" and DT is None:
            raise ValueError("rk4 needs DT")
        weights = [(np.ascontiguousarray(W, np.float64), np.ascontiguousarray(b, np.float64)) for W, b in weights]
        if len(weights) > _lib.MAX_LAYERS:
            raise ValueError("too many layers")
        desc = _lib.NempcDesc()
        desc.x_dim, desc.u_dim, desc.horizon, desc.n_layers = self.x_dim, self.u_dim, self.H, len(weights)
        desc.tvp_dim, desc.p_dim = self.tvp_dim, self.p_dim
        fan_in = self.d + self.tvp_dim + self.p_dim            # the network sees [x, u, tvp, p] (model/tensorflow.py:39-47)
        for l, (W, b) in enumerate(weights):
            if W.ndim != 2 or W.shape[0] != fan_in or b.shape != (W.shape[1],):
                raise ValueError(f"layer {l}: expected kernel ({fan_in}, out) and bias (out,), got {W.shape} {b.shape}")
            desc.widths[l] = W.shape[1]
            fan_in = W.shape[1]
        desc.activation = _lib.ACTIVATIONS[activation]
        desc.integrator = _lib.INTEGRATORS[integrator]
        desc.dt = 0.0 if DT is None else float(DT)
        desc.compute_dtype = _lib.F64 if self.compute_dtype == "float64" else _lib.F32
        desc.io_dtype = _lib.F64 if self.io_dtype == "float64" else _lib.F32
        desc.device = self.device
        desc.kernel = _lib.KERNELS[kernel]
        h = ctypes.c_void_p()
        _lib.check(self.lib.nempc_create(ctypes.byref(desc), ctypes.byref(h)), None, "nempc_create")
        self._h = h
        for l, (W, b) in enumerate(weights):
            self._check(self.lib.nempc_set_weights(h, l, W.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p)),
                        "nempc_set_weights")
        self.weights = weights
        self.tdtype = torch.float64 if self.io_dtype == "float64" else torch.float32
        self.tdevice = torch.device("cuda", self.device)
        self._objective = None
        self.exo_token, self.bound_to = None, None      # see set_exogenous / CudaIpoptProblem._bind
        self._refresh_dims()
        self._pinned = {}

    def set_weights(self, layer, W, b):
        """replace the kernel / bias of one dense layer (same shapes), e.g. after the dynamics model was re-trained online;
        the reference would rebuild its KerasTFModel (model/tensorflow.py:9-29)."""
        W = np.ascontiguousarray(W, np.float64); b = np.ascontiguousarray(b, np.float64)
        if W.shape != self.weights[layer][0].shape or b.shape != self.weights[layer][1].shape:
            raise ValueError(f"layer {layer}: expected shapes {self.weights[layer][0].shape} {self.weights[layer][1].shape}")
        self._check(self.lib.nempc_set_weights(self._h, layer, W.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p)),
                    "nempc_set_weights")
        self.weights[layer] = (W, b)

    def set_exogenous(self, tvp=None, p=None):
        """time-varying / constant model inputs of the following evaluations and solves (``NMPC.next(x0, p=, tvp=)``,
        controller.py:65-113).  ``tvp``: (H, tvp_dim) shared by every problem of a batch, or (B, H, tvp_dim) [model_eval: (N, tvp_dim)];
        ``p``: (p_dim,) shared, or (B, p_dim).  numpy arrays or CUDA tensors (float64)."""
        def prep(a, dim, what):
            if dim == 0:
                if a is not None:
                    raise ValueError(f"the model has no {what} input")
                return None, 0, None
            if a is None:
                raise ValueError(f"the model has a {what} input of width {dim}")
            if self._torch.is_tensor(a):
                a = a.to(self._torch.float64).contiguous()
                if a.shape[-1] != dim:
                    raise ValueError(f"{what} must have last dim {dim}")
                return a, a.numel() // dim, ctypes.c_void_p(a.data_ptr())
            a = np.ascontiguousarray(a, np.float64)
            if a.shape[-1] != dim:
                raise ValueError(f"{what} must have last dim {dim}")
            return a, a.size // dim, a.ctypes.data_as(ctypes.c_void_p)
        ta, tr, tp = prep(tvp, self.tvp_dim, "tvp")
        pa, pr, pp = prep(p, self.p_dim, "p")
        self._check(self.lib.nempc_set_exogenous(self._h, tr, tp, pr, pp), "nempc_set_exogenous")
        # whoever memoised "my rows are the ones on the device" (Integrator._set_exogenous, CudaIpoptProblem._bind) must look again
        self.exo_token = object()
        self.bound_to = None

    # ---- lifetime / errors ------------------------------------------------------------------------------
    def _check(self, rc, what):
        _lib.check(rc, self._h, what)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.nempc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 -- interpreter shutdown
            pass

    # ---- static information -------------------------------------------------------------------------------
    def _refresh_dims(self):
        n, m, nj, nh = (ctypes.c_int64() for _ in range(4))
        self._check(self.lib.nempc_dims(self._h, ctypes.byref(n), ctypes.byref(m), ctypes.byref(nj), ctypes.byref(nh)), "nempc_dims")
        self.n, self.m, self.nnz_jac, self.nnz_hes = n.value, m.value, nj.value, nh.value
        jr, jc = np.empty(self.nnz_jac, np.int32), np.empty(self.nnz_jac, np.int32)
        hr, hc = np.empty(self.nnz_hes, np.int32), np.empty(self.nnz_hes, np.int32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self._check(self.lib.nempc_structure(self._h, p(jr), p(jc), p(hr), p(hc)), "nempc_structure")
        self.jac_rows, self.jac_cols, self.hes_rows, self.hes_cols = jr, jc, hr, hc

    def set_objective(self, lin=None, quad=None, ref=None):
        """separable cost ``sum lin*z + quad*(z-ref)^2`` (see objective.py); arrays of length n or None."""
        arrs = []
        for a in (lin, quad, ref):
            arrs.append(None if a is None else np.ascontiguousarray(np.broadcast_to(np.asarray(a, np.float64).ravel(), (self.n,))))
        p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        self._check(self.lib.nempc_set_objective(self._h, p(arrs[0]), p(arrs[1]), p(arrs[2])), "nempc_set_objective")
        self._objective = arrs
        self.bound_to = None                  # a problem that bound its cost to this evaluator re-binds at its next evaluation
        self._refresh_dims()

    @property
    def has_objective(self):
        return self._objective is not None

    @property
    def kernel_name(self):
        return self.lib.nempc_kernel_name(self._h).decode()

    @property
    def launch_count(self):
        return int(self.lib.nempc_launch_count(self._h))

    @property
    def flops_per_step(self):
        return float(self.lib.nempc_flops_per_step(self._h))

    def bytes_per_step(self):
        """algorithmic bytes per horizon step: inputs z_t, lambda_t, x_t in; resid, A|B, tril Hessian block out."""
        s = 8 if self.io_dtype == "float64" else 4
        x, d = self.x_dim, self.d
        return s * (d + x + x + x * d + d * (d + 1) // 2)

    # ---- device-resident evaluation ---------------------------------------------------------------------------
    def _as_dev(self, a, shape):
        torch = self._torch
        if a is None:
            return None
        if not isinstance(a, torch.Tensor):
            a = torch.as_tensor(np.asarray(a), dtype=self.tdtype)
        a = a.to(device=self.tdevice, dtype=self.tdtype).contiguous()
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
        return a

    def alloc_outputs(self, B, want=("resid", "jac", "hes", "obj", "grad")):
        torch = self._torch
        shapes = {"resid": (B, self.m), "jac": (B, self.nnz_jac), "hes": (B, self.nnz_hes), "obj": (B,), "grad": (B, self.n)}
        return {k: torch.empty(shapes[k], dtype=self.tdtype, device=self.tdevice) for k in want}

    def eval(self, z, x0, lam=None, obj_factor=1.0, want=("resid", "jac", "hes", "obj", "grad"), out=None, stream=None):
        """z (B,n), x0 (B,x), lam (B,m) CUDA tensors of io dtype -> dict of CUDA tensors (asynchronous)."""
        torch = self._torch
        B = int(z.shape[0])
        z = self._as_dev(z, (B, self.n))
        x0 = self._as_dev(x0, (B, self.x_dim))
        lam = self._as_dev(lam, (B, self.m))
        want = tuple(w for w in want if not (w in ("obj", "grad") and not self.has_objective))
        if "hes" in want and lam is None:
            raise ValueError("'hes' needs lam")
        sig_t, sig_s = None, 1.0
        if isinstance(obj_factor, torch.Tensor) or np.ndim(obj_factor) > 0:
            sig_t = self._as_dev(obj_factor, (B,))
        else:
            sig_s = float(obj_factor)
        if out is None:
            out = self.alloc_outputs(B, want)
        s = torch.cuda.current_stream(self.tdevice).cuda_stream if stream is None else stream
        g = lambda k: _vp(out[k]) if k in want else None
        self._check(self.lib.nempc_eval(self._h, B, _vp(z), _vp(x0), _vp(lam), _vp(sig_t), sig_s,
                                        g("resid"), g("jac"), g("hes"), g("obj"), g("grad"), ctypes.c_void_p(s)), "nempc_eval")
        return out

    def eval_blocks(self, z, x0, want_hessian=True):
        """per-step blocks before sparse assembly: Phi (B,H,x), dPhi/d(x_prev,u) (B,H,x,d), per-output Hessians (B,H,x,d,d)."""
        torch = self._torch
        B = int(z.shape[0])
        z = self._as_dev(z, (B, self.n))
        x0 = self._as_dev(x0, (B, self.x_dim))
        pred = torch.empty((B, self.H, self.x_dim), dtype=self.tdtype, device=self.tdevice)
        AB = torch.empty((B, self.H, self.x_dim, self.d), dtype=self.tdtype, device=self.tdevice)
        Hb = torch.empty((B, self.H, self.x_dim, self.d, self.d), dtype=self.tdtype, device=self.tdevice) if want_hessian else None
        s = torch.cuda.current_stream(self.tdevice).cuda_stream
        self._check(self.lib.nempc_eval_blocks(self._h, B, _vp(z), _vp(x0), _vp(pred), _vp(AB), _vp(Hb), ctypes.c_void_p(s)), "nempc_eval_blocks")
        return pred, AB, Hb

    def model_eval(self, zin, want_jac=True, want_hes=True):
        """raw network value (N,x), Jacobian (N,x,d), per-output Hessian (N,x,d,d) of stacked inputs zin (N,d)."""
        torch = self._torch
        N = int(zin.shape[0])
        zin = self._as_dev(zin, (N, self.d))
        f = torch.empty((N, self.x_dim), dtype=self.tdtype, device=self.tdevice)
        J = torch.empty((N, self.x_dim, self.d), dtype=self.tdtype, device=self.tdevice) if (want_jac or want_hes) else None
        Hs = torch.empty((N, self.x_dim, self.d, self.d), dtype=self.tdtype, device=self.tdevice) if want_hes else None
        s = torch.cuda.current_stream(self.tdevice).cuda_stream
        self._check(self.lib.nempc_model_eval(self._h, N, _vp(zin), _vp(f), _vp(J), _vp(Hs), ctypes.c_void_p(s)), "nempc_model_eval")
        return f, J, Hs

    # ---- host-buffer evaluation (the solver-callback situation) ---------------------------------------------------
    def _pin(self, key, shape):
        torch = self._torch
        t = self._pinned.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=self.tdtype).pin_memory()
            self._pinned[key] = t
        return t

    def pinned_buffers(self, B, want=("resid", "jac", "hes", "obj", "grad"), with_lambda=True, per_problem_factor=False):
        """persistent page-locked host buffers (numpy views) for a batch of B problems: write the iterate into
        ``z / x0 / lam (/ sig)``, call :meth:`eval_pinned`, read the results from the output views -- no staging copy."""
        want = tuple(w for w in want if not (w in ("obj", "grad") and not self.has_objective))
        shapes = {"z": (B, self.n), "x0": (B, self.x_dim), "resid": (B, self.m), "jac": (B, self.nnz_jac),
                  "hes": (B, self.nnz_hes), "obj": (B,), "grad": (B, self.n)}
        keys = ["z", "x0"] + list(want)
        if with_lambda:
            shapes["lam"] = (B, self.m); keys.append("lam")
        if per_problem_factor:
            shapes["sig"] = (B,); keys.append("sig")
        return {k: self._pin(k, shapes[k]).numpy() for k in keys}

    def eval_pinned(self, B, obj_factor=1.0, want=("resid", "jac", "hes", "obj", "grad"), per_problem_factor=False):
        """``nempc_eval_host`` on the buffers of :meth:`pinned_buffers`: chunk-pipelined H2D -> kernels -> D2H, then
        a synchronise.  Returns the dict of output views."""
        want = tuple(w for w in want if not (w in ("obj", "grad") and not self.has_objective))
        pz, px = self._pinned["z"], self._pinned["x0"]
        if tuple(pz.shape) != (B, self.n):
            raise ValueError("call pinned_buffers(B, ...) first")
        pl = self._pinned.get("lam") if "hes" in want else None
        if "hes" in want and (pl is None or tuple(pl.shape) != (B, self.m)):
            raise ValueError("'hes' needs the lam buffer")
        ps = self._pinned.get("sig") if per_problem_factor else None
        outs = {k: self._pinned[k] for k in want}
        g = lambda k: _vp(outs[k]) if k in want else None
        self._check(self.lib.nempc_eval_host(self._h, B, _vp(pz), _vp(px), _vp(pl), _vp(ps), float(obj_factor),
                                             g("resid"), g("jac"), g("hes"), g("obj"), g("grad")), "nempc_eval_host")
        return {k: v.numpy() for k, v in outs.items()}

    def eval_host(self, z, x0, lam=None, obj_factor=1.0, want=("resid", "jac", "hes", "obj", "grad")):
        """numpy in -> numpy out (the solver-callback situation): the arrays are copied into the persistent pinned
        buffers, then :meth:`eval_pinned`.  The returned arrays are views of pinned buffers, valid until the next call."""
        npdt = _NP[self.io_dtype]
        z = np.asarray(z, npdt)
        if z.ndim == 1:
            z = z[None]
        B = z.shape[0]
        want = tuple(w for w in want if not (w in ("obj", "grad") and not self.has_objective))
        if "hes" in want and lam is None:
            raise ValueError("'hes' needs lam")
        per_problem = np.ndim(obj_factor) > 0
        buf = self.pinned_buffers(B, want, with_lambda="hes" in want, per_problem_factor=per_problem)
        buf["z"][...] = z
        buf["x0"][...] = np.asarray(x0, npdt).reshape(B, self.x_dim)
        if "hes" in want:
            buf["lam"][...] = np.asarray(lam, npdt).reshape(B, self.m)
        if per_problem:
            buf["sig"][...] = np.asarray(obj_factor, npdt)
        return self.eval_pinned(B, 1.0 if per_problem else float(obj_factor), want, per_problem)

    # ---- batched on-device solver ------------------------------------------------------------------
    def solve(self, x0, lb, ub, z_init=None, stream=None, **options):
        """Solve ``B`` independent NMPC problems ``min f(z) s.t. c(z)=0, lb<=z<=ub`` on the device (``nempc_solve``:
        primal-dual interior point, Riccati KKT solves on the block-banded values of ``eval``).

        x0: (B, x_dim) initial states (numpy or CUDA tensor); lb / ub: length-n bound vectors in ``DomainConstraint``
        order (``get_lower_bounds(H)``), +-inf allowed; z_init: optional (B, n) warm start.  options: fields of
        ``nempc_solver_opts`` (max_iter, tol, mu_init, ...).  Returns a dict of CUDA tensors ``z`` (B,n), ``lam`` (B,m),
        ``status`` (0 converged / 1 iteration limit / 2 failed), ``iterations``, ``kkt_error``, ``outer_iterations``, and the
        ``nempc_solve_stats`` of the call: ``used_graph`` (the iteration loop ran on the device as one CUDA graph) and
        ``unaccepted_steps`` (steps taken although the line search ran out of halvings)."""
        torch = self._torch
        if self.io_dtype != "float64":
            raise ValueError("solve() needs io_dtype='float64'")
        if not self.has_objective:
            raise ValueError("solve() needs set_objective(...)")
        x0 = torch.as_tensor(np.asarray(x0, np.float64)) if not isinstance(x0, torch.Tensor) else x0
        x0 = x0.to(device=self.tdevice, dtype=torch.float64).reshape(-1, self.x_dim).contiguous()
        B = int(x0.shape[0])
        lb = np.ascontiguousarray(np.asarray(lb, np.float64).ravel())
        ub = np.ascontiguousarray(np.asarray(ub, np.float64).ravel())
        if lb.shape != (self.n,) or ub.shape != (self.n,):
            raise ValueError(f"lb / ub must have length n = {self.n}")
        if z_init is None:
            z = torch.empty((B, self.n), dtype=torch.float64, device=self.tdevice)
        else:
            z = self._as_dev(z_init, (B, self.n)).clone()
        lam = torch.empty((B, self.m), dtype=torch.float64, device=self.tdevice)
        status = torch.empty(B, dtype=torch.int32, device=self.tdevice)
        iters = torch.empty(B, dtype=torch.int32, device=self.tdevice)
        kkt = torch.empty(B, dtype=torch.float64, device=self.tdevice)
        opts = _lib.SolverOpts()
        self._check(self.lib.nempc_solver_defaults(ctypes.byref(opts)), "nempc_solver_defaults")
        for k, v in options.items():
            if not hasattr(opts, k):
                raise TypeError(f"unknown solver option {k!r}")
            setattr(opts, k, v)
        outer = ctypes.c_int32()
        s = torch.cuda.current_stream(self.tdevice).cuda_stream if stream is None else stream
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self._check(self.lib.nempc_solve(self._h, B, _vp(x0), p(lb), p(ub), _vp(z), int(z_init is not None), _vp(lam), _vp(status),
                                         _vp(iters), _vp(kkt), ctypes.byref(opts), ctypes.byref(outer), ctypes.c_void_p(s)), "nempc_solve")
        used_graph, unacc = ctypes.c_int32(), ctypes.c_int64()
        self._check(self.lib.nempc_solve_stats(self._h, ctypes.byref(used_graph), ctypes.byref(unacc)), "nempc_solve_stats")
        return dict(z=z, lam=lam, status=status, iterations=iters, kkt_error=kkt, outer_iterations=outer.value,
                    used_graph=bool(used_graph.value), unaccepted_steps=int(unacc.value))

    def host_io_bytes(self, B, want=("resid", "jac", "hes", "obj", "grad"), with_lambda=True):
        s = 8 if self.io_dtype == "float64" else 4
        h2d = B * (self.n + self.x_dim + (self.m if with_lambda else 0)) * s
        sizes = {"resid": self.m, "jac": self.nnz_jac, "hes": self.nnz_hes, "obj": 1, "grad": self.n}
        d2h = B * sum(sizes[k] for k in want) * s
        return h2d, d2h


def measure_fma_peak(device=0, dtype="float32", millis=300):
    """sustained FMA-pipe throughput (TFLOP/s) of the device: the compute-roofline denominator."""
    lib = _lib.load()
    out = ctypes.c_double()
    _lib.check(lib.nempc_measure_fma_peak(int(device), _lib.F64 if dtype == "float64" else _lib.F32, int(millis), ctypes.byref(out)),
               None, "nempc_measure_fma_peak")
    return out.value
