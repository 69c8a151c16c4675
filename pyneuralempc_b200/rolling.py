"""Rolling-window (NARX) models on the device.

Reference: ``KerasTFModelRollingInput`` (``model/tensorflow.py:131-340``) and ``DiffDiscretJaxModelRollingWindow``
(``model/jax.py:93-259``).  The network of model row ``i`` reads ``w = rolling_window`` consecutive rows of the history-extended
inputs ``[prev_x ; x]``, ``[prev_u ; u]`` (``set_prev_data`` supplies the ``w - 1`` rows of history): its input vector is

    [x_ext[i], ..., x_ext[i + w - 1] | u_ext[i], ..., u_ext[i + w - 1]]        (oldest first; reversed when forward_rolling=False,
                                                                                 tensorflow.py:119-127)

Inside the transcription (``integrator/discret.py:13-81``, ``unity.py:15-81``: ``model(x_{t-1} rows, u rows)``) constraint row-block
``t`` therefore depends on the states ``x_{t-w+1} .. x_t`` (``x_0`` and older ones are data) and controls ``u_{t-w+1} .. u_t``: the
Jacobian band and the Hessian blocks are ``w`` steps wide instead of one.

``RollingNlpEvaluator`` is the batched device evaluator for that structure, with the attribute / method surface the problem classes use
of ``engine.NlpEvaluator``.  Per evaluation: window gather kernel -> ``nempc_model_eval`` on the ``dw = w (x + u)``-input network
(per-row value, Jacobian, per-output Hessian) -> banded sparse assembly kernel (``csrc/nempc_rolling.cuh``).  torch only carries device
buffers and streams."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .engine import NlpEvaluator

INT32_MIN = -2 ** 31


def window_columns(H, x_dim, u_dim, w, forward_rolling=True, model_level=False):
    """For every model row t and network input j < dw: the variable it reads.  Returns an (H, dw) int array of gather codes:
    ``>= 0``: index into z = [x_1..x_H | u_0..u_{H-1}];  ``-1 - q``: index q into aux = [x0 | prev_x (w-1, x) | prev_u (w-1, u)].
    ``model_level``: the Model-interface view instead (tensorflow.py:49-109 style call ``model.jacobian(x, u)``): every row of ``x`` is a
    variable, codes ``>= 0`` index ``[x.ravel() | u.ravel()]`` and the x0 slot of aux is unused."""
    dw = w * (x_dim + u_dim)
    code = np.zeros((H, dw), np.int64)
    if model_level:
        for t in range(H):
            for k in range(w):
                pos = k if forward_rolling else w - 1 - k
                src = t + k - (w - 1)
                for c in range(x_dim):
                    code[t, pos * x_dim + c] = src * x_dim + c if src >= 0 else -1 - (x_dim + (src + (w - 1)) * x_dim + c)
                for c in range(u_dim):
                    code[t, w * x_dim + pos * u_dim + c] = (H * x_dim + src * u_dim + c if src >= 0
                                                            else -1 - (x_dim + (w - 1) * x_dim + (src + (w - 1)) * u_dim + c))
        return code
    for t in range(H):
        for k in range(w):                                   # k-th row of the window, oldest first: extended row t + k
            pos = k if forward_rolling else w - 1 - k        # its place inside the network input (tensorflow.py:119-127)
            src = t + k - (w - 1)                            # row of x_{t-1}-array / u-array; negative = history
            for c in range(x_dim):
                j = pos * x_dim + c
                if src >= 1:
                    code[t, j] = (src - 1) * x_dim + c                        # x_src is z's state block src - 1
                elif src == 0:
                    code[t, j] = -1 - c                                        # x0
                else:
                    code[t, j] = -1 - (x_dim + (src + (w - 1)) * x_dim + c)    # prev_x row src + w - 1
            for c in range(u_dim):
                j = w * x_dim + pos * u_dim + c
                if src >= 0:
                    code[t, j] = H * x_dim + src * u_dim + c
                else:
                    code[t, j] = -1 - (x_dim + (w - 1) * x_dim + (src + (w - 1)) * u_dim + c)
    return code


def rolling_structure(H, x_dim, u_dim, w, integrator, quad_mask=None, forward_rolling=True):
    """closed-form sparsity + assembly tables of the rolling-window NLP.  Value order = the reference's: row-major non-zeros of the dense
    Jacobian; ``np.nonzero(np.tril(objective_map + integrator_map))`` for the Hessian (``optimizer/ipopt.py:55-62``)."""
    d = x_dim + u_dim
    dw, n, m = w * d, H * d, H * x_dim
    code = window_columns(H, x_dim, u_dim, w, forward_rolling)
    unity = integrator == "unity"
    # ---- Jacobian: per row, its columns in ascending order
    jr, jc, jsrc, jadd = [], [], [], []
    for t in range(H):
        for p in range(x_dim):
            cols = {}
            for j in range(dw):
                if code[t, j] >= 0:
                    cols[int(code[t, j])] = [(t * x_dim + p) * dw + j, 0.0]
            cols[t * x_dim + p] = [-1, -1.0]                                   # - x_t
            if not unity and t >= 1:                                           # x_{t-1} + f: identity on the newest window state
                cols[(t - 1) * x_dim + p][1] += 1.0
            for c in sorted(cols):
                jr.append(t * x_dim + p); jc.append(c); jsrc.append(cols[c][0]); jadd.append(cols[c][1])
    # ---- Hessian: union of the window blocks (variables only) + objective diagonal, lower triangle, row-major
    contrib = {}
    for t in range(H):
        var = [(int(code[t, j]), j) for j in range(dw) if code[t, j] >= 0]
        for ca, ja in var:
            for cb, jb in var:
                if cb <= ca:
                    contrib.setdefault((ca, cb), []).append(t * dw * dw + ja * dw + jb)
    if quad_mask is not None:
        for i in np.nonzero(np.asarray(quad_mask))[0]:
            contrib.setdefault((int(i), int(i)), [])
    keys = sorted(contrib)
    hr = np.array([k[0] for k in keys], np.int32)
    hc = np.array([k[1] for k in keys], np.int32)
    ptr = np.zeros(len(keys) + 1, np.int32)
    src = []
    for i, k in enumerate(keys):
        src.extend(contrib[k])
        ptr[i + 1] = len(src)
    base = np.full(m, INT32_MIN, np.int64)
    if not unity:
        for t in range(H):
            for p in range(x_dim):
                base[t * x_dim + p] = (t - 1) * x_dim + p if t >= 1 else -1 - p
    return dict(n=n, m=m, dw=dw, gather=code.reshape(-1).astype(np.int32), resid_base=base.astype(np.int32),
                jac_rows=np.array(jr, np.int32), jac_cols=np.array(jc, np.int32), jac_src=np.array(jsrc, np.int32),
                jac_add=np.array(jadd, np.float64), hes_rows=hr, hes_cols=hc, hes_ptr=ptr, hes_src=np.array(src, np.int32))


class RollingNlpEvaluator:
    """batched evaluator of the rolling-window transcription; same surface as ``NlpEvaluator`` where the problem classes touch it"""

    def __init__(self, weights, x_dim, u_dim, H, integrator, rolling_window, forward_rolling=True, activation="tanh",
                 compute_dtype="float32", io_dtype="float64", device=0):
        import torch
        if integrator not in ("discrete", "unity"):
            raise NotImplementedError("rolling-window models run under the discrete / unity integrators (the reference's RK4 keeps only the "
                                      "diagonal blocks of the model Jacobian, integrator/rk4.py:85-110, which drops the window coupling)")
        self._torch = torch
        self.x_dim, self.u_dim, self.H, self.w = int(x_dim), int(u_dim), int(H), int(rolling_window)
        self.tvp_dim = self.p_dim = 0
        self.integrator, self.forward_rolling = integrator, bool(forward_rolling)
        self.dw = self.w * (self.x_dim + self.u_dim)
        if self.dw > 16:
            raise NotImplementedError(f"rolling_window * (x_dim + u_dim) = {self.dw} exceeds the 16 differentiated inputs the kernels take")
        if weights[0][0].shape[0] != self.dw:
            raise ValueError(f"the network needs rolling_window * (x_dim + u_dim) = {self.dw} inputs, it has {weights[0][0].shape[0]}")
        # the window network as a plain model: x outputs, dw differentiated inputs
        self.net = NlpEvaluator(weights, self.x_dim, self.dw - self.x_dim, 1, "unity", activation=activation,
                                compute_dtype=compute_dtype, io_dtype=io_dtype, device=device, kernel="generic")
        self.lib = self.net.lib
        self.tdevice, self.tdtype = self.net.tdevice, self.net.tdtype
        self.io = _lib.F64 if io_dtype == "float64" else _lib.F32
        self.naux = self.x_dim + (self.w - 1) * (self.x_dim + self.u_dim)
        self._objective = None
        self.exo_token, self.bound_to = None, None
        self._launches = 0
        self.prev_source = None                  # a model object whose prev_x / prev_u are read at every evaluation (integrator hook-up)
        self.set_prev_data(np.zeros((self.w - 1, self.x_dim)), np.zeros((self.w - 1, self.u_dim)))
        self._build(None)

    # ---- structure ---------------------------------------------------------------------------------------------------------
    def _build(self, quad):
        st = rolling_structure(self.H, self.x_dim, self.u_dim, self.w, self.integrator, None if quad is None else quad != 0,
                               self.forward_rolling)
        self.n, self.m = st["n"], st["m"]
        self.jac_rows, self.jac_cols, self.hes_rows, self.hes_cols = st["jac_rows"], st["jac_cols"], st["hes_rows"], st["hes_cols"]
        self.nnz_jac, self.nnz_hes = len(self.jac_rows), len(self.hes_rows)
        hes_obj = np.zeros(self.nnz_hes)
        if quad is not None:
            diag = self.hes_rows == self.hes_cols
            hes_obj[diag] = 2.0 * quad[self.hes_rows[diag]]
        t = self._torch
        dev = self.tdevice
        self._tab = {k: t.as_tensor(st[k], device=dev) for k in ("gather", "resid_base", "jac_src", "jac_add", "hes_ptr", "hes_src")}
        self._tab["hes_obj"] = t.as_tensor(hes_obj, device=dev)

    def set_objective(self, lin=None, quad=None, ref=None):
        arrs = [None if a is None else np.ascontiguousarray(np.broadcast_to(np.asarray(a, np.float64).ravel(), (self.n,))) for a in (lin, quad, ref)]
        self._objective = [np.zeros(self.n) if a is None else a for a in arrs]
        self._obj_dev = [self._torch.as_tensor(np.array(a), dtype=self._torch.float64, device=self.tdevice) for a in self._objective]
        self.bound_to = None
        self._build(self._objective[1])

    @property
    def has_objective(self):
        return self._objective is not None

    def set_prev_data(self, x_prev, u_prev):
        """history rows (w-1, x_dim) / (w-1, u_dim) shared by the batch, or (B, w-1, .) per problem (tensorflow.py:174-185)"""
        xp, up = np.asarray(x_prev, np.float64), np.asarray(u_prev, np.float64)
        assert xp.shape[-2:] == (self.w - 1, self.x_dim), f"Your x prev tensor must have the following shape {(self.w - 1, self.x_dim)} (received : {xp.shape})"
        assert up.shape[-2:] == (self.w - 1, self.u_dim), f"Your u prev tensor must have the following shape {(self.w - 1, self.u_dim)} (received : {up.shape})"
        self._prev = (xp, up)
        self.exo_token = object()

    def set_exogenous(self, tvp=None, p=None):
        if tvp is not None or p is not None:
            raise NotImplementedError("rolling-window models take no tvp / p inputs here")

    @property
    def launch_count(self):
        return self._launches + self.net.launch_count

    @property
    def kernel_name(self):
        return f"nempc_rolling_gather/assemble_kernel (window {self.w}) + " + self.net.kernel_name

    def close(self):
        self.net.close()

    # ---- evaluation ------------------------------------------------------------------------------------------------------------
    def _aux(self, x0):
        t = self._torch
        B = x0.shape[0]
        xp, up = self._prev
        if self.prev_source is not None and self.w > 1:
            assert self.prev_source.prev_x is not None and self.prev_source.prev_u is not None, \
                "You must give history window with set_prev_data before calling any inferance function."     # tensorflow.py:189
            xp, up = self.prev_source.prev_x, self.prev_source.prev_u
        parts = [x0]
        if self.w == 1:
            return x0
        for a, dim in ((xp, self.x_dim), (up, self.u_dim)):
            a = t.as_tensor(a, dtype=self.tdtype, device=self.tdevice).reshape(-1, (self.w - 1) * dim)
            parts.append(a.expand(B, -1) if a.shape[0] == 1 else a)
        return t.cat(parts, dim=1).contiguous()

    def alloc_outputs(self, B, want=("resid", "jac", "hes", "obj", "grad")):
        t = self._torch
        shp = {"resid": (B, self.m), "jac": (B, self.nnz_jac), "hes": (B, self.nnz_hes), "obj": (B,), "grad": (B, self.n)}
        return {k: t.empty(shp[k], dtype=self.tdtype, device=self.tdevice) for k in want}

    def eval(self, z, x0, lam=None, sigma=1.0, want=("resid", "jac", "hes", "obj", "grad"), out=None):
        t = self._torch
        z = t.as_tensor(z, dtype=self.tdtype, device=self.tdevice).reshape(-1, self.n).contiguous()
        B = z.shape[0]
        x0 = t.as_tensor(x0, dtype=self.tdtype, device=self.tdevice).reshape(B, self.x_dim).contiguous()
        want = tuple(k for k in want if not (k == "hes" and lam is None) and not (k in ("obj", "grad") and not self.has_objective))
        out = out if out is not None else self.alloc_outputs(B, want)
        p = lambda a: None if a is None else ctypes.c_void_p(a.data_ptr())
        s = ctypes.c_void_p(t.cuda.current_stream(self.tdevice).cuda_stream)
        need_jac, need_hes = "jac" in want, "hes" in want
        if "resid" in want or need_jac or need_hes:
            aux = self._aux(x0)
            zin = t.empty((B * self.H, self.dw), dtype=self.tdtype, device=self.tdevice)
            _lib.check(self.lib.nempc_rolling_gather(self.io, B, self.n, self.naux, self.H * self.dw, p(self._tab["gather"]), p(z), p(aux), p(zin), s),
                       None, "nempc_rolling_gather")
            f, J, Hs = self.net.model_eval(zin, need_jac or need_hes, need_hes)
            lam_t = None if not need_hes else t.as_tensor(lam, dtype=self.tdtype, device=self.tdevice).reshape(B, self.m).contiguous()
            sig_t = sigma if t.is_tensor(sigma) else None
            if sig_t is not None:
                sig_t = sig_t.to(dtype=self.tdtype, device=self.tdevice).reshape(B).contiguous()
            elif not np.isscalar(sigma):
                sig_t = t.as_tensor(np.asarray(sigma), dtype=self.tdtype, device=self.tdevice).reshape(B).contiguous()
            tb = self._tab
            _lib.check(self.lib.nempc_rolling_assemble(
                self.io, B, self.H, self.x_dim, self.dw, self.n, self.naux, self.nnz_jac, self.nnz_hes, p(tb["resid_base"]), p(tb["jac_src"]),
                p(tb["jac_add"]), p(tb["hes_ptr"]), p(tb["hes_src"]), p(tb["hes_obj"]), p(z), p(aux), p(f), p(J if need_jac else None),
                p(Hs if need_hes else None), p(lam_t), p(sig_t), float(sigma) if sig_t is None else 1.0,
                p(out.get("resid")), p(out.get("jac")), p(out.get("hes")), s), None, "nempc_rolling_assemble")
            self._launches += 2
        if "obj" in want or "grad" in want:
            lin, quad, ref = self._obj_dev
            _lib.check(self.lib.nempc_objective_eval(self.io, B, self.n, p(z), p(lin), p(quad), p(ref), p(out.get("obj")), p(out.get("grad")), s),
                       None, "nempc_objective_eval")
            self._launches += 1
        return {k: out[k] for k in want}

    def eval_host(self, z, x0, lam=None, obj_factor=1.0, want=("resid", "jac", "hes", "obj", "grad")):
        """numpy in, numpy out (the solver-callback form): (B, .) arrays"""
        z = np.atleast_2d(np.asarray(z, np.float64))
        x0 = np.atleast_2d(np.asarray(x0, np.float64))
        lam = None if lam is None else np.atleast_2d(np.asarray(lam, np.float64))
        out = self.eval(z, x0, lam, obj_factor, want)
        self._torch.cuda.synchronize(self.tdevice)
        return {k: v.cpu().numpy().astype(np.float64) for k, v in out.items()}
