"""Integrator layer: equality constraints ``c_t = Phi(x_{t-1}, u_t) - x_t`` evaluated by CUDA kernels.

Mirrors ``/root/reference/pyNeuralEMPC/integrator/``: ``Integrator`` (base.py:5-123), ``DiscretIntegrator``
(discret.py), ``UnityIntegrator`` (unity.py) and ``RK4Integrator`` (rk4.py) -- same constructors, attributes
(``H``, ``model``, ``nb_contraints`` [sic]), call signatures, assertion behaviour and dense return layouts, so a
reference ``IpoptProblem`` / ``SlsqpProblem`` can drive them unchanged:

* ``forward(x, u, x0)``  -> (m,)            discret.py:13-30, unity.py:15-32, rk4.py:57-83
* ``jacobian(x, u, x0)`` -> dense (m, n)     discret.py:32-58, rk4.py:113-178
* ``hessian(x, u, x0)``  -> dense (m, n, n)  discret.py:61-81, rk4.py:181-285 (any x_dim+u_dim, not only 3)
* ``hessianstructure()`` -> (n, n) 0/1 map   base.py:83-115 (analytic instead of numerically probed)

The dense arrays exist for drop-in compatibility only (O(H^3) like the reference).  The scalable interface is
``.evaluator`` (sparse values in the structure order; see ``optimizer.ipopt.CudaIpoptProblem``).
"""
from __future__ import annotations

import numpy as np

from ..engine import NlpEvaluator
from ..model import CudaMLPModel, CudaMLPModelRollingInput, Model


class Integrator:
    """Abstract integrator (reference integrator/base.py:5-123)."""

    def __init__(self, model, H: int, nb_contraints: int):
        if not isinstance(model, (Model,)):
            raise ValueError("The model provided isn't a Model object !")
        self.H = H
        self.model = model
        self.nb_contraints = nb_contraints
        self.hessian_structure_cache = None

    def forward(self, x, u, x0, p=None, tvp=None):
        raise NotImplementedError("")

    def jacobian(self, x, u, x0, p=None, tvp=None):
        raise NotImplementedError("")

    def hessian(self, x, u, x0, p=None, tvp=None):
        raise NotImplementedError("")

    def hessianstructure(self):
        if self.hessian_structure_cache is None:
            self.hessian_structure_cache = self._compute_hessianstructure()
        return self.hessian_structure_cache

    def get_lower_bounds(self, _):
        return [0.0, ] * self.nb_contraints

    def get_upper_bounds(self, _):
        return [0.0, ] * self.nb_contraints


class _CudaIntegrator(Integrator):
    KIND = None

    def __init__(self, model, H, DT=None, cache_mode=False, cache_size=2):
        if not isinstance(model, (Model,)):
            raise ValueError("The model provided isn't a Model object !")
        super().__init__(model, H, model.x_dim * H)
        if not isinstance(model, (CudaMLPModel, CudaMLPModelRollingInput)):
            raise ValueError("CUDA integrators need a CudaMLPModel (the network is fused into the integrator kernel)")
        self.DT = DT
        self.cache_mode = cache_mode
        self._cache, self._cache_size = [], max(1, int(cache_size))
        self._exo_key, self._exo_token = None, None
        self.rolling = isinstance(model, CudaMLPModelRollingInput)
        if self.rolling:
            # rolling-window (NARX) model: banded structure, its own evaluator (rolling.py); the history rows are read from the model at
            # every evaluation (KerasTFModelRollingInput.set_prev_data is called between NMPC.next calls)
            from ..rolling import RollingNlpEvaluator
            self.evaluator = RollingNlpEvaluator(model.weights, model.x_dim, model.u_dim, H, self.KIND, model.rolling_window,
                                                 forward_rolling=model.forward_rolling, activation=model.activation,
                                                 compute_dtype=model.dtype, io_dtype="float64", device=model.device)
            self.evaluator.prev_source = model
            return
        self.evaluator = NlpEvaluator(model.weights, model.x_dim, model.u_dim, H, self.KIND, DT=DT,
                                      activation=model.activation, compute_dtype=model.dtype, io_dtype="float64",
                                      device=model.device, kernel=model.kernel,
                                      tvp_dim=model.tvp_dim or 0, p_dim=model.p_dim or 0)

    def _set_exogenous(self, p, tvp):
        """hand the model's tvp (H, tvp_dim) / p (p_dim,) rows to the evaluator when they changed (they are fixed during one solve:
        controller.py:65-113 sets them once per NMPC.next)."""
        m = self.model
        if not (m.tvp_dim or m.p_dim):
            assert p is None and tvp is None, "the model declares no p / tvp input"
            return
        assert (tvp is None) == (not m.tvp_dim) and (p is None) == (not m.p_dim), "p / tvp must match the model's p_dim / tvp_dim"
        tvp = None if tvp is None else np.asarray(tvp, np.float64).reshape(self.H, m.tvp_dim)
        p = None if p is None else np.asarray(p, np.float64).reshape(m.p_dim)
        key = (b"" if tvp is None else tvp.tobytes()) + b"|" + (b"" if p is None else p.tobytes())
        if key != self._exo_key or getattr(self.evaluator, "exo_token", None) is not self._exo_token:
            self.evaluator.set_exogenous(tvp, p)          # also when somebody else (another problem, BatchedNMPC) replaced the rows
            self._exo_key, self._exo_token = key, self.evaluator.exo_token
            self._cache = []

    # ---- helpers --------------------------------------------------------------------------------------------
    def _pack(self, x, u, x0):
        assert len(x.shape) == 2 and len(u.shape) == 2, "x and u tensor must have dim 2"       # discret.py:15
        x0 = np.asarray(x0, np.float64)
        z = np.concatenate([np.asarray(x, np.float64).reshape(-1), np.asarray(u, np.float64).reshape(-1)])
        return z, x0.reshape(-1)

    def _first_order(self, x, u, x0, want):
        """residual (+ Jacobian values) at one iterate; with ``cache_mode`` the pair is computed together and
        memoised on the exact input bytes (the reference memoises k1..k4 on a str() hash, rk4.py:20-43)."""
        z, x0v = self._pack(x, u, x0)
        if not self.cache_mode:
            return self.evaluator.eval_host(z, x0v, want=want)
        key = z.tobytes() + x0v.tobytes()
        for k, v in self._cache:
            if k == key:
                return v
        out = {k: v.copy() for k, v in self.evaluator.eval_host(z, x0v, want=("resid", "jac")).items()}
        self._cache.append((key, out))
        del self._cache[:-self._cache_size]
        return out

    # ---- reference interface ------------------------------------------------------------------------------------
    def forward(self, x, u, x0, p=None, tvp=None):
        assert len(np.shape(x0)) == 1, "x0 shape must have dim 1"                            # discret.py:17
        self._set_exogenous(p, tvp)
        return self._first_order(x, u, x0, ("resid",))["resid"][0].copy()

    def jacobian(self, x, u, x0, p=None, tvp=None):
        ev = self.evaluator
        self._set_exogenous(p, tvp)
        vals = self._first_order(x, u, x0, ("jac",))["jac"][0]
        J = np.zeros((ev.m, ev.n))
        J[ev.jac_rows, ev.jac_cols] = vals
        return J

    def hessian_blocks(self, x, u, x0, p=None, tvp=None):
        """per-step, per-output second derivatives (H, x_dim, d, d) -- what rk4.py:266 calls ``model_H``."""
        self._set_exogenous(p, tvp)
        z, x0v = self._pack(x, u, x0)
        _, _, Hb = self.evaluator.eval_blocks(z[None], x0v[None])
        return Hb[0].cpu().numpy()

    def hessian(self, x, u, x0, p=None, tvp=None):
        H, xd, ud = self.H, self.model.x_dim, self.model.u_dim
        n = H * (xd + ud)
        if self.rolling:
            # banded model Hessian (H, x, n, n) over (x_{t-1} rows, u rows) -> shift the state index by one step (x_0 is data),
            # the slicing of discret.py:61-81 / unity.py:61-81
            xprev = np.concatenate([np.asarray(x0, np.float64).reshape(1, -1), np.asarray(x, np.float64)], axis=0)[:-1]
            mh = self.model.hessian(xprev, np.asarray(u, np.float64)).reshape(-1, n, n)
            out = np.zeros_like(mh)
            sx = xd * H
            out[:, :sx - xd, :sx - xd] = mh[:, xd:sx, xd:sx]
            out[:, sx:, sx:] = mh[:, sx:, sx:]
            out[:, :sx - xd, sx:] = mh[:, xd:sx, sx:]
            out[:, sx:, :sx - xd] = mh[:, sx:, xd:sx]
            return out
        blk = self.hessian_blocks(x, u, x0, p, tvp)
        out = np.zeros((H, xd, n, n))
        off = xd * H
        for t in range(H):                                   # same scatter as rk4.py:270-283 / discret.py:70-78
            cu = slice(off + t * ud, off + (t + 1) * ud)
            out[t, :, cu, cu] = blk[t, :, xd:, xd:]
            if t > 0:
                cx = slice(xd * (t - 1), xd * t)
                out[t, :, cx, cx] = blk[t, :, :xd, :xd]
                out[t, :, cx, cu] = blk[t, :, :xd, xd:]
                out[t, :, cu, cx] = blk[t, :, xd:, :xd]
        return out.reshape(-1, n, n)

    def _compute_hessianstructure(self):
        H, xd, ud = self.H, self.model.x_dim, self.model.u_dim
        n = H * (xd + ud)
        m = np.zeros((n, n))
        if self.rolling:
            from ..rolling import rolling_structure
            st = rolling_structure(H, xd, ud, self.model.rolling_window, self.KIND, None, self.model.forward_rolling)
            m[st["hes_rows"], st["hes_cols"]] = 1.0
            m[st["hes_cols"], st["hes_rows"]] = 1.0
            return m
        for t in range(H):
            cu = slice(H * xd + t * ud, H * xd + (t + 1) * ud)
            m[cu, cu] = 1.0
            if t > 0:
                cx = slice((t - 1) * xd, t * xd)
                m[cx, cx] = 1.0
                m[cx, cu] = 1.0
                m[cu, cx] = 1.0
        return m


class CudaDiscretIntegrator(_CudaIntegrator):
    """``x_{t-1} + f(x_{t-1}, u_t) - x_t`` (reference integrator/discret.py:8-81)."""
    KIND = "discrete"

    def __init__(self, model, H):
        super().__init__(model, H)


class CudaUnityIntegrator(_CudaIntegrator):
    """``f(x_{t-1}, u_t) - x_t`` (reference integrator/unity.py:9-81)."""
    KIND = "unity"

    def __init__(self, model, H):
        super().__init__(model, H)


class CudaRK4Integrator(_CudaIntegrator):
    """classical RK4 on ``xdot = f(x, u)``, zero-order hold on u (reference integrator/rk4.py:46-285)."""
    KIND = "rk4"

    def __init__(self, model, H, DT, cache_mode=False, cache_size=2):
        super().__init__(model, H, DT=DT, cache_mode=cache_mode, cache_size=cache_size)


# reference spellings
DiscretIntegrator = CudaDiscretIntegrator
UnityIntegrator = CudaUnityIntegrator
RK4Integrator = CudaRK4Integrator
