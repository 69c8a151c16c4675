"""CPU restatement of the reference's rolling-window (NARX) model layer.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's CPU leg).

Follows ``KerasTFModelRollingInput`` (``/root/reference/pyNeuralEMPC/model/tensorflow.py:112-340``) with the network derivatives taken
from ``oracle.mlp_np.MLP`` in closed form instead of TensorFlow (absent here: parity against TensorFlow's autodiff itself is unpinned,
like for the plain model; what IS pinned is everything around it -- ``tests/golden/ref_rolling_*.npz`` come from the reference's
unmodified ``DiscretIntegrator`` / ``UnityIntegrator`` / ``IpoptProblem`` driven by this class through a subclass of the reference's
``Model``):

* ``_gather_input`` / ``_gather_input_V2``  (tensorflow.py:187-236)  history-extended arrays and the sliding windows, both orders;
* ``rolling_input``                          (tensorflow.py:112-129)  network input ``[x window | u window]``;
* ``jacobian``                               (tensorflow.py:247-275)  per-row network Jacobian times the window projection, history
                                                                      columns dropped, columns ``[all x | all u]``;
* ``hessian``                                (tensorflow.py:298-340)  ``P^T H P`` with the 0/1 projection matrix built exactly like the
                                                                      reference's ``project_mat`` loop, then the same column selection.
The product code (pyneuralempc_b200/rolling.py) uses gather-code tables instead of projection matrices; the two agree or a test fails.
``RollingBlockEvaluator`` is the sparse NLP assembly on top (dense reference-literal integrator slicing, then the structure gather)."""
from __future__ import annotations

import numpy as np

from .mlp_np import MLP


class RollingMLP:
    def __init__(self, weights, x_dim, u_dim, rolling_window=2, forward_rolling=True, activation="tanh", dtype=np.float64):
        self.x_dim, self.u_dim, self.p_dim, self.tvp_dim = int(x_dim), int(u_dim), 0, 0
        self.rolling_window, self.forward_rolling = int(rolling_window), bool(forward_rolling)
        self.dw = self.rolling_window * (self.x_dim + self.u_dim)
        self.net = MLP(weights, self.x_dim, self.dw - self.x_dim, activation=activation, dtype=dtype)   # the window network, dw inputs
        self.weights, self.activation = self.net.weights, activation
        self.prev_x = self.prev_u = None

    def set_prev_data(self, x_prev, u_prev, tvp_prev=None):                                           # tensorflow.py:174-185
        w = self.rolling_window
        assert x_prev.shape == (w - 1, self.x_dim) and u_prev.shape == (w - 1, self.u_dim)
        self.prev_x, self.prev_u = np.asarray(x_prev, np.float64), np.asarray(u_prev, np.float64)

    # ---- windows ---------------------------------------------------------------------------------------------------------
    def _windows(self, x, u):                                                                          # tensorflow.py:203-236
        assert self.prev_x is not None and self.prev_u is not None
        w = self.rolling_window
        xe = np.concatenate([self.prev_x, np.asarray(x, np.float64)], axis=0)
        ue = np.concatenate([self.prev_u, np.asarray(u, np.float64)], axis=0)
        N = x.shape[0]
        if self.forward_rolling:
            xr = np.stack([xe[i:i + w].reshape(-1) for i in range(N)], axis=0)
            ur = np.stack([ue[i:i + w].reshape(-1) for i in range(N)], axis=0)
        else:
            xr = np.stack([xe[i:i + w][::-1].reshape(-1) for i in range(N)], axis=0)
            ur = np.stack([ue[i:i + w][::-1].reshape(-1) for i in range(N)], axis=0)
        return np.concatenate([xr, ur], axis=1)

    def _projection(self, N):
        """0/1 matrix (N * dw, (N + w - 1) * d) taking the history-extended variables [all x_ext | all u_ext] to the stacked window
        inputs -- the ``project_mat`` loop of tensorflow.py:313-317 (forward order), rows permuted for the reversed order."""
        w, xd, ud = self.rolling_window, self.x_dim, self.u_dim
        Ne = N + w - 1
        P = np.zeros((N * self.dw, Ne * (xd + ud)))
        for t in range(N):
            for k in range(w):
                pos = k if self.forward_rolling else w - 1 - k
                for c in range(xd):
                    P[t * self.dw + pos * xd + c, (t + k) * xd + c] = 1.0
                for c in range(ud):
                    P[t * self.dw + w * xd + pos * ud + c, Ne * xd + (t + k) * ud + c] = 1.0
        return P

    def _keep(self, N):
        """columns of the extended variable vector that belong to the arguments (history dropped): tensorflow.py:269-273, 331-335"""
        w, xd, ud = self.rolling_window, self.x_dim, self.u_dim
        Ne = N + w - 1
        return list(range(xd * (w - 1), xd * Ne)) + list(range(xd * Ne + ud * (w - 1), xd * Ne + ud * Ne))

    # ---- the reference's Model interface -------------------------------------------------------------------------------------
    def forward(self, x, u, p=None, tvp=None):
        return self.net.forward_z(self._windows(x, u))

    def jacobian(self, x, u, p=None, tvp=None):
        N = x.shape[0]
        _, J, _ = self.net.blocks(self._windows(x, u), need_hessian=False)          # (N, x, dw)
        P = self._projection(N)
        full = np.zeros((N * self.x_dim, P.shape[1]))
        for t in range(N):
            full[t * self.x_dim:(t + 1) * self.x_dim] = J[t] @ P[t * self.dw:(t + 1) * self.dw]
        return full[:, self._keep(N)]

    dense_jacobian = jacobian

    def hessian(self, x, u, p=None, tvp=None):
        N = x.shape[0]
        _, _, Hs = self.net.blocks(self._windows(x, u))                              # (N, x, dw, dw)
        P = self._projection(N)
        keep = self._keep(N)
        out = np.zeros((N, self.x_dim, len(keep), len(keep)))
        for t in range(N):
            Pt = P[t * self.dw:(t + 1) * self.dw][:, keep]                            # (dw, N d)
            out[t] = np.einsum("ac,pab,bd->pcd", Pt, Hs[t], Pt)
        return out

    dense_hessian = hessian


class RollingBlockEvaluator:
    """sparse NLP values of the rolling-window transcription for a batch: the dense reference-literal assembly
    (``oracle.dense_ref.DenseIntegrator`` slices any banded model the way integrator/discret.py:32-81 does) gathered at the analytic
    structure.  Small sizes only (dense O(H^3))."""

    def __init__(self, model, kind, H, objective=None):
        from .dense_ref import DenseIntegrator
        assert kind in ("discrete", "unity")
        self.model, self.kind, self.H, self.objective = model, kind, int(H), objective
        self.integ = DenseIntegrator(model, H, kind)
        xd, ud, w = model.x_dim, model.u_dim, model.rolling_window
        n = H * (xd + ud)
        self.n, self.m = n, H * xd
        # structure: variables of constraint block t = states x_{t-w+1..t} with index >= 1, controls u_{t-w+1..t} with index >= 0
        jm = np.zeros((self.m, n), bool)
        hm = np.zeros((n, n), bool)
        for t in range(H):
            cols = []
            for k in range(w):
                s = t - k
                if s >= 1:
                    cols += list(range((s - 1) * xd, s * xd))
                if s >= 0:
                    cols += list(range(H * xd + s * ud, H * xd + (s + 1) * ud))
            rows = slice(t * xd, (t + 1) * xd)
            jm[rows, cols] = True
            jm[rows, t * xd:(t + 1) * xd] |= np.eye(xd, dtype=bool)
            hm[np.ix_(cols, cols)] = True
        if objective is not None:
            hm |= objective.hessianstructure() != 0
        self.jac_rows, self.jac_cols = np.nonzero(jm)
        self.hes_rows, self.hes_cols = np.nonzero(np.tril(hm))

    def evaluate(self, Z, X0, lam=None, obj_factor=1.0):
        Z, X0 = np.atleast_2d(Z), np.atleast_2d(X0)
        B, H, xd, ud = Z.shape[0], self.H, self.model.x_dim, self.model.u_dim
        sig = np.broadcast_to(np.asarray(obj_factor, np.float64), (B,))
        out = {"resid": np.zeros((B, self.m)), "jac_vals": np.zeros((B, len(self.jac_rows)))}
        if lam is not None:
            out["hes_vals"] = np.zeros((B, len(self.hes_rows)))
        if self.objective is not None:
            out["obj"], out["grad"] = np.zeros(B), np.zeros((B, self.n))
        for b in range(B):
            s, u = Z[b, :H * xd].reshape(H, xd), Z[b, H * xd:].reshape(H, ud)
            out["resid"][b] = self.integ.forward(s, u, X0[b])
            out["jac_vals"][b] = self.integ.jacobian(s, u, X0[b])[self.jac_rows, self.jac_cols]
            if lam is not None:
                Hd = np.einsum("i,ijk->jk", lam[b], self.integ.hessian(s, u, X0[b]))
                if self.objective is not None:
                    Hd = Hd + sig[b] * self.objective.hessian(s, u)
                out["hes_vals"][b] = Hd[self.hes_rows, self.hes_cols]
            if self.objective is not None:
                out["obj"][b] = self.objective.forward(s, u)
                out["grad"][b] = self.objective.gradient(s, u)
        return out
