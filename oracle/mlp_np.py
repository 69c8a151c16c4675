"""Analytic feed-forward network oracle (numpy).  TEST INFRASTRUCTURE ONLY.

Restates what the reference obtains from TensorFlow for a Keras ``Sequential`` of
``Dense`` layers (``/root/reference/pyNeuralEMPC/model/tensorflow.py``):

* ``forward``           <- ``KerasTFModel.forward``  (:49-51, ``model.predict``)
* ``dense_jacobian``    <- ``KerasTFModel.jacobian`` (:53-75, ``GradientTape.jacobian``
                           then column reorder to ``[all x | all u]``)
* ``dense_hessian``     <- ``KerasTFModel.hessian``  (:89-109, one ``tf.hessians`` per
                           (sample, output) mask, both trailing axes reordered)

TensorFlow itself is a third-party dependency that is absent here (unpinned in
``setup.py:20``; the fixture was saved by Keras 2.4.0), so the autodiff results are
restated in closed form: for pre-activations ``a_l``, tangents ``T_l = d a_l / d z``
and per-output adjoints ``G_l = d f / d h_l``

    J   = W_out^T diag(s'(a_L)) T_L
    H_p = sum_l T_l^T diag(s''(a_l) * G_l[:, p]) T_l

Keras kernel layout: ``W[in][out]``, ``y = x @ W + b``.
"""
from __future__ import annotations

import numpy as np

ACTIVATIONS = ("tanh", "sigmoid", "softplus", "relu")


def _act(name, a):
    """value, first and second derivative of the activation at ``a``."""
    if name == "tanh":
        t = np.tanh(a)
        s1 = 1.0 - t * t
        return t, s1, -2.0 * t * s1
    if name == "sigmoid":
        s = 1.0 / (1.0 + np.exp(-a))
        s1 = s * (1.0 - s)
        return s, s1, s1 * (1.0 - 2.0 * s)
    if name == "softplus":
        s = 1.0 / (1.0 + np.exp(-a))
        return np.logaddexp(0.0, a), s, s * (1.0 - s)
    if name == "relu":                      # piecewise linear; derivative 0 at a = 0 like tf.nn.relu's gradient
        on = (a > 0.0).astype(a.dtype)
        return a * on, on, np.zeros_like(a)
    raise ValueError(f"unknown activation {name!r}")


class MLP:
    """Dense network ``d -> h_1 -> ... -> h_L -> x_dim`` with a linear output layer."""

    def __init__(self, weights, x_dim, u_dim, activation="tanh", dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.weights = [(np.asarray(W, self.dtype), np.asarray(b, self.dtype)) for W, b in weights]
        self.x_dim, self.u_dim = int(x_dim), int(u_dim)
        self.d = self.x_dim + self.u_dim
        self.activation = activation
        assert self.weights[0][0].shape[0] == self.d, "first layer fan-in must be x_dim+u_dim"
        assert self.weights[-1][0].shape[1] == self.x_dim, "last layer fan-out must be x_dim"
        for (Wa, ba), (Wb, _) in zip(self.weights[:-1], self.weights[1:]):
            assert Wa.shape[1] == Wb.shape[0] == ba.shape[0]

    # ---- construction helpers -------------------------------------------------
    @staticmethod
    def glorot(layer_dims, x_dim, u_dim, seed=0, activation="tanh", dtype=np.float64, bias_scale=0.1):
        """Glorot-uniform kernels (the Keras default of the fixture) and U(-b, b) biases."""
        rng = np.random.default_rng(seed)
        ws = []
        for fan_in, fan_out in zip(layer_dims[:-1], layer_dims[1:]):
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            W = rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)
            b = rng.uniform(-bias_scale, bias_scale, size=(fan_out,)).astype(np.float32)
            ws.append((W, b))
        return MLP(ws, x_dim, u_dim, activation=activation, dtype=dtype)

    @property
    def layer_dims(self):
        return [self.d] + [W.shape[1] for W, _ in self.weights]

    def astype(self, dtype):
        return MLP(self.weights, self.x_dim, self.u_dim, self.activation, dtype)

    # ---- per-sample blocks -----------------------------------------------------
    def forward_z(self, z):
        """``z``: (N, d) stacked ``[x, u]`` rows -> (N, x_dim)."""
        h = np.asarray(z, self.dtype)
        for W, b in self.weights[:-1]:
            h = _act(self.activation, h @ W + b)[0]
        W, b = self.weights[-1]
        return h @ W + b

    def blocks(self, z, need_hessian=True):
        """value (N,x), Jacobian (N,x,d) and per-output Hessians (N,x,d,d) at rows ``z``."""
        z = np.asarray(z, self.dtype)
        N = z.shape[0]
        h = z
        T = np.broadcast_to(np.eye(self.d, dtype=self.dtype), (N, self.d, self.d))  # d h / d z, (N, width, d)
        tangents, s1s, s2s = [], [], []
        for W, b in self.weights[:-1]:
            a = h @ W + b
            Ta = np.einsum("io,nic->noc", W, T)          # d a_l / d z
            h, s1, s2 = _act(self.activation, a)
            tangents.append(Ta)
            s1s.append(s1)
            s2s.append(s2)
            T = s1[:, :, None] * Ta
        Wo, bo = self.weights[-1]
        f = h @ Wo + bo
        J = np.einsum("jp,njc->npc", Wo, T)
        if not need_hessian:
            return f, J, None
        # reverse sweep of the per-output adjoints G_l = d f / d h_l  (N, width_l, x)
        Hs = np.zeros((N, self.x_dim, self.d, self.d), self.dtype)
        G = np.broadcast_to(Wo, (N,) + Wo.shape)
        for l in range(len(self.weights) - 2, -1, -1):
            coef = s2s[l][:, :, None] * G                                     # (N, width, x)
            Hs += np.einsum("njp,njc,nje->npce", coef, tangents[l], tangents[l])
            if l > 0:
                W = self.weights[l][0]                                        # (width_{l-1}, width_l)
                G = np.einsum("ij,njp->nip", W, s1s[l][:, :, None] * G)
        return f, J, Hs

    # ---- the reference's Model interface (dense layouts) ------------------------
    def forward(self, x, u, p=None, tvp=None):
        return self.forward_z(np.concatenate([x, u], axis=1))

    def dense_jacobian(self, x, u):
        """(N*x_dim, N*d), columns ``[x_0..x_{N-1} | u_0..u_{N-1}]`` (tensorflow.py:68-73)."""
        N = x.shape[0]
        xd, ud = self.x_dim, self.u_dim
        _, J, _ = self.blocks(np.concatenate([x, u], axis=1), need_hessian=False)
        out = np.zeros((N * xd, N * (xd + ud)), self.dtype)
        for i in range(N):
            out[i * xd:(i + 1) * xd, i * xd:(i + 1) * xd] = J[i, :, :xd]
            out[i * xd:(i + 1) * xd, N * xd + i * ud:N * xd + (i + 1) * ud] = J[i, :, xd:]
        return out

    def dense_hessian(self, x, u):
        """(N, x_dim, N*d, N*d) with both trailing axes in ``[all x | all u]`` order
        (tensorflow.py:101-107)."""
        N = x.shape[0]
        xd, ud = self.x_dim, self.u_dim
        _, _, Hs = self.blocks(np.concatenate([x, u], axis=1))
        out = np.zeros((N, xd, N * (xd + ud), N * (xd + ud)), self.dtype)
        for i in range(N):
            sx = slice(i * xd, (i + 1) * xd)
            su = slice(N * xd + i * ud, N * xd + (i + 1) * ud)
            out[i, :, sx, sx] = Hs[i, :, :xd, :xd]
            out[i, :, sx, su] = Hs[i, :, :xd, xd:]
            out[i, :, su, sx] = Hs[i, :, xd:, :xd]
            out[i, :, su, su] = Hs[i, :, xd:, xd:]
        return out


class ExoMLP:
    """Network over the gathered input ``[x, u, tvp, p]`` (``model/tensorflow.py:39-47``), bound to the exogenous rows of the
    samples it is evaluated at.  Derivatives are taken w.r.t. the whole input and the tvp / p columns dropped, as the reference
    does (``model/tensorflow.py:65-66, 97-98``).  The reference's ``p`` branch is mis-shaped (``:44-45``, its own TODO); the
    intended meaning -- the same ``p`` row appended to every sample -- is what is restated.

    Duck-types the part of :class:`MLP` that ``blocks_np`` / ``dense_ref`` use (``x_dim, u_dim, d, dtype, forward_z, blocks,
    forward, dense_jacobian, dense_hessian``)."""

    def __init__(self, weights, x_dim, u_dim, tvp_dim=0, p_dim=0, activation="tanh", dtype=np.float64, ext=None):
        self.tvp_dim, self.p_dim = int(tvp_dim), int(p_dim)
        self.full = MLP(weights, x_dim, u_dim + self.tvp_dim + self.p_dim, activation, dtype)
        self.x_dim, self.u_dim = int(x_dim), int(u_dim)
        self.d = self.x_dim + self.u_dim
        self.dtype, self.activation, self.weights = self.full.dtype, activation, self.full.weights
        self.ext = ext

    def astype(self, dtype):
        return ExoMLP(self.weights, self.x_dim, self.u_dim, self.tvp_dim, self.p_dim, self.activation, dtype, self.ext)

    def bind(self, tvp=None, p=None, B=1, H=None):
        """``tvp``: (H, tvp_dim) shared by the B problems or (B, H, tvp_dim); ``p``: (p_dim,) or (B, p_dim); ``H`` is needed when there
        is no tvp to read it from.  Rows are laid
        out like the steps ``blocks_np.BlockEvaluator`` stacks: problem-major, step-minor."""
        cols = []
        if self.tvp_dim:
            tvp = np.asarray(tvp, np.float64)
            if tvp.ndim == 2:
                tvp = np.broadcast_to(tvp, (B,) + tvp.shape)
            H = tvp.shape[1]
            cols.append(tvp.reshape(-1, self.tvp_dim))
        if self.p_dim:
            p = np.asarray(p, np.float64)
            if p.ndim == 1:
                p = np.broadcast_to(p, (B, self.p_dim))
            reps = H if H is not None else 1
            cols.append(np.repeat(p, reps, axis=0))
        ext = np.concatenate(cols, axis=1) if cols else None
        return ExoMLP(self.weights, self.x_dim, self.u_dim, self.tvp_dim, self.p_dim, self.activation, self.dtype, ext)

    def _gather(self, z):
        z = np.asarray(z, self.dtype)
        if self.ext is None:
            assert self.tvp_dim + self.p_dim == 0, "bind() the exogenous rows first"
            return z
        assert self.ext.shape[0] == z.shape[0], "one exogenous row per sample"
        return np.concatenate([z, self.ext.astype(self.dtype)], axis=1)

    def forward_z(self, z):
        return self.full.forward_z(self._gather(z))

    def blocks(self, z, need_hessian=True):
        f, J, Hs = self.full.blocks(self._gather(z), need_hessian=need_hessian)
        d = self.d
        return f, J[:, :, :d], (None if Hs is None else Hs[:, :, :d, :d])

    def forward(self, x, u, p=None, tvp=None):
        return self.forward_z(np.concatenate([x, u], axis=1))

    dense_jacobian = MLP.dense_jacobian
    dense_hessian = MLP.dense_hessian


class DenseModelView:
    """Duck-typed ``Model`` (``model/base.py:3-18`` as actually called: ``discret.py:27,48,64``,
    ``rk4.py:69-72,86,97``) serving the dense layouts from an :class:`MLP`.  ``net_dtype``
    float32 mimics the reference's casts (``tensorflow.py:91``; Keras predicts in float32)."""

    def __init__(self, mlp, p_dim=0, tvp_dim=0):
        self.mlp = mlp
        self.x_dim, self.u_dim = mlp.x_dim, mlp.u_dim
        self.p_dim, self.tvp_dim = p_dim, tvp_dim

    def forward(self, x, u, p=None, tvp=None):
        return self.mlp.forward(x, u)

    def jacobian(self, x, u, p=None, tvp=None):
        return self.mlp.dense_jacobian(x, u)

    def hessian(self, x, u, p=None, tvp=None):
        return self.mlp.dense_hessian(x, u)


# ---- fixture ------------------------------------------------------------------
LV_H5_SHA256 = "fc3ee0ee3e2f0e15936acab756ff9c898fcb26d64491eb91eec4506bc8768a10"
_LV_H5_LAYOUT = (  # dataset -> (shape, byte offset); HDF5 superblock v0, contiguous little-endian f32
    ("dense/kernel", (3, 30), 9416), ("dense/bias", (30,), 9776),
    ("dense_1/kernel", (30, 30), 14216), ("dense_1/bias", (30,), 9896),
    ("dense_2/kernel", (30, 2), 10016), ("dense_2/bias", (2,), 10256),
)


def read_lv_fixture_h5(path):
    """Weights of ``examples/lotka_volterra/nn_model.h5`` (Keras 2.4.0 Sequential
    3 -> 30 tanh -> 30 tanh -> 2 linear) without h5py, by fixed offsets."""
    import hashlib
    buf = open(path, "rb").read()
    if hashlib.sha256(buf).hexdigest() != LV_H5_SHA256:
        raise ValueError("unexpected nn_model.h5 contents; offsets are only valid for the shipped fixture")
    arrs = {}
    for name, shape, off in _LV_H5_LAYOUT:
        arrs[name] = np.frombuffer(buf, "<f4", int(np.prod(shape)), off).reshape(shape).copy()
    return [(arrs["dense/kernel"], arrs["dense/bias"]),
            (arrs["dense_1/kernel"], arrs["dense_1/bias"]),
            (arrs["dense_2/kernel"], arrs["dense_2/bias"])]


def load_lv_fixture_npz(path):
    """The same weights from the committed ``tests/golden/lv_mlp_weights.npz``."""
    d = np.load(path)
    return [(d[f"W{i}"], d[f"b{i}"]) for i in range(3)]
