"""Dynamics-model layer: the reference's ``Model`` contract with the network evaluated by CUDA kernels.

Mirrors ``/root/reference/pyNeuralEMPC/model/base.py:3-18`` (``Model``) and the behaviour of
``model/tensorflow.py:8-109`` (``KerasTFModel``): same constructor checks, same call signature as the
integrators use it (``forward(x, u, p=, tvp=)``), same dense output layouts:

* ``forward``  -> (N, x_dim)
* ``jacobian`` -> (N*x_dim, N*(x_dim+u_dim)), columns ``[x_0..x_{N-1} | u_0..u_{N-1}]`` (tensorflow.py:68-73)
* ``hessian``  -> (N, x_dim, N*d, N*d), same ordering on both trailing axes (tensorflow.py:101-107)

The per-sample blocks behind those dense arrays (``blocks``) come from ``nempc_model_eval``; the dense scatter is
only there for drop-in compatibility with reference-style integrators -- the CUDA integrators never use it.
"""
from __future__ import annotations

import numpy as np

from ..engine import NlpEvaluator


from .base import Model, adopt_reference_base  # noqa: E402,F401


class CudaMLPModel(Model):
    """Feed-forward network (Keras ``Sequential`` of ``Dense`` layers, linear last layer) on the GPU.

    ``weights``: list of ``(kernel[in, out], bias[out])`` in Keras layout.  ``dtype`` is the arithmetic type of the
    network ('float32' = what Keras/TensorFlow computes in, 'float64' for the 1e-10 parity mode); inputs and
    outputs are float64 like the reference's numpy arrays."""

    def __init__(self, weights, x_dim: int, u_dim: int, p_dim=0, tvp_dim=0, activation="tanh", dtype="float32",
                 device=0, kernel="auto", standardScaler=None):
        if standardScaler is not None:
            raise NotImplementedError("This feature isn't supported yet !")        # tensorflow.py:11-12
        p_dim, tvp_dim = int(p_dim or 0), int(tvp_dim or 0)
        weights = [(np.asarray(W, np.float64), np.asarray(b, np.float64)) for W, b in weights]
        if weights[-1][0].shape[1] != x_dim:                                       # tensorflow.py:20-21
            raise ValueError("Your model do not provide a suitable output dim ! \n It must get the same dim as the state dim.")
        if weights[0][0].shape[0] != x_dim + u_dim + p_dim + tvp_dim:              # tensorflow.py:23-24
            raise ValueError("Your model do not provide a suitable input dim ! \n It must get the same dim as the sum of all input vars (x, u, p, tvp).")
        adopt_reference_base()                  # when the reference package is loaded too: pass its isinstance(model, Model) check
        super().__init__(x_dim, u_dim, p_dim, tvp_dim)
        self.weights = weights
        self.activation, self.dtype, self.device, self.kernel = activation, dtype, device, kernel
        self._ev = None

    # ---- constructors -------------------------------------------------------------------------------------
    # ---- constructors from common containers (importers.py) -------------------------------------------------------------
    @classmethod
    def from_keras(cls, keras_model, x_dim, u_dim, **kw):
        """a live Keras ``Sequential`` of Dense layers -- the object the reference hands to KerasTFModel (model/tensorflow.py:9)."""
        from .. import importers
        weights, act = importers.from_keras_model(keras_model)
        return cls(weights, x_dim, u_dim, activation=act, **kw)

    @classmethod
    def from_safetensors(cls, path, x_dim, u_dim, activation="tanh", prefix="", **kw):
        """a safetensors checkpoint of an ``nn.Sequential`` (``<k>.weight`` [out, in], ``<k>.bias``)."""
        from .. import importers
        weights, act = importers.from_state_dict(importers.read_safetensors(path), activation, prefix)
        return cls(weights, x_dim, u_dim, activation=act, **kw)

    @classmethod
    def from_npz(cls, path, x_dim, u_dim, **kw):
        """weights stored as W0,b0,W1,b1,... (tests/golden/lv_mlp_weights.npz has this layout)."""
        d = np.load(path)
        n = len([k for k in d.files if k.startswith("W")])
        return cls([(d[f"W{i}"], d[f"b{i}"]) for i in range(n)], x_dim, u_dim, **kw)

    @classmethod
    def from_torch_sequential(cls, seq, x_dim, u_dim, **kw):
        """``torch.nn.Sequential`` of ``Linear`` (+ Tanh/Sigmoid/Softplus) modules; Linear stores ``[out, in]``."""
        import torch
        ws, act = [], kw.pop("activation", None)
        names = {torch.nn.Tanh: "tanh", torch.nn.Sigmoid: "sigmoid", torch.nn.Softplus: "softplus", torch.nn.ReLU: "relu"}
        for mod in seq:
            if isinstance(mod, torch.nn.Linear):
                b = mod.bias.detach().cpu().double().numpy() if mod.bias is not None else np.zeros(mod.out_features)
                ws.append((mod.weight.detach().cpu().double().numpy().T.copy(), b))
            elif type(mod) in names:
                if act not in (None, names[type(mod)]):
                    raise ValueError("mixed activations are not supported")
                act = names[type(mod)]
            else:
                raise ValueError(f"unsupported module {type(mod).__name__}")
        return cls(ws, x_dim, u_dim, activation=act or "tanh", **kw)

    def __getstate__(self):                      # picklable without the device handle (tensorflow.py:31-37)
        st = dict(self.__dict__)
        st["_ev"] = None
        return st

    # ---- evaluation ----------------------------------------------------------------------------------------------
    def evaluator(self):
        if self._ev is None:
            self._ev = NlpEvaluator(self.weights, self.x_dim, self.u_dim, 1, "unity", activation=self.activation,
                                    compute_dtype=self.dtype, io_dtype="float64", device=self.device, kernel=self.kernel,
                                    tvp_dim=self.tvp_dim, p_dim=self.p_dim)
        return self._ev

    def _gather_input(self, x, u, p=None, tvp=None):          # tensorflow.py:39-47
        """(x, u) rows go to the kernel as the differentiated input; tvp / p are handed over as exogenous rows (the network input is
        [x, u, tvp, p], its tvp / p derivative columns are never formed -- the slicing of tensorflow.py:65-66, 97-98)."""
        if (tvp is None) != (self.tvp_dim == 0) or (p is None) != (self.p_dim == 0):
            raise ValueError("p / tvp must be given exactly when the model declares p_dim / tvp_dim")
        if self.tvp_dim or self.p_dim:
            tvp = None if tvp is None else np.asarray(tvp, np.float64).reshape(np.asarray(x).shape[0], self.tvp_dim)
            p = None if p is None else np.asarray(p, np.float64).reshape(self.p_dim)
            self.evaluator().set_exogenous(tvp, p)
        return np.concatenate([np.asarray(x, np.float64), np.asarray(u, np.float64)], axis=1)

    def blocks(self, x, u, p=None, tvp=None, want_jac=True, want_hes=True):
        """per-sample value (N,x), Jacobian (N,x,d), per-output Hessian (N,x,d,d) as numpy arrays."""
        f, J, Hs = self.evaluator().model_eval(self._gather_input(x, u, p, tvp), want_jac, want_hes)
        c = lambda t: None if t is None else t.cpu().numpy()
        return c(f), c(J), c(Hs)

    def forward(self, x, u, p=None, tvp=None):
        return self.blocks(x, u, p, tvp, False, False)[0]

    def jacobian(self, x, u, p=None, tvp=None):
        N, xd, ud = x.shape[0], self.x_dim, self.u_dim
        _, J, _ = self.blocks(x, u, p, tvp, True, False)
        out = np.zeros((N * xd, N * (xd + ud)))
        for i in range(N):
            out[i * xd:(i + 1) * xd, i * xd:(i + 1) * xd] = J[i, :, :xd]
            out[i * xd:(i + 1) * xd, N * xd + i * ud:N * xd + (i + 1) * ud] = J[i, :, xd:]
        return out

    def hessian(self, x, u, p=None, tvp=None):
        N, xd, ud = x.shape[0], self.x_dim, self.u_dim
        _, _, Hs = self.blocks(x, u, p, tvp, True, True)
        out = np.zeros((N, xd, N * (xd + ud), N * (xd + ud)))
        for i in range(N):
            sx, su = slice(i * xd, (i + 1) * xd), slice(N * xd + i * ud, N * xd + (i + 1) * ud)
            out[i, :, sx, sx] = Hs[i, :, :xd, :xd]
            out[i, :, sx, su] = Hs[i, :, :xd, xd:]
            out[i, :, su, sx] = Hs[i, :, xd:, :xd]
            out[i, :, su, su] = Hs[i, :, xd:, xd:]
        return out


class CudaMLPModelRollingInput(Model):
    """Rolling-window (NARX) network on the GPU: the reference's ``KerasTFModelRollingInput`` (model/tensorflow.py:131-340) /
    ``DiffDiscretJaxModelRollingWindow`` (model/jax.py:93-259).  Row ``i`` of a call reads the ``rolling_window`` latest rows of the
    history-extended inputs (``set_prev_data`` provides the ``rolling_window - 1`` rows before the first one); the network has
    ``rolling_window * (x_dim + u_dim)`` inputs ordered ``[x window | u window]``, oldest row first (newest first when
    ``forward_rolling=False``).  Dense outputs have the layouts of ``CudaMLPModel`` with the band filled in:
    ``jacobian`` (N x_dim, N (x_dim + u_dim)), ``hessian`` (N, x_dim, N d, N d); history columns are dropped like the reference does."""

    def __init__(self, weights, x_dim: int, u_dim: int, p_dim=0, tvp_dim=0, rolling_window=2, forward_rolling=True, activation="tanh",
                 dtype="float32", device=0, standardScaler=None):
        if standardScaler is not None:
            raise NotImplementedError("This feature isn't supported yet !")                      # tensorflow.py:134-135
        if p_dim or tvp_dim:
            raise NotImplementedError("rolling-window models with p / tvp inputs are not supported (the reference's own TODO, tensorflow.py:114)")
        if not isinstance(rolling_window, int) or rolling_window < 1:
            raise ValueError("Your rolling windows need to be an integer gretter than 1.")          # tensorflow.py:151-152
        weights = [(np.asarray(W, np.float64), np.asarray(b, np.float64)) for W, b in weights]
        if weights[-1][0].shape[1] != x_dim:                                                       # tensorflow.py:145-146
            raise ValueError("Your Keras model do not provide a suitable output dim ! \n It must get the same dim as the state dim.")
        if weights[0][0].shape[0] != rolling_window * (x_dim + u_dim):
            raise ValueError("Your model do not provide a suitable input dim ! \n It must be rolling_window * (x_dim + u_dim).")
        adopt_reference_base()
        super().__init__(x_dim, u_dim, 0, 0)
        self.weights, self.activation, self.dtype, self.device = weights, activation, dtype, device
        self.rolling_window, self.forward_rolling = rolling_window, forward_rolling
        self.prev_x, self.prev_u, self.prev_tvp = None, None, None
        self._ev = None

    def __getstate__(self):                                  # tensorflow.py:162-169
        st = dict(self.__dict__)
        st["_ev"], st["prev_x"], st["prev_u"], st["prev_tvp"] = None, None, None, None
        return st

    def set_prev_data(self, x_prev, u_prev, tvp_prev=None):   # tensorflow.py:174-185
        w = self.rolling_window
        x_prev, u_prev = np.asarray(x_prev, np.float64), np.asarray(u_prev, np.float64)
        assert x_prev.shape == (w - 1, self.x_dim), f"Your x prev tensor must have the following shape {(w - 1, self.x_dim)} (received : {x_prev.shape})"
        assert u_prev.shape == (w - 1, self.u_dim), f"Your u prev tensor must have the following shape {(w - 1, self.u_dim)} (received : {u_prev.shape})"
        self.prev_x, self.prev_u = x_prev, u_prev

    def evaluator(self):
        if self._ev is None:
            dw = self.rolling_window * (self.x_dim + self.u_dim)
            self._ev = NlpEvaluator(self.weights, self.x_dim, dw - self.x_dim, 1, "unity", activation=self.activation,
                                    compute_dtype=self.dtype, io_dtype="float64", device=self.device, kernel="generic")
        return self._ev

    def _rows(self, x, u):
        """window rows (N, dw) of the network input and the (N, dw) column codes of its entries"""
        from ..rolling import window_columns
        assert (self.prev_x is not None) and (self.prev_u is not None), \
            "You must give history window with set_prev_data before calling any inferance function."       # tensorflow.py:189
        x, u = np.asarray(x, np.float64), np.asarray(u, np.float64)
        N = x.shape[0]
        code = window_columns(N, self.x_dim, self.u_dim, self.rolling_window, self.forward_rolling, model_level=True)
        flat = np.concatenate([x.reshape(-1), u.reshape(-1)])
        aux = np.concatenate([np.zeros(self.x_dim), self.prev_x.reshape(-1), self.prev_u.reshape(-1)])
        zin = np.where(code >= 0, flat[np.clip(code, 0, None)], aux[np.clip(-1 - code, 0, None)])
        return zin, code

    def blocks(self, x, u, want_jac=True, want_hes=True):
        zin, code = self._rows(x, u)
        f, J, Hs = self.evaluator().model_eval(zin, want_jac, want_hes)
        c = lambda t: None if t is None else t.cpu().numpy()
        return c(f), c(J), c(Hs), code

    def forward(self, x, u, p=None, tvp=None):
        return self.blocks(x, u, False, False)[0]

    def jacobian(self, x, u, p=None, tvp=None):
        N, xd, d = np.asarray(x).shape[0], self.x_dim, self.x_dim + self.u_dim
        _, J, _, code = self.blocks(x, u, True, False)
        out = np.zeros((N * xd, N * d))
        for i in range(N):
            keep = code[i] >= 0
            out[i * xd:(i + 1) * xd, code[i][keep]] = J[i][:, keep]
        return out

    def hessian(self, x, u, p=None, tvp=None):
        N, xd, d = np.asarray(x).shape[0], self.x_dim, self.x_dim + self.u_dim
        _, _, Hs, code = self.blocks(x, u, True, True)
        out = np.zeros((N, xd, N * d, N * d))
        for i in range(N):
            keep = np.nonzero(code[i] >= 0)[0]
            cols = code[i][keep]
            out[i][:, cols[:, None], cols[None, :]] = Hs[i][:, keep[:, None], keep[None, :]]
        return out
