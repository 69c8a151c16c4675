"""``pyNeuralEMPC.model.tensorflow`` under its reference name (``model/tensorflow.py:8-109``): ``KerasTFModel(keras_model, x_dim, u_dim)``
with the network evaluated by the CUDA kernels instead of TensorFlow.

``model`` may be what the reference takes -- a live Keras ``Sequential`` of ``Dense`` layers (only ``layers`` / ``get_weights()`` /
``activation.__name__`` are read; TensorFlow is not imported here) -- or, since TensorFlow is not needed any more, the path of the
saved model itself (``.h5`` / ``.hdf5`` Keras file read by ``h5lite``, ``.npz`` with W0,b0,..., ``.safetensors``), a
``torch.nn.Sequential``, or a plain list of ``(kernel[in, out], bias[out])`` pairs."""
from __future__ import annotations

import os

from .. import importers
from . import CudaMLPModel, CudaMLPModelRollingInput


class KerasTFModel(CudaMLPModel):
    def __init__(self, model, x_dim: int, u_dim: int, p_dim=0, tvp_dim=0, standardScaler=None, activation=None, **cuda_options):
        if standardScaler is not None:
            raise NotImplementedError("This feature isn't supported yet !")        # tensorflow.py:11-12
        weights, act = importers.load_any(model, activation)
        super().__init__(weights, x_dim, u_dim, p_dim=p_dim, tvp_dim=tvp_dim, activation=act, **cuda_options)
        self.model = model if not isinstance(model, (str, os.PathLike)) else os.fspath(model)

    def __getstate__(self):                                                         # tensorflow.py:31-37: the Keras object is not pickled
        st = super().__getstate__()
        if not isinstance(st.get("model"), str):
            st["model"] = None
        return st


class KerasTFModelRollingInput(CudaMLPModelRollingInput):
    """``KerasTFModelRollingInput(model, x_dim, u_dim, rolling_window=2, forward_rolling=True)`` (model/tensorflow.py:131-340) on the GPU;
    ``model`` as for ``KerasTFModel``."""

    def __init__(self, model, x_dim: int, u_dim: int, p_dim=0, tvp_dim=0, rolling_window=2, forward_rolling=True, standardScaler=None,
                 activation=None, **cuda_options):
        if standardScaler is not None:
            raise NotImplementedError("This feature isn't supported yet !")
        weights, act = importers.load_any(model, activation)
        super().__init__(weights, x_dim, u_dim, p_dim=p_dim, tvp_dim=tvp_dim, rolling_window=rolling_window,
                         forward_rolling=forward_rolling, activation=act, **cuda_options)
        self.model = model if not isinstance(model, (str, os.PathLike)) else os.fspath(model)

    def __getstate__(self):
        st = super().__getstate__()
        if not isinstance(st.get("model"), str):
            st["model"] = None
        return st
