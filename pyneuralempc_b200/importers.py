"""Weight importers for ``CudaMLPModel`` (the user-facing step before the hot path).

The reference wraps a live Keras model (``model/tensorflow.py:9-29``; ``examples/lotka_volterra/run.py:56,68`` loads
``nn_model.h5``) and reads its layers through TensorFlow.  The CUDA path only needs the dense kernels and biases in Keras layout
``(kernel[in, out], bias[out])``; these helpers produce that list from the containers people actually have:

* ``from_keras_model``        a live ``keras.Sequential`` (duck-typed: ``get_weights()`` + layer activations), no TensorFlow import here;
* ``from_torch_sequential``   a ``torch.nn.Sequential`` of ``Linear`` / activation modules (``Linear.weight`` is ``[out, in]`` -> transposed);
* ``from_state_dict``         a mapping name -> array with ``*.weight`` / ``*.bias`` entries (torch checkpoints, safetensors);
* ``read_safetensors``        the safetensors container itself (8-byte header length, JSON header, raw little-endian tensors);
* ``from_keras_h5``           a Keras ``.h5`` / ``.hdf5`` file (``model.save`` or ``save_weights``) read by ``h5lite`` -- no h5py, no TensorFlow;
* ``from_keras_archive``      a Keras-3 ``.keras`` zip archive (``config.json`` + ``model.weights.h5``);
* ``load_any``                dispatch on what the caller holds (object, path, or a ready weight list).

Each returns ``(weights, activation)``; every hidden layer must use the same activation (tanh / sigmoid / softplus) and the last layer
must be linear -- what ``nempc_create`` supports."""
from __future__ import annotations

import json
import struct

import numpy as np

_ACT_NAMES = {"tanh": "tanh", "sigmoid": "sigmoid", "softplus": "softplus", "relu": "relu", "linear": None, "identity": None, None: None}
_ST_DTYPES = {"F64": "<f8", "F32": "<f4", "F16": "<f2", "I64": "<i8", "I32": "<i4", "I16": "<i2", "I8": "i1", "U8": "u1", "BOOL": "?"}


def _check(weights, acts):
    hidden = set(a for a in acts[:-1])
    if len(hidden) > 1:
        raise ValueError(f"all hidden layers must share one activation, got {sorted(map(str, hidden))}")
    act = hidden.pop() if hidden else "tanh"
    if act is None:
        raise ValueError("hidden layers need a non-linear activation (tanh, sigmoid or softplus)")
    if acts[-1] is not None:
        raise ValueError("the last layer must be linear (model/tensorflow.py expects the raw next-state prediction)")
    for (Wa, ba), (Wb, _) in zip(weights[:-1], weights[1:]):
        if Wa.shape[1] != Wb.shape[0] or ba.shape != (Wa.shape[1],):
            raise ValueError("layer shapes do not chain")
    return weights, act


def from_keras_model(model):
    """``keras.Sequential`` of ``Dense`` layers -> (weights, activation).  Only attributes are read: ``layers``, ``get_weights()``,
    ``activation.__name__`` -- so any object with that surface works (and no TensorFlow is needed to call this)."""
    weights, acts = [], []
    for layer in model.layers:
        ws = layer.get_weights()
        if not ws:
            continue                                           # InputLayer, Dropout, ...
        if len(ws) != 2 or np.ndim(ws[0]) != 2:
            raise ValueError(f"layer {getattr(layer, 'name', layer)} is not a Dense layer with a bias")
        name = getattr(getattr(layer, "activation", None), "__name__", None)
        if name not in _ACT_NAMES:
            raise ValueError(f"unsupported activation {name!r}")
        weights.append((np.asarray(ws[0], np.float64), np.asarray(ws[1], np.float64)))
        acts.append(_ACT_NAMES[name])
    return _check(weights, acts)


def from_torch_sequential(seq):
    """``torch.nn.Sequential`` of ``Linear`` layers separated by ``Tanh`` / ``Sigmoid`` / ``Softplus`` modules."""
    weights, acts = [], []
    for mod in seq:
        cls = type(mod).__name__
        if cls == "Linear":
            if mod.bias is None:
                raise ValueError("Linear layers need a bias")
            weights.append((mod.weight.detach().cpu().double().numpy().T.copy(), mod.bias.detach().cpu().double().numpy().copy()))
            acts.append(None)
        elif cls in ("Tanh", "Sigmoid", "Softplus", "ReLU"):
            if not weights or acts[-1] is not None:
                raise ValueError("an activation module must follow a Linear layer")
            if cls == "Softplus" and (getattr(mod, "beta", 1) != 1):
                raise ValueError("only Softplus(beta=1) is supported")
            acts[-1] = cls.lower()
        elif cls in ("Identity", "Flatten"):
            continue
        else:
            raise ValueError(f"unsupported module {cls}")
    return _check(weights, acts)


def from_state_dict(state, activation="tanh", prefix=""):
    """mapping ``name -> array`` holding ``<prefix><k>.weight`` ([out, in], torch layout) and ``<prefix><k>.bias`` entries; layers are
    taken in the numeric order of ``k`` (``0.weight, 2.weight, ...`` of an ``nn.Sequential`` checkpoint)."""
    names = sorted((k for k in state if k.startswith(prefix) and k.endswith(".weight")),
                   key=lambda k: [int(t) if t.isdigit() else t for t in k[len(prefix):].split(".")])
    if not names:
        raise ValueError("no '*.weight' entries found")
    weights = []
    for k in names:
        W = np.asarray(state[k], np.float64)
        bk = k[:-len("weight")] + "bias"
        if W.ndim != 2 or bk not in state:
            raise ValueError(f"{k}: expected a 2-D weight with a matching bias")
        weights.append((W.T.copy(), np.asarray(state[bk], np.float64).copy()))
    return _check(weights, [activation] * (len(weights) - 1) + [None])


def read_safetensors(path):
    """minimal safetensors reader -> dict name -> numpy array (no dependency on the safetensors package)."""
    with open(path, "rb") as f:
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen).decode("utf-8"))
        blob = f.read()
    out = {}
    for name, meta in header.items():
        if name == "__metadata__":
            continue
        if meta["dtype"] not in _ST_DTYPES:
            raise ValueError(f"{name}: unsupported dtype {meta['dtype']}")
        lo, hi = meta["data_offsets"]
        out[name] = np.frombuffer(blob[lo:hi], dtype=_ST_DTYPES[meta["dtype"]]).reshape(meta["shape"]).copy()
    return out


def from_keras_h5(path, activation=None):
    """Keras HDF5 file (``examples/lotka_volterra/run.py:56`` loads ``nn_model.h5`` through TensorFlow; here: ``h5lite``).  The Dense
    activations come from the embedded model config; a weights-only file needs ``activation=``."""
    from . import h5lite
    weights, acts = h5lite.read_keras_dense_stack(path)
    if acts is None:
        if activation is None:
            raise ValueError(f"{path}: no model config in the file (save_weights?): pass activation=")
        acts = [activation] * (len(weights) - 1) + ["linear"]
    bad = [a for a in acts if a not in _ACT_NAMES]
    if bad:
        raise ValueError(f"unsupported activation(s) {bad}")
    return _check(weights, [_ACT_NAMES[a] for a in acts])


def from_keras_archive(path, activation=None):
    """Keras-3 ``.keras`` archive: a zip holding ``config.json`` (layer classes and activations) and ``model.weights.h5``
    (``layers/<name>/vars/0`` = kernel, ``vars/1`` = bias)."""
    import zipfile
    from . import h5lite
    with zipfile.ZipFile(path) as zf:
        cfg = json.loads(zf.read("config.json").decode("utf-8"))
        f = h5lite.H5File(zf.read("model.weights.h5"))
    dense = [l for l in cfg.get("config", {}).get("layers", []) if l.get("class_name") == "Dense"]
    by_layer = {}
    for p, a in f.walk():
        parts = p.split("/")
        if len(parts) >= 3 and parts[-2] == "vars":
            by_layer.setdefault("/".join(parts[:-2]), {})[parts[-1]] = f.dataset(a)
    weights, acts = [], []
    for l in dense:
        name = l["config"]["name"]
        key = next((k for k in by_layer if k.split("/")[-1] == name), None)
        if key is None or set(by_layer[key]) != {"0", "1"}:
            raise ValueError(f"{path}: no kernel/bias pair for Dense layer {name!r}")
        weights.append((by_layer[key]["0"].astype(np.float64), by_layer[key]["1"].astype(np.float64)))
        a = l["config"].get("activation", "linear")
        if a not in _ACT_NAMES:
            raise ValueError(f"unsupported activation {a!r}")
        acts.append(_ACT_NAMES[a])
    if not weights:
        raise ValueError(f"{path}: no Dense layers in config.json")
    return _check(weights, acts)


def load_any(model, activation=None):
    """(weights, activation) from whatever the caller holds: a path (.h5 / .hdf5 / .keras / .npz / .safetensors), a live Keras model,
    a ``torch.nn.Sequential``, a state dict, or a list of ``(kernel[in, out], bias[out])`` pairs."""
    import os
    if isinstance(model, (str, os.PathLike)):
        path = os.fspath(model)
        ext = os.path.splitext(path)[1].lower()
        if ext in (".h5", ".hdf5"):
            return from_keras_h5(path, activation)
        if ext == ".keras":
            return from_keras_archive(path, activation)
        if ext == ".safetensors":
            return from_state_dict(read_safetensors(path), activation or "tanh")
        if ext == ".npz":
            d = np.load(path)
            n = len([k for k in d.files if k.startswith("W")])
            return _check([(np.asarray(d[f"W{i}"], np.float64), np.asarray(d[f"b{i}"], np.float64)) for i in range(n)],
                          [activation or "tanh"] * (n - 1) + [None])
        raise ValueError(f"{path}: unknown model file type {ext!r}")
    if isinstance(model, dict):
        return from_state_dict(model, activation or "tanh")
    if isinstance(model, (list, tuple)):
        ws = [(np.asarray(W, np.float64), np.asarray(b, np.float64)) for W, b in model]
        return _check(ws, [activation or "tanh"] * (len(ws) - 1) + [None])
    if hasattr(model, "layers") and hasattr(model.layers[0] if len(model.layers) else None, "get_weights"):
        return from_keras_model(model)
    if type(model).__name__ == "Sequential" and hasattr(model, "__iter__"):
        return from_torch_sequential(model)
    raise ValueError(f"cannot read network weights from a {type(model).__name__}")
