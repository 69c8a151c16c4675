"""Weight importers (pyneuralempc_b200/importers.py): Keras-like objects, torch.nn.Sequential, state dicts, safetensors files
all give the Keras-layout list the C ABI takes (nempc_set_weights), and the oracle MLP built from them matches torch."""
import json
import struct

import numpy as np
import pytest
import torch

from oracle.mlp_np import MLP
from pyneuralempc_b200 import importers


def _seq(act=torch.nn.Tanh):
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(3, 7), act(), torch.nn.Linear(7, 5), act(), torch.nn.Linear(5, 2)).double()


@pytest.mark.parametrize("act,name", [(torch.nn.Tanh, "tanh"), (torch.nn.Sigmoid, "sigmoid"), (torch.nn.Softplus, "softplus")])
def test_torch_sequential_roundtrip(act, name):
    seq = _seq(act)
    weights, activation = importers.from_torch_sequential(seq)
    assert activation == name and [W.shape for W, _ in weights] == [(3, 7), (7, 5), (5, 2)]
    z = np.random.default_rng(0).uniform(-1, 1, (11, 3))
    ref = seq(torch.as_tensor(z)).detach().numpy()
    got = MLP(weights, 2, 1, activation=activation).forward_z(z)
    np.testing.assert_allclose(got, ref, atol=1e-12)


def test_state_dict_and_safetensors(tmp_path):
    seq = _seq()
    sd = {k: v.detach().numpy() for k, v in seq.state_dict().items()}
    w1, a1 = importers.from_state_dict(sd)
    w0, _ = importers.from_torch_sequential(seq)
    for (Wa, ba), (Wb, bb) in zip(w0, w1):
        np.testing.assert_array_equal(Wa, Wb); np.testing.assert_array_equal(ba, bb)
    # write a safetensors file by hand (float32 payload) and read it back
    header, blob = {}, b""
    for k, v in sd.items():
        raw = np.ascontiguousarray(v, "<f4").tobytes()
        header[k] = {"dtype": "F32", "shape": list(v.shape), "data_offsets": [len(blob), len(blob) + len(raw)]}
        blob += raw
    hj = json.dumps(header).encode()
    path = tmp_path / "net.safetensors"
    path.write_bytes(struct.pack("<Q", len(hj)) + hj + blob)
    w2, a2 = importers.from_state_dict(importers.read_safetensors(path))
    assert a2 == "tanh"
    for (Wa, _), (Wb, _) in zip(w0, w2):
        np.testing.assert_allclose(Wa, Wb, rtol=1e-6)


def test_keras_like_object_and_errors():
    class Act:
        def __init__(self, n): self.__name__ = n

    class Layer:
        def __init__(self, W, b, act): self.W, self.b, self.activation, self.name = W, b, Act(act), "dense"
        def get_weights(self): return [self.W, self.b]

    class Model:
        def __init__(self, layers): self.layers = layers

    rng = np.random.default_rng(1)
    mk = lambda i, o, a: Layer(rng.standard_normal((i, o)).astype(np.float32), rng.standard_normal(o).astype(np.float32), a)
    weights, act = importers.from_keras_model(Model([mk(3, 30, "tanh"), mk(30, 30, "tanh"), mk(30, 2, "linear")]))   # the LV fixture's shape
    assert act == "tanh" and [W.shape for W, _ in weights] == [(3, 30), (30, 30), (30, 2)] and weights[0][0].dtype == np.float64
    with pytest.raises(ValueError):
        importers.from_keras_model(Model([mk(3, 4, "tanh"), mk(4, 4, "sigmoid"), mk(4, 2, "linear")]))     # mixed activations
    with pytest.raises(ValueError):
        importers.from_keras_model(Model([mk(3, 4, "tanh"), mk(4, 2, "tanh")]))                           # non-linear last layer
    with pytest.raises(ValueError):
        importers.from_keras_model(Model([mk(3, 4, "relu"), mk(4, 2, "linear")]))                         # unsupported activation
    with pytest.raises(ValueError):
        importers.from_torch_sequential(torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.ReLU(), torch.nn.Linear(4, 2)))
