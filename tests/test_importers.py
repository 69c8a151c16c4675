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
        importers.from_keras_model(Model([mk(3, 4, "gelu"), mk(4, 2, "linear")]))                         # unsupported activation
    with pytest.raises(ValueError):
        importers.from_torch_sequential(torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.GELU(), torch.nn.Linear(4, 2)))
    assert importers.from_keras_model(Model([mk(3, 4, "relu"), mk(4, 2, "linear")]))[1] == "relu"
    assert importers.from_torch_sequential(torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.ReLU(), torch.nn.Linear(4, 2)))[1] == "relu"


def test_reference_module_paths_and_names():
    """user scripts import pyNeuralEMPC.integrator.rk4.RK4Integrator, model.tensorflow.KerasTFModel, objective.jax.JAXObjectifFunc ...
    (examples/lotka_volterra/run.py:56-84, test.py:5): the same dotted paths exist here"""
    import pyneuralempc_b200 as nEMPC
    from pyneuralempc_b200.integrator import discret, rk4, unity
    from pyneuralempc_b200.model import base as mbase, jax as mjax, tensorflow as mtf
    from pyneuralempc_b200.objective import base as obase, jax as ojax
    assert rk4.RK4Integrator is nEMPC.integrator.RK4Integrator and discret.DiscretIntegrator is nEMPC.integrator.DiscretIntegrator
    assert unity.UnityIntegrator is nEMPC.integrator.UnityIntegrator
    assert issubclass(mtf.KerasTFModel, mbase.Model) and issubclass(ojax.JAXObjectifFunc, obase.ObjectiveFunc)
    assert nEMPC.optimizer.Slsqp and nEMPC.optimizer.Ipopt and nEMPC.constraints.DomainConstraint and nEMPC.controller.NMPC
    with pytest.raises(NotImplementedError):
        mjax.DiffDiscretJaxModel(lambda x, u, p=None, tvp=None: x, 2, 1)


def test_keras_tf_model_takes_a_keras_like_object_or_a_weight_list():
    from pyneuralempc_b200.model.tensorflow import KerasTFModel
    rng = np.random.default_rng(0)
    ws = [(rng.standard_normal((3, 8)), rng.standard_normal(8)), (rng.standard_normal((8, 2)), rng.standard_normal(2))]

    class Act:
        def __init__(self, n): self.__name__ = n

    class Layer:
        def __init__(self, W, b, a): self.W, self.b, self.activation, self.name = W, b, Act(a), "dense"
        def get_weights(self): return [self.W, self.b]

    class Keras:
        layers = [Layer(*ws[0], "relu"), Layer(*ws[1], "linear")]

    m = KerasTFModel(Keras(), x_dim=2, u_dim=1)
    assert m.activation == "relu" and m.x_dim == 2 and m.u_dim == 1 and len(m.weights) == 2
    m2 = KerasTFModel(ws, 2, 1)
    assert m2.activation == "tanh"
    with pytest.raises(ValueError):
        KerasTFModel(ws, 3, 1)                                   # output width != x_dim (tensorflow.py:20-21)
    with pytest.raises(NotImplementedError):
        KerasTFModel(ws, 2, 1, standardScaler=object())
    import pickle
    assert pickle.loads(pickle.dumps(m)).model is None           # the Keras object is not pickled (tensorflow.py:31-37)


def test_jax_objectif_func_identifies_the_shipped_costs():
    """the two costs the reference ships (run.py:79-87, test.py:55-60) are recognised from the callable alone; a cost with cross terms is refused"""
    from pyneuralempc_b200.objective.jax import JAXObjectifFunc
    H, xd, ud = 5, 2, 1
    cost_vec = np.full(H, 1.1)
    lotka = JAXObjectifFunc(lambda x, u, p=None, tvp=None: np.sum(u.reshape(-1) * cost_vec.reshape(-1)))
    lotka.prepare(H, xd, ud)
    assert np.allclose(lotka.lin, [0.0] * (H * xd) + [1.1] * H) and not lotka.quad.any() and lotka.offset == 0.0
    setp = JAXObjectifFunc(lambda x, u, p=None, tvp=None: np.sum((u - 2.0) ** 2).astype(np.float32))
    setp.prepare(H, xd, ud)
    assert np.allclose(setp.quad, [0.0] * (H * xd) + [1.0] * H) and np.allclose(setp.lin[H * xd:], -4.0) and abs(setp.offset - 4.0 * H) < 1e-5
    assert (setp.hessianstructure() == np.diag([0.0] * (H * xd) + [1.0] * H)).all()
    with pytest.raises(NotImplementedError):
        JAXObjectifFunc(lambda x, u, p=None, tvp=None: np.sum(np.diff(u[:, 0]) ** 2)).prepare(H, xd, ud)     # rate penalty: cross terms
    with pytest.raises(NotImplementedError):
        JAXObjectifFunc(lambda x, u, p=None, tvp=None: np.sum(x ** 4)).prepare(H, xd, ud)


def _dense_stack(rng, dims):
    return [(rng.standard_normal((a, b)).astype(np.float32), rng.standard_normal(b).astype(np.float32)) for a, b in zip(dims[:-1], dims[1:])]


def test_h5lite_reads_keras_h5_and_keras_archive(tmp_path):
    """the dependency-free HDF5 reader (pyneuralempc_b200/h5lite.py) and the .h5 / .keras importers on files built by the test-side
    mini writer (tests/h5mini_writer.py): weights, layer order, activations, error paths"""
    import h5mini_writer as W
    from pyneuralempc_b200 import h5lite
    from pyneuralempc_b200.model.tensorflow import KerasTFModel
    rng = np.random.default_rng(0)
    ws = _dense_stack(rng, [3, 30, 30, 2])
    p = W.write_keras_h5(str(tmp_path / "m.h5"), ws, ["tanh", "tanh", "linear"])
    f = h5lite.H5File(p)
    assert sorted(f.members()) == ["model_config_blob", "model_weights"]
    assert [n.decode() for n in f.attributes(f.members()["model_weights"])["layer_names"]] == ["dense", "dense_1", "dense_2"]
    got, act = importers.from_keras_h5(p)
    assert act == "tanh" and len(got) == 3
    for (Wg, bg), (Wr, br) in zip(got, ws):
        np.testing.assert_array_equal(Wg, Wr.astype(np.float64)); np.testing.assert_array_equal(bg, br.astype(np.float64))
    m = KerasTFModel(p, x_dim=2, u_dim=1)
    assert m.activation == "tanh" and m.weights[1][0].shape == (30, 30)
    # a weights-only file (save_weights) carries no config: the activation has to be given
    p2 = W.write_keras_h5(str(tmp_path / "w.h5"), ws, None, with_config=False)
    with pytest.raises(ValueError):
        importers.from_keras_h5(p2)
    assert importers.from_keras_h5(p2, "softplus")[1] == "softplus"
    # relu stack through the Keras-3 archive
    ws2 = _dense_stack(rng, [5, 16, 16, 16, 4])
    p3 = W.write_keras_archive(str(tmp_path / "m.keras"), ws2, ["relu", "relu", "relu", "linear"])
    got3, act3 = importers.load_any(p3)
    assert act3 == "relu" and [w.shape for w, _ in got3] == [(5, 16), (16, 16), (16, 16), (16, 4)]
    np.testing.assert_array_equal(got3[2][0], ws2[2][0].astype(np.float64))
    with pytest.raises(ValueError):
        h5lite.H5File(b"not an hdf5 file at all")
    with pytest.raises(ValueError):
        importers.load_any(str(tmp_path / "m.unknown"))
