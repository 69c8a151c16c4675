"""Container-only checks against the UNMODIFIED reference under /root/reference (imported through oracle/shim.py with its absent
third-party imports stubbed).  Skipped where the reference tree does not exist (the GPU box): there the committed goldens stand in.
What is checked: the committed golden files are what the reference computes TODAY (bit-identical regeneration of a sample of them by
the committed generator), and the oracle's dense restatement equals the live reference on fresh random inputs."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import shim
from oracle.dense_ref import DenseIntegrator
from oracle.mlp_np import MLP, DenseModelView, load_lv_fixture_npz

pytestmark = pytest.mark.skipif(not shim.source_tree_available(), reason="/root/reference is not present on this machine")

HERE = os.path.dirname(os.path.abspath(__file__))


def _generator():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _same(rec, path):
    g = np.load(path)
    assert set(g.files) == set(rec), sorted(set(g.files) ^ set(rec))
    for k in g.files:
        np.testing.assert_array_equal(np.asarray(rec[k]), g[k], err_msg=k)


def test_goldens_regenerate_bit_identically_from_the_reference(golden_dir):
    mk, ref = _generator(), shim.load_reference()
    lv = MLP(load_lv_fixture_npz(os.path.join(golden_dir, "lv_mlp_weights.npz")), 2, 1, dtype=np.float64)
    _same(mk.record(ref, lv, "rk4", 6, 102, "tracking", True), os.path.join(golden_dir, "ref_rk4_H6.npz"))
    _same(mk.record(ref, lv, "discrete", 25, 200, "setpoint", False), os.path.join(golden_dir, "ref_discrete_H25.npz"))
    _same(mk.record_wide(ref, [3, 128, 128, 2], 2, 1, "rk4", 6, 700, 21), os.path.join(golden_dir, "ref_rk4_w128_H6.npz"))
    _same(mk.record_constraints(ref, lv), os.path.join(golden_dir, "ref_constraints_H6.npz"))


@pytest.mark.parametrize("kind", ("discrete", "unity", "rk4"))
def test_dense_restatement_equals_the_live_reference(kind, lv_weights):
    """oracle/dense_ref.py (the timed CPU baseline, kind "port") against the reference's own integrator on inputs no golden holds"""
    ref = shim.load_reference()
    mk = _generator()
    mlp = MLP(lv_weights, 2, 1, dtype=np.float64)
    H = 7
    rng = np.random.default_rng(77)
    xs, us, x0 = rng.uniform(-1, 1, (H, 2)), rng.uniform(-1, 1, (H, 1)), rng.uniform(-1, 1, 2)
    theirs = mk.make_integrator(ref, kind, shim.make_reference_model(mlp), H)
    ours = DenseIntegrator(DenseModelView(mlp), H, kind, DT=mk.DT_RK4)
    np.testing.assert_allclose(ours.forward(xs, us, x0), theirs.forward(xs, us, x0), rtol=0, atol=1e-13)
    np.testing.assert_allclose(ours.jacobian(xs, us, x0), theirs.jacobian(xs, us, x0), rtol=0, atol=1e-13)
    np.testing.assert_allclose(ours.hessian(xs, us, x0), theirs.hessian(xs, us, x0), rtol=0, atol=1e-12)


def test_product_h5_reader_on_the_reference_fixture(golden_dir):
    """pyneuralempc_b200.h5lite (no h5py, no TensorFlow) on examples/lotka_volterra/nn_model.h5: the Dense stack and its activations"""
    from pyneuralempc_b200 import importers
    from pyneuralempc_b200.model.tensorflow import KerasTFModel
    path = os.path.join(shim.REFERENCE_ROOT, "examples", "lotka_volterra", "nn_model.h5")
    weights, act = importers.from_keras_h5(path)
    g = np.load(os.path.join(golden_dir, "lv_mlp_weights.npz"))
    assert act == "tanh" and len(weights) == 3
    for i, (W, b) in enumerate(weights):
        np.testing.assert_array_equal(W.astype(np.float32), g[f"W{i}"])
        np.testing.assert_array_equal(b.astype(np.float32), g[f"b{i}"])
    m = KerasTFModel(path, x_dim=2, u_dim=1)                       # what run.py:56,68 does through TensorFlow
    assert m.activation == "tanh" and m.weights[1][0].shape == (30, 30)


def test_cuda_model_passes_the_reference_isinstance_check(lv_weights):
    """integrator/base.py:16-17 of the reference requires isinstance(model, pyNeuralEMPC.model.base.Model)"""
    ref = shim.load_reference()
    from pyneuralempc_b200.model import CudaMLPModel
    from pyneuralempc_b200.model.base import Model
    m = CudaMLPModel(lv_weights, 2, 1)                             # no device is touched before the first evaluation
    assert isinstance(m, ref.model.base.Model) and isinstance(m, Model)
    integ = ref.integrator.discret.DiscretIntegrator(m, 5)
    assert integ.model is m and integ.H == 5


def test_reference_bytecode_build_and_sourceless_import(tmp_path):
    """oracle/build_ref.py: the reference compiled to bytecode (what travels to the GPU box under oracle/_ref) imports without its sources
    and computes what the source tree computes"""
    import subprocess
    import sys
    from oracle.build_ref import build_reference_bytecode
    out = tmp_path / "_ref"
    n = build_reference_bytecode(out=str(out))
    import zipfile
    names = zipfile.ZipFile(out / "pyNeuralEMPC_bytecode.zip").namelist()
    assert n >= 15 and "pyNeuralEMPC/integrator/rk4.pyc" in names
    assert all(nm.endswith(".pyc") for nm in names) and not list(out.rglob("*.py"))     # bytecode only: no reference source is copied
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from oracle import shim; shim.REFERENCE_ROOT = '/nonexistent'; shim.BYTECODE_ROOT = %r;"
            "ref = shim.load_reference(); assert ref.__file__.endswith('.pyc'), ref.__file__;"
            "from oracle.mlp_np import MLP, load_lv_fixture_npz;"
            "mlp = MLP(load_lv_fixture_npz(%r), 2, 1);"
            "integ = ref.integrator.rk4.RK4Integrator(shim.make_reference_model(mlp), 6, 0.1);"
            "g = np.load(%r); z = g['z'];"
            "print(float(np.abs(integ.forward(z[:12].reshape(6, 2), z[12:].reshape(6, 1), g['x0']) - g['integrator_forward']).max()))") % (
        os.path.dirname(HERE), str(out / "pyNeuralEMPC_bytecode.zip"), os.path.join(HERE, "golden", "lv_mlp_weights.npz"), os.path.join(HERE, "golden", "ref_rk4_H6.npz"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-800:]
    assert float(r.stdout.strip().splitlines()[-1]) == 0.0
