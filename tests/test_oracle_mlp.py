"""Pins oracle.mlp_np (the closed-form restatement of the TensorFlow autodiff results used by
KerasTFModel, model/tensorflow.py:49-109) against two independent derivations: torch.func autodiff
and central finite differences.  CPU only."""
import numpy as np
import pytest

from oracle.mlp_np import MLP, DenseModelView


def _torch_net(mlp):
    import torch
    Ws = [(torch.tensor(W, dtype=torch.float64), torch.tensor(b, dtype=torch.float64)) for W, b in mlp.weights]
    act = {"tanh": torch.tanh, "sigmoid": torch.sigmoid, "softplus": torch.nn.functional.softplus}[mlp.activation]

    def f(z):
        h = z
        for W, b in Ws[:-1]:
            h = act(h @ W + b)
        return h @ Ws[-1][0] + Ws[-1][1]
    return f


@pytest.mark.parametrize("dims,xd,ud,act", [([3, 30, 30, 2], 2, 1, "tanh"), ([5, 16, 12, 9, 4], 4, 1, "tanh"),
                                            ([16, 24, 12], 12, 4, "tanh"), ([3, 7, 2], 2, 1, "sigmoid"),
                                            ([3, 7, 5, 2], 2, 1, "softplus")])
def test_blocks_match_torch_autodiff(dims, xd, ud, act):
    import torch
    from torch.func import hessian, jacrev, vmap
    mlp = MLP.glorot(dims, xd, ud, seed=3, activation=act)
    z = np.random.default_rng(0).uniform(-1.5, 1.5, size=(11, xd + ud))
    f, J, Hs = mlp.blocks(z)
    net = _torch_net(mlp)
    zt = torch.tensor(z, dtype=torch.float64)
    np.testing.assert_allclose(f, vmap(net)(zt).numpy(), atol=1e-14)
    np.testing.assert_allclose(J, vmap(jacrev(net))(zt).numpy(), atol=1e-13)
    np.testing.assert_allclose(Hs, vmap(hessian(net))(zt).numpy(), atol=1e-12)


def test_blocks_match_finite_differences(lv_weights):
    mlp = MLP(lv_weights, 2, 1)
    z = np.random.default_rng(1).uniform(-1, 1, size=(5, 3))
    _, J, Hs = mlp.blocks(z)
    eps = 1e-5
    for c in range(3):
        dz = np.zeros(3); dz[c] = eps
        fp, Jp, _ = mlp.blocks(z + dz); fm, Jm, _ = mlp.blocks(z - dz)
        np.testing.assert_allclose((fp - fm) / (2 * eps), J[:, :, c], atol=1e-9)
        np.testing.assert_allclose((Jp - Jm) / (2 * eps), Hs[:, :, :, c], atol=1e-9)


def test_dense_layouts_follow_keras_tf_model(lv_weights):
    """(N*x, N*d) Jacobian with columns [all x | all u]; (N, x, N*d, N*d) Hessian, same order on
    both trailing axes (model/tensorflow.py:68-73,101-107)."""
    mlp = MLP(lv_weights, 2, 1)
    rng = np.random.default_rng(2)
    N = 4
    x, u = rng.uniform(-1, 1, (N, 2)), rng.uniform(-1, 1, (N, 1))
    view = DenseModelView(mlp)
    _, J, Hs = mlp.blocks(np.concatenate([x, u], 1))
    dj, dh = view.jacobian(x, u), view.hessian(x, u)
    assert dj.shape == (N * 2, N * 3) and dh.shape == (N, 2, N * 3, N * 3)
    for i in range(N):
        cols = [2 * i, 2 * i + 1, 2 * N + i]
        np.testing.assert_array_equal(dj[2 * i:2 * i + 2][:, cols], J[i])
        np.testing.assert_array_equal(dh[i][:, cols][:, :, cols], Hs[i])
    assert np.count_nonzero(dj) == N * 6 and np.count_nonzero(dh) == N * 2 * 9
    np.testing.assert_allclose(view.forward(x, u), mlp.forward_z(np.concatenate([x, u], 1)))
