"""Arithmetic + output indexing of the device kernel bodies, executed on the HOST by the test-only emulation in
tests/hostsim (same nempc_generic.cuh / nempc_fast.cuh source the CUDA kernels compile), against the oracle.
This is a GPU-less development check; the parity tests proper are the `-m gpu` tests."""
import numpy as np
import pytest

import hostsim_util as hs
from oracle.blocks_np import BlockEvaluator, step_blocks
from oracle.mlp_np import MLP
from oracle.objectives_np import SeparableQuadraticObjective

CASES = [("discrete", [3, 30, 30, 2], 2, 1, 6, "tanh"), ("unity", [3, 30, 30, 2], 2, 1, 5, "tanh"),
         ("rk4", [3, 30, 30, 2], 2, 1, 7, "tanh"), ("rk4", [5, 12, 9, 7, 4], 4, 1, 4, "tanh"),
         ("rk4", [6, 40, 3], 3, 3, 3, "tanh"), ("discrete", [16, 20, 12], 12, 4, 2, "tanh"),
         ("rk4", [2, 6, 1], 1, 1, 1, "tanh"), ("rk4", [3, 9, 8, 2], 2, 1, 3, "sigmoid"),
         ("unity", [4, 6, 6, 6, 2], 2, 2, 2, "softplus"), ("rk4", [3, 9, 8, 2], 2, 1, 3, "relu")]


def _problem(kind, dims, xd, ud, H, act, seed=0, with_obj=True):
    rng = np.random.default_rng(seed)
    mlp = MLP.glorot(dims, xd, ud, seed=seed + 1, activation=act)
    n, m = H * (xd + ud), H * xd
    obj = None
    if with_obj:
        obj = SeparableQuadraticObjective.tracking(H, xd, ud, rng.uniform(0.5, 2, xd), rng.uniform(0.1, 1, ud),
                                                   x_ref=rng.uniform(-1, 1, (H, xd)))
    B = 3
    return mlp, obj, rng.uniform(-1, 1, (B, n)), rng.uniform(-1, 1, (B, xd)), rng.standard_normal((B, m)), rng.uniform(0.5, 1.5, B)


@pytest.mark.parametrize("kind,dims,xd,ud,H,act", CASES)
def test_generic_body_f64(kind, dims, xd, ud, H, act):
    mlp, obj, Z, X0, lam, sig = _problem(kind, dims, xd, ud, H, act)
    ref = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    got = hs.run(mlp, kind, H, 0.1, Z, X0, lam, sig, quad=obj.quad, compute_f64=True)
    np.testing.assert_allclose(got["resid"], ref["resid"], atol=1e-13)
    np.testing.assert_allclose(got["jac_vals"], ref["jac_vals"], atol=1e-12)
    np.testing.assert_allclose(got["hes_vals"], ref["hes_vals"], atol=1e-11)


@pytest.mark.parametrize("kind,dims,xd,ud,H,act", CASES[:5])
def test_generic_body_f32(kind, dims, xd, ud, H, act):
    mlp, obj, Z, X0, lam, sig = _problem(kind, dims, xd, ud, H, act, seed=3)
    ref = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    got = hs.run(mlp, kind, H, 0.1, Z, X0, lam, sig, quad=obj.quad, compute_f64=False)
    for k in ("resid", "jac_vals", "hes_vals"):
        scale = max(1.0, np.abs(ref[k]).max())
        assert np.abs(got[k] - ref[k]).max() < 1e-5 * scale, k


@pytest.mark.parametrize("kind", ("discrete", "unity", "rk4"))
@pytest.mark.parametrize("h", (30, 32, 16))
def test_fast_body(kind, h, lv_weights):
    H = 6
    mlp = MLP(lv_weights, 2, 1) if h == 30 else MLP.glorot([3, h, h, 2], 2, 1, seed=h)
    rng = np.random.default_rng(h)
    obj = SeparableQuadraticObjective.tracking(H, 2, 1, [1.0, 0.5], [0.3], x_ref=rng.uniform(-1, 1, (H, 2)))
    B = 4
    Z, X0 = rng.uniform(-1, 1, (B, H * 3)), rng.uniform(-1, 1, (B, 2))
    lam, sig = rng.standard_normal((B, H * 2)), rng.uniform(0.5, 1.5, B)
    ref = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    got = hs.run(mlp, kind, H, 0.1, Z, X0, lam, sig, quad=obj.quad, compute_f64=False, kernel="fast")
    for k in ("resid", "jac_vals", "hes_vals"):
        scale = max(1.0, np.abs(ref[k]).max())
        assert np.abs(got[k] - ref[k]).max() < 1e-5 * scale, k
    # the reduced modes give the same numbers
    g1 = hs.run(mlp, kind, H, 0.1, Z, X0, None, 1.0, quad=obj.quad, compute_f64=False, kernel="fast", want_hes=False)
    np.testing.assert_allclose(g1["jac_vals"], got["jac_vals"], rtol=1e-6, atol=1e-7)
    g0 = hs.run(mlp, kind, H, 0.1, Z, X0, None, 1.0, compute_f64=False, kernel="fast", want_jac=False, want_hes=False)
    np.testing.assert_allclose(g0["resid"], got["resid"], rtol=1e-6, atol=1e-7)


def test_blocks_and_model_modes():
    mlp, _, Z, X0, _, _ = _problem("rk4", [5, 12, 9, 4], 4, 1, 4, "tanh")
    ev = BlockEvaluator(mlp, "rk4", 4, DT=0.1)
    _, _, xprev = ev.split(Z, X0)
    u = Z[:, 4 * 4:].reshape(3, 4, 1)
    zz = np.concatenate([xprev, u], axis=2).reshape(-1, 5)
    pred, AB, Hb = step_blocks(mlp, "rk4", 0.1, zz)
    p2, AB2, Hb2 = hs.run(mlp, "rk4", 4, 0.1, Z, X0, what="blocks")
    np.testing.assert_allclose(p2, pred + zz[:, :4], atol=1e-13)
    np.testing.assert_allclose(AB2, AB + np.eye(4, 5)[None], atol=1e-12)
    np.testing.assert_allclose(Hb2, Hb, atol=1e-11)
    f, J, Hs = mlp.blocks(zz)
    f2, J2, Hs2 = hs.run(mlp, "rk4", 4, 0.1, zz, None, what="model")
    np.testing.assert_allclose(f2, f, atol=1e-13)
    np.testing.assert_allclose(J2, J, atol=1e-12)
    np.testing.assert_allclose(Hs2, Hs, atol=1e-11)


EXO_CASES = [("discrete", 2, 1, 2, 1, 5), ("rk4", 2, 1, 2, 0, 4), ("unity", 3, 2, 0, 2, 3), ("rk4", 4, 1, 1, 1, 3)]


def _exo_problem(kind, xd, ud, td, pd, H, seed, B=3, shared=False):
    from oracle.mlp_np import ExoMLP
    rng = np.random.default_rng(seed)
    full = MLP.glorot([xd + ud + td + pd, 10, 7, xd], xd, ud + td + pd, seed=seed + 1)
    exo = ExoMLP(full.weights, xd, ud, td, pd)
    tvp = None if td == 0 else rng.uniform(-1, 1, (H, td) if shared else (B, H, td))
    p = None if pd == 0 else rng.uniform(-1, 1, (pd,) if shared else (B, pd))
    n, m = H * (xd + ud), H * xd
    return exo, tvp, p, rng.uniform(-1, 1, (B, n)), rng.uniform(-1, 1, (B, xd)), rng.standard_normal((B, m))


@pytest.mark.parametrize("kind,xd,ud,td,pd,H", EXO_CASES)
@pytest.mark.parametrize("shared", (False, True))
def test_generic_body_with_tvp_and_p(kind, xd, ud, td, pd, H, shared):
    """exogenous (tvp / p) network inputs: a per-step shift of the first layer's pre-activation, not a decision variable"""
    exo, tvp, p, Z, X0, lam = _exo_problem(kind, xd, ud, td, pd, H, 21, shared=shared)
    ref = BlockEvaluator(exo.bind(tvp, p, B=Z.shape[0], H=H), kind, H, DT=0.1).evaluate(Z, X0, lam)
    got = hs.run(exo, kind, H, 0.1, Z, X0, lam, compute_f64=True, tvp=tvp, p=p)
    np.testing.assert_allclose(got["resid"], ref["resid"], atol=1e-13)
    np.testing.assert_allclose(got["jac_vals"], ref["jac_vals"], atol=1e-12)
    np.testing.assert_allclose(got["hes_vals"], ref["hes_vals"], atol=1e-11)
    got32 = hs.run(exo, kind, H, 0.1, Z, X0, lam, compute_f64=False, tvp=tvp, p=p)
    for k in ("resid", "jac_vals", "hes_vals"):
        assert np.abs(got32[k] - ref[k]).max() < 1e-5 * max(1.0, np.abs(ref[k]).max()), k
