"""The batched interior-point solver bodies (pyneuralempc_b200/csrc/nempc_solver.cuh) executed on the host by the test-only
emulation, against (1) the numpy statement of the same algorithm (oracle/solver_np.py) and (2) SciPy SLSQP on the oracle
callbacks.  CPU only; the GPU run of nempc_solve is checked in tests/test_gpu_solver.py."""
import warnings

import numpy as np
import pytest
from scipy.optimize import Bounds, minimize

import hostsim_util as hs
from oracle.blocks_np import BlockEvaluator
from oracle.dense_ref import DenseIntegrator, DenseIpoptProblem
from oracle.mlp_np import MLP, DenseModelView
from oracle.objectives_np import SeparableQuadraticObjective
from oracle.solver_np import BatchedIPM

CASES = [("unity", "lv", 2, 1, 10, None, 1e9), ("rk4", [3, 30, 30, 2], 2, 1, 20, 0.1, 5.0), ("rk4", [5, 32, 32, 4], 4, 1, 15, 0.05, 5.0),
         ("discrete", [16, 24, 24, 12], 12, 4, 6, None, 5.0), ("unity", [5, 20, 20, 4], 4, 1, 12, None, 2.0)]


def _setup(kind, dims, x, u, H, DT, xb, lv_weights, seed=1):
    rng = np.random.default_rng(seed)
    if dims == "lv":
        mlp = MLP(lv_weights, 2, 1)
        obj = SeparableQuadraticObjective.tracking(H, 2, 1, [1.0, 1.0], [0.1], x_ref=np.array([0.5, -0.7]))
        lb = np.array([-np.inf, -np.inf] * H + [-1.0] * H); ub = np.array([1.0, np.inf] * H + [0.2] * H)     # run.py:72-74
        X0 = np.vstack([[0.66, -0.9], rng.uniform(-0.8, 0.8, (4, 2))])
    else:
        mlp = MLP.glorot(dims, x, u, seed=3)
        W, b = mlp.weights[-1]; mlp.weights[-1] = (W * 0.5, b * 0.5)
        obj = SeparableQuadraticObjective.tracking(H, x, u, np.ones(x), 0.05 * np.ones(u), x_ref=rng.uniform(-0.3, 0.3, (H, x)))
        lb = np.array([-xb] * (H * x) + [-0.3] * (H * u)); ub = np.array([xb] * (H * x) + [0.3] * (H * u))
        X0 = rng.uniform(-0.8, 0.8, (5, x))
    return mlp, obj, lb, ub, X0


@pytest.mark.parametrize("kind,dims,x,u,H,DT,xb", CASES)
def test_solver_bodies_match_numpy_statement(kind, dims, x, u, H, DT, xb, lv_weights):
    mlp, obj, lb, ub, X0 = _setup(kind, dims, x, u, H, DT, xb, lv_weights)
    ev = BlockEvaluator(mlp, kind, H, DT=DT, objective=obj)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Zr, lr, info = BatchedIPM(ev, lb, ub).solve(X0)
    Z, lam, st = hs.solve(mlp, kind, H, DT, obj, X0, lb, ub)
    assert info["converged"].all() and (st["status"] == 0).all()
    np.testing.assert_array_equal(st["iterations"], info["iterations"])          # same algorithm, same iteration counts
    np.testing.assert_allclose(Z, Zr, atol=1e-8)
    np.testing.assert_allclose(lam, lr, atol=1e-6)
    out = ev.evaluate(Z, X0, lam, 1.0)
    assert np.abs(out["resid"]).max() < 1e-6
    assert (Z >= lb - 1e-12).all() and (Z <= ub + 1e-12).all()


def test_solver_finds_the_slsqp_optimum(lv_weights):
    """the shipped Lotka-Volterra setting (fixture network, bounds of run.py:72-74): same optimum as SciPy SLSQP."""
    mlp, obj, lb, ub, X0 = _setup("unity", "lv", 2, 1, 10, None, 0, lv_weights)
    Z, lam, st = hs.solve(mlp, "unity", 10, None, obj, X0, lb, ub)
    ev = BlockEvaluator(mlp, "unity", 10, objective=obj)
    f = ev.evaluate(Z, X0, lam, 1.0)["obj"]
    for b in range(3):
        pb = DenseIpoptProblem(X0[b], obj, DenseIntegrator(DenseModelView(mlp), 10, "unity"))
        x_init = np.concatenate([np.tile(X0[b], 10), np.zeros(10)])
        r = minimize(pb.objective, x_init, method="SLSQP", jac=pb.gradient, bounds=Bounds(lb, ub),
                     constraints=[{"type": "eq", "fun": pb.constraints, "jac": pb.jacobian}], options={"maxiter": 300, "ftol": 1e-12})
        assert r.success and abs(r.fun - f[b]) < 1e-6 and np.abs(r.x - Z[b]).max() < 1e-4


def test_infeasible_problems_fail_cleanly():
    """state bounds that the dynamics cannot respect: status FAILED (or iteration limit), all values finite --
    the reference returns Optimizer.FAIL / (None, None) in that situation (optimizer/ipopt.py:191-195)."""
    rng = np.random.default_rng(1)
    H, x, u = 25, 2, 1
    mlp = MLP.glorot([3, 30, 30, 2], x, u, seed=3)
    W, b = mlp.weights[-1]; mlp.weights[-1] = (W * 0.5, b * 0.5)
    obj = SeparableQuadraticObjective.tracking(H, x, u, np.ones(x), 0.05 * np.ones(u), x_ref=rng.uniform(-0.3, 0.3, (H, x)))
    lb = np.array([-1.0] * (H * x) + [-0.3] * (H * u)); ub = np.array([1.0] * (H * x) + [0.3] * (H * u))
    X0 = rng.uniform(-0.8, 0.8, (6, x))
    Z, lam, st = hs.solve(mlp, "discrete", H, None, obj, X0, lb, ub, max_iter=80)
    assert np.isfinite(Z).all() and np.isfinite(lam).all()
    assert set(st["status"]) <= {0, 1, 2} and (st["status"] != 0).any() and (st["status"] == 0).any()


def test_warm_start_takes_fewer_iterations(lv_weights):
    mlp, obj, lb, ub, X0 = _setup("unity", "lv", 2, 1, 10, None, 0, lv_weights)
    Z, lam, st = hs.solve(mlp, "unity", 10, None, obj, X0, lb, ub)
    Z2, _, st2 = hs.solve(mlp, "unity", 10, None, obj, X0, lb, ub, Z_init=Z)
    assert (st2["status"] == 0).all() and st2["iterations"].sum() <= st["iterations"].sum()
    np.testing.assert_allclose(Z2, Z, atol=1e-5)
