"""Rolling-window (NARX) models on the GPU (reference model/tensorflow.py:112-340, model/jax.py:93-259, test.py:20-79):
window gather kernel + nempc_model_eval + banded sparse assembly kernel against the goldens recorded from the reference's unmodified
integrators / IpoptProblem and against the oracle restatement on batches; drop-in classes under the reference's names; the test.py
known-answer solve."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from parity_metric import elem_err  # noqa: E402

from oracle.mlp_np import MLP  # noqa: E402
from oracle.objectives_np import SeparableQuadraticObjective  # noqa: E402
from oracle.rolling_np import RollingBlockEvaluator, RollingMLP  # noqa: E402

GOLDENS = ("ref_rolling_discrete_w2.npz", "ref_rolling_unity_w3.npz")


def _golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    return g, [(g[f"net_W{i}"], g[f"net_b{i}"]) for i in range(3)]


@pytest.mark.parametrize("name", GOLDENS)
@pytest.mark.parametrize("dtype,tol", (("float64", 1e-10), ("float32", 1e-5)))
def test_rolling_problem_callbacks_vs_reference_golden(golden_dir, name, dtype, tol):
    from pyneuralempc_b200 import integrator as I
    from pyneuralempc_b200.model.tensorflow import KerasTFModelRollingInput
    from pyneuralempc_b200.objective import CudaSeparableObjective
    from pyneuralempc_b200.optimizer.ipopt import CudaIpoptProblem
    g, ws = _golden(golden_dir, name)
    H, w, kind = int(g["H"]), int(g["rolling_window"]), str(g["kind"])
    model = KerasTFModelRollingInput(ws, 2, 1, rolling_window=w, forward_rolling=bool(g["forward_rolling"]), dtype=dtype)
    model.set_prev_data(g["x_prev"], g["u_prev"])
    integ = (I.DiscretIntegrator if kind == "discrete" else I.UnityIntegrator)(model, H)
    pb = CudaIpoptProblem(g["x0"], CudaSeparableObjective(g["obj_lin"], g["obj_quad"], g["obj_ref"]), [], integ, use_hessian=True)
    z = g["z"]
    r, c = pb.hessianstructure()
    np.testing.assert_array_equal(r, g["hes_rows"]); np.testing.assert_array_equal(c, g["hes_cols"])
    jr, jc = pb.jacobianstructure()
    J = np.zeros_like(g["jacobian"]); J[jr, jc] = pb.jacobian(z)
    assert elem_err(pb.constraints(z), g["constraints"]) < tol
    assert elem_err(J, g["jacobian"]) < tol
    assert elem_err(pb.hessian(z, g["lam"], float(g["sigma"])), g["hessian_values"]) < tol
    assert elem_err(pb.gradient(z), g["gradient"]) < 1e-12 and abs(pb.objective(z) - float(g["objective"])) < 1e-12
    # dense drop-in interfaces: the Model layouts (tensorflow.py:247-340) and the integrator's (m, n, n) Hessian
    s, u = z[:2 * H].reshape(H, 2), z[2 * H:].reshape(H, 1)
    xp = np.concatenate([g["x0"][None], s])[:-1]
    assert elem_err(model.forward(xp, u), g["model_forward"]) < tol
    assert elem_err(model.jacobian(xp, u), g["model_jacobian"]) < tol
    assert elem_err(model.hessian(xp, u), g["model_hessian"]) < 2 * tol
    assert elem_err(integ.hessian(s, u, g["x0"]), g["integrator_hessian"]) < 2 * tol
    assert elem_err(integ.jacobian(s, u, g["x0"]), g["jacobian"]) < tol and elem_err(integ.forward(s, u, g["x0"]), g["constraints"]) < tol
    st = integ.hessianstructure()
    dense = np.zeros_like(st); dense[r, c] = 1.0; dense[c, r] = 1.0
    quad_only = np.diag((g["obj_quad"] != 0).astype(float))
    assert ((st != 0) <= (dense != 0)).all() and ((dense - quad_only > 0) <= (st != 0)).all()


@pytest.mark.parametrize("kind,w,fwd,xd,ud,H,B", [("discrete", 2, True, 2, 1, 9, 7), ("unity", 3, False, 2, 1, 6, 5), ("discrete", 4, True, 2, 2, 5, 3),
                                                ("unity", 1, True, 3, 1, 4, 2)])
def test_rolling_batch_vs_oracle(kind, w, fwd, xd, ud, H, B):
    """a batch with PER-PROBLEM history rows against the oracle, problem by problem; float64 network -> 1e-10"""
    import torch
    from pyneuralempc_b200.rolling import RollingNlpEvaluator
    rng = np.random.default_rng(10 * w + H)
    dw = w * (xd + ud)
    net = MLP.glorot([dw, 12, 9, xd], xd, dw - xd, seed=w + 5)
    n, m = H * (xd + ud), H * xd
    obj = SeparableQuadraticObjective.tracking(H, xd, ud, rng.uniform(0.5, 2, xd), rng.uniform(0.1, 1, ud), x_ref=rng.uniform(-1, 1, (H, xd)))
    Z, X0, lam, sig = rng.uniform(-1, 1, (B, n)), rng.uniform(-1, 1, (B, xd)), rng.standard_normal((B, m)), rng.uniform(0.5, 1.5, B)
    XP, UP = rng.uniform(-1, 1, (B, w - 1, xd)), rng.uniform(-1, 1, (B, w - 1, ud))
    ev = RollingNlpEvaluator(net.weights, xd, ud, H, kind, w, forward_rolling=fwd, compute_dtype="float64")
    ev.set_objective(obj.lin, obj.quad, obj.ref)
    ev.set_prev_data(XP, UP)
    out = {k: v.cpu().numpy() for k, v in ev.eval(Z, X0, lam, torch.as_tensor(sig).cuda()).items()}
    assert ev.launch_count >= 4 and "nempc_rolling" in ev.kernel_name
    for b in range(B):
        roll = RollingMLP(net.weights, xd, ud, w, fwd)
        roll.set_prev_data(XP[b], UP[b])
        be = RollingBlockEvaluator(roll, kind, H, obj)
        if b == 0:
            np.testing.assert_array_equal(ev.hes_rows, be.hes_rows); np.testing.assert_array_equal(ev.hes_cols, be.hes_cols)
            np.testing.assert_array_equal(ev.jac_rows, be.jac_rows); np.testing.assert_array_equal(ev.jac_cols, be.jac_cols)
        ref = be.evaluate(Z[b], X0[b], lam[b][None], sig[b])
        for kr, kg in (("resid", "resid"), ("jac_vals", "jac"), ("hes_vals", "hes"), ("obj", "obj"), ("grad", "grad")):
            assert elem_err(out[kg][b], ref[kr][0]) < 1e-10, (b, kg)
    # reduced request sets, and determinism
    o1 = ev.eval(Z, X0, want=("resid", "jac"))
    assert set(o1) == {"resid", "jac"} and torch.equal(o1["jac"], ev.eval(Z, X0, want=("jac",))["jac"])
    again = ev.eval(Z, X0, lam, torch.as_tensor(sig).cuda())
    assert np.array_equal(again["hes"].cpu().numpy(), out["hes"])
    ev.close()


def test_reference_test_py_known_answer():
    """the reference's only runnable script (test.py:20-79): a rolling-window model (window 2, x_dim 2, u_dim 1), discrete integrator,
    H = 10, cost sum((u - 2)^2), no bounds -> the optimum is u == 2 with cost 0 whatever the dynamics (the states are free to follow).
    Here the dynamics are a tanh network; JAXObjectifFunc identifies the cost from the callable."""
    import warnings
    from pyneuralempc_b200.constraints import DomainConstraint
    from pyneuralempc_b200.controller import NMPC
    from pyneuralempc_b200.integrator.discret import DiscretIntegrator
    from pyneuralempc_b200.model.tensorflow import KerasTFModelRollingInput
    from pyneuralempc_b200.objective.jax import JAXObjectifFunc
    from pyneuralempc_b200.optimizer import Slsqp, TrustConstr
    H = 10
    net = MLP.glorot([6, 16, 16, 2], 2, 4, seed=1)
    W, b = net.weights[-1]; net.weights[-1] = (0.3 * W, 0.3 * b)
    model = KerasTFModelRollingInput(net.weights, 2, 1, forward_rolling=True, dtype="float64")
    model.set_prev_data(np.array([[0.2, 0.1]]), np.array([[0.0]]))                     # test.py:39-40, 49-50
    integ = DiscretIntegrator(model, H)
    cost = JAXObjectifFunc(lambda x, u, p=None, tvp=None: np.sum(np.square(u.reshape(-1) - 2.0)))   # test.py:59-60
    dom = DomainConstraint(states_constraint=[[-np.inf, np.inf], [-np.inf, np.inf]], control_constraint=[[-np.inf, np.inf]])
    for opt in (Slsqp(verbose=0), TrustConstr()):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pred, u = NMPC(integ, cost, [dom], H, 1, optimizer=opt, use_hessian=isinstance(opt, TrustConstr)).next(np.array([0.2, 0.1]))
        assert pred is not None and pred.shape == (H, 2) and u.shape == (H, 1)
        # SLSQP stops on its ftol = 0.5e-6 (the reference's default, optimizer/slsqp.py:117): u is then within ~1e-3 of the optimum
        assert np.abs(u - 2.0).max() < (2e-3 if isinstance(opt, Slsqp) else 1e-5)
        assert abs(opt.last_result.fun) < (1e-5 if isinstance(opt, Slsqp) else 1e-9)
        # the predicted states follow the window dynamics
        roll = RollingMLP(net.weights, 2, 1, 2, True)
        roll.set_prev_data(np.array([[0.2, 0.1]]), np.array([[0.0]]))
        xp = np.concatenate([np.array([[0.2, 0.1]]), pred])[:-1]
        assert np.abs(xp + roll.forward(xp, u) - pred).max() < 1e-6


def test_rolling_model_errors():
    from pyneuralempc_b200 import integrator as I
    from pyneuralempc_b200.model import CudaMLPModelRollingInput
    net = MLP.glorot([6, 8, 2], 2, 4, seed=0)
    with pytest.raises(ValueError):
        CudaMLPModelRollingInput(net.weights, 2, 1, rolling_window=3)                  # 9 inputs expected
    with pytest.raises(ValueError):
        CudaMLPModelRollingInput(net.weights, 2, 1, rolling_window=0)
    m = CudaMLPModelRollingInput(net.weights, 2, 1, rolling_window=2)
    with pytest.raises(AssertionError):
        m.forward(np.zeros((3, 2)), np.zeros((3, 1)))                                   # set_prev_data missing (tensorflow.py:189)
    with pytest.raises(AssertionError):
        m.set_prev_data(np.zeros((2, 2)), np.zeros((1, 1)))
    with pytest.raises(NotImplementedError):
        I.RK4Integrator(m, 5, 0.1)
