"""tvp / p model inputs on the CUDA path: C-ABI parity with the oracle, the reference's own integrators (golden
file recorded from /root/reference called with ``p=, tvp=``), the chunked host call, the drop-in classes and the on-device solver."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from parity_metric import elem_err  # noqa: E402

from oracle.blocks_np import BlockEvaluator  # noqa: E402
from oracle.mlp_np import MLP, ExoMLP  # noqa: E402
from oracle.objectives_np import SeparableQuadraticObjective  # noqa: E402

TOL32, TOL64 = 1e-5, 1e-10


def _relerr(got, ref):
    return elem_err(got, ref)    # elementwise: |d| <= tol |ref| + 0.1 tol max|ref| (tests/parity_metric.py)


def _exo(xd, ud, td, pd, hidden, seed):
    full = MLP.glorot([xd + ud + td + pd] + hidden + [xd], xd, ud + td + pd, seed=seed)
    return ExoMLP(full.weights, xd, ud, td, pd)


def _evaluator(exo, kind, H, compute, obj=None, kernel="auto"):
    from pyneuralempc_b200 import NlpEvaluator
    ev = NlpEvaluator(exo.weights, exo.x_dim, exo.u_dim, H, kind, DT=0.1, compute_dtype=compute, kernel=kernel,
                      tvp_dim=exo.tvp_dim, p_dim=exo.p_dim)
    if obj is not None:
        ev.set_objective(obj.lin, obj.quad, obj.ref)
    return ev


CASES = [("discrete", 2, 1, 2, 1, [30, 30], 25), ("rk4", 2, 1, 2, 0, [30, 30], 10), ("unity", 3, 2, 0, 2, [12, 9], 4),
         ("rk4", 4, 1, 1, 1, [128, 128, 128], 5), ("discrete", 12, 4, 3, 2, [64, 64], 3), ("discrete", 2, 1, 2, 1, [64, 64], 30),
         ("rk4", 4, 1, 2, 0, [32, 32, 32], 9)]
TC_SHAPES = {(4, 1, 128, 3), (2, 1, 64, 2), (4, 1, 32, 3)}        # (x, u, width, hidden layers) of the cases the tensor-core kernel serves


@pytest.mark.parametrize("kind,xd,ud,td,pd,hidden,H", CASES)
@pytest.mark.parametrize("compute", ("float64", "float32"))
@pytest.mark.parametrize("shared", (False, True))
def test_eval_with_tvp_and_p_vs_oracle(kind, xd, ud, td, pd, hidden, H, compute, shared):
    B = 7
    rng = np.random.default_rng(31)
    exo = _exo(xd, ud, td, pd, hidden, 5)
    tvp = None if td == 0 else rng.uniform(-1, 1, (H, td) if shared else (B, H, td))
    p = None if pd == 0 else rng.uniform(-1, 1, (pd,) if shared else (B, pd))
    n, m = H * (xd + ud), H * xd
    obj = SeparableQuadraticObjective.tracking(H, xd, ud, rng.uniform(0.5, 2, xd), rng.uniform(0.1, 1, ud))
    Z, X0, lam, sig = rng.uniform(-1, 1, (B, n)), rng.uniform(-1, 1, (B, xd)), rng.standard_normal((B, m)), rng.uniform(0.5, 1.5, B)
    ref = BlockEvaluator(exo.bind(tvp, p, B=B, H=H), kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    ev = _evaluator(exo, kind, H, compute, obj)
    on_tc = compute == "float32" and (xd, ud, hidden[0], len(hidden)) in TC_SHAPES
    assert ("tcgen05" if on_tc else "generic") in ev.kernel_name
    ev.set_exogenous(tvp, p)
    got = ev.eval_host(Z, X0, lam, sig)
    tol = TOL64 if compute == "float64" else TOL32
    for kr, kg in (("resid", "resid"), ("jac_vals", "jac"), ("hes_vals", "hes"), ("obj", "obj"), ("grad", "grad")):
        assert _relerr(got[kg], ref[kr]) < tol, (kg, _relerr(got[kg], ref[kr]))
    # new exogenous rows, same iterate: the outputs must follow (the replayed graph of the host call is dropped)
    if not shared:
        tvp2 = None if tvp is None else tvp[::-1].copy()
        p2 = None if p is None else p[::-1].copy()
        ev.set_exogenous(tvp2, p2)
        ref2 = BlockEvaluator(exo.bind(tvp2, p2, B=B, H=H), kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
        got2 = ev.eval_host(Z, X0, lam, sig)
        assert _relerr(got2["hes"], ref2["hes_vals"]) < tol and _relerr(got2["resid"], ref2["resid"]) < tol
    ev.close()


def test_chunked_host_call_offsets_the_exogenous_rows():
    """a batch large enough for the chunk pipeline of nempc_eval_host: chunk c must read the tvp / p rows of ITS problems"""
    H, B = 10, 4096
    rng = np.random.default_rng(2)
    exo = _exo(2, 1, 2, 1, [30, 30], 8)
    tvp, p = rng.uniform(-1, 1, (B, H, 2)), rng.uniform(-1, 1, (B, 1))
    Z, X0, lam = rng.uniform(-1, 1, (B, H * 3)), rng.uniform(-1, 1, (B, 2)), rng.standard_normal((B, H * 2))
    ref = BlockEvaluator(exo.bind(tvp, p, B=B, H=H), "rk4", H, DT=0.1).evaluate(Z, X0, lam)
    ev = _evaluator(exo, "rk4", H, "float64")
    ev.set_exogenous(tvp, p)
    for _ in range(3):                       # third call replays the captured graph
        got = ev.eval_host(Z, X0, lam, want=("resid", "jac", "hes"))
        assert _relerr(got["hes"], ref["hes_vals"]) < TOL64 and _relerr(got["jac"], ref["jac_vals"]) < TOL64
    ev.close()


def test_exogenous_rows_given_as_cuda_tensors_and_batched_solver():
    """tvp / p as device tensors (copied device-to-device), then the batched on-device solver with per-problem exogenous rows: every
    problem's solution satisfies the oracle's constraints for ITS rows"""
    import torch
    H, B = 6, 5
    rng = np.random.default_rng(8)
    exo = _exo(2, 1, 1, 1, [16, 16], 3)
    exo.weights[0] = (exo.weights[0][0] * 0.5, exo.weights[0][1])
    tvp, p = rng.uniform(-1, 1, (B, H, 1)), rng.uniform(-1, 1, (B, 1))
    Z, X0, lam = rng.uniform(-1, 1, (B, H * 3)), rng.uniform(-0.5, 0.5, (B, 2)), rng.standard_normal((B, H * 2))
    ev = _evaluator(exo, "discrete", H, "float64")
    ev.set_exogenous(torch.as_tensor(tvp).cuda(), torch.as_tensor(p, dtype=torch.float32).cuda())     # float32 tensor: converted
    p32 = p.astype(np.float32).astype(np.float64)
    ref = BlockEvaluator(exo.bind(tvp, p32, B=B, H=H), "discrete", H).evaluate(Z, X0, lam)
    got = ev.eval_host(Z, X0, lam, want=("resid", "jac", "hes"))
    assert _relerr(got["hes"], ref["hes_vals"]) < TOL64 and _relerr(got["resid"], ref["resid"]) < TOL64
    n = H * 3
    quad = np.concatenate([np.ones(H * 2), 0.1 * np.ones(H)])
    ev.set_objective(np.zeros(n), quad, np.zeros(n))
    lb = np.concatenate([np.full(H * 2, -5.0), np.full(H, -1.0)])
    out = ev.solve(X0, lb, -lb, max_iter=80, tol=1e-7)
    assert bool((out["status"] == 0).all())
    zs = out["z"].cpu().numpy()
    res = BlockEvaluator(exo.bind(tvp, p32, B=B, H=H), "discrete", H).evaluate(zs, X0, need_jac=False, need_hes=False)["resid"]
    assert np.abs(res).max() < 1e-6
    ev.close()


def test_errors():
    from pyneuralempc_b200 import NlpEvaluator
    from pyneuralempc_b200._lib import NempcError
    exo = _exo(2, 1, 2, 1, [30, 30], 8)
    with pytest.raises(NempcError):          # the register-resident kernel has no exogenous inputs
        _evaluator(exo, "rk4", 5, "float32", kernel="fast")
    ev = _evaluator(exo, "rk4", 5, "float32")
    Z, X0 = np.zeros((3, 15)), np.zeros((3, 2))
    with pytest.raises(NempcError):          # evaluated before nempc_set_exogenous
        ev.eval_host(Z, X0, want=("resid",))
    ev.set_exogenous(np.zeros((2, 5, 2)), np.zeros((3, 1)))
    with pytest.raises(NempcError):          # rows for two problems, batch of three
        ev.eval_host(Z, X0, want=("resid",))
    with pytest.raises(ValueError):
        ev.set_exogenous(np.zeros((5, 3)), np.zeros(1))
    ev.close()
    plain = NlpEvaluator(MLP.glorot([3, 8, 2], 2, 1).weights, 2, 1, 4, "discrete")
    with pytest.raises(ValueError):
        plain.set_exogenous(np.zeros((4, 1)), None)
    plain.close()


@pytest.mark.parametrize("kind", ("discrete", "unity", "rk4"))
def test_dropin_integrators_with_tvp_and_p_vs_reference_golden(golden_dir, kind):
    from pyneuralempc_b200 import integrator as I
    from pyneuralempc_b200.model import CudaMLPModel
    g = np.load(os.path.join(golden_dir, "ref_exo_H6.npz"))
    H = int(g["H"])
    weights = [(g[f"net_W{i}"], g[f"net_b{i}"]) for i in range(3)]
    model = CudaMLPModel(weights, 2, 1, p_dim=1, tvp_dim=2, dtype="float64")
    integ = {"discrete": lambda: I.DiscretIntegrator(model, H), "unity": lambda: I.UnityIntegrator(model, H),
             "rk4": lambda: I.RK4Integrator(model, H, float(g["DT"]))}[kind]()
    s, u, x0, p, tvp = g["states"], g["u"], g["x0"], g["p"], g["tvp"]
    assert _relerr(integ.forward(s, u, x0, p=p, tvp=tvp), g[f"{kind}_forward"]) < TOL64
    assert _relerr(integ.jacobian(s, u, x0, p=p, tvp=tvp), g[f"{kind}_jacobian"]) < TOL64
    assert _relerr(integ.hessian(s, u, x0, p=p, tvp=tvp), g[f"{kind}_hessian"]) < TOL64
    # the model interface itself (KerasTFModel layouts with the tvp / p columns sliced away, model/tensorflow.py:65-66, 97-98)
    exo = ExoMLP(weights, 2, 1, 2, 1).bind(tvp, p)
    xp = np.concatenate([x0[None], s[:-1]])
    assert _relerr(model.forward(xp, u, p=p, tvp=tvp), exo.forward(xp, u)) < TOL64
    assert _relerr(model.jacobian(xp, u, p=p, tvp=tvp), exo.dense_jacobian(xp, u)) < TOL64
    assert _relerr(model.hessian(xp, u, p=p, tvp=tvp), exo.dense_hessian(xp, u)) < TOL64
    with pytest.raises(ValueError):
        model.forward(xp, u)


def test_nmpc_with_tvp_solves_and_tracks_the_exogenous_signal():
    """closed loop through NMPC.next(x0, p=, tvp=): the optimum found with the CUDA callbacks satisfies the oracle's constraints
    for THOSE exogenous rows, and a different tvp gives a different plan."""
    from pyneuralempc_b200 import integrator as I
    from pyneuralempc_b200.constraints import DomainConstraint
    from pyneuralempc_b200.controller import NMPC
    from pyneuralempc_b200.model import CudaMLPModel
    from pyneuralempc_b200.objective import CudaQuadraticObjective
    from pyneuralempc_b200.optimizer import CudaIpm
    H = 8
    rng = np.random.default_rng(4)
    exo = _exo(2, 1, 1, 1, [16, 16], 3)
    exo.weights[0] = (exo.weights[0][0] * 0.5, exo.weights[0][1])
    model = CudaMLPModel(exo.weights, 2, 1, p_dim=1, tvp_dim=1, dtype="float64")
    integ = I.DiscretIntegrator(model, H)
    obj = CudaQuadraticObjective(H, 2, 1, [1.0, 1.0], [0.1], x_ref=np.zeros((H, 2)))
    dom = DomainConstraint(states_constraint=[[-5.0, 5.0]] * 2, control_constraint=[[-1.0, 1.0]])
    mpc = NMPC(integ, obj, [dom], H, 1.0, optimizer=CudaIpm(max_iteration=80, tolerance=1e-7))
    x0 = np.array([0.4, -0.3])
    plans = []
    for tvp in (np.full((H, 1), 0.8), np.full((H, 1), -0.8)):
        p = np.array([0.2])
        xs, us = mpc.next(x0, p=p, tvp=tvp)
        assert xs is not None
        z = np.concatenate([xs.ravel(), us.ravel()])[None]
        res = BlockEvaluator(exo.bind(tvp, p, B=1, H=H), "discrete", H).evaluate(z, x0[None], need_jac=False, need_hes=False)["resid"]
        assert np.abs(res).max() < 1e-6
        plans.append(us.copy())
    assert np.abs(plans[0] - plans[1]).max() > 1e-3
