"""The oracle against what the unmodified reference computed (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference).  CPU only."""
import os

import numpy as np
import pytest

from oracle.blocks_np import BlockEvaluator
from oracle.dense_ref import DenseIntegrator, DenseIpoptProblem
from oracle.mlp_np import MLP, DenseModelView
from oracle.objectives_np import SeparableQuadraticObjective
from oracle import structure as S

KINDS = ("discrete", "unity", "rk4")


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _setup(g, weights, dtype=np.float64):
    mlp = MLP(weights, int(g["x_dim"]), int(g["u_dim"]), dtype=dtype)
    obj = SeparableQuadraticObjective(g["obj_lin"], g["obj_quad"], g["obj_ref"])
    kind, H = str(g["kind"]), int(g["H"])
    integ = DenseIntegrator(DenseModelView(mlp), H, kind, DT=float(g["DT"]))
    return mlp, obj, kind, H, integ


@pytest.mark.parametrize("kind", KINDS)
def test_dense_restatement_equals_reference_H6(golden_dir, lv_weights, kind):
    g = _load(golden_dir, f"ref_{kind}_H6.npz")
    mlp, obj, kind, H, integ = _setup(g, lv_weights)
    z, x0 = g["z"], g["x0"]
    s, u = z[:H * 2].reshape(H, 2), z[H * 2:].reshape(H, 1)
    np.testing.assert_allclose(integ.forward(s, u, x0), g["integrator_forward"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(integ.jacobian(s, u, x0), g["integrator_jacobian"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(integ.hessian(s, u, x0), g["integrator_hessian"], rtol=0, atol=1e-12)
    # numerically probed structure of the reference == probed structure of the restatement == analytic map
    np.testing.assert_array_equal(integ.hessianstructure(), g["integrator_structure"])
    np.testing.assert_array_equal(S.integrator_hessian_map(H, 2, 1), g["integrator_structure"])


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("H", (6, 25))
def test_ipopt_callbacks_equal_reference(golden_dir, lv_weights, kind, H):
    g = _load(golden_dir, f"ref_{kind}_H{H}.npz")
    mlp, obj, kind, H, integ = _setup(g, lv_weights)
    pb = DenseIpoptProblem(g["x0"], obj, integ)
    z = g["z"]
    assert abs(pb.objective(z) - float(g["objective"])) < 1e-12
    np.testing.assert_allclose(pb.gradient(z), g["gradient"], atol=1e-13)
    np.testing.assert_allclose(pb.constraints(z), g["constraints"], atol=1e-14)
    np.testing.assert_allclose(pb.jacobian(z), g["jacobian"], atol=1e-13)
    r, c = pb.hessianstructure()
    np.testing.assert_array_equal(r, g["hes_rows"])
    np.testing.assert_array_equal(c, g["hes_cols"])
    np.testing.assert_allclose(pb.hessian(z, g["lam"], float(g["sigma"])), g["hessian_values"], atol=1e-12)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("H", (6, 25))
def test_block_oracle_equals_reference(golden_dir, lv_weights, kind, H):
    """sparse per-step formulation: identical indices (bit-exact) and values."""
    g = _load(golden_dir, f"ref_{kind}_H{H}.npz")
    mlp, obj, kind, H, _ = _setup(g, lv_weights)
    ev = BlockEvaluator(mlp, kind, H, DT=float(g["DT"]), objective=obj)
    np.testing.assert_array_equal(ev.hes_rows, g["hes_rows"])
    np.testing.assert_array_equal(ev.hes_cols, g["hes_cols"])
    jr, jc = np.nonzero(g["jacobian"])
    np.testing.assert_array_equal(ev.jac_rows, jr)
    np.testing.assert_array_equal(ev.jac_cols, jc)
    out = ev.evaluate(g["z"][None], g["x0"][None], g["lam"][None], float(g["sigma"]))
    np.testing.assert_allclose(out["resid"][0], g["constraints"], atol=1e-14)
    np.testing.assert_allclose(out["jac_vals"][0], g["jacobian"][jr, jc], atol=1e-13)
    np.testing.assert_allclose(out["hes_vals"][0], g["hessian_values"], atol=1e-12)
    np.testing.assert_allclose(out["grad"][0], g["gradient"], atol=1e-13)
    assert abs(out["obj"][0] - float(g["objective"])) < 1e-12


def test_reference_shipped_problem_sizes(golden_dir):
    """C1 of BASELINE.json: n=75, m=50, tril nnz of the integrator part 145."""
    g = _load(golden_dir, "ref_rk4_H25.npz")          # linear objective -> integrator pattern only
    assert g["jacobian"].shape == (50, 75)
    assert len(g["hes_rows"]) == 145 == S.nnz_hessian_integrator(25, 2, 1)
    assert np.count_nonzero(g["jacobian"]) == 196 == S.nnz_jacobian(25, 2, 1)


def test_float32_network_mimic(golden_dir, lv_weights):
    g = _load(golden_dir, "ref_rk4_f32_H6.npz")
    mlp, obj, kind, H, integ = _setup(g, lv_weights, dtype=np.float32)
    ev = BlockEvaluator(mlp, kind, H, DT=float(g["DT"]), objective=obj)
    out = ev.evaluate(g["z"][None], g["x0"][None], g["lam"][None], float(g["sigma"]))
    # same float32 network arithmetic, different summation order: ~1e-6 relative
    np.testing.assert_allclose(out["resid"][0], g["constraints"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["hes_vals"][0], g["hessian_values"], rtol=1e-4, atol=1e-5)
    # ... and the float64 oracle agrees with the float32 reference run to float32 accuracy
    ev64 = BlockEvaluator(mlp.astype(np.float64), kind, H, DT=float(g["DT"]), objective=obj)
    o64 = ev64.evaluate(g["z"][None], g["x0"][None], g["lam"][None], float(g["sigma"]))
    scale = np.abs(g["hessian_values"]).max()
    assert np.abs(o64["hes_vals"][0] - g["hessian_values"]).max() < 1e-5 * scale


def test_d5_network_discrete_unity_and_rk4_first_order(golden_dir):
    g = _load(golden_dir, "ref_discrete_d5_H5.npz")
    weights = [(g[f"net_W{i}"], g[f"net_b{i}"]) for i in range(3)]
    mlp = MLP(weights, 4, 1)
    obj = SeparableQuadraticObjective(g["obj_lin"], g["obj_quad"], g["obj_ref"])
    H = 5
    z, x0 = g["z"], g["x0"]
    s, u = z[:H * 4].reshape(H, 4), z[H * 4:].reshape(H, 1)
    for kind, pre in (("discrete", ""), ("unity", "unity_")):
        integ = DenseIntegrator(DenseModelView(mlp), H, kind)
        np.testing.assert_allclose(integ.forward(s, u, x0), g[pre + "integrator_forward"], atol=1e-14)
        np.testing.assert_allclose(integ.jacobian(s, u, x0), g[pre + "integrator_jacobian"], atol=1e-13)
        np.testing.assert_allclose(integ.hessian(s, u, x0), g[pre + "integrator_hessian"], atol=1e-12)
        ev = BlockEvaluator(mlp, kind, H, objective=obj)
        out = ev.evaluate(z[None], x0[None], g["lam"][None], float(g["sigma"]))
        np.testing.assert_allclose(out["hes_vals"][0], g[pre + "hessian_values"], atol=1e-12)
    # the reference's RK4 Hessian only works for x_dim+u_dim == 3 (rk4.py:246,255,261)
    assert "none" != str(g["rk4_d5_error"])
    integ = DenseIntegrator(DenseModelView(mlp), H, "rk4", DT=float(g["DT"]))
    np.testing.assert_allclose(integ.forward(g["rk4_x"], g["rk4_u"], g["rk4_x0"]), g["rk4_forward"], atol=1e-14)
    np.testing.assert_allclose(integ.jacobian(g["rk4_x"], g["rk4_u"], g["rk4_x0"]), g["rk4_jacobian"], atol=1e-13)


@pytest.mark.parametrize("kind", KINDS)
def test_block_oracle_with_tvp_and_p_equals_reference(golden_dir, kind):
    """time-varying (tvp) and constant (p) model inputs through the reference's own integrators (called with ``p=, tvp=``)
    against the per-step oracle bound to the same exogenous rows."""
    from oracle.mlp_np import ExoMLP
    g = _load(golden_dir, "ref_exo_H6.npz")
    H, xd, ud = int(g["H"]), int(g["x_dim"]), int(g["u_dim"])
    weights = [(g[f"net_W{i}"], g[f"net_b{i}"]) for i in range(3)]
    exo = ExoMLP(weights, xd, ud, int(g["tvp_dim"]), int(g["p_dim"])).bind(g["tvp"], g["p"], B=1)
    ev = BlockEvaluator(exo, kind, H, DT=float(g["DT"]))
    z = np.concatenate([g["states"].ravel(), g["u"].ravel()])[None, :]
    lam = g["lam"]
    out = ev.evaluate(z, g["x0"][None, :], lam=lam[None, :])
    np.testing.assert_allclose(out["resid"][0], g[f"{kind}_forward"], rtol=0, atol=1e-14)
    dense_j = g[f"{kind}_jacobian"]
    np.testing.assert_allclose(out["jac_vals"][0], dense_j[ev.jac_rows, ev.jac_cols], rtol=0, atol=1e-13)
    mask = np.zeros_like(dense_j, bool)
    mask[ev.jac_rows, ev.jac_cols] = True
    assert np.all(dense_j[~mask] == 0.0)
    dense_h = np.einsum("i,iab->ab", lam, g[f"{kind}_hessian"])
    np.testing.assert_allclose(out["hes_vals"][0], dense_h[ev.hes_rows, ev.hes_cols], rtol=0, atol=1e-12)
    hm = np.zeros_like(dense_h, bool)
    hm[ev.hes_rows, ev.hes_cols] = True
    assert np.all(np.tril(dense_h)[~hm] == 0.0)
