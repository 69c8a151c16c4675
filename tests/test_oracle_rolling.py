"""Rolling-window (NARX) models, CPU side: the oracle restatement (oracle/rolling_np.py) against finite differences and against the goldens
recorded from the unmodified reference's integrators / IpoptProblem; the product's closed-form band structure (pyneuralempc_b200/rolling.py,
pure numpy) against the reference's numerically probed one -- bit-identical indices."""
import os

import numpy as np
import pytest

from oracle.mlp_np import MLP
from oracle.objectives_np import SeparableQuadraticObjective
from oracle.rolling_np import RollingBlockEvaluator, RollingMLP

GOLDENS = ("ref_rolling_discrete_w2.npz", "ref_rolling_unity_w3.npz")


def _load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    ws = [(g[f"net_W{i}"], g[f"net_b{i}"]) for i in range(3)]
    roll = RollingMLP(ws, int(g["x_dim"]), int(g["u_dim"]), int(g["rolling_window"]), bool(g["forward_rolling"]))
    roll.set_prev_data(g["x_prev"], g["u_prev"])
    return g, ws, roll


@pytest.mark.parametrize("w,fwd", [(1, True), (2, True), (3, False), (4, True)])
def test_rolling_model_derivatives_vs_finite_differences(w, fwd):
    xd, ud, N = 2, 1, 5
    rng = np.random.default_rng(w)
    net = MLP.glorot([w * (xd + ud), 7, 6, xd], xd, w * (xd + ud) - xd, seed=w)
    roll = RollingMLP(net.weights, xd, ud, w, fwd)
    roll.set_prev_data(rng.uniform(-1, 1, (w - 1, xd)), rng.uniform(-1, 1, (w - 1, ud)))
    x, u = rng.uniform(-1, 1, (N, xd)), rng.uniform(-1, 1, (N, ud))
    flat = np.concatenate([x.ravel(), u.ravel()])
    f = lambda v: roll.forward(v[:N * xd].reshape(N, xd), v[N * xd:].reshape(N, ud)).ravel()
    g = lambda v: roll.jacobian(v[:N * xd].reshape(N, xd), v[N * xd:].reshape(N, ud))
    J, Hs = roll.jacobian(x, u), roll.hessian(x, u).reshape(N * xd, N * (xd + ud), N * (xd + ud))
    eps = 1e-6
    for i in range(len(flat)):
        e = np.zeros_like(flat); e[i] = eps
        np.testing.assert_allclose((f(flat + e) - f(flat - e)) / (2 * eps), J[:, i], atol=1e-8)
        np.testing.assert_allclose((g(flat + e) - g(flat - e)) / (2 * eps), Hs[:, :, i], atol=1e-7)
    # the band: row block i only touches rows i-w+1..i
    for i in range(N):
        for j in range(N):
            blk = J[i * xd:(i + 1) * xd, j * xd:(j + 1) * xd]
            assert (np.abs(blk).max() > 0) == (i - w + 1 <= j <= i)


@pytest.mark.parametrize("name", GOLDENS)
def test_oracle_rolling_assembly_equals_the_reference(golden_dir, name):
    g, ws, roll = _load(golden_dir, name)
    H, kind = int(g["H"]), str(g["kind"])
    obj = SeparableQuadraticObjective(g["obj_lin"], g["obj_quad"], g["obj_ref"])
    be = RollingBlockEvaluator(roll, kind, H, obj)
    np.testing.assert_array_equal(be.hes_rows, g["hes_rows"]); np.testing.assert_array_equal(be.hes_cols, g["hes_cols"])
    jr, jc = np.nonzero(g["jacobian"])
    np.testing.assert_array_equal(be.jac_rows, jr); np.testing.assert_array_equal(be.jac_cols, jc)
    o = be.evaluate(g["z"], g["x0"], g["lam"][None], float(g["sigma"]))
    np.testing.assert_allclose(o["resid"][0], g["constraints"], atol=1e-13)
    np.testing.assert_allclose(o["jac_vals"][0], g["jacobian"][jr, jc], atol=1e-13)
    np.testing.assert_allclose(o["hes_vals"][0], g["hessian_values"], atol=1e-12)
    np.testing.assert_allclose(o["grad"][0], g["gradient"], atol=1e-13)


@pytest.mark.parametrize("name", GOLDENS)
def test_product_band_structure_is_the_reference_structure(golden_dir, name):
    from pyneuralempc_b200.rolling import rolling_structure, window_columns
    g, ws, roll = _load(golden_dir, name)
    H, w, kind, fwd = int(g["H"]), int(g["rolling_window"]), str(g["kind"]), bool(g["forward_rolling"])
    st = rolling_structure(H, 2, 1, w, kind, g["obj_quad"] != 0, fwd)
    np.testing.assert_array_equal(st["hes_rows"], g["hes_rows"]); np.testing.assert_array_equal(st["hes_cols"], g["hes_cols"])
    jr, jc = np.nonzero(g["jacobian"])
    np.testing.assert_array_equal(st["jac_rows"], jr); np.testing.assert_array_equal(st["jac_cols"], jc)
    # the gather codes reproduce the oracle's window rows
    z, x0 = g["z"], g["x0"]
    aux = np.concatenate([x0, g["x_prev"].ravel(), g["u_prev"].ravel()])
    code = window_columns(H, 2, 1, w, fwd)
    zin = np.where(code >= 0, z[np.clip(code, 0, None)], aux[np.clip(-1 - code, 0, None)])
    xp = np.concatenate([x0[None], z[:2 * H].reshape(H, 2)])[:-1]
    np.testing.assert_array_equal(zin, roll._windows(xp, z[2 * H:].reshape(H, 1)))
    # window 1 degenerates to the plain one-step structure
    from pyneuralempc_b200.structure import nlp_structure
    s1 = rolling_structure(5, 2, 1, 1, "discrete", None)
    jr1, jc1, hr1, hc1 = nlp_structure(5, 2, 1, None)
    np.testing.assert_array_equal(s1["jac_rows"], jr1); np.testing.assert_array_equal(s1["jac_cols"], jc1)
    np.testing.assert_array_equal(s1["hes_rows"], hr1); np.testing.assert_array_equal(s1["hes_cols"], hc1)
