"""Internal consistency of the oracle: the O(H) block formulation equals the dense reference-literal
restatement for every integrator and for dimensions the reference's RK4 cannot handle (d != 3), and
both agree with central finite differences of the residual.  CPU only."""
import numpy as np
import pytest

from oracle.blocks_np import BlockEvaluator
from oracle.dense_ref import DenseIntegrator, DenseIpoptProblem
from oracle.mlp_np import MLP, DenseModelView
from oracle.objectives_np import SeparableQuadraticObjective

CASES = [("discrete", [3, 10, 10, 2], 2, 1, 7), ("unity", [3, 10, 10, 2], 2, 1, 7), ("rk4", [3, 10, 10, 2], 2, 1, 7),
         ("rk4", [5, 12, 9, 4], 4, 1, 5), ("rk4", [6, 8, 3], 3, 3, 4), ("discrete", [16, 20, 12], 12, 4, 3),
         ("rk4", [2, 6, 1], 1, 1, 1), ("unity", [4, 6, 6, 6, 2], 2, 2, 2)]


@pytest.mark.parametrize("kind,dims,xd,ud,H", CASES)
def test_block_vs_dense(kind, dims, xd, ud, H):
    rng = np.random.default_rng(11)
    mlp = MLP.glorot(dims, xd, ud, seed=5)
    n, m = H * (xd + ud), H * xd
    obj = SeparableQuadraticObjective.tracking(H, xd, ud, rng.uniform(0.5, 2, xd), rng.uniform(0.1, 1, ud),
                                               x_ref=rng.uniform(-1, 1, (H, xd)))
    obj.lin[:] = rng.uniform(-1, 1, n)
    B = 3
    Z, X0, lam = rng.uniform(-1, 1, (B, n)), rng.uniform(-1, 1, (B, xd)), rng.standard_normal((B, m))
    sig = rng.uniform(0.5, 1.5, B)
    ev = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj)
    out = ev.evaluate(Z, X0, lam, sig)
    for b in range(B):
        integ = DenseIntegrator(DenseModelView(mlp), H, kind, DT=0.1)
        pb = DenseIpoptProblem(X0[b], obj, integ)
        np.testing.assert_allclose(out["resid"][b], pb.constraints(Z[b]), atol=1e-13)
        Jd = pb.jacobian(Z[b])
        np.testing.assert_allclose(out["jac_vals"][b], Jd[ev.jac_rows, ev.jac_cols], atol=1e-12)
        assert np.count_nonzero(Jd) == len(ev.jac_rows)
        r, c = pb.hessianstructure()
        np.testing.assert_array_equal(r, ev.hes_rows)
        np.testing.assert_array_equal(c, ev.hes_cols)
        np.testing.assert_allclose(out["hes_vals"][b], pb.hessian(Z[b], lam[b], sig[b]), atol=1e-11)
        np.testing.assert_allclose(out["grad"][b], pb.gradient(Z[b]), atol=1e-13)
        assert abs(out["obj"][b] - pb.objective(Z[b])) < 1e-12


@pytest.mark.parametrize("kind,dims,xd,ud,H", [c for c in CASES if c[0] == "rk4"])
def test_generalised_rk4_against_finite_differences(kind, dims, xd, ud, H):
    """the eye(d) generalisation of rk4.py:246,255,261 is checked independently of the restatement."""
    rng = np.random.default_rng(3)
    mlp = MLP.glorot(dims, xd, ud, seed=9)
    n, m = H * (xd + ud), H * xd
    ev = BlockEvaluator(mlp, kind, H, DT=0.1)
    Z, X0, lam = rng.uniform(-1, 1, (1, n)), rng.uniform(-1, 1, (1, xd)), rng.standard_normal((1, m))
    out = ev.evaluate(Z, X0, lam)
    J = np.zeros((m, n)); J[ev.jac_rows, ev.jac_cols] = out["jac_vals"][0]
    Hl = np.zeros((n, n)); Hl[ev.hes_rows, ev.hes_cols] = out["hes_vals"][0]
    Hl = Hl + np.tril(Hl, -1).T
    eps = 1e-5
    Jfd, Hfd = np.zeros((m, n)), np.zeros((n, n))
    for i in range(n):
        dz = np.zeros((1, n)); dz[0, i] = eps
        op, om = ev.evaluate(Z + dz, X0, lam), ev.evaluate(Z - dz, X0, lam)
        Jfd[:, i] = (op["resid"][0] - om["resid"][0]) / (2 * eps)
        Jp = np.zeros((m, n)); Jp[ev.jac_rows, ev.jac_cols] = op["jac_vals"][0]
        Jm = np.zeros((m, n)); Jm[ev.jac_rows, ev.jac_cols] = om["jac_vals"][0]
        Hfd[:, i] = lam[0] @ (Jp - Jm) / (2 * eps)
    np.testing.assert_allclose(J, Jfd, atol=1e-9)
    np.testing.assert_allclose(Hl, Hfd, atol=1e-8)


def test_reference_tril_nnz_formula():
    ev = BlockEvaluator(MLP.glorot([3, 4, 2], 2, 1), "rk4", 25, DT=0.1)
    assert len(ev.hes_rows) == 145 and len(ev.jac_rows) == 196


def test_quadratic_form_objective_oracle_and_product_matrices_agree():
    """non-separable quadratic costs: the oracle's literal cost definition, its separately assembled matrices and the product's
    CudaQuadraticFormObjective.from_blocks (host-side scipy assembly) describe the same function"""
    from oracle.objectives_np import QuadraticFormObjective
    from pyneuralempc_b200.objective import CudaQuadraticFormObjective
    rng = np.random.default_rng(0)
    H, xd, ud = 5, 3, 2
    A = lambda r, c: rng.standard_normal((r, c))
    blocks = dict(Q=A(xd, xd), R=A(ud, ud), Qf=A(xd, xd), S=A(ud, ud), N=A(xd, ud), x_ref=rng.standard_normal((H, xd)),
                  u_ref=rng.standard_normal(ud), lin=rng.standard_normal(H * (xd + ud)))
    o = QuadraticFormObjective(H, xd, ud, **blocks)
    s, u = rng.standard_normal((H, xd)), rng.standard_normal((H, ud))
    g, Hd = o._numeric(s, u)
    P, q, c = o.matrices()
    z = np.concatenate([s.ravel(), u.ravel()])
    assert abs(o.forward(s, u) - (0.5 * z @ P @ z + q @ z + c)) < 1e-11
    np.testing.assert_allclose(g, o.gradient(s, u), atol=1e-11)
    np.testing.assert_allclose(Hd, P, atol=1e-10)
    prod = CudaQuadraticFormObjective.from_blocks(H, xd, ud, **blocks)
    np.testing.assert_allclose(prod.P.toarray(), P, atol=1e-13)
    np.testing.assert_allclose(prod.q, q, atol=1e-12)
    assert abs(prod.c - c) < 1e-11
    # union pattern = np.nonzero(np.tril(objective_map + integrator_map)) (ipopt.py:55-62)
    from oracle import structure as S
    hr, hc = S.hessian_structure(H, xd, ud, None)
    r, cc, src, pval = prod.merge_tables(hr, hc)
    m = np.zeros((len(z), len(z)))
    m[hr, hc] = 1.0
    m += np.tril(P != 0)
    rr, rc = np.nonzero(m)
    np.testing.assert_array_equal(r, rr); np.testing.assert_array_equal(cc, rc)
    assert (src >= 0).sum() == len(hr) and np.array_equal(np.sort(src[src >= 0]), np.arange(len(hr)))
    np.testing.assert_allclose(pval, P[r, cc], atol=1e-13)
    with pytest.raises(ValueError):
        CudaQuadraticFormObjective(np.triu(P) + 1.0)                                   # not symmetric
