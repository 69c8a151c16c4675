"""Parity of the CUDA path (through the C ABI, libnempc.so) with the CPU oracle and the golden files recorded
from the unmodified reference.  Tolerances (ELEMENTWISE, tests/parity_metric.py): network arithmetic float32 -> every element within
1e-5 |ref| + 1e-6 max|ref|; float64 -> 1e-10 |ref| + 1e-11 max|ref| (BASELINE.json north_star).  Sparsity indices are compared bit-exactly."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from parity_metric import elem_err  # noqa: E402

from oracle.blocks_np import BlockEvaluator, step_blocks  # noqa: E402
from oracle.mlp_np import MLP  # noqa: E402
from oracle.objectives_np import SeparableQuadraticObjective  # noqa: E402

TOL32, TOL64 = 1e-5, 1e-10
KEYS = (("resid", "resid"), ("jac_vals", "jac"), ("hes_vals", "hes"), ("obj", "obj"), ("grad", "grad"))


def _relerr(got, ref):
    return elem_err(got, ref)    # elementwise: |d| <= tol |ref| + 0.1 tol max|ref| (tests/parity_metric.py)


def _evaluator(mlp, kind, H, compute, kernel="auto", obj=None, io="float64"):
    from pyneuralempc_b200 import NlpEvaluator
    ev = NlpEvaluator(mlp.weights, mlp.x_dim, mlp.u_dim, H, kind, DT=0.1, activation=mlp.activation,
                      compute_dtype=compute, io_dtype=io, kernel=kernel)
    if obj is not None:
        ev.set_objective(obj.lin, obj.quad, obj.ref)
    return ev


def _problem(dims, xd, ud, H, B, act="tanh", seed=0):
    rng = np.random.default_rng(seed)
    mlp = MLP.glorot(dims, xd, ud, seed=seed + 1, activation=act)
    n, m = H * (xd + ud), H * xd
    obj = SeparableQuadraticObjective.tracking(H, xd, ud, rng.uniform(0.5, 2, xd), rng.uniform(0.1, 1, ud),
                                               x_ref=rng.uniform(-1, 1, (H, xd)))
    obj.lin[:] = rng.uniform(-1, 1, n)
    return mlp, obj, rng.uniform(-1, 1, (B, n)), rng.uniform(-1, 1, (B, xd)), rng.standard_normal((B, m)), rng.uniform(0.5, 1.5, B)


def _run(ev, Z, X0, lam, sig):
    import torch
    t = lambda a: torch.as_tensor(a, dtype=ev.tdtype).cuda()
    out = ev.eval(t(Z), t(X0), t(lam), t(sig))
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


CASES = [("discrete", [3, 30, 30, 2], 2, 1, 25, "tanh"), ("unity", [3, 30, 30, 2], 2, 1, 25, "tanh"),
         ("rk4", [3, 30, 30, 2], 2, 1, 50, "tanh"), ("rk4", [5, 12, 9, 7, 4], 4, 1, 10, "tanh"),
         ("rk4", [6, 40, 3], 3, 3, 5, "tanh"), ("discrete", [16, 24, 12], 12, 4, 4, "tanh"),
         ("rk4", [2, 6, 1], 1, 1, 1, "tanh"), ("rk4", [3, 9, 8, 2], 2, 1, 3, "sigmoid"),
         ("unity", [4, 6, 6, 6, 2], 2, 2, 2, "softplus"), ("rk4", [3, 16, 16, 2], 2, 1, 7, "relu"), ("discrete", [5, 40, 40, 4], 4, 1, 5, "relu"), ("rk4", [5, 128, 128, 128, 4], 4, 1, 6, "tanh"),
         ("discrete", [16, 256, 256, 256, 256, 12], 12, 4, 3, "tanh"), ("rk4", [16, 64, 64, 12], 12, 4, 3, "tanh")]


@pytest.mark.parametrize("kind,dims,xd,ud,H,act", CASES)
@pytest.mark.parametrize("compute", ("float64", "float32"))
def test_generic_kernel_vs_oracle(kind, dims, xd, ud, H, act, compute):
    B = 5
    mlp, obj, Z, X0, lam, sig = _problem(dims, xd, ud, H, B, act)
    ref = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    ev = _evaluator(mlp, kind, H, compute, "generic", obj)
    np.testing.assert_array_equal(ev.hes_rows, BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).hes_rows)
    got = _run(ev, Z, X0, lam, sig)
    tol = TOL64 if compute == "float64" else TOL32
    for kr, kg in KEYS:
        assert _relerr(got[kg], ref[kr]) < tol, (kg, _relerr(got[kg], ref[kr]))
    assert ev.launch_count == 2
    ev.close()


@pytest.mark.parametrize("kind", ("discrete", "unity", "rk4"))
@pytest.mark.parametrize("h", (30, 32, 16))
def test_fast_kernel_vs_oracle(kind, h, lv_weights):
    H, B = 50, 300                      # 15000 steps: several CTAs, ragged last block
    mlp = MLP(lv_weights, 2, 1) if h == 30 else MLP.glorot([3, h, h, 2], 2, 1, seed=h)
    rng = np.random.default_rng(h)
    obj = SeparableQuadraticObjective.tracking(H, 2, 1, [1.0, 0.5], [0.3], x_ref=rng.uniform(-1, 1, (H, 2)))
    Z, X0 = rng.uniform(-1, 1, (B, H * 3)), rng.uniform(-1, 1, (B, 2))
    lam, sig = rng.standard_normal((B, H * 2)), rng.uniform(0.5, 1.5, B)
    ref = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    ev = _evaluator(mlp, kind, H, "float32", "fast", obj)
    assert "fast" in ev.kernel_name
    got = _run(ev, Z, X0, lam, sig)
    for kr, kg in KEYS:
        assert _relerr(got[kg], ref[kr]) < TOL32, (kg, _relerr(got[kg], ref[kr]))
    # reduced request sets run the cheaper kernel instantiations and must agree
    import torch
    t = lambda a: torch.as_tensor(a).cuda()
    o1 = ev.eval(t(Z), t(X0), want=("resid", "jac"))
    o0 = ev.eval(t(Z), t(X0), want=("resid",))
    torch.cuda.synchronize()
    assert _relerr(o1["jac"].cpu().numpy(), ref["jac_vals"]) < TOL32
    assert _relerr(o1["resid"].cpu().numpy(), ref["resid"]) < TOL32
    assert _relerr(o0["resid"].cpu().numpy(), ref["resid"]) < TOL32
    ev.close()


TC_CASES = [("rk4", [5, 128, 128, 128, 4], 4, 1, 13, 7), ("discrete", [5, 128, 128, 4], 4, 1, 9, 30),
            ("unity", [3, 128, 128, 128, 2], 2, 1, 25, 11), ("rk4", [3, 128, 128, 2], 2, 1, 50, 5),
            ("rk4", [4, 128, 128, 128, 3], 3, 1, 6, 9), ("discrete", [6, 128, 128, 4], 4, 2, 3, 17),
            ("rk4", [6, 128, 128, 128, 4], 4, 2, 4, 3),
            # hidden width 64: N = 64 MMAs, four K steps
            ("rk4", [5, 64, 64, 64, 4], 4, 1, 13, 7), ("discrete", [5, 64, 64, 4], 4, 1, 9, 30),
            ("unity", [3, 64, 64, 64, 2], 2, 1, 25, 11), ("rk4", [3, 64, 64, 2], 2, 1, 50, 5),
            # hidden width 32, three hidden layers (two-layer 32-wide LV nets belong to the register-resident kernel)
            ("rk4", [3, 32, 32, 32, 2], 2, 1, 50, 9), ("discrete", [5, 32, 32, 32, 4], 4, 1, 7, 12)]


@pytest.mark.parametrize("kind,dims,xd,ud,H,B", TC_CASES)
def test_tensor_core_kernel_vs_oracle(kind, dims, xd, ud, H, B):
    """tcgen05 kernel (split-f16 operands, forward second-order rows) against the float64 oracle: several tiles per
    launch with a ragged last tile; the reduced request sets run the residual-only / Jacobian-only row stacks."""
    import torch
    mlp, obj, Z, X0, lam, sig = _problem(dims, xd, ud, H, B, seed=len(dims) + H)
    ref = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    ev = _evaluator(mlp, kind, H, "float32", "tc", obj)
    assert "tcgen05" in ev.kernel_name
    got = _run(ev, Z, X0, lam, sig)
    for kr, kg in KEYS:
        assert _relerr(got[kg], ref[kr]) < TOL32, (kg, _relerr(got[kg], ref[kr]))
    t = lambda a: torch.as_tensor(a).cuda()
    o1 = ev.eval(t(Z), t(X0), want=("resid", "jac"))
    o0 = ev.eval(t(Z), t(X0), want=("resid",))
    torch.cuda.synchronize()
    assert _relerr(o1["jac"].cpu().numpy(), ref["jac_vals"]) < TOL32
    assert _relerr(o1["resid"].cpu().numpy(), ref["resid"]) < TOL32
    assert _relerr(o0["resid"].cpu().numpy(), ref["resid"]) < TOL32
    ev.close()
    # "auto" picks the tensor-core kernel for this class, and float32 I/O agrees too
    ev32 = _evaluator(mlp, kind, H, "float32", "auto", obj, io="float32")
    assert "tcgen05" in ev32.kernel_name
    got32 = _run(ev32, Z, X0, lam, sig)
    for kr, kg in KEYS[:3]:
        assert _relerr(got32[kg], ref[kr]) < TOL32, (kg, _relerr(got32[kg], ref[kr]))
    ev32.close()


@pytest.mark.parametrize("kind", ("discrete", "unity", "rk4"))
@pytest.mark.parametrize("H", (6, 25))
def test_against_reference_goldens(golden_dir, lv_weights, kind, H):
    """values recorded from the unmodified reference (IpoptProblem callbacks) -- both kernels."""
    g = np.load(os.path.join(golden_dir, f"ref_{kind}_H{H}.npz"))
    mlp = MLP(lv_weights, 2, 1)
    obj = SeparableQuadraticObjective(g["obj_lin"], g["obj_quad"], g["obj_ref"])
    jr, jc = np.nonzero(g["jacobian"])
    for kernel, compute, tol in (("fast", "float32", TOL32), ("generic", "float32", TOL32), ("generic", "float64", TOL64)):
        ev = _evaluator(mlp, kind, H, compute, kernel, obj)
        np.testing.assert_array_equal(ev.hes_rows, g["hes_rows"])
        np.testing.assert_array_equal(ev.hes_cols, g["hes_cols"])
        np.testing.assert_array_equal(ev.jac_rows, jr)
        np.testing.assert_array_equal(ev.jac_cols, jc)
        got = _run(ev, g["z"][None], g["x0"][None], g["lam"][None], np.asarray([float(g["sigma"])]))
        assert _relerr(got["resid"][0], g["constraints"]) < tol
        assert _relerr(got["jac"][0], g["jacobian"][jr, jc]) < tol
        assert _relerr(got["hes"][0], g["hessian_values"]) < tol
        assert _relerr(got["grad"][0], g["gradient"]) < tol
        assert abs(got["obj"][0] - float(g["objective"])) < tol * max(1.0, abs(float(g["objective"])))
        ev.close()


def test_float32_io_and_host_call(lv_weights):
    H, B = 25, 64
    mlp = MLP(lv_weights, 2, 1)
    rng = np.random.default_rng(5)
    obj = SeparableQuadraticObjective.control_setpoint(H, 2, 1, 2.0)
    Z, X0 = rng.uniform(-1, 1, (B, H * 3)), rng.uniform(-1, 1, (B, 2))
    lam, sig = rng.standard_normal((B, H * 2)), rng.uniform(0.5, 1.5, B)
    ref = BlockEvaluator(mlp, "rk4", H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    for io in ("float32", "float64"):
        ev = _evaluator(mlp, "rk4", H, "float32", "auto", obj, io=io)
        got = _run(ev, Z, X0, lam, sig)
        host = ev.eval_host(Z, X0, lam, sig)
        for kr, kg in KEYS:
            assert _relerr(got[kg], ref[kr]) < 2e-5, kg
            np.testing.assert_array_equal(host[kg], got[kg])          # host-buffer call == device call, bit for bit
        # scalar objective factor and single-problem (1-D) input
        one = ev.eval_host(Z[0], X0[0], lam[0], 0.25)
        r1 = BlockEvaluator(mlp, "rk4", H, DT=0.1, objective=obj).evaluate(Z[:1], X0[:1], lam[:1], 0.25)
        assert _relerr(one["hes"], r1["hes_vals"]) < 2e-5
        ev.close()


def test_blocks_and_model_entry_points():
    import torch
    mlp, _, Z, X0, _, _ = _problem([5, 12, 9, 4], 4, 1, 4, 3)
    ev_o = BlockEvaluator(mlp, "rk4", 4, DT=0.1)
    _, _, xprev = ev_o.split(Z, X0)
    zz = np.concatenate([xprev, Z[:, 16:].reshape(3, 4, 1)], axis=2).reshape(-1, 5)
    pred, AB, Hb = step_blocks(mlp, "rk4", 0.1, zz)
    ev = _evaluator(mlp, "rk4", 4, "float64", "generic")
    p2, AB2, Hb2 = ev.eval_blocks(torch.as_tensor(Z).cuda(), torch.as_tensor(X0).cuda())
    assert _relerr(p2.cpu().numpy().reshape(-1, 4), pred + zz[:, :4]) < TOL64
    assert _relerr(AB2.cpu().numpy().reshape(-1, 4, 5), AB + np.eye(4, 5)[None]) < TOL64
    assert _relerr(Hb2.cpu().numpy().reshape(-1, 4, 5, 5), Hb) < TOL64
    f, J, Hs = mlp.blocks(zz)
    f2, J2, Hs2 = ev.model_eval(torch.as_tensor(zz).cuda())
    assert _relerr(f2.cpu().numpy(), f) < TOL64 and _relerr(J2.cpu().numpy(), J) < TOL64 and _relerr(Hs2.cpu().numpy(), Hs) < TOL64
    ev.close()


def test_edge_cases(lv_weights):
    import torch
    from pyneuralempc_b200 import NlpEvaluator
    from pyneuralempc_b200._lib import NempcError
    mlp = MLP(lv_weights, 2, 1)
    ev = _evaluator(mlp, "rk4", 1, "float32")                  # H = 1: no state-state block at all
    assert ev.nnz_hes == 1 and ev.nnz_jac == 2 * 2
    z = torch.zeros((0, 3), dtype=torch.float64).cuda()        # empty batch is a no-op
    out = ev.eval(z, torch.zeros((0, 2), dtype=torch.float64).cuda(), torch.zeros((0, 2), dtype=torch.float64).cuda())
    assert out["resid"].shape == (0, 2)
    with pytest.raises(ValueError):
        ev.eval(torch.zeros((1, 3), dtype=torch.float64).cuda(), torch.zeros((1, 2), dtype=torch.float64).cuda(), want=("hes",))
    ev.close()
    with pytest.raises(NempcError):                             # fast kernel requested for a shape it does not cover
        NlpEvaluator(MLP.glorot([3, 7, 7, 2], 2, 1).weights, 2, 1, 5, "rk4", DT=0.1, kernel="fast")
    with pytest.raises(ValueError):
        NlpEvaluator(mlp.weights, 2, 1, 5, "rk4")               # RK4 without DT


def test_full_size_c2_properties():
    """BASELINE configs[1] at full size (B=4096, H=50: 204800 steps), size-independent checks:
    fast == generic kernel; Hessian linear in lambda and in the objective factor; batch-permutation equivariance;
    a random subsample equals the oracle."""
    import torch
    H, B = 50, 4096
    rng = np.random.default_rng(1234)
    mlp = MLP.glorot([3, 30, 30, 2], 2, 1, seed=0)
    obj = SeparableQuadraticObjective.tracking(H, 2, 1, [1.0, 2.0], [0.1])
    Z, X0 = rng.uniform(-1, 1, (B, H * 3)), rng.uniform(-1, 1, (B, 2))
    l1, l2 = rng.standard_normal((B, H * 2)), rng.standard_normal((B, H * 2))
    fast = _evaluator(mlp, "rk4", H, "float32", "fast", obj)
    gen = _evaluator(mlp, "rk4", H, "float32", "generic", obj)
    ones, zeros = np.ones(B), np.zeros(B)
    a = _run(fast, Z, X0, l1, ones)
    g = _run(gen, Z, X0, l1, ones)
    for k in ("resid", "jac", "hes", "obj", "grad"):
        assert _relerr(a[k], g[k]) < TOL32, k
    b = _run(fast, Z, X0, l2, zeros)
    c = _run(fast, Z, X0, l1 + l2, ones)
    assert _relerr(c["hes"], a["hes"] + b["hes"]) < TOL32
    perm = rng.permutation(B)
    p = _run(fast, Z[perm], X0[perm], l1[perm], ones)
    for k in ("resid", "jac", "hes", "obj", "grad"):
        np.testing.assert_array_equal(p[k], a[k][perm])
    idx = rng.choice(B, 16, replace=False)
    ref = BlockEvaluator(mlp, "rk4", H, DT=0.1, objective=obj).evaluate(Z[idx], X0[idx], l1[idx], 1.0)
    for kr, kg in KEYS:
        assert _relerr(a[kg][idx], ref[kr]) < TOL32, kg
    # constant structural entries
    assert (a["jac"][:, fast.jac_rows == fast.jac_cols] == -1.0).all()
    fast.close(); gen.close()


def test_eval_host_graph_replay_tracks_inputs_weights_and_objective(lv_weights):
    """nempc_eval_host replays its chunk pipeline as a CUDA graph from the third call with unchanged buffers: every replay
    must read the CURRENT contents of the pinned inputs, and new weights / a new objective must invalidate the graph
    (kernel parameters are captured by value)."""
    H, B = 20, 700                                     # >= 512 problems: the multi-chunk pipeline
    mlp = MLP(lv_weights, 2, 1)
    rng = np.random.default_rng(5)
    obj = SeparableQuadraticObjective.tracking(H, 2, 1, [1.0, 0.5], [0.3], x_ref=rng.uniform(-1, 1, (H, 2)))
    ev = _evaluator(mlp, "rk4", H, "float32", "auto", obj)
    buf = ev.pinned_buffers(B)
    launches = []
    for it in range(5):
        Z, X0, lam = rng.uniform(-1, 1, (B, H * 3)), rng.uniform(-1, 1, (B, 2)), rng.standard_normal((B, H * 2))
        buf["z"][...] = Z; buf["x0"][...] = X0; buf["lam"][...] = lam
        l0 = ev.launch_count
        out = ev.eval_pinned(B, 0.7)
        launches.append(ev.launch_count - l0)
        ref = BlockEvaluator(mlp, "rk4", H, DT=0.1, objective=obj).evaluate(Z, X0, lam, 0.7)
        for kr, kg in KEYS:
            assert _relerr(out[kg], ref[kr]) < TOL32, (it, kg)
    assert len(set(launches)) == 1 and launches[0] > 0          # replays account for the same kernels as direct submission
    # new weights and a new objective: same buffers, different answers
    mlp2 = MLP.glorot([3, 30, 30, 2], 2, 1, seed=77)
    for l, (W, b) in enumerate(mlp2.weights):
        ev.set_weights(l, W, b)
    obj2 = SeparableQuadraticObjective.tracking(H, 2, 1, [2.0, 0.1], [0.9], x_ref=rng.uniform(-1, 1, (H, 2)))
    ev.set_objective(obj2.lin, obj2.quad, obj2.ref)
    out = ev.eval_pinned(B, 0.7)
    ref = BlockEvaluator(mlp2, "rk4", H, DT=0.1, objective=obj2).evaluate(Z, X0, lam, 0.7)
    for kr, kg in KEYS:
        assert _relerr(out[kg], ref[kr]) < TOL32, ("after update", kg)
    ev.close()


def test_c3_full_size_properties_tensor_core():
    """BASELINE config C3 at full size (cart-pole MLP 5-128-128-128-4, RK4, H=100, B=16384 -> 1 638 400 horizon steps,
    tensor-core kernel): size-independent properties plus oracle / generic-kernel agreement on a sample of problems.
    float32 I/O keeps the value arrays of two evaluations (2 x 6.2 GB in float64) small."""
    import torch
    H, B = 100, 16384
    rng = np.random.default_rng(3)
    mlp = MLP.glorot([5, 128, 128, 128, 4], 4, 1, seed=0)
    obj = SeparableQuadraticObjective.tracking(H, 4, 1, [1.0, 2.0, 0.5, 1.5], [0.1])
    tc = _evaluator(mlp, "rk4", H, "float32", "auto", obj, io="float32")
    assert "tcgen05" in tc.kernel_name
    t = lambda a: torch.as_tensor(a, dtype=torch.float32).cuda()
    Z, X0 = t(rng.uniform(-1, 1, (B, H * 5))), t(rng.uniform(-1, 1, (B, 4)))
    l1, l2 = t(rng.standard_normal((B, H * 4))), t(rng.standard_normal((B, H * 4)))
    a = {k: v.clone() for k, v in tc.eval(Z, X0, l1, 1.0).items()}
    # (1) Hessian is linear in lambda and in the objective factor: hes(l1 + l2, 1) = hes(l1, 1) + hes(l2, 0)
    b = tc.eval(Z, X0, l2, 0.0, want=("hes",))["hes"].clone()
    c = tc.eval(Z, X0, l1 + l2, 1.0, want=("hes",))["hes"]
    scale = float(a["hes"].abs().max())
    assert float((c - (a["hes"] + b)).abs().max()) < 2e-5 * scale            # f32 I/O rounds each stored value
    del b, c
    # (2) batch-permutation equivariance (bit-exact: a problem's values do not depend on its position in the batch)
    perm = torch.as_tensor(rng.permutation(B)).cuda()
    p = tc.eval(Z[perm], X0[perm], l1[perm], 1.0)
    for k in ("resid", "jac", "hes"):
        assert torch.equal(p[k], a[k][perm]), k
    # (3) constant structural entries and finiteness everywhere
    jr, jc = torch.as_tensor(tc.jac_rows).cuda(), torch.as_tensor(tc.jac_cols).cuda()
    assert bool((a["jac"][:, jr == jc] == -1.0).all())
    assert all(bool(torch.isfinite(a[k]).all()) for k in ("resid", "jac", "hes", "obj", "grad"))
    # (4) a sample of problems against the float64 oracle and against the FFMA generic kernel
    idx = rng.choice(B, 6, replace=False)
    Zs, X0s, ls = (v[torch.as_tensor(idx).cuda()].double().cpu().numpy() for v in (Z, X0, l1))
    ref = BlockEvaluator(mlp, "rk4", H, DT=0.1, objective=obj).evaluate(Zs, X0s, ls, 1.0)
    for kr, kg in KEYS:
        assert _relerr(a[kg][torch.as_tensor(idx).cuda()].double().cpu().numpy(), ref[kr]) < TOL32, kg
    gen = _evaluator(mlp, "rk4", H, "float32", "generic", obj)
    g = _run(gen, Zs, X0s, ls, np.ones(len(idx)))
    for kr, kg in KEYS[:3]:
        assert _relerr(g[kg], ref[kr]) < TOL32, kg
    tc.close(); gen.close()


@pytest.mark.parametrize("kind", ("discrete", "unity", "rk4"))
@pytest.mark.parametrize("h", (30, 32, 16))
def test_small_batch_warp_kernel_vs_oracle(kind, h, lv_weights):
    """below 4096 horizon steps `auto` runs the warp-per-step kernel (nempc_small.cuh): oracle parity for all request sets, and
    agreement with the thread-per-step kernel (different summation order, so float32 rounding, not bit equality)."""
    import torch
    H, B = 25, 3                         # BASELINE config C1 shape; 75 steps = 75 warps, ragged last CTA
    mlp = MLP(lv_weights, 2, 1) if h == 30 else MLP.glorot([3, h, h, 2], 2, 1, seed=h)
    rng = np.random.default_rng(h + 1)
    obj = SeparableQuadraticObjective.tracking(H, 2, 1, [1.0, 0.5], [0.3], x_ref=rng.uniform(-1, 1, (H, 2)))
    Z, X0 = rng.uniform(-1, 1, (B, H * 3)), rng.uniform(-1, 1, (B, 2))
    lam, sig = rng.standard_normal((B, H * 2)), rng.uniform(0.5, 1.5, B)
    ref = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    ev = _evaluator(mlp, kind, H, "float32", "auto", obj)
    assert "nempc_small_kernel" in ev.kernel_name
    got = _run(ev, Z, X0, lam, sig)
    for kr, kg in KEYS:
        assert _relerr(got[kg], ref[kr]) < TOL32, (kg, _relerr(got[kg], ref[kr]))
    t = lambda a: torch.as_tensor(a).cuda()
    o1 = ev.eval(t(Z), t(X0), want=("resid", "jac"))
    o0 = ev.eval(t(Z), t(X0), want=("resid",))
    torch.cuda.synchronize()
    assert _relerr(o1["jac"].cpu().numpy(), ref["jac_vals"]) < TOL32
    assert _relerr(o0["resid"].cpu().numpy(), ref["resid"]) < TOL32
    fast = _evaluator(mlp, kind, H, "float32", "fast", obj)
    gf = _run(fast, Z, X0, lam, sig)
    for k in ("resid", "jac", "hes"):
        assert _relerr(got[k], gf[k]) < 2e-6, k
    ev.close(); fast.close()
    ev32 = _evaluator(mlp, kind, H, "float32", "auto", obj, io="float32")      # float32 I/O instantiation of the same kernel
    g32 = _run(ev32, Z, X0, lam, sig)
    for kr, kg in KEYS[:3]:
        assert _relerr(g32[kg], ref[kr]) < TOL32, kg
    # and through the host-buffer call: zero-copy on the mapped pinned buffers (a single solver callback)
    buf = ev32.pinned_buffers(B, per_problem_factor=True)
    buf["z"][...] = Z; buf["x0"][...] = X0; buf["lam"][...] = lam; buf["sig"][...] = sig
    out = ev32.eval_pinned(B, per_problem_factor=True)
    for kr, kg in KEYS:
        assert _relerr(out[kg], ref[kr]) < TOL32, kg
    ev32.close()


@pytest.mark.parametrize("kind,dims,xd,ud,H,B", [("rk4", [5, 128, 128, 128, 4], 4, 1, 1, 1), ("discrete", [3, 128, 128, 2], 2, 1, 1, 3),
                                                  ("unity", [6, 128, 128, 128, 4], 4, 2, 2, 1), ("rk4", [4, 128, 128, 3], 3, 1, 5, 2),
                                                  ("rk4", [5, 64, 64, 64, 4], 4, 1, 1, 1), ("discrete", [3, 64, 64, 2], 2, 1, 1, 3),
                                                  ("rk4", [3, 32, 32, 32, 2], 2, 1, 2, 1)])
def test_tensor_core_kernel_edge_sizes(kind, dims, xd, ud, H, B):
    """fewer horizon steps than one row tile holds (the second tile group of the CTA stays idle), H = 1 (no A block, no x-x Hessian
    block), and an empty batch."""
    import torch
    mlp, obj, Z, X0, lam, sig = _problem(dims, xd, ud, H, B, seed=H + B)
    ref = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    ev = _evaluator(mlp, kind, H, "float32", "tc", obj)
    got = _run(ev, Z, X0, lam, sig)
    for kr, kg in KEYS:
        assert got[kg].shape == ref[kr].shape
        assert _relerr(got[kg], ref[kr]) < TOL32, (kg, _relerr(got[kg], ref[kr]))
    empty = ev.eval(torch.zeros((0, ev.n), dtype=torch.float64).cuda(), torch.zeros((0, xd), dtype=torch.float64).cuda(),
                    torch.zeros((0, ev.m), dtype=torch.float64).cuda(), 1.0)
    assert all(v.shape[0] == 0 for v in empty.values())
    ev.close()


@pytest.mark.parametrize("kernel,dims,xd,ud", [("tc", [5, 128, 128, 4], 4, 1), ("fast", [3, 30, 30, 2], 2, 1), ("auto", [3, 30, 30, 2], 2, 1),
                                                ("generic", [5, 12, 9, 4], 4, 1)])
def test_evaluation_without_an_objective(kernel, dims, xd, ud):
    """no nempc_set_objective: residual / Jacobian / Hessian still evaluate (the Hessian pattern has no objective-only diagonal of
    x_H, obj / grad are refused) -- the reference's IpoptProblem with a zero cost (optimizer/ipopt.py:55-62)."""
    import torch
    from pyneuralempc_b200._lib import NempcError
    H, B = 7, 9
    mlp, _, Z, X0, lam, sig = _problem(dims, xd, ud, H, B, seed=11)
    oe = BlockEvaluator(mlp, "rk4", H, DT=0.1, objective=None)
    ref = oe.evaluate(Z, X0, lam, sig)
    ev = _evaluator(mlp, "rk4", H, "float32", kernel, None)
    np.testing.assert_array_equal(ev.hes_rows, oe.hes_rows)
    np.testing.assert_array_equal(ev.hes_cols, oe.hes_cols)
    t = lambda a: torch.as_tensor(a).cuda()
    got = ev.eval(t(Z), t(X0), t(lam), t(sig), want=("resid", "jac", "hes"))
    torch.cuda.synchronize()
    for kr, kg in KEYS[:3]:
        assert _relerr(got[kg].cpu().numpy(), ref[kr]) < TOL32, kg
    assert "obj" not in ev.eval(t(Z), t(X0), t(lam), t(sig), want=("resid", "obj", "grad"))     # the Python layer drops them ...
    buf = (ctypes.c_double * B)()
    rc = ev.lib.nempc_eval(ev._h, B, ctypes.c_void_p(t(Z).data_ptr()), ctypes.c_void_p(t(X0).data_ptr()), None, None, 1.0,
                           None, None, None, ctypes.c_void_p(t(np.zeros(B)).data_ptr()), None, None)
    assert rc == -4                                                                              # ... and the C ABI answers NEMPC_ESTATE
    del buf, NempcError
    ev.close()


@pytest.mark.parametrize("hw", (128, 64, 32))
def test_tensor_core_kernel_is_deterministic(hw):
    """two row tiles share a CTA through named barriers, mbarriers and tensor memory: repeated evaluations of the same batch
    must be bit-identical (a race between the tile groups would show up here long before it shows up in a tolerance)."""
    import torch
    H, B = 50, 300
    mlp, obj, Z, X0, lam, sig = _problem([5, hw, hw, hw, 4], 4, 1, H, B, seed=9)
    ev = _evaluator(mlp, "rk4", H, "float32", "tc", obj)
    t = lambda a: torch.as_tensor(a).cuda()
    z, x0, lm, sg = t(Z), t(X0), t(lam), t(sig)
    ref = {k: v.clone() for k, v in ev.eval(z, x0, lm, sg).items()}
    for _ in range(12):
        out = ev.eval(z, x0, lm, sg)
        torch.cuda.synchronize()
        for k in ref:
            assert torch.equal(out[k], ref[k]), k
    ev.close()


def _wide_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    dims = [int(v) for v in g["dims"]]
    mlp = MLP.glorot(dims, int(g["x_dim"]), int(g["u_dim"]), seed=int(g["net_seed"]), dtype=np.float32)
    chk = float(sum(np.abs(np.asarray(W, np.float64)).sum() + np.abs(np.asarray(b, np.float64)).sum() for W, b in mlp.weights))
    assert chk == float(g["weights_checksum"]), "MLP.glorot no longer regenerates the weights the golden file was recorded with"
    return g, mlp


@pytest.mark.parametrize("name,kernel_tag", (("ref_rk4_w128_H6.npz", "nempc_tc_kernel"), ("ref_discrete_w256_H6.npz", "nempc_wide_kernel"),
                                             ("ref_rk4_w256_H6.npz", "nempc_wide_kernel")))
def test_tensor_core_kernels_against_reference_goldens(golden_dir, name, kernel_tag):
    """the two tcgen05 kernels pinned DIRECTLY to the unmodified reference (not only to the oracle): IpoptProblem callbacks of
    a 3-128-128-2 network under the reference's RK4 integrator (nempc_tc.cuh, TcCfg<2,1,2,...,128>) and of a 5-256-256-256-4 network
    under its DiscretIntegrator (nempc_wide.cuh); tests/golden/make_golden.py record_wide()."""
    g, mlp = _wide_golden(golden_dir, name)
    kind, H = str(g["kind"]), int(g["H"])
    obj = SeparableQuadraticObjective(g["obj_lin"], g["obj_quad"], g["obj_ref"])
    jr, jc = np.nonzero(g["jacobian"])
    for kernel in ("tc", "auto", "generic"):
        ev = _evaluator(mlp, kind, H, "float32", kernel, obj)
        assert (kernel_tag in ev.kernel_name) == (kernel != "generic"), ev.kernel_name
        np.testing.assert_array_equal(ev.hes_rows, g["hes_rows"]); np.testing.assert_array_equal(ev.hes_cols, g["hes_cols"])
        np.testing.assert_array_equal(ev.jac_rows, jr); np.testing.assert_array_equal(ev.jac_cols, jc)
        got = _run(ev, g["z"][None], g["x0"][None], g["lam"][None], np.asarray([float(g["sigma"])]))
        for k, ref in (("resid", g["constraints"]), ("jac", g["jacobian"][jr, jc]), ("hes", g["hessian_values"]), ("grad", g["gradient"])):
            assert _relerr(got[k][0], ref) < TOL32, (kernel, k, _relerr(got[k][0], ref))
        assert abs(got["obj"][0] - float(g["objective"])) < TOL32 * max(1.0, abs(float(g["objective"])))
        ev.close()


WIDE_CASES = [("discrete", [16, 256, 256, 256, 256, 12], 12, 4, 5, 3),        # BASELINE config C4's network (quadrotor 16 -> 256 x 4 -> 12)
              ("discrete", [16, 256, 256, 256, 256, 12], 12, 4, 37, 9),       # several super-tiles of 128 steps, ragged tail, odd tile pairing
              ("unity", [16, 256, 256, 12], 12, 4, 7, 2),
              ("discrete", [5, 256, 256, 256, 4], 4, 1, 11, 5), ("discrete", [3, 256, 256, 2], 2, 1, 6, 4),
              ("discrete", [8, 256, 256, 256, 6], 6, 2, 6, 4), ("unity", [3, 256, 256, 256, 2], 2, 1, 300, 3),
              ("discrete", [16, 256, 256, 12], 12, 4, 1, 1), ("rk4", [16, 256, 256, 12], 12, 4, 1, 1),      # one horizon step in the whole launch
              # RK4: forward sweep (k_s, dk_s), last stage with curvature, backward sweep with w_s = c_s lambda + a_{s+1} J_{s+1,x}^T w_{s+1}
              ("rk4", [16, 256, 256, 256, 256, 12], 12, 4, 5, 3), ("rk4", [16, 256, 256, 256, 256, 12], 12, 4, 29, 10),
              ("rk4", [5, 256, 256, 256, 4], 4, 1, 11, 5), ("rk4", [3, 256, 256, 2], 2, 1, 50, 4), ("rk4", [8, 256, 256, 256, 6], 6, 2, 6, 4),
              # shapes without a specialised instantiation: dimensions read at run time, 4 / 8 / 16 tangent rows
              ("discrete", [4, 256, 256, 3], 3, 1, 9, 4), ("rk4", [4, 256, 256, 3], 3, 1, 9, 4), ("unity", [2, 256, 256, 1], 1, 1, 6, 3),
              ("discrete", [7, 256, 256, 256, 5], 5, 2, 8, 3), ("rk4", [7, 256, 256, 5], 5, 2, 8, 3),
              ("discrete", [12, 256, 256, 10], 10, 2, 5, 3), ("rk4", [14, 256, 256, 256, 8], 8, 6, 7, 3), ("unity", [16, 256, 256, 9], 9, 7, 5, 2),
              ("discrete", [16, 256, 256, 16], 16, 0, 4, 2) if False else ("discrete", [16, 256, 256, 15], 15, 1, 4, 2)]


@pytest.mark.parametrize("kind,dims,xd,ud,H,B", WIDE_CASES)
def test_wide_kernel_vs_oracle(kind, dims, xd, ud, H, B):
    """width-256 tcgen05 kernel (adjoint form, streamed split-f16 weights, CTA pairs; RK4 = two sweeps over the stages) against the
    float64 oracle, all request sets"""
    import torch
    mlp, obj, Z, X0, lam, sig = _problem(dims, xd, ud, H, B, seed=len(dims) + H)
    ref = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    ev = _evaluator(mlp, kind, H, "float32", "auto", obj)
    assert "nempc_wide_kernel" in ev.kernel_name
    got = _run(ev, Z, X0, lam, sig)
    for kr, kg in KEYS:
        assert _relerr(got[kg], ref[kr]) < TOL32, (kg, _relerr(got[kg], ref[kr]))
    t = lambda a: torch.as_tensor(a).cuda()
    o1 = ev.eval(t(Z), t(X0), want=("resid", "jac"))
    o0 = ev.eval(t(Z), t(X0), want=("resid",))
    torch.cuda.synchronize()
    assert _relerr(o1["jac"].cpu().numpy(), ref["jac_vals"]) < TOL32
    assert _relerr(o1["resid"].cpu().numpy(), ref["resid"]) < TOL32
    assert _relerr(o0["resid"].cpu().numpy(), ref["resid"]) < TOL32
    # bit-determinism: the same launch twice
    again = _run(ev, Z, X0, lam, sig)
    for k in ("resid", "jac", "hes"):
        np.testing.assert_array_equal(again[k], got[k])
    ev.close()


def test_c4_properties_wide_kernel():
    """BASELINE config C4's network and horizon (16 -> 256 x 4 -> 12, H = 200) on a batch the oracle cannot follow: size-independent
    properties -- linearity of the Hessian values in lambda, batch-permutation equivariance (bit-exact), the structural -1 entries, and
    a sample of problems against the oracle."""
    import torch
    H, B, xd, ud = 200, 512, 12, 4
    mlp = MLP.glorot([16, 256, 256, 256, 256, 12], xd, ud, seed=0, dtype=np.float32)
    rng = np.random.default_rng(9)
    n, m = H * (xd + ud), H * xd
    Z, X0 = rng.uniform(-1, 1, (B, n)), rng.uniform(-1, 1, (B, xd))
    l1, l2 = rng.standard_normal((B, m)), rng.standard_normal((B, m))
    ev = _evaluator(mlp, "discrete", H, "float32", "auto")
    assert "nempc_wide_kernel" in ev.kernel_name
    t = lambda a: torch.as_tensor(a).cuda()
    h = lambda lam: ev.eval(t(Z), t(X0), t(lam), 0.0, want=("resid", "jac", "hes"))
    o1 = {k: v.clone() for k, v in h(l1).items()}
    o2 = {k: v.clone() for k, v in h(l2).items()}
    o3 = {k: v.clone() for k, v in h(2.0 * l1 - 0.5 * l2).items()}
    lin = 2.0 * o1["hes"] - 0.5 * o2["hes"]
    assert float((o3["hes"] - lin).abs().max()) < 2e-5 * float(lin.abs().max())
    assert torch.equal(o1["jac"], o2["jac"]) and torch.equal(o1["resid"], o3["resid"])          # lambda does not touch them
    perm = rng.permutation(B)
    op = ev.eval(t(Z[perm]), t(X0[perm]), t(l1[perm]), 0.0, want=("resid", "jac", "hes"))
    for k in ("resid", "jac", "hes"):
        assert torch.equal(op[k], o1[k][torch.as_tensor(perm).cuda()]), k
    minus1 = np.nonzero((ev.jac_cols < H * xd) & (ev.jac_cols // xd == ev.jac_rows // xd))[0]
    assert len(minus1) == m and bool((o1["jac"][:, torch.as_tensor(minus1).cuda()] == -1.0).all())
    pick = [0, 255, 511]
    ref = BlockEvaluator(mlp, "discrete", H, DT=0.1).evaluate(Z[pick], X0[pick], l1[pick], 0.0)
    for kr, kg in KEYS[:3]:
        assert _relerr(o1[kg][pick].cpu().numpy(), ref[kr]) < TOL32, (kg, _relerr(o1[kg][pick].cpu().numpy(), ref[kr]))
    ev.close()


@pytest.mark.parametrize("kind", ("discrete", "unity", "rk4"))
@pytest.mark.parametrize("h", (30, 32, 16))
def test_float64_register_resident_kernel_vs_oracle(kind, h, lv_weights):
    """nempc_fast64_kernel (thread per step, DFMA): the 1e-10 parity mode at register-resident speed"""
    import torch
    H, B = 50, 300
    mlp = MLP(lv_weights, 2, 1) if h == 30 else MLP.glorot([3, h, h, 2], 2, 1, seed=h)
    rng = np.random.default_rng(h + 1)
    obj = SeparableQuadraticObjective.tracking(H, 2, 1, [1.0, 0.5], [0.3], x_ref=rng.uniform(-1, 1, (H, 2)))
    Z, X0 = rng.uniform(-1, 1, (B, H * 3)), rng.uniform(-1, 1, (B, 2))
    lam, sig = rng.standard_normal((B, H * 2)), rng.uniform(0.5, 1.5, B)
    ref = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    for kernel in ("auto", "fast"):
        ev = _evaluator(mlp, kind, H, "float64", kernel, obj)
        assert "nempc_fast64_kernel" in ev.kernel_name
        got = _run(ev, Z, X0, lam, sig)
        for kr, kg in KEYS:
            assert _relerr(got[kg], ref[kr]) < TOL64, (kg, _relerr(got[kg], ref[kr]))
        t = lambda a: torch.as_tensor(a).cuda()
        o1 = ev.eval(t(Z), t(X0), want=("resid", "jac"))
        o0 = ev.eval(t(Z), t(X0), want=("resid",))
        torch.cuda.synchronize()
        assert _relerr(o1["jac"].cpu().numpy(), ref["jac_vals"]) < TOL64 and _relerr(o0["resid"].cpu().numpy(), ref["resid"]) < TOL64
        ev.close()
    # a forced register-resident evaluation of ONE small problem against the reference golden (auto would take the generic kernel here)
    if h == 30 and kind == "rk4":
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_rk4_H25.npz"))
        gobj = SeparableQuadraticObjective(g["obj_lin"], g["obj_quad"], g["obj_ref"])
        ev = _evaluator(mlp, "rk4", 25, "float64", "fast", gobj)
        jr, jc = np.nonzero(g["jacobian"])
        got = _run(ev, g["z"][None], g["x0"][None], g["lam"][None], np.asarray([float(g["sigma"])]))
        assert _relerr(got["resid"][0], g["constraints"]) < TOL64 and _relerr(got["jac"][0], g["jacobian"][jr, jc]) < TOL64
        assert _relerr(got["hes"][0], g["hessian_values"]) < TOL64
        ev.close()


@pytest.mark.parametrize("kind,dims,xd,ud,H,B", [("rk4", [5, 128, 128, 128, 4], 4, 1, 13, 7), ("discrete", [5, 128, 128, 4], 4, 1, 9, 30),
                                                ("unity", [3, 128, 128, 128, 2], 2, 1, 25, 11), ("rk4", [3, 128, 128, 2], 2, 1, 50, 5)])
def test_adjoint_form_kernel_at_hidden_width_128(kind, dims, xd, ud, H, B, monkeypatch):
    """nempc_wide_kernel<..., HW = 128>: by default the single-stage integrators send their Hessian evaluations there (1.37x faster than the
    forward second-order kernel on that case) while RK4 and Jacobian-only calls stay on nempc_tc_kernel; NEMPC_WIDE128=1 routes everything
    to it, =0 nothing.  All three routings against the oracle."""
    mlp, obj, Z, X0, lam, sig = _problem(dims, xd, ud, H, B, seed=len(dims) + H)
    ref = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    for setting in (None, "1", "0"):
        if setting is None:
            monkeypatch.delenv("NEMPC_WIDE128", raising=False)
        else:
            monkeypatch.setenv("NEMPC_WIDE128", setting)
        ev = _evaluator(mlp, kind, H, "float32", "auto", obj)
        name = ev.kernel_name
        if setting == "1":
            assert name.startswith("nempc_wide_kernel") and "x128" in name
        elif setting == "0":
            assert name.startswith("nempc_tc_kernel") and "nempc_wide_kernel" not in name
        else:
            assert name.startswith("nempc_tc_kernel") and (("nempc_wide_kernel" in name) == (kind != "rk4"))
        got = _run(ev, Z, X0, lam, sig)
        for kr, kg in KEYS:
            assert _relerr(got[kg], ref[kr]) < TOL32, (setting, kg, _relerr(got[kg], ref[kr]))
        ev.close()


def test_c4_named_size_full_batch():
    """BASELINE config C4 at its NAMED size (16 -> 256 x 4 -> 12, H = 200, B = 65 536: 13.1 M horizon steps, 37 GB of values in one launch):
    structural entries over the whole batch, a sample of problems against the oracle, and a sub-batch evaluated alone must reproduce its
    slice of the big launch bit for bit (the rows of a GEMM tile do not see each other, wherever a step lands in the tiling)."""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 80e9:
        pytest.skip("needs ~45 GB of device memory")
    H, B, xd, ud = 200, 65536, 12, 4
    mlp = MLP.glorot([16, 256, 256, 256, 256, 12], xd, ud, seed=0, dtype=np.float32)
    n, m = H * (xd + ud), H * xd
    ev = _evaluator(mlp, "discrete", H, "float32", "auto")
    assert "nempc_wide_kernel" in ev.kernel_name
    gen = torch.Generator(device="cuda"); gen.manual_seed(7)
    z = torch.rand((B, n), dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    x0 = torch.rand((B, xd), dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    lam = torch.randn((B, m), dtype=torch.float64, device="cuda", generator=gen)
    out = ev.eval(z, x0, lam, 0.0, want=("resid", "jac", "hes"))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out["hes"]).all()) and bool(torch.isfinite(out["jac"]).all())
    minus1 = torch.as_tensor(np.nonzero((ev.jac_cols < H * xd) & (ev.jac_cols // xd == ev.jac_rows // xd))[0]).cuda()
    assert len(minus1) == m and bool((out["jac"][:, minus1] == -1.0).all())
    lo, hi = 1000, 1512                                       # 200 000 steps in: not aligned to the 128-step tiles
    sub = ev.eval(z[lo:hi].contiguous(), x0[lo:hi].contiguous(), lam[lo:hi].contiguous(), 0.0, want=("resid", "jac", "hes"))
    for k in ("resid", "jac", "hes"):
        assert torch.equal(sub[k], out[k][lo:hi]), k
    pick = [0, 31337, B - 1]
    ref = BlockEvaluator(mlp, "discrete", H, DT=0.1).evaluate(z[pick].cpu().numpy(), x0[pick].cpu().numpy(), lam[pick].cpu().numpy(), 0.0)
    for kr, kg in KEYS[:3]:
        assert _relerr(out[kg][pick].cpu().numpy(), ref[kr]) < TOL32, (kg, _relerr(out[kg][pick].cpu().numpy(), ref[kr]))
    ev.close()
    del out, z, lam
    torch.cuda.empty_cache()


@pytest.mark.parametrize("kind", ("discrete", "rk4"))
@pytest.mark.parametrize("scale", (1e6, 1e-7))
def test_wide_kernel_over_the_range_of_solver_multipliers(kind, scale):
    """IPOPT's multipliers span many decades, the adjoint of the width-256 kernel travels through f16-split operands: each step works with
    lambda / max|lambda| and scales its Hessian block back, so 1e6 x and 1e-7 x the usual multipliers (and a step with lambda = 0, and one
    whose multipliers differ by 1e6 among themselves) give the same RELATIVE accuracy"""
    dims, xd, ud, H, B = [16, 256, 256, 12], 12, 4, 7, 3
    mlp, obj, Z, X0, lam, sig = _problem(dims, xd, ud, H, B, seed=17)
    lam = lam * scale
    lam[0, :xd] = 0.0                                          # a step without multipliers
    lam[1, xd:2 * xd] *= np.logspace(0, 6, xd)                 # a step whose multipliers span six decades
    obj.quad[:] = 0.0                                          # the constraint part alone: the cost's diagonal would mask small entries
    ref = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    ev = _evaluator(mlp, kind, H, "float32", "auto", obj)
    assert "nempc_wide_kernel" in ev.kernel_name
    got = _run(ev, Z, X0, lam, sig)
    assert np.isfinite(got["hes"]).all()
    for b in range(B):                                         # per problem: the blocks of problem 1 are 1e6 x larger than the others'
        assert _relerr(got["hes"][b], ref["hes_vals"][b]) < TOL32, (b, _relerr(got["hes"][b], ref["hes_vals"][b]))
    assert _relerr(got["jac"], ref["jac_vals"]) < TOL32
    ev.close()


DMMA_CASES = [("rk4", [5, 128, 128, 128, 4], 4, 1, 7, 9), ("discrete", [3, 128, 128, 2], 2, 1, 5, 30), ("unity", [4, 64, 64, 3], 3, 1, 6, 11),
              ("rk4", [6, 64, 64, 64, 64, 4], 4, 2, 5, 13), ("discrete", [8, 128, 128, 6], 6, 2, 4, 40), ("rk4", [8, 128, 128, 128, 6], 6, 2, 3, 7),
              ("rk4", [3, 128, 128, 128, 128, 2], 2, 1, 50, 12), ("discrete", [5, 64, 64, 4], 4, 1, 1, 1)]


@pytest.mark.parametrize("kind,dims,xd,ud,H,B", DMMA_CASES)
def test_float64_tensor_core_path_vs_oracle(kind, dims, xd, ud, H, B):
    """nempc_dmma_net_kernel (FP64 tensor cores, mma.sync.m8n8k4.f64) + nempc_dmma_stage_kernel: the 1e-10 parity mode of the wide tanh networks;
    all three output sets (residual only / + Jacobian / + Hessian run different row stacks), tiles that end inside a problem, the generic
    kernel as a second opinion"""
    import torch
    mlp, obj, Z, X0, lam, sig = _problem(dims, xd, ud, H, B)
    ref = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).evaluate(Z, X0, lam, sig)
    ev = _evaluator(mlp, kind, H, "float64", "tc", obj)
    assert "nempc_dmma_net_kernel" in ev.kernel_name and "DMMA" in ev.kernel_name
    np.testing.assert_array_equal(ev.hes_rows, BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).hes_rows)
    got = _run(ev, Z, X0, lam, sig)
    for kr, kg in KEYS:
        assert _relerr(got[kg], ref[kr]) < TOL64, (kg, _relerr(got[kg], ref[kr]))
    t = lambda a: torch.as_tensor(a).cuda()
    o1 = ev.eval(t(Z), t(X0), want=("resid", "jac"))
    o0 = ev.eval(t(Z), t(X0), want=("resid",))
    torch.cuda.synchronize()
    assert _relerr(o1["jac"].cpu().numpy(), ref["jac_vals"]) < TOL64 and _relerr(o1["resid"].cpu().numpy(), ref["resid"]) < TOL64
    assert _relerr(o0["resid"].cpu().numpy(), ref["resid"]) < TOL64
    gen = _run(_evaluator(mlp, kind, H, "float64", "generic", obj), Z, X0, lam, sig)
    for _, kg in KEYS:
        assert _relerr(got[kg], gen[kg]) < TOL64, kg
    ev.close()


def test_float64_tensor_core_path_dispatch_and_chunks():
    """AUTO sends float64 wide networks to the DMMA path from 512 horizon steps on (generic below); a batch above one scratch chunk
    (2^18 steps) is evaluated chunk by chunk -- problems on both sides of the chunk boundary against the oracle; unsupported shapes are
    refused under kernel='tc' and served by the generic kernel under 'auto'"""
    import torch
    from pyneuralempc_b200._lib import NempcError
    H, B = 100, 2700                                             # 270 000 steps > 262 144
    mlp, obj, _, _, _, _ = _problem([3, 64, 64, 2], 2, 1, H, 1)
    rng = np.random.default_rng(5)
    Z, X0, lam = rng.uniform(-1, 1, (B, H * 3)), rng.uniform(-1, 1, (B, 2)), rng.standard_normal((B, H * 2))
    ev = _evaluator(mlp, "rk4", H, "float64", "auto", obj)
    assert "nempc_dmma_net_kernel" in ev.kernel_name
    t = lambda a: torch.as_tensor(a).cuda()
    out = ev.eval(t(Z), t(X0), t(lam), 1.0)
    torch.cuda.synchronize()
    assert ev.launch_count == 2 * 9 + 1                          # two chunks x (init + 4 network + 4 stage kernels) + objective
    pick = np.array([0, 1, 2620, 2621, 2622, 2699])              # 2621 * 100 = 262 100: problem 2621 straddles the chunk boundary
    ref = BlockEvaluator(mlp, "rk4", H, DT=0.1, objective=obj).evaluate(Z[pick], X0[pick], lam[pick], 1.0)
    for kr, kg in KEYS:
        assert _relerr(out[kg][torch.as_tensor(pick).cuda()].cpu().numpy(), ref[kr]) < TOL64, kg
    n0 = ev.launch_count
    small = ev.eval(t(Z[:2]), t(X0[:2]), t(lam[:2]), 1.0)        # 200 steps: generic kernel (one launch + objective)
    torch.cuda.synchronize()
    assert ev.launch_count - n0 == 2
    for kr, kg in KEYS:
        assert _relerr(small[kg].cpu().numpy(), ref[kr][:2]) < TOL64, kg
    ev.close()
    m2 = MLP.glorot([3, 128, 96, 2], 2, 1, seed=2)
    with pytest.raises(NempcError):
        _evaluator(m2, "rk4", 5, "float64", "tc")
    assert "generic" in _evaluator(m2, "rk4", 5, "float64", "auto").kernel_name


@pytest.mark.parametrize("compute,kind,dims,xd,ud,H,B,tag", [("float64", "rk4", [3, 64, 64, 2], 2, 1, 8, 1100, "nempc_dmma_net_kernel"),
                                                           ("float32", "discrete", [16, 256, 256, 12], 12, 4, 3, 1100, "nempc_wide_kernel"),
                                                           ("float32", "rk4", [5, 256, 256, 4], 4, 1, 4, 1300, "nempc_wide_kernel")])
def test_host_chunk_pipeline_of_kernels_with_handle_scratch(compute, kind, dims, xd, ud, H, B, tag):
    """nempc_eval_host cuts B >= 512 problems into chunks on three streams (H2D | kernels | D2H overlap).  The width-256 kernel and the
    float64 tensor-core path keep scratch in the handle, so the kernels of consecutive chunks are chained with an event: the host call
    must return the bits of the one-launch device call -- directly issued (first call), and replayed as the captured graph (third call)."""
    mlp, obj, Z, X0, lam, sig = _problem(dims, xd, ud, H, B)
    ev = _evaluator(mlp, kind, H, compute, "auto", obj)
    assert tag in ev.kernel_name
    got = _run(ev, Z, X0, lam, sig)
    for rep in range(3):
        host = ev.eval_host(Z, X0, lam, sig)
        for _, kg in KEYS:
            np.testing.assert_array_equal(host[kg], got[kg], err_msg=f"{kg} (call {rep})")
    pick = np.array([0, 274, 275, 549, 550, B - 1])                       # both sides of the chunk boundaries (chunks of 275 / 325 problems)
    ref = BlockEvaluator(mlp, kind, H, DT=0.1, objective=obj).evaluate(Z[pick], X0[pick], lam[pick], sig[pick])
    tol = TOL64 if compute == "float64" else TOL32
    for kr, kg in KEYS:
        assert _relerr(host[kg][pick], ref[kr]) < tol, kg
    ev.close()
