"""Generate the golden fixtures under ``tests/golden/`` by running the UNMODIFIED reference
(``/root/reference/pyNeuralEMPC`` imported through ``oracle.shim``) on seeded inputs.

Run once in the build container:  ``python tests/golden/make_golden.py``.
The GPU box has no ``/root/reference``; tests there only read the committed ``.npz`` files.

What is recorded (all float64 unless noted):
  lv_mlp_weights.npz      weights of examples/lotka_volterra/nn_model.h5 (float32, by offset)
  ref_<kind>_H6.npz       full dense integrator outputs (forward, jacobian (m,n), hessian (m,n,n),
                          hessianstructure (n,n)) + IpoptProblem callbacks, LV fixture net, H=6
  ref_<kind>_H25.npz      IpoptProblem callbacks at the shipped problem size (C1: H=25)
  ref_discrete_d5_H5.npz  x_dim=4,u_dim=1 network through Discret/Unity (RK4 of the reference
                          crashes unless x_dim+u_dim == 3 -- recorded as ``rk4_d5_error``)
  ref_rk4_f32_H6.npz      RK4 with the network evaluated in float32 (TensorFlow numerics mimic)
  ref_constraints_H6.npz  IpoptProblem callbacks with one extra InequalityConstraint that has a Hessian (ipopt.py:49-50, 75-80, 93-94)
  ref_exo_H6.npz          tvp / p inputs: network over [x, u, tvp, p] (2+1+2+1 -> 8 -> 8 -> 2) through the reference's Discret /
                          Unity / RK4 integrators called with ``p=, tvp=`` (integrator/rk4.py:69-72, discret.py:27,48,64)
  ref_rk4_w128_H6.npz     3 -> 128 -> 128 -> 2 network (float32 weights, regenerated from the recorded seed) through the reference's RK4
                          integrator + IpoptProblem: pins the tcgen05 kernel of nempc_tc.cuh (TcCfg<2,1,2,...,128>) to the reference
  ref_discrete_w256_H6.npz  5 -> 256 -> 256 -> 256 -> 4 network through the reference's DiscretIntegrator + IpoptProblem: pins the
                          width-256 kernel of nempc_wide.cuh to the reference (its RK4 Hessian only runs for x_dim + u_dim == 3)
  ref_rk4_w256_H6.npz     3 -> 256 -> 256 -> 2 network through the reference's RK4 integrator (two-sweep stage schedule of nempc_wide.cuh)
  ref_closed_loop_c1.npz  BASELINE config C1 as named (examples/lotka_volterra/run.py:38-87): the reference's own NMPC.next with its Slsqp
                          optimizer, H = 25, cost 1.1 * sum(u), u in [-1, 0.2], x_0 <= 1, LV fixture network, discrete and RK4 (DT 0.1)
                          integrators: iteration count, final cost, solution
  ref_rolling_<kind>_w<w>.npz  rolling-window (NARX) models, SURVEY 8f rank 2: a window network (w (x+u) -> 10 -> 10 -> x) behind a subclass of
                          the reference's Model with the layouts of KerasTFModelRollingInput (oracle/rolling_np.py restates
                          model/tensorflow.py:112-340), through the reference's Discret / Unity integrators and IpoptProblem
  ref_receding_horizon.npz  six receding-horizon steps of the reference's NMPC + warm-started Slsqp (unity transcription, H = 15, tracking cost)
  ref_quadform_H6.npz     IpoptProblem callbacks with a NON-SEPARABLE quadratic cost (full stage / terminal weights, control-rate penalty,
                          state-control cross term: oracle.objectives_np.QuadraticFormObjective behind the reference's ObjectiveFunc):
                          the Hessian pattern is the union np.nonzero(np.tril(objective_map + integrator_map)) of ipopt.py:55-62
The dynamics model handed to the reference is ``oracle.mlp_np.MLP`` wrapped in a subclass of the
reference's ``Model`` (TensorFlow is not installed), so these files pin the integrator / IPOPT
glue of the oracle, not TensorFlow's autodiff.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import shim  # noqa: E402
from oracle.mlp_np import MLP, ExoMLP, read_lv_fixture_h5  # noqa: E402
from oracle.objectives_np import SeparableQuadraticObjective  # noqa: E402

DT_RK4 = 0.1  # examples/lotka_volterra/run.py:77


def ref_objective(ref, sep):
    class Obj(ref.objective.base.ObjectiveFunc):
        def forward(self, states, u, p=None, tvp=None):
            return sep.forward(states, u)

        def gradient(self, states, u, p=None, tvp=None):
            return sep.gradient(states, u)

        def hessian(self, states, u, p=None, tvp=None):
            return sep.hessian(states, u)

        def hessianstructure(self, H, model):
            return sep.hessianstructure(H, model)

    return Obj()


def make_integrator(ref, kind, model, H):
    if kind == "discrete":
        return ref.integrator.discret.DiscretIntegrator(model, H)
    if kind == "unity":
        return ref.integrator.unity.UnityIntegrator(model, H)
    return ref.integrator.rk4.RK4Integrator(model, H, DT_RK4)


def record(ref, mlp, kind, H, seed, objective_kind, full_dense):
    rng = np.random.default_rng(seed)
    xd, ud = mlp.x_dim, mlp.u_dim
    n, m = H * (xd + ud), H * xd
    z = rng.uniform(-1, 1, size=n)
    x0 = rng.uniform(-1, 1, size=xd)
    lam = rng.standard_normal(m)
    sigma = 0.7
    if objective_kind == "linear":
        sep = SeparableQuadraticObjective.linear_in_u(H, xd, ud, 1.1)          # run.py:83-87
    elif objective_kind == "setpoint":
        sep = SeparableQuadraticObjective.control_setpoint(H, xd, ud, 2.0)     # test.py:59-60
    else:
        sep = SeparableQuadraticObjective.tracking(H, xd, ud, rng.uniform(0.5, 2, xd), rng.uniform(0.1, 1, ud),
                                                   x_ref=rng.uniform(-1, 1, (H, xd)))
    model = shim.make_reference_model(mlp)
    integ = make_integrator(ref, kind, model, H)
    np.random.seed(seed)                                   # integrator/base.py:96 uses the global RNG
    pb = ref.optimizer.ipopt.IpoptProblem(x0, ref_objective(ref, sep), [], integ, use_hessian=True)
    out = dict(kind=kind, H=H, x_dim=xd, u_dim=ud, DT=DT_RK4, z=z, x0=x0, lam=lam, sigma=sigma,
               obj_lin=sep.lin, obj_quad=sep.quad, obj_ref=sep.ref, net_dtype=str(mlp.dtype))
    out["objective"] = pb.objective(z)
    out["gradient"] = pb.gradient(z)
    out["constraints"] = pb.constraints(z)
    out["jacobian"] = pb.jacobian(z)
    r, c = pb.hessianstructure()
    out["hes_rows"], out["hes_cols"] = r, c
    out["hessian_values"] = pb.hessian(z, lam, sigma)
    out["integrator_structure"] = integ.hessianstructure()
    if full_dense:
        states, u = z[:H * xd].reshape(H, xd), z[H * xd:].reshape(H, ud)
        out["integrator_forward"] = integ.forward(states, u, x0)
        out["integrator_jacobian"] = integ.jacobian(states, u, x0)
        out["integrator_hessian"] = integ.hessian(states, u, x0)
    return out


def record_exo(ref, H=6, seed=500):
    """the reference integrators with time-varying (tvp) and constant (p) model inputs; the model adapter gathers
    [x, u, tvp, p] and drops the tvp / p derivative columns like KerasTFModel (model/tensorflow.py:39-47, 65-66, 97-98)."""
    rng = np.random.default_rng(seed)
    xd, ud, td, pd = 2, 1, 2, 1
    full = MLP.glorot([xd + ud + td + pd, 8, 8, xd], xd, ud + td + pd, seed=11)
    exo = ExoMLP(full.weights, xd, ud, td, pd)

    class ExoBackedModel(ref.model.base.Model):
        def __init__(self):
            super().__init__(xd, ud, pd, td)

        def forward(self, x, u, p=None, tvp=None):
            return exo.bind(tvp, p).forward(x, u)

        def jacobian(self, x, u, p=None, tvp=None):
            return exo.bind(tvp, p).dense_jacobian(x, u)

        def hessian(self, x, u, p=None, tvp=None):
            return exo.bind(tvp, p).dense_hessian(x, u)

    states, u, x0 = rng.uniform(-1, 1, (H, xd)), rng.uniform(-1, 1, (H, ud)), rng.uniform(-1, 1, xd)
    tvp, p = rng.uniform(-1, 1, (H, td)), rng.uniform(-1, 1, pd)
    out = dict(H=H, x_dim=xd, u_dim=ud, tvp_dim=td, p_dim=pd, DT=DT_RK4, states=states, u=u, x0=x0, tvp=tvp, p=p,
               lam=rng.standard_normal(H * xd))
    out.update({f"net_W{i}": W for i, (W, _) in enumerate(full.weights)})
    out.update({f"net_b{i}": b for i, (_, b) in enumerate(full.weights)})
    for kind in ("discrete", "unity", "rk4"):
        integ = make_integrator(ref, kind, ExoBackedModel(), H)
        out[f"{kind}_forward"] = integ.forward(states, u, x0, p=p, tvp=tvp)
        out[f"{kind}_jacobian"] = integ.jacobian(states, u, x0, p=p, tvp=tvp)
        out[f"{kind}_hessian"] = integ.hessian(states, u, x0, p=p, tvp=tvp)
    return out


def circle_constraint(base, H, xd, ud):
    """c_t = x_t[0]^2 + u_t[0]^2 >= 0 rows (an InequalityConstraint with a Hessian), used by record_constraints and by the GPU test"""
    n = H * (xd + ud)

    class Circle(base):
        def forward(self, x, u, p=None, tvp=None):
            return x[:, 0] ** 2 + u[:, 0] ** 2

        def jacobian(self, x, u, p=None, tvp=None):
            J = np.zeros((H, n))
            for t in range(H):
                J[t, t * xd] = 2.0 * x[t, 0]
                J[t, H * xd + t * ud] = 2.0 * u[t, 0]
            return J

        def hessian(self, x, u, p=None, tvp=None):
            Hc = np.zeros((H, n, n))
            for t in range(H):
                Hc[t, t * xd, t * xd] = 2.0
                Hc[t, H * xd + t * ud, H * xd + t * ud] = 2.0
            return Hc

        def get_dim(self, H_=None):
            return H

    return Circle()


def record_constraints(ref, lv, H=6, seed=600):
    """IpoptProblem with one extra user constraint (optimizer/ipopt.py:49-50, 75-80, 93-94)"""
    rng = np.random.default_rng(seed)
    xd, ud = lv.x_dim, lv.u_dim
    n, m = H * (xd + ud), H * xd
    sep = SeparableQuadraticObjective.tracking(H, xd, ud, rng.uniform(0.5, 2, xd), rng.uniform(0.1, 1, ud),
                                               x_ref=rng.uniform(-1, 1, (H, xd)))
    ctr = circle_constraint(ref.constraints.InequalityConstraint, H, xd, ud)
    integ = make_integrator(ref, "rk4", shim.make_reference_model(lv), H)
    z, x0, lam, sigma = rng.uniform(-1, 1, n), rng.uniform(-1, 1, xd), rng.standard_normal(m + H), 0.6
    np.random.seed(seed)
    pb = ref.optimizer.ipopt.IpoptProblem(x0, ref_objective(ref, sep), [ctr], integ, use_hessian=True)
    r, c = pb.hessianstructure()
    return dict(H=H, DT=DT_RK4, z=z, x0=x0, lam=lam, sigma=sigma, obj_lin=sep.lin, obj_quad=sep.quad, obj_ref=sep.ref,
                constraints=pb.constraints(z), jacobian=pb.jacobian(z), hes_rows=r, hes_cols=c,
                hessian_values=pb.hessian(z, lam, sigma), cl=pb.get_constraint_lower_bounds(), cu=pb.get_constraint_upper_bounds())


def record_wide(ref, dims, xd, ud, kind, H, seed, net_seed):
    """a wide network (weights NOT stored: MLP.glorot(dims, seed=net_seed, float32) regenerates them, `weights_checksum` guards it)
    through the reference's integrator + IpoptProblem"""
    mlp = MLP.glorot(dims, xd, ud, seed=net_seed, dtype=np.float32)
    out = record(ref, mlp.astype(np.float64), kind, H, seed, "tracking", False)
    out.update(dims=np.array(dims), net_seed=net_seed,
               weights_checksum=float(sum(np.abs(np.asarray(W, np.float64)).sum() + np.abs(np.asarray(b, np.float64)).sum() for W, b in mlp.weights)))
    return out


def record_closed_loop_c1(ref, lv, H=25):
    """the reference's NMPC.next + Slsqp on C1 as the shipped script names it (run.py:38-54, 72-87)"""
    import scipy
    from scipy.optimize import minimize as sp_minimize
    out = dict(H=H, DT=DT_RK4, x0=np.array([0.66, -0.9]), scipy_version=scipy.__version__)
    sep = SeparableQuadraticObjective.linear_in_u(H, 2, 1, 1.1)                        # run.py:80-87: sum(u * 1.1)
    dom = ref.constraints.DomainConstraint(states_constraint=[[-np.inf, 1.0], [-np.inf, np.inf]], control_constraint=[[-1.0, 0.2]])
    seen = []

    def spy(*a, **kw):
        r = sp_minimize(*a, **kw)
        seen.append(r)
        return r

    ref.optimizer.slsqp.minimize = spy
    try:
        for kind in ("discrete", "rk4", "unity"):           # unity: the well-posed transcription for this next-state network
            del seen[:]
            integ = make_integrator(ref, kind, shim.make_reference_model(lv), H)
            opt = ref.optimizer.slsqp.Slsqp(verbose=0)
            mpc = ref.controller.NMPC(integ, ref_objective(ref, sep), [dom], H, DT_RK4, optimizer=opt)
            xs, us = mpc.next(out["x0"])
            r = seen[-1]
            out.update({f"{kind}_nit": r.nit, f"{kind}_fun": r.fun, f"{kind}_success": bool(r.success), f"{kind}_minimize_calls": len(seen),
                        f"{kind}_returned_none": xs is None, f"{kind}_z": np.asarray(r.x),
                        f"{kind}_all_nit": np.array([q.nit for q in seen]), f"{kind}_all_fun": np.array([q.fun for q in seen]),
                        f"{kind}_all_status": np.array([q.status for q in seen]),
                        f"{kind}_eq_violation": float(np.abs(integ.forward(r.x[:2 * H].reshape(H, 2), r.x[2 * H:].reshape(H, 1), out["x0"])).max())})
            if xs is not None:
                out.update({f"{kind}_x": np.asarray(xs), f"{kind}_u": np.asarray(us)})
    finally:
        ref.optimizer.slsqp.minimize = sp_minimize
    return out


def record_rolling(ref, kind, w, forward_rolling, H=6, seed=800):
    """reference integrator + IpoptProblem over a rolling-window model (banded model Jacobian / Hessian)"""
    from oracle.rolling_np import RollingMLP
    rng = np.random.default_rng(seed + w)
    xd, ud = 2, 1
    net = MLP.glorot([w * (xd + ud), 10, 10, xd], xd, w * (xd + ud) - xd, seed=30 + w)
    roll = RollingMLP(net.weights, xd, ud, w, forward_rolling)

    class RollingBackedModel(ref.model.base.Model):
        def __init__(self):
            super().__init__(xd, ud, 0, 0)

        def forward(self, x, u, p=None, tvp=None):
            return roll.forward(x, u)

        def jacobian(self, x, u, p=None, tvp=None):
            return roll.jacobian(x, u)

        def hessian(self, x, u, p=None, tvp=None):
            return roll.hessian(x, u)

    x_prev, u_prev = rng.uniform(-1, 1, (w - 1, xd)), rng.uniform(-1, 1, (w - 1, ud))
    roll.set_prev_data(x_prev, u_prev)
    n, m = H * (xd + ud), H * xd
    z, x0, lam, sigma = rng.uniform(-1, 1, n), rng.uniform(-1, 1, xd), rng.standard_normal(m), 0.8
    sep = SeparableQuadraticObjective.tracking(H, xd, ud, rng.uniform(0.5, 2, xd), rng.uniform(0.1, 1, ud), x_ref=rng.uniform(-1, 1, (H, xd)))
    integ = make_integrator(ref, kind, RollingBackedModel(), H)
    np.random.seed(seed)
    pb = ref.optimizer.ipopt.IpoptProblem(x0, ref_objective(ref, sep), [], integ, use_hessian=True)
    out = dict(kind=kind, H=H, x_dim=xd, u_dim=ud, rolling_window=w, forward_rolling=forward_rolling, x_prev=x_prev, u_prev=u_prev,
               z=z, x0=x0, lam=lam, sigma=sigma, obj_lin=sep.lin, obj_quad=sep.quad, obj_ref=sep.ref)
    out.update({f"net_W{i}": W for i, (W, _) in enumerate(net.weights)})
    out.update({f"net_b{i}": b for i, (_, b) in enumerate(net.weights)})
    out["objective"], out["gradient"], out["constraints"], out["jacobian"] = pb.objective(z), pb.gradient(z), pb.constraints(z), pb.jacobian(z)
    out["hes_rows"], out["hes_cols"] = pb.hessianstructure()
    out["hessian_values"] = pb.hessian(z, lam, sigma)
    states, u = z[:H * xd].reshape(H, xd), z[H * xd:].reshape(H, ud)
    out["integrator_hessian"] = integ.hessian(states, u, x0)
    xp = np.concatenate([x0[None], states])[:-1]
    out["model_forward"], out["model_jacobian"], out["model_hessian"] = roll.forward(xp, u), roll.jacobian(xp, u), roll.hessian(xp, u)
    return out


def quadform_case(H, xd, ud, seed=900):
    """the non-separable cost of ref_quadform_H6.npz (also rebuilt by the GPU test from the stored blocks)"""
    rng = np.random.default_rng(seed)
    A = lambda r, c: rng.uniform(-1, 1, (r, c))
    spd = lambda k: (lambda M: M @ M.T + 0.1 * np.eye(k))(A(k, k))
    return dict(Q=spd(xd), R=spd(ud), Qf=spd(xd), S=spd(ud), N=0.2 * A(xd, ud), x_ref=rng.uniform(-1, 1, (H, xd)), u_ref=rng.uniform(-1, 1, ud),
                lin=0.3 * rng.uniform(-1, 1, H * (xd + ud)))


def record_quadform(ref, lv, H=6, seed=900):
    from oracle.objectives_np import QuadraticFormObjective
    rng = np.random.default_rng(seed + 1)
    xd, ud = lv.x_dim, lv.u_dim
    n, m = H * (xd + ud), H * xd
    blocks = quadform_case(H, xd, ud, seed)
    qf = QuadraticFormObjective(H, xd, ud, **blocks)
    integ = make_integrator(ref, "rk4", shim.make_reference_model(lv), H)
    z, x0, lam, sigma = rng.uniform(-1, 1, n), rng.uniform(-1, 1, xd), rng.standard_normal(m), 0.65
    np.random.seed(seed)
    pb = ref.optimizer.ipopt.IpoptProblem(x0, ref_objective(ref, qf), [], integ, use_hessian=True)
    r, c = pb.hessianstructure()
    out = dict(H=H, DT=DT_RK4, z=z, x0=x0, lam=lam, sigma=sigma, objective=pb.objective(z), gradient=pb.gradient(z),
               constraints=pb.constraints(z), jacobian=pb.jacobian(z), hes_rows=r, hes_cols=c, hessian_values=pb.hessian(z, lam, sigma))
    out.update({f"cost_{k}": v for k, v in blocks.items()})
    return out


def record_receding_horizon(ref, lv, H=15, nsteps=6):
    """receding-horizon run of the reference's NMPC (unity transcription, Slsqp warm-started from the shifted previous solution,
    optimizer/slsqp.py:155-160): the plant is the network itself, x_{k+1} = first predicted state; per step: iterations, cost, control"""
    from scipy.optimize import minimize as sp_minimize
    sep = SeparableQuadraticObjective.tracking(H, 2, 1, [1.0, 1.0], [0.1], x_ref=np.array([0.5, -0.7]))
    dom = ref.constraints.DomainConstraint(states_constraint=[[-np.inf, 1.0], [-np.inf, np.inf]], control_constraint=[[-1.0, 0.2]])
    seen = []

    def spy(*a, **kw):
        r = sp_minimize(*a, **kw)
        seen.append(r)
        return r

    ref.optimizer.slsqp.minimize = spy
    try:
        integ = make_integrator(ref, "unity", shim.make_reference_model(lv), H)
        opt = ref.optimizer.slsqp.Slsqp(verbose=0, init_with_last_result=True)
        mpc = ref.controller.NMPC(integ, ref_objective(ref, sep), [dom], H, DT_RK4, optimizer=opt)
        x = np.array([0.66, -0.9])
        xs, us, nits, funs = [x.copy()], [], [], []
        for _ in range(nsteps):
            pred, u = mpc.next(x)
            nits.append(seen[-1].nit); funs.append(seen[-1].fun); us.append(u[0].copy())
            x = np.asarray(pred[0], np.float64)
            xs.append(x.copy())
    finally:
        ref.optimizer.slsqp.minimize = sp_minimize
    return dict(H=H, nsteps=nsteps, obj_lin=sep.lin, obj_quad=sep.quad, obj_ref=sep.ref, x_traj=np.array(xs), u_traj=np.array(us),
                nit=np.array(nits), fun=np.array(funs), minimize_calls=len(seen))


def main():
    ref = shim.load_reference()
    h5 = os.path.join(shim.REFERENCE_ROOT, "examples", "lotka_volterra", "nn_model.h5")
    weights = read_lv_fixture_h5(h5)
    np.savez(os.path.join(HERE, "lv_mlp_weights.npz"),
             **{f"W{i}": W for i, (W, _) in enumerate(weights)},
             **{f"b{i}": b for i, (_, b) in enumerate(weights)})
    lv = MLP(weights, 2, 1, dtype=np.float64)
    objective_of = {"discrete": "setpoint", "unity": "tracking", "rk4": "linear"}
    for i, kind in enumerate(("discrete", "unity", "rk4")):
        np.savez_compressed(os.path.join(HERE, f"ref_{kind}_H6.npz"),
                            **record(ref, lv, kind, 6, 100 + i, "tracking", True))
        np.savez_compressed(os.path.join(HERE, f"ref_{kind}_H25.npz"),
                            **record(ref, lv, kind, 25, 200 + i, objective_of[kind], False))
    # float32 network arithmetic (what Keras / tf.hessians do, model/tensorflow.py:51,91)
    np.savez_compressed(os.path.join(HERE, "ref_rk4_f32_H6.npz"),
                        **record(ref, lv.astype(np.float32), "rk4", 6, 300, "linear", True))
    # a d = 5 network: Discret / Unity work, the reference RK4 Hessian does not (rk4.py:246)
    net5 = MLP.glorot([5, 8, 8, 4], 4, 1, seed=7)
    out = record(ref, net5, "discrete", 5, 400, "tracking", True)
    out.update({f"net_W{i}": W for i, (W, _) in enumerate(net5.weights)})
    out.update({f"net_b{i}": b for i, (_, b) in enumerate(net5.weights)})
    uni = record(ref, net5, "unity", 5, 400, "tracking", True)
    out.update({f"unity_{k}": v for k, v in uni.items() if k.startswith(("integrator_", "hessian_values", "jacobian", "constraints"))})
    try:
        model = shim.make_reference_model(net5)
        integ = ref.integrator.rk4.RK4Integrator(model, 5, DT_RK4)
        rng = np.random.default_rng(1)
        integ.hessian(rng.uniform(size=(5, 4)), rng.uniform(size=(5, 1)), rng.uniform(size=4))
        out["rk4_d5_error"] = "none"
    except Exception as e:  # noqa: BLE001 -- the reference's own failure is the datum
        out["rk4_d5_error"] = f"{type(e).__name__}: {e}"
    # ... but its RK4 forward/jacobian are dimension-generic: record them for d = 5
    rng = np.random.default_rng(401)
    xs, us, x0 = rng.uniform(-1, 1, (5, 4)), rng.uniform(-1, 1, (5, 1)), rng.uniform(-1, 1, 4)
    integ = ref.integrator.rk4.RK4Integrator(shim.make_reference_model(net5), 5, DT_RK4)
    out.update(rk4_x=xs, rk4_u=us, rk4_x0=x0, rk4_forward=integ.forward(xs, us, x0),
               rk4_jacobian=integ.jacobian(xs, us, x0))
    np.savez_compressed(os.path.join(HERE, "ref_discrete_d5_H5.npz"), **out)
    np.savez_compressed(os.path.join(HERE, "ref_exo_H6.npz"), **record_exo(ref))
    np.savez_compressed(os.path.join(HERE, "ref_constraints_H6.npz"), **record_constraints(ref, lv))
    np.savez_compressed(os.path.join(HERE, "ref_rk4_w128_H6.npz"), **record_wide(ref, [3, 128, 128, 2], 2, 1, "rk4", 6, 700, 21))
    np.savez_compressed(os.path.join(HERE, "ref_discrete_w256_H6.npz"), **record_wide(ref, [5, 256, 256, 256, 4], 4, 1, "discrete", 6, 710, 22))
    np.savez_compressed(os.path.join(HERE, "ref_rk4_w256_H6.npz"), **record_wide(ref, [3, 256, 256, 2], 2, 1, "rk4", 6, 720, 23))
    np.savez_compressed(os.path.join(HERE, "ref_closed_loop_c1.npz"), **record_closed_loop_c1(ref, lv))
    np.savez_compressed(os.path.join(HERE, "ref_receding_horizon.npz"), **record_receding_horizon(ref, lv))
    np.savez_compressed(os.path.join(HERE, "ref_quadform_H6.npz"), **record_quadform(ref, lv))
    np.savez_compressed(os.path.join(HERE, "ref_rolling_discrete_w2.npz"), **record_rolling(ref, "discrete", 2, True))
    np.savez_compressed(os.path.join(HERE, "ref_rolling_unity_w3.npz"), **record_rolling(ref, "unity", 3, False))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
