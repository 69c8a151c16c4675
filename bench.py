#!/usr/bin/env python
"""Benchmark of the NLP-evaluation hot path (BASELINE.json metric: Jacobian+Hessian NLP evals/s in horizon-steps/s).

One "step" of this bench = ONE NLP evaluation of the whole batch at one iterate: constraint residual + sparse
Jacobian values + lambda-contracted sparse Lagrangian-Hessian values + objective value/gradient for B independent
problems (2 kernel launches).  Workload at N=1 is BASELINE.json configs[1] ("C2"): Lotka-Volterra-shaped MLP
3->30->30->2 (tanh), RK4 integrator DT=0.1, horizon 50, batch 4096 problems PER GPU (weak scaling, no collective on
the data path -- problems are independent).

  python bench.py [--gpus N --steps K --warmup W]            # this framework (CUDA kernels)
  python bench.py --impl reference [...]                     # the reference CPU algorithm (oracle port) on host cores
  torchrun --nproc-per-node N bench.py --gpus N ...          # one rank per GPU

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = same metric through the host-buffer
plug-in call (pinned host -> device -> host inside the timed region); `roofline` = dominant kernel vs the measured
FP32 FMA peak (the path is compute bound: ~560 flop/byte); `cpu_baseline` = the dense reference-literal port timed
on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "jac_hess_nlp_eval_horizon_steps_per_s"
UNIT = "horizon-steps/s"

WORKLOADS = {
    # name: (layer dims, x, u, integrator, DT, H, B per GPU)
    "C2": dict(dims=[3, 30, 30, 2], x=2, u=1, integ="rk4", DT=0.1, H=50, B=4096,
               desc="Lotka-Volterra MLP 3-30-30-2 tanh, RK4 DT=0.1, H=50, B=4096 problems/GPU (BASELINE configs[1])"),
    "C1": dict(dims=[3, 30, 30, 2], x=2, u=1, integ="discrete", DT=None, H=25, B=1,
               desc="examples/lotka_volterra: discrete integrator, H=25, single problem (BASELINE configs[0])"),
    "C3": dict(dims=[5, 128, 128, 128, 4], x=4, u=1, integ="rk4", DT=0.1, H=100, B=16384,
               desc="cart-pole MLP 5-128-128-128-4 tanh, RK4, H=100, B=16384 (BASELINE configs[2])"),
    "C4": dict(dims=[16, 256, 256, 256, 256, 12], x=12, u=4, integ="discrete", DT=None, H=200, B=65536,
               desc="quadrotor MLP 16-256x4-12, discrete, H=200, B=65536 (BASELINE configs[3])"),
    "C4rk4": dict(dims=[16, 256, 256, 256, 256, 12], x=12, u=4, integ="rk4", DT=0.1, H=200, B=65536,
                  desc="quadrotor MLP 16-256x4-12, RK4 DT=0.1, H=200, B=65536 (BASELINE configs[3] names no integrator: the RK4 variant)"),
}


def make_problem(wl, B, seed=1234):
    """synthetic weights (Glorot-uniform, seed 0) and iterates: z,x0 ~ U(-1,1), lambda ~ N(0,1), sigma=1."""
    from oracle.mlp_np import MLP
    from oracle.objectives_np import SeparableQuadraticObjective
    mlp = MLP.glorot(wl["dims"], wl["x"], wl["u"], seed=0, dtype=np.float32)
    H, xd, ud = wl["H"], wl["x"], wl["u"]
    n, m = H * (xd + ud), H * xd
    rng = np.random.default_rng(seed)
    Z = rng.uniform(-1, 1, (B, n))
    X0 = rng.uniform(-1, 1, (B, xd))
    lam = rng.standard_normal((B, m))
    obj = SeparableQuadraticObjective.tracking(H, xd, ud, np.linspace(1.0, 2.0, xd), np.linspace(0.1, 0.2, ud))
    return mlp, obj, Z, X0, lam


# ------------------------------------------------------------------------------------------------------
# CPU baseline: the dense reference-literal port (oracle/dense_ref.py) on host cores
# ------------------------------------------------------------------------------------------------------
def reference_usable(wl):
    """the reference's own classes can run this workload on the CPU: its package (sources in the build container, bytecode from
    oracle/_ref on the GPU box) is importable and its RK4 Hessian supports the shape (x_dim + u_dim == 3 only: integrator/rk4.py:246)."""
    from oracle import shim
    if not shim.reference_available() or not (wl["integ"] != "rk4" or wl["x"] + wl["u"] == 3):
        return False
    try:                                                   # a broken / partial copy must not cost the CPU leg: fall back to the port
        ref = shim.load_reference()
        return hasattr(ref.optimizer.ipopt, "IpoptProblem") and hasattr(ref.integrator.rk4, "RK4Integrator")
    except Exception:                                      # noqa: BLE001
        return False


def _reference_problem(wl):
    """the UNMODIFIED reference's integrator + IpoptProblem factory for this workload (network evaluated by oracle.mlp_np: TensorFlow is
    not installable); cached per process"""
    from oracle import shim
    from oracle.mlp_np import MLP
    from oracle.objectives_np import SeparableQuadraticObjective
    ref = shim.load_reference()
    mlp = MLP.glorot(wl["dims"], wl["x"], wl["u"], seed=0, dtype=np.float32)
    sep = SeparableQuadraticObjective.tracking(wl["H"], wl["x"], wl["u"], np.linspace(1.0, 2.0, wl["x"]), np.linspace(0.1, 0.2, wl["u"]))

    class Obj(ref.objective.base.ObjectiveFunc):
        def forward(self, states, u, p=None, tvp=None): return sep.forward(states, u)
        def gradient(self, states, u, p=None, tvp=None): return sep.gradient(states, u)
        def hessian(self, states, u, p=None, tvp=None): return sep.hessian(states, u)
        def hessianstructure(self, H, model): return sep.hessianstructure(H, model)

    model = shim.make_reference_model(mlp)
    if wl["integ"] == "rk4":
        integ = ref.integrator.rk4.RK4Integrator(model, wl["H"], wl["DT"])
    elif wl["integ"] == "unity":
        integ = ref.integrator.unity.UnityIntegrator(model, wl["H"])
    else:
        integ = ref.integrator.discret.DiscretIntegrator(model, wl["H"])
    integ.hessianstructure()                # structure probing is a one-off in the reference (cached on the integrator)
    return lambda x0: ref.optimizer.ipopt.IpoptProblem(x0, Obj(), [], integ, use_hessian=True)


def _cpu_problem_eval(args):
    """one problem, one iterate: objective, gradient, constraints, dense Jacobian, Lagrangian Hessian -- the reference's
    optimizer/ipopt.py callbacks over its dense O(H^3) integrator arrays (use_ref: the reference's own classes; else the literal port)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    wl, z, x0, lam, use_ref = args
    cache = _cpu_problem_eval.__dict__.setdefault("cache", {})
    key = json.dumps(wl, sort_keys=True) + str(use_ref)
    if key not in cache:
        if use_ref:
            cache[key] = _reference_problem(wl)
        else:
            from oracle.dense_ref import DenseIntegrator, DenseIpoptProblem
            from oracle.mlp_np import MLP, DenseModelView
            from oracle.objectives_np import SeparableQuadraticObjective
            mlp = MLP.glorot(wl["dims"], wl["x"], wl["u"], seed=0, dtype=np.float32)
            obj = SeparableQuadraticObjective.tracking(wl["H"], wl["x"], wl["u"], np.linspace(1.0, 2.0, wl["x"]), np.linspace(0.1, 0.2, wl["u"]))
            integ = DenseIntegrator(DenseModelView(mlp), wl["H"], "discrete" if wl["integ"] == "discrete" else wl["integ"], DT=wl["DT"])
            integ.hessianstructure()            # structure probing is a one-off in the reference too (cached)
            cache[key] = lambda x0, obj=obj, integ=integ: DenseIpoptProblem(x0, obj, integ)
    pb = cache[key](x0)
    pb.objective(z); pb.gradient(z); pb.constraints(z); pb.jacobian(z)
    return float(np.sum(pb.hessian(z, lam, 1.0)))


def solver_bounds(wl):
    """control bounds of the shipped example (u in [-1, 0.2], examples/lotka_volterra/run.py:72-74 after normalisation),
    states unbounded."""
    H, xd, ud = wl["H"], wl["x"], wl["u"]
    return np.array([-np.inf] * (H * xd) + [-1.0] * (H * ud)), np.array([np.inf] * (H * xd) + [0.2] * (H * ud))


def _cpu_problem_solve(args):
    """one MPC solve the reference way: SciPy SLSQP (the reference's own Slsqp optimizer, optimizer/slsqp.py:172-173,
    default ftol 0.5e-6, maxiter 200) on the dense reference-literal callbacks."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import warnings
    from scipy.optimize import Bounds, minimize
    wl, x0 = args
    from oracle.dense_ref import DenseIntegrator, DenseIpoptProblem
    from oracle.mlp_np import MLP, DenseModelView
    from oracle.objectives_np import SeparableQuadraticObjective
    mlp = MLP.glorot(wl["dims"], wl["x"], wl["u"], seed=0, dtype=np.float32)
    obj = SeparableQuadraticObjective.tracking(wl["H"], wl["x"], wl["u"], np.linspace(1.0, 2.0, wl["x"]), np.linspace(0.1, 0.2, wl["u"]))
    pb = DenseIpoptProblem(x0, obj, DenseIntegrator(DenseModelView(mlp), wl["H"], wl["integ"], DT=wl["DT"]))
    lb, ub = solver_bounds(wl)
    x_init = np.concatenate([np.tile(x0, wl["H"]), np.zeros(wl["H"] * wl["u"])])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = minimize(pb.objective, x_init, method="SLSQP", jac=pb.gradient, bounds=Bounds(lb, ub),
                     constraints=[{"type": "eq", "fun": pb.constraints, "jac": pb.jacobian}], options={"maxiter": 200, "ftol": 0.5e-6})
    return bool(r.success), int(r.nit), float(r.fun)


def cpu_solver_run(wl_name, procs=None):
    import multiprocessing as mp
    wl = {k: v for k, v in WORKLOADS[wl_name].items() if k != "desc"}
    procs = procs or os.cpu_count() or 1
    _, _, _, X0, _ = make_problem(wl, procs)
    with mp.get_context("fork").Pool(procs) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_problem_solve, [(wl, X0[i]) for i in range(procs)], chunksize=1)
        dt = time.perf_counter() - t0
    return dict(value=procs / dt, unit="solves/s", cores=procs, kind="port", converged=sum(r[0] for r in res), iterations_mean=float(np.mean([r[1] for r in res])),
                sample=f"{procs} problems (one per core), SciPy SLSQP = the reference's Slsqp optimizer on the dense reference-literal callbacks")


def cpu_reference_run(wl_name, sample, steps, warmup, procs=None, budget_s=None):
    """times `steps` evaluations of a bounded sample of problems on `procs` worker processes.  With `budget_s` the
    sample size is chosen so that the whole run lasts about that long (the per-problem cost is measured first)."""
    import multiprocessing as mp
    wl = {k: v for k, v in WORKLOADS[wl_name].items() if k != "desc"}
    procs = procs or os.cpu_count() or 1
    use_ref = reference_usable(wl)
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        _, _, Z, X0, lam = make_problem(wl, max(sample, 4 * procs))
        for _ in range(max(1, min(warmup, 3))):
            pool.map(_cpu_problem_eval, [(wl, Z[i], X0[i], lam[i], use_ref) for i in range(procs)], chunksize=1)   # per-process caches
        if budget_s:
            t0 = time.perf_counter()
            pool.map(_cpu_problem_eval, [(wl, Z[i], X0[i], lam[i], use_ref) for i in range(2 * procs)], chunksize=1)
            per_problem = (time.perf_counter() - t0) / (2 * procs)          # wall seconds per problem with all workers busy
            sample = int(max(procs, min(wl["B"], 512, budget_s / max(1, steps) / per_problem)))
            _, _, Z, X0, lam = make_problem(wl, sample)
        jobs = [(wl, Z[i], X0[i], lam[i], use_ref) for i in range(sample)]
        chunk = max(1, sample // (procs * 4))
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_cpu_problem_eval, jobs, chunksize=chunk)
        dt = time.perf_counter() - t0
    ms = dt / steps * 1e3
    how = ("the reference's own integrator + IpoptProblem (unmodified pyNeuralEMPC, bytecode in oracle/_ref; network evaluated by oracle.mlp_np: TensorFlow is absent)"
           if use_ref else "dense reference-literal port (oracle/dense_ref.py: O(H^3) dense Jacobian/Hessian like integrator/rk4.py + optimizer/ipopt.py)")
    return dict(value=sample * wl["H"] / (dt / steps), ms_per_step=ms, cores=procs, kind="reference" if use_ref else "port",
                sample=f"{sample} of {wl['B']} problems per step ({sample * wl['H']} horizon-steps), {how}, float32 network")


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def wait_ready(self, timeout=3.0):
        """block until nvidia-smi has delivered its first sample: its start-up (NVML initialisation takes driver locks) must not fall into
        the timed region, where it delayed kernel launches of rank 0 by ~1 ms in a 5 ms region (profiles/r2s_bench_n2.json)"""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def tensor_roofline(ev, wl, steps_per_eval, k_ms, peaks):
    """roofline object of a tensor-core kernel: `achieved` = ALGORITHMIC flop (SURVEY 8d formula) / kernel time; the denominator is the
    measured dense 16-bit tensor peak DIVIDED BY 3 -- every product is three f16 MMAs (hi*hi + lo*hi + hi*lo) for f32-grade accuracy."""
    tpeak = peaks.get("bf16_tflops", 1590.0)
    flops = ev.flops_per_step * steps_per_eval
    achieved = flops / (k_ms * 1e-3) / 1e12
    d_in = wl["x"] + wl["u"]
    stages = 4 if wl["integ"] == "rk4" else 1
    nmm = len(wl["dims"]) - 3
    hw = wl["dims"][1]
    if "nempc_wide" in ev.kernel_name:       # adjoint form: primal + adjoint + d tangent rows (RK4: two sweeps), 3 products
        rows = (2 + d_in) if stages == 1 else (7 + 4 + 7 * d_in)      # RK4: 7 primal, 4 adjoint and 7 tangent passes (two sweeps over the stages)
        form = "adjoint form: %d GEMM rows per step" % rows
    else:                                    # forward second order: 1 + d + d(d+1)/2 rows per stage
        rows = (1 + d_in + d_in * (d_in + 1) // 2) * stages
        form = "forward second-order rows: %d GEMM rows per step" % rows
    executed = 3 * 2.0 * hw * hw * rows * nmm * steps_per_eval
    return {"bound": "tensor", "kernel": ev.kernel_name, "achieved": achieved, "peak": tpeak / 3.0, "unit": "TFLOP/s",
            "frac": achieved / (tpeak / 3.0), "frac_of_unsplit_peak": achieved / tpeak,
            "peak_source": ("MEASURED_PEAKS.json bf16_tflops" if peaks else "fallback 1590 TFLOP/s") + " / 3 (three f16 MMAs per product: split operands for f32-grade accuracy)",
            "kernel_ms": k_ms, "flops_per_horizon_step": ev.flops_per_step, "bytes_per_horizon_step": ev.bytes_per_step(),
            "executed_mma_tflops": executed / (k_ms * 1e-3) / 1e12, "executed_over_algorithmic": executed / flops,
            "hbm_achieved_gbs": ev.bytes_per_step() * steps_per_eval / (k_ms * 1e-3) / 1e9,
            "note": "tcgen05 kind::f16, split operands; " + form + "; see DESIGN.md 5.4 / 5.8", "traffic": wide_traffic(ev, wl, steps_per_eval)}


def wide_traffic(ev, wl, steps_per_eval):
    """DRAM bytes per launch of the wide kernel from the committed ncu capture (bytes per horizon step x steps of this launch), or None"""
    if "nempc_wide" not in ev.kernel_name or wl["integ"] == "rk4":
        return None
    try:
        per_step = json.load(open(os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json"))).get("C4_per_horizon_step")
    except (OSError, ValueError):
        return None
    return None if per_step is None else per_step * steps_per_eval


def _cpu_block_chunk(args):
    """a chunk of problems through the O(H) per-step block restatement (oracle/blocks_np.py): the best-effort vectorised CPU path"""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    wl, Z, X0, lam = args
    from oracle.blocks_np import BlockEvaluator
    from oracle.mlp_np import MLP
    from oracle.objectives_np import SeparableQuadraticObjective
    cache = _cpu_block_chunk.__dict__.setdefault("cache", {})
    key = json.dumps(wl, sort_keys=True)
    if key not in cache:
        mlp = MLP.glorot(wl["dims"], wl["x"], wl["u"], seed=0, dtype=np.float32)
        obj = SeparableQuadraticObjective.tracking(wl["H"], wl["x"], wl["u"], np.linspace(1.0, 2.0, wl["x"]), np.linspace(0.1, 0.2, wl["u"]))
        cache[key] = BlockEvaluator(mlp, wl["integ"], wl["H"], DT=wl["DT"], objective=obj)
    out = cache[key].evaluate(Z, X0, lam, 1.0)
    return float(out["hes_vals"].sum())


def cpu_blocks_run(wl_name, per_proc=64, procs=None):
    """SURVEY 8d "R2": the same evaluation restated in O(H) per-step blocks + sparse assembly (what the kernels compute), vectorised numpy,
    one process per core -- reported so that the speed-up is not inflated by the O(H^3) density of the reference's own assembly."""
    import multiprocessing as mp
    wl = {k: v for k, v in WORKLOADS[wl_name].items() if k != "desc"}
    procs = procs or os.cpu_count() or 1
    _, _, Z, X0, lam = make_problem(wl, per_proc * procs)
    jobs = [(wl, Z[i * per_proc:(i + 1) * per_proc], X0[i * per_proc:(i + 1) * per_proc], lam[i * per_proc:(i + 1) * per_proc]) for i in range(procs)]
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_cpu_block_chunk, [(wl, Z[:2], X0[:2], lam[:2])] * procs, chunksize=1)      # per-process caches, imports
        t0 = time.perf_counter()
        pool.map(_cpu_block_chunk, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return dict(value=per_proc * procs * wl["H"] / dt, unit=UNIT, cores=procs, kind="port",
                sample=f"{per_proc * procs} of {wl['B']} problems, O(H) per-step block restatement of the reference algorithm (oracle/blocks_np.py, vectorised numpy, "
                       f"one process per core): the best-effort CPU path, without the reference's O(H^3) dense assembly")


def block_cpu_baseline(name, nprob=2):
    """CPU figure beside a wide workload: the O(H) per-step block restatement (oracle/blocks_np.py, numpy, one core) on `nprob`
    problems.  The reference's own dense path is O(H^3): 0.8 GB per problem for C3, 197 GB for C4 -- not runnable."""
    from oracle.blocks_np import BlockEvaluator
    wl = {k: v for k, v in WORKLOADS[name].items() if k != "desc"}
    mlp, obj, Z, X0, lam = make_problem(wl, nprob, seed=5)
    be = BlockEvaluator(mlp, wl["integ"], wl["H"], DT=wl["DT"], objective=obj)
    be.evaluate(Z[:1], X0[:1], lam[:1], 1.0)
    t0 = time.perf_counter()
    be.evaluate(Z, X0, lam, 1.0)
    dt = time.perf_counter() - t0
    return {"value": nprob * wl["H"] / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{nprob} of {wl['B']} problems, O(H) block restatement of the reference algorithm (oracle/blocks_np.py, numpy); the reference's dense "
                      f"O(H^3) assembly is not runnable at this size"}


def named_workload(name, world, rank, local, peaks, with_cpu, steps=3, e2e_cap=8192):
    """BASELINE configs C3 / C4 at their NAMED batch, beside the headline: the fixed batch is SPLIT over the ranks (strong scaling,
    `sharding.shard_range`; no collective on the data path).  Device-resident value, dominant-kernel roofline, e2e through the
    host-buffer call (on at most `e2e_cap` problems per rank: C4's value arrays are 0.6 MB per problem) and a CPU figure."""
    import torch
    import torch.distributed as dist
    from oracle.mlp_np import MLP
    from oracle.objectives_np import SeparableQuadraticObjective
    from pyneuralempc_b200 import NlpEvaluator
    from pyneuralempc_b200.sharding import shard_range
    wl = WORKLOADS[name]
    lo, hi = shard_range(wl["B"], rank, world)
    B, H, xd, ud = hi - lo, wl["H"], wl["x"], wl["u"]
    n, m = H * (xd + ud), H * xd
    mlp = MLP.glorot(wl["dims"], xd, ud, seed=0, dtype=np.float32)
    obj = SeparableQuadraticObjective.tracking(H, xd, ud, np.linspace(1.0, 2.0, xd), np.linspace(0.1, 0.2, ud))
    ev = NlpEvaluator(mlp.weights, xd, ud, H, wl["integ"], DT=wl["DT"], compute_dtype="float32", io_dtype="float64", device=local)
    ev.set_objective(obj.lin, obj.quad, obj.ref)
    dev = ev.tdevice
    gen = torch.Generator(device=dev); gen.manual_seed(4321 + lo)
    z = torch.rand((B, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    x0 = torch.rand((B, xd), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    lam = torch.randn((B, m), dtype=torch.float64, device=dev, generator=gen)
    out = ev.alloc_outputs(B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); a.record()
        for _ in range(reps):
            fn()
        b.record(); barrier()
        t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(1 if name == "C4rk4" else 2):
        ev.eval(z, x0, lam, 1.0, out=out)
    l0 = ev.launch_count
    ms = timed(lambda: ev.eval(z, x0, lam, 1.0, out=out), steps)
    launches = ev.launch_count - l0
    k_ms = timed(lambda: ev.eval(z, x0, lam, 1.0, want=("resid", "jac", "hes"), out=out), steps)
    # e2e: pinned host buffers in and out through nempc_eval_host, H2D + D2H inside the timed region
    Be = min(B, e2e_cap)
    buf = ev.pinned_buffers(Be)
    buf["z"][...] = z[:Be].cpu().numpy(); buf["x0"][...] = x0[:Be].cpu().numpy(); buf["lam"][...] = lam[:Be].cpu().numpy()
    ev.eval_pinned(Be, 1.0); ev.eval_pinned(Be, 1.0)
    barrier()
    t0 = time.perf_counter()
    chk = 0.0
    for _ in range(2):
        chk += float(ev.eval_pinned(Be, 1.0)["obj"][0])
    torch.cuda.synchronize()
    te = torch.tensor([(time.perf_counter() - t0) / 2], dtype=torch.float64, device=dev)
    bt = torch.tensor([float(Be)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(bt, op=dist.ReduceOp.SUM)
    h2d, d2h = ev.host_io_bytes(Be)
    steps_total = wl["B"] * H
    res = {"workload": name + ": " + wl["desc"], "metric": METRIC, "unit": UNIT, "n_gpus": world, "scaling": "strong",
           "batch_total": wl["B"], "batch_per_gpu": B, "horizon": H, "value": steps_total / (ms * 1e-3), "ms_per_eval": ms,
           "steps": steps, "gpu_launches": launches,
           "eval": "residual + sparse Jacobian + lambda-contracted sparse Lagrangian Hessian + objective value/gradient, float64 I/O",
           "l2": "outputs of one evaluation (%.1f GB per GPU) >> 126 MB L2" % (sum(t.numel() * 8 for t in out.values()) / 1e9),
           "roofline": tensor_roofline(ev, wl, B * H, k_ms, peaks),
           "e2e": {"value": float(bt.item()) * H / float(te.item()), "unit": UNIT, "batch_per_gpu": Be, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": float(te.item()) * 1e3, "api": "NlpEvaluator.eval_pinned -> nempc_eval_host (pinned host buffers, H2D + D2H inside the timed region)"},
           "cpu_baseline": None}
    ev.close()
    del z, x0, lam, out
    torch.cuda.empty_cache()
    if with_cpu and rank == 0:
        try:
            res["cpu_baseline"] = block_cpu_baseline(name)
        except Exception as exc:          # noqa: BLE001 -- never lose the line to a side measurement
            res["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "failed: " + repr(exc)[:160]}
    return res


def c2_float64(world, rank, local, steps=10, name="C2"):
    """C2 (or, name="C3", the cart-pole class at its named batch split over the ranks) with FLOAT64 network arithmetic (the reference assembles
    in float64, optimizer/ipopt.py:66-86; this is the 1e-10 parity mode): nempc_fast64_kernel / the DMMA path against the measured FP64 FMA
    peak (B200's FP64 tensor cores have the rate of its FP64 FMA pipe, so one denominator serves both)."""
    import torch
    import torch.distributed as dist
    from pyneuralempc_b200 import NlpEvaluator
    from pyneuralempc_b200.engine import measure_fma_peak
    from pyneuralempc_b200.sharding import shard_range
    wl = WORKLOADS[name]
    B = wl["B"]
    if name != "C2":
        lo, hi = shard_range(B, rank, world)
        B = hi - lo
    mlp, obj, Z, X0, lam = make_problem(wl, B, seed=99 + rank)
    ev = NlpEvaluator(mlp.weights, wl["x"], wl["u"], wl["H"], wl["integ"], DT=wl["DT"], compute_dtype="float64", io_dtype="float64", device=local)
    ev.set_objective(obj.lin, obj.quad, obj.ref)
    dev = ev.tdevice
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    z, x0, lm = t(Z), t(X0), t(lam)
    out = ev.alloc_outputs(B)
    for _ in range(3):
        ev.eval(z, x0, lm, 1.0, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a.record()
    for _ in range(steps):
        ev.eval(z, x0, lm, 1.0, want=("resid", "jac", "hes"), out=out)
    b.record()
    torch.cuda.synchronize()
    tm = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    k_ms = float(tm.item())
    peak = measure_fma_peak(local, "float64", 200)
    flops = ev.flops_per_step * B * wl["H"]
    tot = world * B if name == "C2" else wl["B"]
    res = {"workload": name + " with float64 network arithmetic: " + wl["desc"], "metric": METRIC, "unit": UNIT, "n_gpus": world,
           "scaling": "weak" if name == "C2" else "strong",
           "value": tot * wl["H"] / (k_ms * 1e-3), "ms_per_eval": k_ms, "steps": steps, "dtype": "f64",
           "roofline": {"bound": "fp64-fma" if name == "C2" else "fp64 tensor (DMMA) = fp64-fma rate", "kernel": ev.kernel_name, "achieved": flops / (k_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                        "frac": flops / (k_ms * 1e-3) / 1e12 / peak, "peak_source": "measured live: register-resident DFMA loop (nempc_measure_fma_peak)",
                        "kernel_ms": k_ms, "flops_per_horizon_step": ev.flops_per_step, "traffic": None}}
    ev.close()
    return res


def callback_latency(device, n=1000):
    """BASELINE config C1 (Lotka-Volterra MLP, H=25, ONE problem): wall time of one solver callback through the host-buffer C-ABI
    call -- all five outputs -- with the shipped RK4 integrator (examples/lotka_volterra/run.py:77) and the discrete one."""
    from pyneuralempc_b200 import NlpEvaluator
    res = {}
    for integ in ("rk4", "discrete"):
        wl = dict(WORKLOADS["C1"]); wl["integ"] = integ; wl["DT"] = 0.1
        mlp, obj, Z, X0, lam = make_problem({k: v for k, v in wl.items() if k != "desc"}, 1, seed=7)
        ev = NlpEvaluator(mlp.weights, wl["x"], wl["u"], wl["H"], integ, DT=0.1, device=device)
        ev.set_objective(obj.lin, obj.quad, obj.ref)
        buf = ev.pinned_buffers(1)
        buf["z"][...] = Z; buf["x0"][...] = X0; buf["lam"][...] = lam
        for _ in range(20):
            ev.eval_pinned(1, 1.0)
        t0 = time.perf_counter()
        for _ in range(n):
            ev.eval_pinned(1, 1.0)
        res[integ] = (time.perf_counter() - t0) / n * 1e6
        ev.close()
    return {"workload": "C1: " + WORKLOADS["C1"]["desc"], "us_per_callback_rk4": res["rk4"], "us_per_callback_discrete": res["discrete"],
            "outputs": "residual + Jacobian + Hessian values + objective + gradient",
            "api": "NlpEvaluator.eval_pinned -> nempc_eval_host: zero-copy on mapped pinned buffers, warp-per-step kernel (nempc_small_kernel)"}


def gpu_run(args):
    import torch
    import torch.distributed as dist
    from pyneuralempc_b200 import NlpEvaluator
    from pyneuralempc_b200.engine import measure_fma_peak

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa_cpus = None
    if world > 1 and not os.environ.get("NEMPC_NO_NUMA_BIND"):
        from pyneuralempc_b200.sharding import bind_host_to_gpu
        numa_cpus = bind_host_to_gpu(local)        # before any pinned allocation: host staging next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = WORKLOADS[args.workload]
    if args.workload in ("C3", "C4", "C4rk4"):
        # wide networks at their named batch: the fixed batch is split over the ranks (strong scaling), data generated on the device
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        r = named_workload(args.workload, world, rank, local, peaks, with_cpu=(world == 1 and not args.no_cpu_baseline), steps=max(1, min(args.steps, 20)))
        if rank == 0:
            line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": r["steps"], "warmup": 3, "ms_per_step": r["ms_per_eval"],
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": {"workload": r["workload"], "batch_total": r["batch_total"], "batch_per_gpu": r["batch_per_gpu"], "horizon": r["horizon"],
                               "io_dtype": "float64", "parallelism": f"fixed batch split over {world} GPU(s) (sharding.shard_range), no data-path collective",
                               "l2": r["l2"], "eval": r["eval"]},
                    "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "clocks": sampler.stop(), "roofline": r["roofline"], "cpu_baseline": r["cpu_baseline"]}
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return
    B = args.batch or wl["B"]
    mlp, obj, Z, X0, lam = make_problem(wl, B, seed=1234 + rank)        # every rank owns different problems
    ev = NlpEvaluator(mlp.weights, wl["x"], wl["u"], wl["H"], wl["integ"], DT=wl["DT"], compute_dtype=args.dtype,
                      io_dtype=args.io_dtype, device=local, kernel=args.kernel)
    ev.set_objective(obj.lin, obj.quad, obj.ref)
    dev = ev.tdevice
    tdt = ev.tdtype
    # rotating input/output sets so successive steps never re-read an L2-resident set (inputs+outputs of all sets >> 126 MB L2)
    nsets = args.sets
    zs = [torch.as_tensor(np.roll(Z, s, axis=0), dtype=tdt, device=dev) for s in range(nsets)]
    x0s = [torch.as_tensor(np.roll(X0, s, axis=0), dtype=tdt, device=dev) for s in range(nsets)]
    lams = [torch.as_tensor(np.roll(lam, s, axis=0), dtype=tdt, device=dev) for s in range(nsets)]
    outs = [ev.alloc_outputs(B) for _ in range(nsets)]
    set_bytes = sum(t.numel() * t.element_size() for t in (zs[0], x0s[0], lams[0])) + sum(t.numel() * t.element_size() for t in outs[0].values())

    def step(i, want=("resid", "jac", "hes", "obj", "grad")):
        s = i % nsets
        ev.eval(zs[s], x0s[s], lams[s], 1.0, want=want, out=outs[s])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(max(3, args.warmup)):
        step(i)
    if rank == 0:
        sampler.wait_ready()
    barrier()
    l0 = ev.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    launches = ev.launch_count - l0
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    steps_per_eval = B * wl["H"]
    value = world * steps_per_eval / (ms_step * 1e-3)

    # ---- dominant kernel alone (residual+Jacobian+Hessian kernel): K back-to-back launches between one event pair ----------
    nk = max(10, args.steps)
    for i in range(3):
        step(i, want=("resid", "jac", "hes"))
    ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ka.record()
    for i in range(nk):
        step(i, want=("resid", "jac", "hes"))
    kb.record()
    torch.cuda.synchronize()
    k_ms = ka.elapsed_time(kb) / nk

    # ---- e2e: host buffers through the plug-in call, H2D + D2H inside the timed region -----------------------------------
    # The step's inputs sit in pinned host memory (NlpEvaluator.pinned_buffers); every timed step uploads them,
    # evaluates and downloads ALL results into pinned host memory (nempc_eval_host), then reads the objective values.
    npdt = np.float64 if args.io_dtype == "float64" else np.float32
    buf = ev.pinned_buffers(B)
    buf["z"][...] = Z.astype(npdt); buf["x0"][...] = X0.astype(npdt); buf["lam"][...] = lam.astype(npdt)
    for _ in range(3):
        ev.eval_pinned(B, 1.0)
    barrier()
    t0 = time.perf_counter()
    chk = 0.0
    for _ in range(args.steps):
        o = ev.eval_pinned(B, 1.0)
        chk += float(o["obj"][0])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * steps_per_eval / (float(te.item()) / args.steps)
    # the numpy-in / numpy-out callback form (adds the pageable -> pinned staging copy), reported beside it
    t0 = time.perf_counter()
    for _ in range(max(3, args.steps // 4)):
        ev.eval_host(Z, X0, lam, 1.0)
    cb_ms = (time.perf_counter() - t0) / max(3, args.steps // 4) * 1e3
    h2d, d2h = ev.host_io_bytes(B)
    clocks = sampler.stop() if rank == 0 else None     # sampled across the timed loop, the kernel-only loop and the e2e loop

    # ---- MPC solves/s: the batched on-device interior-point solver (nempc_solve) on the same B problems --------------------------
    solves = None
    if args.io_dtype == "float64" and not args.no_solver:
        lb, ub = solver_bounds(wl)
        x0d = torch.as_tensor(X0, dtype=torch.float64, device=dev)
        sopt = dict(tol=1e-4, max_iter=40)          # the reference's IPOPT acceptable_tol (optimizer/ipopt.py:185)
        ev.solve(x0d, lb, ub, **sopt)
        barrier()
        times = []                                  # median of 9 solves: the host-driven iteration loop jitters by +-15 % run to run
        for _ in range(9):
            t0 = time.perf_counter()
            so = ev.solve(x0d, lb, ub, **sopt)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        ts = torch.tensor([statistics.median(times)], dtype=torch.float64, device=dev)
        conv = torch.tensor([float((so["status"] == 0).sum().item())], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            dist.all_reduce(conv, op=dist.ReduceOp.SUM)
        # copy-inclusive: x0 from pinned host memory up, the solution (x_pred, u) and the status back into pinned host memory, every solve
        x0_pin = torch.as_tensor(X0, dtype=torch.float64).pin_memory()
        z_pin = torch.empty((B, ev.n), dtype=torch.float64).pin_memory()
        st_pin = torch.empty(B, dtype=torch.int32).pin_memory()

        def solve_e2e():
            so_ = ev.solve(x0_pin.to(dev, non_blocking=True), lb, ub, **sopt)
            z_pin.copy_(so_["z"], non_blocking=True)
            st_pin.copy_(so_["status"], non_blocking=True)
            torch.cuda.synchronize()
            return int((st_pin == 0).sum())

        solve_e2e()
        barrier()
        times_e = []
        for _ in range(9):
            t0 = time.perf_counter()
            solve_e2e()
            times_e.append(time.perf_counter() - t0)
        tse = torch.tensor([statistics.median(times_e)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tse, op=dist.ReduceOp.MAX)
        solves = {"metric": "mpc_solves_per_s", "value": world * B / float(ts.item()), "unit": "solves/s", "batch_per_gpu": B,
                  "e2e": {"value": world * B / float(tse.item()), "unit": "solves/s", "ms_per_batch": float(tse.item()) * 1e3,
                          "h2d_bytes_per_step": int(x0_pin.numel() * 8), "d2h_bytes_per_step": int(z_pin.numel() * 8 + st_pin.numel() * 4),
                          "api": "NlpEvaluator.solve: x0 H2D from pinned memory, (x_pred, u) and status D2H into pinned memory inside the timed region"},
                  "ms_per_batch": float(ts.item()) * 1e3, "converged_frac": float(conv.item()) / (world * B),
                  "ipm_iterations_mean": float(so["iterations"].double().mean().item()), "outer_iterations": so["outer_iterations"],
                  "tol": sopt["tol"], "solver": "nempc_solve: primal-dual interior point, Riccati KKT sweep on the block-banded values, x0 device-resident",
                  "bounds": "u in [-1, 0.2] (run.py:72-74), states free",
                  "loop": "device: one CUDA graph, nested WHILE nodes (cudaGraphSetConditional)" if so.get("used_graph") else "host-issued kernels, one synchronisation per iteration",
                  "unaccepted_steps": so.get("unaccepted_steps")}
        if rank == 0:
            # ONE problem (what a closed-loop controller solves per sample): latency with the device-side loop and with the host-issued one
            one = {}
            for mode, name in (("1", "device_loop_ms"), ("0", "host_loop_ms")):
                os.environ["NEMPC_SOLVE_GRAPH"] = mode
                for _ in range(3):
                    ev.solve(x0d[:1], lb, ub, **sopt)
                tt = []
                for _ in range(15):
                    t0 = time.perf_counter()
                    s1 = ev.solve(x0d[:1], lb, ub, **sopt)
                    torch.cuda.synchronize()
                    tt.append(time.perf_counter() - t0)
                one[name] = statistics.median(tt) * 1e3
                one["iterations"] = int(s1["iterations"][0].item())
            os.environ.pop("NEMPC_SOLVE_GRAPH", None)
            solves["single_problem"] = one
    ev_name, ev_flops, ev_bytes = ev.kernel_name, ev.flops_per_step, ev.bytes_per_step()
    side = None
    if args.workload == "C2" and not args.no_side_workloads:
        # BASELINE configs C3 / C4 at their named batch (split over the ranks: strong scaling); every rank takes part
        side = {}
        peaks_side = {}
        try:
            peaks_side = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        ev.close()                                      # release the headline workload's device buffers first
        del zs, x0s, lams, outs
        torch.cuda.empty_cache()
        for nm in ("C3", "C4", "C4rk4"):
            try:
                side[nm] = named_workload(nm, world, rank, local, peaks_side, with_cpu=(world == 1 and not args.no_cpu_baseline),
                                          steps=2 if nm == "C4rk4" else 3)
            except Exception as exc:                  # a side measurement must never cost the headline line
                side[nm] = {"error": repr(exc)[:300]}
        try:
            side["C2_f64"] = c2_float64(world, rank, local)
        except Exception as exc:                      # noqa: BLE001
            side["C2_f64"] = {"error": repr(exc)[:300]}
        try:
            side["C3_f64"] = c2_float64(world, rank, local, steps=2, name="C3")
        except Exception as exc:                      # noqa: BLE001
            side["C3_f64"] = {"error": repr(exc)[:300]}
        if rank == 0:
            try:
                side["C1_callback"] = callback_latency(local)
            except Exception as exc:                  # noqa: BLE001
                side["C1_callback"] = {"error": repr(exc)[:200]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel ------------------------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    fma_peak = measure_fma_peak(local, args.dtype, 300)
    flops = ev_flops * steps_per_eval
    achieved = flops / (k_ms * 1e-3) / 1e12
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    alg_bytes = ev_bytes * steps_per_eval
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
    if os.path.exists(tr_path):
        try:
            traffic = json.load(open(tr_path)).get(args.workload)
        except (OSError, ValueError):
            traffic = None
    roofline = {"bound": "fp32-fma" if args.dtype == "float32" else "fp64-fma",
                "note": "compute bound on the CUDA-core FMA pipe (arithmetic intensity %.0f flop/B); neither HBM nor tensor cores bound this kernel" % (flops / alg_bytes),
                "kernel": ev_name, "achieved": achieved, "peak": fma_peak, "unit": "TFLOP/s", "frac": achieved / fma_peak,
                "peak_source": "measured live: register-resident FMA loop (nempc_measure_fma_peak)",
                "kernel_ms": k_ms, "flops_per_horizon_step": ev_flops, "bytes_per_horizon_step": ev_bytes,
                "hbm_achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak,
                "hbm_peak_source": "MEASURED_PEAKS.json" if peaks else "fallback", "hbm_frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                "traffic": traffic}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-worker", "--workload", args.workload,
                            "--cpu-sample", str(args.cpu_sample)] + (["--no-solver"] if args.no_solver else []), capture_output=True, text=True)
        try:
            c = json.loads(r.stdout.strip().splitlines()[-1])
            cpu = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": c.get("kind", "port"), "sample": c["sample"]}
            if solves is not None and "solver" in c:
                solves["cpu_baseline"] = c["solver"]
            cpu["vectorised_port"] = c.get("blocks")       # SURVEY 8d "R2": best-effort vectorised CPU path beside the reference-literal one
        except (ValueError, IndexError):
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": "failed: " + r.stderr[-300:]}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "batch_per_gpu": B, "horizon": wl["H"], "io_dtype": args.io_dtype,
                       "horizon_steps_per_eval_per_gpu": steps_per_eval, "parallelism": f"independent problems sharded over {world} GPU(s), no data-path collective",
                       "host_numa_bind": (f"rank 0 on {len(numa_cpus)} cores next to its GPU" if numa_cpus else "none"),
                       "l2": f"{nsets} rotating input/output sets, {set_bytes * nsets / 1e6:.0f} MB total > 126 MB L2",
                       "eval": "residual + sparse Jacobian + lambda-contracted sparse Lagrangian Hessian + objective value/gradient"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(te.item()) / args.steps * 1e3, "api": "NlpEvaluator.eval_pinned -> nempc_eval_host (pinned host buffers, chunk-pipelined H2D|kernels|D2H, replayed as a CUDA graph)",
                    "numpy_callback_ms_per_step": cb_ms},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "mpc_solves": solves, "other_workloads": side}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def reference_run(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    if args.workload in ("C3", "C4", "C4rk4"):
        # the reference's dense O(H^3) assembly needs 0.8 GB (C3) / 197 GB (C4) per problem: the O(H) block restatement is what can be timed
        b = block_cpu_baseline(args.workload, nprob=4)
        r = {"value": b["value"], "ms_per_step": 4 * wl["H"] / b["value"] * 1e3, "cores": 1, "sample": b["sample"], "kind": "port"}
    else:
        r = cpu_reference_run(args.workload, args.cpu_sample, max(1, args.steps), max(1, args.warmup), budget_s=args.cpu_budget)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", str(args.gpus))),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "note": "reference CPU path (the reference's own integrator + IpoptProblem where its bytecode is present, else the oracle port; TensorFlow/JAX/cyipopt are not installable here), bounded sample per step"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r.get("kind", "port"), "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="problems per GPU (default: the workload's)")
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"], help="arithmetic type of the network/chain rule")
    ap.add_argument("--io-dtype", default="float64", choices=["float32", "float64"], help="element type of z/lambda/values (reference: float64)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "fast", "tc"])
    ap.add_argument("--no-side-workloads", action="store_true", help="skip the short C3 (tensor-core kernel) measurement reported beside the headline")
    ap.add_argument("--sets", type=int, default=8)
    ap.add_argument("--cpu-sample", type=int, default=128, help="problems per CPU-baseline step")
    ap.add_argument("--cpu-budget", type=float, default=90.0, help="--impl reference: target wall seconds of the timed loop (sets the sample size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solver", action="store_true", help="skip the MPC-solves/s leg")
    ap.add_argument("--cpu-baseline-worker", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_baseline_worker:
        # the solver leg first: importing the reference (shim) leaves stub `tensorflow` / `jax` modules in sys.modules, and SciPy's
        # array-API dispatch then probes them on every call (measured: SLSQP 30x slower in that process)
        solver = None if args.no_solver else cpu_solver_run(args.workload)
        blocks = cpu_blocks_run(args.workload)
        out = cpu_reference_run(args.workload, args.cpu_sample, 2, 1)
        out["blocks"] = blocks
        if solver is not None:
            out["solver"] = solver
        print(json.dumps(out))
        return
    if args.impl == "reference":
        reference_run(args)
    else:
        gpu_run(args)


if __name__ == "__main__":
    main()
