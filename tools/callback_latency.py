#!/usr/bin/env python
"""latency of ONE solver callback through the host-buffer C-ABI call (nempc_eval_host) for a single problem: BASELINE config C1
(Lotka-Volterra, H=25, B=1), the situation of the reference's IpoptProblem callbacks (optimizer/ipopt.py:30-96)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from bench import WORKLOADS, make_problem
    from pyneuralempc_b200 import NlpEvaluator
    for name, integ in (("C1", "discrete"), ("C1", "rk4")):
        wl = dict(WORKLOADS[name]); wl["integ"] = integ; wl["DT"] = 0.1
        mlp, obj, Z, X0, lam = make_problem({k: v for k, v in wl.items() if k != "desc"}, 1)
        ev = NlpEvaluator(mlp.weights, wl["x"], wl["u"], wl["H"], integ, DT=0.1)
        ev.set_objective(obj.lin, obj.quad, obj.ref)
        buf = ev.pinned_buffers(1)
        buf["z"][...] = Z; buf["x0"][...] = X0; buf["lam"][...] = lam
        for want in (("resid", "jac", "hes", "obj", "grad"), ("resid", "jac", "obj", "grad"), ("hes",)):
            for _ in range(20):
                ev.eval_pinned(1, 1.0, want=want)
            n = 2000
            t0 = time.perf_counter()
            for _ in range(n):
                ev.eval_pinned(1, 1.0, want=want)
            us = (time.perf_counter() - t0) / n * 1e6
            print(f"{name} {integ:8s} want={','.join(want):28s} {us:7.1f} us per callback (pinned buffers, {ev.kernel_name.split('<')[0]})", flush=True)
        t0 = time.perf_counter()
        for _ in range(500):
            ev.eval_host(Z, X0, lam, 1.0)
        print(f"{name} {integ:8s} numpy in/out (eval_host)              {(time.perf_counter() - t0) / 500 * 1e6:7.1f} us per callback", flush=True)
        ev.close()


if __name__ == "__main__" and "--solve" not in sys.argv:
    main()


def single_problem_solve():
    """one NMPC solve of ONE problem (the reference's NMPC.next situation) with the on-device interior-point solver"""
    import torch
    from bench import WORKLOADS, make_problem, solver_bounds
    from pyneuralempc_b200 import NlpEvaluator
    for integ in ("discrete", "rk4"):
        wl = dict(WORKLOADS["C1"]); wl["integ"] = integ; wl["DT"] = 0.1
        mlp, obj, Z, X0, lam = make_problem({k: v for k, v in wl.items() if k != "desc"}, 1)
        ev = NlpEvaluator(mlp.weights, wl["x"], wl["u"], wl["H"], integ, DT=0.1)
        ev.set_objective(obj.lin, obj.quad, obj.ref)
        lb, ub = solver_bounds(wl)
        x0 = torch.as_tensor(X0).cuda()
        for _ in range(3):
            so = ev.solve(x0, lb, ub, tol=1e-4, max_iter=40)
        torch.cuda.synchronize()
        n = 50
        t0 = time.perf_counter()
        for _ in range(n):
            so = ev.solve(x0, lb, ub, tol=1e-4, max_iter=40)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / n * 1e3
        print(f"C1 {integ:8s} single-problem NMPC solve (nempc_solve, tol 1e-4): {ms:.2f} ms, {int(so['iterations'][0])} IPM iterations, status {int(so['status'][0])}", flush=True)
        ev.close()


if __name__ == "__main__" and "--solve" in sys.argv:
    single_problem_solve()
