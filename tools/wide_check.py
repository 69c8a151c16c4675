#!/usr/bin/env python
"""Development check of nempc_wide_kernel (hidden width 256, tcgen05, adjoint form) on a GPU box: every output against the
numpy oracle (float64) and against the generic FFMA kernel, worst element printed; then a timing of the C4 shape.

  python tools/wide_check.py [--modes resid,jac,hes] [--time B]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worst(got, ref):
    d = np.abs(got - ref)
    tol = 1e-5 * np.abs(ref) + 1e-6 * max(1.0, float(np.abs(ref).max()))
    i = int(np.argmax(d / tol))
    return float((d / tol).flat[i]), float(d.flat[i]), float(ref.flat[i])


def check(dims, x, u, H, B, integ, seed=0, DT=0.1):
    import torch
    from oracle.blocks_np import BlockEvaluator
    from oracle.mlp_np import MLP
    from oracle.objectives_np import SeparableQuadraticObjective
    from pyneuralempc_b200 import NlpEvaluator
    rng = np.random.default_rng(seed)
    mlp = MLP.glorot(dims, x, u, seed=seed + 1, dtype=np.float32)
    obj = SeparableQuadraticObjective.tracking(H, x, u, np.linspace(1.0, 2.0, x), np.linspace(0.1, 0.2, u), x_ref=rng.uniform(-1, 1, (H, x)))
    n, m = H * (x + u), H * x
    Z, X0 = rng.uniform(-1, 1, (B, n)), rng.uniform(-1, 1, (B, x))
    lam, sig = rng.standard_normal((B, m)), rng.uniform(0.5, 1.5, B)
    ref = BlockEvaluator(mlp, integ, H, DT=DT, objective=obj).evaluate(Z, X0, lam, sig)
    ok = True
    for kernel in ("tc", "generic"):
        ev = NlpEvaluator(mlp.weights, x, u, H, integ, DT=DT, compute_dtype="float32", io_dtype="float64", kernel=kernel)
        ev.set_objective(obj.lin, obj.quad, obj.ref)
        t = [torch.as_tensor(a).cuda() for a in (Z, X0, lam, sig)]
        for want in (("resid",), ("resid", "jac"), ("resid", "jac", "hes")):
            out = ev.eval(*t, want=want)
            torch.cuda.synchronize()
            for k in want:
                w, d, r = worst(out[k].cpu().numpy(), ref[{"resid": "resid", "jac": "jac_vals", "hes": "hes_vals"}[k]])
                flag = "" if w <= 1.0 else "   <-- FAIL"
                if w > 1.0 and kernel == "tc":
                    ok = False
                print(f"{dims} {integ:8s} H={H} B={B} {kernel:8s} want={'+'.join(want):14s} {k:5s} worst |d|/tol = {w:9.3g} (|d| = {d:.3g} at ref {r:.3g}){flag}", flush=True)
        print("   kernel:", ev.kernel_name, flush=True)
        ev.close()
    return ok


def timing(B, H=200, integ="discrete", reps=3):
    import torch
    from oracle.mlp_np import MLP
    from pyneuralempc_b200 import NlpEvaluator
    x, u = 12, 4
    dims = [16, 256, 256, 256, 256, 12]
    mlp = MLP.glorot(dims, x, u, seed=0, dtype=np.float32)
    rng = np.random.default_rng(1)
    n, m = H * (x + u), H * x
    z = torch.as_tensor(rng.uniform(-1, 1, (B, n))).cuda()
    x0 = torch.as_tensor(rng.uniform(-1, 1, (B, x))).cuda()
    lam = torch.as_tensor(rng.standard_normal((B, m))).cuda()
    ev = NlpEvaluator(mlp.weights, x, u, H, integ, DT=0.1, kernel="tc")
    for want in (("resid",), ("resid", "jac"), ("resid", "jac", "hes")):
        out = ev.alloc_outputs(B, want)
        ev.eval(z, x0, lam, 1.0, want=want, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            ev.eval(z, x0, lam, 1.0, want=want, out=out)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        if hasattr(ev.lib, "nempc_debug_wide_profile"):
            import ctypes
            buf = (ctypes.c_ulonglong * 16)()
            ev.lib.nempc_debug_wide_profile(buf)
            v = [int(x) for x in buf]
            ng = max(1, v[3])
            names = ["iss wait A", "iss wait ring", "iss issue", "gemms", "epi pre", "epi wait D", "epi body", "epi between", "prod wait slot", "prod issue"]
            print("   profile (cycles per GEMM, all launches since the last read): " + ", ".join(f"{n} {x / ng:.0f}" for n, x in zip(names, v[:10]) if n != "gemms") + f"; GEMMs {ng}")
        print(f"C4-shape {integ} B={B} H={H} want={'+'.join(want):14s}: {ms:9.3f} ms  {B * H / ms * 1e3:.3e} steps/s  "
              f"{ev.flops_per_step * B * H / ms / 1e9:.1f} TFLOP/s algorithmic (Jac+Hes count)", flush=True)
    ev.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", type=int, default=0)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    ok = True
    if not a.no_check:
        ok &= check([16, 256, 256, 256, 256, 12], 12, 4, 5, 3, "discrete")
    if not a.quick and not a.no_check:
        ok &= check([16, 256, 256, 256, 256, 12], 12, 4, 37, 9, "discrete", seed=3)       # several super-tiles, ragged tail
        ok &= check([16, 256, 256, 12], 12, 4, 7, 2, "unity", seed=4)
        ok &= check([5, 256, 256, 256, 4], 4, 1, 11, 5, "discrete", seed=5)
        ok &= check([3, 256, 256, 2], 2, 1, 6, 4, "discrete", seed=6)
        ok &= check([8, 256, 256, 256, 6], 6, 2, 6, 4, "discrete", seed=7)
        ok &= check([3, 256, 256, 2], 2, 1, 6, 4, "rk4", seed=8)
        ok &= check([16, 256, 256, 256, 256, 12], 12, 4, 5, 3, "rk4", seed=9)
        ok &= check([16, 256, 256, 256, 256, 12], 12, 4, 29, 10, "rk4", seed=10)
    print("ALL OK" if ok else "FAILURES", flush=True)
    if a.time:
        timing(a.time)
        timing(max(64, a.time // 8), integ="rk4")
    sys.exit(0 if ok else 1)
