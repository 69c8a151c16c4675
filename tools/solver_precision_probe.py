#!/usr/bin/env python
"""How far can nempc_solve be driven with float32 network arithmetic?  LV fixture network, unity, H = 10, 2048 problems (the case of
tests/test_gpu_solver.py::test_float32_network_and_large_batch): float32 solves at several KKT tolerances against the float64 solve."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if __name__ == "__main__":
    from oracle.mlp_np import load_lv_fixture_npz
    import test_gpu_solver as T
    lv = load_lv_fixture_npz(os.path.join(ROOT, "tests", "golden", "lv_mlp_weights.npz"))
    mlp, obj, lb, ub, X0 = T._setup("unity", "lv", 2, 1, 10, None, 0, lv, B=2048)
    o64 = T._ev(mlp, "unity", 10, None, obj, "float64").solve(X0, lb, ub, tol=1e-9)
    z64 = o64["z"].cpu().numpy()
    print("f64 tol 1e-9: converged", int((o64["status"] == 0).sum()), "iterations mean", float(o64["iterations"].double().mean()))
    for tol in (1e-4, 1e-5, 3e-6):
        a = T._ev(mlp, "unity", 10, None, obj, "float32").solve(X0, lb, ub, tol=tol, max_iter=100)
        b = T._ev(mlp, "unity", 10, None, obj, "float64").solve(X0, lb, ub, tol=tol, max_iter=100)
        ok = ((a["status"] == 0) & (b["status"] == 0)).cpu().numpy()
        same_it = (a["iterations"] == b["iterations"]).cpu().numpy()
        dz = np.abs(a["z"].cpu().numpy() - b["z"].cpu().numpy()).max(axis=1)
        print(f"same tol {tol:g}: both converged {int(ok.sum())}, same iteration count {int(same_it.sum())}, max |z32 - z64| all {dz[ok].max():.3e}, "
              f"where the counts agree {dz[ok & same_it].max():.3e}", flush=True)
    for comp in ("float64", "float32"):
        for tol in (1e-4, 1e-5, 3e-6, 1e-6, 3e-7):
            o = T._ev(mlp, "unity", 10, None, obj, comp).solve(X0, lb, ub, tol=tol, max_iter=100)
            ok = (o["status"] == 0).cpu().numpy()
            z = o["z"].cpu().numpy()
            print(f"{comp} tol {tol:g}: converged {int(ok.sum())}/2048, iterations mean {float(o['iterations'].double().mean()):.2f} max {int(o['iterations'].max())}, "
                  f"max |z - z64| over converged {np.abs(z - z64)[ok].max():.3e}, kkt max {float(o['kkt_error'].max()):.3e}", flush=True)
