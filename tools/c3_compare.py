#!/usr/bin/env python
"""C3 (cart-pole 5-128x3-4, RK4, H=100) on the two tensor-core kernels that serve hidden width 128: run once per setting of
NEMPC_PREFER_WIDE128 (0: forward second-order nempc_tc_kernel, 1: adjoint-form nempc_wide_kernel<..., HW=128>): parity vs the oracle on a
small batch, then the evaluation time at B problems.    NEMPC_PREFER_WIDE128=1 python tools/c3_compare.py [B]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

if __name__ == "__main__":
    import torch
    from oracle.mlp_np import MLP
    from pyneuralempc_b200 import NlpEvaluator
    import wide_check
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    ok = True
    for integ, dims, x, u, H, b in (("rk4", [5, 128, 128, 128, 4], 4, 1, 13, 7), ("discrete", [5, 128, 128, 4], 4, 1, 9, 30), ("rk4", [3, 128, 128, 2], 2, 1, 50, 5)):
        ok &= wide_check.check(dims, x, u, H, b, integ, seed=H)
    print("ALL OK" if ok else "FAILURES")
    H = 100
    mlp = MLP.glorot([5, 128, 128, 128, 4], 4, 1, seed=0, dtype=np.float32)
    rng = np.random.default_rng(1)
    z = torch.as_tensor(rng.uniform(-1, 1, (B, H * 5))).cuda()
    x0 = torch.as_tensor(rng.uniform(-1, 1, (B, 4))).cuda()
    lam = torch.as_tensor(rng.standard_normal((B, H * 4))).cuda()
    for integ in ("rk4", "discrete"):
        ev = NlpEvaluator(mlp.weights, 4, 1, H, integ, DT=0.1, kernel="tc")
        for want in (("resid", "jac"), ("resid", "jac", "hes")):
            out = ev.alloc_outputs(B, want)
            ev.eval(z, x0, lam, 1.0, want=want, out=out)
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                ev.eval(z, x0, lam, 1.0, want=want, out=out)
            b_.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b_) / 3
            print(f"C3-shape {integ} B={B} want={'+'.join(want):14s}: {ms:8.3f} ms {B * H / ms * 1e3:.3e} steps/s   {ev.kernel_name[:60]}", flush=True)
        ev.close()
