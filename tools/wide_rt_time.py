"""Time nempc_wide_kernel on one shape with the specialised instantiation and with the run-time-shape one
(NEMPC_WIDE_RUNTIME_SHAPES=1, read once per process -- so this script re-executes itself).
usage: python tools/wide_rt_time.py [x u nhid hw H B kind]"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(x, u, nhid, hw, H, B, kind):
    import numpy as np
    import torch
    from oracle.mlp_np import MLP
    from pyneuralempc_b200.engine import NlpEvaluator
    mlp = MLP.glorot([x + u] + [hw] * nhid + [x], x, u, seed=0, dtype=np.float32)
    ev = NlpEvaluator(mlp.weights, x, u, H, integrator=kind, DT=0.1, compute_dtype="float32", io_dtype="float32")
    g = torch.Generator(device="cuda").manual_seed(1)
    Z = torch.rand((B, H * (x + u)), device="cuda", generator=g, dtype=torch.float32) * 2 - 1
    X0 = torch.rand((B, x), device="cuda", generator=g) * 2 - 1
    lam = torch.randn((B, H * x), device="cuda", generator=g)
    for want in (("resid",), ("resid", "jac"), ("resid", "jac", "hes")):
        for _ in range(3):
            ev.eval(Z, X0, lam, 1.0, want=want)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ev.eval(Z, X0, lam, 1.0, want=want)
        e1.record()
        torch.cuda.synchronize()
        print(f"  {'+'.join(want):16s} {e0.elapsed_time(e1) / 5:8.3f} ms   {ev.kernel_name[:60]}")
    ev.close()


if __name__ == "__main__":
    if os.environ.get("_WIDE_RT_CHILD"):
        a = sys.argv[1:]
        child(int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]), int(a[5]), a[6])
    else:
        args = sys.argv[1:] or ["4", "1", "3", "256", "100", "8192", "discrete"]
        for rt in ("0", "1"):
            print("NEMPC_WIDE_RUNTIME_SHAPES=" + rt, " ".join(args))
            env = dict(os.environ, _WIDE_RT_CHILD="1", NEMPC_WIDE_RUNTIME_SHAPES=rt)
            subprocess.run([sys.executable, os.path.abspath(__file__)] + args, env=env, check=True, timeout=600)
