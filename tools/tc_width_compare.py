import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from pyneuralempc_b200 import NlpEvaluator
from oracle.mlp_np import MLP
for hw in (32, 64, 128):
  for kern in ("generic", "tc"):
    mlp = MLP.glorot([5, hw, hw, hw, 4], 4, 1, seed=1)
    H, B = 100, 4096
    ev = NlpEvaluator(mlp.weights, 4, 1, H, "rk4", DT=0.1, compute_dtype="float32", io_dtype="float32", kernel=kern)
    g = torch.Generator(device="cuda").manual_seed(0)
    Z = torch.rand(B, ev.n, device="cuda", generator=g) * 2 - 1
    X0 = torch.rand(B, 4, device="cuda", generator=g) * 2 - 1
    lam = torch.randn(B, ev.m, device="cuda", generator=g)
    for _ in range(3): ev.eval(Z, X0, lam, want=("resid", "jac", "hes"))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ev.eval(Z, X0, lam, want=("resid", "jac", "hes"))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"hw={hw} {kern}: {ms:.3f} ms  {B*H/ms*1e3:.3e} steps/s  {ev.kernel_name}")
