"""5 -> hw x 3 -> 4 RK4 (H=100, B=4096) on the generic vs the tensor-core kernel for hidden widths 32 / 64 / 128: device-resident CUDA-event
timing.  usage: python tools/tc_width_compare.py [widths] [kernels], e.g. `64 tc` for a profiler run."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from pyneuralempc_b200 import NlpEvaluator
from oracle.mlp_np import MLP
widths = [int(a) for a in sys.argv[1].split(',')] if len(sys.argv) > 1 else (32, 64, 128)
kernels = sys.argv[2].split(',') if len(sys.argv) > 2 else ("generic", "tc")
for hw in widths:
  for kern in kernels:
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
