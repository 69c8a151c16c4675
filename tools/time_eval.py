#!/usr/bin/env python
"""Device-resident timing of one NLP evaluation (residual + Jacobian + Hessian launch) for a bench.py workload and a chosen
kernel: CUDA events over back-to-back launches after warm-up.  Development tool (profiles/ tables), not the driver's bench."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--horizon", type=int, default=0)
    ap.add_argument("--integ", default="")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--want", default="resid,jac,hes")
    ap.add_argument("--compute", default="float32")
    args = ap.parse_args()
    import torch
    from bench import WORKLOADS, make_problem
    from pyneuralempc_b200 import NlpEvaluator
    wl = dict(WORKLOADS[args.workload])
    if args.horizon: wl["H"] = args.horizon
    if args.integ: wl["integ"] = args.integ
    B = args.batch or wl["B"]
    mlp, obj, Z, X0, lam = make_problem({k: v for k, v in wl.items() if k != "desc"}, B)
    ev = NlpEvaluator(mlp.weights, wl["x"], wl["u"], wl["H"], wl["integ"], DT=wl["DT"] or 0.1, compute_dtype=args.compute,
                      io_dtype="float64", kernel=args.kernel)
    ev.set_objective(obj.lin, obj.quad, obj.ref)
    want = tuple(args.want.split(","))
    z, x0, lm = (torch.as_tensor(a).cuda() for a in (Z, X0, lam))
    out = ev.alloc_outputs(B, want)
    for _ in range(2):
        ev.eval(z, x0, lm, 1.0, want=want, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(args.reps):
        ev.eval(z, x0, lm, 1.0, want=want, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.reps
    steps = B * wl["H"]
    print(f"{args.workload} {wl['integ']} B={B} H={wl['H']} want={args.want} kernel={ev.kernel_name}\n"
          f"  {ms:.3f} ms/eval  {steps / ms * 1e3:.4g} horizon-steps/s  {ev.flops_per_step * steps / ms / 1e9:.2f} TFLOP/s algorithmic", flush=True)
    if hasattr(ev.lib, "nempc_debug_tc_profile"):      # -DNEMPC_TC_PROFILE build (NEMPC_LIB_PATH): cycles per phase, thread 0 of each CTA
        import ctypes
        buf = (ctypes.c_ulonglong * 16)()
        ev.lib.nempc_debug_tc_profile(buf)
        names = ["init", "l0 tanh", "pass2", "sync->mma", "mma issue", "mma wait", "pass1", "tanh", "out combine", "algebra", "outputs", "loop"]
        tot = sum(buf[:16]) or 1
        print("  phase cycles (share): " + ", ".join(f"{n} {buf[i] / tot:.1%}" for i, n in enumerate(names)), flush=True)
        ntl = (steps + 5) // 6 * (args.reps + 2) / 2          # thread 0 belongs to tile group 0: half of the tiles
        print(f"  cycles per tile of one group (C3 SPT=6; two groups in flight per SM): {tot / ntl:.0f}")
        print("  mma issue detail (cycles per batch): setup %.0f, corr-1 x8 %.0f, corr-2 x8 %.0f, main x8 + commit %.0f" % tuple(buf[i] / (ntl * 8) for i in (12, 13, 14, 4)))
    ev.close()


if __name__ == "__main__":
    main()
