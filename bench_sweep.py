#!/usr/bin/env python
"""BASELINE.json configs[4]: sweep hidden width x horizon x batch, FP32 vs FP64, Jacobian+Hessian eval roofline.
Not the driver's benchmark (that is bench.py); writes a CSV for profiles/.  Depth 3, cart-pole dims
(x=4,u=1) for widths >= 64 and Lotka-Volterra dims (x=2,u=1) below; every row is one device-resident evaluation
(residual + Jacobian + Hessian kernel) timed with CUDA events over back-to-back launches."""
import argparse
import csv
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/sweep.csv")
    ap.add_argument("--max-seconds", type=float, default=1.5, help="skip rows whose single evaluation is expected to take longer (from the kernel's typical rate)")
    ap.add_argument("--max-gb", type=float, default=60.0, help="skip rows whose inputs + outputs exceed this many GB")
    ap.add_argument("--dtypes", default="float32,float64")
    ap.add_argument("--widths", default="30,32,64,128,256,512")
    args = ap.parse_args()
    import torch
    from oracle.mlp_np import MLP
    from pyneuralempc_b200 import NlpEvaluator
    from pyneuralempc_b200.engine import measure_fma_peak
    peak = {"float32": measure_fma_peak(0, "float32", 200), "float64": measure_fma_peak(0, "float64", 200)}
    rows = []
    rng = np.random.default_rng(0)
    for dtype in args.dtypes.split(","):
        for integ in ("discrete", "rk4"):
            for width in [int(w) for w in args.widths.split(",")]:
                xd, ud = (2, 1) if width < 64 else (4, 1)
                depth = 2 if width == 30 else 3
                dims = [xd + ud] + [width] * depth + [xd]
                mlp = MLP.glorot(dims, xd, ud, seed=0, dtype=np.float32)
                for H in (10, 50, 200, 500):
                    for B in (1, 256, 4096, 65536, 262144):
                        ev = NlpEvaluator(mlp.weights, xd, ud, H, integ, DT=0.1, compute_dtype=dtype, io_dtype="float64")
                        gflop = ev.flops_per_step * B * H / 1e9
                        kname = ("dmma" if "nempc_dmma" in ev.kernel_name else "fast64" if "fast64" in ev.kernel_name else "fast" if "nempc_fast" in ev.kernel_name else
                                 "wide" if "nempc_wide" in ev.kernel_name else "tc" if "tcgen05" in ev.kernel_name else "generic")
                        if (kname == "fast64" and B * H < 4096) or (kname == "dmma" and B * H < 512):
                            kname = "generic"                    # AUTO hands small float64 batches to the generic kernel
                        typical_tf = {"dmma": 8.0, "fast": 25.0, "fast64": 8.0, "wide": 90.0, "tc": 35.0, "generic": 2.0 if dtype == "float32" else 1.2}[kname]
                        gbytes = B * (ev.n + ev.m * 2 + ev.nnz_jac + ev.nnz_hes) * 8 / 1e9
                        if gflop / typical_tf / 1e3 > args.max_seconds or gbytes > args.max_gb:
                            ev.close(); continue
                        z = torch.as_tensor(rng.uniform(-1, 1, (B, ev.n))).cuda()
                        x0 = torch.as_tensor(rng.uniform(-1, 1, (B, xd))).cuda()
                        lam = torch.as_tensor(rng.standard_normal((B, ev.m))).cuda()
                        out = ev.alloc_outputs(B, ("resid", "jac", "hes"))
                        for _ in range(2):
                            ev.eval(z, x0, lam, 1.0, want=("resid", "jac", "hes"), out=out)
                        reps = int(max(2, min(50, 20.0 * typical_tf / 25.0 / max(gflop, 1e-3))))
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        torch.cuda.synchronize(); a.record()
                        for _ in range(reps):
                            ev.eval(z, x0, lam, 1.0, want=("resid", "jac", "hes"), out=out)
                        b.record(); torch.cuda.synchronize()
                        ms = a.elapsed_time(b) / reps
                        tf = gflop / ms
                        rows.append(dict(dtype=dtype, integrator=integ, width=width, depth=depth, x=xd, u=ud, H=H, B=B, steps=B * H,
                                         kernel=kname, ms=round(ms, 4),
                                         steps_per_s=round(B * H / ms * 1e3), tflops=round(tf, 3), frac_of_fma_peak=round(tf / peak[dtype], 4)))
                        print(rows[-1], flush=True)
                        ev.close()
    with open(args.out, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0]))
        w.writeheader(); w.writerows(rows)
    print("fma peaks TFLOP/s:", peak, file=sys.stderr)


if __name__ == "__main__":
    sys.path.insert(0, ".")
    main()
