#!/usr/bin/env python
"""one small tensor-core-kernel evaluation checked against the oracle (used under compute-sanitizer: smallest case that runs
every phase of nempc_tc_kernel: several tiles per group, ragged last tile, RK4, all three modes)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from oracle.blocks_np import BlockEvaluator
    from oracle.mlp_np import MLP
    from oracle.objectives_np import SeparableQuadraticObjective
    from pyneuralempc_b200 import NlpEvaluator
    H, B = 7, 5
    rng = np.random.default_rng(0)
    mlp = MLP.glorot([5, 128, 128, 128, 4], 4, 1, seed=1)
    obj = SeparableQuadraticObjective.tracking(H, 4, 1, [1.0, 2.0, 0.5, 1.5], [0.1])
    Z, X0, lam = rng.uniform(-1, 1, (B, H * 5)), rng.uniform(-1, 1, (B, 4)), rng.standard_normal((B, H * 4))
    ref = BlockEvaluator(mlp, "rk4", H, DT=0.1, objective=obj).evaluate(Z, X0, lam, 1.0)
    ev = NlpEvaluator(mlp.weights, 4, 1, H, "rk4", DT=0.1, kernel="tc")
    ev.set_objective(obj.lin, obj.quad, obj.ref)
    t = lambda a: torch.as_tensor(a).cuda()
    out = ev.eval(t(Z), t(X0), t(lam), 1.0)
    o1 = ev.eval(t(Z), t(X0), want=("resid", "jac"))
    o0 = ev.eval(t(Z), t(X0), want=("resid",))
    torch.cuda.synchronize()
    for kr, kg in (("resid", "resid"), ("jac_vals", "jac"), ("hes_vals", "hes")):
        err = np.abs(out[kg].cpu().numpy() - ref[kr]).max() / max(1.0, np.abs(ref[kr]).max())
        assert err < 1e-5, (kg, err)
    assert np.abs(o1["jac"].cpu().numpy() - ref["jac_vals"]).max() < 1e-5 and np.abs(o0["resid"].cpu().numpy() - ref["resid"]).max() < 1e-5
    ev.close()
    print("tc small case OK")


if __name__ == "__main__":
    main()
