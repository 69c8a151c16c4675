"""bench.py contract checks that need no GPU: the reference arm runs the CPU restatement on the host cores and prints ONE JSON line
with the keys the driver reads; without a CUDA device the GPU arm refuses loudly instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-budget", "2"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "jac_hess_nlp_eval_horizon_steps_per_s" and d["unit"] == "horizon-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    # "reference" = the unmodified reference's integrator + IpoptProblem (sources here, bytecode under oracle/_ref on the GPU box); "port"
    # only where neither is present
    from oracle import shim
    assert cb["kind"] == ("reference" if shim.reference_available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "problems per step" in cb["sample"]
    assert d["config"]["workload"].startswith("C2")


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
