#!/usr/bin/env python
"""C2 kernel time of the library NEMPC_LIB_PATH points at (tools/fast_variants.sh)."""
import json
import os
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "50", "--warmup", "5", "--no-side-workloads", "--no-solver", "--no-cpu-baseline"],
                   capture_output=True, text=True)
d = json.loads(r.stdout.strip().splitlines()[-1])
print("%s: kernel %.4f ms frac %.4f value %.4g" % (os.path.basename(os.environ.get("NEMPC_LIB_PATH", "in-tree")), d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["value"]))
