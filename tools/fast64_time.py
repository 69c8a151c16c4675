#!/usr/bin/env python
"""C2 in float64 on the library NEMPC_LIB_PATH points at (tools/fast64_variants.sh): evaluation time and fraction of the FP64 FMA peak."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["bench"]
import bench
r = bench.c2_float64(1, 0, 0)
print("%s: C2_f64 %.4g steps/s frac %.3f %.4f ms" % (os.path.basename(os.environ.get("NEMPC_LIB_PATH", "in-tree")), r["value"], r["roofline"]["frac"], r["ms_per_eval"]))
