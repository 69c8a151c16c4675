#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libnempc.so (cuobjdump -sass): the evidence that the tensor-core kernels issue tcgen05 MMAs
(UTCHMMA), read / write tensor memory (LDTM / STTM), stream weights with the TMA unit (UBLKCP = cp.async.bulk, UTMALDG = tensor-map
loads), that the register-resident kernel runs on packed FFMA2, and that the float64 wide-network path runs on the
FP64 tensor cores (DMMA) fed by cp.async (LDGSTS).  Kernels are grouped by family (template arguments dropped).

  python tools/sass_opcodes.py [path/to/libnempc.so] > profiles/<round>_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FFMA", "FMUL2", "DFMA", "DMMA", "LDGSTS", "HMMA", "MUFU", "LDCU", "LDS", "STS",
       "LDG", "STG", "BAR", "SHFL")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pyneuralempc_b200", "csrc", "libnempc.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    fam = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            base = re.sub(r"<.*", "", name.replace("void ", ""))
            cur = fam.setdefault(base, {"n": 0, "ops": collections.Counter(), "instr": 0})
            cur["n"] += 1
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            cur["instr"] += 1
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    cur["ops"][o] += 1
                    break
    print("# SASS opcode histogram of libnempc.so (sm_100a), per kernel family\n")
    print("`cuobjdump -sass` of the in-tree library; counts are static instructions summed over all instantiations of a family.\n")
    print("| kernel family | instantiations | instructions | " + " | ".join(OPS) + " |")
    print("|---|---|---|" + "---|" * len(OPS))
    for base, d in fam.items():
        print(f"| `{base}` | {d['n']} | {d['instr']} | " + " | ".join(str(d["ops"].get(o, 0)) for o in OPS) + " |")


if __name__ == "__main__":
    main()
