"""Build recipe for ``oracle/_ref/``: the UNMODIFIED reference package compiled to bytecode.  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure Python, so "compiling it from the sources where they lie under /root/reference" (the rule for C / C++ references)
means ``py_compile``: every ``pyNeuralEMPC/**/*.py`` becomes a ``pyNeuralEMPC/**/*.pyc`` entry of the archive
``oracle/_ref/pyNeuralEMPC_bytecode.zip`` (sourceless layout; ``zipimport`` loads it once the archive is on ``sys.path`` -- an archive because
loose ``.pyc`` files are dropped by the snapshot that carries the repository to the GPU box).
No reference source is copied into the repository; ``oracle/_ref/`` is git-ignored and travels to the GPU box with the snapshot like the
built ``libnempc.so`` does.  ``oracle/shim.py`` imports the package from there when ``/root/reference`` does not exist, which lets
``bench.py --impl reference`` time the reference's OWN integrators and ``IpoptProblem`` on the box's host cores
(``cpu_baseline.kind = "reference"``; the network itself is still evaluated by ``oracle.mlp_np`` because TensorFlow is not installable).

  python -m oracle.build_ref          # also run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SOURCE_ROOT = "/root/reference"


ARCHIVE = "pyNeuralEMPC_bytecode.zip"


def build_reference_bytecode(source_root=SOURCE_ROOT, out=OUT):
    """returns the number of modules compiled (0 when the reference tree is absent: nothing is touched then)."""
    import tempfile
    import zipfile
    pkg = os.path.join(source_root, "pyNeuralEMPC")
    if not os.path.isdir(pkg):
        return 0
    os.makedirs(out, exist_ok=True)
    n = 0
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(os.path.join(out, ARCHIVE), "w", zipfile.ZIP_STORED) as zf:
        for root, _dirs, files in sorted(os.walk(pkg)):
            rel = os.path.relpath(root, pkg)
            for f in sorted(files):
                if f.endswith(".py"):
                    arc = os.path.normpath(os.path.join("pyNeuralEMPC", rel, f + "c"))          # module.py -> module.pyc (sourceless import layout)
                    dst = os.path.join(tmp, str(n) + ".pyc")
                    py_compile.compile(os.path.join(root, f), cfile=dst, dfile=os.path.normpath(os.path.join("pyNeuralEMPC", rel, f)), doraise=True)
                    zf.write(dst, arc)
                    n += 1
    with open(os.path.join(out, "BUILD_INFO.txt"), "w") as fh:
        fh.write(f"bytecode of the unmodified reference package, compiled from {pkg} by oracle/build_ref.py with Python "
                 f"{sys.version_info.major}.{sys.version_info.minor}; {n} modules in {ARCHIVE}\n")
    return n


if __name__ == "__main__":
    print(f"compiled {build_reference_bytecode()} reference modules into {OUT}")
