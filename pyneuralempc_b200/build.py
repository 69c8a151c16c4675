"""In-tree build of libnempc.so (hand-written sm_100a CUDA + the C ABI of include/nempc.h) with nvcc."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_PATH = os.environ.get("NEMPC_LIB_PATH") or os.path.join(CSRC, "libnempc.so")   # env override: kernel-variant experiments
SOURCES = ["nempc_lib.cu", "nempc_tc_tu.cu", "nempc_wide_tu.cu", "nempc_wide_rt_tu.cu", "nempc_fast64_tu.cu", "nempc_dmma_tu.cu"]
# per-source extra flags (see the header of nempc_fast64_tu.cu)
SOURCE_FLAGS = {"nempc_fast64_tu.cu": ["--split-compile=0"]}
HEADERS = ["nempc_generic.cuh", "nempc_fast.cuh", "nempc_fast64.cuh", "nempc_small.cuh", "nempc_tc.cuh", "nempc_wide.cuh", "nempc_wide_launch.cuh", "nempc_dmma.cuh", "nempc_rolling.cuh", "nempc_tc_ptx.cuh", "nempc_solver.cuh", "nempc_layout.h", os.path.join("..", "..", "include", "nempc.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


# NEMPC_BUILD_SPLIT=1: let nvcc optimise the kernels of the one translation unit in parallel (1.5 instead of 4 minutes on 8 cores).  For
# development only: the code it produces measured 8 % slower on nempc_wide_kernel and 2 % slower on nempc_fast_kernel (B200, round 2).
if os.environ.get("NEMPC_BUILD_SPLIT"):
    NVCC_FLAGS = NVCC_FLAGS + ["--split-compile=0"]
if os.environ.get("NEMPC_NVCC_EXTRA"):                 # kernel-variant experiments, e.g. "-DNEMPC_FAST64_SMEM_WEIGHTS=1"
    NVCC_FLAGS = NVCC_FLAGS + os.environ["NEMPC_NVCC_EXTRA"].split()


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libnempc.so cannot be built (there is no CPU fallback)")


def source_hash():
    """sha256 over the CUDA sources and headers the library is built from; compiled into the library (nempc_source_hash) and written
    next to it, so that staleness does not depend on file times (which a copy to another machine does not preserve)"""
    import hashlib
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read() + b"\0")
    return h.hexdigest()[:32]


def is_stale():
    if os.environ.get("NEMPC_LIB_PATH"):
        return False
    if not os.path.exists(LIB_PATH) or not os.path.exists(LIB_PATH + ".srchash"):
        return True
    with open(LIB_PATH + ".srchash") as fh:
        return fh.read().strip() != source_hash()


def build_library(force=False, verbose=False):
    """compile pyneuralempc_b200/csrc/libnempc.so for sm_100a; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    sh = source_hash()
    nvcc = find_nvcc()
    import tempfile
    objdir = tempfile.mkdtemp(prefix="nempc_build_")
    common = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + [f'-DNEMPC_SOURCE_HASH="{sh}"']
    procs = []
    for src in SOURCES:                                    # the translation units compile side by side, then one link
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + common + [fl for fl in SOURCE_FLAGS.get(src, []) if fl not in common] + ["-c", "-o", obj, os.path.join(CSRC, src)]
        procs.append((obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    logs = []
    for obj, pr in procs:
        out, _ = pr.communicate()
        logs.append(out)
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + out)
    link = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH] + [o for o, _ in procs],
                          capture_output=True, text=True)
    if link.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + link.stdout + link.stderr)
    proc = type("R", (), {"stderr": "\n".join(logs)})()
    shutil.rmtree(objdir, ignore_errors=True)
    with open(LIB_PATH + ".srchash", "w") as fh:
        fh.write(sh + "\n")
    if verbose:
        print(proc.stderr)
    return LIB_PATH
