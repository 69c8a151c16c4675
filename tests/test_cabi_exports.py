"""The C-ABI library loads and exports every symbol include/nempc.h declares; the host-only entry points
(structure generation) work without a GPU; compute entry points fail loudly without one.  CPU only."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "nempc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nempc_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from pyneuralempc_b200.build import build_library
    build_library()
    from pyneuralempc_b200 import _lib
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from pyneuralempc_b200 import _lib
    names = _declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"libnempc.so does not export {n}"
    assert sorted(_lib.EXPORTS) == names
    assert b"sm_100a" in lib.nempc_version()


def test_desc_struct_matches_header():
    from pyneuralempc_b200._lib import NempcDesc
    # 4 int32 + 8 int32 widths + 2 int32 + (pad) double + 6 int32
    assert ctypes.sizeof(NempcDesc) == 4 * 4 + 8 * 4 + 2 * 4 + 8 + 6 * 4
    assert NempcDesc.tvp_dim.offset == NempcDesc.kernel.offset + 4 and NempcDesc.p_dim.offset == NempcDesc.kernel.offset + 8
    assert NempcDesc.dt.offset % 8 == 0


@pytest.mark.parametrize("H,x,u", [(25, 2, 1), (1, 2, 1), (7, 4, 1), (3, 12, 4), (2, 1, 1), (200, 12, 4)])
def test_structure_generator_matches_oracle(lib, H, x, u):
    from oracle import structure as S
    from pyneuralempc_b200.structure import nlp_structure
    n = H * (x + u)
    for mask in (None, np.ones(n), np.arange(n) % 2, np.r_[np.zeros(n - 1), 1.0]):
        jr, jc, hr, hc = nlp_structure(H, x, u, mask)
        r, c = S.jacobian_structure(H, x, u)
        np.testing.assert_array_equal(jr, r); np.testing.assert_array_equal(jc, c)
        r, c = S.hessian_structure(H, x, u, mask)
        np.testing.assert_array_equal(hr, r); np.testing.assert_array_equal(hc, c)
    assert len(jr) == S.nnz_jacobian(H, x, u)
    assert len(nlp_structure(H, x, u)[2]) == S.nnz_hessian_integrator(H, x, u)


def test_structure_matches_reference_golden(golden_dir):
    from pyneuralempc_b200.structure import nlp_structure
    for kind in ("discrete", "unity", "rk4"):
        g = np.load(os.path.join(golden_dir, f"ref_{kind}_H25.npz"))
        jr, jc, hr, hc = nlp_structure(25, 2, 1, g["obj_quad"])
        np.testing.assert_array_equal(hr, g["hes_rows"]); np.testing.assert_array_equal(hc, g["hes_cols"])
        r, c = np.nonzero(g["jacobian"])
        np.testing.assert_array_equal(jr, r); np.testing.assert_array_equal(jc, c)


def test_bad_dims_are_rejected(lib):
    from pyneuralempc_b200._lib import NempcError
    from pyneuralempc_b200.structure import nlp_structure
    with pytest.raises(NempcError):
        nlp_structure(5, 12, 8)            # x+u > NEMPC_MAX_D
    with pytest.raises(NempcError):
        nlp_structure(0, 2, 1)


def test_no_cpu_fallback(lib):
    """without a usable CUDA device every compute path raises instead of silently computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pyneuralempc_b200 import NlpEvaluator
    from pyneuralempc_b200._lib import NempcError
    from pyneuralempc_b200.engine import measure_fma_peak
    w = [(np.zeros((3, 4)), np.zeros(4)), (np.zeros((4, 2)), np.zeros(2))]
    with pytest.raises(NempcError, match="no CPU path"):
        NlpEvaluator(w, 2, 1, 5)
    with pytest.raises(NempcError):
        measure_fma_peak(0)


def test_product_never_imports_oracle():
    """the product package neither imports nor includes the oracle or the test-only host emulation (comments that cite
    them as documentation are fine)."""
    pkg = os.path.join(ROOT, "pyneuralempc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                for line in open(path):
                    code = line.split("#")[0]
                    assert not re.search(r"\b(import|from)\s+(oracle|hostsim_util|tests)\b", code), f"{f}: {line.strip()}"
                    assert "libnempc_hostsim" not in code, f
            elif f.endswith((".cu", ".cuh", ".h", ".cpp")):
                for line in open(path):
                    if line.lstrip().startswith("#include"):
                        assert "oracle" not in line and "hostsim" not in line and "tests/" not in line, f"{f}: {line.strip()}"
