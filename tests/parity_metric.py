"""Elementwise parity metric shared by the GPU tests and ``__graft_entry__.smoke()``.

An output array passes at relative tolerance ``rtol`` when EVERY element satisfies

    |got - ref| <= rtol * |ref| + 0.1 * rtol * blockmax,        blockmax = max |ref| over the array

(for rtol = 1e-5: 1e-5 relative plus an absolute floor of 1e-6 of the array's largest magnitude -- the floor is what float32
cancellation in a chain-rule sum can leave on an entry that is itself close to zero; BASELINE.json north_star: "values within 1e-5
relative (FP32) or 1e-10 (FP64)").  An array-max-normalised error would let an entry 100x below the maximum be 1e-3 wrong.
``elem_err`` returns the worst ratio  |d| / (|ref| + 0.1 blockmax)  so that call sites keep the form ``elem_err(got, ref) < rtol``;
``worst_element`` names the offending element for the assertion message."""
import numpy as np


def _den(ref):
    ref = np.asarray(ref, dtype=np.float64)
    bm = float(np.abs(ref).max()) if ref.size else 0.0
    return np.abs(ref) + 0.1 * (bm if bm > 0.0 else 1.0)


def elem_err(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    return float((np.abs(got - ref) / _den(ref)).max())


def worst_element(got, ref):
    """(flat index, got, ref, ratio) of the element with the largest |d| / (|ref| + 0.1 blockmax)."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return (-1, 0.0, 0.0, 0.0)
    r = np.abs(got - ref) / _den(ref)
    i = int(np.argmax(r))
    return (i, float(got.flat[i]), float(ref.flat[i]), float(r.flat[i]))


def assert_close(got, ref, rtol, what=""):
    e = elem_err(got, ref)
    if not e < rtol:
        i, g, r, ratio = worst_element(got, ref)
        raise AssertionError(f"{what}: worst element #{i}: got {g!r}, reference {r!r}, |d|/(|ref| + 0.1 max|ref|) = {ratio:.3e} >= {rtol:g}")
