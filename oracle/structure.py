"""Analytic sparsity of the transcription NLP in the reference's ordering.
TEST INFRASTRUCTURE ONLY.

Decision vector ``z = [x_1..x_H | u_0..u_{H-1}]`` (``optimizer/ipopt.py:20-28``); constraint
row ``r = t*x_dim + p`` (``discret.py:30``).

Jacobian pattern = non-zeros of the dense matrix the reference returns
(``discret.py:38-56``, ``rk4.py:120-176``): ``-1`` at ``(r, t*x_dim+p)``, the full
``x_dim x x_dim`` block on state block ``t-1`` for ``t >= 1`` and the full ``x_dim x u_dim``
block on control block ``t``; enumerated row-major like ``np.nonzero``.

Hessian pattern = ``np.nonzero(np.tril(objective_map + integrator_map))``
(``optimizer/ipopt.py:55-62``) where the integrator map couples, for every step ``t``,
``{state block t-1 (t >= 1), control block t}`` (``discret.py:70-78``, ``rk4.py:270-283``).
"""
from __future__ import annotations

import numpy as np


def jacobian_structure(H, x_dim, u_dim):
    rows, cols = [], []
    for t in range(H):
        for p in range(x_dim):
            r = t * x_dim + p
            if t > 0:
                for q in range(x_dim):
                    rows.append(r); cols.append((t - 1) * x_dim + q)
            rows.append(r); cols.append(r)
            for q in range(u_dim):
                rows.append(r); cols.append(H * x_dim + t * u_dim + q)
    return np.asarray(rows, np.int32), np.asarray(cols, np.int32)


def integrator_hessian_map(H, x_dim, u_dim):
    n = H * (x_dim + u_dim)
    m = np.zeros((n, n))
    for t in range(H):
        cu = slice(H * x_dim + t * u_dim, H * x_dim + (t + 1) * u_dim)
        m[cu, cu] = 1.0
        if t > 0:
            cx = slice((t - 1) * x_dim, t * x_dim)
            m[cx, cx] = 1.0
            m[cx, cu] = 1.0
            m[cu, cx] = 1.0
    return m


def hessian_structure(H, x_dim, u_dim, objective_diag_mask=None):
    """rows, cols (int32) of the lower-triangular Lagrangian-Hessian pattern."""
    m = integrator_hessian_map(H, x_dim, u_dim)
    if objective_diag_mask is not None:
        m = m + np.diag(np.asarray(objective_diag_mask, np.float64))
    r, c = np.nonzero(np.tril(m.astype(bool)))
    return r.astype(np.int32), c.astype(np.int32)


def nnz_jacobian(H, x_dim, u_dim):
    return H * x_dim * (1 + u_dim) + (H - 1) * x_dim * x_dim


def nnz_hessian_integrator(H, x_dim, u_dim):
    return (H - 1) * (x_dim * (x_dim + 1) // 2 + x_dim * u_dim) + H * (u_dim * (u_dim + 1) // 2)
