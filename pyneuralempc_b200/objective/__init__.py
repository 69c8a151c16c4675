"""Objective layer: cost value / gradient / Hessian with the value and gradient evaluated on the GPU.

Mirrors ``/root/reference/pyNeuralEMPC/objective/base.py:4-18`` (``ObjectiveFunc``) and the call contract of
``objective/jax.py:28-65`` (``JAXObjectifFunc``): ``forward(states, u, p, tvp) -> scalar``, ``gradient -> (n,)``
ordered ``[d/dstates | d/du]``, ``hessian -> dense (n, n)``, ``hessianstructure(H, model) -> (n, n)`` 0/1.

The reference differentiates an arbitrary JAX function; the CUDA path implements the separable family

    f(z) = sum_i lin_i z_i + quad_i (z_i - ref_i)^2,       z = [states.ravel() | u.ravel()]

which covers every cost that ships with the reference (run.py:83-84 linear in u; test.py:59-60 squared set-point)
and diagonal tracking costs.  Its Hessian ``diag(2 quad)`` is constant, so the reference's numeric structure probing
(objective/jax.py:67-90) reduces to ``quad != 0``.
"""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _lib


class ObjectiveFunc:
    """Abstract cost (reference objective/base.py:4-18)."""

    def __init__(self):
        pass

    def forward(self, states, u, p=None, tvp=None):
        raise NotImplementedError("")

    def gradient(self, states, u, p=None, tvp=None):
        raise NotImplementedError("")

    def hessian(self, states, u, p=None, tvp=None):
        raise NotImplementedError("")

    def hessianstructure(self, H, model):
        raise NotImplementedError("")


class CudaSeparableObjective(ObjectiveFunc):
    def __init__(self, lin, quad, ref, device=0):
        super().__init__()
        self.lin = np.ascontiguousarray(lin, np.float64).ravel()
        self.quad = np.ascontiguousarray(quad, np.float64).ravel()
        self.ref = np.ascontiguousarray(ref, np.float64).ravel()
        if not (self.lin.shape == self.quad.shape == self.ref.shape):
            raise ValueError("lin, quad and ref must have the same length n = H*(x_dim+u_dim)")
        self.n = self.lin.shape[0]
        self.device = device
        self.offset = 0.0                          # constant term of the cost (objective/jax.py identifies one); not on the device
        self._dev = None

    def prepare(self, H, x_dim, u_dim, p=None, tvp=None):
        """hook for costs that are only known once the problem dimensions are (objective/jax.py); nothing to do here"""

    def _device_params(self):
        if self._dev is None:
            import torch
            dev = torch.device("cuda", self.device)
            self._dev = tuple(torch.as_tensor(a, dtype=torch.float64, device=dev) for a in (self.lin, self.quad, self.ref))
        return self._dev

    def _eval(self, states, u, want_grad):
        import torch
        z = np.concatenate([np.asarray(states, np.float64).reshape(-1), np.asarray(u, np.float64).reshape(-1)])
        if z.shape[0] != self.n:
            raise ValueError(f"objective built for n={self.n}, got {z.shape[0]} variables")
        lin, quad, ref = self._device_params()
        zd = torch.as_tensor(z, dtype=torch.float64, device=lin.device).reshape(1, -1)
        obj = torch.empty(1, dtype=torch.float64, device=lin.device)
        grad = torch.empty((1, self.n), dtype=torch.float64, device=lin.device) if want_grad else None
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        s = torch.cuda.current_stream(lin.device).cuda_stream
        _lib.check(_lib.load().nempc_objective_eval(_lib.F64, 1, self.n, p(zd), p(lin), p(quad), p(ref), p(obj), p(grad),
                                                    ctypes.c_void_p(s)), None, "nempc_objective_eval")
        return obj, grad

    def forward(self, states, u, p=None, tvp=None):
        return float(self._eval(states, u, False)[0].item())

    def gradient(self, states, u, p=None, tvp=None):
        g = self._eval(states, u, True)[1][0].cpu().numpy()
        return np.nan_to_num(g, nan=0.0)                      # objective/jax.py:40

    def hessian(self, states, u, p=None, tvp=None):
        return np.diag(2.0 * self.quad)

    def hessianstructure(self, H=None, model=None):
        return (np.diag(2.0 * self.quad) != 0.0).astype(np.float64)


def CudaLinearObjective(H, x_dim, u_dim, cost_vec, device=0):
    """``sum(u.ravel() * cost_vec)`` -- the shipped Lotka-Volterra cost (examples/lotka_volterra/run.py:79-87)."""
    n = H * (x_dim + u_dim)
    lin = np.zeros(n)
    lin[H * x_dim:] = np.broadcast_to(np.asarray(cost_vec, np.float64).ravel(), (H * u_dim,))
    return CudaSeparableObjective(lin, np.zeros(n), np.zeros(n), device)


def CudaSetpointObjective(H, x_dim, u_dim, target, device=0):
    """``sum((u - target)^2)`` -- the cost of the reference's test.py:55-60."""
    n = H * (x_dim + u_dim)
    quad, ref = np.zeros(n), np.zeros(n)
    quad[H * x_dim:] = 1.0
    ref[H * x_dim:] = target
    return CudaSeparableObjective(np.zeros(n), quad, ref, device)


def CudaQuadraticObjective(H, x_dim, u_dim, q_diag, r_diag, x_ref=None, u_ref=None, device=0):
    """diagonal tracking cost ``sum_t (x_t-xr_t)' Q (x_t-xr_t) + (u_t-ur_t)' R (u_t-ur_t)``."""
    q = np.tile(np.asarray(q_diag, np.float64), H)
    r = np.tile(np.asarray(r_diag, np.float64), H)
    xr = np.zeros((H, x_dim)) if x_ref is None else np.broadcast_to(x_ref, (H, x_dim))
    ur = np.zeros((H, u_dim)) if u_ref is None else np.broadcast_to(u_ref, (H, u_dim))
    return CudaSeparableObjective(np.zeros(H * (x_dim + u_dim)), np.concatenate([q, r]),
                                  np.concatenate([np.ravel(xr), np.ravel(ur)]), device)


class CudaQuadraticFormObjective(ObjectiveFunc):
    """General quadratic cost ``f(z) = 1/2 z' P z + q' z + c`` with a sparse symmetric ``P`` -- the non-separable costs an arbitrary
    ``JAXObjectifFunc`` function (objective/jax.py:28-57) typically is: control-rate penalties, full-matrix stage / terminal weights,
    state-control cross terms.  Value and gradient come from ``nempc_quadform_eval`` (one warp per problem, CSR rows of ``P``); the
    constant Hessian ``P`` joins the Lagrangian Hessian on the union pattern through ``nempc_hessian_merge`` (optimizer/ipopt.py).
    ``z = [states.ravel() | u.ravel()]`` as everywhere (ipopt.py:20-28)."""

    def __init__(self, P, q=None, c=0.0, device=0):
        super().__init__()
        from scipy import sparse
        P = sparse.csr_matrix(P, dtype=np.float64)
        if P.shape[0] != P.shape[1]:
            raise ValueError("P must be square (n x n), n = H*(x_dim+u_dim)")
        if abs(P - P.T).max() > 1e-12 * max(1.0, abs(P).max()):
            raise ValueError("P must be symmetric")
        P = ((P + P.T) * 0.5).tocsr()
        P.eliminate_zeros(); P.sort_indices()
        self.P = P
        self.n = P.shape[0]
        self.q = np.zeros(self.n) if q is None else np.ascontiguousarray(q, np.float64).ravel()
        if self.q.shape != (self.n,):
            raise ValueError("q must have length n")
        self.c, self.offset = float(c), 0.0
        self.device = device
        self._dev = None

    @classmethod
    def from_blocks(cls, H, x_dim, u_dim, Q=None, R=None, Qf=None, S=None, x_ref=None, u_ref=None, N=None, lin=None, device=0):
        """``sum_{t=1..H} (x_t - xr_t)' Q (x_t - xr_t)``  (``Qf`` instead of ``Q`` on the terminal state x_H when given)
        ``+ sum_{t=0..H-1} (u_t - ur_t)' R (u_t - ur_t)  +  sum_{t=1..H-1} (u_t - u_{t-1})' S (u_t - u_{t-1})``
        ``+ sum_{t=1..H-1} 2 (x_t - xr_t)' N (u_t - ur_t)  +  lin' z``;  ``Q, Qf`` (x, x), ``R, S`` (u, u), ``N`` (x, u) full matrices."""
        from scipy import sparse
        n = H * (x_dim + u_dim)
        P = sparse.lil_matrix((n, n))
        zr = np.zeros(n)
        if x_ref is not None:
            zr[:H * x_dim] = np.broadcast_to(np.asarray(x_ref, np.float64), (H, x_dim)).ravel()
        if u_ref is not None:
            zr[H * x_dim:] = np.broadcast_to(np.asarray(u_ref, np.float64), (H, u_dim)).ravel()
        sx = lambda t: slice((t - 1) * x_dim, t * x_dim)                    # x_t, t = 1..H
        su = lambda t: slice(H * x_dim + t * u_dim, H * x_dim + (t + 1) * u_dim)
        sym = lambda M: 0.5 * (np.asarray(M, np.float64) + np.asarray(M, np.float64).T)
        for t in range(1, H + 1):
            W = Qf if (t == H and Qf is not None) else Q
            if W is not None:
                P[sx(t), sx(t)] += 2.0 * sym(W)
        for t in range(H):
            if R is not None:
                P[su(t), su(t)] += 2.0 * sym(R)
            if S is not None and t >= 1:
                Ss = 2.0 * sym(S)
                P[su(t), su(t)] += Ss; P[su(t - 1), su(t - 1)] += Ss
                P[su(t), su(t - 1)] -= Ss; P[su(t - 1), su(t)] -= Ss
            if N is not None and 1 <= t <= H - 1:
                Nm = 2.0 * np.asarray(N, np.float64)
                P[sx(t), su(t)] += Nm; P[su(t), sx(t)] += Nm.T
        P = P.tocsr()
        q = -(P @ zr)
        c = 0.5 * float(zr @ (P @ zr))
        if lin is not None:
            q = q + np.asarray(lin, np.float64).ravel()
        return cls(P, q, c, device)

    def prepare(self, H, x_dim, u_dim, p=None, tvp=None):
        if H * (x_dim + u_dim) != self.n:
            raise ValueError(f"objective built for n={self.n}, the problem has {H * (x_dim + u_dim)} variables")

    def _device_tables(self):
        if self._dev is None:
            import torch
            dev = torch.device("cuda", self.device)
            t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
            self._dev = (t(self.P.indptr, torch.int32), t(self.P.indices, torch.int32), t(self.P.data, torch.float64), t(self.q, torch.float64))
        return self._dev

    def eval_device(self, z, want_grad=True):
        """``z``: CUDA tensor (B, n) float64 -> (obj (B,), grad (B, n) or None) on the device"""
        import torch
        ptr, idx, val, q = self._device_tables()
        z = z.to(torch.float64).contiguous()
        B = z.shape[0]
        obj = torch.empty(B, dtype=torch.float64, device=z.device)
        grad = torch.empty((B, self.n), dtype=torch.float64, device=z.device) if want_grad else None
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        s = torch.cuda.current_stream(z.device).cuda_stream
        _lib.check(_lib.load().nempc_quadform_eval(_lib.F64, B, self.n, p(z), p(ptr), p(idx), p(val), p(q), self.c, p(obj), p(grad),
                                                   ctypes.c_void_p(s)), None, "nempc_quadform_eval")
        return obj, grad

    def _eval(self, states, u, want_grad):
        import torch
        z = np.concatenate([np.asarray(states, np.float64).reshape(-1), np.asarray(u, np.float64).reshape(-1)])
        if z.shape[0] != self.n:
            raise ValueError(f"objective built for n={self.n}, got {z.shape[0]} variables")
        return self.eval_device(torch.as_tensor(z, device=torch.device("cuda", self.device)).reshape(1, -1), want_grad)

    def forward(self, states, u, p=None, tvp=None):
        return float(self._eval(states, u, False)[0].item())

    def gradient(self, states, u, p=None, tvp=None):
        return np.nan_to_num(self._eval(states, u, True)[1][0].cpu().numpy(), nan=0.0)      # objective/jax.py:40

    def hessian(self, states, u, p=None, tvp=None):
        return self.P.toarray()

    def hessianstructure(self, H=None, model=None):
        return (self.P.toarray() != 0.0).astype(np.float64)

    # ---- the Lagrangian-Hessian merge (ipopt.py:55-62, 66-86) ------------------------------------------------------------------
    def merge_tables(self, hes_rows, hes_cols):
        """union of the constraint pattern (rows / cols of the evaluator WITHOUT a device objective) and tril(P):
        -> (rows, cols, src_slot, p_val), row-major sorted like np.nonzero(np.tril(objective_map + integrator_map))"""
        n = self.n
        Pl = self.P.tocoo()
        keep = Pl.row >= Pl.col
        key_k = np.asarray(hes_rows, np.int64) * n + np.asarray(hes_cols, np.int64)
        key_p = Pl.row[keep].astype(np.int64) * n + Pl.col[keep].astype(np.int64)
        keys = np.union1d(key_k, key_p)
        src = np.full(len(keys), -1, np.int32)
        src[np.searchsorted(keys, key_k)] = np.arange(len(key_k), dtype=np.int32)
        pval = np.zeros(len(keys))
        pval[np.searchsorted(keys, key_p)] = Pl.data[keep]
        return (keys // n).astype(np.int32), (keys % n).astype(np.int32), src, pval

    def merge_hessian(self, kern_vals, src_slot, p_val, sigma):
        """device: ``out[b, s] = kern_vals[b, src_slot[s]] + sigma_b * p_val[s]``; ``sigma`` a float or a CUDA tensor (B,)"""
        import torch
        B, nk = kern_vals.shape
        out = torch.empty((B, src_slot.shape[0]), dtype=kern_vals.dtype, device=kern_vals.device)
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        sig_t = sigma.to(kern_vals.dtype).contiguous() if torch.is_tensor(sigma) else None
        io = _lib.F64 if kern_vals.dtype == torch.float64 else _lib.F32
        s = torch.cuda.current_stream(kern_vals.device).cuda_stream
        _lib.check(_lib.load().nempc_hessian_merge(io, B, nk, src_slot.shape[0], p(kern_vals.contiguous()), p(src_slot), p(p_val), p(sig_t),
                                                   1.0 if sig_t is not None else float(sigma), p(out), ctypes.c_void_p(s)), None, "nempc_hessian_merge")
        return out
