"""Per-step block formulation of the integrator constraints (O(H) memory, batched).
TEST INFRASTRUCTURE ONLY.

Same mathematics as ``dense_ref`` (and therefore as ``integrator/discret.py``,
``integrator/unity.py``, ``integrator/rk4.py`` of the reference) but without the dense
``(m, n, n)`` scatter: each horizon step only touches ``(x_{t-1}, u_t, x_t)``, so the
oracle evaluates per-step blocks for all ``B*H`` steps at once and assembles the values in
the sparsity order of ``oracle.structure``.  ``tests/test_oracle_consistency.py`` proves it
equal to ``dense_ref`` (and the golden files prove ``dense_ref`` equal to the reference).

Per step, ``z = (x_{t-1}, u_t)``, ``d = x_dim + u_dim``, ``E = [I; 0]``  (SURVEY 7.3):

    stage s:  z_s = z + a_s E k_{s-1},  k_s = f(z_s),  J_s = df/dz(z_s),  Hf_p = d2 f_p/dz2(z_s)
              dk_s = J_s R_s,  R_0 = I,  R_{s+1} = I + a_{s+1} E dk_s          (rk4.py:143-157)
              h_s[p] = R_s^T Hf_p R_s + a_s sum_{k<x} J_s[p,k] h_{s-1}[k]       (rk4.py:246-263)
    pred  = sum_s c_s k_s,   [A|B] = sum_s c_s dk_s,   Hblk[p] = sum_s c_s h_s[p]

with ``a = (0, DT/2, DT/2, DT)``, ``c = DT/6 (1,2,2,1)`` for RK4 and a single stage
``a = 0, c = 1`` for the discrete / unity integrators.
"""
from __future__ import annotations

import numpy as np

from . import structure as S


def stage_tables(kind, DT):
    if kind == "rk4":
        return (0.0, DT / 2.0, DT / 2.0, DT), (DT / 6.0, DT / 3.0, DT / 3.0, DT / 6.0)
    if kind in ("discrete", "unity", "model"):
        return (0.0,), (1.0,)
    raise ValueError(kind)


def step_blocks(mlp, kind, DT, z, need_jac=True, need_hes=True):
    """``z`` (N, d) -> pred (N,x), AB (N,x,d) or None, Hblk (N,x,d,d) or None.

    ``pred`` excludes the ``+ x_{t-1}`` / ``- x_t`` terms and ``AB`` excludes the ``+I`` of the
    discrete / RK4 integrators (added at assembly).  ``kind='model'`` returns the raw model
    value / Jacobian / per-output Hessian (``Model.forward/jacobian/hessian`` per sample)."""
    z = np.asarray(z, np.float64)
    N, d = z.shape
    xd = mlp.x_dim
    a_tab, c_tab = stage_tables(kind, DT)
    need_jac = need_jac or need_hes
    pred = np.zeros((N, xd))
    AB = np.zeros((N, xd, d)) if need_jac else None
    Hb = np.zeros((N, xd, d, d)) if need_hes else None
    R = np.broadcast_to(np.eye(d), (N, d, d)).copy()
    k_prev = np.zeros((N, xd))
    h_prev = None
    for s, (a_s, c_s) in enumerate(zip(a_tab, c_tab)):
        zs = z.copy()
        zs[:, :xd] += a_s * k_prev
        # the network sees (and rounds to) its own dtype, like model.predict / tf.constant(float32)
        f, J, Hf = mlp.blocks(zs.astype(mlp.dtype), need_hessian=need_hes) if need_jac else \
            (mlp.forward_z(zs.astype(mlp.dtype)), None, None)
        f = f.astype(np.float64)
        pred += c_s * f
        if need_jac:
            J = J.astype(np.float64)
            dk = J @ R
            AB += c_s * dk
        if need_hes:
            Hf = Hf.astype(np.float64)
            h = np.einsum("nkj,npkl,nlm->npjm", R, Hf, R)
            if s > 0:
                h += a_s * np.einsum("npk,nkab->npab", J[:, :, :xd], h_prev)
            Hb += c_s * h
            h_prev = h
        if need_jac and s + 1 < len(a_tab):
            R = np.broadcast_to(np.eye(d), (N, d, d)).copy()
            R[:, :xd, :] += a_tab[s + 1] * dk
        k_prev = f
    return pred, AB, Hb


class BlockEvaluator:
    """Batched residual / sparse Jacobian values / sparse Lagrangian-Hessian values /
    objective for ``B`` independent problems -- the CPU statement of what ``nempc_eval``
    computes on the device (layouts documented in ``include/nempc.h``)."""

    def __init__(self, mlp, kind, H, DT=None, objective=None):
        self.mlp, self.kind, self.H, self.DT = mlp, kind, int(H), DT
        self.xd, self.ud = mlp.x_dim, mlp.u_dim
        self.d = self.xd + self.ud
        self.n = self.H * self.d
        self.m = self.H * self.xd
        self.objective = objective
        quad = None if objective is None else objective.quad
        self.jac_rows, self.jac_cols = S.jacobian_structure(self.H, self.xd, self.ud)
        self.hes_rows, self.hes_cols = S.hessian_structure(
            self.H, self.xd, self.ud, None if quad is None else (quad != 0.0))
        self._build_maps()

    def _gidx(self, t, a):
        """global variable index of local coordinate ``a`` of step ``t`` (or -1: x0 is data)."""
        if a < self.xd:
            return (t - 1) * self.xd + a if t > 0 else -1
        return self.H * self.xd + t * self.ud + (a - self.xd)

    def _build_maps(self):
        H, xd, d = self.H, self.xd, self.d
        jpos = {(int(r), int(c)): k for k, (r, c) in enumerate(zip(self.jac_rows, self.jac_cols))}
        hpos = {(int(r), int(c)): k for k, (r, c) in enumerate(zip(self.hes_rows, self.hes_cols))}
        self.jmap = -np.ones((H, xd, d), np.int64)       # where AB[t,p,c] lands
        self.jdiag = np.zeros((H, xd), np.int64)         # where the -1 lands
        self.hmap = -np.ones((H, d, d), np.int64)        # where Hc[t,a,b] lands (lower triangle)
        for t in range(H):
            for p in range(xd):
                r = t * xd + p
                self.jdiag[t, p] = jpos[(r, r)]
                for c in range(d):
                    g = self._gidx(t, c)
                    if g >= 0:
                        self.jmap[t, p, c] = jpos[(r, g)]
            for a in range(d):
                ga = self._gidx(t, a)
                for b in range(d):
                    gb = self._gidx(t, b)
                    if ga >= 0 and gb >= 0 and ga >= gb:
                        self.hmap[t, a, b] = hpos[(ga, gb)]
        self.hdiag = np.asarray([hpos.get((i, i), -1) for i in range(self.n)], np.int64)

    def split(self, Z, X0):
        B = Z.shape[0]
        H, xd, ud = self.H, self.xd, self.ud
        states = Z[:, :H * xd].reshape(B, H, xd)
        u = Z[:, H * xd:].reshape(B, H, ud)
        xprev = np.concatenate([X0.reshape(B, 1, xd), states[:, :-1]], axis=1)
        return states, u, xprev

    def evaluate(self, Z, X0, lam=None, obj_factor=1.0, need_jac=True, need_hes=True):
        Z = np.asarray(Z, np.float64)
        X0 = np.asarray(X0, np.float64)
        B = Z.shape[0]
        H, xd, d = self.H, self.xd, self.d
        states, u, xprev = self.split(Z, X0)
        zz = np.concatenate([xprev, u], axis=2).reshape(B * H, d)
        need_hes = need_hes and lam is not None
        pred, AB, Hb = step_blocks(self.mlp, self.kind, self.DT, zz, need_jac, need_hes)
        out = {}
        pred = pred.reshape(B, H, xd)
        base = 0.0 if self.kind == "unity" else xprev
        out["resid"] = (base + pred - states).reshape(B, self.m)
        if need_jac:
            AB = AB.reshape(B, H, xd, d).copy()
            if self.kind != "unity":
                AB[:, :, np.arange(xd), np.arange(xd)] += 1.0
            jv = np.zeros((B, len(self.jac_rows)))
            ok = self.jmap >= 0
            jv[:, self.jmap[ok]] = AB[:, ok]
            jv[:, self.jdiag.ravel()] = -1.0
            out["jac_vals"] = jv
        if need_hes:
            lam = np.asarray(lam, np.float64).reshape(B, H, xd)
            Hc = np.einsum("ntp,ntpab->ntab", lam, Hb.reshape(B, H, xd, d, d))
            hv = np.zeros((B, len(self.hes_rows)))
            ok = self.hmap >= 0
            hv[:, self.hmap[ok]] = Hc[:, ok]
            if self.objective is not None:
                sig = np.broadcast_to(np.asarray(obj_factor, np.float64), (B,))
                dd = self.hdiag >= 0
                hv[:, self.hdiag[dd]] += sig[:, None] * (2.0 * self.objective.quad[dd])[None, :]
            out["hes_vals"] = hv
        if self.objective is not None:
            o = self.objective
            dz = Z - o.ref[None, :]
            out["obj"] = np.sum(o.lin[None, :] * Z + o.quad[None, :] * dz * dz, axis=1)
            out["grad"] = o.lin[None, :] + 2.0 * o.quad[None, :] * dz
        return out
