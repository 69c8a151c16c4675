"""Batched primal-dual interior-point NMPC solver -- numpy statement (CPU).  TEST INFRASTRUCTURE ONLY.

This is the oracle of ``nempc_solve`` (the on-device solver loop).  It is NOT a restatement of the
reference: the reference hands the NLP to IPOPT (``optimizer/ipopt.py:162-189``) or SciPy SLSQP
(``optimizer/slsqp.py:172``), both third-party and sequential.  What it shares with them is the problem:

    min f(z)   s.t.  c(z) = 0  (transcription constraints),   lb <= z <= ub  (DomainConstraint, constraints.py:26-30)

Algorithm (IPOPT-style, simplified): barrier problem for mu -> 0; Newton step on the primal-dual equations with the
bound duals eliminated,

    [ W + Sigma   J^T ] [ dz  ]     [ grad f - mu/(z-lb) + mu/(ub-z) + J^T lambda ]
    [ J           0   ] [ dlam] = - [ c                                           ]

solved EXACTLY by a Riccati recursion because the KKT matrix of the transcription is block tridiagonal in the
horizon (stage k couples X_k = x_{k-1}, U_k = u_k and X_{k+1}); stage-wise regularisation of the reduced Hessian
F_uu when it is not positive definite; fraction-to-the-boundary rule; l1-merit backtracking line search; monotone
barrier update.  Every quantity is per problem, so a batch is just a leading axis.
"""
from __future__ import annotations

import numpy as np

DEFAULTS = dict(max_iter=60, tol=1e-6, mu_init=0.1, mu_min=1e-9, kappa_eps=10.0, kappa_mu=0.2, theta_mu=1.5,
                tau_min=0.99, bound_push=1e-2, eta=1e-4, max_backtrack=8, reg_init=1e-8, reg_max=1e10)


def _chol_solve_batched(F, rhs, opt):
    """solve F X = rhs for a batch of small SPD matrices; stage-wise regularisation F += delta I where the Cholesky
    factorisation breaks down.  rhs: (B, u, k).  Returns X and the delta used."""
    B, u, _ = F.shape
    delta = np.zeros(B)
    eye = np.eye(u)[None]
    Fr = F.copy()
    for _ in range(40):
        bad = np.zeros(B, bool)
        L = np.zeros_like(Fr)
        for i in range(u):
            s = Fr[:, i, i] - np.sum(L[:, i, :i] ** 2, axis=1)
            scale = np.maximum(np.abs(Fr[:, i, i]), 1e-300)
            bad |= ~(s > 1e-12 * scale) | ~np.isfinite(s)
            s = np.where(s > 0, s, 1.0)
            L[:, i, i] = np.sqrt(s)
            for j in range(i + 1, u):
                L[:, j, i] = (Fr[:, j, i] - np.sum(L[:, j, :i] * L[:, i, :i], axis=1)) / L[:, i, i]
        if not bad.any():
            break
        delta = np.where(bad, np.where(delta == 0, opt["reg_init"], delta * 10.0), delta)
        delta = np.minimum(delta, opt["reg_max"])
        Fr = F + delta[:, None, None] * eye
    # forward / backward substitution
    Y = np.zeros_like(rhs)
    for i in range(u):
        Y[:, i] = (rhs[:, i] - np.einsum("bj,bjk->bk", L[:, i, :i], Y[:, :i])) / L[:, i, i][:, None]
    Xs = np.zeros_like(rhs)
    for i in range(u - 1, -1, -1):
        Xs[:, i] = (Y[:, i] - np.einsum("bj,bjk->bk", L[:, i + 1:, i], Xs[:, i + 1:])) / L[:, i, i][:, None]
    return Xs, delta


class BatchedIPM:
    def __init__(self, evaluator, lb, ub, **options):
        """``evaluator``: ``oracle.blocks_np.BlockEvaluator`` (with an objective); lb / ub: length-n bound vectors
        (``DomainConstraint.get_lower_bounds(H)`` order), +-inf allowed."""
        self.ev = evaluator
        self.lb = np.asarray(lb, np.float64)
        self.ub = np.asarray(ub, np.float64)
        self.opt = dict(DEFAULTS, **options)
        self.hasL, self.hasU = np.isfinite(self.lb), np.isfinite(self.ub)

    # ---- per-stage views of the sparse value arrays --------------------------------------------------------------
    def _stage_blocks(self, jv, hv):
        ev = self.ev
        H, x, u, d = ev.H, ev.xd, ev.ud, ev.d
        B = jv.shape[0]
        AB = np.zeros((B, H, x, d))
        ok = ev.jmap >= 0
        AB[:, ok] = jv[:, ev.jmap[ok]]
        Wl = np.zeros((B, H, d, d))
        okh = ev.hmap >= 0
        Wl[:, okh] = hv[:, ev.hmap[okh]]
        W = Wl + np.transpose(np.tril(Wl, -1), (0, 1, 3, 2))
        return AB, W

    def _kkt_step(self, Z, X0, lam, zL, zU, mu, out):
        """Riccati solution of the reduced primal-dual system.  Returns dz, lam_new, dzL, dzU, regularisation used."""
        ev, opt = self.ev, self.opt
        H, x, u, d, n = ev.H, ev.xd, ev.ud, ev.d, ev.n
        B = Z.shape[0]
        AB, W = self._stage_blocks(out["jac_vals"], out["hes_vals"])
        c = out["resid"].reshape(B, H, x)
        dL = np.where(self.hasL, Z - self.lb, 1.0)
        dU = np.where(self.hasU, self.ub - Z, 1.0)
        sig = np.where(self.hasL, zL / dL, 0.0) + np.where(self.hasU, zU / dU, 0.0)
        g = out["grad"] - np.where(self.hasL, mu[:, None] / dL, 0.0) + np.where(self.hasU, mu[:, None] / dU, 0.0)
        sx, su = sig[:, :H * x].reshape(B, H, x), sig[:, H * x:].reshape(B, H, u)
        gx, gu = g[:, :H * x].reshape(B, H, x), g[:, H * x:].reshape(B, H, u)
        hd = np.where(ev.hdiag >= 0, out["hes_vals"][:, np.maximum(ev.hdiag, 0)], 0.0)     # objective-only diagonal of x_H
        ix = np.arange(x); iu = np.arange(u)
        # terminal node X_H
        P = np.zeros((B, x, x)); P[:, ix, ix] = hd[:, (H - 1) * x:H * x] + sx[:, H - 1]
        p = gx[:, H - 1].copy()
        K = np.zeros((B, H, u, x)); kf = np.zeros((B, H, u)); reg = np.zeros(B)
        for k in range(H - 1, -1, -1):
            A, Bm = AB[:, k, :, :x], AB[:, k, :, x:]
            h = p + np.einsum("bij,bj->bi", P, c[:, k])
            Fuu = W[:, k, x:, x:] + np.einsum("bji,bjk,bkl->bil", Bm, P, Bm)
            Fuu[:, iu, iu] += su[:, k]
            fu = gu[:, k] + np.einsum("bji,bj->bi", Bm, h)
            if k > 0:
                Fux = W[:, k, x:, :x] + np.einsum("bji,bjk,bkl->bil", Bm, P, A)
                sol, dlt = _chol_solve_batched(Fuu, np.concatenate([Fux, fu[:, :, None]], axis=2), opt)
                K[:, k], kf[:, k] = -sol[:, :, :x], -sol[:, :, x]
                Fxx = W[:, k, :x, :x] + np.einsum("bji,bjk,bkl->bil", A, P, A)
                Fxx[:, ix, ix] += sx[:, k - 1]
                fx = gx[:, k - 1] + np.einsum("bji,bj->bi", A, h)
                P = Fxx + np.einsum("bji,bjk->bik", Fux, K[:, k])
                P = 0.5 * (P + np.transpose(P, (0, 2, 1)))
                p = fx + np.einsum("bji,bj->bi", Fux, kf[:, k])
            else:
                sol, dlt = _chol_solve_batched(Fuu, fu[:, :, None], opt)
                kf[:, k] = -sol[:, :, 0]
            reg = np.maximum(reg, dlt)
        # forward sweep needs the value functions of the NEXT node for the multipliers: recompute them on the way
        stX = np.zeros((B, H + 1, x)); stU = np.zeros((B, H, u))
        for k in range(H):
            stU[:, k] = np.einsum("bij,bj->bi", K[:, k], stX[:, k]) + kf[:, k]
            stX[:, k + 1] = np.einsum("bij,bj->bi", AB[:, k, :, :x], stX[:, k]) * (k > 0) + np.einsum("bij,bj->bi", AB[:, k, :, x:], stU[:, k]) + c[:, k]
        dz = np.concatenate([stX[:, 1:].reshape(B, H * x), stU.reshape(B, H * u)], axis=1)
        # multipliers from stationarity w.r.t. X_{k+1}:  lam_k = (W+Sigma) dz + g restricted to X_{k+1} + Abar_{k+1}^T lam_{k+1}
        lam_new = np.zeros((B, H, x))
        nxt = np.zeros((B, x))
        for k in range(H - 1, -1, -1):
            node = k + 1                                   # X_node = state block k
            r = gx[:, k] + sx[:, k] * stX[:, node]
            if node < H:                                   # Hessian block of stage `node` couples (X_node, U_node)
                r = r + np.einsum("bij,bj->bi", W[:, node, :x, :x], stX[:, node]) + np.einsum("bji,bj->bi", W[:, node, x:, :x], stU[:, node])
                r = r + np.einsum("bji,bj->bi", AB[:, node, :, :x], nxt)
            else:
                r = r + hd[:, (H - 1) * x:H * x] * stX[:, node]
            lam_new[:, k] = r
            nxt = r
        dzL = np.where(self.hasL, mu[:, None] / dL - zL - np.where(self.hasL, zL / dL, 0.0) * dz, 0.0)
        dzU = np.where(self.hasU, mu[:, None] / dU - zU + np.where(self.hasU, zU / dU, 0.0) * dz, 0.0)
        return dz, lam_new.reshape(B, H * x), dzL, dzU, reg, g

    def _merit(self, Z, X0, mu, nu):
        out = self.ev.evaluate(Z, X0, None, 1.0, need_jac=False, need_hes=False)
        dL = np.where(self.hasL, Z - self.lb, 1.0)
        dU = np.where(self.hasU, self.ub - Z, 1.0)
        bar = -mu * (np.sum(np.where(self.hasL, np.log(np.maximum(dL, 1e-300)), 0.0), axis=1)
                     + np.sum(np.where(self.hasU, np.log(np.maximum(dU, 1e-300)), 0.0), axis=1))
        return out["obj"] + bar + nu * np.abs(out["resid"]).sum(axis=1), np.abs(out["resid"]).sum(axis=1)

    def solve(self, X0, Z_init=None):
        ev, opt = self.ev, self.opt
        X0 = np.atleast_2d(np.asarray(X0, np.float64))
        B, n, m, H, x = X0.shape[0], ev.n, ev.m, ev.H, ev.xd
        if Z_init is None:                                   # [x0 tiled | zeros]: optimizer/ipopt.py:149
            Z = np.concatenate([np.tile(X0, (1, H)), np.zeros((B, n - H * x))], axis=1)
        else:
            Z = np.array(Z_init, np.float64).reshape(B, n)
        push = opt["bound_push"]
        with np.errstate(invalid="ignore"):
            lbf, ubf = np.where(self.hasL, self.lb, 0.0), np.where(self.hasU, self.ub, 0.0)
        lo = np.where(self.hasL, lbf + push * np.maximum(1.0, np.abs(lbf)), -np.inf)
        hi = np.where(self.hasU, ubf - push * np.maximum(1.0, np.abs(ubf)), np.inf)
        both = self.hasL & self.hasU
        mid = np.where(both, 0.5 * (lbf + ubf), 0.0)
        lo = np.where(both & (lo > hi), mid, lo); hi = np.where(both & (lo > hi), mid, hi)
        Z = np.minimum(np.maximum(Z, lo), hi)
        lam = np.zeros((B, m)); zL = np.where(self.hasL, 1.0, 0.0) * np.ones((B, n)); zU = np.where(self.hasU, 1.0, 0.0) * np.ones((B, n))
        mu = np.full(B, opt["mu_init"]); nu = np.ones(B)
        done = np.zeros(B, bool); failed = np.zeros(B, bool); iters = np.zeros(B, int); err = np.full(B, np.inf)
        for it in range(opt["max_iter"]):
            out = ev.evaluate(Z, X0, lam, 1.0)
            dL = np.where(self.hasL, Z - self.lb, 1.0); dU = np.where(self.hasU, self.ub - Z, 1.0)
            # dual infeasibility: grad f + J^T lam - zL + zU
            J = np.zeros((B, m, n)); J[:, ev.jac_rows, ev.jac_cols] = out["jac_vals"]
            rd = out["grad"] + np.einsum("bmn,bm->bn", J, lam) - zL + zU
            comp = np.maximum(np.abs(np.where(self.hasL, dL * zL, 0.0)).max(axis=1, initial=0.0), np.abs(np.where(self.hasU, dU * zU, 0.0)).max(axis=1, initial=0.0))
            cinf = np.abs(out["resid"]).max(axis=1)
            err = np.maximum(np.abs(rd).max(axis=1), np.maximum(cinf, comp))
            newly = (~done) & (err <= opt["tol"])
            done |= newly
            if (done | failed).all():
                break
            compmu = np.maximum(np.abs(np.where(self.hasL, dL * zL - mu[:, None], 0.0)).max(axis=1, initial=0.0),
                                np.abs(np.where(self.hasU, dU * zU - mu[:, None], 0.0)).max(axis=1, initial=0.0))
            emu = np.maximum(np.abs(rd).max(axis=1), np.maximum(cinf, compmu))
            shrink = (~done) & (emu <= opt["kappa_eps"] * mu)
            mu = np.where(shrink, np.maximum(opt["mu_min"], np.minimum(opt["kappa_mu"] * mu, mu ** opt["theta_mu"])), mu)
            dz, lam_new, dzL, dzU, reg, g = self._kkt_step(Z, X0, lam, zL, zU, mu, out)
            # a non-finite step (infeasible / unbounded sub-problem, regularisation exhausted) freezes that problem: FAIL
            failed |= (~done) & ~(np.isfinite(dz).all(axis=1) & np.isfinite(lam_new).all(axis=1) & np.isfinite(dzL).all(axis=1) & np.isfinite(dzU).all(axis=1))
            dz = np.where(failed[:, None], 0.0, dz); lam_new = np.where(failed[:, None], lam, lam_new)
            dzL = np.where(failed[:, None], 0.0, dzL); dzU = np.where(failed[:, None], 0.0, dzU)
            tau = np.maximum(opt["tau_min"], 1.0 - mu)[:, None]
            with np.errstate(divide="ignore", invalid="ignore"):
                aP = np.minimum(np.where(self.hasL & (dz < 0), -tau * dL / dz, np.inf).min(axis=1), np.where(self.hasU & (dz > 0), tau * dU / dz, np.inf).min(axis=1))
                aD = np.minimum(np.where(self.hasL & (dzL < 0), -tau * zL / dzL, np.inf).min(axis=1), np.where(self.hasU & (dzU < 0), -tau * zU / dzU, np.inf).min(axis=1))
            aP = np.minimum(1.0, aP); aD = np.minimum(1.0, aD)
            nu = np.maximum(nu, np.abs(lam_new).max(axis=1) + 1.0)
            phi0, c1 = self._merit(Z, X0, mu, nu)
            dphi = np.einsum("bn,bn->b", g, dz) - nu * c1
            alpha = aP.copy(); accepted = np.zeros(B, bool)
            for _ in range(opt["max_backtrack"]):
                phi, _ = self._merit(Z + alpha[:, None] * dz, X0, mu, nu)
                ok = (np.isfinite(phi) & (phi <= phi0 + opt["eta"] * alpha * np.minimum(dphi, 0.0))) | accepted
                accepted |= ok
                if accepted.all():
                    break
                alpha = np.where(accepted, alpha, 0.5 * alpha)
            upd = ~done & ~failed
            Z = np.where(upd[:, None], Z + alpha[:, None] * dz, Z)
            lam = np.where(upd[:, None], lam + alpha[:, None] * (lam_new - lam), lam)
            zL = np.where(upd[:, None], zL + aD[:, None] * dzL, zL)
            zU = np.where(upd[:, None], zU + aD[:, None] * dzU, zU)
            # keep the duals within a factor of mu / slack (IPOPT eq. 16)
            dL = np.where(self.hasL, Z - self.lb, 1.0); dU = np.where(self.hasU, self.ub - Z, 1.0)
            ks = 1e10
            zL = np.where(self.hasL, np.clip(zL, mu[:, None] / (ks * dL), ks * mu[:, None] / dL), 0.0)
            zU = np.where(self.hasU, np.clip(zU, mu[:, None] / (ks * dU), ks * mu[:, None] / dU), 0.0)
            iters += upd
        info = dict(converged=done & ~failed, failed=failed, iterations=iters, kkt_error=err, mu=mu)
        return Z, lam, info
