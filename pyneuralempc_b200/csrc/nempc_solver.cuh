// Batched primal-dual interior-point NMPC solver: per-problem bodies (one thread per problem).
//
// once NLP evaluations are ~1e5x faster than the reference, "MPC solves/s" is bounded by the host
// solver (IPOPT via cyipopt, optimizer/ipopt.py:189, or SciPy SLSQP, optimizer/slsqp.py:172).  This replaces that
// loop for batches of independent problems: the KKT system of the transcription NLP is block tridiagonal in the horizon
// (stage k couples X_k = x_{k-1}, U_k = u_k, X_{k+1} = x_k), so the Newton step of the barrier problem is solved
// EXACTLY by a Riccati sweep that consumes the block-banded Jacobian / Hessian value arrays of nempc_eval in place.
//
// Algorithm = oracle/solver_np.py (IPOPT-style, simplified): reduced primal-dual system with eliminated bound duals,
// stage-wise regularisation of the reduced Hessian F_uu, fraction-to-the-boundary rule, l1-merit backtracking, monotone
// barrier update.  All solver state is float64.  Plain C++ over one problem index, so tests/hostsim can run it on the host.
#pragma once
#include <math.h>
#include <stdint.h>

#include "nempc_layout.h"

#define NEMPC_SOLVER_XM 16
#define NEMPC_SOLVER_UM 16

enum : int { NEMPC_ST_RUNNING = -1, NEMPC_ST_CONVERGED = 0, NEMPC_ST_MAXITER = 1, NEMPC_ST_FAILED = 2 };
// slots of the solver's device counters (nempc_handle::sv_counts): [0] problems still running after the KKT kernel, [1] problems whose line-search
// trial was not accepted, then the state of the device-side iteration loop and the statistics nempc_solve_stats reports
enum : int { NEMPC_SV_RUNNING = 0, NEMPC_SV_NOTACC = 1, NEMPC_SV_TRIAL = 2, NEMPC_SV_IT = 3, NEMPC_SV_LSX = 4, NEMPC_SV_TRIALS_TOTAL = 5, NEMPC_SV_COUNT = 8 };

struct SolverOpts {
    int max_iter, max_backtrack;
    double tol, mu_init, mu_min, kappa_eps, kappa_mu, theta_mu, tau_min, bound_push, eta, reg_init, reg_max;
};

inline SolverOpts solver_defaults() {
    SolverOpts o;
    o.max_iter = 60; o.max_backtrack = 8; o.tol = 1e-6; o.mu_init = 0.1; o.mu_min = 1e-9; o.kappa_eps = 10.0;
    o.kappa_mu = 0.2; o.theta_mu = 1.5; o.tau_min = 0.99; o.bound_push = 1e-2; o.eta = 1e-4; o.reg_init = 1e-8; o.reg_max = 1e10;
    return o;
}

struct SolverWs {                       // device (or host-emulation) arrays, row-major over problems
    const double* x0;                   // (B, x)
    const double *lb, *ub;              // (n) shared bounds, +-inf allowed
    double *z, *lam, *zL, *zU;          // iterate
    double *dz, *lamn, *dzL, *dzU;      // step (lamn = new multipliers of the QP)
    double *grad, *resid, *jac, *hes, *obj;     // nempc_eval outputs at the iterate
    double *zt, *residt, *objt;         // line-search trial point and its residual / objective
    double *K, *kf;                     // Riccati feedback (B, H, u, x), (B, H, u)
    double *mu, *nu, *alpha, *alphaD, *phi0, *dphi, *err;    // (B) scalars
    int *status, *iters, *accepted;     // (B)
};

NEMPC_HD bool nempc_finite(double v) { return v == v && v - v == 0.0; }

// log-barrier of the bound slacks at point zz
NEMPC_HD double ipm_barrier(const NlpLayout& L, const SolverWs& w, const double* zz) {
    double s = 0.0;
    for (int i = 0; i < L.n; ++i) {
        if (nempc_finite(w.lb[i])) { const double d = zz[i] - w.lb[i]; s += log(d > 1e-300 ? d : 1e-300); }
        if (nempc_finite(w.ub[i])) { const double d = w.ub[i] - zz[i]; s += log(d > 1e-300 ? d : 1e-300); }
    }
    return s;
}

NEMPC_HD void ipm_init_problem(const NlpLayout& L, const SolverWs& w, long long b, const SolverOpts& o, bool has_init) {
    double* z = w.z + b * L.n;
    const int nx = L.H * L.x;
    for (int i = 0; i < L.n; ++i) {
        double v = has_init ? z[i] : (i < nx ? w.x0[b * L.x + i % L.x] : 0.0);      // [x0 tiled | zeros]: optimizer/ipopt.py:149
        const bool hl = nempc_finite(w.lb[i]), hu = nempc_finite(w.ub[i]);
        double lo = hl ? w.lb[i] + o.bound_push * fmax(1.0, fabs(w.lb[i])) : -INFINITY;
        double hi = hu ? w.ub[i] - o.bound_push * fmax(1.0, fabs(w.ub[i])) : INFINITY;
        if (hl && hu && lo > hi) lo = hi = 0.5 * (w.lb[i] + w.ub[i]);
        v = fmin(fmax(v, lo), hi);
        z[i] = v;
        w.zL[b * L.n + i] = hl ? 1.0 : 0.0;
        w.zU[b * L.n + i] = hu ? 1.0 : 0.0;
    }
    for (int i = 0; i < L.m; ++i) w.lam[b * L.m + i] = 0.0;
    w.mu[b] = o.mu_init; w.nu[b] = 1.0; w.alpha[b] = 0.0; w.alphaD[b] = 0.0; w.err[b] = INFINITY;
    w.status[b] = NEMPC_ST_RUNNING; w.iters[b] = 0; w.accepted[b] = 1;
}

// solve F X = R (u x u SPD, nr right-hand sides, row-major R[u][nr]) with F += delta I until Cholesky succeeds
template <int UM>
NEMPC_HD double ipm_chol_solve(int u, const double* F, double* R, int nr, const SolverOpts& o) {
    double Lc[UM * UM];
    double delta = 0.0;
    for (int attempt = 0; attempt < 40; ++attempt) {
        bool bad = false;
        for (int i = 0; i < u; ++i) {
            double s = F[i * u + i] + delta;
            for (int k = 0; k < i; ++k) s -= Lc[i * u + k] * Lc[i * u + k];
            const double scale = fmax(fabs(F[i * u + i] + delta), 1e-300);
            if (!(s > 1e-12 * scale) || !nempc_finite(s)) bad = true;
            if (!(s > 0.0)) s = 1.0;
            Lc[i * u + i] = sqrt(s);
            for (int j = i + 1; j < u; ++j) {
                double t = F[j * u + i];
                for (int k = 0; k < i; ++k) t -= Lc[j * u + k] * Lc[i * u + k];
                Lc[j * u + i] = t / Lc[i * u + i];
            }
        }
        if (!bad) break;
        delta = (delta == 0.0) ? o.reg_init : delta * 10.0;
        if (delta > o.reg_max) delta = o.reg_max;
    }
    for (int c = 0; c < nr; ++c) {
        for (int i = 0; i < u; ++i) {
            double t = R[i * nr + c];
            for (int k = 0; k < i; ++k) t -= Lc[i * u + k] * R[k * nr + c];
            R[i * nr + c] = t / Lc[i * u + i];
        }
        for (int i = u - 1; i >= 0; --i) {
            double t = R[i * nr + c];
            for (int k = i + 1; k < u; ++k) t -= Lc[k * u + i] * R[k * nr + c];
            R[i * nr + c] = t / Lc[i * u + i];
        }
    }
    return delta;
}

// ---- Newton step of the barrier problem: Riccati sweeps over the block-tridiagonal KKT system ----------------------------
// sig(i) = Sigma_i (bound-dual curvature), gb(i) = gradient of the barrier objective; functors so that the one-thread body
// evaluates them on the fly while the warp-cooperative kernel reads arrays it filled in parallel (same values either way).
// XM / UM: compile-time capacity of the per-thread matrices (x_dim <= XM, u_dim <= UM).
// XC / UC: exact x_dim / u_dim when known at compile time (0 = run time): every inner loop unrolls and the small matrices
// live in registers instead of local memory.
template <int XM, int UM, int XC, int UC, class SigF, class GbF>
NEMPC_HD void ipm_kkt_riccati(const NlpLayout& L, const SolverOpts& o, const double* c, const double* jv, const double* hv,
                              double* dz, double* lamn, double* Kb, double* kfb, SigF sig, GbF gb) {
    const int H = L.H, x = XC ? XC : L.x, u = UC ? UC : L.u, nx = H * x;
#define A_(k, p, q) ((k) > 0 ? jv[jac_slot_A(L, (k), (p), (q))] : 0.0)
#define B_(k, p, q) (jv[jac_slot_B(L, (k), (p), (q))])
#define WXX_(k, p, q) (hv[hes_slot_xx(L, (k), (p) >= (q) ? (p) : (q), (p) >= (q) ? (q) : (p))])
#define WUX_(k, q, p) (hv[hes_slot_ux(L, (k), (q), (p))])
#define WUU_(k, q, r) (hv[hes_slot_uu(L, (k), (q) >= (r) ? (q) : (r), (q) >= (r) ? (r) : (q))])
#define XI_(node, p) (((node) - 1) * x + (p))          /* variable index of (X_node)_p, node = 1..H */
#define UI_(k, q) (nx + (k) * u + (q))
    // ---- backward Riccati sweep ------------------------------------------------------------------------------------------
    double P[XM * XM], pv[XM], h[XM];
    double PA[XM * XM], PB[XM * UM];
    double Fuu[UM * UM], R[UM * (XM + 1)];
    double Fxx[XM * XM], fx[XM];
    for (int p = 0; p < x; ++p) {
        for (int q = 0; q < x; ++q) P[p * x + q] = 0.0;
        const int i = XI_(H, p);
        P[p * x + p] = (L.hes_last_slot[p] >= 0 ? hv[L.hes_last_slot[p]] : 0.0) + sig(i);
        pv[p] = gb(i);
    }
    for (int k = H - 1; k >= 0; --k) {
        for (int p = 0; p < x; ++p) { double t = pv[p]; for (int q = 0; q < x; ++q) t += P[p * x + q] * c[k * x + q]; h[p] = t; }
        for (int p = 0; p < x; ++p)
            for (int q = 0; q < u; ++q) { double t = 0.0; for (int r = 0; r < x; ++r) t += P[p * x + r] * B_(k, r, q); PB[p * u + q] = t; }
        const int nr = (k > 0) ? x + 1 : 1;               // right-hand sides: [Fux | fu]
        for (int q = 0; q < u; ++q) {
            for (int r = 0; r < u; ++r) { double t = WUU_(k, q, r); for (int p = 0; p < x; ++p) t += B_(k, p, q) * PB[p * u + r]; Fuu[q * u + r] = t; }
            Fuu[q * u + q] += sig(UI_(k, q));
            double t = gb(UI_(k, q));
            for (int p = 0; p < x; ++p) t += B_(k, p, q) * h[p];
            R[q * nr + (nr - 1)] = t;                      // fu
        }
        if (k > 0) {
            for (int p = 0; p < x; ++p)
                for (int q = 0; q < x; ++q) { double t = 0.0; for (int r = 0; r < x; ++r) t += P[p * x + r] * A_(k, r, q); PA[p * x + q] = t; }
            for (int q = 0; q < u; ++q)
                for (int p = 0; p < x; ++p) { double t = WUX_(k, q, p); for (int r = 0; r < x; ++r) t += B_(k, r, q) * PA[r * x + p]; R[q * nr + p] = t; }   // Fux
            // keep Fux (before the solve overwrites R) for the P / p update
            double Fux[UM * XM];
            for (int q = 0; q < u; ++q) for (int p = 0; p < x; ++p) Fux[q * x + p] = R[q * nr + p];
            ipm_chol_solve<UM>(u, Fuu, R, nr, o);
            for (int q = 0; q < u; ++q) {
                for (int p = 0; p < x; ++p) Kb[(k * u + q) * x + p] = -R[q * nr + p];
                kfb[k * u + q] = -R[q * nr + x];
            }
            for (int p = 0; p < x; ++p) {
                for (int q = 0; q < x; ++q) { double t = WXX_(k, p, q); for (int r = 0; r < x; ++r) t += A_(k, r, p) * PA[r * x + q]; Fxx[p * x + q] = t; }
                Fxx[p * x + p] += sig(XI_(k, p));
                double t = gb(XI_(k, p));
                for (int r = 0; r < x; ++r) t += A_(k, r, p) * h[r];
                fx[p] = t;
            }
            for (int p = 0; p < x; ++p) {
                for (int q = 0; q < x; ++q) { double t = Fxx[p * x + q]; for (int r = 0; r < u; ++r) t += Fux[r * x + p] * Kb[(k * u + r) * x + q]; P[p * x + q] = t; }
                double t = fx[p];
                for (int r = 0; r < u; ++r) t += Fux[r * x + p] * kfb[k * u + r];
                pv[p] = t;
            }
            for (int p = 0; p < x; ++p) for (int q = 0; q < p; ++q) { const double s = 0.5 * (P[p * x + q] + P[q * x + p]); P[p * x + q] = s; P[q * x + p] = s; }
        } else {
            ipm_chol_solve<UM>(u, Fuu, R, 1, o);
            for (int q = 0; q < u; ++q) kfb[q] = -R[q];
        }
    }
    // ---- forward sweep -----------------------------------------------------------------------------------------------------
    double xs[XM], xn[XM];
    for (int p = 0; p < x; ++p) xs[p] = 0.0;
    for (int k = 0; k < H; ++k) {
        for (int q = 0; q < u; ++q) {
            double t = kfb[k * u + q];
            if (k > 0) for (int p = 0; p < x; ++p) t += Kb[(k * u + q) * x + p] * xs[p];
            dz[UI_(k, q)] = t;
        }
        for (int p = 0; p < x; ++p) {
            double t = c[k * x + p];
            for (int q = 0; q < x; ++q) t += A_(k, p, q) * xs[q];
            for (int q = 0; q < u; ++q) t += B_(k, p, q) * dz[UI_(k, q)];
            xn[p] = t;
        }
        for (int p = 0; p < x; ++p) { xs[p] = xn[p]; dz[XI_(k + 1, p)] = xn[p]; }
    }
    // ---- multipliers of the QP from stationarity w.r.t. X_{k+1} -------------------------------------------------------------
    for (int k = H - 1; k >= 0; --k) {
        const int node = k + 1;
        for (int p = 0; p < x; ++p) {
            const int i = XI_(node, p);
            double r = gb(i) + sig(i) * dz[i];
            if (node < H) {
                for (int q = 0; q < x; ++q) r += WXX_(node, p, q) * dz[XI_(node, q)];
                for (int q = 0; q < u; ++q) r += WUX_(node, q, p) * dz[UI_(node, q)];
                for (int q = 0; q < x; ++q) r += A_(node, q, p) * lamn[node * x + q];
            } else if (L.hes_last_slot[p] >= 0) {
                r += hv[L.hes_last_slot[p]] * dz[i];
            }
            lamn[k * x + p] = r;
        }
    }
#undef WXX_
#undef WUX_
#undef WUU_
}

// dual-infeasibility residual of variable block (k, j): j < x -> (X_{k+1})_j, else (U_k)_{j-x}
NEMPC_HD double ipm_dual_residual(const NlpLayout& L, int x, int u, const double* gr, const double* lam, const double* zL,
                                  const double* zU, const double* jv, int k, int j) {
    const int H = L.H, nx = H * x;
    if (j < x) {
        const int p = j, i = XI_(k + 1, p);
        double r = gr[i] - lam[k * x + p] - zL[i] + zU[i];
        if (k + 1 < H) for (int q = 0; q < x; ++q) r += A_(k + 1, q, p) * lam[(k + 1) * x + q];
        return r;
    }
    const int q = j - x, i = UI_(k, q);
    double r = gr[i] - zL[i] + zU[i];
    for (int p = 0; p < x; ++p) r += B_(k, p, q) * lam[k * x + p];
    return r;
}
#undef A_
#undef B_
#undef XI_
#undef UI_

NEMPC_HD double ipm_sigma(const SolverWs& w, const double* z, const double* zL, const double* zU, int i) {
    double s = 0.0;
    if (nempc_finite(w.lb[i])) s += zL[i] / (z[i] - w.lb[i]);
    if (nempc_finite(w.ub[i])) s += zU[i] / (w.ub[i] - z[i]);
    return s;
}
NEMPC_HD double ipm_gbar(const SolverWs& w, const double* z, const double* gr, double mu, int i) {
    double g = gr[i];
    if (nempc_finite(w.lb[i])) g -= mu / (z[i] - w.lb[i]);
    if (nempc_finite(w.ub[i])) g += mu / (w.ub[i] - z[i]);
    return g;
}
// bound-dual steps of variable i and its fraction-to-the-boundary limits (aP, aD are running minima)
NEMPC_HD void ipm_bound_step(const SolverWs& w, const double* z, const double* zL, const double* zU, const double* dz, double mu,
                             double tau, int i, double& dl, double& du, double& aP, double& aD) {
    dl = 0.0; du = 0.0;
    if (nempc_finite(w.lb[i])) {
        const double d = z[i] - w.lb[i];
        dl = mu / d - zL[i] - zL[i] / d * dz[i];
        if (dz[i] < 0.0) aP = fmin(aP, -tau * d / dz[i]);
        if (dl < 0.0) aD = fmin(aD, -tau * zL[i] / dl);
    }
    if (nempc_finite(w.ub[i])) {
        const double d = w.ub[i] - z[i];
        du = mu / d - zU[i] + zU[i] / d * dz[i];
        if (dz[i] > 0.0) aP = fmin(aP, tau * d / dz[i]);
        if (du < 0.0) aD = fmin(aD, -tau * zU[i] / du);
    }
}

// One interior-point iteration up to (and including) the first line-search trial point: ONE THREAD per problem.
template <int XM, int UM, int XC = 0, int UC = 0>
NEMPC_HD void ipm_kkt_problem(const NlpLayout& L, const SolverWs& w, long long b, const SolverOpts& o) {
    if (w.status[b] != NEMPC_ST_RUNNING) { w.accepted[b] = 1; return; }
    const int H = L.H, x = XC ? XC : L.x, u = UC ? UC : L.u, n = L.n, m = L.m;
    const double* z = w.z + b * n; const double* lam = w.lam + b * m;
    const double* zL = w.zL + b * n; const double* zU = w.zU + b * n;
    const double* gr = w.grad + b * n; const double* c = w.resid + b * m;
    const double* jv = w.jac + b * L.nnz_jac; const double* hv = w.hes + b * L.nnz_hes;
    double* dz = w.dz + b * n; double* lamn = w.lamn + b * m; double* dzL = w.dzL + b * n; double* dzU = w.dzU + b * n;
    double* Kb = w.K + b * (long long)H * u * x; double* kfb = w.kf + b * (long long)H * u;
    double mu = w.mu[b];

    // ---- optimality error (max-norm of dual infeasibility, constraint violation, complementarity) -------------------
    double e_d = 0.0, e_c = 0.0, e_comp = 0.0, e_compmu = 0.0;
    for (int i = 0; i < m; ++i) e_c = fmax(e_c, fabs(c[i]));
    for (int k = 0; k < H; ++k)
        for (int j = 0; j < x + u; ++j) e_d = fmax(e_d, fabs(ipm_dual_residual(L, x, u, gr, lam, zL, zU, jv, k, j)));
    for (int i = 0; i < n; ++i) {
        if (nempc_finite(w.lb[i])) { const double v = (z[i] - w.lb[i]) * zL[i]; e_comp = fmax(e_comp, fabs(v)); e_compmu = fmax(e_compmu, fabs(v - mu)); }
        if (nempc_finite(w.ub[i])) { const double v = (w.ub[i] - z[i]) * zU[i]; e_comp = fmax(e_comp, fabs(v)); e_compmu = fmax(e_compmu, fabs(v - mu)); }
    }
    const double err = fmax(e_d, fmax(e_c, e_comp));
    w.err[b] = err;
    if (err <= o.tol) { w.status[b] = NEMPC_ST_CONVERGED; w.accepted[b] = 1; return; }
    if (fmax(e_d, fmax(e_c, e_compmu)) <= o.kappa_eps * mu) {
        mu = fmax(o.mu_min, fmin(o.kappa_mu * mu, pow(mu, o.theta_mu)));
        w.mu[b] = mu;
    }

    ipm_kkt_riccati<XM, UM, XC, UC>(L, o, c, jv, hv, dz, lamn, Kb, kfb,
                                    [&](int i) { return ipm_sigma(w, z, zL, zU, i); }, [&](int i) { return ipm_gbar(w, z, gr, mu, i); });

    // ---- bound-dual steps, fraction to the boundary, merit ------------------------------------------------------------------
    const double tau = fmax(o.tau_min, 1.0 - mu);
    double aP = 1.0, aD = 1.0, lmax = 0.0, gdz = 0.0, c1 = 0.0;
    bool finite = true;
    for (int i = 0; i < n; ++i) {
        double dl, du;
        ipm_bound_step(w, z, zL, zU, dz, mu, tau, i, dl, du, aP, aD);
        dzL[i] = dl; dzU[i] = du;
        gdz += ipm_gbar(w, z, gr, mu, i) * dz[i];
        finite = finite && nempc_finite(dz[i]) && nempc_finite(dl) && nempc_finite(du);
    }
    for (int i = 0; i < m; ++i) { lmax = fmax(lmax, fabs(lamn[i])); c1 += fabs(c[i]); finite = finite && nempc_finite(lamn[i]); }
    if (!finite) { w.status[b] = NEMPC_ST_FAILED; w.accepted[b] = 1; return; }
    const double nu = fmax(w.nu[b], lmax + 1.0);
    w.nu[b] = nu;
    w.phi0[b] = w.obj[b] - mu * ipm_barrier(L, w, z) + nu * c1;
    w.dphi[b] = gdz - nu * c1;
    w.alpha[b] = aP; w.alphaD[b] = aD; w.accepted[b] = 0;
    double* zt = w.zt + b * n;
    for (int i = 0; i < n; ++i) zt[i] = z[i] + aP * dz[i];
}

#if defined(__CUDACC__)
// The same iteration by ONE WARP per problem on a shared-memory copy of the problem's rows (nempc_ipm_kkt_staged_kernel).
// Everything that is independent per variable -- the residual norms, Sigma and the barrier gradient (two divisions each),
// the bound-dual steps (six divisions), the barrier logarithms -- is spread over the 32 lanes; maxima and minima are exact
// under any order, and the three running SUMS (g^T dz, |c|_1, the barrier) are added by lane 0 in the sequential order from
// terms computed in parallel, so every value equals the one-thread body bit for bit.  Lane 0 alone runs the Riccati sweeps.
// s1, s2, s3: three scratch arrays of n doubles in shared memory.  All pointers are the problem's own rows (no b offset).
struct KktRows {
    const double *z, *lam, *zL, *zU, *gr;                     // global rows (read by the lane-parallel phases only)
    const double *c, *jv, *hv;                                // shared-memory copies (read by the sequential Riccati lane)
    double *dz, *lamn, *Kb, *kfb;                             // shared memory, copied out by the caller
    double *dzL, *dzU, *zt;                                   // global rows (write only, coalesced)
    double *s1, *s2, *s3;
};
__device__ __forceinline__ double warp_max_f64(double v) { for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ double warp_min_f64(double v) { for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }

template <int XM, int UM, int XC, int UC>
__device__ void ipm_kkt_warp(const NlpLayout& L, const SolverWs& w, long long b, const SolverOpts& o, const KktRows& r, int lane) {
    const int H = L.H, x = XC ? XC : L.x, u = UC ? UC : L.u, n = L.n, m = L.m, d = x + u;
    double mu = w.mu[b];
    // ---- optimality error ---------------------------------------------------------------------------------------------------
    double e_d = 0.0, e_c = 0.0, e_comp = 0.0, e_compmu = 0.0;
    for (int i = lane; i < m; i += 32) e_c = fmax(e_c, fabs(r.c[i]));
    for (int e = lane; e < H * d; e += 32) e_d = fmax(e_d, fabs(ipm_dual_residual(L, x, u, r.gr, r.lam, r.zL, r.zU, r.jv, e / d, e % d)));
    for (int i = lane; i < n; i += 32) {
        if (nempc_finite(w.lb[i])) { const double v = (r.z[i] - w.lb[i]) * r.zL[i]; e_comp = fmax(e_comp, fabs(v)); e_compmu = fmax(e_compmu, fabs(v - mu)); }
        if (nempc_finite(w.ub[i])) { const double v = (w.ub[i] - r.z[i]) * r.zU[i]; e_comp = fmax(e_comp, fabs(v)); e_compmu = fmax(e_compmu, fabs(v - mu)); }
    }
    e_d = warp_max_f64(e_d); e_c = warp_max_f64(e_c); e_comp = warp_max_f64(e_comp); e_compmu = warp_max_f64(e_compmu);
    const double err = fmax(e_d, fmax(e_c, e_comp));
    if (lane == 0) w.err[b] = err;
    if (err <= o.tol) { if (lane == 0) { w.status[b] = NEMPC_ST_CONVERGED; w.accepted[b] = 1; } return; }
    if (fmax(e_d, fmax(e_c, e_compmu)) <= o.kappa_eps * mu) {
        mu = fmax(o.mu_min, fmin(o.kappa_mu * mu, pow(mu, o.theta_mu)));
        if (lane == 0) w.mu[b] = mu;
    }
    // ---- Sigma and barrier gradient of every variable, in parallel -------------------------------------------------------------
    for (int i = lane; i < n; i += 32) { r.s1[i] = ipm_sigma(w, r.z, r.zL, r.zU, i); r.s2[i] = ipm_gbar(w, r.z, r.gr, mu, i); }
    __syncwarp();
    if (lane == 0) {
        const double* sg = r.s1; const double* gbv = r.s2;
        ipm_kkt_riccati<XM, UM, XC, UC>(L, o, r.c, r.jv, r.hv, r.dz, r.lamn, r.Kb, r.kfb, [sg](int i) { return sg[i]; }, [gbv](int i) { return gbv[i]; });
    }
    __syncwarp();
    // ---- bound-dual steps, fraction to the boundary, merit ------------------------------------------------------------------------
    const double tau = fmax(o.tau_min, 1.0 - mu);
    double aP = 1.0, aD = 1.0, lmax = 0.0;
    bool finite = true;
    for (int i = lane; i < n; i += 32) {
        double dl, du;
        ipm_bound_step(w, r.z, r.zL, r.zU, r.dz, mu, tau, i, dl, du, aP, aD);
        r.dzL[i] = dl; r.dzU[i] = du;
        const double dzi = r.dz[i], zi = r.z[i];
        r.s3[i] = r.s2[i] * dzi;                                             // gb(i) * dz[i]
        finite = finite && nempc_finite(dzi) && nempc_finite(dl) && nempc_finite(du);
        // barrier terms at z (ipm_barrier adds the lower-bound term, then the upper-bound term, variable by variable)
        double tl = 0.0, tu = 0.0;
        if (nempc_finite(w.lb[i])) { const double dd = zi - w.lb[i]; tl = log(dd > 1e-300 ? dd : 1e-300); }
        if (nempc_finite(w.ub[i])) { const double dd = w.ub[i] - zi; tu = log(dd > 1e-300 ? dd : 1e-300); }
        r.s1[i] = tl; r.s2[i] = tu;
    }
    for (int i = lane; i < m; i += 32) { lmax = fmax(lmax, fabs(r.lamn[i])); finite = finite && nempc_finite(r.lamn[i]); }
    aP = warp_min_f64(aP); aD = warp_min_f64(aD); lmax = warp_max_f64(lmax);
    finite = __all_sync(0xffffffffu, finite);
    if (!finite) { if (lane == 0) { w.status[b] = NEMPC_ST_FAILED; w.accepted[b] = 1; } return; }
    __syncwarp();
    if (lane == 0) {
        double gdz = 0.0, c1 = 0.0, bar = 0.0;
        for (int i = 0; i < n; ++i) gdz += r.s3[i];
        for (int i = 0; i < m; ++i) c1 += fabs(r.c[i]);
        for (int i = 0; i < n; ++i) { bar += r.s1[i]; bar += r.s2[i]; }      // terms of infinite bounds are +0.0: adding them changes no bit
        const double nu = fmax(w.nu[b], lmax + 1.0);
        w.nu[b] = nu;
        w.phi0[b] = w.obj[b] - mu * bar + nu * c1;
        w.dphi[b] = gdz - nu * c1;
        w.alpha[b] = aP; w.alphaD[b] = aD; w.accepted[b] = 0;
    }
    for (int i = lane; i < n; i += 32) r.zt[i] = r.z[i] + aP * r.dz[i];
}
#endif

// after residt / objt were evaluated at the trial point: Armijo test on the l1 merit; halve the step otherwise
NEMPC_HD void ipm_linesearch_problem(const NlpLayout& L, const SolverWs& w, long long b, const SolverOpts& o) {
    if (w.status[b] != NEMPC_ST_RUNNING || w.accepted[b]) return;
    const int n = L.n, m = L.m;
    const double* zt = w.zt + b * n; const double* ct = w.residt + b * m;
    double c1 = 0.0;
    for (int i = 0; i < m; ++i) c1 += fabs(ct[i]);
    const double phi = w.objt[b] - w.mu[b] * ipm_barrier(L, w, zt) + w.nu[b] * c1;
    const double a = w.alpha[b];
    if (nempc_finite(phi) && phi <= w.phi0[b] + o.eta * a * fmin(w.dphi[b], 0.0)) { w.accepted[b] = 1; return; }
    const double an = 0.5 * a;
    w.alpha[b] = an;
    const double* z = w.z + b * n; const double* dz = w.dz + b * n; double* ztw = w.zt + b * n;
    for (int i = 0; i < n; ++i) ztw[i] = z[i] + an * dz[i];
}

NEMPC_HD void ipm_update_problem(const NlpLayout& L, const SolverWs& w, long long b, const SolverOpts& o) {
    if (w.status[b] != NEMPC_ST_RUNNING) return;
    const int n = L.n, m = L.m;
    const double a = w.alpha[b], aD = w.alphaD[b], mu = w.mu[b];
    double* z = w.z + b * n; double* lam = w.lam + b * m; double* zL = w.zL + b * n; double* zU = w.zU + b * n;
    const double* dz = w.dz + b * n; const double* lamn = w.lamn + b * m; const double* dzL = w.dzL + b * n; const double* dzU = w.dzU + b * n;
    const double ks = 1e10;
    for (int i = 0; i < n; ++i) {
        z[i] += a * dz[i];
        if (nempc_finite(w.lb[i])) { const double d = z[i] - w.lb[i]; zL[i] = fmin(fmax(zL[i] + aD * dzL[i], mu / (ks * d)), ks * mu / d); }
        if (nempc_finite(w.ub[i])) { const double d = w.ub[i] - z[i]; zU[i] = fmin(fmax(zU[i] + aD * dzU[i], mu / (ks * d)), ks * mu / d); }
    }
    for (int i = 0; i < m; ++i) lam[i] += a * (lamn[i] - lam[i]);
    w.iters[b] += 1;
}
