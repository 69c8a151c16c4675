// Sparse output layout of one problem: where every per-step block entry lands in the Jacobian /
// Hessian value arrays.  Shared by the kernels (device) and the structure generator (host), so the
// indices handed to the solver and the slots the kernels write cannot drift apart.
//
// Orders reproduce the reference: Jacobian = np.nonzero of the dense matrix of integrator/discret.py:32-58
// (row-major); Hessian = np.nonzero(np.tril(map)) of optimizer/ipopt.py:55-62.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define NEMPC_HD __host__ __device__ __forceinline__
#else
#define NEMPC_HD inline
#endif

#define NEMPC_LAYOUT_MAX_X 16

struct NlpLayout {
    int H, x, u, d;
    int n, m;
    int tri_x, tri_u;          // x(x+1)/2, u(u+1)/2
    int jac_row_first;         // entries per Jacobian row of step 0:  1 + u
    int jac_row_rest;          // entries per Jacobian row of step >= 1: x + 1 + u
    int hes_last_base;         // first slot of the objective-only diagonal entries of state block H-1
    int hes_last_count;
    int hes_last_slot[NEMPC_LAYOUT_MAX_X];   // slot of (x_H)_p diagonal or -1
    int hes_u_base;            // first slot of the control rows
    int hes_u_row_rest;        // x*u + tri_u : slots per control block for t >= 1
    long long nnz_jac, nnz_hes;
};

// quad_mask: n bytes (non-zero = objective has a diagonal Hessian entry there) or nullptr.
inline void nlp_layout_init(NlpLayout& L, int H, int x, int u, const uint8_t* quad_mask) {
    L.H = H; L.x = x; L.u = u; L.d = x + u;
    L.n = H * (x + u); L.m = H * x;
    L.tri_x = x * (x + 1) / 2; L.tri_u = u * (u + 1) / 2;
    L.jac_row_first = 1 + u;
    L.jac_row_rest = x + 1 + u;
    L.nnz_jac = (long long)x * L.jac_row_first + (long long)(H - 1) * x * L.jac_row_rest;
    L.hes_last_base = (H - 1) * L.tri_x;
    L.hes_last_count = 0;
    for (int p = 0; p < NEMPC_LAYOUT_MAX_X; ++p) L.hes_last_slot[p] = -1;
    for (int p = 0; p < x; ++p)
        if (quad_mask && quad_mask[(H - 1) * x + p]) L.hes_last_slot[p] = L.hes_last_base + L.hes_last_count++;
    L.hes_u_base = L.hes_last_base + L.hes_last_count;
    L.hes_u_row_rest = x * u + L.tri_u;
    L.nnz_hes = (long long)L.hes_u_base + L.tri_u + (long long)(H - 1) * L.hes_u_row_rest;
}

// ---- Jacobian ---------------------------------------------------------------------------------------
// first slot of constraint row (t, p)
NEMPC_HD int jac_row_start(const NlpLayout& L, int t, int p) {
    return t == 0 ? p * L.jac_row_first : L.x * L.jac_row_first + ((t - 1) * L.x + p) * L.jac_row_rest;
}
// slot of d c_{t,p} / d (x_{t-1})_q   (t >= 1 only)
NEMPC_HD int jac_slot_A(const NlpLayout& L, int t, int p, int q) { return jac_row_start(L, t, p) + q; }
// slot of the constant -1 = d c_{t,p} / d (x_t)_p
NEMPC_HD int jac_slot_minus1(const NlpLayout& L, int t, int p) { return jac_row_start(L, t, p) + (t == 0 ? 0 : L.x); }
// slot of d c_{t,p} / d (u_t)_q
NEMPC_HD int jac_slot_B(const NlpLayout& L, int t, int p, int q) {
    return jac_row_start(L, t, p) + (t == 0 ? 1 : L.x + 1) + q;
}

// ---- Hessian (lower triangle) --------------------------------------------------------------------------
// (x_{t-1})_p x (x_{t-1})_q, q <= p, written by step t >= 1 (state block t-1)
NEMPC_HD int hes_slot_xx(const NlpLayout& L, int t, int p, int q) { return (t - 1) * L.tri_x + p * (p + 1) / 2 + q; }
// first slot of control row (t, q)
NEMPC_HD int hes_urow_start(const NlpLayout& L, int t, int q) {
    return t == 0 ? L.hes_u_base + q * (q + 1) / 2
                  : L.hes_u_base + L.tri_u + (t - 1) * L.hes_u_row_rest + q * L.x + q * (q + 1) / 2;
}
// (u_t)_q x (x_{t-1})_p, t >= 1
NEMPC_HD int hes_slot_ux(const NlpLayout& L, int t, int q, int p) { return hes_urow_start(L, t, q) + p; }
// (u_t)_q x (u_t)_r, r <= q
NEMPC_HD int hes_slot_uu(const NlpLayout& L, int t, int q, int r) {
    return hes_urow_start(L, t, q) + (t == 0 ? 0 : L.x) + r;
}
