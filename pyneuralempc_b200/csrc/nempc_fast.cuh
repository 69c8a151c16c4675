// Register-resident kernel body for small two-hidden-layer networks (the Lotka-Volterra class of
// BASELINE configs C1/C2: 3 -> 30 tanh -> 30 tanh -> 2): ONE THREAD PER HORIZON STEP.
//
// Design (B200): the whole weight set (a few KB) is passed as a __grid_constant__ kernel parameter, so every
// weight is an immediate constant-bank operand of an FFMA -- no shared-memory staging, no load instructions,
// no barriers.  All loops are fully unrolled with compile-time indices, activations / tangents live in
// registers, and the layer-2 pre-activation (value + d tangent rows) is accumulated input-stationary so that
// layer 1 is produced one neuron at a time and never stored.  The last hidden layer is consumed on the fly
// (output value, local Jacobian, curvature, adjoint seed), then one more sweep over W2 gives the layer-1
// adjoint and curvature.  Per-output ("forward-only") second-order chain through the RK4 stages as in
// nempc_generic.cuh, which needs no reverse sweep over stages and therefore no stored stage state.
//
// Same maths and references as nempc_generic.cuh (integrator/rk4.py:113-285, model/tensorflow.py:49-109).
#pragma once
#include <math.h>

#include "nempc_generic.cuh"

template <int X, int U, int H1, int H2> struct FastWeights {
    static constexpr int D = X + U;
    static constexpr int NS = D * (D + 1) / 2;
    float W1[D][H1];
    float b1[H1];
    float W2[H1][H2];
    float b2[H2];
    float W3[H2][X];
    float b3[X];
    float P1[NS][H1];   // W1[c][i] * W1[c2][i], packed lower triangle e = c(c+1)/2 + c2: layer-1 tangents are constant
    float W23[X][H1][H2];   // W2[i][j] * W3[j][p]: the per-output layer-1 adjoint is sum_j W23[p][i][j] * s'(a2_j)
};

// host-side fill from Keras-layout double arrays (W[in][out])
template <int X, int U, int H1, int H2>
inline void fill_fast_weights(FastWeights<X, U, H1, H2>& f, const double* W1, const double* b1, const double* W2,
                              const double* b2, const double* W3, const double* b3) {
    constexpr int D = X + U;
    for (int c = 0; c < D; ++c) for (int i = 0; i < H1; ++i) f.W1[c][i] = (float)W1[c * H1 + i];
    for (int i = 0; i < H1; ++i) f.b1[i] = (float)b1[i];
    for (int i = 0; i < H1; ++i) for (int j = 0; j < H2; ++j) f.W2[i][j] = (float)W2[i * H2 + j];
    for (int j = 0; j < H2; ++j) f.b2[j] = (float)b2[j];
    for (int j = 0; j < H2; ++j) for (int p = 0; p < X; ++p) f.W3[j][p] = (float)W3[j * X + p];
    for (int p = 0; p < X; ++p) f.b3[p] = (float)b3[p];
    for (int c = 0; c < D; ++c) for (int c2 = 0; c2 <= c; ++c2) for (int i = 0; i < H1; ++i)
        f.P1[c * (c + 1) / 2 + c2][i] = f.W1[c][i] * f.W1[c2][i];
    for (int p = 0; p < X; ++p) for (int i = 0; i < H1; ++i) for (int j = 0; j < H2; ++j)
        f.W23[p][i][j] = f.W2[i][j] * f.W3[j][p];
}

// per-thread scratch in shared memory for COLD state (touched once per stage, not in the FMA loops):
// element e of thread tid lives at scr[e * stride] with stride = blockDim.x -> bank = tid, conflict-free.
template <int X, int U, int H1, int H2> struct FastScratch {
    static constexpr int NS = (X + U) * (X + U + 1) / 2;
    static constexpr int H1_OFF = 0;                 // layer-1 activations, written in phase A, read in phase C
    static constexpr int HPREV_OFF = H1;             // h_{s-1}[p] packed
    static constexpr int HACC_OFF = H1 + X * NS;     // sum_s c_s h_s[p]
    static constexpr int COUNT = H1 + 2 * X * NS;
};

NEMPC_HD constexpr int tri_index(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

// MODE: 0 residual only, 1 + Jacobian, 2 + Hessian
template <int X, int U, int H1, int H2, int MODE, typename TIO>
NEMPC_HD void fast_step(const FastWeights<X, U, H1, H2>& w, const StageTable<float>& st, const NlpLayout& L,
                        const EvalArgs<TIO>& ar, long long step, float* scr, const int sstride) {
    constexpr int D = X + U, NS = D * (D + 1) / 2;
    constexpr bool JAC = MODE >= 1, HES = MODE >= 2;
    typedef FastScratch<X, U, H1, H2> SC;
    typedef typename WideOf<float, TIO>::type TW;
    const bool unity = (ar.flags & NEMPC_UNITY) != 0;
    const long long b = step / L.H;
    const int t = (int)(step - b * L.H);
    const TIO* zb = ar.z + b * (long long)L.n;

    float z[D];
#pragma unroll
    for (int c = 0; c < X; ++c) z[c] = (float)((t == 0) ? ar.x0[b * X + c] : zb[(t - 1) * X + c]);
#pragma unroll
    for (int c = 0; c < U; ++c) z[X + c] = (float)zb[L.H * X + t * U + c];

    float Rt[X][D];          // top X rows of R_s = I + a_s E dk_{s-1}; the lower U rows stay [0 I]
    float kprev[X], kacc[X], dkacc[X][D];
#pragma unroll
    for (int p = 0; p < X; ++p) {
        kprev[p] = 0.f; kacc[p] = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) { Rt[p][c] = (p == c) ? 1.f : 0.f; dkacc[p][c] = 0.f; }
        if (HES) {
#pragma unroll
            for (int e = 0; e < NS; ++e) { scr[(SC::HPREV_OFF + p * NS + e) * sstride] = 0.f; scr[(SC::HACC_OFF + p * NS + e) * sstride] = 0.f; }
        }
    }

#pragma unroll 1
    for (int s = 0; s < st.S; ++s) {
        const float a_s = st.a[s], c_s = st.c[s];
        float zs[D];
#pragma unroll
        for (int c = 0; c < D; ++c) zs[c] = (c < X) ? fmaf(a_s, kprev[c < X ? c : 0], z[c]) : z[c];

        // ---- phase A: layer 1 neuron by neuron, accumulated straight into the layer-2 pre-activations ------------
        float av[H2], at[JAC ? D : 1][H2];
#pragma unroll
        for (int j = 0; j < H2; ++j) {
            av[j] = w.b2[j];
            if (JAC) {
#pragma unroll
                for (int c = 0; c < D; ++c) at[c][j] = 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < H1; ++i) {
            float a1 = w.b1[i];
#pragma unroll
            for (int c = 0; c < D; ++c) a1 = fmaf(w.W1[c][i], zs[c], a1);
            const float t1 = tanhf(a1);
            if (HES) scr[(SC::H1_OFF + i) * sstride] = t1;
            float v[D];
            if (JAC) {
                const float sp = fmaf(-t1, t1, 1.f);
#pragma unroll
                for (int c = 0; c < D; ++c) v[c] = sp * w.W1[c][i];
            }
#pragma unroll
            for (int j = 0; j < H2; ++j) {
                av[j] = fmaf(w.W2[i][j], t1, av[j]);
                if (JAC) {
#pragma unroll
                    for (int c = 0; c < D; ++c) at[c][j] = fmaf(w.W2[i][j], v[c], at[c][j]);
                }
            }
        }
        // ---- phase B: consume layer 2 on the fly ----------------------------------------------------------------------
        float k[X], J[X][D], M[X][NS], sp2[HES ? H2 : 1];
#pragma unroll
        for (int p = 0; p < X; ++p) {
            k[p] = w.b3[p];
#pragma unroll
            for (int c = 0; c < D; ++c) J[p][c] = 0.f;
#pragma unroll
            for (int e = 0; e < NS; ++e) M[p][e] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < H2; ++j) {
            const float t2 = tanhf(av[j]);
#pragma unroll
            for (int p = 0; p < X; ++p) k[p] = fmaf(w.W3[j][p], t2, k[p]);
            if (JAC) {
                const float sp = fmaf(-t2, t2, 1.f);
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const float vt = sp * at[c][j];
#pragma unroll
                    for (int p = 0; p < X; ++p) J[p][c] = fmaf(w.W3[j][p], vt, J[p][c]);
                }
                if (HES) {
                    const float spp = -2.f * t2 * sp;
                    float pp[NS];
#pragma unroll
                    for (int c = 0; c < D; ++c)
#pragma unroll
                        for (int c2 = 0; c2 <= c; ++c2) pp[c * (c + 1) / 2 + c2] = at[c][j] * at[c2][j];
#pragma unroll
                    for (int p = 0; p < X; ++p) {
                        const float q = spp * w.W3[j][p];
#pragma unroll
                        for (int e = 0; e < NS; ++e) M[p][e] = fmaf(q, pp[e], M[p][e]);
                    }
                    sp2[j] = sp;
                }
            }
        }
        // ---- phase C: layer-1 adjoint (per output) and its curvature ------------------------------------------------------
        if (HES) {
#pragma unroll
            for (int i = 0; i < H1; ++i) {
                const float t1 = scr[(SC::H1_OFF + i) * sstride];
                const float sp = fmaf(-t1, t1, 1.f);
                const float spp = -2.f * t1 * sp;
#pragma unroll
                for (int p = 0; p < X; ++p) {
                    float g = 0.f;
#pragma unroll
                    for (int j = 0; j < H2; ++j) g = fmaf(w.W23[p][i][j], sp2[j], g);
                    const float cf = spp * g;
#pragma unroll
                    for (int e = 0; e < NS; ++e) M[p][e] = fmaf(cf, w.P1[e][i], M[p][e]);
                }
            }
        }
        // ---- stage algebra -------------------------------------------------------------------------------------------------
#define NEMPC_RF(kk, cc) ((kk) < X ? Rt[(kk) < X ? (kk) : 0][cc] : ((kk) == (cc) ? 1.f : 0.f))
        float dk[X][D];
        if (JAC) {
#pragma unroll
            for (int p = 0; p < X; ++p)
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    float acc = (c >= X) ? J[p][c] : 0.f;
#pragma unroll
                    for (int kk = 0; kk < X; ++kk) acc = fmaf(J[p][kk], Rt[kk][c], acc);
                    dk[p][c] = acc;
                    dkacc[p][c] = fmaf(c_s, acc, dkacc[p][c]);
                }
        }
        if (HES) {
            float hs[X][NS];
#pragma unroll
            for (int p = 0; p < X; ++p) {
                float tm[D][D];                                   // M_p R
#pragma unroll
                for (int kk = 0; kk < D; ++kk)
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        float acc = 0.f;
#pragma unroll
                        for (int l2 = 0; l2 < D; ++l2) acc = fmaf(M[p][tri_index(kk, l2)], NEMPC_RF(l2, c), acc);
                        tm[kk][c] = acc;
                    }
#pragma unroll
                for (int a = 0; a < D; ++a)
#pragma unroll
                    for (int c = 0; c <= a; ++c) {
                        float acc = 0.f;
#pragma unroll
                        for (int kk = 0; kk < D; ++kk) acc = fmaf(NEMPC_RF(kk, a), tm[kk][c], acc);
#pragma unroll
                        for (int kk = 0; kk < X; ++kk) acc = fmaf(a_s * J[p][kk], scr[(SC::HPREV_OFF + kk * NS + a * (a + 1) / 2 + c) * sstride], acc);
                        hs[p][a * (a + 1) / 2 + c] = acc;
                    }
            }
#pragma unroll
            for (int p = 0; p < X; ++p)
#pragma unroll
                for (int e = 0; e < NS; ++e) {
                    scr[(SC::HPREV_OFF + p * NS + e) * sstride] = hs[p][e];
                    float* ha = scr + (SC::HACC_OFF + p * NS + e) * sstride;
                    *ha = fmaf(c_s, hs[p][e], *ha);
                }
        }
#undef NEMPC_RF
#pragma unroll
        for (int p = 0; p < X; ++p) { kacc[p] = fmaf(c_s, k[p], kacc[p]); kprev[p] = k[p]; }
        if (JAC && s + 1 < st.S) {
            const float an = st.a[s + 1];
#pragma unroll
            for (int p = 0; p < X; ++p)
#pragma unroll
                for (int c = 0; c < D; ++c) Rt[p][c] = fmaf(an, dk[p][c], (p == c) ? 1.f : 0.f);
        }
    }

    // ---- outputs --------------------------------------------------------------------------------------------------------------
    if (ar.resid) {
#pragma unroll
        for (int p = 0; p < X; ++p) {
            const TW xt = (TW)zb[t * X + p];
            const TW xp = unity ? (TW)0 : (TW)((t == 0) ? ar.x0[b * X + p] : zb[(t - 1) * X + p]);
            ar.resid[b * L.m + t * X + p] = (TIO)(xp + (TW)kacc[p] - xt);
        }
    }
    if (JAC && ar.jac) {
        TIO* jv = ar.jac + b * L.nnz_jac;
#pragma unroll
        for (int p = 0; p < X; ++p) {
            jv[jac_slot_minus1(L, t, p)] = (TIO)-1;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const TW v = (TW)dkacc[p][c] + ((!unity && c == p) ? (TW)1 : (TW)0);
                if (c < X) { if (t > 0) jv[jac_slot_A(L, t, p, c)] = (TIO)v; }
                else jv[jac_slot_B(L, t, p, c - X)] = (TIO)v;
            }
        }
    }
    if (HES && ar.hes) {
        TIO* hv = ar.hes + b * L.nnz_hes;
        const TW sig = ar.sigma ? (TW)ar.sigma[b] : (TW)ar.sigma_scalar;
        float lam[X];
#pragma unroll
        for (int p = 0; p < X; ++p) lam[p] = (float)ar.lam[b * L.m + t * X + p];
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int c = 0; c <= a; ++c) {
                if (t == 0 && c < X) continue;
                float acc = 0.f;
#pragma unroll
                for (int p = 0; p < X; ++p) acc = fmaf(lam[p], scr[(SC::HACC_OFF + p * NS + a * (a + 1) / 2 + c) * sstride], acc);
                TW v = (TW)acc;
                int slot;
                if (a < X) {
                    slot = hes_slot_xx(L, t, a, c);
                    if (a == c && ar.quad) v += sig * (TW)2 * (TW)ar.quad[(t - 1) * X + a];
                } else if (c < X) {
                    slot = hes_slot_ux(L, t, a - X, c);
                } else {
                    slot = hes_slot_uu(L, t, a - X, c - X);
                    if (a == c && ar.quad) v += sig * (TW)2 * (TW)ar.quad[L.H * X + t * U + (a - X)];
                }
                hv[slot] = (TIO)v;
            }
        if (t == L.H - 1) {
#pragma unroll
            for (int p = 0; p < X; ++p)
                if (L.hes_last_slot[p] >= 0) hv[L.hes_last_slot[p]] = (TIO)(sig * (TW)2 * (TW)ar.quad[(L.H - 1) * X + p]);
        }
    }
}
