// Register-resident kernel body for small two-hidden-layer networks (the Lotka-Volterra class of
// BASELINE configs C1/C2: 3 -> 30 tanh -> 30 tanh -> 2): ONE THREAD PER HORIZON STEP.
//
// Design (B200):
//  * the whole weight set (~13 KB incl. padding and pre-multiplied products) is a __grid_constant__ kernel
//    parameter; every loop over neurons is warp-uniform, so weights arrive as LDCU.128 (uniform constant loads
//    into UNIFORM registers) and feed FFMA as a uniform operand -- no per-thread registers, no shared-memory
//    staging, no barriers;
//  * loops over INPUT neurons are real loops (compact code: the round-1 profile profiles/r1a showed the fully
//    unrolled variant stalled 3.9 cycles/issue on instruction fetch), loops over the register-resident OUTPUT
//    chunk are unrolled (compile-time register indices);
//  * layer-2 pre-activations (value + D tangent rows) are accumulated in registers for a chunk of H2/NCHUNK
//    output neurons at a time and consumed on the fly (output value, local Jacobian, curvature, adjoint seed);
//  * cold per-thread state (layer-1 activations, s'(a2), per-output Hessian accumulators) lives in a
//    [element][thread] shared-memory scratch: bank = thread, conflict-free;
//  * per-output ("forward-only") second-order chain through the RK4 stages as in nempc_generic.cuh: no reverse
//    sweep over stages, hence no stored stage state.
//
// Same maths and references as nempc_generic.cuh (integrator/rk4.py:113-285, model/tensorflow.py:49-109).
#pragma once
#include <math.h>
#include <string.h>

#include "nempc_generic.cuh"

// ---- packed f32x2 arithmetic (sm_100a FFMA2 / FMUL2) --------------------------------------------------------------
// On the device an f2 is a 64-bit register pair and fma2 is `fma.rn.f32x2`; ptxas folds pk(s, s) into the scalar-
// broadcast operand form and takes constant-bank pairs as a uniform-register operand
// (`FFMA2 R, R.F32, UR.F32x2, R.F32x2`), so a packed FMA costs ONE issue slot for two FMAs per lane.
// On the host (tests/hostsim) the same names are plain float pairs.
#if defined(__CUDA_ARCH__)
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float f2lo(f2 a) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); return lo; }
__device__ __forceinline__ float f2hi(f2 a) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); return hi; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v)); return d; }
// tanh on the MUFU pipe: 1 - 2 / (exp(2|x|) + 1) with EX2 + RCP (absolute error ~1e-7, saturates correctly to +-1) for |x| >= 0.25;
// below that the formula cancels -- its RELATIVE error grows like 1e-7 / |x|, and s'' = -2 h (1 - h^2) inherits it, which an
// elementwise parity check of the Hessian sees -- so a degree-9 odd polynomial takes over (truncation 2e-9 at 0.25): a few 1e-7
// relative everywhere.  ~12 instructions instead of ~20 for libdevice tanhf.
__device__ __forceinline__ float fast_tanh(float x) {
    const float ax = fabsf(x);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * 2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    const float big = fmaf(-2.0f, r, 1.0f);
    const float x2 = ax * ax;
    float p = fmaf(x2, 0.021869488536155202f, -0.05396825396825397f);
    p = fmaf(x2, p, 0.13333333333333333f);
    p = fmaf(x2, p, -0.3333333333333333f);
    p = fmaf(ax * x2, p, ax);
    return copysignf(ax < 0.25f ? p : big, x);
}
// the register-resident kernel's variant, 11 instructions instead of 15 (60 tanh per RK4 stage are ~4 % of its instructions): the MUFU
// formula on the signed argument (saturates by itself) for |x| >= 1/8 -- relative error <= 1e-6 there -- and x - x^3/3 + 2 x^5/15 below
// (truncation 17/315 x^7: 2e-7 relative at 1/8)
__device__ __forceinline__ float fast_tanh_lv(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    const float big = fmaf(-2.0f, r, 1.0f);
    const float x2 = x * x;
    const float p = fmaf(x2 * x, fmaf(x2, 0.13333333333333333f, -0.3333333333333333f), x);
    return x2 < 0.015625f ? p : big;
}
#else
struct f2 { float lo, hi; };
inline f2 pk(float lo, float hi) { f2 r; r.lo = lo; r.hi = hi; return r; }
inline float f2lo(f2 a) { return a.lo; }
inline float f2hi(f2 a) { return a.hi; }
inline f2 fma2(f2 a, f2 b, f2 c) { return pk(fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)); }
inline f2 mul2(f2 a, f2 b) { return pk(a.lo * b.lo, a.hi * b.hi); }
inline float fast_tanh(float x) {
    const float ax = fabsf(x), e = exp2f(ax * 2.885390081777927f), big = fmaf(-2.0f, 1.0f / (e + 1.0f), 1.0f), x2 = ax * ax;
    float p = fmaf(x2, 0.021869488536155202f, -0.05396825396825397f);
    p = fmaf(x2, p, 0.13333333333333333f);
    p = fmaf(x2, p, -0.3333333333333333f);
    p = fmaf(ax * x2, p, ax);
    return copysignf(ax < 0.25f ? p : big, x);
}
inline float fast_tanh_lv(float x) {
    const float e = exp2f(x * 2.885390081777927f), big = fmaf(-2.0f, 1.0f / (e + 1.0f), 1.0f), x2 = x * x;
    const float p = fmaf(x2 * x, fmaf(x2, 0.13333333333333333f, -0.3333333333333333f), x);
    return x2 < 0.015625f ? p : big;
}
#endif

// NEMPC_FAST_CONTRACT_LAST = 1: the LAST integrator stage (the only one of the discrete / unity integrators) contracts the multipliers into
// the layer-1 adjoint -- w_S = c_S lambda is known there, so ONE adjoint row (15 FFMA2 per layer-1 neuron against the W2 pairs of the forward
// loop) replaces the two per-output rows (30 against the W2 W3 table), and the stage algebra runs once instead of per output.  The earlier
// RK4 stages need the Jacobians of later stages for their multipliers (w_s = c_s lambda + a_{s+1} J_{s+1,x}^T w_{s+1}) and keep the
// per-output chain.
#ifndef NEMPC_FAST_CONTRACT_LAST
#define NEMPC_FAST_CONTRACT_LAST 1
#endif

// 16-byte weight quad: read as ONE uniform 128-bit constant load (LDCU.128) and consumed as two FFMA2 operand pairs
struct alignas(16) nf4 { float x, y, z, w; };

template <int X, int U, int H1, int H2, int NCHUNK> struct FastWeights {
    static constexpr int D = X + U;
    static constexpr int NS = D * (D + 1) / 2;
    static constexpr int DP = (D + 1 + 3) / 4 * 4;        // W1 column + bias, padded to 128 bit
    static constexpr int JC = H2 / NCHUNK;                // output neurons per register chunk
    static constexpr int JCP = (JC + 3) / 4 * 4;
    static constexpr int H2P = (H2 + 3) / 4 * 4;
    static constexpr int NSP = (NS + 3) / 4 * 4;
    static_assert(H2 % NCHUNK == 0, "H2 must be divisible by NCHUNK");
    static_assert(X <= 4, "W3T packs x_dim into one 128-bit slot");
    static_assert(DP == 4 && X == 2, "the register-resident kernel is instantiated for x_dim = 2, x_dim+u_dim = 3");
    nf4 W1T[H1];                            // (W1[0][i], W1[1][i], W1[2][i], b1[i])
    nf4 W2C[NCHUNK][H1][JCP / 4];           // [jc][i][q] = W2[i][jc*JC + 4q .. 4q+3]  (zero padded)
    alignas(16) float b2[NCHUNK][JCP];
    alignas(16) float W3T[NCHUNK][JC][4];   // [jc][jj][p] = W3[j][p]
    alignas(16) float b3[4];
    static constexpr int XQ = (X + 1) / 2 * 2;            // outputs padded to an even count (packed pairs over p)
    nf4 W23T[H1][H2P / 2];                  // [i][h] = (W2[i][2h]W3[2h][0], W2[i][2h]W3[2h][1], W2[i][2h+1]W3[2h+1][0], W2[i][2h+1]W3[2h+1][1])
    nf4 P1T[H1][NSP / 4];                   // W1[c][i] * W1[c2][i], e = c(c+1)/2 + c2 (zero padded): layer-1 tangents are constant
};

// host-side fill from Keras-layout double arrays (W[in][out])
template <int X, int U, int H1, int H2, int NCHUNK>
inline void fill_fast_weights(FastWeights<X, U, H1, H2, NCHUNK>& f, const double* W1, const double* b1, const double* W2,
                              const double* b2, const double* W3, const double* b3) {
    typedef FastWeights<X, U, H1, H2, NCHUNK> FW;
    constexpr int D = X + U, JC = FW::JC;
    memset(&f, 0, sizeof(FW));
    auto at = [](nf4* base, int idx) -> float& { return (&base[idx / 4].x)[idx % 4]; };
    for (int i = 0; i < H1; ++i) {
        float w1[D];
        for (int c = 0; c < D; ++c) { w1[c] = (float)W1[c * H1 + i]; at(&f.W1T[i], c) = w1[c]; }
        at(&f.W1T[i], D) = (float)b1[i];
        for (int c = 0; c < D; ++c)
            for (int c2 = 0; c2 <= c; ++c2) at(f.P1T[i], c * (c + 1) / 2 + c2) = w1[c] * w1[c2];
    }
    for (int jc = 0; jc < NCHUNK; ++jc)
        for (int jj = 0; jj < JC; ++jj) {
            const int j = jc * JC + jj;
            f.b2[jc][jj] = (float)b2[j];
            for (int i = 0; i < H1; ++i) at(f.W2C[jc][i], jj) = (float)W2[i * H2 + j];
            for (int p = 0; p < X; ++p) f.W3T[jc][jj][p] = (float)W3[j * X + p];
        }
    for (int p = 0; p < X; ++p) f.b3[p] = (float)b3[p];
    for (int i = 0; i < H1; ++i)
        for (int j = 0; j < H2; ++j)
            for (int p = 0; p < X; ++p) at(f.W23T[i], 2 * j + p) = (float)W2[i * H2 + j] * (float)W3[j * X + p];
}

// per-thread scratch in shared memory for COLD state: element e of thread tid lives at scr[e * stride] with
// stride = blockDim.x -> bank = tid, conflict-free.
template <int X, int U, int H1, int H2> struct FastScratch {
    static constexpr int NS = (X + U) * (X + U + 1) / 2;
    static constexpr int H1_OFF = 0;                     // layer-1 activations
    static constexpr int SP2_OFF = H1;                   // s'(a2_j)
    static constexpr int HPREV_OFF = H1 + H2;            // h_{s-1}[p] packed
    static constexpr int HACC_OFF = H1 + H2 + X * NS;    // sum_s c_s h_s[p]
    static constexpr int COUNT = H1 + H2 + 2 * X * NS;
    static constexpr int count(int mode) { return mode >= 2 ? COUNT : H1; }
};

// MODE: 0 residual only, 1 + Jacobian, 2 + Hessian
template <int X, int U, int H1, int H2, int NCHUNK, int MODE, typename TIO>
NEMPC_HD void fast_step(const FastWeights<X, U, H1, H2, NCHUNK>& w, const StageTable<float>& st, const NlpLayout& L,
                        const EvalArgs<TIO>& ar, long long step, float* scr, const int sstride) {
    typedef FastWeights<X, U, H1, H2, NCHUNK> FW;
    typedef FastScratch<X, U, H1, H2> SC;
    constexpr int D = X + U, NS = D * (D + 1) / 2, JC = FW::JC, NR = (MODE >= 1) ? 1 + D : 1;
    constexpr bool JAC = MODE >= 1, HES = MODE >= 2;
    static_assert(!NEMPC_FAST_CONTRACT_LAST || JC % 2 == 0, "the contracted adjoint pairs the layer-2 neurons inside a register chunk");
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

    float lamf[X];
#pragma unroll
    for (int p = 0; p < X; ++p) lamf[p] = (HES && ar.lam) ? (float)ar.lam[b * L.m + t * X + p] : 0.f;
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

        // ---- layer 1: activations to scratch -------------------------------------------------------------------
#pragma unroll 2
        for (int i = 0; i < H1; ++i) {
            const nf4 w1 = w.W1T[i];
            const float a1 = fmaf(w1.x, zs[0], fmaf(w1.y, zs[1], fmaf(w1.z, zs[2], w1.w)));
            scr[(SC::H1_OFF + i) * sstride] = fast_tanh_lv(a1);
        }

        // packed state: k2 = (k[0],k[1]), J2[c] = (J[0][c],J[1][c]) [pairs over outputs p], M2[p][e/2] [pairs over e]
        constexpr int XP = (X + 1) / 2, NSH = (NS + 1) / 2, JCH = FW::JCP / 2;
        f2 k2[XP], J2[XP][D], M2[X][NSH];
#pragma unroll
        for (int q = 0; q < XP; ++q) {
            k2[q] = pk(w.b3[2 * q], w.b3[2 * q + 1]);
#pragma unroll
            for (int c = 0; c < D; ++c) J2[q][c] = pk(0.f, 0.f);
        }
#pragma unroll
        for (int p = 0; p < X; ++p)
#pragma unroll
            for (int e = 0; e < NSH; ++e) M2[p][e] = pk(0.f, 0.f);
        // ---- layer 2, one register chunk of JC output neurons at a time ---------------------------------------------
#pragma unroll 1
        for (int jc = 0; jc < NCHUNK; ++jc) {
            f2 acc[NR][JCH];                       // acc[r][h] = rows (value, tangents) x neuron pair (2h, 2h+1)
#pragma unroll
            for (int h = 0; h < JCH; ++h) {
                acc[0][h] = pk(w.b2[jc][2 * h], w.b2[jc][2 * h + 1]);
#pragma unroll
                for (int r = 1; r < NR; ++r) acc[r][h] = pk(0.f, 0.f);
            }
#pragma unroll 2
            for (int i = 0; i < H1; ++i) {
                const float t1 = scr[(SC::H1_OFF + i) * sstride];
                float v[D];
                if (JAC) {
                    const float sp = fmaf(-t1, t1, 1.f);
                    const nf4 w1 = w.W1T[i];
                    v[0] = sp * w1.x; v[1] = sp * w1.y; v[2] = sp * w1.z;     // post-activation tangent of layer 1
                }
#pragma unroll
                for (int h = 0; h < JCH; ++h) {
                    if (2 * h >= JC) continue;                                // padding pair of the last quad
                    const nf4 wq = w.W2C[jc][i][h / 2];
                    const f2 wp = (h & 1) ? pk(wq.z, wq.w) : pk(wq.x, wq.y);
                    acc[0][h] = fma2(pk(t1, t1), wp, acc[0][h]);
                    if (JAC) {
#pragma unroll
                        for (int c = 0; c < D; ++c) acc[JAC ? 1 + c : 0][h] = fma2(pk(v[c], v[c]), wp, acc[JAC ? 1 + c : 0][h]);
                    }
                }
            }
            // consume the chunk: output value, local Jacobian, layer-2 curvature, adjoint seed
#pragma unroll
            for (int jj = 0; jj < JC; ++jj) {
                const float a2 = (jj & 1) ? f2hi(acc[0][jj / 2]) : f2lo(acc[0][jj / 2]);
                const float t2 = fast_tanh_lv(a2);
                f2 w3[XP];
#pragma unroll
                for (int q = 0; q < XP; ++q) {
                    w3[q] = pk(w.W3T[jc][jj][2 * q], w.W3T[jc][jj][2 * q + 1]);
                    k2[q] = fma2(pk(t2, t2), w3[q], k2[q]);
                }
                if (JAC) {
                    const float sp = fmaf(-t2, t2, 1.f);
                    float tg[D];
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        tg[c] = (jj & 1) ? f2hi(acc[JAC ? 1 + c : 0][jj / 2]) : f2lo(acc[JAC ? 1 + c : 0][jj / 2]);
                        const float vt = sp * tg[c];
#pragma unroll
                        for (int q = 0; q < XP; ++q) J2[q][c] = fma2(pk(vt, vt), w3[q], J2[q][c]);
                    }
                    if (HES) {
                        const float spp = -2.f * t2 * sp;
                        float pp[2 * NSH];
#pragma unroll
                        for (int c = 0; c < D; ++c)
#pragma unroll
                            for (int c2 = 0; c2 <= c; ++c2) pp[c * (c + 1) / 2 + c2] = tg[c] * tg[c2];
                        if (NS & 1) pp[NS] = 0.f;
#pragma unroll
                        for (int p = 0; p < X; ++p) {
                            const float q = spp * w.W3T[jc][jj][p];
#pragma unroll
                            for (int e = 0; e < NSH; ++e) M2[p][e] = fma2(pk(q, q), pk(pp[2 * e], pp[2 * e + 1]), M2[p][e]);
                        }
                        scr[(SC::SP2_OFF + jc * JC + jj) * sstride] = sp;
                    }
                }
            }
        }
        // ---- layer-1 adjoint (per output) and its curvature ---------------------------------------------------------------
        // g2[q] = (g[2q], g[2q+1])[i] = sum_j s'(a2_j) * W23T[i][j][2q..2q+1]: scalar-broadcast x uniform pair (the FFMA2 form
        // of the layer-2 loop), four independent chains; s'(a2) stays in registers for the whole loop.
        const bool lastc = NEMPC_FAST_CONTRACT_LAST && HES && (s + 1 == st.S);
        f2 Ml2[NSH];                                   // contracted layer-1 curvature of the last stage (pairs over e)
#pragma unroll
        for (int e = 0; e < NSH; ++e) Ml2[e] = pk(0.f, 0.f);
        if (HES && lastc) {
            // y_j = s'(a2_j) * (W3[j][.] . lambda): one adjoint row, contracted with the multipliers
            f2 y2[H2 / 2];
#pragma unroll
            for (int h = 0; h < H2 / 2; ++h) {
                float yv[2];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int j = 2 * h + r, jc = j / JC, jj = j - jc * JC;
                    float wl = 0.f;
#pragma unroll
                    for (int p = 0; p < X; ++p) wl = fmaf(w.W3T[jc][jj][p], lamf[p], wl);
                    yv[r] = scr[(SC::SP2_OFF + j) * sstride] * wl;
                }
                y2[h] = pk(yv[0], yv[1]);
            }
#pragma unroll 1
            for (int i = 0; i < H1; ++i) {
                const float t1 = scr[(SC::H1_OFF + i) * sstride];
                const float sp = fmaf(-t1, t1, 1.f);
                const float spp = -2.f * t1 * sp;
                f2 g[2];
                g[0] = pk(0.f, 0.f); g[1] = pk(0.f, 0.f);
#pragma unroll
                for (int h = 0; h < H2 / 2; ++h) {
                    const int j = 2 * h, jc = j / JC, hh = (j - jc * JC) / 2;          // JC is even: a pair never straddles two chunks
                    const nf4 wq = w.W2C[jc][i][hh / 2];
                    g[h & 1] = fma2(y2[h], (hh & 1) ? pk(wq.z, wq.w) : pk(wq.x, wq.y), g[h & 1]);
                }
                const float gl = (f2lo(g[0]) + f2hi(g[0])) + (f2lo(g[1]) + f2hi(g[1]));
                const float cf = spp * gl;
#pragma unroll
                for (int e = 0; e < NSH; ++e) {
                    const nf4 pq = w.P1T[i][e / 2];
                    Ml2[e] = fma2(pk(cf, cf), (e & 1) ? pk(pq.z, pq.w) : pk(pq.x, pq.y), Ml2[e]);
                }
            }
        }
        if (HES && !lastc) {
            float sp2[H2];
#pragma unroll
            for (int j = 0; j < H2; ++j) sp2[j] = scr[(SC::SP2_OFF + j) * sstride];
#pragma unroll 1
            for (int i = 0; i < H1; ++i) {
                const float t1 = scr[(SC::H1_OFF + i) * sstride];
                const float sp = fmaf(-t1, t1, 1.f);
                const float spp = -2.f * t1 * sp;
#pragma unroll
                for (int q = 0; q < XP; ++q) {
                    f2 g[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) g[r] = pk(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < H2; ++j) {
                        const nf4 wq = w.W23T[i][j / 2];
                        g[j & 3] = fma2(pk(sp2[j], sp2[j]), (j & 1) ? pk(wq.z, wq.w) : pk(wq.x, wq.y), g[j & 3]);
                    }
                    const float glo = (f2lo(g[0]) + f2lo(g[1])) + (f2lo(g[2]) + f2lo(g[3]));
                    const float ghi = (f2hi(g[0]) + f2hi(g[1])) + (f2hi(g[2]) + f2hi(g[3]));
                    const float cf0 = spp * glo, cf1 = spp * ghi;
#pragma unroll
                    for (int e = 0; e < NSH; ++e) {
                        const nf4 pq = w.P1T[i][e / 2];
                        const f2 p1 = (e & 1) ? pk(pq.z, pq.w) : pk(pq.x, pq.y);
                        M2[2 * q][e] = fma2(pk(cf0, cf0), p1, M2[2 * q][e]);
                        if (2 * q + 1 < X) M2[2 * q + 1 < X ? 2 * q + 1 : 0][e] = fma2(pk(cf1, cf1), p1, M2[2 * q + 1 < X ? 2 * q + 1 : 0][e]);
                    }
                }
            }
        }
        // unpack for the (small) stage algebra
        float k[X], J[X][D], M[X][NS];
#pragma unroll
        for (int p = 0; p < X; ++p) {
            k[p] = (p & 1) ? f2hi(k2[p / 2]) : f2lo(k2[p / 2]);
#pragma unroll
            for (int c = 0; c < D; ++c) J[p][c] = (p & 1) ? f2hi(J2[p / 2][c]) : f2lo(J2[p / 2][c]);
#pragma unroll
            for (int e = 0; e < NS; ++e) M[p][e] = (e & 1) ? f2hi(M2[p][e / 2]) : f2lo(M2[p][e / 2]);
        }
        // ---- stage algebra -------------------------------------------------------------------------------------------------
#define NEMPC_RF(kk, cc) ((kk) < X ? Rt[(kk) < X ? (kk) : 0][cc] : ((kk) == (cc) ? 1.f : 0.f))
        float dk[X][D];
        if (JAC) {
#pragma unroll
            for (int p = 0; p < X; ++p)
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    float a = (c >= X) ? J[p][c] : 0.f;
#pragma unroll
                    for (int kk = 0; kk < X; ++kk) a = fmaf(J[p][kk], Rt[kk][c], a);
                    dk[p][c] = a;
                    dkacc[p][c] = fmaf(c_s, a, dkacc[p][c]);
                }
        }
        if (HES && lastc) {
            // last stage, contracted: out = sum_p lambda_p hacc[p] + c_s (R^T M(lambda) R + a_s sum_k (lambda^T J)[k] h_{s-1}[k]),
            // M(lambda) = sum_p lambda_p M_p (layer-2 curvature, per output above) + the contracted layer-1 curvature
            float Ml[NS], jl[X];
#pragma unroll
            for (int e = 0; e < NS; ++e) {
                float a = (e & 1) ? f2hi(Ml2[e / 2]) : f2lo(Ml2[e / 2]);
#pragma unroll
                for (int p = 0; p < X; ++p) a = fmaf(lamf[p], M[p][e], a);
                Ml[e] = a;
            }
#pragma unroll
            for (int kk = 0; kk < X; ++kk) {
                float a = 0.f;
#pragma unroll
                for (int p = 0; p < X; ++p) a = fmaf(lamf[p], J[p][kk], a);
                jl[kk] = a_s * a;
            }
            float tm[D][D];
#pragma unroll
            for (int kk = 0; kk < D; ++kk)
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    float a = 0.f;
#pragma unroll
                    for (int l2 = 0; l2 < D; ++l2) a = fmaf(Ml[l2 <= kk ? kk * (kk + 1) / 2 + l2 : l2 * (l2 + 1) / 2 + kk], NEMPC_RF(l2, c), a);
                    tm[kk][c] = a;
                }
#pragma unroll
            for (int a2 = 0; a2 < D; ++a2)
#pragma unroll
                for (int c = 0; c <= a2; ++c) {
                    const int e = a2 * (a2 + 1) / 2 + c;
                    float a = 0.f;
#pragma unroll
                    for (int kk = 0; kk < D; ++kk) a = fmaf(NEMPC_RF(kk, a2), tm[kk][c], a);
#pragma unroll
                    for (int kk = 0; kk < X; ++kk) a = fmaf(jl[kk], scr[(SC::HPREV_OFF + kk * NS + e) * sstride], a);
                    float o = c_s * a;
#pragma unroll
                    for (int p = 0; p < X; ++p) o = fmaf(lamf[p], scr[(SC::HACC_OFF + p * NS + e) * sstride], o);
                    scr[(SC::HACC_OFF + e) * sstride] = o;         // e < NS: slot of output 0, read (p = 0, e) just above by this thread only
                }
        }
        if (HES && !lastc) {
            float hs[X][NS];
#pragma unroll
            for (int p = 0; p < X; ++p) {
                float tm[D][D];                                   // M_p R
#pragma unroll
                for (int kk = 0; kk < D; ++kk)
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        float a = 0.f;
#pragma unroll
                        for (int l2 = 0; l2 < D; ++l2) a = fmaf(M[p][l2 <= kk ? kk * (kk + 1) / 2 + l2 : l2 * (l2 + 1) / 2 + kk], NEMPC_RF(l2, c), a);
                        tm[kk][c] = a;
                    }
#pragma unroll
                for (int a2 = 0; a2 < D; ++a2)
#pragma unroll
                    for (int c = 0; c <= a2; ++c) {
                        float a = 0.f;
#pragma unroll
                        for (int kk = 0; kk < D; ++kk) a = fmaf(NEMPC_RF(kk, a2), tm[kk][c], a);
                        hs[p][a2 * (a2 + 1) / 2 + c] = a;
                    }
            }
            // + a_s sum_k J[p][k] h_{s-1}[k]   (h_{s-1} is read from scratch BEFORE it is overwritten)
#pragma unroll
            for (int e = 0; e < NS; ++e) {
                float hp[X];
#pragma unroll
                for (int kk = 0; kk < X; ++kk) hp[kk] = scr[(SC::HPREV_OFF + kk * NS + e) * sstride];
#pragma unroll
                for (int p = 0; p < X; ++p) {
                    float a = hs[p][e];
#pragma unroll
                    for (int kk = 0; kk < X; ++kk) a = fmaf(a_s * J[p][kk], hp[kk], a);
                    hs[p][e] = a;
                }
#pragma unroll
                for (int p = 0; p < X; ++p) {
                    scr[(SC::HPREV_OFF + p * NS + e) * sstride] = hs[p][e];
                    float* ha = scr + (SC::HACC_OFF + p * NS + e) * sstride;
                    *ha = fmaf(c_s, hs[p][e], *ha);
                }
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
                if (NEMPC_FAST_CONTRACT_LAST) acc = scr[(SC::HACC_OFF + a * (a + 1) / 2 + c) * sstride];      // contracted by the last stage
                else {
#pragma unroll
                    for (int p = 0; p < X; ++p) acc = fmaf(lam[p], scr[(SC::HACC_OFF + p * NS + a * (a + 1) / 2 + c) * sstride], acc);
                }
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
