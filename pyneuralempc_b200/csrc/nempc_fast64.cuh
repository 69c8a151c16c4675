// FLOAT64 register-resident kernel for the small two-hidden-layer networks of BASELINE configs C1 / C2 (3 -> 30 tanh -> 30 tanh -> 2):
// ONE THREAD PER HORIZON STEP, every multiply-add a DFMA.  This is the 1e-10 parity mode (the reference assembles in float64,
// optimizer/ipopt.py:66-86) at register-resident speed; before it, float64 arithmetic only ran on the generic kernel (shared-memory
// workspace, 1.5 - 3.3 TFLOP/s).
//
// Same mathematics and the same loop structure as nempc_fast.cuh (the "forward-only per-output" second-order chain of
// integrator/rk4.py:113-285: no stored stage state), without the f32x2 packing:
//   * weights (25 - 28 KB of doubles incl. the pre-multiplied W2 W3 and W1 (x) W1 tables) arrive as a __grid_constant__ kernel parameter and
//     are copied to shared memory once per CTA: the neuron loops are warp-uniform, so every weight is ONE broadcast LDS.64 (reading them
//     straight from the constant bank measured 14 % slower: the image overflows the immediate-constant cache);
//   * layer-2 pre-activations (value + d tangent rows) accumulate in registers for JC output neurons at a time and are consumed on the
//     fly (output value, local Jacobian, curvature, adjoint seed);
//   * cold per-thread state (layer-1 activations, s'(a2), per-output Hessian accumulators: 84 doubles) sits in an [element][thread]
//     shared-memory scratch (bank pair = thread, conflict free);
//   * tanh is libdevice's double-precision tanh (1 ulp): 60 per stage.
// Roofline: the FP64 FMA pipe (nempc_measure_fma_peak(F64)).
#pragma once
#include "nempc_fast.cuh"

template <int X, int U, int H1, int H2> struct Fast64Weights {
    static constexpr int D = X + U, NS = D * (D + 1) / 2;
    static_assert(D == 3 && X == 2, "the register-resident kernels are instantiated for x_dim = 2, x_dim + u_dim = 3");
    double W1[H1][4];             // (W1[0][i], W1[1][i], W1[2][i], b1[i])
    double W2[H1][H2];
    double b2[H2];
    double W3[H2][X];
    double b3[X];
    double W23[H1][H2][X];        // W2[i][j] W3[j][p]
    double P1[H1][NS];            // W1[c][i] W1[c2][i], e = c (c + 1) / 2 + c2
};

template <int X, int U, int H1, int H2>
inline void fill_fast64_weights(Fast64Weights<X, U, H1, H2>& f, const double* W1, const double* b1, const double* W2,
                                const double* b2, const double* W3, const double* b3) {
    constexpr int D = X + U;
    memset(&f, 0, sizeof(f));
    for (int i = 0; i < H1; ++i) {
        for (int c = 0; c < D; ++c) f.W1[i][c] = W1[c * H1 + i];
        f.W1[i][D] = b1[i];
        for (int c = 0; c < D; ++c)
            for (int c2 = 0; c2 <= c; ++c2) f.P1[i][c * (c + 1) / 2 + c2] = W1[c * H1 + i] * W1[c2 * H1 + i];
        for (int j = 0; j < H2; ++j) {
            f.W2[i][j] = W2[i * H2 + j];
            for (int p = 0; p < X; ++p) f.W23[i][j][p] = W2[i * H2 + j] * W3[j * X + p];
        }
    }
    for (int j = 0; j < H2; ++j) {
        f.b2[j] = b2[j];
        for (int p = 0; p < X; ++p) f.W3[j][p] = W3[j * X + p];
    }
    for (int p = 0; p < X; ++p) f.b3[p] = b3[p];
}

#if defined(__CUDACC__)
// MODE: 0 residual only, 1 + Jacobian, 2 + Hessian;  JC = layer-2 neurons per register chunk (divides H2)
template <int X, int U, int H1, int H2, int JC, int MODE, typename TIO>
__device__ __forceinline__ void fast64_step(const Fast64Weights<X, U, H1, H2>& w, const StageTable<double>& st, const NlpLayout& L,
                                            const EvalArgs<TIO>& ar, long long step, double* scr, const int sstride) {
    typedef FastScratch<X, U, H1, H2> SC;
    constexpr int D = X + U, NS = D * (D + 1) / 2, NR = (MODE >= 1) ? 1 + D : 1;
    constexpr bool JAC = MODE >= 1, HES = MODE >= 2;
    static_assert(H2 % JC == 0, "JC must divide H2");
    const bool unity = (ar.flags & NEMPC_UNITY) != 0;
    const long long b = step / L.H;
    const int t = (int)(step - b * L.H);
    const TIO* zb = ar.z + b * (long long)L.n;

    double z[D];
#pragma unroll
    for (int c = 0; c < X; ++c) z[c] = (double)((t == 0) ? ar.x0[b * X + c] : zb[(t - 1) * X + c]);
#pragma unroll
    for (int c = 0; c < U; ++c) z[X + c] = (double)zb[L.H * X + t * U + c];

    double lamd[X];          // multipliers of this step's rows: the LAST stage contracts them into its layer-1 adjoint (NEMPC_FAST_CONTRACT_LAST, nempc_fast.cuh)
#pragma unroll
    for (int p = 0; p < X; ++p) lamd[p] = (HES && ar.lam) ? (double)ar.lam[b * L.m + t * X + p] : 0.0;
    double Rt[X][D];         // top X rows of R_s = I + a_s E dk_{s-1}; the lower U rows stay [0 I]
    double kprev[X], kacc[X], dkacc[X][D];
#pragma unroll
    for (int p = 0; p < X; ++p) {
        kprev[p] = 0.0; kacc[p] = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) { Rt[p][c] = (p == c) ? 1.0 : 0.0; dkacc[p][c] = 0.0; }
        if (HES) {
#pragma unroll
            for (int e = 0; e < NS; ++e) { scr[(SC::HPREV_OFF + p * NS + e) * sstride] = 0.0; scr[(SC::HACC_OFF + p * NS + e) * sstride] = 0.0; }
        }
    }

#pragma unroll 1
    for (int s = 0; s < st.S; ++s) {
        const double a_s = st.a[s], c_s = st.c[s];
        double zs[D];
#pragma unroll
        for (int c = 0; c < D; ++c) zs[c] = (c < X) ? fma(a_s, kprev[c < X ? c : 0], z[c]) : z[c];

        // ---- layer 1: activations to scratch
#pragma unroll 2
        for (int i = 0; i < H1; ++i) {
            const double a1 = fma(w.W1[i][0], zs[0], fma(w.W1[i][1], zs[1], fma(w.W1[i][2], zs[2], w.W1[i][3])));
            scr[(SC::H1_OFF + i) * sstride] = tanh(a1);
        }
        double k[X], J[X][D], M[X][NS];
#pragma unroll
        for (int p = 0; p < X; ++p) {
            k[p] = w.b3[p];
#pragma unroll
            for (int c = 0; c < D; ++c) J[p][c] = 0.0;
#pragma unroll
            for (int e = 0; e < NS; ++e) M[p][e] = 0.0;
        }
        // ---- layer 2, one register chunk of JC output neurons at a time
#pragma unroll 1
        for (int j0 = 0; j0 < H2; j0 += JC) {
            double acc[NR][JC];
#pragma unroll
            for (int jj = 0; jj < JC; ++jj) {
                acc[0][jj] = w.b2[j0 + jj];
#pragma unroll
                for (int r = 1; r < NR; ++r) acc[r][jj] = 0.0;
            }
#pragma unroll 2
            for (int i = 0; i < H1; ++i) {
                const double t1 = scr[(SC::H1_OFF + i) * sstride];
                double v[D];
                if (JAC) {
                    const double sp = fma(-t1, t1, 1.0);
#pragma unroll
                    for (int c = 0; c < D; ++c) v[c] = sp * w.W1[i][c];          // post-activation tangent of layer 1
                }
#pragma unroll
                for (int jj = 0; jj < JC; ++jj) {
                    const double wv = w.W2[i][j0 + jj];
                    acc[0][jj] = fma(t1, wv, acc[0][jj]);
                    if (JAC) {
#pragma unroll
                        for (int c = 0; c < D; ++c) acc[JAC ? 1 + c : 0][jj] = fma(v[c], wv, acc[JAC ? 1 + c : 0][jj]);
                    }
                }
            }
            // consume the chunk: output value, local Jacobian, layer-2 curvature, adjoint seed
#pragma unroll
            for (int jj = 0; jj < JC; ++jj) {
                const int j = j0 + jj;
                const double t2 = tanh(acc[0][jj]);
#pragma unroll
                for (int p = 0; p < X; ++p) k[p] = fma(t2, w.W3[j][p], k[p]);
                if (JAC) {
                    const double sp = fma(-t2, t2, 1.0);
                    double tg[D];
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        tg[c] = acc[JAC ? 1 + c : 0][jj];
                        const double vt = sp * tg[c];
#pragma unroll
                        for (int p = 0; p < X; ++p) J[p][c] = fma(vt, w.W3[j][p], J[p][c]);
                    }
                    if (HES) {
                        const double spp = -2.0 * t2 * sp;
                        double pp[NS];
#pragma unroll
                        for (int c = 0; c < D; ++c)
#pragma unroll
                            for (int c2 = 0; c2 <= c; ++c2) pp[c * (c + 1) / 2 + c2] = tg[c] * tg[c2];
#pragma unroll
                        for (int p = 0; p < X; ++p) {
                            const double q = spp * w.W3[j][p];
#pragma unroll
                            for (int e = 0; e < NS; ++e) M[p][e] = fma(q, pp[e], M[p][e]);
                        }
                        scr[(SC::SP2_OFF + j) * sstride] = sp;
                    }
                }
            }
        }
        // ---- layer-1 adjoint (per output) and its curvature: g[p][i] = sum_j s'(a2_j) W2[i][j] W3[j][p]
        const bool lastc = NEMPC_FAST_CONTRACT_LAST && HES && (s + 1 == st.S);
        double Ml[NS];                                            // contracted layer-1 curvature of the last stage
#pragma unroll
        for (int e = 0; e < NS; ++e) Ml[e] = 0.0;
        if (HES && lastc) {
            // y_j = s'(a2_j) (W3[j][.] . lambda) replaces s'(a2_j) in the scratch: ONE adjoint row g[i] = sum_j y_j W2[i][j]
#pragma unroll 2
            for (int j = 0; j < H2; ++j) {
                double wl = 0.0;
#pragma unroll
                for (int p = 0; p < X; ++p) wl = fma(w.W3[j][p], lamd[p], wl);
                scr[(SC::SP2_OFF + j) * sstride] *= wl;
            }
#pragma unroll 1
            for (int i = 0; i < H1; ++i) {
                const double t1 = scr[(SC::H1_OFF + i) * sstride];
                const double sp = fma(-t1, t1, 1.0);
                const double spp = -2.0 * t1 * sp;
                double g[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int j = 0; j < H2; ++j) g[j & 3] = fma(scr[(SC::SP2_OFF + j) * sstride], w.W2[i][j], g[j & 3]);
                const double cf = spp * ((g[0] + g[1]) + (g[2] + g[3]));
#pragma unroll
                for (int e = 0; e < NS; ++e) Ml[e] = fma(cf, w.P1[i][e], Ml[e]);
            }
        }
        if (HES && !lastc) {
#pragma unroll 1
            for (int i = 0; i < H1; ++i) {
                const double t1 = scr[(SC::H1_OFF + i) * sstride];
                const double sp = fma(-t1, t1, 1.0);
                const double spp = -2.0 * t1 * sp;
                double g[X][2];
#pragma unroll
                for (int p = 0; p < X; ++p) g[p][0] = g[p][1] = 0.0;
#pragma unroll
                for (int j = 0; j < H2; ++j) {
                    const double s2 = scr[(SC::SP2_OFF + j) * sstride];
#pragma unroll
                    for (int p = 0; p < X; ++p) g[p][j & 1] = fma(s2, w.W23[i][j][p], g[p][j & 1]);
                }
#pragma unroll
                for (int p = 0; p < X; ++p) {
                    const double cf = spp * (g[p][0] + g[p][1]);
#pragma unroll
                    for (int e = 0; e < NS; ++e) M[p][e] = fma(cf, w.P1[i][e], M[p][e]);
                }
            }
        }
        // ---- stage algebra ---------------------------------------------------------------------------------------
#define NEMPC_RF64(kk, cc) ((kk) < X ? Rt[(kk) < X ? (kk) : 0][cc] : ((kk) == (cc) ? 1.0 : 0.0))
        double dk[X][D];
        if (JAC) {
#pragma unroll
            for (int p = 0; p < X; ++p)
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    double a = (c >= X) ? J[p][c] : 0.0;
#pragma unroll
                    for (int kk = 0; kk < X; ++kk) a = fma(J[p][kk], Rt[kk][c], a);
                    dk[p][c] = a;
                    dkacc[p][c] = fma(c_s, a, dkacc[p][c]);
                }
        }
        if (HES && lastc) {
            // out = sum_p lambda_p hacc[p] + c_s (R^T M(lambda) R + a_s sum_k (lambda^T J)[k] h_{s-1}[k]), M(lambda) = sum_p lambda_p M_p + Ml
            double jl[X];
#pragma unroll
            for (int e = 0; e < NS; ++e) {
#pragma unroll
                for (int p = 0; p < X; ++p) Ml[e] = fma(lamd[p], M[p][e], Ml[e]);
            }
#pragma unroll
            for (int kk = 0; kk < X; ++kk) {
                double a = 0.0;
#pragma unroll
                for (int p = 0; p < X; ++p) a = fma(lamd[p], J[p][kk], a);
                jl[kk] = a_s * a;
            }
            double tm[D][D];
#pragma unroll
            for (int kk = 0; kk < D; ++kk)
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    double a = 0.0;
#pragma unroll
                    for (int l2 = 0; l2 < D; ++l2) a = fma(Ml[l2 <= kk ? kk * (kk + 1) / 2 + l2 : l2 * (l2 + 1) / 2 + kk], NEMPC_RF64(l2, c), a);
                    tm[kk][c] = a;
                }
#pragma unroll
            for (int a2 = 0; a2 < D; ++a2)
#pragma unroll
                for (int c = 0; c <= a2; ++c) {
                    const int e = a2 * (a2 + 1) / 2 + c;
                    double a = 0.0;
#pragma unroll
                    for (int kk = 0; kk < D; ++kk) a = fma(NEMPC_RF64(kk, a2), tm[kk][c], a);
#pragma unroll
                    for (int kk = 0; kk < X; ++kk) a = fma(jl[kk], scr[(SC::HPREV_OFF + kk * NS + e) * sstride], a);
                    double o = c_s * a;
#pragma unroll
                    for (int p = 0; p < X; ++p) o = fma(lamd[p], scr[(SC::HACC_OFF + p * NS + e) * sstride], o);
                    scr[(SC::HACC_OFF + e) * sstride] = o;
                }
        }
        if (HES && !lastc) {
            double hs[X][NS];
#pragma unroll
            for (int p = 0; p < X; ++p) {
                double tm[D][D];                                  // M_p R
#pragma unroll
                for (int kk = 0; kk < D; ++kk)
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        double a = 0.0;
#pragma unroll
                        for (int l2 = 0; l2 < D; ++l2) a = fma(M[p][l2 <= kk ? kk * (kk + 1) / 2 + l2 : l2 * (l2 + 1) / 2 + kk], NEMPC_RF64(l2, c), a);
                        tm[kk][c] = a;
                    }
#pragma unroll
                for (int a2 = 0; a2 < D; ++a2)
#pragma unroll
                    for (int c = 0; c <= a2; ++c) {
                        double a = 0.0;
#pragma unroll
                        for (int kk = 0; kk < D; ++kk) a = fma(NEMPC_RF64(kk, a2), tm[kk][c], a);
                        hs[p][a2 * (a2 + 1) / 2 + c] = a;
                    }
            }
            // + a_s sum_k J[p][k] h_{s-1}[k]   (h_{s-1} is read from scratch BEFORE it is overwritten)
#pragma unroll
            for (int e = 0; e < NS; ++e) {
                double hp[X];
#pragma unroll
                for (int kk = 0; kk < X; ++kk) hp[kk] = scr[(SC::HPREV_OFF + kk * NS + e) * sstride];
#pragma unroll
                for (int p = 0; p < X; ++p) {
                    double a = hs[p][e];
#pragma unroll
                    for (int kk = 0; kk < X; ++kk) a = fma(a_s * J[p][kk], hp[kk], a);
                    hs[p][e] = a;
                }
#pragma unroll
                for (int p = 0; p < X; ++p) {
                    scr[(SC::HPREV_OFF + p * NS + e) * sstride] = hs[p][e];
                    double* ha = scr + (SC::HACC_OFF + p * NS + e) * sstride;
                    *ha = fma(c_s, hs[p][e], *ha);
                }
            }
        }
#undef NEMPC_RF64
#pragma unroll
        for (int p = 0; p < X; ++p) { kacc[p] = fma(c_s, k[p], kacc[p]); kprev[p] = k[p]; }
        if (JAC && s + 1 < st.S) {
            const double an = st.a[s + 1];
#pragma unroll
            for (int p = 0; p < X; ++p)
#pragma unroll
                for (int c = 0; c < D; ++c) Rt[p][c] = fma(an, dk[p][c], (p == c) ? 1.0 : 0.0);
        }
    }

    // ---- outputs (slots of nempc_layout.h, same as every other kernel) -------------------------------------------------------------
    if (ar.resid) {
#pragma unroll
        for (int p = 0; p < X; ++p) {
            const double xt = (double)zb[t * X + p];
            const double xp = unity ? 0.0 : (double)((t == 0) ? ar.x0[b * X + p] : zb[(t - 1) * X + p]);
            ar.resid[b * L.m + t * X + p] = (TIO)(xp + kacc[p] - xt);
        }
    }
    if (JAC && ar.jac) {
        TIO* jv = ar.jac + b * L.nnz_jac;
#pragma unroll
        for (int p = 0; p < X; ++p) {
            jv[jac_slot_minus1(L, t, p)] = (TIO)-1;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const double v = dkacc[p][c] + ((!unity && c == p) ? 1.0 : 0.0);
                if (c < X) { if (t > 0) jv[jac_slot_A(L, t, p, c)] = (TIO)v; }
                else jv[jac_slot_B(L, t, p, c - X)] = (TIO)v;
            }
        }
    }
    if (HES && ar.hes) {
        TIO* hv = ar.hes + b * L.nnz_hes;
        const double sig = ar.sigma ? (double)ar.sigma[b] : (double)ar.sigma_scalar;
        double lam[X];
#pragma unroll
        for (int p = 0; p < X; ++p) lam[p] = (double)ar.lam[b * L.m + t * X + p];
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int c = 0; c <= a; ++c) {
                if (t == 0 && c < X) continue;
                double v = 0.0;
                if (NEMPC_FAST_CONTRACT_LAST) v = scr[(SC::HACC_OFF + a * (a + 1) / 2 + c) * sstride];          // contracted by the last stage
                else {
#pragma unroll
                    for (int p = 0; p < X; ++p) v = fma(lam[p], scr[(SC::HACC_OFF + p * NS + a * (a + 1) / 2 + c) * sstride], v);
                }
                int slot;
                if (a < X) {
                    slot = hes_slot_xx(L, t, a, c);
                    if (a == c && ar.quad) v += sig * 2.0 * ar.quad[(t - 1) * X + a];
                } else if (c < X) {
                    slot = hes_slot_ux(L, t, a - X, c);
                } else {
                    slot = hes_slot_uu(L, t, a - X, c - X);
                    if (a == c && ar.quad) v += sig * 2.0 * ar.quad[L.H * X + t * U + (a - X)];
                }
                hv[slot] = (TIO)v;
            }
        if (t == L.H - 1) {
#pragma unroll
            for (int p = 0; p < X; ++p)
                if (L.hes_last_slot[p] >= 0) hv[L.hes_last_slot[p]] = (TIO)(sig * 2.0 * ar.quad[(L.H - 1) * X + p]);
        }
    }
}

#ifndef NEMPC_FAST64_THREADS
#define NEMPC_FAST64_THREADS 128
#endif
#ifndef NEMPC_FAST64_SMEM_WEIGHTS
#define NEMPC_FAST64_SMEM_WEIGHTS 1      // copy the weight image to shared memory and read it with broadcast LDS.64 (0: constant bank -- the 26 KB image
                                         // overflows the immediate-constant cache: hit rate 73 %, 0.29 instead of 0.33 of the FP64 peak on C2)
#endif
template <int X, int U, int H1, int H2, int JC, int MODE, typename TIO>
__global__ void __launch_bounds__(NEMPC_FAST64_THREADS)
nempc_fast64_kernel(const __grid_constant__ Fast64Weights<X, U, H1, H2> w, const StageTable<double> st, const NlpLayout L, const EvalArgs<TIO> ar) {
    extern __shared__ __align__(16) unsigned char fast64_smem[];
#if NEMPC_FAST64_SMEM_WEIGHTS
    typedef Fast64Weights<X, U, H1, H2> FW;
    constexpr int WB = (int)((sizeof(FW) + 15) / 16 * 16);
    {
        const double* src = reinterpret_cast<const double*>(&w);
        double* dst = reinterpret_cast<double*>(fast64_smem);
        for (int i = threadIdx.x; i < (int)(sizeof(FW) / 8); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    const FW& ws = *reinterpret_cast<const FW*>(fast64_smem);
    double* scr = reinterpret_cast<double*>(fast64_smem + WB) + threadIdx.x;
    for (long long step = (long long)blockIdx.x * blockDim.x + threadIdx.x; step < ar.nsteps; step += (long long)gridDim.x * blockDim.x)
        fast64_step<X, U, H1, H2, JC, MODE, TIO>(ws, st, L, ar, step, scr, NEMPC_FAST64_THREADS);     // the launch uses exactly this block size
#else
    double* scr = reinterpret_cast<double*>(fast64_smem) + threadIdx.x;
    for (long long step = (long long)blockIdx.x * blockDim.x + threadIdx.x; step < ar.nsteps; step += (long long)gridDim.x * blockDim.x)
        fast64_step<X, U, H1, H2, JC, MODE, TIO>(w, st, L, ar, step, scr, NEMPC_FAST64_THREADS);
#endif
}
#endif  // __CUDACC__
