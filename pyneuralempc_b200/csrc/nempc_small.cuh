// Low-latency kernel for SMALL batches of the Lotka-Volterra class (two hidden layers of <= 32 neurons, x_dim + u_dim = 3):
// ONE WARP PER HORIZON STEP, lane j = hidden neuron j of both layers.
//
// Why: nempc_fast.cuh gives every horizon step to one thread; that is the throughput-optimal mapping for >= 10^5 steps, but a
// single IPOPT callback of the reference (BASELINE config C1: one problem, H = 25) is 25 steps, and one thread needs ~56 us for
// an RK4 step (a dependent chain of ~2x10^4 FMAs).  Here the 30-wide layers are spread over the lanes of a warp: a layer is a
// loop of register-to-register broadcasts (__shfl_sync) against the lane's own weight column, sums over neurons are butterfly
// reductions, and the (tiny) RK4 stage algebra is replicated on all lanes.  Same mathematics as nempc_fast.cuh /
// nempc_generic.cuh (reference integrator/rk4.py:113-285, model/tensorflow.py:49-109); summation order differs (tree instead of
// sequential), so results agree with the other kernels to float32 rounding, not bit for bit.
#pragma once
#include "nempc_fast.cuh"
#include "nempc_generic.cuh"

#if defined(__CUDACC__)
__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// W1 [D][H1], b1 [H1], W2 [H1][H2], b2 [H2], W3 [H2][X], b3 [X]: the float32 device copies of the generic kernel
template <int X, int U, int H1, int H2, int MODE, typename TIO>
__global__ void __launch_bounds__(128)
nempc_small_kernel(const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                   const float* __restrict__ W3, const float* __restrict__ b3, const StageTable<float> st, const NlpLayout L,
                   const EvalArgs<TIO> ar) {
    static_assert(H1 <= 32 && H2 <= 32 && X == 2 && X + U == 3, "one lane per neuron; the stage algebra below is written for x = 2, d = 3");
    constexpr int D = X + U, NS = D * (D + 1) / 2;
    constexpr bool JAC = MODE >= 1, HES = MODE >= 2;
    typedef typename WideOf<float, TIO>::type TW;
    const int lane = threadIdx.x & 31;
    const long long step = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (step >= ar.nsteps) return;                                   // whole warps leave together
    const bool unity = (ar.flags & NEMPC_UNITY) != 0;
    const long long b = step / L.H;
    const int t = (int)(step - b * L.H);
    const TIO* zb = ar.z + b * (long long)L.n;

    // ---- this lane's weights (zero for lanes beyond the layer width, so they drop out of every sum) ----------------------------
    const bool in1 = lane < H1, in2 = lane < H2;
    float w1[D], w2col[H1], w3[X], w23row[HES ? H2 : 1][X];
#pragma unroll
    for (int c = 0; c < D; ++c) w1[c] = in1 ? W1[c * H1 + lane] : 0.f;
    const float bb1 = in1 ? b1[lane] : 0.f, bb2 = in2 ? b2[lane] : 0.f;
#pragma unroll
    for (int i = 0; i < H1; ++i) w2col[i] = in2 ? W2[i * H2 + lane] : 0.f;          // column `lane` of W2: layer-2 forward
#pragma unroll
    for (int p = 0; p < X; ++p) w3[p] = in2 ? W3[lane * X + p] : 0.f;
    if (HES) {
#pragma unroll
        for (int j = 0; j < H2; ++j)                                                     // row `lane` of W2 times W3: layer-1 adjoint
#pragma unroll
            for (int p = 0; p < X; ++p) w23row[HES ? j : 0][p] = in1 ? W2[lane * H2 + j] * W3[j * X + p] : 0.f;
    }
    float p1[NS];                                                                        // W1[c] W1[c2] of this neuron
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
        for (int c2 = 0; c2 <= c; ++c2) p1[c * (c + 1) / 2 + c2] = w1[c] * w1[c2];

    float z[D];
#pragma unroll
    for (int c = 0; c < X; ++c) z[c] = (float)((t == 0) ? ar.x0[b * X + c] : zb[(t - 1) * X + c]);
#pragma unroll
    for (int c = 0; c < U; ++c) z[X + c] = (float)zb[L.H * X + t * U + c];

    float Rt[X][D], kprev[X], kacc[X], dkacc[X][D], hprev[X][NS], hacc[X][NS];
#pragma unroll
    for (int p = 0; p < X; ++p) {
        kprev[p] = 0.f; kacc[p] = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) { Rt[p][c] = (p == c) ? 1.f : 0.f; dkacc[p][c] = 0.f; }
#pragma unroll
        for (int e = 0; e < NS; ++e) { hprev[p][e] = 0.f; hacc[p][e] = 0.f; }
    }

#pragma unroll 1
    for (int s = 0; s < st.S; ++s) {
        const float a_s = st.a[s], c_s = st.c[s];
        float zs[D];
#pragma unroll
        for (int c = 0; c < D; ++c) zs[c] = (c < X) ? fmaf(a_s, kprev[c < X ? c : 0], z[c]) : z[c];
        // ---- layer 1: this lane's neuron -------------------------------------------------------------------------------------
        float a1 = bb1;
#pragma unroll
        for (int c = 0; c < D; ++c) a1 = fmaf(w1[c], zs[c], a1);
        const float t1 = in1 ? fast_tanh(a1) : 0.f;
        const float sp1 = fmaf(-t1, t1, 1.f);
        float v1[D];
#pragma unroll
        for (int c = 0; c < D; ++c) v1[c] = sp1 * w1[c];
        // ---- layer 2: broadcast layer-1 neuron i, multiply with this lane's weight column ------------------------------------------
        float a2 = bb2, tg[D];
#pragma unroll
        for (int c = 0; c < D; ++c) tg[c] = 0.f;
#pragma unroll
        for (int i = 0; i < H1; ++i) {
            a2 = fmaf(w2col[i], __shfl_sync(0xffffffffu, t1, i), a2);
            if (JAC) {
#pragma unroll
                for (int c = 0; c < D; ++c) tg[c] = fmaf(w2col[i], __shfl_sync(0xffffffffu, v1[c], i), tg[c]);
            }
        }
        const float t2 = in2 ? fast_tanh(a2) : 0.f;
        const float sp2 = fmaf(-t2, t2, 1.f), spp2 = -2.f * t2 * sp2;
        // ---- output layer: sums over the lanes --------------------------------------------------------------------------------------
        float k[X], J[X][D], M[X][NS];
        float pp[NS];
        if (HES) {
#pragma unroll
            for (int c = 0; c < D; ++c)
#pragma unroll
                for (int c2 = 0; c2 <= c; ++c2) pp[c * (c + 1) / 2 + c2] = tg[c] * tg[c2];
        }
#pragma unroll
        for (int p = 0; p < X; ++p) {
            k[p] = warp_sum_f32(w3[p] * t2) + b3[p];
            if (JAC) {
#pragma unroll
                for (int c = 0; c < D; ++c) J[p][c] = warp_sum_f32(w3[p] * sp2 * tg[c]);
            }
        }
        if (HES) {
            // layer-1 adjoint g[p] of THIS lane's neuron (sum over layer-2 neurons j), then both curvature terms in one reduction
            float g[X];
#pragma unroll
            for (int p = 0; p < X; ++p) g[p] = 0.f;
#pragma unroll
            for (int j = 0; j < H2; ++j) {
                const float sj = __shfl_sync(0xffffffffu, sp2, j);
#pragma unroll
                for (int p = 0; p < X; ++p) g[p] = fmaf(sj, w23row[HES ? j : 0][p], g[p]);
            }
            const float spp1 = -2.f * t1 * sp1;
#pragma unroll
            for (int p = 0; p < X; ++p) {
                const float q2 = spp2 * w3[p], q1 = spp1 * g[p];
#pragma unroll
                for (int e = 0; e < NS; ++e) M[p][e] = warp_sum_f32(fmaf(q2, pp[e], q1 * p1[e]));
            }
        }
        // ---- stage algebra, replicated on every lane (all lanes hold the same sums) ---------------------------------------------------
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
        if (HES) {
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
                for (int a2i = 0; a2i < D; ++a2i)
#pragma unroll
                    for (int c = 0; c <= a2i; ++c) {
                        float a = 0.f;
#pragma unroll
                        for (int kk = 0; kk < D; ++kk) a = fmaf(NEMPC_RF(kk, a2i), tm[kk][c], a);
                        hs[p][a2i * (a2i + 1) / 2 + c] = a;
                    }
            }
#pragma unroll
            for (int e = 0; e < NS; ++e) {
                float hn[X];
#pragma unroll
                for (int p = 0; p < X; ++p) {
                    float a = hs[p][e];
#pragma unroll
                    for (int kk = 0; kk < X; ++kk) a = fmaf(a_s * J[p][kk], hprev[kk][e], a);
                    hn[p] = a;
                }
#pragma unroll
                for (int p = 0; p < X; ++p) { hprev[p][e] = hn[p]; hacc[p][e] = fmaf(c_s, hn[p], hacc[p][e]); }
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

    // ---- outputs (lane 0; same slots as nempc_fast.cuh) ---------------------------------------------------------------------------------
    if (lane != 0) return;
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
                for (int p = 0; p < X; ++p) acc = fmaf(lam[p], hacc[p][a * (a + 1) / 2 + c], acc);
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
#endif  // __CUDACC__
