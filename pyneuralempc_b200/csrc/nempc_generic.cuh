// Generic per-step NLP-evaluation body: any layer count / width (<= limits), f32 or f64 arithmetic,
// discrete / unity / RK4 integrators, tanh / sigmoid / softplus.
//
// One "slot" = a group of `tps` threads that cooperates on one horizon step; per-step state lives in a
// workspace (shared memory when it fits, a global scratch otherwise).  Lanes map to output neurons, so
// weight reads W[i][j] are coalesced over j and activations are broadcast reads.
//
// The body is plain C++ over (lt, tps): with lt = 0, tps = 1 and no-op barriers it is also compiled for
// the host by tests/hostsim (a TEST-ONLY emulation used to validate indexing/maths on GPU-less CI; the
// product library never contains or calls it).
//
// Mathematics per step (reference integrator/rk4.py:113-285 generalised to any d):
//   stage s:  z_s = z + a_s E k_{s-1};  forward a_l, h_l, local tangents T_l = d a_l / d z_s
//             k_s = f(z_s), J_s = W_out^T diag(s') T_L            (model/tensorflow.py:49-75)
//             per-output adjoints G_l = d f / d h_l, local Hessians
//             M_p = sum_l T_l^T diag(s''(a_l) * G_l[:,p]) T_l     (model/tensorflow.py:77-109)
//             dk_s = J_s R_s;  h_s[p] = R_s^T M_p R_s + a_s sum_{k<x} J_s[p,k] h_{s-1}[k]
//             R_{s+1} = I + a_{s+1} E dk_s
//   pred = sum c_s k_s;  [A|B] = sum c_s dk_s;  Hblk[p] = sum c_s h_s[p];  Hc = sum_p lambda_p Hblk[p]
#pragma once
#include <math.h>
#include <stdint.h>

#include "nempc_layout.h"

#define NEMPC_MAXL 8
#define NEMPC_ACT_TANH_ 0
#define NEMPC_ACT_SIGMOID_ 1
#define NEMPC_ACT_SOFTPLUS_ 2
#define NEMPC_ACT_RELU_ 3

enum : int { NEMPC_WANT_JAC = 1, NEMPC_WANT_HES = 2, NEMPC_MODE_MODEL = 4, NEMPC_MODE_BLOCKS = 8, NEMPC_UNITY = 16 };

template <typename T> struct NetView {
    int L;                        // dense layers (hidden layers = L-1)
    int dims[NEMPC_MAXL + 1];     // dims[0] = d, dims[l+1] = fan-out of dense layer l
    int hoff[NEMPC_MAXL];         // neuron offset of hidden layer l inside the concatenated hidden vector
    int sum_h, hmax;
    int act;
    int tvp_dim, p_dim;           // exogenous inputs appended to (x, u) (model/tensorflow.py:39-47); W[0] then has d + tvp_dim + p_dim rows
    const T* W[NEMPC_MAXL];       // [in][out]   (Keras kernel layout)
    const T* WT[NEMPC_MAXL];      // [out][in]   (for the adjoint sweep: coalesced over `in`)
    const T* b[NEMPC_MAXL];
};

template <typename T> struct StageTable {
    int S;
    T a[4], c[4];
};

// a = (0, DT/2, DT/2, DT), c = DT/6 (1,2,2,1) for RK4 (integrator/rk4.py:69-78); one stage a=0, c=1 otherwise
template <typename T> inline StageTable<T> make_stage_table(bool rk4, double dt) {
    StageTable<T> st{};
    if (rk4) {
        st.S = 4;
        const double a[4] = {0.0, dt / 2.0, dt / 2.0, dt};
        const double c[4] = {dt / 6.0, dt / 3.0, dt / 3.0, dt / 6.0};
        for (int i = 0; i < 4; ++i) { st.a[i] = (T)a[i]; st.c[i] = (T)c[i]; }
    } else {
        st.S = 1;
        for (int i = 0; i < 4; ++i) { st.a[i] = (T)0; st.c[i] = (T)0; }
        st.c[0] = (T)1;
    }
    return st;
}

struct SlotLayout {
    int z, zs, lam, kprev, kcur, kacc, R, J, dk, dkacc, M, tmp, Hprev, Hacc, act, Tl, V0, V1, G0, G1, coef, ext, total;
};

inline SlotLayout make_slot_layout(int x, int d, int sum_h, int hmax, int n_ext = 0) {
    SlotLayout s;
    int o = 0;
    auto take = [&](int cnt) { int r = o; o += (cnt + 3) & ~3; return r; };
    s.z = take(d); s.zs = take(d); s.lam = take(x); s.kprev = take(x); s.kcur = take(x); s.kacc = take(x);
    s.R = take(d * d); s.J = take(x * d); s.dk = take(x * d); s.dkacc = take(x * d);
    s.M = take(x * d * d); s.tmp = take(x * d * d); s.Hprev = take(x * d * d); s.Hacc = take(x * d * d);
    s.act = take(sum_h); s.Tl = take(sum_h * d);
    s.V0 = take(hmax * d); s.V1 = take(hmax * d);
    s.G0 = take(hmax * x); s.G1 = take(hmax * x); s.coef = take(hmax * x);
    s.ext = take(n_ext);
    s.total = o;
    return s;
}

template <typename TIO> struct EvalArgs {
    const TIO* z;        // (B, n)   [MODEL mode: (N, d) stacked network inputs]
    const TIO* x0;       // (B, x)
    const TIO* lam;      // (B, m) or null
    const TIO* sigma;    // (B) or null
    double sigma_scalar;
    const double* quad;  // (n) objective diagonal weights or null
    TIO* resid;          // (B, m)
    TIO* jac;            // (B, nnz_jac)
    TIO* hes;            // (B, nnz_hes)
    TIO* pred;           // blocks / model: (N, x)
    TIO* AB;             // blocks / model: (N, x, d)
    TIO* Hblk;           // blocks / model: (N, x, d, d)
    long long nsteps;    // B*H  (MODEL mode: N)
    int flags;
    // exogenous model inputs (always double; not differentiated): tvp row of step (b, t) at tvp + b * tvp_bstride + t * tvp_dim
    // (MODEL mode: sample i at tvp + i * tvp_dim), p row of problem b at p + b * p_bstride; a stride of 0 shares one set between problems
    const double* tvp; const double* p;
    long long tvp_bstride, p_bstride;
    // optional gate (the batched solver): problem b is evaluated only where gate[b] == gate_value (null: every problem).  Honoured by the
    // register-resident f32 kernel and the objective kernel; the other kernels evaluate everything (their callers ignore the surplus).
    const int* gate; int gate_value;
};

template <typename A, typename B> struct WideOf { typedef double type; };
template <> struct WideOf<float, float> { typedef float type; };

// ---- four consecutive workspace values with one 16-byte (float) / two 16-byte (double) shared-memory loads ---------------------
// p must be 16-byte aligned: true for rows of the tangent buffers whenever d is a multiple of 4 (offsets are multiples of 4 elements)
template <typename T> NEMPC_HD void ld4(const T* p, T* v) { v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3]; }
#if defined(__CUDA_ARCH__)
template <> __device__ __forceinline__ void ld4<float>(const float* p, float* v) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
template <> __device__ __forceinline__ void ld4<double>(const double* p, double* v) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
#endif

// ---- slot barrier ------------------------------------------------------------------------------------
NEMPC_HD void slot_barrier(int bar_id, int tps) {
#if defined(__CUDA_ARCH__)
    if (tps == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(tps) : "memory");
#else
    (void)bar_id; (void)tps;
#endif
}

// ---- activations: value from pre-activation; derivatives recovered from the VALUE -----------------------
template <typename T> NEMPC_HD T act_value(int act, T a) {
    if (act == NEMPC_ACT_TANH_) return (T)tanh(a);
    if (act == NEMPC_ACT_SIGMOID_) return (T)1 / ((T)1 + (T)exp(-a));
    if (act == NEMPC_ACT_RELU_) return a > (T)0 ? a : (T)0;
    return a > (T)0 ? a + (T)log1p(exp(-a)) : (T)log1p(exp(a));   // softplus
}
template <> NEMPC_HD float act_value<float>(int act, float a) {
    if (act == NEMPC_ACT_TANH_) return tanhf(a);
    if (act == NEMPC_ACT_SIGMOID_) return 1.0f / (1.0f + expf(-a));
    if (act == NEMPC_ACT_RELU_) return a > 0.0f ? a : 0.0f;
    return a > 0.0f ? a + log1pf(expf(-a)) : log1pf(expf(a));
}
template <typename T> NEMPC_HD void act_derivs(int act, T h, T& s1, T& s2) {
    if (act == NEMPC_ACT_TANH_) { s1 = (T)1 - h * h; s2 = (T)-2 * h * s1; }
    else if (act == NEMPC_ACT_SIGMOID_) { s1 = h * ((T)1 - h); s2 = s1 * ((T)1 - (T)2 * h); }
    else if (act == NEMPC_ACT_RELU_) { s1 = h > (T)0 ? (T)1 : (T)0; s2 = (T)0; }   // piecewise linear: the Hessian VALUES are zero, its structure is kept
    else { s1 = (T)1 - (T)exp(-h); s2 = s1 * ((T)1 - s1); }          // softplus: sigmoid(a) = 1 - exp(-softplus(a))
}

// ---- one horizon step ------------------------------------------------------------------------------------
template <typename T, typename TIO, int DMAX>
NEMPC_HD void generic_step(const NetView<T>& net, const StageTable<T>& st, const NlpLayout& L, const SlotLayout& sl,
                           const EvalArgs<TIO>& ar, long long step, T* ws, int lt, int tps, int bar_id) {
    typedef typename WideOf<T, TIO>::type TW;
    // weight loads in flight per thread in the layer loops: wide-input networks (the C4 class, 256-wide layers whose weights live in
    // L2) are bound by the latency of those loads (ncu r1u: long_scoreboard 60 %), small ones prefer the shorter schedule
    constexpr int WU = DMAX >= 16 ? 16 : 4;
    const int x = L.x, d = L.d, dd = d * d;
    const int flags = ar.flags;
    const bool model_mode = (flags & NEMPC_MODE_MODEL) != 0;
    const bool want_hes = (flags & NEMPC_WANT_HES) != 0;
    const bool want_jac = want_hes || (flags & NEMPC_WANT_JAC) != 0;
    const bool unity = (flags & NEMPC_UNITY) != 0;
    // Single-stage integrators (discrete / unity) have no stage recursion, so the Lagrangian Hessian block is sum_p lambda_p M_p and
    // the multipliers can be contracted into the adjoint SEED: one adjoint column and one curvature matrix per step instead of x_dim
    // of each (12x less Hessian work for the quadrotor network of BASELINE config C4).  RK4 needs the per-output form (h_s recursion),
    // and so do the block / model entry points, which return per-output Hessians.
    const bool contract = want_hes && st.S == 1 && !model_mode && !(flags & NEMPC_MODE_BLOCKS);
    const int xh = contract ? 1 : x;
    const long long b = model_mode ? 0 : step / L.H;
    const int t = model_mode ? 0 : (int)(step - b * L.H);
    const TIO* zb = model_mode ? ar.z + step * d : ar.z + b * (long long)L.n;

    T* z = ws + sl.z; T* zs = ws + sl.zs; T* kprev = ws + sl.kprev; T* kcur = ws + sl.kcur; T* kacc = ws + sl.kacc;
    T* R = ws + sl.R; T* J = ws + sl.J; T* dk = ws + sl.dk; T* dkacc = ws + sl.dkacc;
    T* M = ws + sl.M; T* tmp = ws + sl.tmp; T* Hprev = ws + sl.Hprev; T* Hacc = ws + sl.Hacc;
    T* act = ws + sl.act; T* Tl = ws + sl.Tl; T* coef = ws + sl.coef;
    T* Vb[2] = {ws + sl.V0, ws + sl.V1};
    T* Gb[2] = {ws + sl.G0, ws + sl.G1};

    // ---- load the step's inputs, reset accumulators -------------------------------------------------------
    for (int c = lt; c < d; c += tps) {
        TIO v;
        if (model_mode) v = zb[c];
        else if (c < x) v = (t == 0) ? ar.x0[b * x + c] : zb[(t - 1) * x + c];
        else v = zb[L.H * x + t * L.u + (c - x)];
        z[c] = (T)v;
    }
    if (contract) for (int p = lt; p < x; p += tps) (ws + sl.lam)[p] = (T)ar.lam[b * L.m + t * x + p];
    const int n_ext = net.tvp_dim + net.p_dim;
    T* ext = ws + sl.ext;
    for (int e = lt; e < n_ext; e += tps)
        ext[e] = (T)(e < net.tvp_dim ? ar.tvp[(model_mode ? step * net.tvp_dim : b * ar.tvp_bstride + (long long)t * net.tvp_dim) + e]
                                     : ar.p[b * ar.p_bstride + (e - net.tvp_dim)]);
    for (int i = lt; i < x; i += tps) { kprev[i] = (T)0; kacc[i] = (T)0; }
    for (int i = lt; i < x * d; i += tps) dkacc[i] = (T)0;
    for (int i = lt; i < dd; i += tps) R[i] = (i / d == i % d) ? (T)1 : (T)0;
    if (want_hes) for (int i = lt; i < xh * dd; i += tps) { Hacc[i] = (T)0; Hprev[i] = (T)0; }
    slot_barrier(bar_id, tps);

    const int nh = net.L - 1;   // hidden layers
    for (int s = 0; s < st.S; ++s) {
        const T a_s = st.a[s], c_s = st.c[s];
        for (int c = lt; c < d; c += tps) zs[c] = z[c] + (c < x ? a_s * kprev[c] : (T)0);
        slot_barrier(bar_id, tps);

        // ---- first dense layer: d -> h_0 ; local tangent of a_0 is W_0 itself ---------------------------------
        {
            const int h0 = net.dims[1];
            const T* W = net.W[0]; const T* bb = net.b[0];
            T* Vout = Vb[0];
            for (int j = lt; j < h0; j += tps) {
                T a = bb[j];
                for (int c = 0; c < d; ++c) a += W[c * h0 + j] * zs[c];
                for (int e = 0; e < n_ext; ++e) a += W[(d + e) * h0 + j] * ext[e];     // tvp / p rows of the first layer: a bias shift
                const T h = act_value<T>(net.act, a);
                act[j] = h;
                if (want_jac) {
                    T s1, s2; act_derivs<T>(net.act, h, s1, s2);
                    for (int c = 0; c < d; ++c) { const T w = W[c * h0 + j]; Tl[j * d + c] = w; Vout[j * d + c] = s1 * w; }
                }
            }
        }
        slot_barrier(bar_id, tps);
        // ---- hidden layers 1 .. nh-1 -----------------------------------------------------------------------------
        for (int l = 1; l < nh; ++l) {
            const int hin = net.dims[l], hout = net.dims[l + 1];
            const T* W = net.W[l]; const T* bb = net.b[l];
            const T* hprev = act + net.hoff[l - 1];
            const T* Vin = Vb[(l - 1) & 1]; T* Vout = Vb[l & 1];
            for (int j = lt; j < hout; j += tps) {
                T acc = bb[j];
                T at[DMAX];
#pragma unroll
                for (int c = 0; c < DMAX; ++c) at[c] = (T)0;
                if (want_jac && (d & 3) == 0) {                           // tangent rows read as 16-byte vectors
#pragma unroll WU
                    for (int i = 0; i < hin; ++i) {                       // unrolled: WU independent weight loads (L2 latency) in flight
                        const T w = W[i * hout + j];
                        acc += w * hprev[i];
#pragma unroll
                        for (int q = 0; q < DMAX / 4; ++q)
                            if (4 * q < d) {
                                T v4[4];
                                ld4<T>(Vin + i * d + 4 * q, v4);
#pragma unroll
                                for (int k = 0; k < 4; ++k) at[4 * q + k] += w * v4[k];
                            }
                    }
                } else if (want_jac) {
#pragma unroll WU
                    for (int i = 0; i < hin; ++i) {
                        const T w = W[i * hout + j];
                        acc += w * hprev[i];
#pragma unroll
                        for (int c = 0; c < DMAX; ++c) if (c < d) at[c] += w * Vin[i * d + c];
                    }
                } else {
#pragma unroll 8
                    for (int i = 0; i < hin; ++i) acc += W[i * hout + j] * hprev[i];
                }
                const T h = act_value<T>(net.act, acc);
                act[net.hoff[l] + j] = h;
                if (want_jac) {
                    T s1, s2; act_derivs<T>(net.act, h, s1, s2);
#pragma unroll
                    for (int c = 0; c < DMAX; ++c) if (c < d) { Tl[(net.hoff[l] + j) * d + c] = at[c]; Vout[j * d + c] = s1 * at[c]; }
                }
            }
            slot_barrier(bar_id, tps);
        }
        // ---- linear output layer: value and local Jacobian -----------------------------------------------------
        {
            const int hin = net.dims[nh];
            const T* Wo = net.W[nh]; const T* bo = net.b[nh];
            const T* hl = act + net.hoff[nh - 1];
            const T* Vl = Vb[(nh - 1) & 1];
            const int items = want_jac ? x * (1 + d) : x;
            for (int idx = lt; idx < items; idx += tps) {
                const int p = want_jac ? idx / (1 + d) : idx, c = want_jac ? idx % (1 + d) : 0;
                if (c == 0) {
                    T acc = bo[p];
                    for (int j = 0; j < hin; ++j) acc += Wo[j * x + p] * hl[j];
                    kcur[p] = acc;
                } else {
                    T acc = (T)0;
                    for (int j = 0; j < hin; ++j) acc += Wo[j * x + p] * Vl[j * d + (c - 1)];
                    J[p * d + (c - 1)] = acc;
                }
            }
            if (want_hes) {
                if (contract) {
                    const T* lamv = ws + sl.lam;
                    for (int j = lt; j < hin; j += tps) {
                        T g = (T)0;
                        for (int p = 0; p < x; ++p) g += lamv[p] * Wo[j * x + p];
                        Gb[0][j] = g;
                    }
                } else {
                    for (int idx = lt; idx < hin * x; idx += tps) Gb[0][idx] = Wo[idx];
                }
                for (int idx = lt; idx < xh * dd; idx += tps) M[idx] = (T)0;
            }
        }
        slot_barrier(bar_id, tps);
        // ---- per-output adjoint sweep + curvature accumulation ---------------------------------------------------
        if (want_hes) {
            const int ntri = d * (d + 1) / 2;
            int cur = 0;
            for (int l = nh - 1; l >= 0; --l) {
                const int hl = net.dims[l + 1];
                T* G = Gb[cur]; T* Gn = Gb[cur ^ 1];
                const T* hv = act + net.hoff[l];
                for (int idx = lt; idx < hl * xh; idx += tps) {
                    T s1, s2; act_derivs<T>(net.act, hv[idx / xh], s1, s2);
                    const T g = G[idx];
                    coef[idx] = s2 * g;
                    G[idx] = s1 * g;
                }
                slot_barrier(bar_id, tps);
                const T* Tt = Tl + (long long)net.hoff[l] * d;
                for (int e = lt; e < xh * ntri; e += tps) {
                    const int p = e / ntri;
                    int r = e - p * ntri, c = 0;
                    while (r > c) { r -= c + 1; ++c; }       // r-th entry of the lower triangle -> (c, r)
                    const int c2 = r;
                    T acc = (T)0;
                    for (int j = 0; j < hl; ++j) acc += coef[j * xh + p] * Tt[j * d + c] * Tt[j * d + c2];
                    M[p * dd + c * d + c2] += acc;
                    if (c != c2) M[p * dd + c2 * d + c] += acc;
                }
                if (l > 0) {
                    const int hp = net.dims[l];
                    const T* WT = net.WT[l];
                    if (xh == 1) {                             // contracted adjoint: one column, no predicated DMAX-wide loop
                        for (int i = lt; i < hp; i += tps) {
                            T a0 = (T)0;
#pragma unroll WU
                            for (int j = 0; j < hl; ++j) a0 += WT[j * hp + i] * G[j];
                            Gn[i] = a0;
                        }
                    } else {
                        for (int i = lt; i < hp; i += tps) {
                            T ap[DMAX];
#pragma unroll
                            for (int p = 0; p < DMAX; ++p) ap[p] = (T)0;
                            if ((xh & 3) == 0) {                   // adjoint rows read as 16-byte vectors
#pragma unroll WU
                                for (int j = 0; j < hl; ++j) {
                                    const T w = WT[j * hp + i];
#pragma unroll
                                    for (int q = 0; q < DMAX / 4; ++q)
                                        if (4 * q < xh) {
                                            T g4[4];
                                            ld4<T>(G + j * xh + 4 * q, g4);
#pragma unroll
                                            for (int k = 0; k < 4; ++k) ap[4 * q + k] += w * g4[k];
                                        }
                                }
                            } else {
#pragma unroll WU
                                for (int j = 0; j < hl; ++j) {
                                    const T w = WT[j * hp + i];
#pragma unroll
                                    for (int p = 0; p < DMAX; ++p) if (p < xh) ap[p] += w * G[j * xh + p];
                                }
                            }
#pragma unroll
                            for (int p = 0; p < DMAX; ++p) if (p < xh) Gn[i * xh + p] = ap[p];
                        }
                    }
                }
                slot_barrier(bar_id, tps);
                cur ^= 1;
            }
        }
        // ---- stage algebra ---------------------------------------------------------------------------------------
        if (want_jac) {
            for (int idx = lt; idx < x * d; idx += tps) {
                const int p = idx / d, c = idx % d;
                T acc = (T)0;
                for (int k = 0; k < d; ++k) acc += J[p * d + k] * R[k * d + c];
                dk[idx] = acc;
            }
            if (want_hes) {
                for (int idx = lt; idx < xh * dd; idx += tps) {     // tmp[p] = M[p] R
                    const int p = idx / dd, k = (idx % dd) / d, c = idx % d;
                    T acc = (T)0;
                    for (int l2 = 0; l2 < d; ++l2) acc += M[p * dd + k * d + l2] * R[l2 * d + c];
                    tmp[idx] = acc;
                }
            }
            slot_barrier(bar_id, tps);
            if (want_hes) {
                for (int idx = lt; idx < xh * dd; idx += tps) {     // h_s[p] = R^T tmp[p] + a_s sum_k J[p,k] h_{s-1}[k]
                    const int p = idx / dd, a = (idx % dd) / d, c = idx % d;
                    T acc = (T)0;
                    for (int k = 0; k < d; ++k) acc += R[k * d + a] * tmp[p * dd + k * d + c];
                    if (s > 0) for (int k = 0; k < x; ++k) acc += a_s * J[p * d + k] * Hprev[k * dd + a * d + c];
                    M[idx] = acc;
                }
            }
            for (int idx = lt; idx < x * d; idx += tps) dkacc[idx] += c_s * dk[idx];
        }
        for (int idx = lt; idx < x; idx += tps) { kacc[idx] += c_s * kcur[idx]; kprev[idx] = kcur[idx]; }
        slot_barrier(bar_id, tps);
        if (want_hes)
            for (int idx = lt; idx < xh * dd; idx += tps) { Hacc[idx] += c_s * M[idx]; Hprev[idx] = M[idx]; }
        if (want_jac && s + 1 < st.S) {
            const T an = st.a[s + 1];
            for (int idx = lt; idx < dd; idx += tps) {
                const int k = idx / d, c = idx % d;
                R[idx] = (k == c ? (T)1 : (T)0) + (k < x ? an * dk[k * d + c] : (T)0);
            }
        }
        slot_barrier(bar_id, tps);
    }

    // ---- outputs ------------------------------------------------------------------------------------------------
    if (model_mode || (flags & NEMPC_MODE_BLOCKS)) {
        const bool addI = !model_mode && !unity;
        if (ar.pred) for (int p = lt; p < x; p += tps) ar.pred[step * x + p] = (TIO)((TW)kacc[p] + (addI ? (TW)z[p] : (TW)0));
        if (ar.AB) for (int i = lt; i < x * d; i += tps)
            ar.AB[step * x * d + i] = (TIO)((TW)dkacc[i] + ((addI && i / d == i % d) ? (TW)1 : (TW)0));
        if (ar.Hblk) for (int i = lt; i < x * dd; i += tps) ar.Hblk[step * x * dd + i] = (TIO)Hacc[i];
        return;
    }
    if (ar.resid) {
        for (int p = lt; p < x; p += tps) {
            const TW xt = (TW)zb[t * x + p];
            const TW xp = unity ? (TW)0 : (TW)((t == 0) ? ar.x0[b * x + p] : zb[(t - 1) * x + p]);
            ar.resid[b * L.m + t * x + p] = (TIO)(xp + (TW)kacc[p] - xt);
        }
    }
    if (ar.jac && want_jac) {
        TIO* jv = ar.jac + b * L.nnz_jac;
        for (int idx = lt; idx < x * (d + 1); idx += tps) {
            const int p = idx / (d + 1), c = idx % (d + 1);
            if (c == d) { jv[jac_slot_minus1(L, t, p)] = (TIO)-1; continue; }
            const TW v = (TW)dkacc[p * d + c] + ((!unity && c == p) ? (TW)1 : (TW)0);
            if (c < x) { if (t > 0) jv[jac_slot_A(L, t, p, c)] = (TIO)v; }
            else jv[jac_slot_B(L, t, p, c - x)] = (TIO)v;
        }
    }
    if (ar.hes && want_hes) {
        TIO* hv = ar.hes + b * L.nnz_hes;
        const TW sig = ar.sigma ? (TW)ar.sigma[b] : (TW)ar.sigma_scalar;
        T* lam = ws + sl.lam;
        if (!contract) for (int p = lt; p < x; p += tps) lam[p] = (T)ar.lam[b * L.m + t * x + p];
        slot_barrier(bar_id, tps);
        for (int idx = lt; idx < dd; idx += tps) {
            const int a = idx / d, c = idx % d;
            if (c > a) continue;
            if (t == 0 && c < x) continue;                      // x0 is data, not a variable (discret.py:70-78)
            T acc = (T)0;
            if (contract) acc = Hacc[a * d + c];                // the multipliers went into the adjoint seed
            else for (int p = 0; p < x; ++p) acc += lam[p] * Hacc[p * dd + a * d + c];
            TW v = (TW)acc;
            int slot;
            if (a < x) {
                slot = hes_slot_xx(L, t, a, c);
                if (a == c && ar.quad) v += sig * (TW)2 * (TW)ar.quad[(t - 1) * x + a];
            } else if (c < x) {
                slot = hes_slot_ux(L, t, a - x, c);
            } else {
                slot = hes_slot_uu(L, t, a - x, c - x);
                if (a == c && ar.quad) v += sig * (TW)2 * (TW)ar.quad[L.H * x + t * L.u + (a - x)];
            }
            hv[slot] = (TIO)v;
        }
        if (t == L.H - 1)                                        // objective-only diagonal of x_H
            for (int p = lt; p < x; p += tps)
                if (L.hes_last_slot[p] >= 0) hv[L.hes_last_slot[p]] = (TIO)(sig * (TW)2 * (TW)ar.quad[(L.H - 1) * x + p]);
    }
}
