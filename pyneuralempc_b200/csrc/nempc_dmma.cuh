// Float64 tensor-core path for wide tanh networks (hidden layers all HW = 128 or 64 wide): the reference assembles in float64
// (optimizer/ipopt.py:66-86) and a Keras model built with float64 layers evaluates in it; this is the 1e-10 mode of the C3 class
// (cart-pole 5 -> 128 x 3 -> 4) at tensor-core speed.  B200 keeps full-rate FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64: the
// tcgen05 family has no f64 kind), so the hidden-to-hidden layers run as warp-level DMMA GEMMs; everything else is DFMA.
//
// Formulation: FORWARD second order, as nempc_tc.cuh -- per horizon step and integrator stage the network is evaluated on a stack of
// RPS = 1 + d + d(d+1)/2 rows (activations, first-order tangents V[., c], second-order tangents Q[., (c,c2)]); every row goes through the
// same weights, so a hidden-to-hidden layer is ONE GEMM [64 rows x HW] x [HW x HW] for SPT = 64 / RPS steps at once:
//   * one persistent CTA per SM, TWO GROUPS of 4 warps, each group with its own 64-row tile A (in place in shared memory for the life of a
//     tile), weight ring and named barrier -- while one group runs an epilogue the other keeps the DMMA pipe busy;
//   * the weights of the layer stream from L2 through a two-slot cp.async ring of 16-row chunks per group (a layer is 128 KB: tiles and
//     weights do not fit together), one group barrier per chunk;
//   * warp tile 32 x 64 at HW = 128: 32 m8n8k4 accumulators (128 registers) per lane, 12 shared-memory fragment loads per 32 DMMAs;
//   * epilogue per (step, neuron): bias, tanh, s', s'', V = s' dV, Q = s' dQ + s'' dV_c dV_c2, written back as the next operand tile;
//   * the first layer (K = d) is DFMA; the output layer (N = x <= 8) is one more DMMA pass over the tile against W_out padded to 8 columns,
//     whose accumulator fragments go straight to global memory (a shuffle-reduced DFMA contraction took 29 % of the kernel, profiles/r2r).
// The kernel is a pure NETWORK kernel: zin (N, d) -> k = f(zin) (N, x), local Jacobian J (N, x, d), per-output local Hessians M (N, x, NS)
// (lower triangle e = c(c+1)/2 + c2).  The integrator's stage algebra (dk = J R, h_s = R^T M R + a_s J h_{s-1}, R_{s+1} = I + a_{s+1} E dk;
// reference integrator/rk4.py:113-285, discret.py:32-81, unity.py:34-81) and the scatter into the block-banded value arrays run in a second,
// thread-per-step kernel between the stages (nempc_dmma_stage_kernel); the few hundred bytes per step that travel between the two kernels
// through L2 are noise next to 5.6 MFLOP per step.
#pragma once
#include "nempc_generic.cuh"

struct DmmaNet {
    int d, x, nhid;                                   // inputs, outputs, hidden layers (all HW wide)
    const double* W[NEMPC_MAXL];                      // [in][out] (Keras kernel layout), W[0] (d x HW), W[1..nhid-1] (HW x HW), W[nhid] (HW x x)
    const double* b[NEMPC_MAXL];
};

template <int HW_> struct DmmaCfg {
    // TWO GROUPS of 128 threads per CTA, each with its own 64-row tile, weight ring and named barrier: while one group runs an epilogue
    // (tanh, chain rule: DFMA) the other keeps the DMMA pipe busy -- the overlap a second CTA per SM would give (a second CTA does not fit:
    // registers).  One tile of 128 rows measured 72 % DMMA-pipe activity, serialised GEMM / epilogue phases (profiles/r2r).
    static constexpr int HW = HW_, NG = 2, GT = 128, THREADS = NG * GT, MT = 64;   // MT: rows per GROUP tile
    static constexpr int LDA = HW + 4;                // row stride of the tile (doubles): a half-warp's fragment loads (4 rows x 4 columns) fall on distinct banks
    static constexpr int LDW = HW + 4;                // row stride of a weight chunk: likewise (HW + 8 measured a two-way conflict)
    static constexpr int LDO = 12, O_DOUBLES = HW * LDO;  // output weights, zero-padded to 8 columns (row stride 12: conflict-free fragments)
    static constexpr int KC = 16, NSLOT = 2;          // ring: chunks of 16 weight rows, two slots per group (chunk kc + 1 streams in under the DMMAs of chunk kc: 8 group barriers per layer; 8-row chunks in three slots measured the same)
    static constexpr int WGN = HW / 64, WGM = 4 / WGN;          // warp grid of a group (4 warps): warp tile (MT / WGM) x 64
    static constexpr int WM = MT / WGM, WN = 64, MI = WM / 8, NI = WN / 8;
    static constexpr int A_DOUBLES = MT * LDA, RING_DOUBLES = NSLOT * KC * LDW;      // per group
    static constexpr int Z_DOUBLES = MT * 8;          // inputs of the tile's steps (d <= 8), per group
    static constexpr int G_DOUBLES = A_DOUBLES + RING_DOUBLES + Z_DOUBLES;
    static constexpr size_t SMEM = (size_t)(NG * G_DOUBLES + O_DOUBLES) * sizeof(double);
    static_assert(HW == 128 || HW == 64, "hidden width 128 or 64");
    static_assert(SMEM <= 232448, "shared-memory map exceeds 227 KB");
};

#if defined(__CUDACC__)
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma_cp_async16(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void dmma_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void dmma_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// rows: 1 (mode 0), 1 + d (mode 1), 1 + d + d(d+1)/2 (mode 2) per step, kind-major inside the tile: row = kind * spt + step
template <class C, int D>
__global__ void __launch_bounds__(C::THREADS, 1)
nempc_dmma_net_kernel(const DmmaNet net, const double* __restrict__ zin, long long N, int mode, double* __restrict__ fo, double* __restrict__ Jo,
                      double* __restrict__ Mo) {
    constexpr int HW = C::HW, MT = C::MT, LDA = C::LDA, LDW = C::LDW, KC = C::KC, NSLOT = C::NSLOT, MI = C::MI, NI = C::NI, GT = C::GT;
    static_assert(D <= 8, "input width");
    extern __shared__ __align__(16) unsigned char dmma_smem[];
    const int grp = threadIdx.x / GT, tid = threadIdx.x - grp * GT, lane = tid & 31, warp = tid >> 5;      // group-local thread / warp index
    double* A = reinterpret_cast<double*>(dmma_smem) + grp * C::G_DOUBLES;
    double* ring = A + C::A_DOUBLES;
    double* zs = ring + C::RING_DOUBLES;
    double* wos = reinterpret_cast<double*>(dmma_smem) + C::NG * C::G_DOUBLES;
    constexpr int d = D, ns = D * (D + 1) / 2;
    const int x = net.x;
    const int rps = 1 + (mode >= 1 ? d : 0) + (mode >= 2 ? ns : 0);
    const int spt = MT / rps, rows = spt * rps;
    const int wm0 = (warp / C::WGN) * C::WM, wn0 = (warp % C::WGN) * C::WN;
    const int g = lane >> 2, q = lane & 3;
    auto gsync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(C::GT) : "memory"); };        // barrier of this group only
    for (int i = tid; i < C::A_DOUBLES; i += GT) A[i] = 0.0;                         // padding rows stay zero for the life of the CTA
    for (int i = threadIdx.x; i < HW * C::LDO; i += C::THREADS) { const int j = i / C::LDO, p = i - j * C::LDO; wos[i] = p < x ? net.W[net.nhid][(size_t)j * x + p] : 0.0; }
    __syncthreads();                                                                 // the only CTA-wide barrier: the groups run on their own from here
    const long long ntiles = (N + spt - 1) / spt;
    for (long long tile = (long long)blockIdx.x * C::NG + grp; tile < ntiles; tile += (long long)gridDim.x * C::NG) {
        const long long s0 = tile * spt;
        const int nst = (int)((N - s0) < spt ? (N - s0) : spt);
        gsync();                                                                     // previous tile's output pass is done with A
        for (int i = tid; i < spt * D; i += GT) zs[i] = (i < nst * D) ? zin[s0 * D + i] : 0.0;
        gsync();
        // ---- first layer (K = d): DFMA ------------------------------------------------------------------------------------------
        for (int idx = tid; idx < spt * HW; idx += GT) {
            const int sl = idx / HW, j = idx - sl * HW;
            double a = net.b[0][j], w0[D];
#pragma unroll
            for (int c = 0; c < D; ++c) { w0[c] = net.W[0][c * HW + j]; a = fma(w0[c], zs[sl * D + c], a); }
            const double h = tanh(a), sp = fma(-h, h, 1.0), spp = -2.0 * h * sp;
            double* col = A + sl * LDA + j;
            const int rs = spt * LDA;                                                // distance between the row kinds of one step
            col[0] = h;
            if (mode >= 1) {
#pragma unroll
                for (int c = 0; c < D; ++c) col[(1 + c) * rs] = sp * w0[c];
                if (mode >= 2) {
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        const double sw = spp * w0[c];
#pragma unroll
                        for (int c2 = 0; c2 <= c; ++c2) col[(1 + D + c * (c + 1) / 2 + c2) * rs] = sw * w0[c2];
                    }
                }
            }
        }
        // ---- hidden-to-hidden layers: DMMA GEMM + epilogue --------------------------------------------------------------------------
        for (int l = 1; l < net.nhid; ++l) {
            const double* __restrict__ Wl = net.W[l];
            double acc[MI][NI][2];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
            auto issue_chunk = [&](int kc) {                                         // weight rows [kc KC, +KC) -> ring slot kc % NSLOT
                double* dst = ring + (kc % NSLOT) * (KC * LDW);
                constexpr int V = HW / 2;                                            // 16-byte pieces per row
                for (int i = tid; i < KC * V; i += GT) {
                    const int r = i / V, v = i - r * V;
                    dmma_cp_async16(dst + r * LDW + 2 * v, Wl + (size_t)(kc * KC + r) * HW + 2 * v);
                }
                dmma_cp_commit();
            };
            constexpr int NCH = HW / KC;
            issue_chunk(0);
            for (int kc = 0; kc < NCH; ++kc) {
                dmma_cp_wait<0>();
                gsync();                                    // chunk kc has landed for every thread of the group; everyone is done with chunk kc - 1 (and, for kc = 0, with writing A)
                if (kc + 1 < NCH) issue_chunk(kc + 1);      // its slot held chunk kc - 1
                const double* wch = ring + (kc % NSLOT) * (KC * LDW);
#pragma unroll
                for (int kk = 0; kk < KC / 4; ++kk) {
                    double af[MI], bf[NI];
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi) af[mi] = A[(wm0 + mi * 8 + g) * LDA + kc * KC + kk * 4 + q];
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni) bf[ni] = wch[(kk * 4 + q) * LDW + wn0 + ni * 8 + g];
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                        for (int ni = 0; ni < NI; ++ni) dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
                }
            }
            gsync();                                        // every warp has read its A fragments: the tile can be overwritten
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni)
                    *reinterpret_cast<double2*>(&A[(wm0 + mi * 8 + g) * LDA + wn0 + ni * 8 + 2 * q]) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            gsync();
            const double* __restrict__ bl = net.b[l];
            // IL steps of one column in flight per thread: the chain tanh -> s' -> s'' is ~100 dependent DFMA-class instructions and an epilogue
            // has only this group's four warps to hide them (measured neutral on C3: the kernel is bound by the DMMA pipe, 3/4 busy)
            constexpr int IL = (D <= 6) ? 2 : 1;
            for (int idx = tid; idx < spt * HW; idx += IL * GT) {
                double* col[IL]; bool on[IL];
                double a[IL], tg[IL][D], qv[IL][ns];
                const int rs = spt * LDA;
#pragma unroll
                for (int u = 0; u < IL; ++u) {
                    const int id = idx + u * GT;
                    on[u] = id < spt * HW;
                    const int sl = on[u] ? id / HW : 0, j = on[u] ? id - sl * HW : 0;
                    col[u] = A + sl * LDA + j;
                    a[u] = col[u][0] + bl[j];
                    if (mode >= 1) {
#pragma unroll
                        for (int c = 0; c < D; ++c) tg[u][c] = col[u][(1 + c) * rs];
                        if (mode >= 2) {
#pragma unroll
                            for (int e = 0; e < ns; ++e) qv[u][e] = col[u][(1 + D + e) * rs];
                        }
                    }
                }
                double h[IL], sp[IL], spp[IL];
#pragma unroll
                for (int u = 0; u < IL; ++u) { h[u] = tanh(a[u]); sp[u] = fma(-h[u], h[u], 1.0); spp[u] = -2.0 * h[u] * sp[u]; }
#pragma unroll
                for (int u = 0; u < IL; ++u) {
                    if (!on[u]) continue;
                    col[u][0] = h[u];
                    if (mode >= 1) {
#pragma unroll
                        for (int c = 0; c < D; ++c) col[u][(1 + c) * rs] = sp[u] * tg[u][c];
                        if (mode >= 2) {
#pragma unroll
                            for (int c = 0; c < D; ++c) {
                                const double st = spp[u] * tg[u][c];
#pragma unroll
                                for (int c2 = 0; c2 <= c; ++c2) { const int e = c * (c + 1) / 2 + c2; col[u][(1 + D + e) * rs] = fma(sp[u], qv[u][e], st * tg[u][c2]); }
                            }
                        }
                    }
                }
            }
            // (the group barrier at the top of the next layer's first chunk / before the output pass orders these writes)
        }
        gsync();
        // ---- output layer (N = x <= 8): one DMMA pass, two row blocks of 8 per warp, two K halves each (four independent accumulator chains) ----
        {
            double oc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
            const int mt0 = warp * 2;                                                // MT / 8 = 8 row blocks over the group's 4 warps
#pragma unroll 4
            for (int k4 = 0; k4 < HW / 8; ++k4) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int kb = (hf * (HW / 8) + k4) * 4;
                    const double bfr = wos[(kb + q) * C::LDO + g];
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi) dmma_m8n8k4(oc[mi][hf][0], oc[mi][hf][1], A[((mt0 + mi) * 8 + g) * LDA + kb + q], bfr);
                }
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const int r = (mt0 + mi) * 8 + g;
                const int kind = r / spt, sl = r - kind * spt;
                if (r < rows && sl < nst) {
                    const long long st = s0 + sl;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int p = 2 * q + i;
                        if (p < x) {
                            const double v = oc[mi][0][i] + oc[mi][1][i];
                            if (kind == 0) fo[st * x + p] = v + net.b[net.nhid][p];
                            else if (kind <= d) Jo[(st * x + p) * d + (kind - 1)] = v;
                            else Mo[(st * x + p) * ns + (kind - 1 - d)] = v;
                        }
                    }
                }
            }
        }
    }
}

// ---- stage algebra + scatter: one thread per horizon step -----------------------------------------------------------------------------
// per-step state that lives across the stages (doubles): kprev X | kacc X | Rt X*D | dkacc X*D | hprev 2 x X*NS (stage s writes half s & 1 and
// reads h_{s-1} of every output from the other half) | hacc X*NS
template <int X, int U> struct DmmaState {
    static constexpr int D = X + U, NS = D * (D + 1) / 2;
    static constexpr int KPREV = 0, KACC = X, RT = 2 * X, DKACC = RT + X * D, HPREV = DKACC + X * D, HACC = HPREV + 2 * X * NS, COUNT = HACC + X * NS;
};

// stage < 0: initialise (zin = z of the step); otherwise consume the network outputs of stage `stage`
template <int X, int U>
__global__ void __launch_bounds__(128)
nempc_dmma_stage_kernel(const StageTable<double> st, const NlpLayout L, const EvalArgs<double> ar, long long base, long long N, int mode, int stage,
                        double* __restrict__ zin, const double* __restrict__ fo, const double* __restrict__ Jo, const double* __restrict__ Mo,
                        double* __restrict__ state) {
    typedef DmmaState<X, U> S;
    constexpr int D = S::D, NS = S::NS;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const long long step = base + i;
    const long long b = step / L.H;
    const int t = (int)(step - b * L.H);
    const double* zb = ar.z + b * (long long)L.n;
    double z[D];
#pragma unroll
    for (int c = 0; c < X; ++c) z[c] = (t == 0) ? ar.x0[b * X + c] : zb[(t - 1) * X + c];
#pragma unroll
    for (int c = 0; c < U; ++c) z[X + c] = zb[L.H * X + t * U + c];
    double* sp = state + i * S::COUNT;
    if (stage < 0) {
#pragma unroll
        for (int c = 0; c < D; ++c) zin[i * D + c] = z[c];
        return;
    }
    const bool JAC = mode >= 1, HES = mode >= 2, first = stage == 0, last = stage + 1 == st.S;
    const double a_s = st.a[stage], c_s = st.c[stage];
    double k[X], J[X][D], Rt[X][D];
#pragma unroll
    for (int p = 0; p < X; ++p) {
        k[p] = fo[i * X + p];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            J[p][c] = JAC ? Jo[(i * X + p) * D + c] : 0.0;
            Rt[p][c] = first ? (p == c ? 1.0 : 0.0) : sp[S::RT + p * D + c];
        }
    }
#define NEMPC_RF(kk, cc) ((kk) < X ? Rt[(kk) < X ? (kk) : 0][cc] : ((kk) == (cc) ? 1.0 : 0.0))
    double dk[X][D], dkacc[X][D];
    if (JAC) {
#pragma unroll
        for (int p = 0; p < X; ++p)
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double a = (c >= X) ? J[p][c] : 0.0;
#pragma unroll
                for (int kk = 0; kk < X; ++kk) a = fma(J[p][kk], Rt[kk][c], a);
                dk[p][c] = a;
                dkacc[p][c] = fma(c_s, a, first ? 0.0 : sp[S::DKACC + p * D + c]);
                if (!last) sp[S::DKACC + p * D + c] = dkacc[p][c];
            }
    }
    if (HES) {
#pragma unroll
        for (int p = 0; p < X; ++p) {
            double M[NS], tm[D][D];
#pragma unroll
            for (int e = 0; e < NS; ++e) M[e] = Mo[(i * X + p) * NS + e];
#pragma unroll
            for (int kk = 0; kk < D; ++kk)
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    double a = 0.0;
#pragma unroll
                    for (int l2 = 0; l2 < D; ++l2) a = fma(M[l2 <= kk ? kk * (kk + 1) / 2 + l2 : l2 * (l2 + 1) / 2 + kk], NEMPC_RF(l2, c), a);
                    tm[kk][c] = a;
                }
#pragma unroll
            for (int a2 = 0; a2 < D; ++a2)
#pragma unroll
                for (int c = 0; c <= a2; ++c) {
                    double a = 0.0;
#pragma unroll
                    for (int kk = 0; kk < D; ++kk) a = fma(NEMPC_RF(kk, a2), tm[kk][c], a);
                    const int e = a2 * (a2 + 1) / 2 + c;
                    // + a_s sum_k J[p][k] h_{s-1}[k]: h_{s-1} of EVERY output is needed, so the new h_s goes to the other half of a double buffer
                    if (!first) {
#pragma unroll
                        for (int kk = 0; kk < X; ++kk) a = fma(a_s * J[p][kk], sp[S::HPREV + ((stage & 1) ? 0 : X * NS) + kk * NS + e], a);
                    }
                    sp[S::HPREV + ((stage & 1) ? X * NS : 0) + p * NS + e] = a;
                    double* ha = sp + S::HACC + p * NS + e;
                    *ha = fma(c_s, a, first ? 0.0 : *ha);
                }
        }
    }
#undef NEMPC_RF
    double kacc[X];
#pragma unroll
    for (int p = 0; p < X; ++p) {
        kacc[p] = fma(c_s, k[p], first ? 0.0 : sp[S::KACC + p]);
        sp[S::KACC + p] = kacc[p];
    }
    if (!last) {
        const double an = st.a[stage + 1];
#pragma unroll
        for (int p = 0; p < X; ++p) {
            zin[i * D + p] = fma(an, k[p], z[p]);
#pragma unroll
            for (int c = 0; c < D; ++c) sp[S::RT + p * D + c] = JAC ? fma(an, dk[p][c], (p == c) ? 1.0 : 0.0) : 0.0;
        }
        return;
    }
    // ---- outputs (the scatter of nempc_fast.cuh, in double) -------------------------------------------------------------------------
    const bool unity = (ar.flags & NEMPC_UNITY) != 0;
    if (ar.resid) {
#pragma unroll
        for (int p = 0; p < X; ++p) {
            const double xt = zb[t * X + p];
            const double xp = unity ? 0.0 : z[p];
            ar.resid[b * L.m + t * X + p] = xp + kacc[p] - xt;
        }
    }
    if (JAC && ar.jac) {
        double* jv = ar.jac + b * L.nnz_jac;
#pragma unroll
        for (int p = 0; p < X; ++p) {
            jv[jac_slot_minus1(L, t, p)] = -1.0;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const double v = dkacc[p][c] + ((!unity && c == p) ? 1.0 : 0.0);
                if (c < X) { if (t > 0) jv[jac_slot_A(L, t, p, c)] = v; }
                else jv[jac_slot_B(L, t, p, c - X)] = v;
            }
        }
    }
    if (HES && ar.hes) {
        double* hv = ar.hes + b * L.nnz_hes;
        const double sig = ar.sigma ? ar.sigma[b] : ar.sigma_scalar;
        double lam[X];
#pragma unroll
        for (int p = 0; p < X; ++p) lam[p] = ar.lam[b * L.m + t * X + p];
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int c = 0; c <= a; ++c) {
                if (t == 0 && c < X) continue;
                double acc = 0.0;
#pragma unroll
                for (int p = 0; p < X; ++p) acc = fma(lam[p], sp[S::HACC + p * NS + a * (a + 1) / 2 + c], acc);
                double v = acc;
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
                hv[slot] = v;
            }
        if (t == L.H - 1) {
#pragma unroll
            for (int p = 0; p < X; ++p)
                if (L.hes_last_slot[p] >= 0) hv[L.hes_last_slot[p]] = sig * 2.0 * ar.quad[(L.H - 1) * X + p];
        }
    }
}
#endif
