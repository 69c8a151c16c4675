// Tensor-core kernel for wide tanh networks (hidden width 128: the cart-pole class of BASELINE config C3,
// 5 -> 128 -> 128 -> 128 -> 4): the per-step chain of nempc_generic.cuh with the hidden-to-hidden layers on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory).
//
// Formulation: FORWARD second order.  Per horizon step and RK4 stage the network is evaluated on a stack of
// RPS = 1 + d + d(d+1)/2 rows
//      P        the activations                 h_l
//      T_c      first-order tangents            V_l[., c]    = s'(a_l) * da_l/dz_c
//      S_(c,c2) second-order tangents           Q_l[., c,c2] = s''(a_l) * da_l/dz_c * da_l/dz_c2 + s'(a_l) * d2a_l/dz_c dz_c2
// Every row goes through the SAME weights, so a hidden-to-hidden layer is one GEMM [128 rows x 128] x [128 x 128] for
// SPT = 128 / RPS steps at once, and the linear output layer turns the three row kinds into k = f(z_s), the local
// Jacobian J_s and the per-output local Hessians M_p.  No adjoint sweep and no per-layer state survives a layer: that
// is what lets both hidden-to-hidden weight matrices stay resident in shared memory for the life of the CTA.
// (The adjoint form needs d*sum_h or x*sum_h floats per step across layers -- 15 KB per step for C3 -- which is why
// nempc_generic.cuh holds only ~15 steps per SM.)  The stage algebra (dk = J R, h_s = R^T M R + a_s J h_{s-1},
// R_{s+1} = I + a_{s+1} E dk) and the sparse scatter are those of nempc_generic.cuh (reference integrator/rk4.py:113-285).
//
// Arithmetic: the reference network is float32 (model/tensorflow.py:85-91); plain TF32/F16 tensor-core inputs are
// ~1e-3 accurate, far from the 1e-5 parity bound.  Operands are therefore split in two f16 terms, x = hi + lo/2^11
// (tcx::split_f16), and a layer is three MMAs per K step into ONE f32 accumulator in tensor memory:
//      D  = A_lo W_hi + A_hi W_lo            (both products carry the factor 2^11)
//      D  = A_hi W_hi + D * 2^-11            (first main MMA, tcgen05.mma scale-input-d = 11), then D += A_hi W_hi for the other K steps
// (a single accumulator matters: tensor memory is read at only ~52 B/clk/SM, tests/tools/tc_tmem_probe.cu),
// which carries ~22 mantissa bits (measured 4e-7 relative, tests/tools/tc_gemm_probe.cu) at 1.5x the cost of one TF32
// pass and HALF the shared-memory footprint of a 3xTF32 scheme -- the footprint is what decides residency here.
//
// (Hidden widths 64 / 32 are the same kernel with N = 64 / 32 MMAs; their tiles need <= 128 tensor-memory columns, so FOUR groups of 128
// threads run.)
// One persistent CTA per SM, NEMPC_TC_THREADS = 512 threads = TWO GROUPS of 256 threads, each working on its own row tile with its
// own named barrier, mbarrier and 256 tensor-memory columns: while one group waits for its MMA batch the other runs its epilogue
// (the overlap a second CTA per SM would give, without a second copy of the weights).  Thread (m = gtid & 127, cq = gtid >> 7) of a
// group owns row m of the tile and the neurons [CPT cq, CPT cq + CPT).  Row order is kind-major (m = kind * SPT + step).
// The operand tile lives in TENSOR MEMORY (tcgen05.st, two f16 K elements per 32-bit column, lane = row; TS-mode MMA), which is
// what makes room for the second tile.  Per layer:
//   the group's first thread issues 8 K-steps x 3 tcgen05.mma (M = 128, N = 128, K = 16) and commits to the group's mbarrier;
//   pass 1: P rows add the bias and park a_l in a side buffer, T rows park their raw tangents (needed by the S rows);
//   tanh  : the SPT x 128 activations are spread over all threads of the group (MUFU tanh, as in nempc_fast.cuh);
//   pass 2: every row forms its post-activation quantity from tensor memory + the side buffers, splits it and stores its part of the
//           next operand tile; after the LAST hidden layer the rows are contracted with W_out in registers instead (exact f32).
// The first layer (K = d) and the output layer (N = x) are too thin for the tensor core and stay on FFMA.
#pragma once
#include "nempc_fast.cuh"
#include "nempc_generic.cuh"
#include "nempc_tc_ptx.cuh"

#include <type_traits>

#define NEMPC_TC_HW 128
#ifndef NEMPC_TC_THREADS
#define NEMPC_TC_THREADS 512            // two tile groups of 256 threads (width 128) or four of 128 (width 64)
#endif
#ifndef NEMPC_TC_NG
#define NEMPC_TC_NG(hw) ((hw) <= 64 ? 4 : 2)      // tile groups per CTA
#endif
#define NEMPC_TC_SMEM_MAX 232448
#ifndef NEMPC_TC_P2_UNROLL
#define NEMPC_TC_P2_UNROLL 2          // unroll factor of the 16-neuron chunk loop of pass 2 (2: both tensor-memory loads in flight; +3% on C3)
#endif

template <int X_, int U_, int NHID_, int MODE_, int HW_ = NEMPC_TC_HW> struct TcCfg {
    static constexpr int X = X_, U = U_, NHID = NHID_, MODE = MODE_;
    static constexpr int D = X + U, NTRI = D * (D + 1) / 2, HW = HW_, NMM = NHID - 1;
    static constexpr bool JAC = MODE >= 1, HES = MODE >= 2;
    static constexpr int RPS = 1 + (JAC ? D : 0) + (HES ? NTRI : 0);      // rows per step
    static constexpr int SPT = (128 / RPS) < 32 ? (128 / RPS) : 32;       // steps per tile
    static constexpr int ROWS = SPT * RPS;
    static constexpr int XP = (X + 3) / 4 * 4;
    static constexpr int SROW = HW + 4;                                   // padded side-buffer row (bank spread)
    // per-step state that lives across the RK4 stages (floats)
    static constexpr int P_Z = 0, P_KPREV = P_Z + D, P_KACC = P_KPREV + X, P_R = P_KACC + X, P_DKACC = P_R + (JAC ? D * D : 0),
                         P_HPREV = P_DKACC + (JAC ? X * D : 0), P_HACC = P_HPREV + (HES ? X * D * D : 0),
                         P_TOTAL = P_HACC + (HES ? X * D * D : 0);
    // per-step temporaries of one stage (floats)
    static constexpr int T_KCUR = 0, T_J = XP, T_DK = T_J + (JAC ? X * D : 0), T_M = T_DK + (JAC ? X * D : 0),
                         T_TMP = T_M + (HES ? X * D * D : 0), T_TOTAL = T_TMP + (HES ? X * D * D : 0);
    // two independent row tiles per CTA ("groups" of GT threads): while one group waits for its MMA batch the other runs its
    // epilogue -- the overlap a second CTA per SM would give, without a second copy of the weights
    // (hidden width 64 needs half the tensor-memory columns per tile, so FOUR groups of 128 threads fit)
    static constexpr int NG = NEMPC_TC_NG(HW), GT = NEMPC_TC_THREADS / NG;
    static constexpr int NQ = GT / 128;                                   // threads per row: each owns CPT consecutive neurons
    static constexpr int CPT = HW / NQ;
    // tensor memory, per group: D (f32 accumulator) | A_hi | A_lo (f16 operand tile, two K elements per 32-bit column)
    static constexpr int TM_GROUP = 512 / NG, TM_D = 0, TM_AHI = HW, TM_ALO = HW + HW / 2, TM_COLS = NG * TM_GROUP;
    // shared-memory map (bytes)
    static constexpr int IMG = HW * HW * 2;                               // one f16 weight image [n = HW][k = HW]: 32 KB at HW = 128
    static constexpr int OFF_W = 0;                                       // NMM x (hi image, lo image)
    static constexpr int OFF_C = NMM * 2 * IMG;                           // f32 constants
    static constexpr int C_W0 = 0, C_WOUT = D * HW, C_B = C_WOUT + HW * XP, C_BOUT = C_B + NHID * HW, C_FLOATS = C_BOUT + XP;
    static constexpr int OFF_G = OFF_C + C_FLOATS * 4;                    // group blocks
    static constexpr int G_SH = 0;                                        // a_l / h_l     [SPT][SROW]
    static constexpr int G_ST = G_SH + SPT * SROW * 4;                    // raw tangents  [D][SPT][SROW]   (Hessian only)
    static constexpr int G_S1 = G_ST + (HES ? D * SPT * SROW * 4 : 0);    // s'(a_l)  and  s''(a_l)  [SPT][SROW] each (Hessian only): computed
    static constexpr int G_S2 = G_S1 + (HES ? SPT * SROW * 4 : 0);        // once per (step, neuron) by the tanh pass, not once per row
    static constexpr int G_P = G_S2 + (HES ? SPT * SROW * 4 : 0);         // per-step state across the stages
    static constexpr int G_I = G_P + SPT * P_TOTAL * 4;                   // (problem, time index) of each step of the tile
    static constexpr int G_PART = G_I + SPT * 8;                          // [NQ-1][128][XP] partial output sums of the upper column groups
    static constexpr int G_TMP = G_PART + (NQ - 1) * 128 * XP * 4;        // stage temporaries
    static constexpr int G_BYTES = (G_TMP + SPT * T_TOTAL * 4 + 15) / 16 * 16;
    static constexpr int TOTAL = OFF_G + NG * G_BYTES;
    static_assert(NEMPC_TC_THREADS % (128 * NG) == 0 && CPT % 16 == 0, "128 rows x NQ column groups of a multiple of 16 neurons per group");
    static_assert(TOTAL <= NEMPC_TC_SMEM_MAX, "tensor-core kernel: shared-memory map exceeds 227 KB");
    static_assert(TM_COLS <= 512 && TM_ALO + HW / 2 <= TM_GROUP, "tensor memory: 512 columns");
    static_assert(NHID >= 2 && NHID <= 3, "two or three hidden layers");
    static_assert(HW == 128 || HW == 64 || HW == 32, "hidden width: 128 (N = 128 MMAs), 64 (N = 64) or 32 (N = 32)");
    static_assert(X <= 16 && RPS <= 128, "row stack too tall");
};

// host: element (n, k) of a K-major no-swizzle operand image with hw rows: 16-byte chunk (n, k/8) at (k/8)*(hw*16) + n*16
inline size_t tc_img_index(int n, int k, int hw = NEMPC_TC_HW) { return (size_t)(k / 8) * ((size_t)hw * 8) + (size_t)n * 8 + (k % 8); }

#if defined(__CUDACC__)
// -DNEMPC_TC_PROFILE: thread 0 of every CTA accumulates the cycles between phase boundaries (development builds only)
#ifdef NEMPC_TC_PROFILE
static __device__ unsigned long long nempc_tc_prof[16];
#define TC_PROF_DECL long long prof_t0 = clock64(); unsigned long long prof_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define TC_PROF(i) do { const long long t_ = clock64(); prof_acc[i] += (unsigned long long)(t_ - prof_t0); prof_t0 = t_; } while (0)
#define TC_PROF_FLUSH do { if (tid == 0) for (int i_ = 0; i_ < 16; ++i_) atomicAdd(&nempc_tc_prof[i_], prof_acc[i_]); } while (0)
#else
#define TC_PROF_DECL
#define TC_PROF(i) do {} while (0)
#define TC_PROF_FLUSH do {} while (0)
#endif

template <class C, typename TIO>
__global__ void __launch_bounds__(NEMPC_TC_THREADS, 1)
nempc_tc_kernel(const __half* __restrict__ wimg, const float* __restrict__ cblk, const StageTable<float> st,
                const NlpLayout L, const EvalArgs<TIO> ar, const float* __restrict__ wexo, const int tvp_dim, const int p_dim) {
    using namespace tcx;
    constexpr int X = C::X, U = C::U, D = C::D, HW = C::HW, NHID = C::NHID, NMM = C::NMM, SPT = C::SPT, XP = C::XP,
                  SROW = C::SROW, DD = D * D;
    constexpr bool JAC = C::JAC, HES = C::HES;
    typedef typename WideOf<float, TIO>::type TW;
    extern __shared__ __align__(128) unsigned char tc_smem[];
    __shared__ uint64_t mbar_store[1 + C::NG];
    __shared__ uint32_t tmem_holder;

    constexpr int GT = C::GT;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int grp = tid / GT, gtid = tid - grp * GT;   // tile group of this thread, index inside the group
    const int m = gtid & 127, cq = gtid >> 7;          // row, column group
    const bool valid = m < C::ROWS;
    const int kind = valid ? m / SPT : 0, sl = valid ? m - (m / SPT) * SPT : 0;     // spare rows of the tile act as a second copy of P row 0 (finite, unused)
    int cT = 0, c1 = 0, c2 = 0;                       // tangent column of a T row; column pair of an S row
    if (valid && kind >= 1) {
        if (kind <= D) cT = kind - 1;
        else {
            int e = kind - 1 - D;
            while (e > c1) { e -= c1 + 1; ++c1; }     // e-th entry of the lower triangle -> (c1, c2)
            c2 = e;
        }
    }
    // pass 2 reads tensor memory only for rows that did not park their values in pass 1: S rows (Hessian mode) or T rows
    // (Jacobian mode); the load is warp-collective, so the test is per lane quadrant
    const int first_tmem_row = HES ? SPT * (1 + D) : SPT;
    const bool quad_reads_tmem = JAC && ((warp & 3) * 32 + 31 >= first_tmem_row) && ((warp & 3) * 32 < C::ROWS);
    const bool is_T = valid && kind >= 1 && kind <= D;
    const bool is_S = valid && kind > D;

    float* cb = reinterpret_cast<float*>(tc_smem + C::OFF_C);
    const float* W0 = cb + C::C_W0;                   // [D][HW]
    const float* Wout = cb + C::C_WOUT;               // [HW][XP]
    const float* bias = cb + C::C_B;                  // [NHID][HW]
    const float* bout = cb + C::C_BOUT;
    unsigned char* gblk = tc_smem + C::OFF_G + grp * C::G_BYTES;     // this group's block
    float* sideH = reinterpret_cast<float*>(gblk + C::G_SH);
    float* sideT = reinterpret_cast<float*>(gblk + C::G_ST);
    float* sideS1 = reinterpret_cast<float*>(gblk + C::G_S1);
    float* sideS2 = reinterpret_cast<float*>(gblk + C::G_S2);
    float* pers = reinterpret_cast<float*>(gblk + C::G_P);
    int* step_b = reinterpret_cast<int*>(gblk + C::G_I);             // problem index of step s_ of the tile, -1 past the end
    int* step_t = step_b + SPT;
    float* part = reinterpret_cast<float*>(gblk + C::G_PART);
    float* temps = reinterpret_cast<float*>(gblk + C::G_TMP);
    // the threads of one group meet on their own named barrier; the two groups never wait for each other inside the tile loop
    auto gsync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(GT) : "memory"); };

    const uint32_t mbar_w = smem_u32(&mbar_store[0]), mbar_mma = smem_u32(&mbar_store[1 + grp]);
    if (tid == 0) { mbar_init(mbar_w, 1); for (int g = 0; g < C::NG; ++g) mbar_init(smem_u32(&mbar_store[1 + g]), 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_holder), C::TM_COLS);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_holder;
    const uint32_t tm_row = tmem + grp * C::TM_GROUP + ((uint32_t)((warp & 3) * 32) << 16);     // this group's columns, this warp's lane quadrant
    // weights: resident for the life of the CTA, brought in by the TMA unit as plain 1-D bulk copies
    if (tid == 0) {
        mbar_expect_tx(mbar_w, NMM * 2 * C::IMG);
        for (int i = 0; i < NMM * 2; ++i)
            bulk_g2s(smem_u32(tc_smem + C::OFF_W + i * C::IMG), wimg + (size_t)i * (C::IMG / 2), C::IMG, mbar_w);
    }
    for (int i = tid; i < C::C_FLOATS; i += NEMPC_TC_THREADS) cb[i] = cblk[i];
    mbar_wait(mbar_w, 0);
    __syncthreads();

    const bool unity = (ar.flags & NEMPC_UNITY) != 0;
    const uint32_t idesc = make_idesc_f16(128, HW);
    uint32_t parity = 0;
    const long long ntiles = (ar.nsteps + SPT - 1) / SPT;

    TC_PROF_DECL
    for (long long tile = (long long)blockIdx.x * C::NG + grp; tile < ntiles; tile += (long long)gridDim.x * C::NG) {
        const long long step0 = tile * SPT;
        TC_PROF(11);
        // ---- inputs and per-step state ------------------------------------------------------------------------------
        if (gtid < SPT) {
            const long long step = step0 + gtid;
            const long long b = step / L.H;
            step_b[gtid] = step < ar.nsteps ? (int)b : -1;
            step_t[gtid] = (int)(step - b * L.H);
        }
        gsync();
        for (int idx = gtid; idx < SPT * C::P_TOTAL; idx += GT) {
            const int s_ = idx / C::P_TOTAL, o = idx - s_ * C::P_TOTAL;
            float v = 0.f;
            if (o < D) {
                const long long b = step_b[s_];
                if (b >= 0) {
                    const int t = step_t[s_];
                    const TIO* zb = ar.z + b * (long long)L.n;
                    if (o < X) v = (float)((t == 0) ? ar.x0[b * X + o] : zb[(t - 1) * X + o]);
                    else v = (float)zb[L.H * X + t * U + (o - X)];
                }
            } else if (JAC && o >= C::P_R && o < C::P_R + DD) {
                const int r = o - C::P_R;
                v = (r / D == r % D) ? 1.f : 0.f;
            }
            pers[idx] = v;
        }
        gsync();
        TC_PROF(0);

        for (int s = 0; s < st.S; ++s) {
            const float a_s = st.a[s], c_s = st.c[s];
            // ---- first layer (K = d, FFMA): activations of all SPT steps spread over the CTA ------------------------------
            for (int e = gtid; e < SPT * HW; e += GT) {
                const int s_ = e / HW, j = e - s_ * HW;
                const float* ps = pers + s_ * C::P_TOTAL;
                float a = bias[j];
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const float zc = (c < X) ? fmaf(a_s, ps[C::P_KPREV + (c < X ? c : 0)], ps[C::P_Z + c]) : ps[C::P_Z + c];
                    a = fmaf(W0[c * HW + j], zc, a);
                }
                if (tvp_dim + p_dim > 0) {        // exogenous inputs [tvp_t, p] (model/tensorflow.py:39-47): a shift of this pre-activation only
                    const long long b = step_b[s_];
                    if (b >= 0) {
                        const double* tv = ar.tvp + b * ar.tvp_bstride + (long long)step_t[s_] * tvp_dim;
                        for (int q = 0; q < tvp_dim; ++q) a = fmaf(__ldg(wexo + q * HW + j), (float)__ldg(tv + q), a);
                        const double* pv = ar.p + b * ar.p_bstride;
                        for (int q = 0; q < p_dim; ++q) a = fmaf(__ldg(wexo + (tvp_dim + q) * HW + j), (float)__ldg(pv + q), a);
                    }
                }
                const float h = fast_tanh(a);
                sideH[s_ * SROW + j] = h;
                if (HES) { const float s1 = fmaf(-h, h, 1.f); sideS1[s_ * SROW + j] = s1; sideS2[s_ * SROW + j] = -2.f * h * s1; }
            }
            if (HES) {      // da_0/dz_c = W0[c]: park it like the raw tangents of any other layer, so pass 2 has ONE code path
                for (int e = gtid; e < D * SPT * (HW / 4); e += GT) {
                    const int c = e / (SPT * (HW / 4)), r = e - c * (SPT * (HW / 4)), s_ = r / (HW / 4), j4 = r - s_ * (HW / 4);
                    *reinterpret_cast<float4*>(sideT + (c * SPT + s_) * SROW + 4 * j4) = *reinterpret_cast<const float4*>(W0 + c * HW + 4 * j4);
                }
            }
            gsync();
            TC_PROF(1);

            constexpr int XH = (X + 1) / 2;
            f2 oacc2[XH];                                          // output-layer partial sums, packed over output pairs
#pragma unroll
            for (int q = 0; q < XH; ++q) oacc2[q] = pk(0.f, 0.f);

            // ---- pass 2 of a layer: post-activation rows -> next operand tile (or, LAST, the output contraction) ---------------
            // Packed f32x2 arithmetic (FFMA2); the row class (P / T / S) is a per-thread branch that only diverges in the two
            // warps that straddle a class boundary (row order is kind-major).  One instance serves every layer but the last
            // (the kernel is kept small on purpose: the lone MMA-issuing thread pays every instruction-cache miss in full).
            auto pass2 = [&](auto last_tag, const bool first) {
                constexpr bool LAST = decltype(last_tag)::value;
                constexpr int P2U = NEMPC_TC_P2_UNROLL;
#pragma unroll P2U
                for (int q4 = 0; q4 < C::CPT / 16; ++q4) {
                    const int col = C::CPT * cq + 16 * q4;
                    f2 v2[8];
                    if (!first && quad_reads_tmem) {               // warp-collective tensor-memory load: every lane of the warp takes part
                        float v[16];
                        tmem_ld16(tm_row + col, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) v2[i] = pk(v[2 * i], v[2 * i + 1]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v2[i] = pk(0.f, 0.f);    // first layer: d2a_0 = 0
                    }
                    f2 out2[8];
                    {
                        const float* hrow = sideH + sl * SROW + col;
                        if (kind == 0) {                           // P: the activations themselves
#pragma unroll
                            for (int i4 = 0; i4 < 4; ++i4) {
                                const float4 h4 = *reinterpret_cast<const float4*>(hrow + 4 * i4);
                                out2[2 * i4] = pk(h4.x, h4.y); out2[2 * i4 + 1] = pk(h4.z, h4.w);
                            }
                        } else if (!is_S) {                        // T_c: s'(a) * da/dz_c
                            // raw tangent: parked in sideT (Hessian mode); else W0[c] for the first layer, tensor memory otherwise
                            const float* traw = HES ? sideT + (cT * SPT + sl) * SROW + col : W0 + cT * HW + col;
                            const float* s1row = sideS1 + sl * SROW + col;
                            const f2 m1 = pk(-1.f, -1.f), one = pk(1.f, 1.f);
#pragma unroll
                            for (int i4 = 0; i4 < 4; ++i4) {
                                f2 va = v2[2 * i4], vb = v2[2 * i4 + 1];
                                if (HES || first) {
                                    const float4 w4 = *reinterpret_cast<const float4*>(traw + 4 * i4);
                                    va = pk(w4.x, w4.y); vb = pk(w4.z, w4.w);
                                }
                                if (HES) {
                                    const float4 s4 = *reinterpret_cast<const float4*>(s1row + 4 * i4);
                                    out2[2 * i4] = mul2(pk(s4.x, s4.y), va);
                                    out2[2 * i4 + 1] = mul2(pk(s4.z, s4.w), vb);
                                } else {
                                    const float4 h4 = *reinterpret_cast<const float4*>(hrow + 4 * i4);
                                    const f2 ha = pk(h4.x, h4.y), hb = pk(h4.z, h4.w);
                                    out2[2 * i4] = mul2(fma2(mul2(ha, m1), ha, one), va);
                                    out2[2 * i4 + 1] = mul2(fma2(mul2(hb, m1), hb, one), vb);
                                }
                            }
                        } else {                                   // S_(c1,c2): s''(a) T_c1 T_c2 + s'(a) d2a
                            const float* t1 = sideT + (c1 * SPT + sl) * SROW + col;
                            const float* t2 = sideT + (c2 * SPT + sl) * SROW + col;
                            const float* s1row = sideS1 + sl * SROW + col;
                            const float* s2row = sideS2 + sl * SROW + col;
#pragma unroll
                            for (int i4 = 0; i4 < 4; ++i4) {
                                const float4 s4 = *reinterpret_cast<const float4*>(s1row + 4 * i4);
                                const float4 r4 = *reinterpret_cast<const float4*>(s2row + 4 * i4);
                                const float4 a4 = *reinterpret_cast<const float4*>(t1 + 4 * i4);
                                const float4 b4 = *reinterpret_cast<const float4*>(t2 + 4 * i4);
                                out2[2 * i4] = fma2(mul2(pk(a4.x, a4.y), pk(b4.x, b4.y)), pk(r4.x, r4.y), mul2(pk(s4.x, s4.y), v2[2 * i4]));
                                out2[2 * i4 + 1] = fma2(mul2(pk(a4.z, a4.w), pk(b4.z, b4.w)), pk(r4.z, r4.w), mul2(pk(s4.z, s4.w), v2[2 * i4 + 1]));
                            }
                        }
                    }
                    if (!LAST) {
                        // x = hi + lo / 2^11, two neurons per instruction; the operand tile lives in TENSOR MEMORY (two f16 K elements
                        // per 32-bit column, lane = row): 8 words = 16 neurons per tcgen05.st, hi and lo image
                        const f2 sc = pk(NEMPC_TC_LO_SCALE, NEMPC_TC_LO_SCALE), nsc = pk(-NEMPC_TC_LO_SCALE, -NEMPC_TC_LO_SCALE);
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const f2 x = out2[i];
                            const __half2 h2 = __floats2half2_rn(f2lo(x), f2hi(x));
                            const float2 hf2 = __half22float2(h2);
                            const f2 r = fma2(x, sc, mul2(pk(hf2.x, hf2.y), nsc));            // (x - hi) * 2^11, exact
                            const __half2 l2 = __floats2half2_rn(f2lo(r), f2hi(r));
                            hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
                            lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
                        }
                        tmem_st8(tm_row + C::TM_AHI + col / 2, hi);           // warp-collective: invalid rows store zeros
                        tmem_st8(tm_row + C::TM_ALO + col / 2, lo);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float o = (i & 1) ? f2hi(out2[i / 2]) : f2lo(out2[i / 2]);
                            const float* wo = Wout + (col + i) * XP;
#pragma unroll
                            for (int q = 0; q < XH; ++q) {
                                const float2 w2 = *reinterpret_cast<const float2*>(wo + 2 * q);
                                oacc2[q] = fma2(pk(o, o), pk(w2.x, w2.y), oacc2[q]);
                            }
                        }
                    }
                }
            };

#pragma unroll 1
            for (int l = 0; l < NHID - 1; ++l) {
                pass2(std::false_type{}, l == 0);
                TC_PROF(2);

                // ---- hidden-to-hidden layer l -> l+1 on the tensor core -------------------------------------------------------
                tmem_st_wait();                                    // this thread's operand-tile stores have landed in tensor memory
                fence_before_sync();
                gsync();
                TC_PROF(3);
                if (gtid < 32) {                                   // first warp of the group
                  if (gtid == 0) {
                    fence_after_sync();
                    // Layer and group offsets must be COMPILE-TIME constants: with run-time indices ptxas cannot prove the operands
                    // warp-uniform inside this single-thread branch and wraps every MMA in an ELECT / R2UR loop (~110 clk per MMA
                    // instead of the tensor core's 64).  Hence one unrolled instance per (layer, group).
                    auto issue = [&](auto layer_tag, auto group_tag) {
                        constexpr int LI = decltype(layer_tag)::value, GI = decltype(group_tag)::value;
                        const uint32_t whi = smem_u32(tc_smem) + C::OFF_W + LI * 2 * C::IMG, wlo = whi + C::IMG;
                        const uint32_t td = tmem + GI * C::TM_GROUP + C::TM_D, ta1 = tmem + GI * C::TM_GROUP + C::TM_AHI, ta2 = tmem + GI * C::TM_GROUP + C::TM_ALO;
                        // correction terms first (both carry the factor 2^11), then the first main MMA folds them in with
                        // scale-input-d: D = A_hi W_hi + D * 2^-11 -- ONE f32 accumulator, half the tensor-memory read volume.
                        // A operand from tensor memory (8 columns = 16 K elements per MMA), B descriptors advanced in their low word.
                        const uint32_t dhi = (128u >> 4) | (1u << 14);                       // SBO = 128 B, descriptor version 1
                        const uint32_t dlo = ((uint32_t)(HW * 16) >> 4) << 16;               // LBO = HW * 16 B: next 8-wide K chunk of the image
                        const uint32_t lw1 = dlo | (whi >> 4), lw2 = dlo | (wlo >> 4);
#define NEMPC_TC_DESC(lo, ks) ((((uint64_t)dhi) << 32) | (uint64_t)((lo) + (ks) * (2u * HW)))
#pragma unroll
                        for (int ks = 0; ks < HW / 16; ++ks) mma_f16_ts(td, ta2 + ks * 8, NEMPC_TC_DESC(lw1, ks), idesc, ks != 0);
#pragma unroll
                        for (int ks = 0; ks < HW / 16; ++ks) mma_f16_ts(td, ta1 + ks * 8, NEMPC_TC_DESC(lw2, ks), idesc, 1);
                        mma_f16_ts_scaled_d<11>(td, ta1, NEMPC_TC_DESC(lw1, 0), idesc);
#pragma unroll
                        for (int ks = 1; ks < HW / 16; ++ks) mma_f16_ts(td, ta1 + ks * 8, NEMPC_TC_DESC(lw1, ks), idesc, 1);
#undef NEMPC_TC_DESC
                    };
                    typedef std::integral_constant<int, 0> I0;
                    typedef std::integral_constant<int, (NMM > 1 ? 1 : 0)> I1;
                    typedef std::integral_constant<int, 1> G1;
                    typedef std::integral_constant<int, (C::NG > 2 ? 2 : 0)> G2;
                    typedef std::integral_constant<int, (C::NG > 2 ? 3 : 0)> G3;
                    if (grp == 0)      { if (l == 0) issue(I0{}, I0{}); else issue(I1{}, I0{}); }
                    else if (grp == 1) { if (l == 0) issue(I0{}, G1{}); else issue(I1{}, G1{}); }
                    else if (grp == 2) { if (l == 0) issue(I0{}, G2{}); else issue(I1{}, G2{}); }
                    else               { if (l == 0) issue(I0{}, G3{}); else issue(I1{}, G3{}); }
                    mma_commit(mbar_mma);
                    TC_PROF(4);
                    mbar_wait(mbar_mma, parity);                   // ONE polling thread; the rest of the group blocks on its barrier below
                  }
                  __syncwarp();
                }
                parity ^= 1;
                gsync();                                   // everybody else blocks on the hardware barrier
                fence_after_sync();
                TC_PROF(5);

                // ---- pass 1 of layer l+1: P rows park a_{l+1} (+ bias), T rows park their raw tangents ---------------------------
                {
                    const int need_rows = SPT * (1 + (HES ? D : 0));      // kind-major: P rows first, then T rows
                    if ((warp & 3) * 32 < need_rows) {
                        const float* bl = bias + (l + 1) * HW;
#pragma unroll 1
                        for (int q4 = 0; q4 < C::CPT / 16; ++q4) {
                            const int col = C::CPT * cq + 16 * q4;
                            float v[16];
                            tmem_ld16(tm_row + col, v);
                            tmem_ld_wait();
                            if (valid && kind == 0) {
#pragma unroll
                                for (int i4 = 0; i4 < 4; ++i4) {
                                    const float4 b4 = *reinterpret_cast<const float4*>(bl + col + 4 * i4);
                                    *reinterpret_cast<float4*>(sideH + sl * SROW + col + 4 * i4) =
                                        make_float4(v[4 * i4] + b4.x, v[4 * i4 + 1] + b4.y, v[4 * i4 + 2] + b4.z, v[4 * i4 + 3] + b4.w);
                                }
                            } else if (HES && is_T) {
#pragma unroll
                                for (int i4 = 0; i4 < 4; ++i4)
                                    *reinterpret_cast<float4*>(sideT + (cT * SPT + sl) * SROW + col + 4 * i4) =
                                        make_float4(v[4 * i4], v[4 * i4 + 1], v[4 * i4 + 2], v[4 * i4 + 3]);
                            }
                        }
                    }
                }
                gsync();
                TC_PROF(6);
                for (int e = gtid; e < SPT * HW; e += GT) {
                    const int s_ = e / HW, j = e - s_ * HW;
                    const float h = fast_tanh(sideH[s_ * SROW + j]);
                    sideH[s_ * SROW + j] = h;
                    if (HES) { const float s1 = fmaf(-h, h, 1.f); sideS1[s_ * SROW + j] = s1; sideS2[s_ * SROW + j] = -2.f * h * s1; }
                }
                gsync();
                TC_PROF(7);
            }
            pass2(std::true_type{}, false);                        // last hidden layer: rows contracted with W_out in registers
            TC_PROF(2);

            // ---- linear output layer: join the two neuron halves; k, J, M_p of this stage -------------------------------------
            float oacc[X];
#pragma unroll
            for (int p = 0; p < X; ++p) oacc[p] = (p & 1) ? f2hi(oacc2[p / 2]) : f2lo(oacc2[p / 2]);
            gsync();                                       // every thread is done with tensor memory and the side buffers
            if (cq > 0 && valid) {
#pragma unroll
                for (int p = 0; p < X; ++p) part[((cq - 1) * 128 + m) * XP + p] = oacc[p];
            }
            gsync();
            if (cq == 0 && valid) {
                float* tp = temps + sl * C::T_TOTAL;
#pragma unroll
                for (int p = 0; p < X; ++p) {
                    float tot = oacc[p];
#pragma unroll
                    for (int g = 0; g + 1 < C::NQ; ++g) tot += part[(g * 128 + m) * XP + p];
                    if (kind == 0) tp[C::T_KCUR + p] = tot + bout[p];
                    else if (kind <= D) tp[C::T_J + p * D + cT] = tot;
                    else { tp[C::T_M + p * DD + c1 * D + c2] = tot; tp[C::T_M + p * DD + c2 * D + c1] = tot; }
                }
            }
            gsync();

            TC_PROF(8);
            // ---- stage algebra (as nempc_generic.cuh), flattened over the SPT steps of the tile ---------------------------------
            if (JAC) {
                for (int idx = gtid; idx < SPT * X * D; idx += GT) {
                    const int s_ = idx / (X * D), r = idx - s_ * (X * D), p = r / D, c = r - p * D;
                    const float* tp = temps + s_ * C::T_TOTAL;
                    const float* R = pers + s_ * C::P_TOTAL + C::P_R;
                    float acc = 0.f;
#pragma unroll
                    for (int k = 0; k < D; ++k) acc = fmaf(tp[C::T_J + p * D + k], R[k * D + c], acc);
                    temps[s_ * C::T_TOTAL + C::T_DK + r] = acc;
                }
                if (HES) {
                    for (int idx = gtid; idx < SPT * X * DD; idx += GT) {
                        const int s_ = idx / (X * DD), r = idx - s_ * (X * DD), p = r / DD, k = (r - p * DD) / D, c = r % D;
                        const float* tp = temps + s_ * C::T_TOTAL;
                        const float* R = pers + s_ * C::P_TOTAL + C::P_R;
                        float acc = 0.f;
#pragma unroll
                        for (int l2 = 0; l2 < D; ++l2) acc = fmaf(tp[C::T_M + p * DD + k * D + l2], R[l2 * D + c], acc);
                        temps[s_ * C::T_TOTAL + C::T_TMP + r] = acc;
                    }
                }
                gsync();
                if (HES) {
                    for (int idx = gtid; idx < SPT * X * DD; idx += GT) {
                        const int s_ = idx / (X * DD), r = idx - s_ * (X * DD), p = r / DD, a = (r - p * DD) / D, c = r % D;
                        float* tp = temps + s_ * C::T_TOTAL;
                        const float* ps = pers + s_ * C::P_TOTAL;
                        float acc = 0.f;
#pragma unroll
                        for (int k = 0; k < D; ++k) acc = fmaf(ps[C::P_R + k * D + a], tp[C::T_TMP + p * DD + k * D + c], acc);
                        if (s > 0) {
#pragma unroll
                            for (int k = 0; k < X; ++k) acc = fmaf(a_s * tp[C::T_J + p * D + k], ps[C::P_HPREV + k * DD + a * D + c], acc);
                        }
                        tp[C::T_M + r] = acc;                      // M_p was consumed before the barrier; it now holds h_s[p]
                    }
                }
                for (int idx = gtid; idx < SPT * X * D; idx += GT) {
                    const int s_ = idx / (X * D), r = idx - s_ * (X * D);
                    float* dkacc = pers + s_ * C::P_TOTAL + C::P_DKACC + r;
                    *dkacc = fmaf(c_s, temps[s_ * C::T_TOTAL + C::T_DK + r], *dkacc);
                }
            }
            gsync();
            for (int idx = gtid; idx < SPT * X; idx += GT) {
                const int s_ = idx / X, p = idx - s_ * X;
                float* ps = pers + s_ * C::P_TOTAL;
                const float kc = temps[s_ * C::T_TOTAL + C::T_KCUR + p];
                ps[C::P_KACC + p] = fmaf(c_s, kc, ps[C::P_KACC + p]);
                ps[C::P_KPREV + p] = kc;
            }
            if (HES) {
                for (int idx = gtid; idx < SPT * X * DD; idx += GT) {
                    const int s_ = idx / (X * DD), r = idx - s_ * (X * DD);
                    float* ps = pers + s_ * C::P_TOTAL;
                    const float hv = temps[s_ * C::T_TOTAL + C::T_M + r];
                    ps[C::P_HACC + r] = fmaf(c_s, hv, ps[C::P_HACC + r]);
                    ps[C::P_HPREV + r] = hv;
                }
            }
            if (JAC && s + 1 < st.S) {
                const float an = st.a[s + 1];
                for (int idx = gtid; idx < SPT * DD; idx += GT) {
                    const int s_ = idx / DD, r = idx - s_ * DD, k = r / D, c = r - k * D;
                    pers[s_ * C::P_TOTAL + C::P_R + r] = (k == c ? 1.f : 0.f) + (k < X ? an * temps[s_ * C::T_TOTAL + C::T_DK + k * D + c] : 0.f);
                }
            }
            gsync();
        }

        TC_PROF(9);
        // ---- outputs (same slots as nempc_generic.cuh) ----------------------------------------------------------------------
        if (ar.resid) {
            for (int idx = gtid; idx < SPT * X; idx += GT) {
                const int s_ = idx / X, p = idx - s_ * X;
                const long long b = step_b[s_];
                if (b < 0) continue;
                const int t = step_t[s_];
                const TIO* zb = ar.z + b * (long long)L.n;
                const TW xt = (TW)zb[t * X + p];
                const TW xp = unity ? (TW)0 : (TW)((t == 0) ? ar.x0[b * X + p] : zb[(t - 1) * X + p]);
                ar.resid[b * L.m + t * X + p] = (TIO)(xp + (TW)pers[s_ * C::P_TOTAL + C::P_KACC + p] - xt);
            }
        }
        if (JAC && ar.jac) {
            for (int idx = gtid; idx < SPT * X * (D + 1); idx += GT) {
                const int s_ = idx / (X * (D + 1)), r = idx - s_ * (X * (D + 1)), p = r / (D + 1), c = r - p * (D + 1);
                const long long b = step_b[s_];
                if (b < 0) continue;
                const int t = step_t[s_];
                TIO* jv = ar.jac + b * L.nnz_jac;
                if (c == D) { jv[jac_slot_minus1(L, t, p)] = (TIO)-1; continue; }
                const TW v = (TW)pers[s_ * C::P_TOTAL + C::P_DKACC + p * D + c] + ((!unity && c == p) ? (TW)1 : (TW)0);
                if (c < X) { if (t > 0) jv[jac_slot_A(L, t, p, c)] = (TIO)v; }
                else jv[jac_slot_B(L, t, p, c - X)] = (TIO)v;
            }
        }
        if (HES && ar.hes) {
            for (int idx = gtid; idx < SPT * DD; idx += GT) {
                const int s_ = idx / DD, r = idx - s_ * DD, a = r / D, c = r - a * D;
                const long long b = step_b[s_];
                if (b < 0 || c > a) continue;
                const int t = step_t[s_];
                if (t == 0 && c < X) continue;                          // x0 is data, not a variable (discret.py:70-78)
                TIO* hv = ar.hes + b * L.nnz_hes;
                const TW sig = ar.sigma ? (TW)ar.sigma[b] : (TW)ar.sigma_scalar;
                const float* ha = pers + s_ * C::P_TOTAL + C::P_HACC;
                float acc = 0.f;
#pragma unroll
                for (int p = 0; p < X; ++p) acc = fmaf((float)ar.lam[b * L.m + t * X + p], ha[p * DD + a * D + c], acc);
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
            for (int idx = gtid; idx < SPT * X; idx += GT) {   // objective-only diagonal of x_H
                const int s_ = idx / X, p = idx - s_ * X;
                const long long b = step_b[s_];
                if (b < 0) continue;
                const int t = step_t[s_];
                if (t != L.H - 1 || L.hes_last_slot[p] < 0) continue;
                const TW sig = ar.sigma ? (TW)ar.sigma[b] : (TW)ar.sigma_scalar;
                ar.hes[b * L.nnz_hes + L.hes_last_slot[p]] = (TIO)(sig * (TW)2 * (TW)ar.quad[(L.H - 1) * X + p]);
            }
        }
        gsync();                                           // per-step state is rewritten by the next tile
        TC_PROF(10);
    }
    TC_PROF_FLUSH;

    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, C::TM_COLS);
}
#endif  // __CUDACC__
