// Tensor-core kernel for WIDE tanh networks (hidden width 256, x_dim + u_dim <= 16: the quadrotor class of BASELINE config C4,
// 16 -> 256 x 4 -> 12): the per-step NLP blocks in ADJOINT FORM, every layer a split-f16 tcgen05 GEMM whose weights are STREAMED
// from L2 through a TMA ring (a 256 x 256 layer is 384 KB of operand images: nothing stays resident).
//
// Formulation (reference integrator/discret.py:32-81, rk4.py:113-285, model/tensorflow.py:49-109).  The forward
// second-order form of nempc_tc.cuh needs 1 + d + d(d+1)/2 = 153 rows per step at d = 16; here a step costs 2 + d rows:
//   phase A  primal forward          128 steps per row tile   h_l = tanh(W_l^T h_{l-1} + b_l)            -> h_l to a per-CTA global scratch
//   phase B  contracted adjoint      128 steps per row tile   g_{l-1} = W_l (s'(a_l) * g_l),  seed g_L = W_out w       (w = lambda for the
//            single-stage integrators, the stage weight w_s of SURVEY 7.3 for RK4)      -> q_l = -2 h_l g_l to the scratch
//   phase C  tangent forward         128 / DP steps per tile, DP tangent rows per step (seed rows R_s, identity for discrete):
//            T_l = W_l^T V_{l-1},  V_l = s'(a_l) * T_l;  J R = W_out^T V_L;  curvature  sum_l T_l^T diag(s''(a_l) g_l) T_l  with
//            s''(a) g = (q s'),  accumulated as 4x4 register blocks on the FFMA2 pipe from a warp-private shared-memory staging of T_l
//            (a warp owns all DP rows of its steps, so no cross-warp traffic until the final sum over the four neuron quarters).
//   RK4: sweep 1 over the stages (A, C without curvature: k_s, dk_s = J_s R_s, R_{s+1} = I + a_{s+1} E dk_s), sweep 2 backwards
//   (B with w_s = c_s lambda + a_{s+1} J_{s+1,x}^T w_{s+1}, C with curvature): H = sum_s R_s^T (sum_p w_{s,p} Hess f_p(z_s)) R_s.
//
// Every product with a weight matrix -- including the thin first layer (K = 16), the output layer (N = 16) and their transposes --
// is one "GEMM" of the same engine:  D[128 rows x N] (f32, tensor memory) = A[128 x K] (tensor memory, TS mode) * B[N x K]^T (ring).
//   * arithmetic: x = hi + lo / 2^11 in f16 (tcx::split_f16); per K step three MMAs  A_hi (2^11 W_hi) + A_lo W_hi + A_hi W_lo  into ONE
//     accumulator that carries the factor 2^11 (the image of 2^11 W_hi replaces the scale-input-d trick of nempc_tc.cuh, which would
//     need a second pass over the streamed W_hi); ~22 mantissa bits.
//   * tensor memory: two 256-column regions X / Y.  GEMM g reads its A operand from one and accumulates into the other; the epilogue
//     converts the accumulator IN PLACE: the 16 f32 columns of neurons [16q, 16q+16) become 8 columns of f16 hi pairs + 8 columns of
//     f16 lo pairs = K step q of the next GEMM's A operand.  Nothing but registers and tensor memory carries a layer to the next.
//   * roles: 16 epilogue warps (lane = row; warp w: lane quadrant w & 3, neuron quarter w >> 2), one TMA producer thread (warp 16)
//     that runs ahead through the ring, one MMA-issuing thread (warp 17).  mbarriers only: ring full / empty, accumulator ready
//     (tcgen05.commit), operand ready (512 arrivals).
#pragma once
#include "nempc_fast.cuh"
#include "nempc_generic.cuh"
#include "nempc_tc_ptx.cuh"

#define NEMPC_WIDE_HW 256
#define NEMPC_WIDE_EPI_WARPS 16
#define NEMPC_WIDE_THREADS (NEMPC_WIDE_EPI_WARPS * 32 + 64)
#ifndef NEMPC_WIDE_NSTAGE
#define NEMPC_WIDE_NSTAGE 5
#endif
#define NEMPC_WIDE_MAXHID 4
#define NEMPC_WIDE_SUP 128                      // steps per super-tile (= rows of a phase A / B tile)

// one streamed operand: `ksteps` ring stages of `stage_bytes` = 96 n bytes: [image 2^11 hi | hi | lo][K chunk 0..1][n][8 halves]
struct WideGemm { uint32_t off, ksteps, n, stage_bytes; };
struct WideNet {
    int nhid;                                   // hidden layers, 2..4, all NEMPC_WIDE_HW wide
    WideGemm in_f, hid_f[NEMPC_WIDE_MAXHID - 1], out_f, out_b, hid_b[NEMPC_WIDE_MAXHID - 1], in_b;
};

// host: half index of element (n, k) of a streamed operand with N rows
inline size_t wide_img_index(int N, int img, int n, int k) {
    return (size_t)(k / 16) * (48 * (size_t)N) + (size_t)img * (16 * (size_t)N) + (size_t)((k % 16) / 8) * (8 * (size_t)N) + (size_t)n * 8 + (k % 8);
}

template <int X_, int U_, int MODE_> struct WideCfg {
    static constexpr int X = X_, U = U_, D = X + U, MODE = MODE_, HW = NEMPC_WIDE_HW;
    static constexpr bool JAC = MODE >= 1, HES = MODE >= 2;
    static constexpr int DP = D <= 4 ? 4 : (D <= 8 ? 8 : 16);          // tangent rows per step (padded to a power of two)
    static constexpr int SPT = 128 / DP;                                 // steps per phase-C tile
    static constexpr int NTILE = NEMPC_WIDE_SUP / SPT;                   // phase-C tiles per super-tile
    static constexpr int NSTAGE = NEMPC_WIDE_NSTAGE, STAGE_BYTES = 96 * HW;
    static constexpr int OFF_RING = 0;
    static constexpr int C_FLOATS = NEMPC_WIDE_MAXHID * HW + 16;        // biases, output bias
    static constexpr int OFF_C = OFF_RING + NSTAGE * STAGE_BYTES;
    static constexpr int STG_WARP = 4 * 40 * 16;                         // per-warp staging of a 32-row x 16-neuron chunk of T_l (4 planes of 40 granules)
    static constexpr int OFF_STG = OFF_C + C_FLOATS * 4;
    static constexpr int OFF_PART = OFF_STG + NEMPC_WIDE_EPI_WARPS * STG_WARP;   // [4 quarters][SPT][DP][DP] partial curvature
    static constexpr int PART_BYTES = HES ? 4 * SPT * DP * DP * 4 : 0;
    static constexpr int TOTAL = OFF_PART + PART_BYTES;
    static constexpr long long SCRATCH_FLOATS = 2LL * NEMPC_WIDE_SUP * NEMPC_WIDE_MAXHID * HW;    // h_l and q_l of one super-tile
    static_assert(D <= 16 && X <= 16, "x_dim + u_dim <= 16");
    static_assert(TOTAL <= 232448, "wide kernel: shared-memory map exceeds 227 KB");
};

#if defined(__CUDACC__)
namespace widex {
using namespace tcx;

__device__ __forceinline__ void mbar_arrive(uint32_t mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory"); }

// tanh with a relative error of a few 1e-7 down to 0: 1 - 2 / (e^{2|x|} + 1) cancels for small |x| (absolute error 1e-7), so a
// degree-9 odd polynomial takes over below 0.25 (truncation 2e-9 there)
__device__ __forceinline__ float tanh_acc(float x) {
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

// 16 f32 values -> 8 words of f16 hi pairs + 8 words of f16 lo pairs (x = hi + lo / 2^11)
__device__ __forceinline__ void split16(const float* x, uint32_t* hi, uint32_t* lo) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const __half2 h2 = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
        const float2 hf = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn((x[2 * i] - hf.x) * NEMPC_TC_LO_SCALE, (x[2 * i + 1] - hf.y) * NEMPC_TC_LO_SCALE);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
}

__device__ __forceinline__ void ld16_global_cg(const float* p, float* v) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(p) + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ void st16_global(float* p, const float* v) {
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// curvature of one 16-neuron chunk: acc[i][i'] += sum_j coef_j T[4 bi + i][j] T[4 bj + i'][j] for the lane's 4x4 block (bi, bj) of
// its step; T (this lane's row, true scale) is exchanged through the warp's staging planes (granule of row r at r + r / 4: the
// eight rows one load instruction touches fall in eight different 16-byte bank groups)
template <int DP>
__device__ __forceinline__ void gram_chunk(float* stg, const float* T, const float* coef, f2* acc, const int lane) {
    constexpr int NB = DP / 4, LPS = NB * NB, ACTIVE = (32 / DP) * LPS;
    {
        const int p = lane + (lane >> 2);
#pragma unroll
        for (int g = 0; g < 4; ++g) *reinterpret_cast<float4*>(stg + (g * 40 + p) * 4) = make_float4(T[4 * g], T[4 * g + 1], T[4 * g + 2], T[4 * g + 3]);
    }
    float cf[16];
    const int sw = (lane / LPS) % (32 / DP);
    if (DP == 16) {
#pragma unroll
        for (int i = 0; i < 16; ++i) cf[i] = coef[i];
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) cf[i] = __shfl_sync(0xffffffffu, coef[i], sw * DP);     // the coefficients of the block's step live in that step's lanes
    }
    __syncwarp();
    if (lane < ACTIVE) {
        const int bl = lane % LPS, bi = bl / NB, bj = bl % NB;
        const int ra = DP * sw + 4 * bi, rb = DP * sw + 4 * bj;
        const float* pa = stg + (ra + (ra >> 2)) * 4;
        const float* pb = stg + (rb + (rb >> 2)) * 4;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const f2 c01 = pk(cf[4 * g], cf[4 * g + 1]), c23 = pk(cf[4 * g + 2], cf[4 * g + 3]);
            f2 ua[4], ub[4], b0[4], b1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 a4 = *reinterpret_cast<const float4*>(pa + (g * 40 + i) * 4);
                ua[i] = mul2(pk(a4.x, a4.y), c01); ub[i] = mul2(pk(a4.z, a4.w), c23);
                const float4 b4 = *reinterpret_cast<const float4*>(pb + (g * 40 + i) * 4);
                b0[i] = pk(b4.x, b4.y); b1[i] = pk(b4.z, b4.w);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[i * 4 + k] = fma2(ub[i], b1[k], fma2(ua[i], b0[k], acc[i * 4 + k]));
        }
    }
    __syncwarp();
}
}  // namespace widex

template <class C, typename TIO>
__global__ void __launch_bounds__(NEMPC_WIDE_THREADS, 1)
nempc_wide_kernel(const unsigned char* __restrict__ blob, const float* __restrict__ cblk, const WideNet net, const StageTable<float> st,
                  const NlpLayout L, const EvalArgs<TIO> ar, float* __restrict__ scratch_all) {
    using namespace widex;
    constexpr int X = C::X, U = C::U, D = C::D, DP = C::DP, SPT = C::SPT, HW = C::HW, NSTAGE = C::NSTAGE;
    constexpr bool JAC = C::JAC, HES = C::HES;
    constexpr float INV = NEMPC_TC_LO_INV;                 // accumulators carry 2^11
    typedef typename WideOf<float, TIO>::type TW;
    extern __shared__ __align__(1024) unsigned char wide_smem[];
    __shared__ uint64_t bars[2 * C::NSTAGE + 2];
    __shared__ uint32_t tmem_holder;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_epi = warp < NEMPC_WIDE_EPI_WARPS, is_prod = warp == NEMPC_WIDE_EPI_WARPS, is_mma = warp == NEMPC_WIDE_EPI_WARPS + 1;
    const int wq = warp & 3, sub = (warp >> 2) & 3;         // lane quadrant of tensor memory, neuron quarter
    const int row = 32 * wq + lane;

    float* cb = reinterpret_cast<float*>(wide_smem + C::OFF_C);
    const float* bias = cb;                                 // [MAXHID][HW]
    const float* bout = cb + NEMPC_WIDE_MAXHID * HW;
    float* stg = reinterpret_cast<float*>(wide_smem + C::OFF_STG + (is_epi ? warp : 0) * C::STG_WARP);
    float* part = reinterpret_cast<float*>(wide_smem + C::OFF_PART);
    float* sh = scratch_all + (long long)blockIdx.x * C::SCRATCH_FLOATS;          // h_l  [128][MAXHID][HW]
    float* sq = sh + (long long)NEMPC_WIDE_SUP * NEMPC_WIDE_MAXHID * HW;           // q_l

    const uint32_t bar0 = smem_u32(&bars[0]);
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
    const uint32_t bar_dready = bar0 + 8u * (2 * NSTAGE), bar_aready = bar_dready + 8u;
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        mbar_init(bar_dready, 1);
        mbar_init(bar_aready, NEMPC_WIDE_EPI_WARPS * 32);
        mbar_fence_init();
    }
    if (is_mma) tmem_alloc(smem_u32(&tmem_holder), 512);
    for (int i = tid; i < C::C_FLOATS; i += NEMPC_WIDE_THREADS) cb[i] = cblk[i];
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_holder;
    const uint32_t lane_off = (uint32_t)(32 * wq) << 16;
    const uint32_t ring = smem_u32(wide_smem + C::OFF_RING);

    const bool unity = (ar.flags & NEMPC_UNITY) != 0;
    const int nhid = net.nhid;
    uint32_t it = 0;            // ring iteration (producer / issuer)
    uint32_t g = 0;             // GEMM counter (all roles): A operand in region g & 1, accumulator in the other
    auto areg = [&](uint32_t gg) { return (gg & 1u) ? 256u : 0u; };
    auto dreg = [&](uint32_t gg) { return (gg & 1u) ? 0u : 256u; };

    // ---- one GEMM, seen by the three roles -----------------------------------------------------------------------------------
    // epilogue role: wait for the accumulator, run `epi(tensor-memory address of this lane quadrant's accumulator rows)`.
    // The A operand of GEMM g must have been published (publish()) exactly once since GEMM g-1.
    auto gemm = [&](const WideGemm& gm, auto&& epi) {
        if (is_prod) {
            if (lane == 0) {
                for (uint32_t ks = 0; ks < gm.ksteps; ++ks, ++it) {
                    const uint32_t slot = it % NSTAGE, round = it / NSTAGE;
                    mbar_wait(bar_empty(slot), (round & 1u) ^ 1u);
                    mbar_expect_tx(bar_full(slot), gm.stage_bytes);
                    bulk_g2s(ring + slot * C::STAGE_BYTES, blob + gm.off + (size_t)ks * gm.stage_bytes, gm.stage_bytes, bar_full(slot));
                }
            }
            __syncwarp();
        } else if (is_mma) {
            if (lane == 0) {
                mbar_wait(bar_aready, g & 1u);
                fence_after_sync();
                const uint32_t ta = tmem + areg(g), td = tmem + dreg(g);
                const uint32_t idesc = make_idesc_f16(128, (int)gm.n);
                const uint32_t lbo = 16u * gm.n, img = 32u * gm.n;
                for (uint32_t ks = 0; ks < gm.ksteps; ++ks, ++it) {
                    const uint32_t slot = it % NSTAGE, round = it / NSTAGE;
                    mbar_wait(bar_full(slot), round & 1u);
                    fence_after_sync();
                    const uint32_t sb = ring + slot * C::STAGE_BYTES;
                    mma_f16_ts(td, ta + 16u * ks, make_desc_kmajor(sb, lbo, 128), idesc, ks != 0);                 // A_hi (2^11 W_hi)
                    mma_f16_ts(td, ta + 16u * ks + 8u, make_desc_kmajor(sb + img, lbo, 128), idesc, 1);           // A_lo W_hi
                    mma_f16_ts(td, ta + 16u * ks, make_desc_kmajor(sb + 2u * img, lbo, 128), idesc, 1);           // A_hi W_lo
                    mma_commit(bar_empty(slot));
                }
                mma_commit(bar_dready);
            }
            __syncwarp();
        } else {
            mbar_wait(bar_dready, g & 1u);
            fence_after_sync();
            epi(tmem + dreg(g) + lane_off);
        }
        ++g;
    };
    // epilogue role: this thread's part of the A operand of GEMM g (region areg(g)) is written
    auto publish = [&]() {
        if (is_epi) { tmem_st_wait(); fence_before_sync(); mbar_arrive(bar_aready); }
    };
    auto epi_sync = [&]() { if (is_epi) asm volatile("bar.sync 1, %0;" ::"n"(NEMPC_WIDE_EPI_WARPS * 32) : "memory"); };
    // K = 16 operand (one K step): hi pairs in columns [0, 8), lo pairs in [8, 16) of region areg(g); written by the quarter-0 warps
    auto put_seed = [&](const float* v16) {
        uint32_t hi[8], lo[8];
        split16(v16, hi, lo);
        const uint32_t a = tmem + areg(g) + lane_off;
        tmem_st8(a, hi);
        tmem_st8(a + 8, lo);
    };

    const long long nsup = (ar.nsteps + NEMPC_WIDE_SUP - 1) / NEMPC_WIDE_SUP;
    for (long long sup = blockIdx.x; sup < nsup; sup += gridDim.x) {
        const long long step_base = sup * NEMPC_WIDE_SUP;
        const int nvalid = (int)((ar.nsteps - step_base) < NEMPC_WIDE_SUP ? (ar.nsteps - step_base) : NEMPC_WIDE_SUP);

        // ============================ phase A: primal forward, row = step ========================================================
        const long long stepA = step_base + row;
        const bool validA = is_epi && row < nvalid;
        long long bA = 0; int tA = 0;
        if (validA) { bA = stepA / L.H; tA = (int)(stepA - bA * L.H); }
        if (is_epi && sub == 0) {
            float zr[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) zr[c] = 0.f;
            if (validA) {
                const TIO* zb = ar.z + bA * (long long)L.n;
#pragma unroll
                for (int c = 0; c < D; ++c)
                    zr[c] = (float)(c < X ? ((tA == 0) ? ar.x0[bA * X + c] : zb[(tA - 1) * X + c]) : zb[L.H * X + tA * U + (c - X)]);
            }
            put_seed(zr);
        }
        publish();
        for (int l = 0; l < nhid; ++l) {
            gemm(l == 0 ? net.in_f : net.hid_f[l - 1], [&](const uint32_t dbase) {
                const float* bl = bias + l * HW;
#pragma unroll 1
                for (int qq = 0; qq < 4; ++qq) {
                    const int col = 64 * sub + 16 * qq;
                    float v[16];
                    tmem_ld16(dbase + col, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = tanh_acc(fmaf(v[i], INV, bl[col + i]));
                    st16_global(sh + ((long long)row * NEMPC_WIDE_MAXHID + l) * HW + col, v);
                    uint32_t hi[8], lo[8];
                    split16(v, hi, lo);
                    tmem_st8(dbase + col, hi);
                    tmem_st8(dbase + col + 8, lo);
                }
            });
            publish();
        }
        gemm(net.out_f, [&](const uint32_t dbase) {
            if (sub != 0) return;
            float v[16];
            tmem_ld16(dbase, v);
            tmem_ld_wait();
            if (validA && ar.resid) {
                const TIO* zb = ar.z + bA * (long long)L.n;
#pragma unroll
                for (int p = 0; p < X; ++p) {
                    const TW xt = (TW)zb[tA * X + p];
                    const TW xp = unity ? (TW)0 : (TW)((tA == 0) ? ar.x0[bA * X + p] : zb[(tA - 1) * X + p]);
                    ar.resid[bA * L.m + tA * X + p] = (TIO)(xp + (TW)fmaf(v[p], INV, bout[p]) - xt);
                }
            }
        });

        // ============================ phase B: lambda-contracted adjoint, row = step =============================================
        if (HES) {
            if (is_epi && sub == 0) {
                float lr[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) lr[c] = 0.f;
                if (validA) {
#pragma unroll
                    for (int p = 0; p < X; ++p) lr[p] = (float)ar.lam[bA * L.m + tA * X + p];
                }
                put_seed(lr);
            }
            publish();
            for (int l = nhid - 1; l >= 0; --l) {
                // accumulator = g_l (adjoint with respect to h_l);  q_l = -2 h_l g_l -> scratch;  next operand u_l = s'(a_l) g_l
                gemm(l == nhid - 1 ? net.out_b : net.hid_b[l], [&](const uint32_t dbase) {
#pragma unroll 1
                    for (int qq = 0; qq < 4; ++qq) {
                        const int col = 64 * sub + 16 * qq;
                        float v[16], h[16];
                        tmem_ld16(dbase + col, v);
                        ld16_global_cg(sh + ((long long)row * NEMPC_WIDE_MAXHID + l) * HW + col, h);
                        tmem_ld_wait();
                        float qv[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float gl = v[i] * INV;
                            qv[i] = -2.f * h[i] * gl;
                            v[i] = fmaf(-h[i], h[i], 1.f) * gl;
                        }
                        st16_global(sq + ((long long)row * NEMPC_WIDE_MAXHID + l) * HW + col, qv);
                        if (l > 0) {
                            uint32_t hi[8], lo[8];
                            split16(v, hi, lo);
                            tmem_st8(dbase + col, hi);
                            tmem_st8(dbase + col + 8, lo);
                        }
                    }
                });
                if (l > 0) publish();
            }
        }
        epi_sync();                                        // the scratch of this super-tile is complete (bar.sync orders it at CTA scope)

        // ============================ phase C: tangent forward (+ curvature), row = (step, tangent column) =========================
        if (JAC) {
            const int ntile = (nvalid + SPT - 1) / SPT;
            for (int ti = 0; ti < ntile; ++ti) {
                const int sidx = ti * SPT + row / DP, cc = row % DP;          // step inside the super-tile, tangent column
                const bool validC = is_epi && sidx < nvalid && cc < D;
                f2 acc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = pk(0.f, 0.f);
                if (is_epi && sub == 0) {
                    float e[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) e[c] = (c == cc && cc < D) ? 1.f : 0.f;
                    put_seed(e);
                }
                publish();
                for (int l = 0; l < nhid; ++l) {
                    gemm(l == 0 ? net.in_f : net.hid_f[l - 1], [&](const uint32_t dbase) {
#pragma unroll 1
                        for (int qq = 0; qq < 4; ++qq) {
                            const int col = 64 * sub + 16 * qq;
                            float v[16], h[16];
                            tmem_ld16(dbase + col, v);
                            ld16_global_cg(sh + ((long long)sidx * NEMPC_WIDE_MAXHID + l) * HW + col, h);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) { v[i] *= INV; h[i] = fmaf(-h[i], h[i], 1.f); }          // raw tangent T_l, s'(a_l)
                            if (HES) {
                                float cf[16];
                                ld16_global_cg(sq + ((long long)sidx * NEMPC_WIDE_MAXHID + l) * HW + col, cf);
#pragma unroll
                                for (int i = 0; i < 16; ++i) cf[i] *= h[i];                                          // s''(a_l) g_l
                                gram_chunk<DP>(stg, v, cf, acc, lane);
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] *= h[i];                                               // V_l
                            uint32_t hi[8], lo[8];
                            split16(v, hi, lo);
                            tmem_st8(dbase + col, hi);
                            tmem_st8(dbase + col + 8, lo);
                        }
                    });
                    publish();
                }
                gemm(net.out_f, [&](const uint32_t dbase) {
                    if (sub != 0) return;
                    float v[16];
                    tmem_ld16(dbase, v);
                    tmem_ld_wait();
                    if (validC && ar.jac) {
                        const long long step = step_base + sidx;
                        const long long b = step / L.H;
                        const int t = (int)(step - b * L.H);
                        TIO* jv = ar.jac + b * L.nnz_jac;
#pragma unroll
                        for (int p = 0; p < X; ++p) {
                            const TW val = (TW)(v[p] * INV) + ((!unity && cc == p) ? (TW)1 : (TW)0);
                            if (cc < X) { if (t > 0) jv[jac_slot_A(L, t, p, cc)] = (TIO)val; }
                            else jv[jac_slot_B(L, t, p, cc - X)] = (TIO)val;
                            if (cc == 0) jv[jac_slot_minus1(L, t, p)] = (TIO)-1;
                        }
                    }
                });
                if (HES) {
                    // ---- sum the four neuron quarters, scatter the lower triangle (same slots as nempc_generic.cuh) -----------------
                    if (is_epi) {
                        constexpr int NB = DP / 4, LPS = NB * NB, ACTIVE = (32 / DP) * LPS;
                        if (lane < ACTIVE) {
                            const int sw = lane / LPS, bl = lane % LPS, bi = bl / NB, bj = bl % NB;
                            const int s8 = (32 * wq) / DP + sw;
                            float* pp = part + ((sub * SPT + s8) * DP + 4 * bi) * DP + 4 * bj;
#pragma unroll
                            for (int i = 0; i < 4; ++i)
#pragma unroll
                                for (int k = 0; k < 4; ++k) pp[i * DP + k] = f2lo(acc[i * 4 + k]) + f2hi(acc[i * 4 + k]);
                        }
                    }
                    epi_sync();
                    if (is_epi && ar.hes) {
                        for (int idx = tid; idx < SPT * DP * DP; idx += NEMPC_WIDE_EPI_WARPS * 32) {
                            const int s8 = idx / (DP * DP), r = idx - s8 * (DP * DP), a = r / DP, c = r - a * DP;
                            const int sx = ti * SPT + s8;
                            if (sx >= nvalid || c > a || a >= D) continue;
                            const long long step = step_base + sx;
                            const long long b = step / L.H;
                            const int t = (int)(step - b * L.H);
                            if (t == 0 && c < X) continue;                          // x0 is data, not a variable (discret.py:70-78)
                            float sum = 0.f;
#pragma unroll
                            for (int k = 0; k < 4; ++k) sum += part[((k * SPT + s8) * DP + a) * DP + c];
                            TIO* hv = ar.hes + b * L.nnz_hes;
                            const TW sig = ar.sigma ? (TW)ar.sigma[b] : (TW)ar.sigma_scalar;
                            TW val = (TW)sum;
                            int slot;
                            if (a < X) {
                                slot = hes_slot_xx(L, t, a, c);
                                if (a == c && ar.quad) val += sig * (TW)2 * (TW)ar.quad[(t - 1) * X + a];
                            } else if (c < X) {
                                slot = hes_slot_ux(L, t, a - X, c);
                            } else {
                                slot = hes_slot_uu(L, t, a - X, c - X);
                                if (a == c && ar.quad) val += sig * (TW)2 * (TW)ar.quad[L.H * X + t * U + (a - X)];
                            }
                            hv[slot] = (TIO)val;
                        }
                        for (int idx = tid; idx < SPT * X; idx += NEMPC_WIDE_EPI_WARPS * 32) {   // objective-only diagonal of x_H
                            const int s8 = idx / X, p = idx - s8 * X;
                            const int sx = ti * SPT + s8;
                            if (sx >= nvalid) continue;
                            const long long step = step_base + sx;
                            const long long b = step / L.H;
                            const int t = (int)(step - b * L.H);
                            if (t != L.H - 1 || L.hes_last_slot[p] < 0) continue;
                            const TW sig = ar.sigma ? (TW)ar.sigma[b] : (TW)ar.sigma_scalar;
                            ar.hes[b * L.nnz_hes + L.hes_last_slot[p]] = (TIO)(sig * (TW)2 * (TW)ar.quad[(L.H - 1) * X + p]);
                        }
                    }
                    epi_sync();                                    // `part` is rewritten by the next tile
                }
            }
        }
        epi_sync();                                                // the scratch is rewritten by the next super-tile
    }

    fence_before_sync();
    __syncthreads();
    if (is_mma) tmem_dealloc(tmem, 512);
}
#endif  // __CUDACC__
