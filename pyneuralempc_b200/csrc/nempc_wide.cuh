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
//            T_l = W_l^T V_{l-1},  V_l = s'(a_l) * T_l;  J R = W_out^T V_L;  curvature  sum_l T_l^T diag(s''(a_l) g_l) T_l  =  sum_l U_l V_l^T
//            with U_l = (-2 h_l g_l) * T_l: per 16-neuron chunk the warp stages the split-f16 rows of U and V in its own shared-memory planes
//            and accumulates the 16 x 16 blocks with warp-level mma.sync.m16n8k16 (three products, as in the big GEMMs) -- a warp owns all DP
//            rows of its steps, so there is no cross-warp traffic until the final sum over the four 16-neuron slices of each quarter.
//            (-DNEMPC_WIDE_GRAM_MMA=0 keeps the first version: 4x4 register blocks on the FFMA2 pipe, 1.44x slower on C4.)
//   RK4: sweep 1 over the stages (A, C without curvature: k_s, dk_s = J_s R_s, R_{s+1} = I + a_{s+1} E dk_s), sweep 2 backwards
//   (B with w_s = c_s lambda + a_{s+1} J_{s+1,x}^T w_{s+1}, C with curvature): H = sum_s R_s^T (sum_p w_{s,p} Hess f_p(z_s)) R_s.
//
// Every product with a weight matrix -- including the thin first layer (K = 16), the output layer (N = 16) and their transposes --
// is one "GEMM" of the same engine:  D[128 rows x N] (f32, tensor memory) = A[128 x K] (tensor memory, TS mode) * B[N x K]^T (ring).
//   * arithmetic: x = hi + lo / 2^11 in f16 (tcx::split_f16); three products  A_hi (2^11 W_hi) + A_lo W_hi + A_hi W_lo  into ONE
//     accumulator that carries the factor 2^11 (the image of 2^11 W_hi replaces the scale-input-d trick of nempc_tc.cuh).  The two
//     correction products of ALL K steps are issued first, while the accumulator is small, then the main products: the tensor core
//     rounds the f32 accumulator once per MMA, so only the 16 main MMAs round at full magnitude (instead of 48 when interleaved);
//     ~22 mantissa bits.
//   * tensor memory: two 256-column regions X / Y.  GEMM g reads its A operand from one and accumulates into the other; the epilogue
//     converts the accumulator IN PLACE: the 16 f32 columns of neurons [16q, 16q+16) become 8 columns of f16 hi pairs + 8 columns of
//     f16 lo pairs = K step q of the next GEMM's A operand.  Nothing but registers and tensor memory carries a layer to the next.
//   * CTA PAIRS (cta_group::2).  One SM pulls ~26 B/clk out of L2 (the same ~6.3 KB/clk chip-wide whatever the path), i.e. 15 k clk for
//     the 384 KB of one layer against 6.1 k clk of MMA time: a single CTA is bound by its weight stream (measured: ring waits, linear
//     scaling with the CTA count).  So two CTAs of a cluster work on two 128-row tiles as ONE M = 256 MMA: each CTA streams only its
//     half of the B rows (tcgen05.mma.cta_group::2 exchanges the halves between the two SMs), which halves the bytes per SM and row.
//     The two CTAs run the same GEMM sequence on their own super-tile (a CTA without work runs it on zero rows).
//   * the operand is handed over QUARTER BY QUARTER: an epilogue warp converts its 16-neuron slice of every 64-neuron quarter in turn and
//     arrives on that quarter's barrier, the issuer consumes quarter q (K steps 4q .. 4q+3, corrections first within the quarter) as soon as
//     all 32 warps of the pair have arrived: the MMAs of layer l+1 run under the rest of the epilogue of layer l.
//   * roles per CTA: 16 epilogue warps (lane = row; warp w: lane quadrant w & 3, 16-neuron slice w >> 2 of each quarter), one TMA producer thread
//     (warp 16) that runs ahead through the ring, and in warp 17 the MMA-issuing thread (leader CTA) or a thread that forwards the
//     peer's ring-full signals to the leader.  mbarriers only: ring full (leader: own TMA + peer's forward) / empty and accumulator
//     ready (tcgen05.commit multicast to both CTAs), operand-quarter ready (2 x 16 warp arrivals on the leader's barriers, the peer's arrive remotely).
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
#ifndef NEMPC_WIDE_GRAM_MMA
#define NEMPC_WIDE_GRAM_MMA 1                  // curvature blocks on the warp-level tensor path (mma.sync, split-f16) instead of FFMA2
#endif

// one streamed operand of a GEMM with n rows and `ksteps` K steps of 16, cut in two halves of nh = n / 2 rows (CTA r of the pair
// streams rows [r nh, (r + 1) nh) from off[r]); an image of one K step = [K chunk 0..1][nh][8 halves] = 32 nh bytes:
//   pass 1, one ring stage per TWO K steps:   [hi (ks) | lo (ks) | hi (ks + 1) | lo (ks + 1)]   128 nh bytes (64 nh for a lone K step)
//   pass 2, one ring stage per FOUR K steps:  [2^11 hi (ks .. ks + 3)]                          128 nh bytes
// (four MMAs per ring stage: the lone issuing thread pays ~300 clk of barrier wait + commit per stage)
struct WideGemm { uint32_t off[2], ksteps, n; };
struct WideNet {
    int nhid;                                   // hidden layers, 2..4, all NEMPC_WIDE_HW wide
    WideGemm in_f, hid_f[NEMPC_WIDE_MAXHID - 1], out_f, out_b, hid_b[NEMPC_WIDE_MAXHID - 1], in_b;
};

#define NEMPC_WIDE_NOUT 32                      // rows of the thin output operands (x_dim or d, zero padded): cta_group::2 needs N % 32 == 0
// host: half index of element (n, k) of one CTA's half (N rows) of a streamed operand with K16 K steps; img 0 = 2^11 hi, 1 = hi, 2 = lo
inline size_t wide_img_index(int N, int K16, int img, int n, int k) {
    const size_t in_img = (size_t)((k % 16) / 8) * (8 * (size_t)N) + (size_t)n * 8 + (k % 8), ks = (size_t)(k / 16);
    if (img == 0) return (size_t)K16 * 32 * N + ks * 16 * N + in_img;
    return ks * 32 * N + (img == 2 ? 16 * (size_t)N : 0) + in_img;
}

// X_ = U_ = 0: x_dim / u_dim are read from the layout at run time (any x_dim + u_dim <= DPR_ <= 16); the thin per-step loops then run to 16
// under a predicate, everything else is the same code
template <int X_, int U_, int MODE_, bool RK4_ = false, int HW_ = NEMPC_WIDE_HW, int DPR_ = 16> struct WideCfg {
    static constexpr int X = X_, U = U_, D = X_ ? X_ + U_ : DPR_, MODE = MODE_, HW = HW_;      // hidden width 256 (C4 class) or 128 (C3 class)
    static constexpr int XM = X_ ? X_ : 16;                                // bound of the unrolled per-output loops
    static constexpr int NQ = HW / 64;                                     // 64-neuron operand quarters (= groups of 4 K steps) per hidden layer
    static_assert(HW == 256 || HW == 128, "hidden width 256 or 128");
    static constexpr bool RK4 = RK4_;                                    // four stages: per-stage k_s, dk_s and adjoint weights in the scratch
    static constexpr bool JAC = MODE >= 1, HES = MODE >= 2;
    static constexpr int DP = D <= 4 ? 4 : (D <= 8 ? 8 : 16);          // tangent rows per step (padded to a power of two)
    static_assert(X_ != 0 || DPR_ == 4 || DPR_ == 8 || DPR_ == 16, "run-time shapes: tangent rows 4, 8 or 16");
    static constexpr int SPT = 128 / DP;                                 // steps per phase-C tile
    static constexpr int NTILE = NEMPC_WIDE_SUP / SPT;                   // phase-C tiles per super-tile
    static constexpr int STAGE_BYTES = 128 * (HW / 2);                   // four K-step images of one CTA's half of the B rows
    static constexpr int C_FLOATS = NEMPC_WIDE_MAXHID * HW + 16;        // biases, output bias
#if NEMPC_WIDE_GRAM_MMA
    static constexpr int STG_WARP = HES ? 4 * 1024 : 0;                  // per-warp staging of a 32-row x 16-neuron chunk: f16 planes U_hi, U_lo, V_hi, V_lo, [K half][row][8 halves]
#else
    static constexpr int STG_WARP = HES ? 4 * 32 * 16 : 0;              // per-warp staging of a 32-row x 16-neuron chunk of T_l (4 planes of 32 granules)
#endif
    static constexpr int SPW = 32 / DP;                                  // steps per epilogue warp in phase C
    static constexpr int SC_WARP = JAC ? (HES ? 2 : 1) * SPW * 64 * 4 : 0;   // per-warp copy of s'(a_l) (and the curvature coefficients) of its steps and neuron quarter
    static constexpr int PART_BYTES = HES ? 4 * SPT * DP * DP * 4 : 0;  // [4 quarters][SPT][DP][DP] partial curvature: aliases the staging area
    static constexpr int STG_BYTES = NEMPC_WIDE_EPI_WARPS * STG_WARP > PART_BYTES ? NEMPC_WIDE_EPI_WARPS * STG_WARP : PART_BYTES;
    static constexpr int LS_BYTES = HES ? NEMPC_WIDE_SUP * 4 : 0;        // per-step multiplier scale of the super-tile (see phase B)
    static constexpr int FIXED = C_FLOATS * 4 + STG_BYTES + NEMPC_WIDE_EPI_WARPS * SC_WARP + LS_BYTES;
    static constexpr int NSTAGE = (232448 - 1024 - FIXED) / STAGE_BYTES < 10 ? (232448 - 1024 - FIXED) / STAGE_BYTES : 10;
    static constexpr int OFF_RING = 0;
    static constexpr int OFF_C = OFF_RING + NSTAGE * STAGE_BYTES;
    static constexpr int OFF_STG = OFF_C + C_FLOATS * 4;
    static constexpr int OFF_PART = OFF_STG;
    static constexpr int OFF_SC = OFF_STG + STG_BYTES;
    static constexpr int OFF_LS = OFF_SC + NEMPC_WIDE_EPI_WARPS * SC_WARP;
    static constexpr int TOTAL = OFF_LS + LS_BYTES;
    static_assert(NSTAGE >= 4, "wide kernel: weight ring too shallow");
    static constexpr long long SCRATCH_NET = 2LL * NEMPC_WIDE_SUP * NEMPC_WIDE_MAXHID * HW;       // h_l and q_l of one super-tile
    // RK4: k_s [4][128][16], adjoint weights w [128][16], dk_s [3][128][16 columns][16], running sum_s c_s dk_s [128][16][16]
    static constexpr long long SCRATCH_FLOATS = SCRATCH_NET + (RK4 ? 5LL * NEMPC_WIDE_SUP * 16 + 4LL * NEMPC_WIDE_SUP * 256 : 0);
    static_assert(D <= 16 && X <= 16, "x_dim + u_dim <= 16");
    static_assert(TOTAL <= 232448, "wide kernel: shared-memory map exceeds 227 KB");
};

#if defined(__CUDACC__)
namespace widex {
using namespace tcx;

// one lane of a converged warp (the branch on it keeps the surrounding values in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory"); }

// ---- CTA pair (cluster of two, cta_group::2) ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t holder_smem, uint32_t ncols) {     // one warp of EACH CTA of the pair, same holder offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 over the pair: A rows [128 r, 128 r + 128) from CTA r's tensor memory, B rows [N/2 r, N/2 (r+1)) from CTA r's shared memory
// (same descriptor offset in both), accumulator rows in each CTA's own tensor memory; issued by one thread of the leader CTA
__device__ __forceinline__ void mma2_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs when all MMAs issued so far have completed
__device__ __forceinline__ void mma2_commit(uint32_t mbar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(mbar), "h"(mask) : "memory");
}

// tanh with a relative error of a few 1e-7 down to 0 (nempc_fast.cuh)
__device__ __forceinline__ float tanh_acc(float x) { return fast_tanh(x); }

// 8 f32 pairs -> 8 words of f16 hi pairs + 8 words of f16 lo pairs (x = hi + lo / 2^11), packed arithmetic
__device__ __forceinline__ void split16p(const f2* x, uint32_t* hi, uint32_t* lo) {
    const f2 sc = pk(NEMPC_TC_LO_SCALE, NEMPC_TC_LO_SCALE), nsc = pk(-NEMPC_TC_LO_SCALE, -NEMPC_TC_LO_SCALE);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const __half2 h2 = __floats2half2_rn(f2lo(x[i]), f2hi(x[i]));
        const float2 hf = __half22float2(h2);
        const f2 r = fma2(x[i], sc, mul2(pk(hf.x, hf.y), nsc));            // (x - hi) * 2^11, exact
        const __half2 l2 = __floats2half2_rn(f2lo(r), f2hi(r));
        hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
}
__device__ __forceinline__ void split16(const float* x, uint32_t* hi, uint32_t* lo) {
    f2 x2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x2[i] = pk(x[2 * i], x[2 * i + 1]);
    split16p(x2, hi, lo);
}

// tensor-memory load without the wait (software pipelining: the next chunk's accumulator is in flight while this one is processed);
// tmem_ld_pin after tcgen05.wait::ld ties every later use of the registers to the wait
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_pin(uint32_t* r) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                      "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
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
// its step.  T = this lane's accumulator row (any common scale: `cfs` carries its inverse square) is exchanged through the warp's
// staging planes (32 granules of 16 bytes per plane, row r at granule r ^ 2 [r / 8 odd]: the eight consecutive rows of a store phase
// and the rows 4 apart of a load phase fall in different 16-byte bank groups); cfs = shared-memory rows [step of the warp][64 neurons of the warp's quarter] of the coefficients, cc0 = first neuron.
template <int DP>
__device__ __forceinline__ void gram_chunk(float* stg, const uint32_t* T, const float* cfs, const int cc0, f2* acc, const int lane) {
    constexpr int NB = DP / 4, LPS = NB * NB, ACTIVE = (32 / DP) * LPS;
    {
        const int p = lane ^ ((lane >> 2) & 2);
#pragma unroll
        for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(stg + (g * 32 + p) * 4) = make_uint4(T[4 * g], T[4 * g + 1], T[4 * g + 2], T[4 * g + 3]);
    }
    __syncwarp();
    if (lane < ACTIVE) {
        const int sw = lane / LPS, bl = lane % LPS, bi = bl / NB, bj = bl % NB;
        const int ra = DP * sw + 4 * bi, rb = DP * sw + 4 * bj;
        // rows ra + i, i < 4, sit at granules ra + (i ^ swz): two runs of two consecutive granules
        const int sza = (ra >> 2) & 2, szb = (rb >> 2) & 2;
        const float* pa[2] = {stg + (ra + sza) * 4, stg + (ra + (2 ^ sza)) * 4};
        const float* pb[2] = {stg + (rb + szb) * 4, stg + (rb + (2 ^ szb)) * 4};
        const float* cf = cfs + sw * 64 + cc0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float4 c4 = *reinterpret_cast<const float4*>(cf + 4 * g);
            const f2 c01 = pk(c4.x, c4.y), c23 = pk(c4.z, c4.w);
            f2 ua[4], ub[4], b0[4], b1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 a4 = *reinterpret_cast<const float4*>(pa[i >> 1] + (g * 32 + (i & 1)) * 4);
                ua[i] = mul2(pk(a4.x, a4.y), c01); ub[i] = mul2(pk(a4.z, a4.w), c23);
                const float4 b4 = *reinterpret_cast<const float4*>(pb[i >> 1] + (g * 32 + (i & 1)) * 4);
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

// ---- curvature blocks on the warp-level tensor path ----------------------------------------------------------------------------
// G[a][b] += sum_j U[a][j] V[b][j] over the 16 neurons of a chunk, for the two 16-row tiles of the warp (U = (-2 h g) T, V = s' T, so
// that U V = s'' g T T), as split-f16 mma.sync.m16n8k16 products:  U_hi V_hi  into accM,  U_lo V_hi + U_hi V_lo  (both carry 2^11) into
// accC;  G = accM + 2^-11 accC.  The rows reach the fragment layout through the warp's staging planes and ldmatrix:
//   plane p (0 U_hi, 1 U_lo, 2 V_hi, 3 V_lo) = 1 KB:  [K half kh][row r][8 halves]  (16 B per row: conflict-free stores and ldmatrix rows)
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t* r) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float* d, const uint32_t* a, const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// accM / accC: [tile 0..1][column half 0..1][4] in the m16n8 accumulator layout (rows lane/4 and lane/4 + 8, columns 2 (lane%4), +1)
__device__ __forceinline__ void gram_mma_chunk(const uint32_t stg_addr, uint4* stg, const uint32_t* uhi, const uint32_t* ulo,
                                               const uint32_t* vhi, const uint32_t* vlo, float* accM, float* accC, const int lane) {
    stg[0 * 64 + lane] = make_uint4(uhi[0], uhi[1], uhi[2], uhi[3]); stg[0 * 64 + 32 + lane] = make_uint4(uhi[4], uhi[5], uhi[6], uhi[7]);
    stg[1 * 64 + lane] = make_uint4(ulo[0], ulo[1], ulo[2], ulo[3]); stg[1 * 64 + 32 + lane] = make_uint4(ulo[4], ulo[5], ulo[6], ulo[7]);
    stg[2 * 64 + lane] = make_uint4(vhi[0], vhi[1], vhi[2], vhi[3]); stg[2 * 64 + 32 + lane] = make_uint4(vhi[4], vhi[5], vhi[6], vhi[7]);
    stg[3 * 64 + lane] = make_uint4(vlo[0], vlo[1], vlo[2], vlo[3]); stg[3 * 64 + 32 + lane] = make_uint4(vlo[4], vlo[5], vlo[6], vlo[7]);
    __syncwarp();
    const int mi = lane >> 3, rr = lane & 7;
    // A fragments: matrices (rows 0-7, kh 0), (rows 8-15, kh 0), (rows 0-7, kh 1), (rows 8-15, kh 1);  B fragments of both column halves:
    // (rows 0-7, kh 0), (rows 0-7, kh 1), (rows 8-15, kh 0), (rows 8-15, kh 1)
    const uint32_t offA = (uint32_t)(((mi >> 1) * 32 + 8 * (mi & 1) + rr) * 16);
    const uint32_t offB = (uint32_t)(((mi & 1) * 32 + 8 * (mi >> 1) + rr) * 16);
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
        uint32_t ah[4], al[4], bh[4], bl[4];
        ldmatrix_x4(stg_addr + 0 * 1024 + offA + tile * 256, ah);
        ldmatrix_x4(stg_addr + 1 * 1024 + offA + tile * 256, al);
        ldmatrix_x4(stg_addr + 2 * 1024 + offB + tile * 256, bh);
        ldmatrix_x4(stg_addr + 3 * 1024 + offB + tile * 256, bl);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            mma_16816(accM + (tile * 2 + h) * 4, ah, bh[2 * h], bh[2 * h + 1]);
            mma_16816(accC + (tile * 2 + h) * 4, al, bh[2 * h], bh[2 * h + 1]);
            mma_16816(accC + (tile * 2 + h) * 4, ah, bl[2 * h], bl[2 * h + 1]);
        }
    }
    __syncwarp();                                              // the planes are rewritten by the next chunk
}
}  // namespace widex

// -DNEMPC_WIDE_PROFILE (development builds): cycles per phase of the three roles, summed over all CTAs
//   0 issuer: wait operand   1 issuer: wait ring stage   2 issuer: issue + commit   3 GEMMs
//   4 epilogue thread 0: pre   5 wait accumulator   6 epilogue body   7 between GEMMs (publish, seeds, scatter)
//   8 producer: wait free slot   9 producer: issue
#ifdef NEMPC_WIDE_PROFILE
static __device__ unsigned long long nempc_wide_prof[16];
#define WPROF_DECL long long wp_t = clock64(); unsigned long long wp_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define WPROF(i) do { const long long t_ = clock64(); wp_acc[i] += (unsigned long long)(t_ - wp_t); wp_t = t_; } while (0)
#define WPROF_COUNT(i) do { wp_acc[i] += 1; } while (0)
#define WPROF_FLUSH do { if (lane == 0 && (warp == 0 || is_prod || is_mma)) for (int i_ = 0; i_ < 16; ++i_) if (wp_acc[i_]) atomicAdd(&nempc_wide_prof[i_], wp_acc[i_]); } while (0)
#else
#define WPROF_DECL
#define WPROF(i) do {} while (0)
#define WPROF_COUNT(i) do {} while (0)
#define WPROF_FLUSH do {} while (0)
#endif

template <class C, typename TIO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NEMPC_WIDE_THREADS, 1)     // 18 warps are allocated as 20: 96 registers per thread (112 does not launch)
nempc_wide_kernel(const unsigned char* __restrict__ blob, const float* __restrict__ cblk, const WideNet net, const StageTable<float> st,
                  const NlpLayout L, const EvalArgs<TIO> ar, float* __restrict__ scratch_all) {
    using namespace widex;
    constexpr int XM = C::XM, DP = C::DP, SPT = C::SPT, SPW = C::SPW, HW = C::HW, NSTAGE = C::NSTAGE;
    const int X = C::X ? C::X : L.x, U = C::X ? C::U : L.u, D = X + U;        // compile-time constants for the listed shapes
#define WIDE_FOR_X(p) _Pragma("unroll") for (int p = 0; p < XM; ++p) if (p < X)
    constexpr bool JAC = C::JAC, HES = C::HES, RK4 = C::RK4;
    constexpr float INV = NEMPC_TC_LO_INV;                 // accumulators carry 2^11
    typedef typename WideOf<float, TIO>::type TW;
    extern __shared__ __align__(1024) unsigned char wide_smem[];
    __shared__ uint64_t bars[2 * C::NSTAGE + 1 + 4];
    __shared__ uint32_t tmem_holder;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_epi = warp < NEMPC_WIDE_EPI_WARPS, is_prod = warp == NEMPC_WIDE_EPI_WARPS, is_mma = warp == NEMPC_WIDE_EPI_WARPS + 1;
    const int wq = warp & 3, sub = (warp >> 2) & 3;         // lane quadrant of tensor memory, neuron quarter
    const int row = 32 * wq + lane;

    float* cb = reinterpret_cast<float*>(wide_smem + C::OFF_C);
    const float* bias = cb;                                 // [MAXHID][HW]
    const float* bout = cb + NEMPC_WIDE_MAXHID * HW;
    float* stg = reinterpret_cast<float*>(wide_smem + C::OFF_STG + (is_epi ? warp : 0) * C::STG_WARP);
    float* scs = reinterpret_cast<float*>(wide_smem + C::OFF_SC + (is_epi ? warp : 0) * C::SC_WARP);   // [s' 2^-11 | coef 2^-22][SPW][64]
    float* part = reinterpret_cast<float*>(wide_smem + C::OFF_PART);
    float* lscale = reinterpret_cast<float*>(wide_smem + C::OFF_LS);      // [128] max |lambda| of each step of the super-tile
    // per-CTA scratch: sa = h_l (phase A -> B), overwritten by s'(a_l) (phase B, or phase A when no Hessian is wanted); sq = s''(a_l) g_l
    const uint32_t rank = cluster_ctarank();                // 0 = leader (issues the MMAs of the pair)
    float* sa = scratch_all + (long long)blockIdx.x * C::SCRATCH_FLOATS;           // [128][MAXHID][HW]
    float* sq = sa + (long long)NEMPC_WIDE_SUP * NEMPC_WIDE_MAXHID * HW;
    float* sks = sa + C::SCRATCH_NET;                                  // RK4 stage state, see WideCfg::SCRATCH_FLOATS
    float* sw = sks + 4 * NEMPC_WIDE_SUP * 16;
    float* sdk = sw + NEMPC_WIDE_SUP * 16;
    float* sdkacc = sdk + 3 * NEMPC_WIDE_SUP * 256;

    const uint32_t bar0 = smem_u32(&bars[0]);
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
    const uint32_t bar_dready = bar0 + 8u * (2 * NSTAGE), bar_aready = bar_dready + 8u;   // bar_aready + 8 q: operand quarter q (K steps 4q .. 4q+3) is ready
    if (tid == 0) {
        // leader: a ring stage is full when its own TMA bytes have landed AND the peer forwarded the same for its half
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full(s), rank == 0 ? 2 : 1); mbar_init(bar_empty(s), 1); }
        mbar_init(bar_dready, 1);
        for (int q = 0; q < 4; ++q) mbar_init(bar_aready + 8u * q, 2 * NEMPC_WIDE_EPI_WARPS);   // one arrival per epilogue warp of BOTH CTAs (used in the leader only)
        mbar_fence_init();
    }
    if (is_mma) tmem_alloc2(smem_u32(&tmem_holder), 512);
    for (int i = tid; i < C::C_FLOATS; i += NEMPC_WIDE_THREADS) cb[i] = cblk[i];
    fence_before_sync();
    __syncthreads();
    cluster_sync_all();                                     // the peer's barriers are initialised before anything arrives on them
    fence_after_sync();
    const uint32_t tmem = tmem_holder;
    const uint32_t lane_off = (uint32_t)(32 * wq) << 16;
    const uint32_t ring = smem_u32(wide_smem + C::OFF_RING);
    const uint32_t aready_leader = mapa_u32(bar_aready, 0);

    const bool unity = (ar.flags & NEMPC_UNITY) != 0;
    const int nhid = net.nhid;
    uint32_t it = 0;            // ring iteration (producer / issuer)
    uint32_t g = 0;             // GEMM counter (all roles): A operand in region g & 1, accumulator in the other
    uint32_t aphase = 0;        // issuer: phase bit of each operand-quarter barrier (quarters 1..3 only see the K = 256 GEMMs)
    auto areg = [&](uint32_t gg) { return (gg & 1u) ? 256u : 0u; };
    auto dreg = [&](uint32_t gg) { return (gg & 1u) ? 0u : 256u; };

    // ---- one GEMM, seen by the three roles -----------------------------------------------------------------------------------
    // epilogue role: `pre()` (work that does not need the accumulator: it overlaps the MMAs), wait for the accumulator, then
    // `epi(tensor-memory address of this lane quadrant's accumulator rows)`.
    // The A operand of GEMM g must have been published (publish()) exactly once since GEMM g-1.
    WPROF_DECL
    auto gemm = [&](const WideGemm& gm, auto&& pre, auto&& epi) {
        if (is_prod) {
            if (lane == 0) {
                const uint32_t img = 16u * gm.n;                                               // one K-step image of this CTA's half of the rows
                const unsigned char* src0 = blob + gm.off[rank];
                const uint32_t main0 = gm.ksteps * 2u * img;                                   // the 2^11 hi images follow the [hi | lo] pairs of all K steps
                for (uint32_t k0 = 0; k0 < gm.ksteps; k0 += 4) {                               // operand quarter by quarter, in the order the issuer consumes
                    const uint32_t k1 = k0 + 4u < gm.ksteps ? k0 + 4u : gm.ksteps;
                    for (uint32_t ks = k0; ks < k1; ks += 2, ++it) {                           // corrections: [hi | lo] of two K steps
                        const uint32_t slot = it % NSTAGE, round = it / NSTAGE, bytes = (ks + 1 < k1 ? 4u : 2u) * img;
                        mbar_wait(bar_empty(slot), (round & 1u) ^ 1u);
                        WPROF(8);
                        mbar_expect_tx(bar_full(slot), bytes);
                        bulk_g2s(ring + slot * C::STAGE_BYTES, src0 + ks * 2u * img, bytes, bar_full(slot));
                        WPROF(9);
                    }
                    {                                                                          // main products: 2^11 hi of the quarter's K steps
                        const uint32_t slot = it % NSTAGE, round = it / NSTAGE, bytes = (k1 - k0) * img;
                        mbar_wait(bar_empty(slot), (round & 1u) ^ 1u);
                        WPROF(8);
                        mbar_expect_tx(bar_full(slot), bytes);
                        bulk_g2s(ring + slot * C::STAGE_BYTES, src0 + main0 + k0 * img, bytes, bar_full(slot));
                        WPROF(9);
                        ++it;
                    }
                }
            }
            __syncwarp();
        } else if (is_mma) {
            if (rank == 0) {
                // the whole warp runs the loop (converged: operands stay in uniform registers), one elected lane issues.
                // The A operand arrives QUARTER BY QUARTER (64 neurons = 4 K steps): the epilogue of the previous layer publishes a quarter as
                // soon as it is converted, so these MMAs run under the rest of that epilogue.  Within a quarter the corrections go first.
                WPROF(2);
                WPROF_COUNT(3);
                const uint32_t ta = tmem + areg(g), td = tmem + dreg(g);
                const uint32_t idesc = make_idesc_f16(256, (int)gm.n);
                const uint32_t lbo = 8u * gm.n, img = 16u * gm.n;                               // per-CTA half: n / 2 rows
                for (uint32_t k0 = 0; k0 < gm.ksteps; k0 += 4) {
                    const uint32_t k1 = k0 + 4u < gm.ksteps ? k0 + 4u : gm.ksteps, q = k0 >> 2;
                    mbar_wait(bar_aready + 8u * q, (aphase >> q) & 1u);
                    aphase ^= 1u << q;
                    fence_after_sync();
                    WPROF(0);
                    for (uint32_t ks = k0; ks < k1; ks += 2, ++it) {
                        const uint32_t slot = it % NSTAGE, round = it / NSTAGE;
                        mbar_wait(bar_full(slot), round & 1u);
                        fence_after_sync();
                        WPROF(1);
                        const uint32_t sb = ring + slot * C::STAGE_BYTES;
                        if (elect_one()) {
                            mma2_f16_ts(td, ta + 16u * ks + 8u, make_desc_kmajor(sb, lbo, 128), idesc, ks != 0);      // A_lo W_hi
                            mma2_f16_ts(td, ta + 16u * ks, make_desc_kmajor(sb + img, lbo, 128), idesc, 1);           // A_hi W_lo
                            if (ks + 1 < k1) {
                                mma2_f16_ts(td, ta + 16u * ks + 24u, make_desc_kmajor(sb + 2u * img, lbo, 128), idesc, 1);
                                mma2_f16_ts(td, ta + 16u * ks + 16u, make_desc_kmajor(sb + 3u * img, lbo, 128), idesc, 1);
                            }
                            mma2_commit(bar_empty(slot));
                        }
                        __syncwarp();
                        WPROF(2);
                    }
                    {                                                                          // main products A_hi (2^11 W_hi)
                        const uint32_t slot = it % NSTAGE, round = it / NSTAGE;
                        mbar_wait(bar_full(slot), round & 1u);
                        fence_after_sync();
                        WPROF(1);
                        const uint32_t sb = ring + slot * C::STAGE_BYTES;
                        if (elect_one()) {
#pragma unroll
                            for (uint32_t j = 0; j < 4; ++j)
                                if (k0 + j < k1) mma2_f16_ts(td, ta + 16u * (k0 + j), make_desc_kmajor(sb + j * img, lbo, 128), idesc, 1);
                            mma2_commit(bar_empty(slot));
                        }
                        __syncwarp();
                        WPROF(2);
                        ++it;
                    }
                }
                if (elect_one()) mma2_commit(bar_dready);
                __syncwarp();
            } else if (lane == 0) {
                // peer CTA: tell the leader when this CTA's half of a ring stage has landed
                uint32_t nst = 0;
                for (uint32_t k0 = 0; k0 < gm.ksteps; k0 += 4) nst += ((k0 + 4u < gm.ksteps ? 4u : gm.ksteps - k0) + 1u) / 2u + 1u;
                for (uint32_t i = 0; i < nst; ++i, ++it) {
                    const uint32_t slot = it % NSTAGE, round = it / NSTAGE;
                    mbar_wait(bar_full(slot), round & 1u);
                    mbar_arrive_cluster(mapa_u32(bar_full(slot), 0));
                }
            }
            __syncwarp();
        } else {
            WPROF(7);
            pre();
            WPROF(4);
            mbar_wait(bar_dready, g & 1u);
            fence_after_sync();
            WPROF(5);
            epi(tmem + dreg(g) + lane_off);
            WPROF(6);
        }
        ++g;
    };
    auto nopre = []() {};
    // epilogue role: this thread's part of the A operand of GEMM g (region areg(g)) is written
    // quarter q (K steps 4q .. 4q+3 = neurons [64 q, 64 q + 64)) of the A operand of GEMM g is written: one arrival per warp
    auto publish_q = [&](const int q) {
        if (is_epi) {
            tmem_st_wait();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(aready_leader + 8u * q);
        }
    };
    auto publish = [&]() { publish_q(0); };                // K = 16 operands (seeds) have one quarter
    auto epi_sync = [&]() { if (is_epi) asm volatile("bar.sync 1, %0;" ::"n"(NEMPC_WIDE_EPI_WARPS * 32) : "memory"); };
    // K = 16 operand (one K step): hi pairs in columns [0, 8), lo pairs in [8, 16) of region areg(g); written by the quarter-0 warps
    auto put_seed = [&](const float* v16) {
        uint32_t hi[8], lo[8];
        split16(v16, hi, lo);
        const uint32_t a = tmem + areg(g) + lane_off;
        tmem_st8(a, hi);
        tmem_st8(a + 8, lo);
    };
    // this warp's 16-neuron chunk of EVERY 64-neuron quarter (columns 64 q + 16 sub), quarter by quarter, the next chunk's accumulator in flight
    // while `body(q, registers)` runs; `feeds`: the body converted the chunk into operand form -- the quarter is published right away
    auto chunks = [&](const uint32_t dbase, const bool feeds, auto&& body) {
        uint32_t va[16], vb[16];
        const uint32_t c0 = dbase + 16 * sub;
        tmem_ld16_nowait(c0, va);
        tmem_ld_wait(); tmem_ld_pin(va);
        tmem_ld16_nowait(c0 + 64, vb);
        body(0, va);
        if (feeds) publish_q(0);
        tmem_ld_wait(); tmem_ld_pin(vb);
        if constexpr (C::NQ == 4) tmem_ld16_nowait(c0 + 128, va);
        body(1, vb);
        if (feeds) publish_q(1);
        if constexpr (C::NQ == 4) {
            tmem_ld_wait(); tmem_ld_pin(va);
            tmem_ld16_nowait(c0 + 192, vb);
            body(2, va);
            if (feeds) publish_q(2);
            tmem_ld_wait(); tmem_ld_pin(vb);
            body(3, vb);
            if (feeds) publish_q(3);
        }
    };

    // the pair works on super-tiles (2 i, 2 i + 1); both CTAs run the GEMM sequence of the fuller one (the leader's)
    const int S = RK4 ? st.S : 1;                            // integrator stages: 1 (discrete / unity: a = 0, c = 1) or 4 (RK4)
    const long long nsup = (ar.nsteps + NEMPC_WIDE_SUP - 1) / NEMPC_WIDE_SUP;
    for (long long sup0 = 2 * (long long)(blockIdx.x >> 1); sup0 < nsup; sup0 += gridDim.x) {
        const long long step_base = (sup0 + rank) * NEMPC_WIDE_SUP;
        const long long left = ar.nsteps - step_base, left0 = ar.nsteps - sup0 * NEMPC_WIDE_SUP;
        const int nvalid = (int)(left < 0 ? 0 : (left < NEMPC_WIDE_SUP ? left : NEMPC_WIDE_SUP));
        const int nvalid_pair = (int)(left0 < NEMPC_WIDE_SUP ? left0 : NEMPC_WIDE_SUP);
        const long long stepA = step_base + row;
        const bool validA = is_epi && row < nvalid;
        long long bA = 0; int tA = 0;
        if (validA) { bA = stepA / L.H; tA = (int)(stepA - bA * L.H); }

        // ============================ phase A: primal forward at stage s, row = step =============================================
        // z_s = z + a_s E k_{s-1};  h_l = tanh(W_l^T h_{l-1} + b_l) -> scratch (keep_h: phase B needs h_l; else s'(a_l) for phase C);
        // k_s -> sks[s] (RK4);  write_resid: the residual with sum_s c_s k_s
        auto phaseA = [&](const int sg, const bool keep_h, const bool write_resid) {
            if (is_epi && sub == 0) {
                float zr[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) zr[c] = 0.f;
                if (validA) {
                    const TIO* zb = ar.z + bA * (long long)L.n;
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        if (c < D) zr[c] = (float)(c < X ? ((tA == 0) ? ar.x0[bA * X + c] : zb[(tA - 1) * X + c]) : zb[L.H * X + tA * U + (c - X)]);
                    if (RK4 && sg > 0) {
                        const float a_s = st.a[sg];
                        const float* kp = sks + ((long long)(sg - 1) * NEMPC_WIDE_SUP + row) * 16;
                        WIDE_FOR_X(p) zr[p] = fmaf(a_s, kp[p], zr[p]);
                    }
                }
                put_seed(zr);
            }
            publish();
            for (int l = 0; l < nhid; ++l) {
                gemm(l == 0 ? net.in_f : net.hid_f[l - 1], nopre, [&](const uint32_t dbase) {
                    const float* bl = bias + l * HW + 16 * sub;
                    float* dst = sa + ((long long)row * NEMPC_WIDE_MAXHID + l) * HW + 16 * sub;
                    chunks(dbase, true, [&](const int qq, const uint32_t* vr) {
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = tanh_acc(fmaf(__uint_as_float(vr[i]), INV, bl[64 * qq + i]));
                        if (HES && keep_h) st16_global(dst + 64 * qq, v);                     // h_l: phase B needs it
                        else if (JAC) {
                            float s1[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) s1[i] = fmaf(-v[i], v[i], 1.f);
                            st16_global(dst + 64 * qq, s1);                                   // s'(a_l) for phase C
                        }
                        uint32_t hi[8], lo[8];
                        split16(v, hi, lo);
                        tmem_st8(dbase + 64 * qq + 16 * sub, hi);
                        tmem_st8(dbase + 64 * qq + 16 * sub + 8, lo);
                    });
                });
            }
            gemm(net.out_f, nopre, [&](const uint32_t dbase) {
                if (sub != 0) return;
                float v[16];
                tmem_ld16(dbase, v);
                tmem_ld_wait();
                float k[XM];
                WIDE_FOR_X(p) k[p] = fmaf(v[p], INV, bout[p]);
                if (RK4) {
                    float* kp = sks + ((long long)sg * NEMPC_WIDE_SUP + row) * 16;
                    WIDE_FOR_X(p) kp[p] = k[p];
                }
                if (write_resid && validA && ar.resid) {
                    const TIO* zb = ar.z + bA * (long long)L.n;
                    WIDE_FOR_X(p) {
                        float kacc = RK4 ? st.c[sg] * k[p] : k[p];
                        if (RK4)
                            for (int s2 = 0; s2 < sg; ++s2) kacc = fmaf(st.c[s2], sks[((long long)s2 * NEMPC_WIDE_SUP + row) * 16 + p], kacc);
                        const TW xt = (TW)zb[tA * X + p];
                        const TW xp = unity ? (TW)0 : (TW)((tA == 0) ? ar.x0[bA * X + p] : zb[(tA - 1) * X + p]);
                        ar.resid[bA * L.m + tA * X + p] = (TIO)(xp + (TW)kacc - xt);
                    }
                }
            });
        };

        // ============================ phase B: contracted adjoint at stage s, row = step ==========================================
        // seed w_s = lambda (single stage) / c_{S-1} lambda (last RK4 stage) / the weights left in `sw` by the previous call (earlier
        // stages);  want_in: continue through the input layer, J_s^T w_s, and leave  w_{s-1} = c_{s-1} lambda + a_s (J_s^T w_s)[:x]  in `sw`
        auto phaseB = [&](const int sg, const bool want_in) {
            if (is_epi && sub == 0) {
                float lr[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) lr[c] = 0.f;
                // The adjoint is linear in the multipliers and travels through f16-split operands (range 6e-5 .. 6e4 for full precision), while an
                // NLP solver's multipliers range over many decades: every step works with lambda / max|lambda| and its Hessian block is scaled
                // back in the scatter (all RK4 stage weights of a step derive from the same lambda, so one scale per step serves them all)
                float lmax = 0.f;
                if (validA) {
                    WIDE_FOR_X(p) lmax = fmaxf(lmax, fabsf((float)ar.lam[bA * L.m + tA * X + p]));
                }
                const float linv = lmax > 0.f ? 1.f / lmax : 0.f;
                lscale[row] = lmax;
                if (validA) {
                    if (RK4 && sg < S - 1) {
                        WIDE_FOR_X(p) lr[p] = sw[row * 16 + p];
                    } else {
                        const float cs = (RK4 ? st.c[sg] : 1.f) * linv;
                        WIDE_FOR_X(p) lr[p] = cs * (float)ar.lam[bA * L.m + tA * X + p];
                    }
                }
                put_seed(lr);
            }
            publish();
            for (int l = nhid - 1; l >= 0; --l) {
                // accumulator = g_l (adjoint with respect to h_l);  s'(a_l) -> sa (over h_l),  s''(a_l) g_l = -2 h_l s'(a_l) g_l -> sq;
                // next operand u_l = s'(a_l) g_l
                const bool more = l > 0 || want_in;
                // (fetching h_l of the next chunk while the MMAs run / the previous chunk is processed was tried: the 16 extra live registers
                // cost more than the hidden latency returns -- 118 -> 125 ms on C4 at B = 16384)
                gemm(l == nhid - 1 ? net.out_b : net.hid_b[l], nopre, [&](const uint32_t dbase) {
                    float* ph = sa + ((long long)row * NEMPC_WIDE_MAXHID + l) * HW + 16 * sub;
                    float* pq = sq + ((long long)row * NEMPC_WIDE_MAXHID + l) * HW + 16 * sub;
                    chunks(dbase, more, [&](const int qq, const uint32_t* vr) {
                        float h[16], u[16], cf[16];
                        ld16_global_cg(ph + 64 * qq, h);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float hv = h[i], s1 = fmaf(-hv, hv, 1.f);
                            const float gv = __uint_as_float(vr[i]) * INV;
                            u[i] = s1 * gv;
                            cf[i] = NEMPC_WIDE_GRAM_MMA ? -2.f * hv * gv : -2.f * hv * u[i];     // -2 h g (times s' T below), or s''(a) g = -2 h s' g
                            h[i] = s1;
                        }
                        st16_global(ph + 64 * qq, h);
                        st16_global(pq + 64 * qq, cf);
                        if (more) {
                            uint32_t hi[8], lo[8];
                            split16(u, hi, lo);
                            tmem_st8(dbase + 64 * qq + 16 * sub, hi);
                            tmem_st8(dbase + 64 * qq + 16 * sub + 8, lo);
                        }
                    });
                });
            }
            if (RK4 && want_in) {
                gemm(net.in_b, nopre, [&](const uint32_t dbase) {
                    if (sub != 0) return;
                    float v[16];
                    tmem_ld16(dbase, v);
                    tmem_ld_wait();
                    const float lmax = lscale[row];
                    const float a_s = st.a[sg], cprev = st.c[sg - 1] * (lmax > 0.f ? 1.f / lmax : 0.f);
                    WIDE_FOR_X(p)
                        sw[row * 16 + p] = validA ? fmaf(a_s, v[p] * INV, cprev * (float)ar.lam[bA * L.m + tA * X + p]) : 0.f;
                });
            }
        };

        // ============================ phase C: tangent forward (+ curvature) at stage s, row = (step, tangent column) =================
        // seed rows = columns of R_s = I + a_s E dk_{s-1} (identity for a single stage);  store_dk: leave dk_s = J_s R_s and the running
        // sum_s c_s dk_s in the scratch;  write_jac: this is the last stage of the forward sweep;  curv: accumulate
        // R_s^T (sum_p w_{s,p} Hess f_p) R_s;  first_hes: store the Hessian values (with the objective's), later stages add to them
        auto phaseC = [&](const int sg, const bool curv, const bool first_hes, const bool write_jac, const bool store_dk) {
            const int ntile = (nvalid_pair + SPT - 1) / SPT;
            for (int ti = 0; ti < ntile; ++ti) {
                const int sidx = ti * SPT + row / DP, cc = row % DP;          // step inside the super-tile, tangent column
                const int sidx0 = ti * SPT + (32 * wq) / DP;                  // first step of this warp
                const bool validC = is_epi && sidx < nvalid && cc < D;
#if NEMPC_WIDE_GRAM_MMA
                float accM[16], accC[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { accM[i] = 0.f; accC[i] = 0.f; }
#else
                f2 acc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = pk(0.f, 0.f);
#endif
                if (is_epi && sub == 0) {
                    float e[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) e[c] = (c == cc && cc < D) ? 1.f : 0.f;
                    if (RK4 && sg > 0 && validC) {
                        const float a_s = st.a[sg];
                        const float* dp = sdk + ((((long long)(sg - 1) * NEMPC_WIDE_SUP + sidx) * 16 + cc) * 16);
                        WIDE_FOR_X(p) e[p] = fmaf(a_s, dp[p], e[p]);
                    }
                    put_seed(e);
                }
                publish();
                for (int l = 0; l < nhid; ++l) {
                    gemm(l == 0 ? net.in_f : net.hid_f[l - 1],
                         [&]() {
                             // while the MMAs run: this warp's s'(a_l) (x 2^-11: the accumulator's scale) and curvature coefficients (x 2^-22) of
                             // its SPW steps and 64 neurons (chunk `sub` of every quarter), global scratch -> warp-private shared memory
                             for (int i = lane; i < SPW * 4 * C::NQ; i += 32) {
                                 const int sw_ = i / (4 * C::NQ), f4 = i % (4 * C::NQ);
                                 const long long o = ((long long)(sidx0 + sw_) * NEMPC_WIDE_MAXHID + l) * HW + 64 * (f4 >> 2) + 16 * sub + 4 * (f4 & 3);
                                 float4 a = __ldcg(reinterpret_cast<const float4*>(sa + o));
                                 a.x *= INV; a.y *= INV; a.z *= INV; a.w *= INV;
                                 *reinterpret_cast<float4*>(scs + sw_ * 64 + 4 * f4) = a;
                                 if (HES && curv) {
                                     float4 q = __ldcg(reinterpret_cast<const float4*>(sq + o));
                                     constexpr float QS = NEMPC_WIDE_GRAM_MMA ? INV : INV * INV;     // one raw tangent factor (2^11) or two
                                     q.x *= QS; q.y *= QS; q.z *= QS; q.w *= QS;
                                     *reinterpret_cast<float4*>(scs + (SPW + sw_) * 64 + 4 * f4) = q;
                                 }
                             }
                             __syncwarp();
                         },
                         [&](const uint32_t dbase) {
                             const float* s1row = scs + (lane / DP) * 64;
                             chunks(dbase, true, [&](const int qq, const uint32_t* vr) {
#if !NEMPC_WIDE_GRAM_MMA
                                 if (HES && curv) gram_chunk<DP>(stg, vr, scs + SPW * 64, 16 * qq, acc, lane);      // raw tangent T_l (x 2^11)
#endif
                                 f2 v2[8];
#pragma unroll
                                 for (int i4 = 0; i4 < 4; ++i4) {
                                     const float4 s4 = *reinterpret_cast<const float4*>(s1row + 16 * qq + 4 * i4);
                                     v2[2 * i4] = mul2(pk(__uint_as_float(vr[4 * i4]), __uint_as_float(vr[4 * i4 + 1])), pk(s4.x, s4.y));       // V_l = s'(a_l) T_l
                                     v2[2 * i4 + 1] = mul2(pk(__uint_as_float(vr[4 * i4 + 2]), __uint_as_float(vr[4 * i4 + 3])), pk(s4.z, s4.w));
                                 }
                                 uint32_t hi[8], lo[8];
                                 split16p(v2, hi, lo);
                                 tmem_st8(dbase + 64 * qq + 16 * sub, hi);
                                 tmem_st8(dbase + 64 * qq + 16 * sub + 8, lo);
#if NEMPC_WIDE_GRAM_MMA
                                 if (HES && curv) {
                                     // U = (-2 h g) T (true scale: the coefficient row carries 2^-11), split like V; G += U V^T on the tensor path
                                     const float* krow = scs + (SPW + lane / DP) * 64 + 16 * qq;
                                     f2 u2[8];
#pragma unroll
                                     for (int i4 = 0; i4 < 4; ++i4) {
                                         const float4 k4 = *reinterpret_cast<const float4*>(krow + 4 * i4);
                                         u2[2 * i4] = mul2(pk(__uint_as_float(vr[4 * i4]), __uint_as_float(vr[4 * i4 + 1])), pk(k4.x, k4.y));
                                         u2[2 * i4 + 1] = mul2(pk(__uint_as_float(vr[4 * i4 + 2]), __uint_as_float(vr[4 * i4 + 3])), pk(k4.z, k4.w));
                                     }
                                     uint32_t uhi[8], ulo[8];
                                     split16p(u2, uhi, ulo);
                                     gram_mma_chunk(smem_u32(stg), reinterpret_cast<uint4*>(stg), uhi, ulo, hi, lo, accM, accC, lane);
                                 }
#endif
                             });
                             __syncwarp();                                      // `scs` is rewritten for the next layer
                         });
                }
                gemm(net.out_f, nopre, [&](const uint32_t dbase) {
                    if (sub != 0) return;
                    float v[16];
                    tmem_ld16(dbase, v);
                    tmem_ld_wait();
                    if (!validC) return;
                    float jv_[XM];                                             // (sum_s c_s dk_s)[p][cc] up to this stage
                    WIDE_FOR_X(p) jv_[p] = v[p] * INV;
                    if (RK4) {
                        const float c_s = st.c[sg];
                        float* dacc = sdkacc + ((long long)sidx * 16 + cc) * 16;
                        if (store_dk) {
                            float* dp = sdk + ((((long long)sg * NEMPC_WIDE_SUP + sidx) * 16 + cc) * 16);
                            WIDE_FOR_X(p) dp[p] = jv_[p];
                        }
                        if (store_dk || write_jac) {
                            WIDE_FOR_X(p) {
                                jv_[p] = sg == 0 ? c_s * jv_[p] : fmaf(c_s, jv_[p], dacc[p]);
                                if (store_dk) dacc[p] = jv_[p];
                            }
                        }
                    }
                    if (write_jac && ar.jac) {
                        const long long step = step_base + sidx;
                        const long long b = step / L.H;
                        const int t = (int)(step - b * L.H);
                        TIO* jv = ar.jac + b * L.nnz_jac;
                        WIDE_FOR_X(p) {
                            const TW val = (TW)jv_[p] + ((!unity && cc == p) ? (TW)1 : (TW)0);
                            if (cc < X) { if (t > 0) jv[jac_slot_A(L, t, p, cc)] = (TIO)val; }
                            else jv[jac_slot_B(L, t, p, cc - X)] = (TIO)val;
                            if (cc == 0) jv[jac_slot_minus1(L, t, p)] = (TIO)-1;
                        }
                    }
                });
                if (HES && curv) {
                    // ---- sum the four neuron quarters, scatter the lower triangle (same slots as nempc_generic.cuh) -----------------
                    epi_sync();                                    // `part` aliases the staging planes: every warp is done with its curvature
#if NEMPC_WIDE_GRAM_MMA
                    if (is_epi) {
                        // accumulator fragments -> [quarter][step][DP][DP]: entry (R, Cc) of a 16-row tile is kept when row and column belong to the same step
#pragma unroll
                        for (int tile = 0; tile < 2; ++tile)
#pragma unroll
                            for (int h = 0; h < 2; ++h)
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int R = 16 * tile + (lane >> 2) + 8 * (e >> 1), Cc = 16 * tile + 8 * h + 2 * (lane & 3) + (e & 1);
                                    if (R / DP == Cc / DP) {
                                        const int s8 = (32 * wq) / DP + R / DP;
                                        const int ai = (tile * 2 + h) * 4 + e;
                                        part[((sub * SPT + s8) * DP + R % DP) * DP + Cc % DP] = fmaf(accC[ai], INV, accM[ai]);
                                    }
                                }
                    }
#else
                    if (is_epi) {
                        constexpr int NB = DP / 4, LPS = NB * NB, ACTIVE = (32 / DP) * LPS;
                        if (lane < ACTIVE) {
                            const int sw_ = lane / LPS, bl = lane % LPS, bi = bl / NB, bj = bl % NB;
                            const int s8 = (32 * wq) / DP + sw_;
                            float* pp = part + ((sub * SPT + s8) * DP + 4 * bi) * DP + 4 * bj;
#pragma unroll
                            for (int i = 0; i < 4; ++i)
#pragma unroll
                                for (int k = 0; k < 4; ++k) pp[i * DP + k] = f2lo(acc[i * 4 + k]) + f2hi(acc[i * 4 + k]);
                        }
                    }
#endif
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
                            TW val = (TW)sum * (TW)lscale[sx];                      // back to the step's multiplier scale
                            int slot;
                            if (a < X) {
                                slot = hes_slot_xx(L, t, a, c);
                                if (first_hes && a == c && ar.quad) val += sig * (TW)2 * (TW)ar.quad[(t - 1) * X + a];
                            } else if (c < X) {
                                slot = hes_slot_ux(L, t, a - X, c);
                            } else {
                                slot = hes_slot_uu(L, t, a - X, c - X);
                                if (first_hes && a == c && ar.quad) val += sig * (TW)2 * (TW)ar.quad[L.H * X + t * U + (a - X)];
                            }
                            if (first_hes) hv[slot] = (TIO)val;
                            else hv[slot] = (TIO)((TW)hv[slot] + val);             // later RK4 stages add their R_s^T M_s R_s (same thread every time)
                        }
                        if (first_hes) {
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
                    }
                    epi_sync();                                    // `part` is rewritten by the next tile
                }
            }
        };

        // ============================ the stage schedule ==========================================================================
        // single stage:  A, B, C.   RK4:  forward sweep  s = 0..2: A, C (tangents only: k_s, dk_s, R_{s+1});  last stage:
        // A, B (w_3 = c_3 lambda, known up front), C with curvature (also the Jacobian);  backward sweep  s = 2..0: A (recomputed), B with
        // w_s = c_s lambda + a_{s+1} J_{s+1,x}^T w_{s+1}, C with curvature:  H = sum_s R_s^T (sum_p w_{s,p} Hess f_p(z_s)) R_s
        // (one call site per phase, so that the three bodies stay inlined)
        // (for a single stage everything below is a compile-time constant and the schedule collapses to A, B, C)
        const int npass = RK4 ? ((HES && S > 1) ? 2 * S - 1 : S) : 1;
        for (int pass = 0; pass < npass; ++pass) {
            const int sg = RK4 ? (pass < S ? pass : 2 * S - 2 - pass) : 0;
            const bool last = RK4 ? pass == S - 1 : true;
            const bool second = RK4 ? pass >= S - 1 : true;              // second-order passes: the last stage and the backward sweep
            phaseA(sg, second, last);
            if (HES && second) phaseB(sg, RK4 && sg > 0);
            epi_sync();                                    // the scratch of this stage is complete (bar.sync orders it at CTA scope)
            if (JAC) phaseC(sg, second, last, last, !second);
            epi_sync();                                    // the scratch is rewritten by the next stage / super-tile
        }
    }

#undef WIDE_FOR_X
    WPROF_FLUSH;
    fence_before_sync();
    __syncthreads();
    cluster_sync_all();                                     // nobody leaves (or frees tensor memory) while the peer may still use this CTA
    if (is_mma) tmem_dealloc2(tmem, 512);
}
#endif  // __CUDACC__
