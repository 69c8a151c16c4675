// libnempc: CUDA kernels (sm_100a) + C ABI (include/nempc.h) of the pyNeuralEMPC NLP-evaluation hot path.
// No CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/nempc.h"
#include "nempc_fast.cuh"
#include "nempc_fast64.cuh"
#include "nempc_generic.cuh"
#include "nempc_layout.h"
#include "nempc_small.cuh"
#include "nempc_solver.cuh"
#include "nempc_tc.cuh"
#include "nempc_wide.cuh"
#include "nempc_rolling.cuh"

// ======================================================================================================
// kernels
// ======================================================================================================
extern __shared__ __align__(16) unsigned char nempc_smem[];

template <typename T, typename TIO, int DMAX>
__global__ void __launch_bounds__(256)
nempc_generic_kernel(const NetView<T> net, const StageTable<T> st, const NlpLayout L, const SlotLayout sl,
                     const EvalArgs<TIO> ar, const int tps, T* __restrict__ gws) {
    const int slots = blockDim.x / tps;
    const int slot = threadIdx.x / tps;
    const int lt = threadIdx.x - slot * tps;
    T* ws = gws ? gws + ((long long)blockIdx.x * slots + slot) * sl.total
                : reinterpret_cast<T*>(nempc_smem) + (long long)slot * sl.total;
    const int bar_id = 1 + slot;
    for (long long tile = blockIdx.x; tile * slots < ar.nsteps; tile += gridDim.x) {
        const long long step = tile * slots + slot;
        if (step < ar.nsteps) {
            generic_step<T, TIO, DMAX>(net, st, L, sl, ar, step, ws, lt, tps, bar_id);
            slot_barrier(bar_id, tps);      // workspace is reused by the next step of this slot
        }
    }
}

#ifndef NEMPC_FAST_THREADS
#define NEMPC_FAST_THREADS 128
#endif
#ifndef NEMPC_FAST_NCHUNK30        // register chunks of the 30-wide second layer: 3 chunks x 10 neurons = 5 packed pairs each
#define NEMPC_FAST_NCHUNK30 3      // (sweep profiles/r1d: 3 chunks @ 5 CTAs/SM 0.299 ms, 2 @ 4 0.307, 2 @ 3 0.475, 1 @ 2 0.551)
#endif
#ifndef NEMPC_FAST_MINBLOCKS       // CTAs/SM the register allocator must allow: occupancy is the main lever of this kernel
#define NEMPC_FAST_MINBLOCKS(JC) ((JC) <= 10 ? 5 : 4)
#endif

template <int X, int U, int H1, int H2, int NCHUNK, int MODE, typename TIO>
__global__ void __launch_bounds__(NEMPC_FAST_THREADS, NEMPC_FAST_MINBLOCKS(H2 / NCHUNK))
nempc_fast_kernel(const __grid_constant__ FastWeights<X, U, H1, H2, NCHUNK> w, const StageTable<float> st,
                  const NlpLayout L, const EvalArgs<TIO> ar) {
    // cold per-thread state (layer-1 activations, per-output Hessian accumulators): [element][thread], bank = thread
    float* scr = reinterpret_cast<float*>(nempc_smem) + threadIdx.x;
    for (long long step = (long long)blockIdx.x * blockDim.x + threadIdx.x; step < ar.nsteps;
         step += (long long)gridDim.x * blockDim.x) {
        if (ar.gate && ar.gate[step / L.H] != ar.gate_value) continue;          // batched solver: problems that are done are not re-evaluated
        fast_step<X, U, H1, H2, NCHUNK, MODE, TIO>(w, st, L, ar, step, scr, NEMPC_FAST_THREADS);   // compile-time scratch stride (the launch uses exactly this block size): LDS / STS with immediate offsets, 2.8 % fewer instructions
    }
}

// objective value + gradient of f(z) = sum lin_i z_i + quad_i (z_i - ref_i)^2 : one warp per problem,
// fixed summation order (deterministic).  Replaces JAXObjectifFunc.forward / .gradient (objective/jax.py:28-41).
template <typename TIO>
__global__ void nempc_objective_kernel(const TIO* __restrict__ z, const double* __restrict__ lin,
                                       const double* __restrict__ quad, const double* __restrict__ ref,
                                       TIO* __restrict__ obj, TIO* __restrict__ grad, int n, long long B,
                                       const int* __restrict__ gate = nullptr, int gate_value = 0) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long b = warp; b < B; b += nwarps) {
        if (gate && gate[b] != gate_value) continue;
        double acc = 0.0;
        for (int i = lane; i < n; i += 32) {
            const double zi = (double)z[b * n + i];
            const double dz = zi - ref[i];
            acc += lin[i] * zi + quad[i] * dz * dz;
            if (grad) grad[b * n + i] = (TIO)(lin[i] + 2.0 * quad[i] * dz);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (obj && lane == 0) obj[b] = (TIO)acc;
    }
}

// objective value + gradient of a general quadratic cost f(z) = 1/2 z' P z + q' z + c with a sparse symmetric P (CSR over all n rows,
// both triangles): rate penalties (u_{t+1} - u_t)' S (u_{t+1} - u_t), full-matrix stage / terminal weights, cross terms.  One warp per
// problem, fixed summation order.  Replaces JAXObjectifFunc.forward / .gradient (objective/jax.py:28-41) for non-separable costs.
template <typename TIO>
__global__ void nempc_quadform_kernel(const TIO* __restrict__ z, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                      const double* __restrict__ val, const double* __restrict__ q, double c,
                                      TIO* __restrict__ obj, TIO* __restrict__ grad, int n, long long B) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long b = warp; b < B; b += nwarps) {
        const TIO* zb = z + b * n;
        double acc = 0.0;
        for (int i = lane; i < n; i += 32) {
            double pz = 0.0;
            for (int k = ptr[i]; k < ptr[i + 1]; ++k) pz += val[k] * (double)zb[idx[k]];
            const double zi = (double)zb[i];
            acc += zi * (0.5 * pz + q[i]);
            if (grad) grad[b * n + i] = (TIO)(pz + q[i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (obj && lane == 0) obj[b] = (TIO)(acc + c);
    }
}

// Lagrangian-Hessian values on the UNION pattern of the constraint Hessian (what the evaluation kernels write, in their closed-form slot
// order) and a constant objective Hessian P:  out[b, s] = kern[b, src[s]] (if src[s] >= 0) + obj_factor_b * pval[s]   (ipopt.py:66-86)
template <typename TIO>
__global__ void nempc_hessian_merge_kernel(const TIO* __restrict__ kern, const int32_t* __restrict__ src, const double* __restrict__ pval,
                                           const TIO* __restrict__ sigma, double sigma_scalar, TIO* __restrict__ out,
                                           long long nnz_kern, long long nnz_out, long long B) {
    const long long total = B * nnz_out;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / nnz_out, s = i - b * nnz_out;
        const int k = src[s];
        const double sg = sigma ? (double)sigma[b] : sigma_scalar;
        out[i] = (TIO)((k >= 0 ? (double)kern[b * nnz_kern + k] : 0.0) + sg * pval[s]);
    }
}

// register-resident FMA loop: sustained FP32 / FP64 FMA-pipe throughput (the compute-roofline denominator)
template <typename T>
__global__ void nempc_fma_peak_kernel(T* out, int iters, T seed) {
    T a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + (T)(threadIdx.x + i);
    const T m = (T)0.999, c = (T)1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = a[i] * m + c;
    }
    T s = (T)0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == (T)123456789) out[0] = s;      // never true; keeps the loop alive
}

// packed variant (fma.rn.f32x2 -> FFMA2): two FMAs per lane per issued instruction
__global__ void nempc_fma2_peak_kernel(float* out, int iters, float seed) {
    f2 a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = pk(seed + (float)(threadIdx.x + i), seed - (float)i);
    const f2 m = pk(0.999f, 0.998f), c = pk(1e-3f, 2e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma2(a[i], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f2lo(a[i]) + f2hi(a[i]);
    if (s == 123456789.f) out[0] = s;
}

// ---- interior-point solver kernels: one thread per problem ------------------------------------------------------------
__global__ void nempc_ipm_init_kernel(const NlpLayout L, const SolverWs w, const SolverOpts o, long long B, int has_init) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) ipm_init_problem(L, w, b, o, has_init != 0);
}
template <int XM, int UM, int XC, int UC>
__global__ void nempc_ipm_kkt_kernel(const NlpLayout L, const SolverWs w, const SolverOpts o, long long B, int* counts) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    ipm_kkt_problem<XM, UM, XC, UC>(L, w, b, o);
    if (w.status[b] == NEMPC_ST_RUNNING) { atomicAdd(&counts[0], 1); if (!w.accepted[b]) atomicAdd(&counts[1], 1); }
}
// Staged variant: ONE WARP PER PROBLEM.  The Riccati sweeps are a long chain of dependent float64 operations per problem;
// with one thread per problem reading its rows straight from global memory (row stride n doubles between threads, nothing
// coalesces) and doing ~2000 divisions and ~100 logarithms inside that chain, the kernel took 1.08 ms for 4096 problems
// (profiles/r1k).  Here the warp copies the problem's rows (iterate, duals, gradient, residual, Jacobian / Hessian values)
// into shared memory with coalesced loads and runs ipm_kkt_warp: per-variable work on 32 lanes, the Riccati recursion on
// lane 0, bit-identical results.  Bounds are staged once per CTA.
template <int XM, int UM, int XC, int UC>
__global__ void __launch_bounds__(512)
nempc_ipm_kkt_staged_kernel(const NlpLayout L, const SolverWs w, const SolverOpts o, long long B, int* counts) {
    extern __shared__ __align__(16) double kkt_sm[];
    const int n = L.n, m = L.m, nj = (int)L.nnz_jac, nh = (int)L.nnz_hes, nK = L.H * L.u * L.x, nk = L.H * L.u;
    // staged per problem: what the sequential Riccati lane reads and writes (residual, Jacobian / Hessian values, step, gains, the
    // Sigma / barrier-gradient rows).  The iterate, its bound duals, the gradient and the multipliers are only touched by the
    // lane-parallel phases (coalesced) and stay in global memory: less shared memory per problem = more problems in flight per SM.
    const int per = 4 * n + 2 * m + nj + nh + nK + nk;
    const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* slb = kkt_sm; double* sub = kkt_sm + n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { slb[i] = w.lb[i]; sub[i] = w.ub[i]; }
    const long long b = (long long)blockIdx.x * wpb + warp;
    double* p = kkt_sm + 2 * n + (size_t)warp * per;
    double* sres = p; p += m; double* sjac = p; p += nj; double* shes = p; p += nh;
    double* sdz = p; p += n; double* slamn = p; p += m; double* sK = p; p += nK; double* skf = p; p += nk;
    double* s1 = p; p += n; double* s2 = p; p += n; double* s3 = p;
    const bool live = b < B && w.status[b < B ? b : 0] == NEMPC_ST_RUNNING;
    if (live) {
        for (int i = lane; i < m; i += 32) sres[i] = w.resid[b * m + i];
        for (int i = lane; i < nj; i += 32) sjac[i] = w.jac[b * L.nnz_jac + i];
        for (int i = lane; i < nh; i += 32) shes[i] = w.hes[b * L.nnz_hes + i];
    } else if (b < B && lane == 0) w.accepted[b] = 1;
    __syncthreads();
    if (!live) return;
    SolverWs w2 = w;
    w2.lb = slb; w2.ub = sub;
    KktRows r;
    r.z = w.z + b * n; r.lam = w.lam + b * m; r.zL = w.zL + b * n; r.zU = w.zU + b * n; r.gr = w.grad + b * n;
    r.c = sres; r.jv = sjac; r.hv = shes;
    r.dz = sdz; r.lamn = slamn; r.Kb = sK; r.kfb = skf;
    r.dzL = w.dzL + b * n; r.dzU = w.dzU + b * n; r.zt = w.zt + b * n;
    r.s1 = s1; r.s2 = s2; r.s3 = s3;
    ipm_kkt_warp<XM, UM, XC, UC>(L, w2, b, o, r, lane);
    __syncwarp();
    if (lane == 0 && w.status[b] == NEMPC_ST_RUNNING) { atomicAdd(&counts[0], 1); if (!w.accepted[b]) atomicAdd(&counts[1], 1); }
    for (int i = lane; i < n; i += 32) w.dz[b * n + i] = sdz[i];
    for (int i = lane; i < m; i += 32) w.lamn[b * m + i] = slamn[i];
    for (int i = lane; i < nK; i += 32) w.K[b * (long long)nK + i] = sK[i];
    for (int i = lane; i < nk; i += 32) w.kf[b * (long long)nk + i] = skf[i];
}

// iterate update, one thread per (problem, variable): same arithmetic as ipm_update_problem, coalesced
// `counts` (device-side iteration loop, see nempc_solve): counts[0] == 0 means every problem converged or failed in this iteration's KKT
// kernel -- the host loop leaves BEFORE the update, the captured loop runs the update kernels as no-ops
__global__ void nempc_ipm_update_flat_kernel(const NlpLayout L, const SolverWs w, long long B, const int* counts) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = L.n, m = L.m;
    if (idx >= B * n || (counts && counts[0] == 0)) return;
    const long long b = idx / n;
    const int i = (int)(idx - b * n);
    if (w.status[b] != NEMPC_ST_RUNNING) return;
    const double a = w.alpha[b], aD = w.alphaD[b], mu = w.mu[b];
    const double ks = 1e10;
    const double zi = w.z[idx] + a * w.dz[idx];
    w.z[idx] = zi;
    if (nempc_finite(w.lb[i])) { const double d = zi - w.lb[i]; w.zL[idx] = fmin(fmax(w.zL[idx] + aD * w.dzL[idx], mu / (ks * d)), ks * mu / d); }
    if (nempc_finite(w.ub[i])) { const double d = w.ub[i] - zi; w.zU[idx] = fmin(fmax(w.zU[idx] + aD * w.dzU[idx], mu / (ks * d)), ks * mu / d); }
    if (i < m) { const long long j = b * m + i; w.lam[j] += a * (w.lamn[j] - w.lam[j]); }
}
// stats[NEMPC_SV_LSX] counts the steps that were taken although no line-search trial passed the Armijo test (the step of the last halving
// is applied as in the numpy statement; nempc_solve_stats reports how often that happened)
__global__ void nempc_ipm_count_iter_kernel(const SolverWs w, long long B, const int* counts, int* stats) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (counts && counts[0] == 0) return;
    if (b < B && w.status[b] == NEMPC_ST_RUNNING) {
        w.iters[b] += 1;
        if (!w.accepted[b]) atomicAdd(stats + NEMPC_SV_LSX, 1);
    }
}
// ---- device-side iteration loop: the conditions of the two host loops of nempc_solve, evaluated by one thread between the kernels of a
// CUDA graph whose WHILE nodes they drive (cudaGraphSetConditional) ------------------------------------------------------------------------
__global__ void nempc_ipm_ls_begin_kernel(cudaGraphConditionalHandle inner, int* stats, int max_backtrack) {
    stats[NEMPC_SV_TRIAL] = 0;
    cudaGraphSetConditional(inner, max_backtrack > 0 ? 1u : 0u);
}
__global__ void nempc_ipm_ls_cond_kernel(cudaGraphConditionalHandle inner, int* stats, int max_backtrack) {
    const int t = ++stats[NEMPC_SV_TRIAL];
    stats[NEMPC_SV_TRIALS_TOTAL] += 1;
    cudaGraphSetConditional(inner, (t < max_backtrack && stats[0] > 0 && stats[1] > 0) ? 1u : 0u);
}
__global__ void nempc_ipm_outer_cond_kernel(cudaGraphConditionalHandle outer, int* stats, int max_iter) {
    unsigned go = 0;
    if (stats[0] > 0) { const int it = ++stats[NEMPC_SV_IT]; go = it < max_iter ? 1u : 0u; }      // stats[0] == 0: the host loop's `break`
    cudaGraphSetConditional(outer, go);
}

// line search, ONE WARP PER PROBLEM: |c|_1 and the barrier terms (a logarithm per bounded variable) are computed on 32 lanes into
// shared memory and added by lane 0 in the order of ipm_linesearch_problem / ipm_barrier, so the Armijo test sees the same bits.
__global__ void __launch_bounds__(256)
nempc_ipm_linesearch_warp_kernel(const NlpLayout L, const SolverWs w, const SolverOpts o, long long B, int* counts) {
    extern __shared__ __align__(16) double ls_sm[];
    const int n = L.n, m = L.m, wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * wpb + warp;
    if (b >= B || w.status[b] != NEMPC_ST_RUNNING || w.accepted[b]) return;          // whole warps leave together
    double* tl = ls_sm + (size_t)warp * (2 * n + m); double* tu = tl + n; double* tc = tu + n;
    const double* zt = w.zt + b * n; const double* ct = w.residt + b * m;
    for (int i = lane; i < m; i += 32) tc[i] = fabs(ct[i]);
    for (int i = lane; i < n; i += 32) {
        double a = 0.0, c = 0.0;
        if (nempc_finite(w.lb[i])) { const double d = zt[i] - w.lb[i]; a = log(d > 1e-300 ? d : 1e-300); }
        if (nempc_finite(w.ub[i])) { const double d = w.ub[i] - zt[i]; c = log(d > 1e-300 ? d : 1e-300); }
        tl[i] = a; tu[i] = c;
    }
    __syncwarp();
    int accept = 0;
    double an = 0.0;
    if (lane == 0) {
        double c1 = 0.0, bar = 0.0;
        for (int i = 0; i < m; ++i) c1 += tc[i];
        for (int i = 0; i < n; ++i) { bar += tl[i]; bar += tu[i]; }          // terms of infinite bounds are +0.0: adding them changes no bit
        const double phi = w.objt[b] - w.mu[b] * bar + w.nu[b] * c1;
        const double a = w.alpha[b];
        if (nempc_finite(phi) && phi <= w.phi0[b] + o.eta * a * fmin(w.dphi[b], 0.0)) { w.accepted[b] = 1; accept = 1; }
        else { an = 0.5 * a; w.alpha[b] = an; atomicAdd(&counts[1], 1); }
    }
    accept = __shfl_sync(0xffffffffu, accept, 0);
    if (accept) return;
    an = __shfl_sync(0xffffffffu, an, 0);
    const double* z = w.z + b * n; const double* dz = w.dz + b * n; double* ztw = w.zt + b * n;
    for (int i = lane; i < n; i += 32) ztw[i] = z[i] + an * dz[i];
}

__global__ void nempc_ipm_linesearch_kernel(const NlpLayout L, const SolverWs w, const SolverOpts o, long long B, int* counts) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    ipm_linesearch_problem(L, w, b, o);
    if (w.status[b] == NEMPC_ST_RUNNING && !w.accepted[b]) atomicAdd(&counts[1], 1);
}
__global__ void nempc_ipm_finish_kernel(const SolverWs w, long long B, int* status, int* iters, double* err) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    status[b] = w.status[b] == NEMPC_ST_RUNNING ? NEMPC_ST_MAXITER : w.status[b];
    if (iters) iters[b] = w.iters[b];
    if (err) err[b] = w.err[b];
}

// ======================================================================================================
// handle
// ======================================================================================================
static thread_local std::string g_err;

struct FastShape { int x, u, h1, h2; };

struct nempc_handle {
    nempc_desc desc{};
    int d = 0, L = 0;
    std::vector<int> dims;                       // d, widths...
    std::vector<std::vector<double>> W, bvec;
    std::vector<bool> wset;
    std::vector<double> lin, quad, ref;
    bool has_objective = false;
    NlpLayout lay{};
    // device state
    void* dW[NEMPC_MAXL] = {}; void* dWT[NEMPC_MAXL] = {}; void* db[NEMPC_MAXL] = {};
    double *dlin = nullptr, *dquad = nullptr, *dref = nullptr;
    int use_fast = 0; int fast_id = -1;
    int fast64_id = -1; std::vector<unsigned char> fast64w;      // float64 register-resident kernel (nempc_fast64.cuh): Fast64Weights<...> blob
    int use_tc = 0; int tc_id = -1; void* d_tcimg = nullptr; float* d_tccb = nullptr; float* d_tcwx = nullptr;   // tensor-core kernel: f16 weight images, f32 constants, first-layer rows of the exogenous inputs
    int dmma_id = -1; double* dmma_scratch = nullptr; size_t dmma_scratch_doubles = 0;      // float64 DMMA path: network outputs + stage state of one chunk of steps
    int use_wide = 0; int wide_hes = 0; int wide_id = -1; unsigned char* d_wblob = nullptr; float* d_wcb = nullptr; WideNet wnet{};    // width-256 tensor-core kernel: streamed operand images, biases
    float* wide_scratch = nullptr; size_t wide_scratch_bytes = 0;
    std::vector<unsigned char> fastw;            // FastWeights<...> blob
    SlotLayout sl{};
    int tps = 32, slots = 1, dmax = 4;
    size_t smem_bytes = 0; bool global_ws = false; void* gws = nullptr; size_t gws_bytes = 0;
    int sm_count = 0; int max_smem_optin = 0;
    cudaStream_t stream = nullptr;               // own stream for *_host calls
    cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};   // chunk pipeline of nempc_eval_host (H2D | kernels | D2H overlap)
    // staging for eval_host
    void* st_buf[10] = {}; size_t st_cap[10] = {};
    long long launches = 0;
    const int* gate = nullptr; int gate_value = 0;   // set by nempc_solve around its internal evaluations (EvalArgs::gate)
    // exogenous model inputs (nempc_set_exogenous): device copies, rows held, and the problem offset of the chunk being issued
    int n_ext = 0; double *d_tvp = nullptr, *d_p = nullptr; size_t tvp_cap = 0, p_cap = 0;
    long long tvp_rows = 0, p_rows = 0, exo_base = 0;
    // nempc_eval_host replay: the chunk pipeline of the last argument set, captured as a CUDA graph (one submission per call)
    struct HostKey { int64_t B; const void* in[4]; void* out[5]; double sigma; bool operator==(const HostKey& o) const { return memcmp(this, &o, sizeof(HostKey)) == 0; } };
    HostKey hk{}; int hk_seen = 0; cudaGraphExec_t hk_exec = nullptr; long long hk_launches = 0; bool hk_disabled = false;
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_kern[3] = {nullptr, nullptr, nullptr};   // chunk pipeline: kernels of chunk c+1 wait for the kernels of chunk c when the evaluation kernels share a per-handle scratch
    // solver workspace
    void* sv_buf = nullptr; size_t sv_cap = 0; double *sv_lb = nullptr, *sv_ub = nullptr; int* sv_counts = nullptr; int* sv_counts_host = nullptr;
    // nempc_solve's iteration loop as a CUDA graph with two nested WHILE nodes, replayed while (batch, options, workspace) stay the same
    struct SolveKey { int64_t B; const void* buf; int32_t max_iter, max_backtrack; double od[11]; bool operator==(const SolveKey& o) const { return memcmp(this, &o, sizeof(SolveKey)) == 0; } };
    SolveKey sk{}; int sk_seen = 0; cudaGraphExec_t sk_exec = nullptr; bool sk_disabled = false; long long sk_launches_kkt = 3, sk_launches_trial = 3;
    int sv_last_graph = 0; long long sv_last_lsx = 0, sv_last_trials = 0;
    std::string err, kname;
};

#define SET_ERR(h, ...)                                   \
    do {                                                  \
        char _b[512]; snprintf(_b, sizeof _b, __VA_ARGS__); \
        if (h) (h)->err = _b; g_err = _b;                 \
    } while (0)
#define CU(h, call)                                                                   \
    do {                                                                              \
        cudaError_t _e = (call);                                                      \
        if (_e != cudaSuccess) {                                                      \
            SET_ERR(h, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return NEMPC_ECUDA;                                                       \
        }                                                                             \
    } while (0)

static size_t dsize(int dt) { return dt == NEMPC_F64 ? 8 : 4; }

static const FastShape kFastShapes[] = {{2, 1, 30, 30}, {2, 1, 32, 32}, {2, 1, 16, 16}};
static const int kNumFastShapes = sizeof(kFastShapes) / sizeof(kFastShapes[0]);

static int fast_shape_id(const nempc_desc& d) {
    if (d.compute_dtype != NEMPC_F32 || d.activation != NEMPC_ACT_TANH || d.n_layers != 3 || d.tvp_dim + d.p_dim > 0) return -1;
    for (int i = 0; i < kNumFastShapes; ++i)
        if (d.x_dim == kFastShapes[i].x && d.u_dim == kFastShapes[i].u && d.widths[0] == kFastShapes[i].h1 &&
            d.widths[1] == kFastShapes[i].h2)
            return i;
    return -1;
}

// float64 arithmetic on the same small networks: nempc_fast64_kernel (float64 I/O only)
static int fast64_shape_id(const nempc_desc& d) {
    if (d.compute_dtype != NEMPC_F64 || d.io_dtype != NEMPC_F64 || d.activation != NEMPC_ACT_TANH || d.n_layers != 3 || d.tvp_dim + d.p_dim > 0) return -1;
    for (int i = 0; i < kNumFastShapes; ++i)
        if (d.x_dim == kFastShapes[i].x && d.u_dim == kFastShapes[i].u && d.widths[0] == kFastShapes[i].h1 &&
            d.widths[1] == kFastShapes[i].h2)
            return i;
    return -1;
}

// tensor-core kernel instantiations: (x, u, hidden layers, width), every hidden layer `hw` wide
struct TcShape { int x, u, nhid, hw; };
static const TcShape kTcShapes[] = {{4, 1, 3, 128}, {4, 1, 2, 128}, {2, 1, 3, 128}, {2, 1, 2, 128}, {3, 1, 3, 128}, {3, 1, 2, 128}, {4, 2, 3, 128}, {4, 2, 2, 128},
                                    {4, 1, 3, 64}, {4, 1, 2, 64}, {2, 1, 3, 64}, {2, 1, 2, 64}, {2, 1, 3, 32}, {4, 1, 3, 32}};
static const int kNumTcShapes = sizeof(kTcShapes) / sizeof(kTcShapes[0]);

static int tc_shape_id(const nempc_desc& d) {
    if (d.compute_dtype != NEMPC_F32 || d.activation != NEMPC_ACT_TANH) return -1;
    for (int l = 0; l + 1 < d.n_layers; ++l) if (d.widths[l] != d.widths[0]) return -1;
    for (int i = 0; i < kNumTcShapes; ++i)
        if (d.x_dim == kTcShapes[i].x && d.u_dim == kTcShapes[i].u && d.n_layers - 1 == kTcShapes[i].nhid && d.widths[0] == kTcShapes[i].hw) return i;
    return -1;
}

// width-256 tensor-core kernel (nempc_wide.cuh): (x, u) instantiations; 2..4 hidden layers, all 256 wide; discrete / unity / RK4
struct WideShape { int x, u, hw; };
static const WideShape kWideShapes[] = {{12, 4, 256}, {4, 1, 256}, {2, 1, 256}, {6, 2, 256}, {4, 1, 128}, {2, 1, 128}};
static const int kNumWideShapes = sizeof(kWideShapes) / sizeof(kWideShapes[0]);

static int wide_shape_id(const nempc_desc& d) {
    if (d.compute_dtype != NEMPC_F32 || d.activation != NEMPC_ACT_TANH || d.tvp_dim + d.p_dim > 0) return -1;
    if (d.n_layers - 1 < 2 || d.n_layers - 1 > NEMPC_WIDE_MAXHID) return -1;
    for (int l = 0; l + 1 < d.n_layers; ++l) if (d.widths[l] != d.widths[0]) return -1;
    if (d.widths[0] != 256 && d.widths[0] != 128) return -1;
    static const bool force_rt = getenv("NEMPC_WIDE_RUNTIME_SHAPES") && atoi(getenv("NEMPC_WIDE_RUNTIME_SHAPES")) != 0;      // experiments
    for (int i = 0; i < kNumWideShapes && !force_rt; ++i)
        if (d.x_dim == kWideShapes[i].x && d.u_dim == kWideShapes[i].u && d.widths[0] == kWideShapes[i].hw) return i;
    // any other shape with x_dim + u_dim <= 16: instantiations that read the dimensions at run time (nempc_wide_tu.cu)
    const int dd = d.x_dim + d.u_dim;
    if (dd > 16) return -1;
    return 100 + (d.widths[0] == 128 ? 3 : 0) + (dd <= 4 ? 0 : (dd <= 8 ? 1 : 2));
}

// float64 tensor-core (DMMA) path (nempc_dmma.cuh): (x, u) instantiations of the stage kernel; 2..4 tanh hidden layers, all 128 or 64 wide
struct DmmaShape { int x, u; };
static const DmmaShape kDmmaShapes[] = {{2, 1}, {3, 1}, {4, 1}, {4, 2}, {6, 2}};
static const int kNumDmmaShapes = sizeof(kDmmaShapes) / sizeof(kDmmaShapes[0]);
static int dmma_shape_id(const nempc_desc& d) {
    if (d.compute_dtype != NEMPC_F64 || d.io_dtype != NEMPC_F64 || d.activation != NEMPC_ACT_TANH || d.tvp_dim + d.p_dim > 0) return -1;
    if (d.n_layers - 1 < 2 || d.n_layers - 1 > 4) return -1;
    for (int l = 0; l + 1 < d.n_layers; ++l) if (d.widths[l] != d.widths[0]) return -1;
    if (d.widths[0] != 128 && d.widths[0] != 64) return -1;
    for (int i = 0; i < kNumDmmaShapes; ++i)
        if (d.x_dim == kDmmaShapes[i].x && d.u_dim == kDmmaShapes[i].u) return i;
    return -1;
}

extern "C" const char* nempc_version(void) { return "nempc 0.2 (sm_100a)"; }
#ifndef NEMPC_SOURCE_HASH
#define NEMPC_SOURCE_HASH "unknown"
#endif
// sha256 (first 32 hex digits) of the sources this binary was built from (pyneuralempc_b200/build.py source_hash)
extern "C" const char* nempc_source_hash(void) { return NEMPC_SOURCE_HASH; }
// ABI generation (bumped whenever a signature or struct layout of include/nempc.h changes) and the struct sizes this binary was built with:
// a binding checks them after dlopen instead of misreading a stale library
extern "C" int32_t nempc_abi_info(int32_t* desc_bytes, int32_t* solver_opts_bytes) {
    if (desc_bytes) *desc_bytes = (int32_t)sizeof(nempc_desc);
    if (solver_opts_bytes) *solver_opts_bytes = (int32_t)sizeof(nempc_solver_opts);
    return NEMPC_ABI_VERSION;
}
extern "C" const char* nempc_last_error(const nempc_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

// ---- structure (host only) -----------------------------------------------------------------------------
static int check_dims(int H, int x, int u) {
    if (H < 1 || x < 1 || u < 1 || x + u > NEMPC_MAX_D || x > NEMPC_LAYOUT_MAX_X) return NEMPC_EINVAL;
    return NEMPC_OK;
}

extern "C" int nempc_structure_counts(int32_t H, int32_t x, int32_t u, const uint8_t* quad_mask, int64_t* nnz_jac,
                                      int64_t* nnz_hes) {
    if (check_dims(H, x, u)) { SET_ERR((nempc_handle*)nullptr, "bad dims H=%d x=%d u=%d", H, x, u); return NEMPC_EINVAL; }
    NlpLayout L; nlp_layout_init(L, H, x, u, quad_mask);
    if (nnz_jac) *nnz_jac = L.nnz_jac;
    if (nnz_hes) *nnz_hes = L.nnz_hes;
    return NEMPC_OK;
}

extern "C" int nempc_structure_fill(int32_t H, int32_t x, int32_t u, const uint8_t* quad_mask, int32_t* jr, int32_t* jc,
                                    int32_t* hr, int32_t* hc) {
    if (check_dims(H, x, u)) { SET_ERR((nempc_handle*)nullptr, "bad dims H=%d x=%d u=%d", H, x, u); return NEMPC_EINVAL; }
    NlpLayout L; nlp_layout_init(L, H, x, u, quad_mask);
    if (jr && jc) {
        for (int t = 0; t < H; ++t)
            for (int p = 0; p < x; ++p) {
                const int r = t * x + p;
                if (t > 0) for (int q = 0; q < x; ++q) { int s = jac_slot_A(L, t, p, q); jr[s] = r; jc[s] = (t - 1) * x + q; }
                { int s = jac_slot_minus1(L, t, p); jr[s] = r; jc[s] = r; }
                for (int q = 0; q < u; ++q) { int s = jac_slot_B(L, t, p, q); jr[s] = r; jc[s] = H * x + t * u + q; }
            }
    }
    if (hr && hc) {
        for (int t = 0; t < H; ++t) {
            if (t > 0)
                for (int p = 0; p < x; ++p)
                    for (int q = 0; q <= p; ++q) { int s = hes_slot_xx(L, t, p, q); hr[s] = (t - 1) * x + p; hc[s] = (t - 1) * x + q; }
            for (int q = 0; q < u; ++q) {
                const int r = H * x + t * u + q;
                if (t > 0) for (int p = 0; p < x; ++p) { int s = hes_slot_ux(L, t, q, p); hr[s] = r; hc[s] = (t - 1) * x + p; }
                for (int r2 = 0; r2 <= q; ++r2) { int s = hes_slot_uu(L, t, q, r2); hr[s] = r; hc[s] = H * x + t * u + r2; }
            }
        }
        for (int p = 0; p < x; ++p)
            if (L.hes_last_slot[p] >= 0) { hr[L.hes_last_slot[p]] = (H - 1) * x + p; hc[L.hes_last_slot[p]] = (H - 1) * x + p; }
    }
    return NEMPC_OK;
}

static void rebuild_layout(nempc_handle* h) {
    std::vector<uint8_t> mask;
    if (h->has_objective) {
        mask.resize(h->lay.n);
        for (size_t i = 0; i < mask.size(); ++i) mask[i] = h->quad[i] != 0.0;
    }
    nlp_layout_init(h->lay, h->desc.horizon, h->desc.x_dim, h->desc.u_dim, mask.empty() ? nullptr : mask.data());
}

extern "C" int nempc_dims(const nempc_handle* h, int64_t* n, int64_t* m, int64_t* nnz_jac, int64_t* nnz_hes) {
    if (!h) return NEMPC_EINVAL;
    if (n) *n = h->lay.n;
    if (m) *m = h->lay.m;
    if (nnz_jac) *nnz_jac = h->lay.nnz_jac;
    if (nnz_hes) *nnz_hes = h->lay.nnz_hes;
    return NEMPC_OK;
}

extern "C" int nempc_structure(const nempc_handle* h, int32_t* jr, int32_t* jc, int32_t* hr, int32_t* hc) {
    if (!h) return NEMPC_EINVAL;
    std::vector<uint8_t> mask;
    if (h->has_objective) { mask.resize(h->lay.n); for (int i = 0; i < h->lay.n; ++i) mask[i] = h->quad[i] != 0.0; }
    return nempc_structure_fill(h->desc.horizon, h->desc.x_dim, h->desc.u_dim, mask.empty() ? nullptr : mask.data(), jr, jc, hr, hc);
}

// ---- lifetime ------------------------------------------------------------------------------------------------
static void free_device(nempc_handle* h) {
    for (int l = 0; l < NEMPC_MAXL; ++l) { cudaFree(h->dW[l]); cudaFree(h->dWT[l]); cudaFree(h->db[l]); }
    cudaFree(h->dlin); cudaFree(h->dquad); cudaFree(h->dref); cudaFree(h->gws); cudaFree(h->d_tcimg); cudaFree(h->d_tccb);
    cudaFree(h->d_tvp); cudaFree(h->d_p); cudaFree(h->d_tcwx); cudaFree(h->d_wblob); cudaFree(h->d_wcb); cudaFree(h->wide_scratch);
    cudaFree(h->sv_buf); cudaFree(h->sv_lb); cudaFree(h->sv_ub); cudaFree(h->sv_counts); if (h->sv_counts_host) cudaFreeHost(h->sv_counts_host);
    for (int i = 0; i < 10; ++i) cudaFree(h->st_buf[i]);
    if (h->hk_exec) cudaGraphExecDestroy(h->hk_exec);
    cudaFree(h->dmma_scratch);
    if (h->sk_exec) cudaGraphExecDestroy(h->sk_exec);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (int i = 0; i < 3; ++i) if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    for (int i = 0; i < 3; ++i) if (h->ev_kern[i]) cudaEventDestroy(h->ev_kern[i]);
    if (h->stream) cudaStreamDestroy(h->stream);
    for (int i = 0; i < 3; ++i) if (h->pipe[i]) cudaStreamDestroy(h->pipe[i]);
}

extern "C" int nempc_create(const nempc_desc* desc, nempc_handle** out) {
    if (!desc || !out) { SET_ERR((nempc_handle*)nullptr, "null argument"); return NEMPC_EINVAL; }
    *out = nullptr;
    const nempc_desc& D = *desc;
    if (check_dims(D.horizon, D.x_dim, D.u_dim)) {
        SET_ERR((nempc_handle*)nullptr, "unsupported dims: H=%d x_dim=%d u_dim=%d (need H>=1, x_dim+u_dim<=%d)", D.horizon, D.x_dim, D.u_dim, NEMPC_MAX_D);
        return NEMPC_EINVAL;
    }
    if (D.n_layers < 2 || D.n_layers > NEMPC_MAX_LAYERS) { SET_ERR((nempc_handle*)nullptr, "n_layers must be in [2,%d]", NEMPC_MAX_LAYERS); return NEMPC_EINVAL; }
    if (D.widths[D.n_layers - 1] != D.x_dim) { SET_ERR((nempc_handle*)nullptr, "last layer width %d != x_dim %d", D.widths[D.n_layers - 1], D.x_dim); return NEMPC_EINVAL; }
    for (int l = 0; l < D.n_layers; ++l) if (D.widths[l] < 1 || D.widths[l] > 4096) { SET_ERR((nempc_handle*)nullptr, "bad width"); return NEMPC_EINVAL; }
    if (D.activation < 0 || D.activation > NEMPC_ACT_RELU || D.integrator < 0 || D.integrator > NEMPC_INTEG_RK4) { SET_ERR((nempc_handle*)nullptr, "bad activation/integrator"); return NEMPC_EINVAL; }
    if (D.integrator == NEMPC_INTEG_RK4 && !(D.dt > 0.0)) { SET_ERR((nempc_handle*)nullptr, "RK4 needs dt > 0"); return NEMPC_EINVAL; }
    if ((D.compute_dtype != NEMPC_F32 && D.compute_dtype != NEMPC_F64) || (D.io_dtype != NEMPC_F32 && D.io_dtype != NEMPC_F64) ||
        (D.compute_dtype == NEMPC_F64 && D.io_dtype == NEMPC_F32)) { SET_ERR((nempc_handle*)nullptr, "unsupported dtype combination"); return NEMPC_EINVAL; }

    if (D.tvp_dim < 0 || D.p_dim < 0 || D.tvp_dim + D.p_dim > NEMPC_MAX_EXO) { SET_ERR((nempc_handle*)nullptr, "tvp_dim + p_dim must be in [0,%d]", NEMPC_MAX_EXO); return NEMPC_EINVAL; }
    if (D.tvp_dim + D.p_dim > 0 && D.kernel == NEMPC_KERNEL_FAST) {
        SET_ERR((nempc_handle*)nullptr, "tvp / p model inputs are served by the generic and tensor-core kernels (kernel must be AUTO, GENERIC or TC)");
        return NEMPC_EUNSUPPORTED;
    }

    nempc_handle* h = new nempc_handle();
    h->desc = D; h->d = D.x_dim + D.u_dim; h->L = D.n_layers; h->n_ext = D.tvp_dim + D.p_dim;
    h->dims.push_back(h->d + h->n_ext);         // fan-in of the first layer: [x, u, tvp, p] (model/tensorflow.py:39-47)
    for (int l = 0; l < D.n_layers; ++l) h->dims.push_back(D.widths[l]);
    h->W.resize(h->L); h->bvec.resize(h->L); h->wset.assign(h->L, false);
    nlp_layout_init(h->lay, D.horizon, D.x_dim, D.u_dim, nullptr);

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= D.device || D.device < 0) {
        SET_ERR((nempc_handle*)nullptr, "no usable CUDA device %d (%s); libnempc has no CPU path", D.device, e != cudaSuccess ? cudaGetErrorString(e) : "ordinal out of range");
        delete h; return NEMPC_ECUDA;
    }
    if ((e = cudaSetDevice(D.device)) != cudaSuccess) { SET_ERR((nempc_handle*)nullptr, "cudaSetDevice: %s", cudaGetErrorString(e)); delete h; return NEMPC_ECUDA; }
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, D.device);
    cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, D.device);
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) { SET_ERR((nempc_handle*)nullptr, "stream: %s", cudaGetErrorString(e)); delete h; return NEMPC_ECUDA; }
    for (int i = 0; i < 3; ++i)
        if ((e = cudaStreamCreateWithFlags(&h->pipe[i], cudaStreamNonBlocking)) != cudaSuccess) { SET_ERR((nempc_handle*)nullptr, "stream: %s", cudaGetErrorString(e)); free_device(h); delete h; return NEMPC_ECUDA; }

    // kernel choice
    h->fast_id = fast_shape_id(D);
    h->fast64_id = (D.kernel == NEMPC_KERNEL_AUTO || D.kernel == NEMPC_KERNEL_FAST) ? fast64_shape_id(D) : -1;
    if (D.kernel == NEMPC_KERNEL_FAST && h->fast_id < 0 && h->fast64_id < 0) {
        SET_ERR((nempc_handle*)nullptr, "NEMPC_KERNEL_FAST requested but no register-resident instantiation matches this network");
        free_device(h); delete h; return NEMPC_EUNSUPPORTED;
    }
    h->use_fast = (h->fast_id >= 0 && (D.kernel == NEMPC_KERNEL_AUTO || D.kernel == NEMPC_KERNEL_FAST)) ? 1 : 0;
    h->tc_id = tc_shape_id(D);
    // hidden width 128 is served by two tensor-core kernels: the forward second-order one (nempc_tc.cuh) and the adjoint-form one
    // (nempc_wide.cuh, instantiated for (4,1) and (2,1)).  Measured on 5-128x3-4, H=100, B=8192 (tools/c3_compare.py): RK4 29.6 vs 39.0 ms
    // (tc wins), discrete with Hessian 9.0 vs 6.6 ms (adjoint form wins), Jacobian only 2.3 vs 3.1 ms (tc wins).  So the handle keeps
    // both images and the single-stage integrators send their Hessian evaluations to the adjoint-form kernel.
    // NEMPC_WIDE128 = 0 never / 1 always / unset: as measured.
    int wide128 = -1;
    if (getenv("NEMPC_WIDE128")) wide128 = atoi(getenv("NEMPC_WIDE128"));
    if (wide128 == 1 && D.widths[0] == 128 && wide_shape_id(D) >= 0) h->tc_id = -1;
    h->use_tc = (h->tc_id >= 0 && !h->use_fast && (D.kernel == NEMPC_KERNEL_AUTO || D.kernel == NEMPC_KERNEL_TC)) ? 1 : 0;
    h->dmma_id = (!h->use_fast && h->fast64_id < 0 && (D.kernel == NEMPC_KERNEL_AUTO || D.kernel == NEMPC_KERNEL_TC)) ? dmma_shape_id(D) : -1;
    h->wide_id = wide_shape_id(D);
    h->use_wide = (h->wide_id >= 0 && !h->use_fast && !h->use_tc && (D.kernel == NEMPC_KERNEL_AUTO || D.kernel == NEMPC_KERNEL_TC)) ? 1 : 0;
    h->wide_hes = (h->use_tc && h->wide_id >= 0 && wide128 != 0 && D.integrator != NEMPC_INTEG_RK4) ? 1 : 0;
    if (D.kernel == NEMPC_KERNEL_TC && !h->use_tc && !h->use_wide && h->dmma_id < 0) {
        SET_ERR((nempc_handle*)nullptr, "NEMPC_KERNEL_TC requested but no tensor-core instantiation matches this network (f32: tanh, hidden layers all 256 / 128 / 64 / 32 wide; f64: all 128 / 64 wide)");
        free_device(h); delete h; return NEMPC_EUNSUPPORTED;
    }

    // generic launch geometry (also used by eval_blocks / model_eval of fast handles)
    int sum_h = 0, hmax = 0;
    for (int l = 0; l + 1 < h->L; ++l) { sum_h += D.widths[l]; hmax = std::max(hmax, D.widths[l]); }
    h->sl = make_slot_layout(D.x_dim, h->d, sum_h, hmax, h->n_ext);
    h->dmax = h->d <= 4 ? 4 : (h->d <= 8 ? 8 : 16);
    h->tps = std::min(256, std::max(32, (hmax + 31) / 32 * 32));
    const size_t per_slot = (size_t)h->sl.total * dsize(D.compute_dtype);
    int slots = std::max(1, 256 / h->tps);
    if (h->tps > 32) slots = std::min(slots, 15);
    const size_t budget = (size_t)std::max(0, h->max_smem_optin - 1024);
    while (slots > 1 && per_slot * slots > budget / 2) --slots;    // keep >= 2 CTAs/SM when possible
    if (per_slot * slots > budget) { h->global_ws = true; h->smem_bytes = 0; }
    else h->smem_bytes = per_slot * slots;
    h->slots = slots;

    char nm[288];
    if (h->use_fast) snprintf(nm, sizeof nm, "nempc_fast_kernel<x=%d,u=%d,h1=%d,h2=%d> f32 (thread/step, weights in constant bank%s)", D.x_dim, D.u_dim, D.widths[0], D.widths[1],
                              D.kernel == NEMPC_KERNEL_AUTO ? "; warp/step nempc_small_kernel for small batches" : "");
    else if (h->fast64_id >= 0) snprintf(nm, sizeof nm, "nempc_fast64_kernel<x=%d,u=%d,h1=%d,h2=%d> f64 (thread/step, DFMA, weights in constant bank%s)", D.x_dim, D.u_dim, D.widths[0], D.widths[1],
                                         D.kernel == NEMPC_KERNEL_AUTO ? "; generic kernel for small batches" : "");
    else if (h->dmma_id >= 0) snprintf(nm, sizeof nm, "nempc_dmma_net_kernel<hidden=%dx%d> f64 DMMA m8n8k4 (forward second order, weights streamed through a cp.async ring) + nempc_dmma_stage_kernel<x=%d,u=%d>%s",
                                      D.n_layers - 1, D.widths[0], D.x_dim, D.u_dim, D.kernel == NEMPC_KERNEL_AUTO ? "; generic kernel for small batches" : "");
    else if (h->use_wide) snprintf(nm, sizeof nm, "nempc_wide_kernel<x=%d,u=%d,hidden=%dx%d> tcgen05 split-f16 (adjoint form, weights streamed through a TMA ring%s)", D.x_dim, D.u_dim, D.n_layers - 1, D.widths[0],
                               h->wide_id >= 100 ? "; dimensions read at run time" : "");
    else if (h->use_tc) snprintf(nm, sizeof nm, "nempc_tc_kernel<x=%d,u=%d,hidden=%dx%d> tcgen05 split-f16 (forward second order, weights resident in smem%s)", D.x_dim, D.u_dim, D.n_layers - 1, D.widths[0],
                                 h->wide_hes ? "; Hessian evaluations on the adjoint-form nempc_wide_kernel<hw=128>" : "");
    else snprintf(nm, sizeof nm, "nempc_generic_kernel<%s,dmax=%d> tps=%d slots=%d %s", D.compute_dtype == NEMPC_F64 ? "f64" : "f32", h->dmax, h->tps, h->slots, h->global_ws ? "global-ws" : "smem-ws");
    h->kname = nm;
    *out = h;
    return NEMPC_OK;
}

extern "C" int nempc_destroy(nempc_handle* h) {
    if (!h) return NEMPC_OK;
    cudaSetDevice(h->desc.device);
    free_device(h);
    delete h;
    return NEMPC_OK;
}

template <typename FW> static FW* fast_blob(nempc_handle* h) {       // 16-byte aligned view into the byte blob
    uintptr_t a = reinterpret_cast<uintptr_t>(h->fastw.data());
    return reinterpret_cast<FW*>((a + 15) & ~uintptr_t(15));
}

template <typename T> static int upload_layer(nempc_handle* h, int l, int fin, int fout) {
    std::vector<T> w(fin * (size_t)fout), wt(fin * (size_t)fout), b(fout);
    for (int i = 0; i < fin; ++i)
        for (int j = 0; j < fout; ++j) { w[i * (size_t)fout + j] = (T)h->W[l][i * (size_t)fout + j]; wt[j * (size_t)fin + i] = (T)h->W[l][i * (size_t)fout + j]; }
    for (int j = 0; j < fout; ++j) b[j] = (T)h->bvec[l][j];
    if (!h->dW[l]) {
        CU(h, cudaMalloc(&h->dW[l], w.size() * sizeof(T)));
        CU(h, cudaMalloc(&h->dWT[l], w.size() * sizeof(T)));
        CU(h, cudaMalloc(&h->db[l], b.size() * sizeof(T)));
    }
    CU(h, cudaMemcpy(h->dW[l], w.data(), w.size() * sizeof(T), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->dWT[l], wt.data(), wt.size() * sizeof(T), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->db[l], b.data(), b.size() * sizeof(T), cudaMemcpyHostToDevice));
    return NEMPC_OK;
}

template <int X, int U, int H1, int H2> static void fill_fast64(nempc_handle* h) {
    typedef Fast64Weights<X, U, H1, H2> FW;
    h->fast64w.assign(sizeof(FW) + 16, 0);
    uintptr_t a = (reinterpret_cast<uintptr_t>(h->fast64w.data()) + 15) & ~(uintptr_t)15;
    fill_fast64_weights<X, U, H1, H2>(*reinterpret_cast<FW*>(a), h->W[0].data(), h->bvec[0].data(), h->W[1].data(), h->bvec[1].data(),
                                      h->W[2].data(), h->bvec[2].data());
}
template <int X, int U, int H1, int H2, int NCHUNK> static void fill_fast(nempc_handle* h) {
    typedef FastWeights<X, U, H1, H2, NCHUNK> FW;
    h->fastw.assign(sizeof(FW) + 16, 0);
    fill_fast_weights<X, U, H1, H2, NCHUNK>(*fast_blob<FW>(h), h->W[0].data(), h->bvec[0].data(),
                                            h->W[1].data(), h->bvec[1].data(), h->W[2].data(), h->bvec[2].data());
}

// f16 hi/lo operand images of the hidden-to-hidden layers + the f32 constant block of nempc_tc_kernel
static int upload_tc(nempc_handle* h) {
    const int nhid = h->L - 1, HW = kTcShapes[h->tc_id].hw, x = h->desc.x_dim, d = h->d, xp = (x + 3) / 4 * 4, nmm = nhid - 1;
    std::vector<__half> img((size_t)nmm * 2 * HW * HW);
    for (int l = 0; l < nmm; ++l) {
        const std::vector<double>& W = h->W[l + 1];                       // [in][out]
        __half* hi = img.data() + (size_t)l * 2 * HW * HW;
        __half* lo = hi + (size_t)HW * HW;
        for (int i = 0; i < HW; ++i)
            for (int j = 0; j < HW; ++j) {
                const float w = (float)W[(size_t)i * HW + j];
                const __half whi = __float2half_rn(w);
                const size_t off = tc_img_index(j, i, HW);                    // B[n = out][k = in]
                hi[off] = whi;
                lo[off] = __float2half_rn((w - __half2float(whi)) * NEMPC_TC_LO_SCALE);
            }
    }
    std::vector<float> cb((size_t)d * HW + (size_t)HW * xp + (size_t)nhid * HW + xp, 0.f);
    float* W0 = cb.data(); float* Wout = W0 + (size_t)d * HW; float* b = Wout + (size_t)HW * xp; float* bout = b + (size_t)nhid * HW;
    for (int c = 0; c < d; ++c) for (int j = 0; j < HW; ++j) W0[c * HW + j] = (float)h->W[0][(size_t)c * HW + j];
    for (int j = 0; j < HW; ++j) for (int p = 0; p < x; ++p) Wout[j * xp + p] = (float)h->W[nhid][(size_t)j * x + p];
    for (int l = 0; l < nhid; ++l) for (int j = 0; j < HW; ++j) b[l * HW + j] = (float)h->bvec[l][j];
    for (int p = 0; p < x; ++p) bout[p] = (float)h->bvec[nhid][p];
    if (!h->d_tcimg) {
        CU(h, cudaMalloc(&h->d_tcimg, img.size() * sizeof(__half)));
        CU(h, cudaMalloc((void**)&h->d_tccb, cb.size() * sizeof(float)));
    }
    CU(h, cudaMemcpy(h->d_tcimg, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->d_tccb, cb.data(), cb.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (h->n_ext > 0) {                                   // rows d .. d+n_ext-1 of the first kernel: the tvp / p inputs
        std::vector<float> wx((size_t)h->n_ext * HW);
        for (int e = 0; e < h->n_ext; ++e) for (int j = 0; j < HW; ++j) wx[(size_t)e * HW + j] = (float)h->W[0][(size_t)(d + e) * HW + j];
        if (!h->d_tcwx) CU(h, cudaMalloc((void**)&h->d_tcwx, wx.size() * sizeof(float)));
        CU(h, cudaMemcpy(h->d_tcwx, wx.data(), wx.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    return NEMPC_OK;
}

// streamed operand images of nempc_wide_kernel: per GEMM and K step [2^11 hi | hi | lo] x [K chunk] x [n] x [8 halves]
static int upload_wide(nempc_handle* h) {
    const int nhid = h->L - 1, HW = h->desc.widths[0], x = h->desc.x_dim, d = h->d;
    std::vector<__half> blob;
    bool range_ok = true;
    // Bm(n, k): row n of the B operand, contraction index k
    auto add = [&](int N, int K, auto&& Bm) {
        WideGemm gm{};
        gm.ksteps = (uint32_t)(K / 16); gm.n = (uint32_t)N;
        const int Nh = N / 2;                                   // CTA r of the pair streams rows [r Nh, (r + 1) Nh)
        for (int r = 0; r < 2; ++r) {
            gm.off[r] = (uint32_t)(blob.size() * sizeof(__half));
            const size_t base = blob.size();
            blob.resize(base + (size_t)(K / 16) * 48 * Nh, __float2half_rn(0.f));
            for (int nl = 0; nl < Nh; ++nl)
                for (int k = 0; k < K; ++k) {
                    const float w = (float)Bm(r * Nh + nl, k);
                    const __half whi = __float2half_rn(w);
                    const float whf = __half2float(whi);
                    if (fabsf(whf) * NEMPC_TC_LO_SCALE > 60000.f) range_ok = false;
                    blob[base + wide_img_index(Nh, K / 16, 0, nl, k)] = __float2half_rn(whf * NEMPC_TC_LO_SCALE);
                    blob[base + wide_img_index(Nh, K / 16, 1, nl, k)] = whi;
                    blob[base + wide_img_index(Nh, K / 16, 2, nl, k)] = __float2half_rn((w - whf) * NEMPC_TC_LO_SCALE);
                }
        }
        return gm;
    };
    WideNet& wn = h->wnet;
    wn = WideNet{};
    wn.nhid = nhid;
    const std::vector<double>& W0 = h->W[0];                 // [d][HW]
    const std::vector<double>& Wo = h->W[nhid];              // [HW][x]
    wn.in_f = add(HW, 16, [&](int n, int k) { return k < d ? W0[(size_t)k * HW + n] : 0.0; });
    for (int l = 1; l < nhid; ++l) { const std::vector<double>& W = h->W[l]; wn.hid_f[l - 1] = add(HW, HW, [&](int n, int k) { return W[(size_t)k * HW + n]; }); }
    wn.out_f = add(NEMPC_WIDE_NOUT, HW, [&](int n, int k) { return n < x ? Wo[(size_t)k * x + n] : 0.0; });
    wn.out_b = add(HW, 16, [&](int n, int k) { return k < x ? Wo[(size_t)n * x + k] : 0.0; });
    for (int l = 1; l < nhid; ++l) { const std::vector<double>& W = h->W[l]; wn.hid_b[l - 1] = add(HW, HW, [&](int n, int k) { return W[(size_t)n * HW + k]; }); }
    wn.in_b = add(NEMPC_WIDE_NOUT, HW, [&](int n, int k) { return n < d ? W0[(size_t)n * HW + k] : 0.0; });
    if (!range_ok) { SET_ERR(h, "nempc_wide_kernel: a weight exceeds the f16 range of the scaled operand image (|w| < 29)"); return NEMPC_EUNSUPPORTED; }
    std::vector<float> cb((size_t)NEMPC_WIDE_MAXHID * HW + 16, 0.f);
    for (int l = 0; l < nhid; ++l) for (int j = 0; j < HW; ++j) cb[(size_t)l * HW + j] = (float)h->bvec[l][j];
    for (int p = 0; p < x; ++p) cb[(size_t)NEMPC_WIDE_MAXHID * HW + p] = (float)h->bvec[nhid][p];
    CU(h, cudaDeviceSynchronize());                       // earlier launches may still stream the old images
    cudaFree(h->d_wblob); h->d_wblob = nullptr;
    CU(h, cudaMalloc((void**)&h->d_wblob, blob.size() * sizeof(__half)));
    if (!h->d_wcb) CU(h, cudaMalloc((void**)&h->d_wcb, cb.size() * sizeof(float)));
    CU(h, cudaMemcpy(h->d_wblob, blob.data(), blob.size() * sizeof(__half), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->d_wcb, cb.data(), cb.size() * sizeof(float), cudaMemcpyHostToDevice));
    return NEMPC_OK;
}

// kernel parameters (weights of the register-resident kernel, the sparse layout) are baked into a captured graph by value
static void drop_solve_graph(nempc_handle* h) {
    if (h->sk_exec) { cudaGraphExecDestroy(h->sk_exec); h->sk_exec = nullptr; }
    h->sk_seen = 0;
}
static void drop_host_graph(nempc_handle* h) {
    if (h->hk_exec) { cudaGraphExecDestroy(h->hk_exec); h->hk_exec = nullptr; }
    h->hk_seen = 0;
    drop_solve_graph(h);
}

extern "C" int nempc_set_weights(nempc_handle* h, int32_t layer, const double* W, const double* b) {
    if (!h || !W || !b || layer < 0 || layer >= h->L) { SET_ERR(h, "nempc_set_weights: bad argument"); return NEMPC_EINVAL; }
    drop_host_graph(h);
    const int fin = h->dims[layer], fout = h->dims[layer + 1];
    h->W[layer].assign(W, W + (size_t)fin * fout);
    h->bvec[layer].assign(b, b + fout);
    CU(h, cudaSetDevice(h->desc.device));
    CU(h, cudaDeviceSynchronize());                       // evaluations still in flight on a caller's stream read the old weight images
    int rc = h->desc.compute_dtype == NEMPC_F64 ? upload_layer<double>(h, layer, fin, fout) : upload_layer<float>(h, layer, fin, fout);
    if (rc) return rc;
    h->wset[layer] = true;
    bool all = true; for (bool s : h->wset) all = all && s;
    if (all && h->fast_id >= 0) {
        switch (h->fast_id) {
            case 0: fill_fast<2, 1, 30, 30, NEMPC_FAST_NCHUNK30>(h); break;
            case 1: fill_fast<2, 1, 32, 32, 2>(h); break;
            case 2: fill_fast<2, 1, 16, 16, 1>(h); break;
        }
    }
    if (all && h->fast64_id >= 0) {
        switch (h->fast64_id) {
            case 0: fill_fast64<2, 1, 30, 30>(h); break;
            case 1: fill_fast64<2, 1, 32, 32>(h); break;
            case 2: fill_fast64<2, 1, 16, 16>(h); break;
        }
    }
    if (all && h->tc_id >= 0) { rc = upload_tc(h); if (rc) return rc; }
    if (all && (h->use_wide || h->wide_hes)) { rc = upload_wide(h); if (rc) return rc; }
    return NEMPC_OK;
}

extern "C" int nempc_set_objective(nempc_handle* h, const double* lin, const double* quad, const double* ref) {
    if (!h) return NEMPC_EINVAL;
    drop_host_graph(h);
    const int n = h->lay.n;
    h->lin.assign(n, 0.0); h->quad.assign(n, 0.0); h->ref.assign(n, 0.0);
    if (lin) h->lin.assign(lin, lin + n);
    if (quad) h->quad.assign(quad, quad + n);
    if (ref) h->ref.assign(ref, ref + n);
    h->has_objective = true;
    rebuild_layout(h);
    CU(h, cudaSetDevice(h->desc.device));
    CU(h, cudaDeviceSynchronize());                       // evaluations / solves still in flight on a caller's stream read the old cost
    if (!h->dlin) {
        CU(h, cudaMalloc(&h->dlin, n * sizeof(double)));
        CU(h, cudaMalloc(&h->dquad, n * sizeof(double)));
        CU(h, cudaMalloc(&h->dref, n * sizeof(double)));
    }
    CU(h, cudaMemcpy(h->dlin, h->lin.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->dquad, h->quad.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->dref, h->ref.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    return NEMPC_OK;
}

extern "C" int nempc_set_exogenous(nempc_handle* h, int64_t tvp_rows, const double* tvp, int64_t p_rows, const double* p) {
    if (!h) return NEMPC_EINVAL;
    const int td = h->desc.tvp_dim, pd = h->desc.p_dim;
    if ((td > 0 && (!tvp || tvp_rows < 1)) || (pd > 0 && (!p || p_rows < 1)) || (td == 0 && tvp) || (pd == 0 && p)) {
        SET_ERR(h, "nempc_set_exogenous: arguments do not match tvp_dim=%d p_dim=%d", td, pd);
        return NEMPC_EINVAL;
    }
    drop_host_graph(h);
    CU(h, cudaSetDevice(h->desc.device));
    CU(h, cudaDeviceSynchronize());                       // earlier launches may still read the old rows
    if (td > 0) {
        const size_t bytes = (size_t)tvp_rows * td * sizeof(double);
        if (bytes > h->tvp_cap) { cudaFree(h->d_tvp); h->d_tvp = nullptr; h->tvp_cap = 0; CU(h, cudaMalloc(&h->d_tvp, bytes)); h->tvp_cap = bytes; }
        CU(h, cudaMemcpy(h->d_tvp, tvp, bytes, cudaMemcpyDefault));
        h->tvp_rows = tvp_rows;
    }
    if (pd > 0) {
        const size_t bytes = (size_t)p_rows * pd * sizeof(double);
        if (bytes > h->p_cap) { cudaFree(h->d_p); h->d_p = nullptr; h->p_cap = 0; CU(h, cudaMalloc(&h->d_p, bytes)); h->p_cap = bytes; }
        CU(h, cudaMemcpy(h->d_p, p, bytes, cudaMemcpyDefault));
        h->p_rows = p_rows;
    }
    return NEMPC_OK;
}

// ---- launches ------------------------------------------------------------------------------------------------
static int ready(nempc_handle* h) {
    if (!h) return NEMPC_EINVAL;
    for (size_t l = 0; l < h->wset.size(); ++l)
        if (!h->wset[l]) { SET_ERR(h, "weights of layer %zu not set", l); return NEMPC_ESTATE; }
    return NEMPC_OK;
}

template <typename T> static NetView<T> make_net(const nempc_handle* h) {
    NetView<T> n{};
    n.L = h->L; n.act = h->desc.activation; n.tvp_dim = h->desc.tvp_dim; n.p_dim = h->desc.p_dim;
    int off = 0, hm = 0;
    for (int l = 0; l <= h->L; ++l) n.dims[l] = h->dims[l];
    for (int l = 0; l + 1 < h->L; ++l) { n.hoff[l] = off; off += h->dims[l + 1]; hm = std::max(hm, h->dims[l + 1]); }
    n.sum_h = off; n.hmax = hm;
    for (int l = 0; l < h->L; ++l) { n.W[l] = (const T*)h->dW[l]; n.WT[l] = (const T*)h->dWT[l]; n.b[l] = (const T*)h->db[l]; }
    return n;
}

// attach the handle's exogenous rows to a launch: problems [exo_base, exo_base + B) of the set given to nempc_set_exogenous
template <typename TIO>
static int bind_exogenous(nempc_handle* h, EvalArgs<TIO>& ax, bool model_mode) {
    const int td = h->desc.tvp_dim, pd = h->desc.p_dim, H = h->desc.horizon;
    if ((td > 0 && !h->d_tvp) || (pd > 0 && !h->d_p)) { SET_ERR(h, "the model has tvp / p inputs: call nempc_set_exogenous first"); return NEMPC_ESTATE; }
    const long long units = model_mode ? ax.nsteps : ax.nsteps / H;          // samples (model) or problems (NLP)
    ax.tvp = h->d_tvp; ax.p = h->d_p; ax.tvp_bstride = 0; ax.p_bstride = 0;
    if (td > 0) {
        const long long per = model_mode ? 1 : H;                             // tvp rows per unit
        if (!model_mode && h->tvp_rows == H) ax.tvp_bstride = 0;              // one (H, tvp_dim) table shared by every problem
        else {
            if (h->tvp_rows < (h->exo_base + units) * per) { SET_ERR(h, "nempc_set_exogenous holds %lld tvp rows, this call needs %lld", h->tvp_rows, (h->exo_base + units) * per); return NEMPC_EINVAL; }
            ax.tvp_bstride = (long long)H * td;
            ax.tvp = h->d_tvp + h->exo_base * per * td;
        }
    }
    if (pd > 0) {
        if (h->p_rows == 1) ax.p_bstride = 0;
        else if (model_mode) { SET_ERR(h, "nempc_model_eval takes one p row"); return NEMPC_EINVAL; }
        else {
            if (h->p_rows < h->exo_base + units) { SET_ERR(h, "nempc_set_exogenous holds %lld p rows, this call needs %lld", h->p_rows, h->exo_base + units); return NEMPC_EINVAL; }
            ax.p_bstride = pd;
            ax.p = h->d_p + h->exo_base * pd;
        }
    }
    return NEMPC_OK;
}

template <typename T, typename TIO, int DMAX>
static int launch_generic_t(nempc_handle* h, const EvalArgs<TIO>& ar, bool model_mode, cudaStream_t s) {
    auto kern = nempc_generic_kernel<T, TIO, DMAX>;
    const int threads = h->tps * h->slots;
    if (h->smem_bytes > 48 * 1024) CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    int occ = 1;
    CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, h->smem_bytes));
    occ = std::max(1, occ);
    const long long tiles = (ar.nsteps + h->slots - 1) / h->slots;
    const long long grid = std::max(1LL, std::min(tiles, (long long)h->sm_count * occ));
    T* gws = nullptr;
    if (h->global_ws) {
        const size_t need = (size_t)grid * h->slots * h->sl.total * sizeof(T);
        if (need > h->gws_bytes) {
            CU(h, cudaStreamSynchronize(s));
            cudaFree(h->gws); h->gws = nullptr; h->gws_bytes = 0;
            CU(h, cudaMalloc(&h->gws, need));
            h->gws_bytes = need;
        }
        gws = (T*)h->gws;
    }
    NetView<T> net = make_net<T>(h);
    StageTable<T> st = make_stage_table<T>(h->desc.integrator == NEMPC_INTEG_RK4 && !model_mode, h->desc.dt);
    EvalArgs<TIO> ax = ar;
    if (h->n_ext > 0) {
        int rc = bind_exogenous(h, ax, model_mode);
        if (rc) return rc;
    }
    kern<<<(unsigned)grid, threads, h->smem_bytes, s>>>(net, st, h->lay, h->sl, ax, h->tps, gws);
    CU(h, cudaGetLastError());
    h->launches++;
    return NEMPC_OK;
}

template <typename T, typename TIO>
static int launch_generic_d(nempc_handle* h, const EvalArgs<TIO>& ar, bool model_mode, cudaStream_t s) {
    switch (h->dmax) {
        case 4: return launch_generic_t<T, TIO, 4>(h, ar, model_mode, s);
        case 8: return launch_generic_t<T, TIO, 8>(h, ar, model_mode, s);
        default: return launch_generic_t<T, TIO, 16>(h, ar, model_mode, s);
    }
}

template <typename TIO> static int launch_generic(nempc_handle* h, const EvalArgs<TIO>& ar, bool model_mode, cudaStream_t s);
template <> int launch_generic<float>(nempc_handle* h, const EvalArgs<float>& ar, bool model_mode, cudaStream_t s) {
    return launch_generic_d<float, float>(h, ar, model_mode, s);
}
template <> int launch_generic<double>(nempc_handle* h, const EvalArgs<double>& ar, bool model_mode, cudaStream_t s) {
    return h->desc.compute_dtype == NEMPC_F64 ? launch_generic_d<double, double>(h, ar, model_mode, s)
                                              : launch_generic_d<float, double>(h, ar, model_mode, s);
}

template <int X, int U, int H1, int H2, int NCHUNK, typename TIO>
static int launch_fast_shape(nempc_handle* h, const EvalArgs<TIO>& ar, int mode, cudaStream_t s) {
    typedef FastWeights<X, U, H1, H2, NCHUNK> FW;
    const FW& w = *fast_blob<FW>(h);
    StageTable<float> st = make_stage_table<float>(h->desc.integrator == NEMPC_INTEG_RK4, h->desc.dt);
    const int threads = NEMPC_FAST_THREADS;
    const long long blocks = std::max(1LL, (ar.nsteps + threads - 1) / threads);
    const unsigned grid = (unsigned)std::min(blocks, (long long)h->sm_count * 64);
    const size_t smem = (size_t)FastScratch<X, U, H1, H2>::count(mode) * threads * sizeof(float);
    switch (mode) {
        case 0: nempc_fast_kernel<X, U, H1, H2, NCHUNK, 0, TIO><<<grid, threads, smem, s>>>(w, st, h->lay, ar); break;
        case 1: nempc_fast_kernel<X, U, H1, H2, NCHUNK, 1, TIO><<<grid, threads, smem, s>>>(w, st, h->lay, ar); break;
        default: nempc_fast_kernel<X, U, H1, H2, NCHUNK, 2, TIO><<<grid, threads, smem, s>>>(w, st, h->lay, ar); break;
    }
    CU(h, cudaGetLastError());
    h->launches++;
    return NEMPC_OK;
}

// float64 arithmetic, register resident (nempc_fast64.cuh): the kernels live in their own translation unit (nempc_fast64_tu.cu, see
// build.py) behind this one internal entry point
#ifndef NEMPC_FAST64_MIN_STEPS
#define NEMPC_FAST64_MIN_STEPS 4096      // below this many horizon steps one thread per step leaves the GPU empty: the generic kernel takes over
#endif
int nempc_fast64_launch(int shape_id, int mode, const void* weights, const StageTable<double>& st, const NlpLayout& L, const EvalArgs<double>& ar,
                        int sm_count, cudaStream_t s);
static int launch_fast64(nempc_handle* h, const EvalArgs<double>& ar, int mode, cudaStream_t s) {
    const uintptr_t a = (reinterpret_cast<uintptr_t>(h->fast64w.data()) + 15) & ~(uintptr_t)15;
    StageTable<double> st = make_stage_table<double>(h->desc.integrator == NEMPC_INTEG_RK4, h->desc.dt);
    const int rc = nempc_fast64_launch(h->fast64_id, mode, reinterpret_cast<const void*>(a), st, h->lay, ar, h->sm_count, s);
    if (rc == -1) { SET_ERR(h, "internal: bad fast64_id"); return NEMPC_EINVAL; }
    if (rc != 0) { SET_ERR(h, "nempc_fast64_kernel launch: %s", cudaGetErrorString((cudaError_t)rc)); return NEMPC_ECUDA; }
    h->launches++;
    return NEMPC_OK;
}
template <typename TIO> static bool take_fast64(nempc_handle*, const EvalArgs<TIO>&) { return false; }
template <> bool take_fast64<double>(nempc_handle* h, const EvalArgs<double>& ar) {
    static const long long min_steps = getenv("NEMPC_FAST64_MIN_STEPS") ? atoll(getenv("NEMPC_FAST64_MIN_STEPS")) : NEMPC_FAST64_MIN_STEPS;
    return h->fast64_id >= 0 && (ar.nsteps >= min_steps || h->desc.kernel == NEMPC_KERNEL_FAST);
}
template <typename TIO> static int run_fast64(nempc_handle* h, const EvalArgs<TIO>&, int, cudaStream_t) { SET_ERR(h, "internal: fast64 needs float64 I/O"); return NEMPC_EINVAL; }
template <> int run_fast64<double>(nempc_handle* h, const EvalArgs<double>& ar, int mode, cudaStream_t s) { return launch_fast64(h, ar, mode, s); }

// small batches of the same networks: one WARP per horizon step (nempc_small.cuh) -- latency instead of throughput
#ifndef NEMPC_SMALL_MAX_STEPS
#define NEMPC_SMALL_MAX_STEPS 6144      // measured crossover with the thread-per-step kernel: 52 vs 63 us at 6400 steps, 91 vs 64 us at 12800
#endif
template <int X, int U, int H1, int H2, typename TIO>
static int launch_small_shape(nempc_handle* h, const EvalArgs<TIO>& ar, int mode, cudaStream_t s) {
    StageTable<float> st = make_stage_table<float>(h->desc.integrator == NEMPC_INTEG_RK4, h->desc.dt);
    const int threads = 128;                                            // 4 warps = 4 steps per CTA
    const unsigned grid = (unsigned)((ar.nsteps * 32 + threads - 1) / threads);
    const float *W1 = (const float*)h->dW[0], *b1 = (const float*)h->db[0], *W2 = (const float*)h->dW[1], *b2 = (const float*)h->db[1],
                *W3 = (const float*)h->dW[2], *b3 = (const float*)h->db[2];
    switch (mode) {
        case 0: nempc_small_kernel<X, U, H1, H2, 0, TIO><<<grid, threads, 0, s>>>(W1, b1, W2, b2, W3, b3, st, h->lay, ar); break;
        case 1: nempc_small_kernel<X, U, H1, H2, 1, TIO><<<grid, threads, 0, s>>>(W1, b1, W2, b2, W3, b3, st, h->lay, ar); break;
        default: nempc_small_kernel<X, U, H1, H2, 2, TIO><<<grid, threads, 0, s>>>(W1, b1, W2, b2, W3, b3, st, h->lay, ar); break;
    }
    CU(h, cudaGetLastError());
    h->launches++;
    return NEMPC_OK;
}

template <typename TIO> static int launch_fast(nempc_handle* h, const EvalArgs<TIO>& ar, int mode, cudaStream_t s) {
    static const long long small_max = getenv("NEMPC_SMALL_MAX_STEPS") ? atoll(getenv("NEMPC_SMALL_MAX_STEPS")) : NEMPC_SMALL_MAX_STEPS;
    if (ar.nsteps <= small_max && h->desc.kernel == NEMPC_KERNEL_AUTO) {
        switch (h->fast_id) {
            case 0: return launch_small_shape<2, 1, 30, 30, TIO>(h, ar, mode, s);
            case 1: return launch_small_shape<2, 1, 32, 32, TIO>(h, ar, mode, s);
            case 2: return launch_small_shape<2, 1, 16, 16, TIO>(h, ar, mode, s);
        }
    }
    switch (h->fast_id) {
        case 0: return launch_fast_shape<2, 1, 30, 30, NEMPC_FAST_NCHUNK30, TIO>(h, ar, mode, s);
        case 1: return launch_fast_shape<2, 1, 32, 32, 2, TIO>(h, ar, mode, s);
        case 2: return launch_fast_shape<2, 1, 16, 16, 1, TIO>(h, ar, mode, s);
    }
    SET_ERR(h, "internal: bad fast_id");
    return NEMPC_EINVAL;
}

// tensor-core kernels: compiled in their own translation units (nempc_tc_tu.cu, nempc_wide_tu.cu; build.py compiles the units side by
// side) behind internal entry points -- 0, a cudaError_t, or -1 for an unknown shape
int nempc_tc_launch(int tc_id, int mode, int io_f64, const void* img, const float* cb, const float* wx, const StageTable<float>& st,
                    const NlpLayout& L, const void* ar, int tvp_dim, int p_dim, int sm_count, cudaStream_t s);
size_t nempc_wide_scratch_floats(int wide_id, int mode, int rk4);
int nempc_wide_launch(int wide_id, int mode, int rk4, int io_f64, const unsigned char* blob, const float* cb, const WideNet& net,
                      const StageTable<float>& st, const NlpLayout& L, const void* ar, float* scratch, int sm_count, cudaStream_t s);

template <typename TIO> static int launch_tc(nempc_handle* h, const EvalArgs<TIO>& ar, int mode, cudaStream_t s) {
    StageTable<float> st = make_stage_table<float>(h->desc.integrator == NEMPC_INTEG_RK4, h->desc.dt);
    EvalArgs<TIO> ax = ar;
    if (h->n_ext > 0) {
        int rc = bind_exogenous(h, ax, false);
        if (rc) return rc;
    }
    const int rc = nempc_tc_launch(h->tc_id, mode, sizeof(TIO) == 8, h->d_tcimg, h->d_tccb, h->d_tcwx, st, h->lay, &ax, h->desc.tvp_dim, h->desc.p_dim,
                                   h->sm_count, s);
    if (rc == -1) { SET_ERR(h, "internal: bad tc_id"); return NEMPC_EINVAL; }
    if (rc != 0) { SET_ERR(h, "nempc_tc_kernel launch: %s", cudaGetErrorString((cudaError_t)rc)); return NEMPC_ECUDA; }
    h->launches++;
    return NEMPC_OK;
}

template <typename TIO> static int launch_wide(nempc_handle* h, const EvalArgs<TIO>& ar, int mode, cudaStream_t s) {
    const int rk4 = h->desc.integrator == NEMPC_INTEG_RK4;
    const size_t fl = nempc_wide_scratch_floats(h->wide_id, mode, rk4);
    if (fl == 0) { SET_ERR(h, "internal: bad wide_id"); return NEMPC_EINVAL; }
    const size_t need = (size_t)h->sm_count * fl * sizeof(float);      // h_l, q_l (and the RK4 stage state) of one super-tile per CTA
    if (need > h->wide_scratch_bytes) {
        CU(h, cudaStreamSynchronize(s));
        cudaFree(h->wide_scratch); h->wide_scratch = nullptr; h->wide_scratch_bytes = 0;
        CU(h, cudaMalloc((void**)&h->wide_scratch, need));
        h->wide_scratch_bytes = need;
    }
    StageTable<float> st = make_stage_table<float>(rk4 != 0, h->desc.dt);
    const int rc = nempc_wide_launch(h->wide_id, mode, rk4, sizeof(TIO) == 8, h->d_wblob, h->d_wcb, h->wnet, st, h->lay, &ar, h->wide_scratch, h->sm_count, s);
    if (rc == -1) { SET_ERR(h, "internal: bad wide_id"); return NEMPC_EINVAL; }
    if (rc != 0) { SET_ERR(h, "nempc_wide_kernel launch: %s", cudaGetErrorString((cudaError_t)rc)); return NEMPC_ECUDA; }
    h->launches++;
    return NEMPC_OK;
}

// float64 tensor-core path (nempc_dmma_tu.cu)
#include "nempc_dmma.cuh"
size_t nempc_dmma_scratch_doubles(int shape_id, long long nsteps);
int nempc_dmma_launch(int shape_id, int hw, int mode, const DmmaNet& net, const StageTable<double>& st, const NlpLayout& L, const EvalArgs<double>& ar,
                      double* scratch, int sm_count, cudaStream_t s, int* launches);
#ifndef NEMPC_DMMA_MIN_STEPS
#define NEMPC_DMMA_MIN_STEPS 512         // below this many horizon steps the row tiles leave most SMs empty: the generic kernel takes over
#endif
template <typename TIO> static bool take_dmma(nempc_handle*, const EvalArgs<TIO>&) { return false; }
template <> bool take_dmma<double>(nempc_handle* h, const EvalArgs<double>& ar) {
    static const long long min_steps = getenv("NEMPC_DMMA_MIN_STEPS") ? atoll(getenv("NEMPC_DMMA_MIN_STEPS")) : NEMPC_DMMA_MIN_STEPS;
    return h->dmma_id >= 0 && (ar.nsteps >= min_steps || h->desc.kernel == NEMPC_KERNEL_TC);
}
template <typename TIO> static int run_dmma(nempc_handle*, const EvalArgs<TIO>&, int, cudaStream_t) { return NEMPC_EINVAL; }
template <> int run_dmma<double>(nempc_handle* h, const EvalArgs<double>& ar, int mode, cudaStream_t s) {
    const size_t need = nempc_dmma_scratch_doubles(h->dmma_id, ar.nsteps);
    if (need == 0) { SET_ERR(h, "internal: bad dmma_id"); return NEMPC_EINVAL; }
    if (need > h->dmma_scratch_doubles) {
        CU(h, cudaStreamSynchronize(s));
        cudaFree(h->dmma_scratch); h->dmma_scratch = nullptr; h->dmma_scratch_doubles = 0;
        CU(h, cudaMalloc((void**)&h->dmma_scratch, need * sizeof(double)));
        h->dmma_scratch_doubles = need;
    }
    DmmaNet net{};
    net.d = h->d; net.x = h->desc.x_dim; net.nhid = h->L - 1;
    for (int l = 0; l < h->L; ++l) { net.W[l] = (const double*)h->dW[l]; net.b[l] = (const double*)h->db[l]; }
    StageTable<double> st = make_stage_table<double>(h->desc.integrator == NEMPC_INTEG_RK4, h->desc.dt);
    int launches = 0;
    const int rc = nempc_dmma_launch(h->dmma_id, h->desc.widths[0], mode, net, st, h->lay, ar, h->dmma_scratch, h->sm_count, s, &launches);
    h->launches += launches;
    if (rc == -1) { SET_ERR(h, "internal: bad dmma shape"); return NEMPC_EINVAL; }
    if (rc != 0) { SET_ERR(h, "nempc_dmma kernels: %s", cudaGetErrorString((cudaError_t)rc)); return NEMPC_ECUDA; }
    return NEMPC_OK;
}

template <typename TIO>
static int eval_t(nempc_handle* h, int64_t B, const void* z, const void* x0, const void* lambda, const void* obj_factor,
                  double sigma, void* resid, void* jac, void* hes, void* obj, void* grad, cudaStream_t s) {
    if (resid || jac || hes) {
        EvalArgs<TIO> ar{};
        ar.z = (const TIO*)z; ar.x0 = (const TIO*)x0; ar.lam = (const TIO*)lambda; ar.sigma = (const TIO*)obj_factor;
        ar.sigma_scalar = sigma; ar.quad = h->has_objective ? h->dquad : nullptr;
        ar.resid = (TIO*)resid; ar.jac = (TIO*)jac; ar.hes = (TIO*)hes;
        ar.nsteps = (long long)B * h->desc.horizon;
        ar.gate = h->gate; ar.gate_value = h->gate_value;
        const int mode = hes ? 2 : (jac ? 1 : 0);
        ar.flags = (mode >= 1 ? NEMPC_WANT_JAC : 0) | (mode >= 2 ? NEMPC_WANT_HES : 0) |
                   (h->desc.integrator == NEMPC_INTEG_UNITY ? NEMPC_UNITY : 0);
        int rc = take_fast64<TIO>(h, ar) ? run_fast64<TIO>(h, ar, mode, s) : take_dmma<TIO>(h, ar) ? run_dmma<TIO>(h, ar, mode, s) : h->use_fast ? launch_fast<TIO>(h, ar, mode, s)
                            : (h->use_tc ? ((h->wide_hes && mode == 2) ? launch_wide<TIO>(h, ar, mode, s) : launch_tc<TIO>(h, ar, mode, s)) : (h->use_wide ? launch_wide<TIO>(h, ar, mode, s) : launch_generic<TIO>(h, ar, false, s)));
        if (rc) return rc;
    }
    if (obj || grad) {
        if (!h->has_objective) { SET_ERR(h, "obj/grad requested but nempc_set_objective was never called"); return NEMPC_ESTATE; }
        const int threads = 256;
        const long long warps_needed = B;
        const unsigned grid = (unsigned)std::max(1LL, std::min((warps_needed * 32 + threads - 1) / threads, (long long)h->sm_count * 16));
        nempc_objective_kernel<TIO><<<grid, threads, 0, s>>>((const TIO*)z, h->dlin, h->dquad, h->dref, (TIO*)obj, (TIO*)grad, h->lay.n, B, h->gate, h->gate_value);
        CU(h, cudaGetLastError());
        h->launches++;
    }
    return NEMPC_OK;
}

extern "C" int nempc_eval(nempc_handle* h, int64_t B, const void* z, const void* x0, const void* lambda,
                          const void* obj_factor, double sigma, void* resid, void* jac, void* hes, void* obj, void* grad,
                          void* stream) {
    int rc = ready(h);
    if (rc) return rc;
    if (B == 0) return NEMPC_OK;
    if (B < 0 || !z || !x0) { SET_ERR(h, "nempc_eval: bad argument"); return NEMPC_EINVAL; }
    if (hes && !lambda) { SET_ERR(h, "nempc_eval: hes_vals needs lambda"); return NEMPC_EINVAL; }
    CU(h, cudaSetDevice(h->desc.device));
    cudaStream_t s = (cudaStream_t)stream;
    return h->desc.io_dtype == NEMPC_F64 ? eval_t<double>(h, B, z, x0, lambda, obj_factor, sigma, resid, jac, hes, obj, grad, s)
                                         : eval_t<float>(h, B, z, x0, lambda, obj_factor, sigma, resid, jac, hes, obj, grad, s);
}

static int stage_reserve(nempc_handle* h, int i, size_t bytes) {
    if (bytes <= h->st_cap[i]) return NEMPC_OK;
    cudaFree(h->st_buf[i]); h->st_buf[i] = nullptr; h->st_cap[i] = 0;
    const size_t cap = bytes + bytes / 4;
    CU(h, cudaMalloc(&h->st_buf[i], cap));
    h->st_cap[i] = cap;
    return NEMPC_OK;
}

// the chunk pipeline of nempc_eval_host: the batch is cut along B and chunk c runs H2D -> kernels -> D2H on stream c % 3, so the
// upload of chunk c+1, the kernels of chunk c and the download of chunk c-1 overlap (separate copy engines; PCIe is full
// duplex).  Small batches and the global-workspace fallback use a single chunk on pipe[0].
static int eval_host_issue(nempc_handle* h, int64_t B, const void* const in[4], void* const outp[5], double sigma, const size_t sz[10]) {
    const size_t es = dsize(h->desc.io_dtype);
    const NlpLayout& L = h->lay;
    const size_t per_problem = ((size_t)2 * L.n + L.x + 2 * L.m + L.nnz_jac + L.nnz_hes + 2) * es;
    int64_t chunk = B;
    if (!h->global_ws && B >= 512) {
        static const int chunk_mb = getenv("NEMPC_HOST_CHUNK_MB") ? std::max(1, atoi(getenv("NEMPC_HOST_CHUNK_MB"))) : 16;
        chunk = std::max<int64_t>(256, (int64_t)(((size_t)chunk_mb << 20) / per_problem));     // ~16 MB of traffic per chunk: copies below ~4 MB lose PCIe efficiency, worst in full duplex (tools/pcie_probe.py; sweep in DESIGN.md 6)
        static const int min_chunks = getenv("NEMPC_HOST_MIN_CHUNKS") ? std::max(1, atoi(getenv("NEMPC_HOST_MIN_CHUNKS"))) : 4;
        chunk = std::min<int64_t>(chunk, (B + min_chunks - 1) / min_chunks);              // at least 4 chunks in flight
    }
    const size_t in_w[4] = {(size_t)L.n * es, (size_t)L.x * es, (size_t)L.m * es, es};
    const size_t out_w[5] = {(size_t)L.m * es, (size_t)L.nnz_jac * es, (size_t)L.nnz_hes * es, es, (size_t)L.n * es};
    int ci = 0;
    for (int64_t c0 = 0; c0 < B; c0 += chunk, ++ci) {
        const int64_t nb = std::min(chunk, B - c0);
        cudaStream_t s = h->pipe[ci % 3];
        void* din[4]; void* dout[5];
        for (int i = 0; i < 4; ++i) {
            din[i] = sz[i] ? (char*)h->st_buf[i] + c0 * in_w[i] : nullptr;
            if (sz[i]) CU(h, cudaMemcpyAsync(din[i], (const char*)in[i] + c0 * in_w[i], nb * in_w[i], cudaMemcpyHostToDevice, s));
        }
        for (int i = 0; i < 5; ++i) dout[i] = sz[4 + i] ? (char*)h->st_buf[4 + i] + c0 * out_w[i] : nullptr;
        // The width-256 / adjoint-form kernels (per-CTA scratch indexed by blockIdx) and the float64 tensor-core path (network outputs + stage
        // state of one chunk of steps) keep scratch in the HANDLE: the tail of chunk c's kernels must not overlap the head of chunk c+1's on
        // another stream.  Their kernels are chained with an event; the copies on either side still overlap them.
        const bool shared_scratch = h->use_wide || h->wide_hes || h->dmma_id >= 0;
        if (shared_scratch && ci > 0) CU(h, cudaStreamWaitEvent(s, h->ev_kern[(ci - 1) % 3], 0));
        h->exo_base = c0;
        int rc = nempc_eval(h, nb, din[0], din[1], din[2], din[3], sigma, dout[0], dout[1], dout[2], dout[3], dout[4], (void*)s);
        h->exo_base = 0;
        if (rc) return rc;
        if (shared_scratch) {
            if (!h->ev_kern[ci % 3]) CU(h, cudaEventCreateWithFlags(&h->ev_kern[ci % 3], cudaEventDisableTiming));
            CU(h, cudaEventRecord(h->ev_kern[ci % 3], s));
        }
        for (int i = 0; i < 5; ++i)
            if (sz[4 + i]) CU(h, cudaMemcpyAsync((char*)outp[i] + c0 * out_w[i], dout[i], nb * out_w[i], cudaMemcpyDeviceToHost, s));
    }
    return NEMPC_OK;
}

static bool is_pinned_host(const void* p) {
    if (!p) return true;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
// device alias of a page-locked, mapped host buffer (nullptr for pageable memory); NULL stays NULL
static bool mapped_alias(const void* p, void** dev) {
    *dev = nullptr;
    if (!p) return true;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
    *dev = a.devicePointer;
    return true;
}

extern "C" int nempc_eval_host(nempc_handle* h, int64_t B, const void* z, const void* x0, const void* lambda,
                               const void* obj_factor, double sigma, void* resid, void* jac, void* hes, void* obj,
                               void* grad) {
    int rc = ready(h);
    if (rc) return rc;
    if (B == 0) return NEMPC_OK;
    if (B < 0 || !z || !x0) { SET_ERR(h, "nempc_eval_host: bad argument"); return NEMPC_EINVAL; }
    if (hes && !lambda) { SET_ERR(h, "nempc_eval_host: hes_vals needs lambda"); return NEMPC_EINVAL; }
    if ((obj || grad) && !h->has_objective) { SET_ERR(h, "obj/grad requested but nempc_set_objective was never called"); return NEMPC_ESTATE; }
    CU(h, cudaSetDevice(h->desc.device));
    const size_t es = dsize(h->desc.io_dtype);
    const NlpLayout& L = h->lay;
    const size_t sz[10] = {(size_t)B * L.n * es, (size_t)B * L.x * es, lambda ? (size_t)B * L.m * es : 0, obj_factor ? (size_t)B * es : 0,
                           resid ? (size_t)B * L.m * es : 0, jac ? (size_t)B * L.nnz_jac * es : 0, hes ? (size_t)B * L.nnz_hes * es : 0,
                           obj ? (size_t)B * es : 0, grad ? (size_t)B * L.n * es : 0, 0};
    const void* in[4] = {z, x0, lambda, obj_factor};
    void* outp[5] = {resid, jac, hes, obj, grad};

    // ---- zero-copy: a single solver callback moves a few KB.  With page-locked (mapped) buffers the kernels read the iterate and
    // write the residual / Jacobian / Hessian values straight through PCIe -- no staging buffers, no copy-engine round trips, one
    // stream synchronise (the "pinned zero-copy handoff" of the callback glue, optimizer/ipopt.py:30-96).
    {
        static const size_t zc_limit = getenv("NEMPC_ZEROCOPY_BYTES") ? (size_t)atoll(getenv("NEMPC_ZEROCOPY_BYTES")) : (size_t)(256u << 10);
        size_t total = 0;
        for (int i = 0; i < 9; ++i) total += sz[i];
        if (total <= zc_limit && !h->global_ws) {
            void* din[4]; void* dout[5];
            bool ok = true;
            for (int i = 0; i < 4 && ok; ++i) ok = mapped_alias(in[i], &din[i]);
            for (int i = 0; i < 5 && ok; ++i) ok = mapped_alias(outp[i], &dout[i]);
            if (ok) {
                rc = nempc_eval(h, B, din[0], din[1], din[2], din[3], sigma, dout[0], dout[1], dout[2], dout[3], dout[4], (void*)h->stream);
                if (rc) return rc;
                CU(h, cudaStreamSynchronize(h->stream));
                return NEMPC_OK;
            }
        }
    }
    // ---- replay: a solver callback (and the bench) calls this with the same pinned buffers over and over.  The ~50 copies and
    // launches of the pipeline then cost more CPU time to submit than PCIe needs to move the data, so the second call with an
    // unchanged argument set captures the pipeline as a CUDA graph and later calls submit that graph (one API call).
    static const bool graph_off = getenv("NEMPC_HOST_GRAPH") && !atoi(getenv("NEMPC_HOST_GRAPH"));
    if (graph_off) h->hk_disabled = true;
    nempc_handle::HostKey key{};
    key.B = B; key.sigma = sigma;
    for (int i = 0; i < 4; ++i) key.in[i] = in[i];
    for (int i = 0; i < 5; ++i) key.out[i] = outp[i];
    if (!h->hk_disabled && h->hk_exec && key == h->hk) {
        CU(h, cudaGraphLaunch(h->hk_exec, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        h->launches += h->hk_launches;
        return NEMPC_OK;
    }
    if (!(key == h->hk)) {
        if (h->hk_exec) { cudaGraphExecDestroy(h->hk_exec); h->hk_exec = nullptr; }
        h->hk = key; h->hk_seen = 0;
    }
    for (int i = 0; i < 9; ++i) if (sz[i]) { rc = stage_reserve(h, i, sz[i]); if (rc) return rc; }
    bool capture = false;
    if (!h->hk_disabled && ++h->hk_seen >= 2 && !h->global_ws) {
        capture = true;
        for (int i = 0; i < 4 && capture; ++i) capture = is_pinned_host(in[i]);
        for (int i = 0; i < 5 && capture; ++i) capture = is_pinned_host(outp[i]);
    }
    if (capture) {
        if (!h->ev_fork) {
            CU(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
            for (int i = 0; i < 3; ++i) CU(h, cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
        }
        const long long l0 = h->launches;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            ok = cudaEventRecord(h->ev_fork, h->stream) == cudaSuccess;
            for (int i = 0; i < 3 && ok; ++i) ok = cudaStreamWaitEvent(h->pipe[i], h->ev_fork, 0) == cudaSuccess;
            if (ok) ok = eval_host_issue(h, B, in, outp, sigma, sz) == NEMPC_OK;
            for (int i = 0; i < 3 && ok; ++i) ok = cudaEventRecord(h->ev_join[i], h->pipe[i]) == cudaSuccess && cudaStreamWaitEvent(h->stream, h->ev_join[i], 0) == cudaSuccess;
            const bool ended = cudaStreamEndCapture(h->stream, &graph) == cudaSuccess;
            ok = ok && ended && graph;
        }
        h->hk_launches = h->launches - l0;
        h->launches = l0;
        if (ok) ok = cudaGraphInstantiate(&h->hk_exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        if (ok) {
            CU(h, cudaGraphLaunch(h->hk_exec, h->stream));
            CU(h, cudaStreamSynchronize(h->stream));
            h->launches += h->hk_launches;
            return NEMPC_OK;
        }
        cudaGetLastError();                          // capture is an optimisation: fall back to direct submission for good
        if (h->hk_exec) { cudaGraphExecDestroy(h->hk_exec); h->hk_exec = nullptr; }
        h->hk_disabled = true;
    }
    rc = eval_host_issue(h, B, in, outp, sigma, sz);
    if (rc) return rc;
    for (int i = 0; i < 3; ++i) CU(h, cudaStreamSynchronize(h->pipe[i]));
    return NEMPC_OK;
}

template <typename TIO>
static int blocks_t(nempc_handle* h, long long N, const void* z, const void* x0, void* pred, void* AB, void* Hblk, bool model_mode, cudaStream_t s) {
    EvalArgs<TIO> ar{};
    ar.z = (const TIO*)z; ar.x0 = (const TIO*)x0; ar.pred = (TIO*)pred; ar.AB = (TIO*)AB; ar.Hblk = (TIO*)Hblk;
    ar.nsteps = N;
    ar.flags = (model_mode ? NEMPC_MODE_MODEL : NEMPC_MODE_BLOCKS) | (AB || Hblk ? NEMPC_WANT_JAC : 0) | (Hblk ? NEMPC_WANT_HES : 0) |
               (h->desc.integrator == NEMPC_INTEG_UNITY ? NEMPC_UNITY : 0);
    return launch_generic<TIO>(h, ar, model_mode, s);
}

extern "C" int nempc_eval_blocks(nempc_handle* h, int64_t B, const void* z, const void* x0, void* pred, void* AB, void* Hblk, void* stream) {
    int rc = ready(h);
    if (rc) return rc;
    if (B < 0 || !z || !x0) { SET_ERR(h, "nempc_eval_blocks: bad argument"); return NEMPC_EINVAL; }
    if (B == 0) return NEMPC_OK;
    CU(h, cudaSetDevice(h->desc.device));
    const long long N = (long long)B * h->desc.horizon;
    return h->desc.io_dtype == NEMPC_F64 ? blocks_t<double>(h, N, z, x0, pred, AB, Hblk, false, (cudaStream_t)stream)
                                         : blocks_t<float>(h, N, z, x0, pred, AB, Hblk, false, (cudaStream_t)stream);
}

extern "C" int nempc_model_eval(nempc_handle* h, int64_t N, const void* zin, void* f, void* jac, void* hes, void* stream) {
    int rc = ready(h);
    if (rc) return rc;
    if (N < 0 || !zin) { SET_ERR(h, "nempc_model_eval: bad argument"); return NEMPC_EINVAL; }
    if (N == 0) return NEMPC_OK;
    CU(h, cudaSetDevice(h->desc.device));
    return h->desc.io_dtype == NEMPC_F64 ? blocks_t<double>(h, N, zin, nullptr, f, jac, hes, true, (cudaStream_t)stream)
                                         : blocks_t<float>(h, N, zin, nullptr, f, jac, hes, true, (cudaStream_t)stream);
}

extern "C" int nempc_objective_eval(int32_t io_dtype, int64_t B, int64_t n, const void* z, const double* lin, const double* quad,
                                    const double* ref, void* obj, void* grad, void* stream) {
    if (B < 0 || n < 1 || !z || !lin || !quad || !ref || (io_dtype != NEMPC_F32 && io_dtype != NEMPC_F64)) {
        SET_ERR((nempc_handle*)nullptr, "nempc_objective_eval: bad argument");
        return NEMPC_EINVAL;
    }
    if (B == 0 || (!obj && !grad)) return NEMPC_OK;
    const int threads = 256;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((B * 32 + threads - 1) / threads, 148 * 16));
    cudaStream_t s = (cudaStream_t)stream;
    if (io_dtype == NEMPC_F64) nempc_objective_kernel<double><<<grid, threads, 0, s>>>((const double*)z, lin, quad, ref, (double*)obj, (double*)grad, (int)n, B);
    else nempc_objective_kernel<float><<<grid, threads, 0, s>>>((const float*)z, lin, quad, ref, (float*)obj, (float*)grad, (int)n, B);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { SET_ERR((nempc_handle*)nullptr, "objective kernel launch: %s", cudaGetErrorString(e)); return NEMPC_ECUDA; }
    return NEMPC_OK;
}

// ---- general quadratic objective: value / gradient, and the merge of its constant Hessian into the Lagrangian-Hessian values ---------
extern "C" int nempc_quadform_eval(int32_t io_dtype, int64_t B, int64_t n, const void* z, const int32_t* p_ptr, const int32_t* p_idx,
                                   const double* p_val, const double* q, double c, void* obj, void* grad, void* stream) {
    if (B < 0 || n < 1 || !z || !p_ptr || !p_idx || !p_val || !q || (io_dtype != NEMPC_F32 && io_dtype != NEMPC_F64)) {
        SET_ERR((nempc_handle*)nullptr, "nempc_quadform_eval: bad argument");
        return NEMPC_EINVAL;
    }
    if (B == 0 || (!obj && !grad)) return NEMPC_OK;
    const int threads = 256;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((B * 32 + threads - 1) / threads, 148 * 16));
    cudaStream_t s = (cudaStream_t)stream;
    if (io_dtype == NEMPC_F64) nempc_quadform_kernel<double><<<grid, threads, 0, s>>>((const double*)z, p_ptr, p_idx, p_val, q, c, (double*)obj, (double*)grad, (int)n, B);
    else nempc_quadform_kernel<float><<<grid, threads, 0, s>>>((const float*)z, p_ptr, p_idx, p_val, q, c, (float*)obj, (float*)grad, (int)n, B);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { SET_ERR((nempc_handle*)nullptr, "quadratic-form kernel launch: %s", cudaGetErrorString(e)); return NEMPC_ECUDA; }
    return NEMPC_OK;
}

extern "C" int nempc_hessian_merge(int32_t io_dtype, int64_t B, int64_t nnz_kern, int64_t nnz_out, const void* kern_vals,
                                   const int32_t* src_slot, const double* p_val, const void* obj_factor, double obj_factor_scalar,
                                   void* out_vals, void* stream) {
    if (B < 0 || nnz_out < 1 || nnz_kern < 0 || !src_slot || !p_val || !out_vals || (nnz_kern > 0 && !kern_vals) ||
        (io_dtype != NEMPC_F32 && io_dtype != NEMPC_F64)) {
        SET_ERR((nempc_handle*)nullptr, "nempc_hessian_merge: bad argument");
        return NEMPC_EINVAL;
    }
    if (B == 0) return NEMPC_OK;
    const int threads = 256;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((B * nnz_out + threads - 1) / threads, 148 * 16));
    cudaStream_t s = (cudaStream_t)stream;
    if (io_dtype == NEMPC_F64) nempc_hessian_merge_kernel<double><<<grid, threads, 0, s>>>((const double*)kern_vals, src_slot, p_val, (const double*)obj_factor, obj_factor_scalar, (double*)out_vals, nnz_kern, nnz_out, B);
    else nempc_hessian_merge_kernel<float><<<grid, threads, 0, s>>>((const float*)kern_vals, src_slot, p_val, (const float*)obj_factor, obj_factor_scalar, (float*)out_vals, nnz_kern, nnz_out, B);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { SET_ERR((nempc_handle*)nullptr, "Hessian merge kernel launch: %s", cudaGetErrorString(e)); return NEMPC_ECUDA; }
    return NEMPC_OK;
}

// ---- rolling-window (NARX) models: window gather + banded sparse assembly (nempc_rolling.cuh) -------------------------------------
extern "C" int nempc_rolling_gather(int32_t io_dtype, int64_t B, int32_t n, int32_t naux, int32_t rows, const int32_t* gidx,
                                    const void* z, const void* aux, void* zin, void* stream) {
    if (B < 0 || n < 1 || naux < 1 || rows < 1 || !gidx || !z || !aux || !zin || (io_dtype != NEMPC_F32 && io_dtype != NEMPC_F64)) {
        SET_ERR((nempc_handle*)nullptr, "nempc_rolling_gather: bad argument");
        return NEMPC_EINVAL;
    }
    if (B == 0) return NEMPC_OK;
    const int threads = 256;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((B * rows + threads - 1) / threads, 148 * 16));
    cudaStream_t s = (cudaStream_t)stream;
    if (io_dtype == NEMPC_F64) nempc_rolling_gather_kernel<double><<<grid, threads, 0, s>>>((const double*)z, (const double*)aux, gidx, (double*)zin, n, naux, rows, B);
    else nempc_rolling_gather_kernel<float><<<grid, threads, 0, s>>>((const float*)z, (const float*)aux, gidx, (float*)zin, n, naux, rows, B);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { SET_ERR((nempc_handle*)nullptr, "rolling gather kernel launch: %s", cudaGetErrorString(e)); return NEMPC_ECUDA; }
    return NEMPC_OK;
}

extern "C" int nempc_rolling_assemble(int32_t io_dtype, int64_t B, int32_t H, int32_t x, int32_t dw, int32_t n, int32_t naux,
                                      int64_t nnz_jac, int64_t nnz_hes, const int32_t* resid_base, const int32_t* jac_src,
                                      const double* jac_add, const int32_t* hes_ptr, const int32_t* hes_src, const double* hes_obj,
                                      const void* z, const void* aux, const void* f, const void* J, const void* Hs, const void* lambda,
                                      const void* obj_factor, double obj_factor_scalar, void* resid, void* jac_vals, void* hes_vals,
                                      void* stream) {
    if (B < 0 || H < 1 || x < 1 || dw < 1 || n < 1 || !z || !aux || (io_dtype != NEMPC_F32 && io_dtype != NEMPC_F64)) {
        SET_ERR((nempc_handle*)nullptr, "nempc_rolling_assemble: bad argument");
        return NEMPC_EINVAL;
    }
    if ((resid && (!f || !resid_base)) || (jac_vals && (!J || !jac_src || !jac_add)) || (hes_vals && (!Hs || !lambda || !hes_ptr || !hes_src || !hes_obj))) {
        SET_ERR((nempc_handle*)nullptr, "nempc_rolling_assemble: an output was requested without its inputs / tables");
        return NEMPC_EINVAL;
    }
    if (B == 0 || (!resid && !jac_vals && !hes_vals)) return NEMPC_OK;
    const RollingTables tb{resid_base, jac_src, jac_add, hes_ptr, hes_src, hes_obj};
    const long long per = (resid ? (long long)H * x : 0) + (jac_vals ? nnz_jac : 0) + (hes_vals ? nnz_hes : 0);
    const int threads = 256;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((B * per + threads - 1) / threads, 148 * 16));
    cudaStream_t s = (cudaStream_t)stream;
    if (io_dtype == NEMPC_F64)
        nempc_rolling_assemble_kernel<double><<<grid, threads, 0, s>>>(tb, (const double*)z, (const double*)aux, (const double*)f, (const double*)J, (const double*)Hs,
                                                                     (const double*)lambda, (const double*)obj_factor, obj_factor_scalar, (double*)resid,
                                                                     (double*)jac_vals, (double*)hes_vals, H, x, dw, n, naux, nnz_jac, nnz_hes, B);
    else
        nempc_rolling_assemble_kernel<float><<<grid, threads, 0, s>>>(tb, (const float*)z, (const float*)aux, (const float*)f, (const float*)J, (const float*)Hs,
                                                                    (const float*)lambda, (const float*)obj_factor, obj_factor_scalar, (float*)resid,
                                                                    (float*)jac_vals, (float*)hes_vals, H, x, dw, n, naux, nnz_jac, nnz_hes, B);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { SET_ERR((nempc_handle*)nullptr, "rolling assemble kernel launch: %s", cudaGetErrorString(e)); return NEMPC_ECUDA; }
    return NEMPC_OK;
}

// ---- batched interior-point solver ---------------------------------------------------------------------------------------
extern "C" int nempc_solver_defaults(nempc_solver_opts* o) {
    if (!o) return NEMPC_EINVAL;
    const SolverOpts d = solver_defaults();
    o->max_iter = d.max_iter; o->max_backtrack = d.max_backtrack; o->tol = d.tol; o->mu_init = d.mu_init; o->mu_min = d.mu_min;
    o->kappa_eps = d.kappa_eps; o->kappa_mu = d.kappa_mu; o->theta_mu = d.theta_mu; o->tau_min = d.tau_min;
    o->bound_push = d.bound_push; o->eta = d.eta; o->reg_init = d.reg_init; o->reg_max = d.reg_max;
    return NEMPC_OK;
}

extern "C" int nempc_solve(nempc_handle* h, int64_t B, const void* x0, const double* lb, const double* ub, void* z, int32_t use_init,
                           void* lambda, int32_t* status, int32_t* iterations, double* kkt_error, const nempc_solver_opts* opts,
                           int32_t* outer_iterations, void* stream) {
    int rc = ready(h);
    if (rc) return rc;
    if (outer_iterations) *outer_iterations = 0;
    if (B == 0) return NEMPC_OK;
    if (B < 0 || !x0 || !lb || !ub || !z || !status) { SET_ERR(h, "nempc_solve: bad argument"); return NEMPC_EINVAL; }
    if (h->desc.io_dtype != NEMPC_F64) { SET_ERR(h, "nempc_solve needs io_dtype = F64"); return NEMPC_EUNSUPPORTED; }
    if (!h->has_objective) { SET_ERR(h, "nempc_solve: nempc_set_objective was never called"); return NEMPC_ESTATE; }
    CU(h, cudaSetDevice(h->desc.device));
    cudaStream_t s = (cudaStream_t)stream;
    const NlpLayout& L = h->lay;
    SolverOpts o = solver_defaults();
    if (opts) {
        o.max_iter = opts->max_iter; o.max_backtrack = opts->max_backtrack; o.tol = opts->tol; o.mu_init = opts->mu_init; o.mu_min = opts->mu_min;
        o.kappa_eps = opts->kappa_eps; o.kappa_mu = opts->kappa_mu; o.theta_mu = opts->theta_mu; o.tau_min = opts->tau_min;
        o.bound_push = opts->bound_push; o.eta = opts->eta; o.reg_init = opts->reg_init; o.reg_max = opts->reg_max;
    }
    // ---- workspace -------------------------------------------------------------------------------------------------------
    // The iterate and x0 live in the handle's own workspace (copied in / out around the loop): every pointer the loop's kernels see is then
    // a function of (workspace, B) alone, so the captured loop below can be replayed for any caller buffers.
    const size_t n = L.n, m = L.m, nj = (size_t)L.nnz_jac, nh = (size_t)L.nnz_hes, Hux = (size_t)L.H * L.u * L.x, Hu = (size_t)L.H * L.u;
    const size_t per = 10 * n + 4 * m + nj + nh + Hux + Hu + L.x + 2 /*obj, objt*/ + 7 /*scalars*/ + 2 /*3 ints, padded*/;
    const size_t need = per * (size_t)B * sizeof(double);
    if (need > h->sv_cap) {
        CU(h, cudaStreamSynchronize(s));
        drop_solve_graph(h);
        cudaFree(h->sv_buf); h->sv_buf = nullptr; h->sv_cap = 0;
        CU(h, cudaMalloc(&h->sv_buf, need));
        h->sv_cap = need;
    }
    if (!h->sv_lb) {
        CU(h, cudaMalloc(&h->sv_lb, n * sizeof(double))); CU(h, cudaMalloc(&h->sv_ub, n * sizeof(double)));
        CU(h, cudaMalloc(&h->sv_counts, NEMPC_SV_COUNT * sizeof(int))); CU(h, cudaMallocHost(&h->sv_counts_host, NEMPC_SV_COUNT * sizeof(int)));
    }
    CU(h, cudaMemcpyAsync(h->sv_lb, lb, n * sizeof(double), cudaMemcpyHostToDevice, s));
    CU(h, cudaMemcpyAsync(h->sv_ub, ub, n * sizeof(double), cudaMemcpyHostToDevice, s));
    double* p = (double*)h->sv_buf;
    auto take = [&](size_t cnt) { double* r = p; p += cnt * (size_t)B; return r; };
    SolverWs w{};
    double* x0w = take(L.x);
    w.x0 = x0w; w.lb = h->sv_lb; w.ub = h->sv_ub;
    w.z = take(n); w.lam = take(m); w.zL = take(n); w.zU = take(n);
    w.dz = take(n); w.lamn = take(m); w.dzL = take(n); w.dzU = take(n);
    w.grad = take(n); w.resid = take(m); w.jac = take(nj); w.hes = take(nh); w.obj = take(1);
    w.zt = take(n); w.residt = take(m); w.objt = take(1);
    w.K = take(Hux); w.kf = take(Hu);
    w.mu = take(1); w.nu = take(1); w.alpha = take(1); w.alphaD = take(1); w.phi0 = take(1); w.dphi = take(1); w.err = take(1);
    int* ip = (int*)take(2);
    w.status = ip; w.iters = ip + B; w.accepted = ip + 2 * B;
    CU(h, cudaMemcpyAsync(x0w, x0, (size_t)B * L.x * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (use_init) CU(h, cudaMemcpyAsync(w.z, z, (size_t)B * n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    const int threads = 128;
    const unsigned grid = (unsigned)((B + threads - 1) / threads);
    // KKT kernel instance: exact (x_dim, u_dim) instantiations for the common small systems, run-time dimensions otherwise;
    // staged = one warp per problem with the problem's rows copied to shared memory first (used whenever a problem fits;
    // NEMPC_KKT_STAGED=0 forces the thread-per-problem kernel)
    typedef void (*kkt_fn)(const NlpLayout, const SolverWs, const SolverOpts, long long, int*);
    kkt_fn kkt_plain, kkt_staged;
#define NEMPC_KKT_PICK(XM_, UM_, XC_, UC_) do { kkt_plain = nempc_ipm_kkt_kernel<XM_, UM_, XC_, UC_>; kkt_staged = nempc_ipm_kkt_staged_kernel<XM_, UM_, XC_, UC_>; } while (0)
    if (L.x == 2 && L.u == 1) NEMPC_KKT_PICK(2, 1, 2, 1);
    else if (L.x == 3 && L.u == 1) NEMPC_KKT_PICK(3, 1, 3, 1);
    else if (L.x == 4 && L.u == 1) NEMPC_KKT_PICK(4, 1, 4, 1);
    else if (L.x == 4 && L.u == 2) NEMPC_KKT_PICK(4, 2, 4, 2);
    else if (L.x <= 4 && L.u <= 2) NEMPC_KKT_PICK(4, 2, 0, 0);
    else NEMPC_KKT_PICK(NEMPC_SOLVER_XM, NEMPC_SOLVER_UM, 0, 0);
#undef NEMPC_KKT_PICK
    // warps (= problems) per CTA: the count that packs the most problems per SM into shared memory (bounds are per CTA)
    const size_t per_bytes = (4 * n + 2 * m + nj + nh + Hux + Hu) * sizeof(double), bnd_bytes = 2 * n * sizeof(double);
    // at most 192 KB of the SM for staging: the Riccati body keeps its small matrices in local memory and needs the L1 that is left
    static const size_t smem_kb = getenv("NEMPC_KKT_SMEM_KB") ? (size_t)std::max(16, std::min(227, atoi(getenv("NEMPC_KKT_SMEM_KB")))) : 192;
    const size_t smem_sm = smem_kb * 1024, smem_cta_max = std::min<size_t>(smem_sm - 1024, (size_t)std::max(0, h->max_smem_optin - 1024));
    int staged_wpb = 0; size_t staged_smem = 0; long long best = 0;
    for (int wpb = 1; wpb <= 16; ++wpb) {
        const size_t need = bnd_bytes + wpb * per_bytes;
        if (need > smem_cta_max) break;
        const long long ctas = std::min<long long>(32, (long long)(smem_sm / (need + 1024))), per_sm = std::min<long long>(ctas * wpb, 64);
        if (per_sm > best) { best = per_sm; staged_wpb = wpb; staged_smem = need; }
    }
    const size_t ls_smem = (size_t)8 * (2 * n + m) * sizeof(double);           // line-search kernel: 8 warps (problems) per CTA
    const char* force = getenv("NEMPC_KKT_STAGED");
    if (force && force[0] == '0') staged_smem = 0;
    if (staged_smem > 48 * 1024) CU(h, cudaFuncSetAttribute(kkt_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem));
    int* const cnt = h->sv_counts;
    // ---- the three segments of one outer iteration, issued on stream `cs` (directly, or while it is being captured) -----------------------
    // A: evaluation at the iterate -> KKT step;  T: one line-search trial;  U: take the step
    auto seg_kkt = [&](cudaStream_t cs) -> int {
        h->gate = w.status; h->gate_value = NEMPC_ST_RUNNING;           // converged / failed problems are not re-evaluated
        const int r = nempc_eval(h, B, w.z, w.x0, w.lam, nullptr, 1.0, w.resid, w.jac, w.hes, w.obj, w.grad, (void*)cs);
        h->gate = nullptr;
        if (r) return r;
        CU(h, cudaMemsetAsync(cnt, 0, 2 * sizeof(int), cs));
        if (staged_smem) kkt_staged<<<(unsigned)((B + staged_wpb - 1) / staged_wpb), 32 * staged_wpb, staged_smem, cs>>>(L, w, o, B, cnt);
        else kkt_plain<<<grid, threads, 0, cs>>>(L, w, o, B, cnt);
        CU(h, cudaGetLastError()); h->launches++;
        return NEMPC_OK;
    };
    auto seg_trial = [&](cudaStream_t cs) -> int {
        h->gate = w.accepted; h->gate_value = 0;               // only the problems whose step is not accepted yet need the trial point
        const int r = nempc_eval(h, B, w.zt, w.x0, nullptr, nullptr, 1.0, w.residt, nullptr, nullptr, w.objt, nullptr, (void*)cs);
        h->gate = nullptr;
        if (r) return r;
        CU(h, cudaMemsetAsync(cnt + 1, 0, sizeof(int), cs));
        if (ls_smem <= 48 * 1024) nempc_ipm_linesearch_warp_kernel<<<(unsigned)((B + 7) / 8), 256, ls_smem, cs>>>(L, w, o, B, cnt);
        else nempc_ipm_linesearch_kernel<<<grid, threads, 0, cs>>>(L, w, o, B, cnt);
        CU(h, cudaGetLastError()); h->launches++;
        return NEMPC_OK;
    };
    auto seg_update = [&](cudaStream_t cs, const int* gate_counts) -> int {
        const long long tot = (long long)B * L.n;
        nempc_ipm_update_flat_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, cs>>>(L, w, B, gate_counts);
        nempc_ipm_count_iter_kernel<<<grid, threads, 0, cs>>>(w, B, gate_counts, cnt);
        CU(h, cudaGetLastError()); h->launches += 2;
        return NEMPC_OK;
    };
    CU(h, cudaMemsetAsync(cnt, 0, NEMPC_SV_COUNT * sizeof(int), s));
    nempc_ipm_init_kernel<<<grid, threads, 0, s>>>(L, w, o, B, use_init);
    CU(h, cudaGetLastError()); h->launches++;
    int it = 0;
    // ---- device-side loop: both host loops below as WHILE nodes of ONE CUDA graph (no host synchronisation and no launch latency between
    // the ~10 kernels of an iteration: what a single problem's solve time consists of).  NEMPC_SOLVE_GRAPH: 0 = never, 1 (default) = from
    // the second solve with the same (batch, options, workspace) on, 2 = from the first (a warm-up evaluation sizes the kernels' scratch).
    const char* gm = getenv("NEMPC_SOLVE_GRAPH");
    const int graph_mode = gm ? atoi(gm) : 1;
    nempc_handle::SolveKey key;
    memset(&key, 0, sizeof key);
    key.B = B; key.buf = h->sv_buf; key.max_iter = o.max_iter; key.max_backtrack = o.max_backtrack;
    const double od[11] = {o.tol, o.mu_init, o.mu_min, o.kappa_eps, o.kappa_mu, o.theta_mu, o.tau_min, o.bound_push, o.eta, o.reg_init, o.reg_max};
    memcpy(key.od, od, sizeof od);
    if (!(key == h->sk)) { drop_solve_graph(h); h->sk = key; }
    h->sk_seen++;
    bool used_graph = false;
    if (graph_mode > 0 && o.max_iter > 0 && !h->sk_disabled && !h->sk_exec && (graph_mode >= 2 || h->sk_seen >= 2)) {
        if (h->sk_seen < 2) {                    // first solve with this key: one evaluation of each kind outside the capture (allocations)
            rc = nempc_eval(h, B, w.z, w.x0, w.lam, nullptr, 1.0, w.resid, w.jac, w.hes, w.obj, w.grad, (void*)s);
            if (!rc) rc = nempc_eval(h, B, w.z, w.x0, nullptr, nullptr, 1.0, w.residt, nullptr, nullptr, w.objt, nullptr, (void*)s);
            if (rc) return rc;
            CU(h, cudaStreamSynchronize(s));
        }
        const long long l0 = h->launches;
        cudaStream_t cs = h->stream;
        cudaGraph_t g = nullptr, tmp = nullptr;
        cudaGraphConditionalHandle hO = 0, hI = 0;
        bool ok = cudaGraphCreate(&g, 0) == cudaSuccess;
        ok = ok && cudaGraphConditionalHandleCreate(&hO, g, 1, cudaGraphCondAssignDefault) == cudaSuccess;
        ok = ok && cudaGraphConditionalHandleCreate(&hI, g, 0, cudaGraphCondAssignDefault) == cudaSuccess;
        cudaGraphNodeParams pO = {cudaGraphNodeTypeConditional};
        pO.type = cudaGraphNodeTypeConditional; pO.conditional.handle = hO; pO.conditional.type = cudaGraphCondTypeWhile; pO.conditional.size = 1;
        cudaGraphNode_t nO = nullptr, nI = nullptr;
        ok = ok && cudaGraphAddNode(&nO, g, nullptr, 0, &pO) == cudaSuccess;
        cudaGraph_t bodyO = ok ? pO.conditional.phGraph_out[0] : nullptr, bodyI = nullptr;
        bool capturing = false;
        if (ok) { ok = cudaStreamBeginCaptureToGraph(cs, bodyO, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess; capturing = ok; }
        if (ok) ok = seg_kkt(cs) == NEMPC_OK;
        h->sk_launches_kkt = h->launches - l0;
        if (ok) { nempc_ipm_ls_begin_kernel<<<1, 1, 0, cs>>>(hI, cnt, o.max_backtrack); ok = cudaGetLastError() == cudaSuccess; }
        if (ok) {                                // the inner WHILE node goes between the segments: after what was captured so far, before the rest
            cudaStreamCaptureStatus st; cudaGraph_t cg = nullptr; const cudaGraphNode_t* deps = nullptr; size_t nd = 0;
            ok = cudaStreamGetCaptureInfo_v2(cs, &st, nullptr, &cg, &deps, &nd) == cudaSuccess && st == cudaStreamCaptureStatusActive;
            cudaGraphNodeParams pI = {cudaGraphNodeTypeConditional};
            pI.type = cudaGraphNodeTypeConditional; pI.conditional.handle = hI; pI.conditional.type = cudaGraphCondTypeWhile; pI.conditional.size = 1;
            ok = ok && cudaGraphAddNode(&nI, cg, deps, nd, &pI) == cudaSuccess;
            if (ok) bodyI = pI.conditional.phGraph_out[0];
            ok = ok && cudaStreamUpdateCaptureDependencies(cs, &nI, 1, cudaStreamSetCaptureDependencies) == cudaSuccess;
        }
        if (ok) ok = seg_update(cs, cnt) == NEMPC_OK;
        if (ok) { nempc_ipm_outer_cond_kernel<<<1, 1, 0, cs>>>(hO, cnt, o.max_iter); ok = cudaGetLastError() == cudaSuccess; }
        if (capturing) { const bool ended = cudaStreamEndCapture(cs, &tmp) == cudaSuccess; ok = ok && ended; capturing = false; }
        if (ok) { ok = cudaStreamBeginCaptureToGraph(cs, bodyI, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess; capturing = ok; }
        const long long l1 = h->launches;
        if (ok) ok = seg_trial(cs) == NEMPC_OK;
        h->sk_launches_trial = h->launches - l1;
        if (ok) { nempc_ipm_ls_cond_kernel<<<1, 1, 0, cs>>>(hI, cnt, o.max_backtrack); ok = cudaGetLastError() == cudaSuccess; }
        if (capturing) { const bool ended = cudaStreamEndCapture(cs, &tmp) == cudaSuccess; ok = ok && ended; }
        h->gate = nullptr;
        h->launches = l0;
        if (ok) ok = cudaGraphInstantiate(&h->sk_exec, g, 0) == cudaSuccess;
        if (g) cudaGraphDestroy(g);
        if (!ok) {                               // the captured loop is an optimisation: keep the host loop for good on this handle
            cudaGetLastError();
            if (h->sk_exec) { cudaGraphExecDestroy(h->sk_exec); h->sk_exec = nullptr; }
            h->sk_disabled = true;
        }
    }
    if (graph_mode > 0 && h->sk_exec && !h->sk_disabled) {
        CU(h, cudaGraphLaunch(h->sk_exec, s));
        used_graph = true;
    } else {
        for (; it < o.max_iter; ++it) {
            rc = seg_kkt(s);
            if (rc) return rc;
            // After a KKT step every running problem needs at least one line-search trial, so the first trial is issued WITHOUT waiting
            // for the counters (one host synchronisation per iteration instead of two); further trials only when some problem backtracks.
            // When nothing is running any more the trial is a wasted residual evaluation, once per solve.
            int trial = 0;
            do {
                if (o.max_backtrack > 0) {
                    rc = seg_trial(s);
                    if (rc) return rc;
                }
                CU(h, cudaMemcpyAsync(h->sv_counts_host, cnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
                CU(h, cudaStreamSynchronize(s));
            } while (++trial < o.max_backtrack && h->sv_counts_host[0] > 0 && h->sv_counts_host[1] > 0);
            if (h->sv_counts_host[0] == 0) break;                      // every problem converged or failed
            rc = seg_update(s, nullptr);
            if (rc) return rc;
        }
    }
    nempc_ipm_finish_kernel<<<grid, threads, 0, s>>>(w, B, status, iterations, kkt_error);
    CU(h, cudaGetLastError()); h->launches++;
    CU(h, cudaMemcpyAsync(z, w.z, (size_t)B * n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (lambda) CU(h, cudaMemcpyAsync(lambda, w.lam, (size_t)B * m * sizeof(double), cudaMemcpyDeviceToDevice, s));
    CU(h, cudaMemcpyAsync(h->sv_counts_host, cnt, NEMPC_SV_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(h, cudaStreamSynchronize(s));
    if (used_graph) {
        it = h->sv_counts_host[NEMPC_SV_IT];
        // kernels the graph ran: per outer iteration segment A + ls_begin + the two update kernels + outer_cond, per trial segment T + ls_cond
        const long long outer_runs = std::min<long long>(it + 1, o.max_iter);
        h->launches += outer_runs * (h->sk_launches_kkt + 4) + (long long)h->sv_counts_host[NEMPC_SV_TRIALS_TOTAL] * (h->sk_launches_trial + 1);
    }
    h->sv_last_graph = used_graph ? 1 : 0;
    h->sv_last_lsx = h->sv_counts_host[NEMPC_SV_LSX];
    h->sv_last_trials = used_graph ? h->sv_counts_host[NEMPC_SV_TRIALS_TOTAL] : -1;
    if (outer_iterations) *outer_iterations = it;
    return NEMPC_OK;
}

/* statistics of the last nempc_solve on this handle */
extern "C" int nempc_solve_stats(const nempc_handle* h, int32_t* used_graph, int64_t* unaccepted_steps) {
    if (!h) return NEMPC_EINVAL;
    if (used_graph) *used_graph = h->sv_last_graph;
    if (unaccepted_steps) *unaccepted_steps = h->sv_last_lsx;
    return NEMPC_OK;
}

// ---- introspection ------------------------------------------------------------------------------------------------
extern "C" int64_t nempc_launch_count(const nempc_handle* h) { return h ? h->launches : 0; }
extern "C" const char* nempc_kernel_name(const nempc_handle* h) { return h ? h->kname.c_str() : ""; }

extern "C" double nempc_flops_per_step(const nempc_handle* h) {
    if (!h) return 0.0;
    const double d = h->d, x = h->desc.x_dim;
    double sumW = 0, sumh = 0;
    for (int l = 0; l < h->L; ++l) sumW += (double)h->dims[l] * h->dims[l + 1];
    for (int l = 0; l + 1 < h->L; ++l) sumh += h->dims[l + 1];
    const double stage = 2 * sumW + 2 * d * (sumW - (double)h->dims[0] * h->dims[1]) + 2 * sumW + d * (d + 1) * sumh;
    const int S = h->desc.integrator == NEMPC_INTEG_RK4 ? 4 : 1;
    return S * stage + (S == 4 ? 6 * x * d * d + 6 * x * x : 0.0);
}

template <typename T> static int fma_peak_t(int millis, double* tflops, bool packed = false) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    T* out = nullptr;
    if (cudaMalloc(&out, sizeof(T)) != cudaSuccess) return NEMPC_ECUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 256, blocks = sms * 8;
    int iters = 2000;
    double best = 0.0;
    float total_ms = 0.f;
    auto launch = [&](int it) {
        if (packed) nempc_fma2_peak_kernel<<<blocks, threads>>>((float*)out, it, 1.f);
        else nempc_fma_peak_kernel<T><<<blocks, threads>>>(out, it, (T)1);
    };
    launch(200);   // warm-up
    cudaDeviceSynchronize();
    while (total_ms < (float)millis) {
        cudaEventRecord(e0);
        launch(iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return NEMPC_ECUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        total_ms += ms;
        const double flops = 2.0 * 16 * 8 * (double)iters * threads * blocks * (packed ? 2.0 : 1.0);
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
        if (ms < 5.f) iters *= 2;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *tflops = best;
    return NEMPC_OK;
}

extern "C" int nempc_measure_fma_peak(int32_t device, int32_t dtype, int32_t millis, double* tflops) {
    if (!tflops) return NEMPC_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) { SET_ERR((nempc_handle*)nullptr, "no usable CUDA device %d", device); return NEMPC_ECUDA; }
    if (dtype == NEMPC_F64) return fma_peak_t<double>(millis, tflops);
    double scalar = 0.0, packed = 0.0;       // FP32: best of scalar FFMA and packed FFMA2 issue
    int rc = fma_peak_t<float>(millis / 2 + 1, &scalar);
    if (rc) return rc;
    rc = fma_peak_t<float>(millis / 2 + 1, &packed, true);
    if (rc) return rc;
    *tflops = scalar > packed ? scalar : packed;
    return NEMPC_OK;
}
