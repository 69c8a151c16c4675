// Launch plumbing shared by the two translation units of nempc_wide_kernel (nempc_wide_tu.cu: the specialised shapes; nempc_wide_rt_tu.cu:
// the instantiations that read x_dim / u_dim at run time).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include "../../include/nempc.h"
#include "nempc_generic.cuh"
#include "nempc_wide.cuh"

struct WideArgs { int wide_id, rk4, query; const unsigned char* blob; const float* cb; const WideNet* net; const StageTable<float>* st; const NlpLayout* L;
                  float* scratch; int sm_count; };

template <int X, int U, int MODE, bool RK4, typename TIO, int HW, int DPR = 16>
static int launch_wide_cfg(const WideArgs& t, const EvalArgs<TIO>& ar, cudaStream_t s) {
    typedef WideCfg<X, U, MODE, RK4, HW, DPR> C;
    auto kern = nempc_wide_kernel<C, TIO>;
    if (t.query) return (int)C::SCRATCH_FLOATS;            // (fits an int: < 2^31 floats)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL);
    if (e != cudaSuccess) return (int)e;
    const long long nsup = (ar.nsteps + NEMPC_WIDE_SUP - 1) / NEMPC_WIDE_SUP;
    static const int grid_cap = getenv("NEMPC_WIDE_GRID") ? std::max(1, atoi(getenv("NEMPC_WIDE_GRID"))) : 1 << 30;      // experiments: fewer CTAs
    const long long npair = (nsup + 1) / 2;                                  // CTA pairs (clusters of two, cta_group::2 MMAs)
    const unsigned grid = 2u * (unsigned)std::max(1LL, std::min(npair, (long long)std::min(t.sm_count, grid_cap) / 2));
    kern<<<grid, NEMPC_WIDE_THREADS, C::TOTAL, s>>>(t.blob, t.cb, *t.net, *t.st, *t.L, ar, t.scratch);
    return (int)cudaGetLastError();
}
template <int X, int U, int MODE, typename TIO, int HW, int DPR = 16>
static int launch_wide_mode(const WideArgs& t, const EvalArgs<TIO>& ar, cudaStream_t s) {
    return t.rk4 ? launch_wide_cfg<X, U, MODE, true, TIO, HW, DPR>(t, ar, s) : launch_wide_cfg<X, U, MODE, false, TIO, HW, DPR>(t, ar, s);
}
template <int X, int U, typename TIO, int HW = NEMPC_WIDE_HW, int DPR = 16>
static int launch_wide_shape(const WideArgs& t, const EvalArgs<TIO>& ar, int mode, cudaStream_t s) {
    switch (mode) {
        case 0: return launch_wide_mode<X, U, 0, TIO, HW, DPR>(t, ar, s);
        case 1: return launch_wide_mode<X, U, 1, TIO, HW, DPR>(t, ar, s);
        default: return launch_wide_mode<X, U, 2, TIO, HW, DPR>(t, ar, s);
    }
}
// nempc_wide_rt_tu.cu: wide_id 100 + 3 * (hidden width 128) + {0, 1, 2} (4 / 8 / 16 tangent rows); -1 for anything else
int nempc_wide_rt_launch_f32(const WideArgs& t, const EvalArgs<float>& ar, int mode, cudaStream_t s);
int nempc_wide_rt_launch_f64(const WideArgs& t, const EvalArgs<double>& ar, int mode, cudaStream_t s);
