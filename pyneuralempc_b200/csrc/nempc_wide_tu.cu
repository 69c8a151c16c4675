// Translation unit of the width-256 tcgen05 kernel (nempc_wide.cuh): 48 instantiations (shape x request set x integrator x I/O type).
// Internal entry points (declared in nempc_lib.cu); not part of the C ABI.
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

template <int X, int U, int MODE, bool RK4, typename TIO, int HW>
static int launch_wide_cfg(const WideArgs& t, const EvalArgs<TIO>& ar, cudaStream_t s) {
    typedef WideCfg<X, U, MODE, RK4, HW> C;
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
template <int X, int U, int MODE, typename TIO, int HW>
static int launch_wide_mode(const WideArgs& t, const EvalArgs<TIO>& ar, cudaStream_t s) {
    return t.rk4 ? launch_wide_cfg<X, U, MODE, true, TIO, HW>(t, ar, s) : launch_wide_cfg<X, U, MODE, false, TIO, HW>(t, ar, s);
}
template <int X, int U, typename TIO, int HW = NEMPC_WIDE_HW>
static int launch_wide_shape(const WideArgs& t, const EvalArgs<TIO>& ar, int mode, cudaStream_t s) {
    switch (mode) {
        case 0: return launch_wide_mode<X, U, 0, TIO, HW>(t, ar, s);
        case 1: return launch_wide_mode<X, U, 1, TIO, HW>(t, ar, s);
        default: return launch_wide_mode<X, U, 2, TIO, HW>(t, ar, s);
    }
}
template <typename TIO> static int launch_wide(const WideArgs& t, const EvalArgs<TIO>& ar, int mode, cudaStream_t s) {
    switch (t.wide_id) {                                  // index into kWideShapes
        case 0: return launch_wide_shape<12, 4, TIO>(t, ar, mode, s);
        case 1: return launch_wide_shape<4, 1, TIO>(t, ar, mode, s);
        case 2: return launch_wide_shape<2, 1, TIO>(t, ar, mode, s);
        case 3: return launch_wide_shape<6, 2, TIO>(t, ar, mode, s);
        case 4: return launch_wide_shape<4, 1, TIO, 128>(t, ar, mode, s);       // hidden width 128 (C3 class)
        case 5: return launch_wide_shape<2, 1, TIO, 128>(t, ar, mode, s);
    }
    return -1;
}


// floats of per-CTA scratch the instantiation needs (0 for an unknown shape)
size_t nempc_wide_scratch_floats(int wide_id, int mode, int rk4) {
    const WideArgs t{wide_id, rk4, 1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    const EvalArgs<double> ar{};
    const int r = launch_wide<double>(t, ar, mode, nullptr);
    return r < 0 ? 0 : (size_t)r;
}

// returns 0, a cudaError_t, or -1 for an unknown shape; wide_id indexes kWideShapes of nempc_lib.cu
int nempc_wide_launch(int wide_id, int mode, int rk4, int io_f64, const unsigned char* blob, const float* cb, const WideNet& net,
                      const StageTable<float>& st, const NlpLayout& L, const void* ar, float* scratch, int sm_count, cudaStream_t s) {
    const WideArgs t{wide_id, rk4, 0, blob, cb, &net, &st, &L, scratch, sm_count};
    return io_f64 ? launch_wide<double>(t, *static_cast<const EvalArgs<double>*>(ar), mode, s)
                  : launch_wide<float>(t, *static_cast<const EvalArgs<float>*>(ar), mode, s);
}

#ifdef NEMPC_WIDE_PROFILE
extern "C" int nempc_debug_wide_profile(unsigned long long* out16) {
    if (cudaDeviceSynchronize() != cudaSuccess) return NEMPC_ECUDA;
    if (cudaMemcpyFromSymbol(out16, nempc_wide_prof, 16 * sizeof(unsigned long long)) != cudaSuccess) return NEMPC_ECUDA;
    unsigned long long z[16] = {};
    cudaMemcpyToSymbol(nempc_wide_prof, z, sizeof z);
    return NEMPC_OK;
}
#endif
