// Translation unit of the tcgen05 kernels for hidden widths 128 / 64 / 32 (nempc_tc.cuh): 84 kernel instantiations that would otherwise
// serialise the build of nempc_lib.cu.  One internal entry point (declared in nempc_lib.cu); not part of the C ABI.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include "../../include/nempc.h"
#include "nempc_generic.cuh"
#include "nempc_tc.cuh"

struct TcArgs { int tc_id; const void* img; const float* cb; const float* wx; const StageTable<float>* st; const NlpLayout* L; int tvp_dim, p_dim, sm_count; };

template <int X, int U, int NHID, int HW, int MODE, typename TIO>
static int launch_tc_mode(const TcArgs& t, const EvalArgs<TIO>& ar, cudaStream_t s) {
    typedef TcCfg<X, U, NHID, MODE, HW> C;
    auto kern = nempc_tc_kernel<C, TIO>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL);
    if (e != cudaSuccess) return (int)e;
    const long long ntiles = (ar.nsteps + C::SPT - 1) / C::SPT;
    const unsigned grid = (unsigned)std::max(1LL, std::min((ntiles + C::NG - 1) / C::NG, (long long)t.sm_count));     // NG tiles in flight per CTA
    kern<<<grid, NEMPC_TC_THREADS, C::TOTAL, s>>>((const __half*)t.img, t.cb, *t.st, *t.L, ar, t.wx, t.tvp_dim, t.p_dim);
    return (int)cudaGetLastError();
}
template <int X, int U, int NHID, int HW, typename TIO>
static int launch_tc_shape(const TcArgs& t, const EvalArgs<TIO>& ar, int mode, cudaStream_t s) {
    switch (mode) {
        case 0: return launch_tc_mode<X, U, NHID, HW, 0, TIO>(t, ar, s);
        case 1: return launch_tc_mode<X, U, NHID, HW, 1, TIO>(t, ar, s);
        default: return launch_tc_mode<X, U, NHID, HW, 2, TIO>(t, ar, s);
    }
}
template <typename TIO> static int launch_tc(const TcArgs& t, const EvalArgs<TIO>& ar, int mode, cudaStream_t s) {
    switch (t.tc_id) {                                    // index into kTcShapes
        case 0: return launch_tc_shape<4, 1, 3, 128, TIO>(t, ar, mode, s);
        case 1: return launch_tc_shape<4, 1, 2, 128, TIO>(t, ar, mode, s);
        case 2: return launch_tc_shape<2, 1, 3, 128, TIO>(t, ar, mode, s);
        case 3: return launch_tc_shape<2, 1, 2, 128, TIO>(t, ar, mode, s);
        case 4: return launch_tc_shape<3, 1, 3, 128, TIO>(t, ar, mode, s);
        case 5: return launch_tc_shape<3, 1, 2, 128, TIO>(t, ar, mode, s);
        case 6: return launch_tc_shape<4, 2, 3, 128, TIO>(t, ar, mode, s);
        case 7: return launch_tc_shape<4, 2, 2, 128, TIO>(t, ar, mode, s);
        case 8: return launch_tc_shape<4, 1, 3, 64, TIO>(t, ar, mode, s);
        case 9: return launch_tc_shape<4, 1, 2, 64, TIO>(t, ar, mode, s);
        case 10: return launch_tc_shape<2, 1, 3, 64, TIO>(t, ar, mode, s);
        case 11: return launch_tc_shape<2, 1, 2, 64, TIO>(t, ar, mode, s);
        case 12: return launch_tc_shape<2, 1, 3, 32, TIO>(t, ar, mode, s);
        case 13: return launch_tc_shape<4, 1, 3, 32, TIO>(t, ar, mode, s);
    }
    return -1;
}


// returns 0, a cudaError_t, or -1 for an unknown shape; tc_id indexes kTcShapes of nempc_lib.cu; `ar` points at an EvalArgs<double> (io_f64) or
// EvalArgs<float> whose exogenous rows are already bound
int nempc_tc_launch(int tc_id, int mode, int io_f64, const void* img, const float* cb, const float* wx, const StageTable<float>& st,
                    const NlpLayout& L, const void* ar, int tvp_dim, int p_dim, int sm_count, cudaStream_t s) {
    const TcArgs t{tc_id, img, cb, wx, &st, &L, tvp_dim, p_dim, sm_count};
    return io_f64 ? launch_tc<double>(t, *static_cast<const EvalArgs<double>*>(ar), mode, s)
                  : launch_tc<float>(t, *static_cast<const EvalArgs<float>*>(ar), mode, s);
}

#ifdef NEMPC_TC_PROFILE
// development builds only: cycles per phase of nempc_tc_kernel summed over all CTAs (thread 0's clock), then reset
extern "C" int nempc_debug_tc_profile(unsigned long long* out16) {
    if (cudaDeviceSynchronize() != cudaSuccess) return NEMPC_ECUDA;
    if (cudaMemcpyFromSymbol(out16, nempc_tc_prof, 16 * sizeof(unsigned long long)) != cudaSuccess) return NEMPC_ECUDA;
    unsigned long long z[16] = {};
    cudaMemcpyToSymbol(nempc_tc_prof, z, sizeof z);
    return NEMPC_OK;
}
#endif
