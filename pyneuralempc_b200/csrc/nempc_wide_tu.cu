// Translation unit of the width-256 tcgen05 kernel (nempc_wide.cuh): 72 instantiations (6 specialised shapes x request set x integrator x I/O type).
// Internal entry points (declared in nempc_lib.cu); not part of the C ABI.
#include "nempc_wide_launch.cuh"

template <typename TIO> static int launch_wide(const WideArgs& t, const EvalArgs<TIO>& ar, int mode, cudaStream_t s) {
    switch (t.wide_id) {                                  // index into kWideShapes
        case 0: return launch_wide_shape<12, 4, TIO>(t, ar, mode, s);
        case 1: return launch_wide_shape<4, 1, TIO>(t, ar, mode, s);
        case 2: return launch_wide_shape<2, 1, TIO>(t, ar, mode, s);
        case 3: return launch_wide_shape<6, 2, TIO>(t, ar, mode, s);
        case 4: return launch_wide_shape<4, 1, TIO, 128>(t, ar, mode, s);       // hidden width 128 (C3 class)
        case 5: return launch_wide_shape<2, 1, TIO, 128>(t, ar, mode, s);
    }
    // any other (x_dim, u_dim) with x_dim + u_dim <= 16: dimensions read from the layout at run time (nempc_wide_rt_tu.cu)
    if constexpr (sizeof(TIO) == 8) return nempc_wide_rt_launch_f64(t, ar, mode, s);
    else return nempc_wide_rt_launch_f32(t, ar, mode, s);
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
