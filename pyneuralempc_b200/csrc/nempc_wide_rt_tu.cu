// Second translation unit of nempc_wide_kernel (nempc_wide.cuh): the instantiations that take (x_dim, u_dim) from the layout at run time,
// WideCfg<0, 0, MODE, RK4, HW, DP> with DP = 4 / 8 / 16 tangent rows -- any shape with x_dim + u_dim <= 16 that has no specialised
// instantiation in nempc_wide_tu.cu.  72 kernels; a separate file so that the two halves compile side by side.
#include "nempc_wide_launch.cuh"

template <typename TIO> static int launch_wide_rt(const WideArgs& t, const EvalArgs<TIO>& ar, int mode, cudaStream_t s) {
    switch (t.wide_id) {
        case 100: return launch_wide_shape<0, 0, TIO, 256, 4>(t, ar, mode, s);
        case 101: return launch_wide_shape<0, 0, TIO, 256, 8>(t, ar, mode, s);
        case 102: return launch_wide_shape<0, 0, TIO, 256, 16>(t, ar, mode, s);
        case 103: return launch_wide_shape<0, 0, TIO, 128, 4>(t, ar, mode, s);
        case 104: return launch_wide_shape<0, 0, TIO, 128, 8>(t, ar, mode, s);
        case 105: return launch_wide_shape<0, 0, TIO, 128, 16>(t, ar, mode, s);
    }
    return -1;
}
int nempc_wide_rt_launch_f32(const WideArgs& t, const EvalArgs<float>& ar, int mode, cudaStream_t s) { return launch_wide_rt<float>(t, ar, mode, s); }
int nempc_wide_rt_launch_f64(const WideArgs& t, const EvalArgs<double>& ar, int mode, cudaStream_t s) { return launch_wide_rt<double>(t, ar, mode, s); }
