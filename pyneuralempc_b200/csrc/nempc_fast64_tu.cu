// Translation unit of the float64 register-resident kernels (nempc_fast64.cuh).  Separate from nempc_lib.cu because these kernels are
// sensitive to the register allocation the compiler settles on: built with nvcc --split-compile they measure 0.29 - 0.33 of the FP64 peak on
// C2, inside the big translation unit without it 0.24 -- while the tensor-core kernels of nempc_lib.cu lose 8 % WITH --split-compile.
// One internal entry point (declared in nempc_lib.cu); not part of the C ABI.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include "nempc_generic.cuh"
#include "nempc_fast64.cuh"

#ifndef NEMPC_FAST64_JC30
#define NEMPC_FAST64_JC30 10       // layer-2 neurons per register chunk of the 30-wide networks (tools/fast64_variants.sh: 5 / 6 / 10 / 15 -> 0.31 / 0.33 / 0.34 / 0.34 of the FP64 peak at 128 threads; 96 or 64 threads per CTA: 0.26 - 0.29)
#endif
template <int X, int U, int H1, int H2, int JC>
static int launch_shape(int mode, const void* weights, const StageTable<double>& st, const NlpLayout& L, const EvalArgs<double>& ar, int sm_count,
                        cudaStream_t s) {
    typedef Fast64Weights<X, U, H1, H2> FW;
    const FW& w = *reinterpret_cast<const FW*>(weights);
    const int threads = NEMPC_FAST64_THREADS;
    const long long blocks = std::max(1LL, (ar.nsteps + threads - 1) / threads);
    const unsigned grid = (unsigned)std::min(blocks, (long long)sm_count * 64);
    const size_t wbytes = NEMPC_FAST64_SMEM_WEIGHTS ? (sizeof(FW) + 15) / 16 * 16 : 0;
    const size_t smem = wbytes + (size_t)FastScratch<X, U, H1, H2>::count(mode) * threads * sizeof(double);
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) {
        static bool once[3] = {false, false, false};     // opt in to > 48 KB of dynamic shared memory, once per instantiation
        if (!once[mode]) {
            if (mode == 0) e = cudaFuncSetAttribute(nempc_fast64_kernel<X, U, H1, H2, JC, 0, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            else if (mode == 1) e = cudaFuncSetAttribute(nempc_fast64_kernel<X, U, H1, H2, JC, 1, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            else e = cudaFuncSetAttribute(nempc_fast64_kernel<X, U, H1, H2, JC, 2, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            once[mode] = true;
        }
    }
    switch (mode) {
        case 0: nempc_fast64_kernel<X, U, H1, H2, JC, 0, double><<<grid, threads, smem, s>>>(w, st, L, ar); break;
        case 1: nempc_fast64_kernel<X, U, H1, H2, JC, 1, double><<<grid, threads, smem, s>>>(w, st, L, ar); break;
        default: nempc_fast64_kernel<X, U, H1, H2, JC, 2, double><<<grid, threads, smem, s>>>(w, st, L, ar); break;
    }
    return (int)cudaGetLastError();
}

// returns 0, a cudaError_t, or -1 for an unknown shape; shapes = kFastShapes of nempc_lib.cu
int nempc_fast64_launch(int shape_id, int mode, const void* weights, const StageTable<double>& st, const NlpLayout& L, const EvalArgs<double>& ar,
                        int sm_count, cudaStream_t s) {
    switch (shape_id) {
        case 0: return launch_shape<2, 1, 30, 30, NEMPC_FAST64_JC30>(mode, weights, st, L, ar, sm_count, s);
        case 1: return launch_shape<2, 1, 32, 32, 8>(mode, weights, st, L, ar, sm_count, s);
        case 2: return launch_shape<2, 1, 16, 16, 8>(mode, weights, st, L, ar, sm_count, s);
    }
    return -1;
}
