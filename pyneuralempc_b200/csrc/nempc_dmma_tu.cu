// Translation unit of the float64 tensor-core (DMMA) path (nempc_dmma.cuh): network kernel + thread-per-step stage kernel, issued stage by
// stage over chunks of the batch.  Two internal entry points (declared in nempc_lib.cu); not part of the C ABI.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include "nempc_dmma.cuh"

#ifndef NEMPC_DMMA_CHUNK_STEPS
#define NEMPC_DMMA_CHUNK_STEPS (1 << 18)       // horizon steps per pass: bounds the scratch (2 KB per step at (x,u) = (4,1)) at ~0.5 GB
#endif

template <int X, int U> static size_t per_step_doubles() {
    constexpr int D = X + U, NS = D * (D + 1) / 2;
    return (size_t)D + X + (size_t)X * D + (size_t)X * NS + DmmaState<X, U>::COUNT;
}

template <int X, int U, int HW>
static int launch_shape(int mode, const DmmaNet& net, const StageTable<double>& st, const NlpLayout& L, const EvalArgs<double>& ar, double* scratch,
                        int sm_count, cudaStream_t s, int* launches) {
    typedef DmmaCfg<HW> C;
    constexpr int D = X + U, NS = D * (D + 1) / 2;
    static bool once = false;
    if (!once) {
        cudaError_t e = cudaFuncSetAttribute(nempc_dmma_net_kernel<C, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return (int)e;
        once = true;
    }
    const int rps = 1 + (mode >= 1 ? D : 0) + (mode >= 2 ? NS : 0), spt = C::MT / rps;
    for (long long base = 0; base < ar.nsteps; base += NEMPC_DMMA_CHUNK_STEPS) {
        const long long N = std::min<long long>(NEMPC_DMMA_CHUNK_STEPS, ar.nsteps - base);
        double* zin = scratch; double* fo = zin + N * D; double* Jo = fo + N * X; double* Mo = Jo + N * X * D; double* state = Mo + N * X * NS;
        const unsigned sgrid = (unsigned)((N + 127) / 128);
        const long long ntiles = (N + spt - 1) / spt;
        const unsigned ngrid = (unsigned)std::min<long long>((ntiles + C::NG - 1) / C::NG, sm_count);      // persistent CTAs, one row tile per group
        nempc_dmma_stage_kernel<X, U><<<sgrid, 128, 0, s>>>(st, L, ar, base, N, mode, -1, zin, fo, Jo, Mo, state);
        ++*launches;
        for (int stage = 0; stage < st.S; ++stage) {
            nempc_dmma_net_kernel<C, D><<<ngrid, C::THREADS, C::SMEM, s>>>(net, zin, N, mode, fo, Jo, Mo);
            nempc_dmma_stage_kernel<X, U><<<sgrid, 128, 0, s>>>(st, L, ar, base, N, mode, stage, zin, fo, Jo, Mo, state);
            *launches += 2;
        }
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

// shapes: the (x, u) pairs of kDmmaShapes in nempc_lib.cu
size_t nempc_dmma_scratch_doubles(int shape_id, long long nsteps) {
    const size_t n = (size_t)std::min<long long>(nsteps, NEMPC_DMMA_CHUNK_STEPS);
    switch (shape_id) {
        case 0: return n * per_step_doubles<2, 1>();
        case 1: return n * per_step_doubles<3, 1>();
        case 2: return n * per_step_doubles<4, 1>();
        case 3: return n * per_step_doubles<4, 2>();
        case 4: return n * per_step_doubles<6, 2>();
    }
    return 0;
}

// returns 0, a cudaError_t, or -1 for an unknown shape / width
int nempc_dmma_launch(int shape_id, int hw, int mode, const DmmaNet& net, const StageTable<double>& st, const NlpLayout& L, const EvalArgs<double>& ar,
                      double* scratch, int sm_count, cudaStream_t s, int* launches) {
#define NEMPC_DMMA_CASE(ID, X_, U_) \
    case ID: return hw == 128 ? launch_shape<X_, U_, 128>(mode, net, st, L, ar, scratch, sm_count, s, launches) \
                  : hw == 64 ? launch_shape<X_, U_, 64>(mode, net, st, L, ar, scratch, sm_count, s, launches) : -1;
    switch (shape_id) {
        NEMPC_DMMA_CASE(0, 2, 1)
        NEMPC_DMMA_CASE(1, 3, 1)
        NEMPC_DMMA_CASE(2, 4, 1)
        NEMPC_DMMA_CASE(3, 4, 2)
        NEMPC_DMMA_CASE(4, 6, 2)
    }
#undef NEMPC_DMMA_CASE
    return -1;
}
