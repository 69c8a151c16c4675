// Micro-benchmarks behind the design of nempc_tc.cuh (TEST TOOL): tensor-memory read bandwidth per SM for a few
// tcgen05.ld shapes, and the duration of a 24-MMA batch (M=128, N=128, K=16, f16, both operands in shared memory).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tc_tmem_probe tests/tools/tc_tmem_probe.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

#include "../../pyneuralempc_b200/csrc/nempc_tc_ptx.cuh"

using namespace tcx;

__device__ __forceinline__ void ld_x32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// mode 0: x16 loads, wait after each pair; mode 1: x32 loads, wait after each; mode 2: x16, one wait per 4 loads
template <int MODE>
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(long long* cycles, float* sink, int iters, int nwarps_active) {
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&holder), 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = holder;
    const uint32_t row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nwarps_active) {
        for (int it = 0; it < iters; ++it) {
            const uint32_t col = ((warp >> 2) * 64 + (it & 3) * 16) & 255;
            if (MODE == 1) {
                uint32_t r[32];
                ld_x32(row + (col & ~31u), r);
                tmem_ld_wait();
                acc += __uint_as_float(r[it & 31]);
            } else if (MODE == 0) {
                float a[16], b[16];
                tmem_ld16(row + col, a);
                tmem_ld16(row + 256 + col, b);
                tmem_ld_wait();
                acc += a[it & 15] + b[it & 15];
            } else {
                float a[16], b[16], c[16], d[16];
                tmem_ld16(row + col, a);
                tmem_ld16(row + 256 + col, b);
                tmem_ld16(row + ((col + 16) & 255), c);
                tmem_ld16(row + 256 + ((col + 16) & 255), d);
                tmem_ld_wait();
                acc += a[it & 15] + b[it & 15] + c[it & 15] + d[it & 15];
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// One batch = the 24 MMAs of a layer (M=128, K=16 each), issued by one thread exactly as nempc_tc_kernel does it: fully
// unrolled, descriptors advanced in their low word (uniform-register arithmetic), then commit + wait.
//   0: two accumulators (main | corr), N=128, interleaved          1: ONE accumulator, N=128, 24 dependent MMAs
//   2: ONE accumulator split in two N=64 halves, chains interleaved 3: ONE accumulator split in four N=32 quarters
//   4: main only (8 MMAs, one accumulator)
template <int VARIANT>
__global__ void __launch_bounds__(128, 1) mma_batch_kernel(long long* cycles, int batches, int random_data) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar_store;
    __shared__ uint32_t holder;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 196608 / 4; i += 128) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        reinterpret_cast<uint32_t*>(smem)[i] = random_data == 2 ? (h & 0x83ff83ffu) : (random_data == 3 ? (h & 0xffffffffu & 0xfbfffbffu) : (random_data ? ((h & 0x83ff83ffu) | 0x38003800u) : 0x3c003c00u));   // 2: subnormals only, 3: any finite f16   // f16 in [0.5, 1) with random signs/mantissas, or ones
    }
    const uint32_t mbar = smem_u32(&mbar_store);
    if (tid == 0) { mbar_init(mbar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&holder), 512);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = holder;
    if (warp == 0) {
        if (tid == 0) {
            const uint32_t ahi = smem_u32(smem), alo = ahi + 32768, whi = ahi + 65536, wlo = whi + 32768;
            const uint32_t dhi = (128u >> 4) | (1u << 14), dlo = (2048u >> 4) << 16;
            const uint32_t la1 = dlo | (ahi >> 4), la2 = dlo | (alo >> 4), lw1 = dlo | (whi >> 4), lw2 = dlo | (wlo >> 4);
#define DESC(lo, ks, rowoff) ((((uint64_t)dhi) << 32) | (uint64_t)((lo) + (ks) * 256u + (rowoff)))
            constexpr int NS = VARIANT == 2 ? 2 : (VARIANT == 3 ? 4 : 1), NN = 128 / NS;
            const uint32_t idesc = make_idesc_f16(128, NN);
            uint32_t parity = 0;
            const long long t0 = clock64();
            for (int b = 0; b < batches; ++b) {
                if (VARIANT == 5) {          // the kernel's sequence: corrections, scale-input-d fold, rest of the main chain
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) mma_f16_ss(tmem, DESC(la2, ks, 0), DESC(lw1, ks, 0), idesc, ks != 0);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) mma_f16_ss(tmem, DESC(la1, ks, 0), DESC(lw2, ks, 0), idesc, 1);
                    mma_f16_ss_scaled_d<11>(tmem, DESC(la1, 0, 0), DESC(lw1, 0, 0), idesc);
#pragma unroll
                    for (int ks = 1; ks < 8; ++ks) mma_f16_ss(tmem, DESC(la1, ks, 0), DESC(lw1, ks, 0), idesc, 1);
                }
#pragma unroll
                for (int ks = 0; ks < 8 && VARIANT != 5; ++ks) {
                    if (VARIANT == 0) {
                        mma_f16_ss(tmem, DESC(la1, ks, 0), DESC(lw1, ks, 0), idesc, ks > 0);
                        mma_f16_ss(tmem + 128, DESC(la2, ks, 0), DESC(lw1, ks, 0), idesc, ks > 0);
                        mma_f16_ss(tmem + 128, DESC(la1, ks, 0), DESC(lw2, ks, 0), idesc, 1);
                    } else if (VARIANT == 4) {
                        mma_f16_ss(tmem, DESC(la1, ks, 0), DESC(lw1, ks, 0), idesc, ks > 0);
                    } else {
#pragma unroll
                        for (int h = 0; h < NS; ++h) mma_f16_ss(tmem + h * NN, DESC(la2, ks, 0), DESC(lw1, ks, h * NN), idesc, ks > 0);
#pragma unroll
                        for (int h = 0; h < NS; ++h) mma_f16_ss(tmem + h * NN, DESC(la1, ks, 0), DESC(lw2, ks, h * NN), idesc, 1);
#pragma unroll
                        for (int h = 0; h < NS; ++h) mma_f16_ss(tmem + h * NN, DESC(la1, ks, 0), DESC(lw1, ks, h * NN), idesc, 1);
                    }
                }
#undef DESC
                mma_commit(mbar);
                mbar_wait(mbar, parity);
                parity ^= 1;
            }
            if (blockIdx.x == 0) cycles[0] = clock64() - t0;
        }
        __syncwarp();
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int VARIANT> static void run_mma(long long* dc, const char* what) {
    cudaFuncSetAttribute(mma_batch_kernel<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
    for (int grid : {1, 148})
        for (int rnd : {0, 1, 2, 3}) {
            mma_batch_kernel<VARIANT><<<grid, 128, 196608>>>(dc, 2000, rnd);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("mma kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
            long long c;
            cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
            printf("mma batch variant %d (%s), %3d CTAs, %s data: %.0f cycles per batch\n", VARIANT, what, grid, rnd == 0 ? "ones" : (rnd == 1 ? "random normal" : (rnd == 2 ? "subnormal" : "any finite")), (double)c / 2000);
        }
}

// As variant 5, but every batch is preceded by what the kernel does: all 512 threads rewrite the 64 KB operand tile with
// generic-proxy stores, fence.proxy.async, barrier.  Only the issue + wait time of thread 0 is accumulated.
__global__ void __launch_bounds__(512, 1) mma_after_stores_kernel(long long* cycles, int batches, int do_stores, int idle_cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar_store;
    __shared__ uint32_t holder;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 196608 / 4; i += 512) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    const uint32_t mbar = smem_u32(&mbar_store);
    if (tid == 0) { mbar_init(mbar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&holder), 128);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = holder;
    long long acc = 0;
    uint32_t parity = 0;
    for (int b = 0; b < batches; ++b) {
        if (do_stores) {
            const int m = tid & 127, cq = tid >> 7;
            for (int kc = 4 * cq; kc < 4 * cq + 4; ++kc) {
                *reinterpret_cast<uint4*>(smem + 131072 + kc * 2048 + m * 16) = make_uint4(0x3c003c00u + b, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
                *reinterpret_cast<uint4*>(smem + 131072 + 32768 + kc * 2048 + m * 16) = make_uint4(0x3c003c00u, 0x3c003c00u + b, 0x3c003c00u, 0x3c003c00u);
            }
        }
        if (idle_cycles) { const long long w0 = clock64(); while (clock64() - w0 < idle_cycles) {} }   // tensor core idle, ALUs busy
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        if (warp == 0) {
            if (tid == 0) {
                fence_after_sync();
                const long long t0 = clock64();
                const uint32_t whi = smem_u32(smem), wlo = whi + 32768, ahi = whi + 131072, alo = ahi + 32768;
                const uint32_t dhi = (128u >> 4) | (1u << 14), dlo = (2048u >> 4) << 16;
                const uint32_t la1 = dlo | (ahi >> 4), la2 = dlo | (alo >> 4), lw1 = dlo | (whi >> 4), lw2 = dlo | (wlo >> 4);
                const uint32_t idesc = make_idesc_f16(128, 128);
#define DESC(lo, ks) ((((uint64_t)dhi) << 32) | (uint64_t)((lo) + (ks) * 256u))
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) mma_f16_ss(tmem, DESC(la2, ks), DESC(lw1, ks), idesc, ks != 0);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) mma_f16_ss(tmem, DESC(la1, ks), DESC(lw2, ks), idesc, 1);
                mma_f16_ss_scaled_d<11>(tmem, DESC(la1, 0), DESC(lw1, 0), idesc);
#pragma unroll
                for (int ks = 1; ks < 8; ++ks) mma_f16_ss(tmem, DESC(la1, ks), DESC(lw1, ks), idesc, 1);
#undef DESC
                mma_commit(mbar);
                mbar_wait(mbar, parity);
                acc += clock64() - t0;
            }
            __syncwarp();
        }
        parity ^= 1;
        __syncthreads();
        fence_after_sync();
    }
    if (tid == 0 && blockIdx.x == 0) cycles[0] = acc;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
    long long* dc; float* ds;
    cudaMalloc(&dc, 8); cudaMalloc(&ds, 4);
    long long c;
    const int iters = 2000;
    for (int nw : {4, 8, 16}) {
        for (int mode = 0; mode < 3; ++mode) {
            if (mode == 0) tmem_read_kernel<0><<<1, 512>>>(dc, ds, iters, nw);
            if (mode == 1) tmem_read_kernel<1><<<1, 512>>>(dc, ds, iters, nw);
            if (mode == 2) tmem_read_kernel<2><<<1, 512>>>(dc, ds, iters, nw);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)nw * iters * (mode == 0 ? 2 : (mode == 1 ? 2 : 4)) * 2048.0;
            printf("tmem read: %2d warps, mode %d (%s): %lld cycles, %.1f B/clk/SM\n", nw, mode,
                   mode == 0 ? "2 x .x16 then wait" : (mode == 1 ? "1 x .x32 then wait" : "4 x .x16 then wait"), c, bytes / c);
        }
    }
    run_mma<0>(dc, "24 MMAs N=128, two accumulators interleaved");
    run_mma<1>(dc, "24 MMAs N=128, ONE accumulator");
    run_mma<2>(dc, "48 MMAs N=64, one accumulator split in two N halves");
    run_mma<3>(dc, "96 MMAs N=32, one accumulator split in four N quarters");
    run_mma<4>(dc, "8 MMAs N=128, one accumulator");
    run_mma<5>(dc, "24 MMAs N=128, ONE accumulator, scale-input-d fold (the kernel's sequence)");
    cudaFuncSetAttribute(mma_after_stores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
    for (int idle : {0, 1000, 5000, 20000}) {
        mma_after_stores_kernel<<<148, 512, 196608>>>(dc, 500, 1, idle);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        printf("mma batch in a 512-thread CTA, tile rewritten, tensor core idle %5d clk between batches: %.0f cycles per batch (issue + wait)\n", idle, (double)c / 500);
    }
    return 0;
}
