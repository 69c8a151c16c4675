// Stand-alone probe of the tcgen05 building blocks used by nempc_tc.cuh (TEST TOOL, not part of the product library):
// D[128 x N] = A[128 x 128] * W[128 x N] with the two-term f16 split (main + scaled correction accumulator), A image written
// by the threads (one row per thread, 16-byte K chunks), B image prepared on the host in the canonical K-major no-swizzle
// layout and brought in with a 1-D bulk copy.  Prints the max relative error against a float64 product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tc_gemm_probe tests/tools/tc_gemm_probe.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../pyneuralempc_b200/csrc/nempc_tc_ptx.cuh"

using namespace tcx;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } \
    } while (0)

constexpr int M = 128, K = 128;

// smem: A hi/lo images (32 KB each), B hi/lo images (N*K*2 each)
template <int N>
__global__ void __launch_bounds__(256, 1) probe_kernel(const float* __restrict__ A, const __half* __restrict__ Bimg, float* __restrict__ D, float* __restrict__ D2, float* __restrict__ D3) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar_store[2];
    __shared__ uint32_t tmem_holder;
    unsigned char* Ahi = smem;
    unsigned char* Alo = smem + M * K * 2;
    unsigned char* Bhi = smem + 2 * M * K * 2;
    unsigned char* Blo = Bhi + N * K * 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t mbar_mma = smem_u32(&mbar_store[0]), mbar_ld = smem_u32(&mbar_store[1]);
    if (tid == 0) { mbar_init(mbar_mma, 1); mbar_init(mbar_ld, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_holder), 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_holder;
    if (tid == 0) {
        mbar_expect_tx(mbar_ld, 2 * N * K * 2);
        bulk_g2s(smem_u32(Bhi), Bimg, 2 * N * K * 2, mbar_ld);
    }
    // A image: thread (row m, half hf) writes K chunks [8*hf, 8*hf+8)
    {
        const int m = tid & 127, hf = tid >> 7;
        for (int kc = 8 * hf; kc < 8 * hf + 8; ++kc) {
            __half hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_f16(A[m * K + kc * 8 + e], hi[e], lo[e]);
            *reinterpret_cast<uint4*>(Ahi + kc * (M * 16) + m * 16) = *reinterpret_cast<uint4*>(hi);
            *reinterpret_cast<uint4*>(Alo + kc * (M * 16) + m * 16) = *reinterpret_cast<uint4*>(lo);
        }
    }
    fence_async_smem();
    mbar_wait(mbar_ld, 0);
    __syncthreads();
    if (tid == 0) {
        fence_after_sync();
        const uint32_t idesc = make_idesc_f16(M, N);
        const uint32_t a_lbo = M * 16, b_lbo = N * 16, sbo = 128;
        // (a) two accumulators: main at column 0, scaled correction at column 128
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t a1 = make_desc_kmajor(smem_u32(Ahi) + ks * 2 * a_lbo, a_lbo, sbo);
            const uint64_t a2 = make_desc_kmajor(smem_u32(Alo) + ks * 2 * a_lbo, a_lbo, sbo);
            const uint64_t w1 = make_desc_kmajor(smem_u32(Bhi) + ks * 2 * b_lbo, b_lbo, sbo);
            const uint64_t w2 = make_desc_kmajor(smem_u32(Blo) + ks * 2 * b_lbo, b_lbo, sbo);
            mma_f16_ss(tmem, a1, w1, idesc, ks > 0);
            mma_f16_ss(tmem + 128, a2, w1, idesc, ks > 0);
            mma_f16_ss(tmem + 128, a1, w2, idesc, 1);
        }
        // (b) ONE accumulator at column 256: correction terms first, then the first main MMA folds them in with
        //     scale-input-d:  D = A_hi W_hi + D * 2^-11
        for (int pass = 0; pass < 3; ++pass)
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint64_t a1 = make_desc_kmajor(smem_u32(Ahi) + ks * 2 * a_lbo, a_lbo, sbo);
                const uint64_t a2 = make_desc_kmajor(smem_u32(Alo) + ks * 2 * a_lbo, a_lbo, sbo);
                const uint64_t w1 = make_desc_kmajor(smem_u32(Bhi) + ks * 2 * b_lbo, b_lbo, sbo);
                const uint64_t w2 = make_desc_kmajor(smem_u32(Blo) + ks * 2 * b_lbo, b_lbo, sbo);
                if (pass == 0) mma_f16_ss(tmem + 256, a2, w1, idesc, ks > 0);
                else if (pass == 1) mma_f16_ss(tmem + 256, a1, w2, idesc, 1);
                else if (ks == 0) mma_f16_ss_scaled_d<11>(tmem + 256, a1, w1, idesc);
                else mma_f16_ss(tmem + 256, a1, w1, idesc, 1);
            }
        mma_commit(mbar_mma);
    }
    mbar_wait(mbar_mma, 0);
    fence_after_sync();
    {
        const int m = (warp & 3) * 32 + lane;
        const int c0 = (warp >> 2) * (N / 2);
        for (int c = c0; c < c0 + N / 2; c += (N >= 32 ? 16 : 8)) {
            if (N < 32 && warp >= 4) break;
            float vm[16], vc[16], vs[16];
            const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (N < 32 ? 0 : c);
            tmem_ld16(ta, vm);
            tmem_ld16(ta + 128, vc);
            tmem_ld16(ta + 256, vs);
            tmem_ld_wait();
            const int base = N < 32 ? 0 : c;
#pragma unroll
            for (int i = 0; i < 16; ++i) { D[m * N + base + i] = vm[i] + vc[i] * NEMPC_TC_LO_INV; D2[m * N + base + i] = vs[i]; }
            if (N < 32) break;
        }
    }
    // ---- phase 2: the A operand in TENSOR MEMORY (TS-mode MMA): thread (row m, half hf) packs its 64 K elements two per
    // 32-bit word and stores them with tcgen05.st at columns 384 + 32 hf (hi image) and 448 + 32 hf (lo image); D at column 0
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    {
        const int m = (warp & 3) * 32 + lane, hf = warp >> 2;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        for (int g = 0; g < 4; ++g) {                         // 4 groups of 16 elements = 8 words
            uint32_t whi[8], wlo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                __half h0, l0, h1, l1;
                split_f16(A[m * K + 64 * hf + 16 * g + 2 * e], h0, l0);
                split_f16(A[m * K + 64 * hf + 16 * g + 2 * e + 1], h1, l1);
                whi[e] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                wlo[e] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            }
            tmem_st8(trow + 384 + 32 * hf + 8 * g, whi);
            tmem_st8(trow + 448 + 32 * hf + 8 * g, wlo);
        }
        tmem_st_wait();
    }
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        fence_after_sync();
        const uint32_t idesc = make_idesc_f16(M, N);
        const uint32_t b_lbo = N * 16, sbo = 128;
        for (int pass = 0; pass < 3; ++pass)
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint32_t a1 = tmem + 384 + ks * 8, a2 = tmem + 448 + ks * 8;
                const uint64_t w1 = make_desc_kmajor(smem_u32(Bhi) + ks * 2 * b_lbo, b_lbo, sbo);
                const uint64_t w2 = make_desc_kmajor(smem_u32(Blo) + ks * 2 * b_lbo, b_lbo, sbo);
                if (pass == 0) mma_f16_ts(tmem, a2, w1, idesc, ks > 0);
                else if (pass == 1) mma_f16_ts(tmem, a1, w2, idesc, 1);
                else if (ks == 0) mma_f16_ts_scaled_d<11>(tmem, a1, w1, idesc);
                else mma_f16_ts(tmem, a1, w1, idesc, 1);
            }
        mma_commit(mbar_mma);
    }
    mbar_wait(mbar_mma, 1);
    fence_after_sync();
    {
        const int m = (warp & 3) * 32 + lane;
        const int c0 = (warp >> 2) * (N / 2);
        for (int c = c0; c < c0 + N / 2; c += (N >= 32 ? 16 : 8)) {
            if (N < 32 && warp >= 4) break;
            float vs[16];
            const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (N < 32 ? 0 : c);
            tmem_ld16(ta, vs);
            tmem_ld_wait();
            const int base = N < 32 ? 0 : c;
#pragma unroll
            for (int i = 0; i < 16; ++i) D3[m * N + base + i] = vs[i];
            if (N < 32) break;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N> static int run(unsigned seed) {
    std::vector<float> A(M * K), W(K * N);
    srand(seed);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : W) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.2f;
    for (int i = 0; i < K; ++i) A[5 * K + i] *= 1e-4f;        // a tiny row and a large row: exercises the scaled lo part
    for (int i = 0; i < K; ++i) A[9 * K + i] *= 300.f;
    // B image: chunk (n, kc) at kc * (N*16) + n*16 bytes, 8 halves = W[kc*8 .. +8][n]; hi image then lo image
    std::vector<__half> img(2 * N * K);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const float w = W[k * N + n];
            const __half hi = __float2half_rn(w);
            const __half lo = __float2half_rn((w - __half2float(hi)) * 2048.f);
            const size_t off = (size_t)(k / 8) * (N * 8) + (size_t)n * 8 + (k % 8);
            img[off] = hi;
            img[(size_t)N * K + off] = lo;
        }
    float *dA, *dD, *dD2, *dD3; __half* dB;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&dD2, M * N * 4)); CK(cudaMalloc(&dD3, M * N * 4)); CK(cudaMalloc(&dB, img.size() * 2));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, M * N * 4));
    const size_t smem = 2 * M * K * 2 + 2 * N * K * 2;
    CK(cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<N><<<1, 256, smem>>>(dA, dB, dD, dD2, dD3);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D(M * N), D2(M * N), D3(M * N);
    CK(cudaMemcpy(D3.data(), dD3, D3.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0.0, worst2 = 0.0, worst3 = 0.0;
    for (int m = 0; m < M; ++m) {
        double rowmax = 0.0, err = 0.0, err2 = 0.0, err3 = 0.0;
        for (int n = 0; n < N; ++n) {
            double ref = 0.0;
            for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)W[k * N + n];
            rowmax = std::max(rowmax, std::fabs(ref));
            err = std::max(err, std::fabs(ref - (double)D[m * N + n]));
            err2 = std::max(err2, std::fabs(ref - (double)D2[m * N + n]));
            err3 = std::max(err3, std::fabs(ref - (double)D3[m * N + n]));
        }
        worst = std::max(worst, err / rowmax);
        worst2 = std::max(worst2, err2 / rowmax);
        worst3 = std::max(worst3, err3 / rowmax);
    }
    printf("N=%d  A operand in tensor memory (TS mode), single accumulator: max row-relative error %.3e  %s\n", N, worst3, worst3 < 2e-6 ? "OK" : "FAIL");
    printf("N=%d  single accumulator with scale-input-d: max row-relative error %.3e  %s\n", N, worst2, worst2 < 2e-6 ? "OK" : "FAIL");
    printf("N=%d  max row-relative error %.3e  D[0][0]=%g D[127][%d]=%g  %s\n", N, worst, D[0], N - 1, D[127 * N + N - 1],
           worst < 2e-6 ? "OK" : "FAIL");
    cudaFree(dA); cudaFree(dD); cudaFree(dB);
    return worst < 2e-6 ? 0 : 1;
}

int main() {
    int rc = run<128>(1);
    rc |= run<16>(2);
    return rc;
}
