// Microbenchmark: cost of issuing tcgen05.mma (kind::f16, cta_group::1, M = 128, K = 16) from one thread, as a function of N,
// of the operand source (shared-memory descriptors, no-swizzle canonical layout / A from TMEM), of the accumulator dependency
// (one accumulator / alternating accumulators) and of the number of issuing warps.  One CTA per SM.
//   issue  = clock64 around the G back-to-back MMA issues + the commit (what the issuing thread is blocked for)
//   done   = until the commit's mbarrier fires (what a consumer waits for)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I point-cloud-audio_b200/csrc -o tools/mb_mma_issue tools/mb_mma_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_prims.cuh"

using namespace pca::tc;

// MODE: 0 = A, B from shared memory, one accumulator; 1 = alternating accumulators; 2 = A from TMEM, one accumulator
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* out, int rounds, int G, int nwarps) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = tmem_slot;
    if (warp < nwarps) {
        const bool leader = elect_one();
        const uint32_t idesc = idesc_bf16(128, N, 0, 0);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 32768;
        long long t_issue = 0, t_done = 0;
        for (int r = 0; r < rounds; ++r) {
            const long long c0 = clock64();
            if (leader) {
                for (int g = 0; g < G; ++g) {
                    const uint32_t acc = tmem_addr(tb, 0, 256 * (warp & 1) + ((MODE == 1) ? (g & 1) * 128 : 0));   // (N <= 128 for MODE 1)
                    const uint64_t da = smem_desc(sa + (g & 3) * 4096, 2048, 128), db = smem_desc(sb + (g & 3) * 8192, 4096, 128);
                    if (MODE == 2) mma_ts(acc, tmem_addr(tb, 0, 448), db, idesc, g > 0);
                    else mma_ss(acc, da, db, idesc, g > (MODE == 1 ? 1 : 0));
                }
                mma_commit(&bar[warp]);
            }
            __syncwarp();
            const long long c1 = clock64();
            mbar_wait(&bar[warp], r & 1);
            const long long c2 = clock64();
            t_issue += c1 - c0;
            t_done += c2 - c0;
        }
        if (blockIdx.x == 0 && lane == 0) { out[2 * warp] = t_issue / rounds; out[2 * warp + 1] = t_done / rounds; }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N, int MODE>
static void run(const char* name, int G, int nwarps) {
    long long* d;
    cudaMalloc(&d, 64);
    cudaMemset(d, 0, 64);
    cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
    k<N, MODE><<<148, 128, 65536, 0>>>(d, 200, G, nwarps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("%-28s N=%3d G=%2d warps=%d: issue %6lld cycles (%5.1f per MMA), done %6lld (%5.1f per MMA)%s\n", name, N, G, nwarps, h[0],
           (double)h[0] / G, h[1], (double)h[1] / G, e == cudaSuccess ? "" : cudaGetErrorString(e));
    if (nwarps > 1) printf("%-28s        second warp: issue %6lld (%5.1f), done %6lld (%5.1f)\n", "", h[2], (double)h[2] / G, h[3], (double)h[3] / G);
    cudaFree(d);
}

int main() {
    for (int G : {1, 4, 8, 16}) {
        run<16, 0>("smem A/B, one accumulator", G, 1);
        run<64, 0>("smem A/B, one accumulator", G, 1);
        run<128, 0>("smem A/B, one accumulator", G, 1);
        run<256, 0>("smem A/B, one accumulator", G, 1);
    }
    run<16, 1>("smem A/B, two accumulators", 8, 1);
    run<64, 1>("smem A/B, two accumulators", 8, 1);
    run<128, 1>("smem A/B, two accumulators", 8, 1);
    run<16, 2>("TMEM A, one accumulator", 8, 1);
    run<64, 2>("TMEM A, one accumulator", 8, 1);
    run<128, 2>("TMEM A, one accumulator", 8, 1);
    run<16, 0>("two issuing warps", 8, 2);
    run<64, 0>("two issuing warps", 8, 2);
    run<128, 0>("two issuing warps", 8, 2);
    return 0;
}
