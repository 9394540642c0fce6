// Micro-benchmark of the hand-off primitives of the persistent tcgen05 kernels (one CTA, clock64):
//   1. mbarrier ping-pong between two warps: suspending try_wait (with the time hint the kernels use) vs spinning test_wait
//   2. tcgen05.mma (128 x 64 x 16, smem operands) -> tcgen05.commit -> mbarrier: issue-to-completion latency seen by the issuer
//   3. st.shared + fence.proxy.async; tcgen05.ld x32 + wait::ld; tcgen05.st x16 + wait::st
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I point-cloud-audio_b200/csrc -o tools/mb_handoff tools/mb_handoff.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_prims.cuh"
using namespace pca::tc;

constexpr int ITERS = 200;

__global__ void __launch_bounds__(128) handoff_kernel(long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); fence_barrier_init(); }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;
    long long t0, t1;
    // ---- 1. ping-pong (warp 0 lane 0 <-> warp 1 lane 0): mode 0 suspending, mode 1 spinning
    for (int mode = 0; mode < 2; ++mode) {
        uint64_t* ping = &bars[2 * mode], *pong = &bars[2 * mode + 1];
        __syncthreads();
        if (lane == 0 && warp < 2) {
            t0 = clock64();
            for (int i = 0; i < ITERS; ++i) {
                if (warp == 0) {
                    mbar_arrive(ping);
                    if (mode == 0) mbar_wait(pong, i & 1); else mbar_spin(pong, i & 1);
                } else {
                    if (mode == 0) mbar_wait(ping, i & 1); else mbar_spin(ping, i & 1);
                    mbar_arrive(pong);
                }
            }
            t1 = clock64();
            if (warp == 0) out[mode] = (t1 - t0) / ITERS;          // one round trip = two hand-offs
        }
    }
    __syncthreads();
    // ---- 2. MMA + commit + wait (issuer's view), suspending and spinning
    for (int mode = 0; mode < 2; ++mode) {
        if (threadIdx.x == 0) {
            const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16384);
            const uint32_t idesc = idesc_bf16(128, 64, 0, 0);
            t0 = clock64();
            for (int i = 0; i < ITERS; ++i) {
                mma_ss(tmem_addr(tb, 0, 0), smem_desc(a, 2048, 128), smem_desc(b, 1024, 128), idesc, 0);
                mma_commit(&bars[4 + mode]);
                if (mode == 0) mbar_wait(&bars[4 + mode], i & 1); else mbar_spin(&bars[4 + mode], i & 1);
            }
            t1 = clock64();
            out[2 + mode] = (t1 - t0) / ITERS;
        }
        __syncthreads();
    }
    // ---- 3. per-thread primitives (warp 0)
    if (warp == 0) {
        uint32_t v[32];
        t0 = clock64();
        for (int i = 0; i < ITERS; ++i) {
            *reinterpret_cast<uint4*>(smem + 32768 + lane * 16) = make_uint4(i, i, i, i);
            fence_async_smem();
        }
        t1 = clock64();
        if (lane == 0) out[4] = (t1 - t0) / ITERS;
        t0 = clock64();
        for (int i = 0; i < ITERS; ++i) { tmem_ld32(tmem_addr(tb, 0, 0), v); tmem_ld_wait32(v); }
        t1 = clock64();
        if (lane == 0) out[5] = (t1 - t0) / ITERS + (v[0] & 0);
        uint32_t w16[16];
        for (int j = 0; j < 16; ++j) w16[j] = v[j];
        t0 = clock64();
        for (int i = 0; i < ITERS; ++i) { tmem_st16(tmem_addr(tb, 0, 64), w16); tmem_st_wait(); }
        t1 = clock64();
        if (lane == 0) out[6] = (t1 - t0) / ITERS;
        t0 = clock64();
        for (int i = 0; i < ITERS; ++i) { fence_before_sync(); fence_after_sync(); }
        t1 = clock64();
        if (lane == 0) out[7] = (t1 - t0) / ITERS;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 256);
}

int main() {
    long long* d;
    cudaMalloc(&d, 64);
    cudaMemset(d, 0, 64);
    cudaFuncSetAttribute(handoff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 256);
    handoff_kernel<<<1, 128, 65536 + 256>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("status %s\n", cudaGetErrorString(e));
    printf("mbarrier round trip (2 hand-offs), suspending try_wait: %lld cycles\n", h[0]);
    printf("mbarrier round trip (2 hand-offs), spinning test_wait : %lld cycles\n", h[1]);
    printf("MMA 128x64x16 + commit + wait, suspending            : %lld cycles\n", h[2]);
    printf("MMA 128x64x16 + commit + wait, spinning              : %lld cycles\n", h[3]);
    printf("st.shared.v4 + fence.proxy.async                     : %lld cycles\n", h[4]);
    printf("tcgen05.ld x32 + wait::ld                            : %lld cycles\n", h[5]);
    printf("tcgen05.st x16 + wait::st                            : %lld cycles\n", h[6]);
    printf("tcgen05.fence before + after thread sync             : %lld cycles\n", h[7]);
    return 0;
}
