// Microbenchmark: ceiling of the per-item softmax instruction stream of the tcgen05 encoder kernels, without any
// MMA / mbarrier dependency.  One CTA per SM, W warps; every warp loops over "items" of 64 score columns:
//   tcgen05.ld x32 x2 -> row max -> 2^(s-m) -> bf16 pack -> tcgen05.st x16 x2
// Variants switch individual pieces off to see which unit binds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I point-cloud-audio_b200/csrc -o tools/mb_softmax tools/mb_softmax.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_prims.cuh"

using namespace pca::tc;

__device__ __forceinline__ float max_chunk32(const uint32_t* v, float mx) {
    float m0 = mx, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        m0 = max3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
        m1 = max3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
    return fmaxf(m0, m1);
}
template <uint32_t POLY>
__device__ __forceinline__ void exp_chunk32(const uint32_t* v, const float2 neg_m2, float2& sum2, uint32_t* pk) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        const float2 x = add2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), neg_m2);
        const float2 pr = ((POLY >> (j >> 1)) & 1u) ? ex2_poly2(x) : make_float2(ex2(x.x), ex2(x.y));
        sum2 = add2(sum2, pr);
        pk[j >> 1] = pack_bf16(pr.x, pr.y);
    }
}

// MODE bits: 1 = TMEM loads, 2 = max pass, 4 = exponentials (else plain multiply), 8 = TMEM stores, 16 = reload (1.5 passes)
template <int MODE, uint32_t POLY, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k(float* out, long long* clk_out, int iters) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = tmem_slot;
    const uint32_t sbase = tmem_addr(tb, 32 * (warp & 3), 64 * ((warp >> 2) & 7));
    uint32_t va[32], vb[32], pk[16];
#pragma unroll
    for (int j = 0; j < 32; ++j) { va[j] = __float_as_uint(threadIdx.x * 1e-3f + j); vb[j] = __float_as_uint(threadIdx.x * 2e-3f - j); }
    {   // initialise the TMEM region with finite values
        tmem_st16(sbase, va); tmem_st16(sbase + 16, va + 16); tmem_st16(sbase + 32, vb); tmem_st16(sbase + 48, vb + 16);
        tmem_st_wait();
    }
    float l = 0.f, m_run = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE & 1) {
            tmem_ld32(sbase, va);
            tmem_ld32(sbase + 32, vb);
            tmem_ld_wait64(va, vb);
        }
        float mx = m_run;
        if (MODE & 2) mx = fmaxf(max_chunk32(vb, max_chunk32(va, -INFINITY)), m_run);
        const float2 neg2 = make_float2(-mx, -mx);
        float2 sum2 = make_float2(0.f, 0.f);
        if (MODE & 4) exp_chunk32<POLY>(va, neg2, sum2, pk);
        else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float2 x = add2(make_float2(__uint_as_float(va[j]), __uint_as_float(va[j + 1])), neg2);
                sum2 = add2(sum2, x);
                pk[j >> 1] = pack_bf16(x.x, x.y);
            }
        }
        if (MODE & 8) tmem_st16(sbase, pk);
        else { uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) x ^= pk[j];
            va[0] ^= x & 1u; }
        if (MODE & 16) { tmem_ld32(sbase + 32, vb); tmem_ld_wait32(vb); }
        if (MODE & 4) exp_chunk32<POLY>(vb, neg2, sum2, pk);
        else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float2 x = add2(make_float2(__uint_as_float(vb[j]), __uint_as_float(vb[j + 1])), neg2);
                sum2 = add2(sum2, x);
                pk[j >> 1] = pack_bf16(x.x, x.y);
            }
        }
        if (MODE & 8) { tmem_st16(sbase + 16, pk); tmem_st_wait(); }
        else { uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) x ^= pk[j];
            vb[0] ^= x & 1u; }
        l += sum2.x + sum2.y;
        m_run = mx * 0.5f;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = l + m_run + __uint_as_float(va[3]) + __uint_as_float(vb[5]);
    if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <int MODE, uint32_t POLY, int warps>
void run(const char* name) {
    float* out; long long* clk;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&clk, 148 * 8);
    const int iters = 4000;
    k<MODE, POLY, warps * 32><<<148, warps * 32>>>(out, clk, 10);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE, POLY, warps * 32><<<148, warps * 32>>>(out, clk, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    const double cyc_item_sm = (double)h[0] / iters;                 // cycles for W warps to do one item each
    const double exps_per_clk = warps * 32.0 * 64.0 / cyc_item_sm;
    printf("%-44s warps=%2d  %8.1f clk per round  -> %6.2f elements/clk/SM   (%.3f ms) %s\n", name, warps, cyc_item_sm, exps_per_clk, ms,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(clk);
}

int main() {
#define RUNALL(W)                                                                 \
    run<1 | 2 | 4 | 8, 0, W>("ld + max + ex2 + st (full item)");               \
    run<1 | 2 | 4 | 8 | 16, 0, W>("full item, 1.5 TMEM passes");                \
    run<2 | 4 | 8, 0, W>("no ld");                                              \
    run<1 | 2 | 4, 0, W>("no st");                                              \
    run<2 | 4, 0, W>("max + ex2 only (registers)");                             \
    run<4, 0, W>("ex2 + sum + pack only");                                      \
    run<1 | 2 | 8, 0, W>("ld + max + (no ex2) + st");                           \
    run<1, 0, W>("ld only (+add/pack)");                                        \
    run<1 | 2 | 4 | 8, 0x5555u, W>("full item, 8/16 pairs polynomial");         \
    run<1 | 2 | 4 | 8, 0x1111u, W>("full item, 4/16 pairs polynomial");         \
    run<1 | 2 | 4 | 8, 0x0421u, W>("full item, 3/16 pairs polynomial");         \
    run<2 | 4, 0x5555u, W>("registers only, 8/16 polynomial");                  \
    run<2 | 4, 0xffffu, W>("registers only, all polynomial");
    RUNALL(4)
    RUNALL(8)
    RUNALL(16)
    return 0;
}
