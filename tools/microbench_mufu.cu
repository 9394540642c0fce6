// Microbenchmark: per-SM throughput of ex2 in f32 / f16x2 / bf16x2 form, f32x2 packed add, 3-input max.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench_mufu.cu && /tmp/mb
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
    uint32_t h0 = threadIdx.x, h1 = h0 + 77, h2 = h0 + 99, h3 = h0 + 3;
    unsigned long long p0 = threadIdx.x, p1 = p0 + 5;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
        } else if (MODE == 1) {
            asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
            asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
        } else if (MODE == 2) {
            asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h0)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h1));
            asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h2)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h3));
        } else if (MODE == 3) {
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(p1)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(p0));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(p1)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(p0));
        } else if (MODE == 4) {
            asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a0) : "f"(a1), "f"(a2)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a1) : "f"(a2), "f"(a3));
            asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a2) : "f"(a3), "f"(a0)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a3) : "f"(a0), "f"(a1));
        } else if (MODE == 5) {
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a0) : "f"(a1), "f"(a2)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a1) : "f"(a2), "f"(a3));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a2) : "f"(a3), "f"(a0)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a3) : "f"(a0), "f"(a1));
        } else if (MODE == 6) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p0) : "l"(p1)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p1) : "l"(p0));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p0) : "l"(p1)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p1) : "l"(p0));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + h0 + h1 + h2 + h3 + (float)p0 + (float)p1;
}

template <int MODE>
void run(const char* name, int elems_per_op) {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    const int iters = 20000, blocks = 148 * 2, threads = 1024;
    k<MODE><<<blocks, threads>>>(out, 10);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<blocks, threads>>>(out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double ops = (double)blocks * threads * iters * 4;
    printf("%-22s %8.3f ms  %7.2f Gop/s/SM  (%5.2f lane-ops/clk/SM at %d MHz nominal; %d element(s) per op)  err=%s\n", name, ms,
           ops / ms / 1e6 / 148, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000, elems_per_op, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    run<0>("ex2.f32", 1);
    run<1>("ex2.f16x2", 2);
    run<2>("ex2.bf16x2", 2);
    run<3>("add.f32x2", 2);
    run<4>("max3.f32", 1);
    run<5>("fma.f32", 1);
    run<6>("fma.f32x2", 2);
    return 0;
}
