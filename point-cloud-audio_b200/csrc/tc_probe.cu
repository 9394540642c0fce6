// Unit probe for the tcgen05 building blocks (descriptor encodings, operand layouts, TMEM A operand,
// commit/mbarrier handshake).  One 128-thread CTA computes D (128 x N) = A (128 x K) * B (K x N) with bf16
// operands / fp32 accumulation for every operand-source combination the encoder kernels use.
#include "common.cuh"
#include "tc_prims.cuh"

namespace pca {
using namespace tc;

// a_mode: 0 smem K-major (A given 128 x K), 1 TMEM (A given 128 x K), 2 smem MN-major (A given K x 128)
// b_mode: 0 smem K-major (B given N x K), 1 smem MN-major (B given K x N)
__global__ void __launch_bounds__(128)
umma_probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K,
                  int a_mode, int b_mode) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sa = smem;                  // up to 128 * 128 * 2 = 32 KB
    uint8_t* sb = smem + 32768;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t d_col = 0, a_col = 128;

    // ---- stage operands
    if (a_mode == 0) {
        for (int i = tid; i < 128 * K; i += 128) {
            const int r = i / K, k = i % K;
            *reinterpret_cast<__nv_bfloat16*>(sa + (k / 8) * 2048 + r * 16 + (k % 8) * 2) = __float2bfloat16(A[i]);
        }
    } else if (a_mode == 2) {
        for (int i = tid; i < 128 * K; i += 128) {
            const int k = i / 128, m = i % 128;          // A given as (K, 128)
            *reinterpret_cast<__nv_bfloat16*>(sa + (m / 8) * (K * 16) + k * 16 + (m % 8) * 2) = __float2bfloat16(A[i]);
        }
    } else {
        // thread = row: pack K bf16 into K/2 columns of TMEM lane tid
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = pack_bf16(A[tid * K + 2 * (c0 + j)], A[tid * K + 2 * (c0 + j) + 1]);
            tmem_st8(tmem_addr(tbase, (uint32_t)(warp * 32), a_col + c0), v);
        }
        tmem_st_wait();
    }
    if (b_mode == 0) {
        for (int i = tid; i < N * K; i += 128) {
            const int n = i / K, k = i % K;              // B given as (N, K)
            *reinterpret_cast<__nv_bfloat16*>(sb + (k / 8) * (N * 16) + n * 16 + (k % 8) * 2) = __float2bfloat16(B[i]);
        }
    } else {
        for (int i = tid; i < N * K; i += 128) {
            const int k = i / N, n = i % N;              // B given as (K, N)
            *reinterpret_cast<__nv_bfloat16*>(sb + (n / 8) * (K * 16) + k * 16 + (n % 8) * 2) = __float2bfloat16(B[i]);
        }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();

    if (tid == 0) {
        fence_after_sync();
        const uint32_t idesc = idesc_bf16(128, N, a_mode == 2, b_mode == 1);
        for (int ks = 0; ks < K / 16; ++ks) {
            uint64_t bd;
            if (b_mode == 0) bd = smem_desc(smem_u32(sb) + ks * 2 * (N * 16), N * 16, 128);
            else bd = smem_desc(smem_u32(sb) + ks * 256, 128, K * 16);
            if (a_mode == 1) {
                mma_ts(tmem_addr(tbase, 0, d_col), tmem_addr(tbase, 0, a_col + ks * 8), bd, idesc, ks > 0);
            } else {
                uint64_t ad;
                if (a_mode == 0) ad = smem_desc(smem_u32(sa) + ks * 2 * 2048, 2048, 128);
                else ad = smem_desc(smem_u32(sa) + ks * 256, 128, K * 16);
                mma_ss(tmem_addr(tbase, 0, d_col), ad, bd, idesc, ks > 0);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();

    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(tmem_addr(tbase, (uint32_t)(warp * 32), d_col + c0), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

int launch_umma_probe(const float* A, const float* B, float* D, int N, int K, int a_mode, int b_mode, cudaStream_t st) {
    if (!A || !B || !D) return fail(PCA_EINVAL, "umma_probe: null pointer");
    if (N < 16 || N > 128 || N % 16 || K < 16 || K > 128 || K % 16) return fail(PCA_EINVAL, "umma_probe: N, K must be multiples of 16 in [16, 128]");
    if (a_mode < 0 || a_mode > 2 || b_mode < 0 || b_mode > 1) return fail(PCA_EINVAL, "umma_probe: bad mode");
    PCA_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    umma_probe_kernel<<<1, 128, 65536, st>>>(A, B, D, N, K, a_mode, b_mode);
    PCA_CHECK_LAUNCH("umma_probe_kernel");
    return 0;
}

}  // namespace pca
