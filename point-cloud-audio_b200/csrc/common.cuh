// Shared helpers for the pcaudio_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/pcaudio_b200.h"

namespace pca {

// thread-local error slot (re-entrant, one caller thread per GPU)
char* err_buf();
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);

// Optional per-kernel device timing (pca_profile_enable): brackets a launch with CUDA events on the
// launching stream and records the algorithmic flops / bytes the launcher states for it.
struct LaunchTimer {
    LaunchTimer(const char* name, cudaStream_t st, double flops, double bytes);
    ~LaunchTimer();
    int slot;
    cudaStream_t st;
};

#define PCA_CHECK_CUDA(expr)                                                  \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return ::pca::cuda_fail(_e, #expr);            \
    } while (0)

#define PCA_CHECK_LAUNCH(name)                                                \
    do {                                                                      \
        ::pca::count_launch();                                                \
        cudaError_t _e = cudaPeekAtLastError();                               \
        if (_e != cudaSuccess) return ::pca::cuda_fail(_e, name);             \
    } while (0)

#define PCA_TRY(expr)                                                         \
    do {                                                                      \
        int _r = (expr);                                                      \
        if (_r != 0) return _r;                                               \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Arena {
    char* base;
    size_t cap;
    size_t off;
    Arena(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
    template <typename T>
    T* take(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        if (base == nullptr) { off += bytes; return nullptr; }   // sizing pass
        if (off + bytes > cap) { off = (size_t)-1 / 2; return nullptr; }
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= cap; }
};

// packed MAB parameter blob (layout documented in include/pcaudio_b200.h)
struct MabParams {
    const float *Wq, *bq, *Wkv, *bkv, *Wo, *bo, *ln0w, *ln0b, *ln1w, *ln1b;
};
__host__ __device__ static inline long long mab_count(int dq, int dk, int D, int ln) {
    return (long long)D * dq + D + 2LL * D * dk + 2LL * D + (long long)D * D + D + (ln ? 4LL * D : 0);
}
__host__ __device__ static inline MabParams mab_slice(const float* p, int dq, int dk, int D, int ln) {
    MabParams m;
    m.Wq = p; p += (long long)D * dq;
    m.bq = p; p += D;
    m.Wkv = p; p += 2LL * D * dk;      // Wk then Wv: one (2D, dk) matrix
    m.bkv = p; p += 2LL * D;           // bk then bv
    m.Wo = p; p += (long long)D * D;
    m.bo = p; p += D;
    m.ln0w = m.ln0b = m.ln1w = m.ln1b = nullptr;
    if (ln) { m.ln0w = p; m.ln0b = p + D; m.ln1w = p + 2 * D; m.ln1b = p + 3 * D; }
    return m;
}

// fp32 encoder launchers shared by the inference (api.cu) and training (encoder_train.cu) orchestration
// img / img_bytes: optional scratch for the weight image of the tcgen05 GEMM (gemm_tc.cu); when it is given and the shape is
// eligible the layer runs on the tensor cores (split-bf16, fp32-grade), otherwise on the CUDA-core kernel.
int launch_linear(const float* X, const float* W, const float* b, float* Y, long long rows, int din, int dout, int mode,
                  cudaStream_t st, float* R = nullptr, void* img = nullptr, size_t img_bytes = 0);
size_t gemm_tc_image_bytes(int N, int K);
bool linear_tc_eligible(long long rows, int K, int N);
int launch_linear_tc(const float* X, const float* W, int trans_w, const float* bias, const float* resid, float* Y, float* R,
                     long long rows, int K, int N, int relu, void* img, size_t img_bytes, cudaStream_t st);
bool grad_weight_tc_eligible(long long rows, int M, int N);
int launch_grad_weight_tc(const float* dY, const float* X, float* dW, long long rows, int M, int N, cudaStream_t st);
void set_gemm_tc(int on);
int launch_attn(const float* Qp, long long q_bstride, const float* KV, int B, int nq, int nk, int D, int H, float* O, float* part,
                const int* key_counts, cudaStream_t st, float* lse = nullptr, float* p_out = nullptr);
size_t attn_part_floats(int B, int nq, int nk, int D, int H);
// attention with one small side (<= 16 queries or keys) as split-bf16 tcgen05 GEMMs (attn_tc.cu); scratch sizes in floats
bool attn_tc_eligible(int B, int nq, int nk, int D, int H);
int attn_tc_kind(int B, int nq, int nk, int D, int H);
size_t attn_tc_fwd_floats(int B, int nq, int nk, int D, int H);
size_t attn_tc_bwd_floats(int B, int nq, int nk, int D, int H);
int launch_attn_tc(const float* Qp, long long q_bstride, const float* KV, int B, int nq, int nk, int D, int H, float* O, float* scratch,
                   cudaStream_t st, float* lse, const int* key_counts = nullptr, float* p_out = nullptr);
size_t attn_tc_p_floats(int B, int nq, int nk, int D, int H);
// inference, shared small query set: the block on the un-projected points (W_k folded into the query image, W_v applied to the
// per-cloud sums P^T X); the K | V projection is never computed
bool attn_fold_eligible(int B, int nq, int nk, int dk, int D, int H);
int launch_attn_folded(const float* Qp, const float* Wkv, const float* bkv, const float* X, int B, int nq, int nk, int dk, int D, int H,
                       float* O, float* scratch, cudaStream_t st, const int* key_counts, float* p_out = nullptr, float* z_out = nullptr,
                       float* gq_out = nullptr);
// the same block in training: forward keeps P / Z / Gq, the backward below works on the un-projected points as well
bool attn_fold_train_on();
size_t attn_fold_z_floats(int B, int nq, int nk, int D, int H);
size_t attn_fold_gq_floats(int B, int nq, int nk, int D, int H);
int attn_fold_rows(int B, int nq, int nk, int D, int H);
int launch_attn_folded_bwd(const float* Qp, const float* Wkv, const float* X, const float* dO, const float* P, const float* Z,
                           const float* Gq, int B, int nq, int nk, int dk, int D, int H, float* dX, int x_acc, float* dWk, float* dQ,
                           float** dox_out, float* scratch, cudaStream_t st);
int launch_attn_bwd_tc(const float* Qp, long long q_bstride, const float* KV, const float* dO, const float* lse, const float* delta,
                       int B, int nq, int nk, int D, int H, float* dQp, float* dKV, float* scratch, cudaStream_t st,
                       const float* p_saved = nullptr, float* db_q = nullptr, float* db_kv = nullptr);
void set_attn_tc(int on);
int launch_layernorm(float* X, long long rows, int D, const float* g, const float* b, cudaStream_t st);

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace pca
