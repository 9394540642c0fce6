// extern "C" surface of libpcaudio_b200.so (see include/pcaudio_b200.h) and the host-side
// orchestration of the kernels.  No torch types, no device allocation, no global mutable state
// beyond the launch counter and the thread-local error slot.
#include "common.cuh"
#include "tma_host.cuh"
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace pca {

// ---- kernels / launchers defined in the other translation units
int launch_stft_logmag(const float*, int, int, int, int, const float*, const float*, float, int, int, float*, cudaStream_t);
int launch_build_clouds(const float*, int, int, int, const float*, const float*, float*, cudaStream_t);
int launch_topk(const float*, int, int, int, const float*, const float*, int, int, int, float, float*, int32_t*, int32_t*, cudaStream_t);
int launch_pool(const float*, int, int, int, int, float*, const int*, cudaStream_t);
int launch_fused_frontend(const float*, int, int, int, int, const float*, const float*, float, int, int, const float*, const float*,
                          int, int, int, float, float*, int32_t*, int32_t*, cudaStream_t);
// subsampling (sampling.cu)
int launch_random_keys(float*, long long, unsigned long long, cudaStream_t);
int launch_gather_points(const float*, int, int, int, const float*, const float*, const int32_t*, int, float*, cudaStream_t);
int launch_importance_map(const float*, int, int, int, const float*, int, const float*, int, float*, float*, cudaStream_t);
int launch_multinomial(const float*, int, int, int, unsigned long long, double*, int32_t*, cudaStream_t);
int launch_resample(const float*, int, int, int, double, const double*, const double*, int, int, float, float*, cudaStream_t);
// training path (encoder_train.cu)
size_t st_train_saved_bytes(const pca_st_dims* d, int B, int N, float dropout_p);
size_t st_train_ws_bytes(const pca_st_dims* d, int B, int N);
int st_train_forward(const float*, const int*, int, int, const pca_st_dims*, const float*, float, unsigned long long, float*, void*, size_t,
                     void*, size_t, cudaStream_t);
int st_train_backward(const float*, const int*, int, int, const pca_st_dims*, const float*, float, unsigned long long, const float*,
                      const void*, size_t, float*, float*, void*, size_t, cudaStream_t, int phase = 0, long long* tail_offset = nullptr);
size_t mab_train_saved_bytes(int B, int qb, int nq, int nk, int D, int H, int ln);
size_t mab_train_ws_bytes(int B, int qb, int nq, int nk, int D, int H, int ln);
int mab_train_forward_api(const float*, int, const float*, int, int, int, int, int, int, int, int, const float*, float*, void*, size_t,
                          void*, size_t, cudaStream_t);
int mab_train_backward_api(const float*, int, const float*, int, int, int, int, int, int, int, int, const float*, const float*,
                           const void*, size_t, float*, float*, float*, void*, size_t, cudaStream_t);
size_t deepset_train_saved_bytes(int B, int N, int dh);
size_t deepset_train_ws_bytes(int B, int N, int dh);
int deepset_train_forward(const float*, int, int, int, int, int, int, const float*, float*, void*, size_t, void*, size_t, cudaStream_t);
int deepset_train_backward(const float*, int, int, int, int, int, int, const float*, const float*, const void*, size_t, float*, float*,
                           void*, size_t, cudaStream_t);
int launch_cross_entropy(const float*, const long long*, int, int, float, float*, int*, float*, cudaStream_t);
int dropout_api(const float*, float*, long long, float, unsigned long long, cudaStream_t);
int linear_bwd_api(const float*, const float*, long long, int, int, const float*, float*, float*, cudaStream_t);
int launch_adam(float*, const float*, float*, float*, long long, float, float, float, float, float, int, float, cudaStream_t);
void set_timeline(long long* p);
void set_tail_max(int t);
void set_reduce_wg(int n);
void set_pool_variant(int v);
void debug_set_stft_generic(int on);
int launch_umma_probe(const float*, const float*, float*, int, int, int, int, cudaStream_t);
// tcgen05 path (encoder_tc.cu)
size_t st_tc_workspace_bytes(const pca_st_dims* d, int B, int N);
int st_tc_supported(const pca_st_dims* d, int N);
int st_tc_forward(const float* X, const int* counts, int B, int N, const pca_st_dims* d, const float* params, float* logits,
                  void* ws, size_t ws_bytes, cudaStream_t st);
int st_tc_accepts_logmag();
int st_tc_forward_logmag(const float* logmag, const float* farr, const float* tarr, int nf, int B, int N, const pca_st_dims* d,
                         const float* params, float* logits, void* ws, size_t ws_bytes, cudaStream_t st);
int st_tc_forward_stages(const float* X, int B, int N, const pca_st_dims* d, const float* params, float* logits,
                         float* H1, float* Y1, float* H2, float* Y2, float* pooled, void* ws, size_t ws_bytes,
                         cudaStream_t st);

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

char* err_buf() { return g_err; }
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

// ------------------------------------------------------------------------------------ TMA tensor maps (tma_host.cuh)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static std::atomic<void*> cached{nullptr};
    void* f = cached.load(std::memory_order_acquire);
    if (f == nullptr) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        cached.store(f, std::memory_order_release);
    }
    return reinterpret_cast<EncodeTiledFn>(f);
}
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, unsigned long long inner, unsigned long long rows,
                      unsigned long long row_stride_bytes, unsigned box_inner, unsigned box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return fail(PCA_EDEVICE, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {inner, rows};
    const cuuint64_t strides[1] = {row_stride_bytes};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PCA_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d (base %p, %llu x %llu, stride %llu)", (int)r, base, inner, rows, row_stride_bytes);
    return 0;
}

// The same (rows, 64) bf16 tensor seen as (8 elements, rows, 8 chunks) with byte strides (2, row_stride, 16): ONE box of
// (8, box_rows, 8) then lands in shared memory as [chunk][row][16 B] -- the canonical no-swizzle UMMA operand layout of a
// 64-wide tile -- where the 2-D view needs eight 16-byte-wide boxes (measured ~90 cycles of issue per TMA instruction).
int make_tmap_chunked_bf16(CUtensorMap* out, const void* base, unsigned long long rows, unsigned long long row_stride_bytes,
                           unsigned box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return fail(PCA_EDEVICE, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {8, rows, 8};
    const cuuint64_t strides[2] = {row_stride_bytes, 16};
    const cuuint32_t box[3] = {8, box_rows, 8};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PCA_EINVAL, "cuTensorMapEncodeTiled (chunked view) failed with CUresult %d (base %p, %llu rows, stride %llu)", (int)r, base, rows, row_stride_bytes);
    return 0;
}

int make_tmap_3d_f32(CUtensorMap* out, const void* base, unsigned long long inner, unsigned long long rows, unsigned long long batch,
                     unsigned long long row_stride_bytes, unsigned long long batch_stride_bytes, unsigned box_inner, unsigned box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return fail(PCA_EDEVICE, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {inner, rows, batch};
    const cuuint64_t strides[2] = {row_stride_bytes, batch_stride_bytes};
    const cuuint32_t box[3] = {box_inner, box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(PCA_EINVAL, "cuTensorMapEncodeTiled (fp32 rows) failed with CUresult %d (base %p, %llu x %llu x %llu, strides %llu / %llu)",
                    (int)r, base, inner, rows, batch, row_stride_bytes, batch_stride_bytes);
    return 0;
}

// ------------------------------------------------------------------------------------ profiling
struct ProfRec { const char* name; cudaEvent_t a, b; double flops, bytes; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;

LaunchTimer::LaunchTimer(const char* name, cudaStream_t s, double flops, double bytes) : slot(-1), st(s) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec r{name, nullptr, nullptr, flops, bytes};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
    slot = (int)g_prof.size() - 1;
}
LaunchTimer::~LaunchTimer() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].b, st);
}

// ------------------------------------------------------------------------------------ params
// ------------------------------------------------------------------------------------ MAB (fp32)
// workspace: Qp (qb*nq*D) | KV (B*nk*2D) | O (B*nq*D) | part
// scratch for the largest weight image of a MAB's tensor-core GEMMs (Wkv with dk <= max(D, 4) ... D x D is the largest eligible)
static size_t mab_img_bytes(int D) { return gemm_tc_image_bytes(2 * D, D); }
static size_t mab_ws_floats(int B, int qb, int nq, int nk, int D, int H) {
    Arena a(nullptr, 0);
    a.take<float>((size_t)qb * nq * D);
    a.take<float>((size_t)B * nk * 2 * D);
    a.take<float>((size_t)B * nq * D);
    a.take<float>(attn_part_floats(B, nq, nk, D, H));
    a.take<uint8_t>(mab_img_bytes(D));
    return a.off;
}

// Q (qb, nq, dq) with qb in {1, B}; K (B, nk, dk); out (B, nq, D)
static int mab_forward(const float* Q, int qb, const float* K, int B, int nq, int nk, int dq, int dk, int D,
                       int H, int ln, const float* params, float* out, void* ws, size_t ws_bytes,
                       cudaStream_t st, const int* key_counts = nullptr) {
    if (qb != 1 && qb != B) return fail(PCA_EINVAL, "MAB: query batch must be 1 or B");
    if (nq <= 0 || nk <= 0) return fail(PCA_EINVAL, "MAB: empty query or key set (nq=%d, nk=%d)", nq, nk);
    const MabParams m = mab_slice(params, dq, dk, D, ln);
    Arena a(ws, ws_bytes);
    float* Qp = a.take<float>((size_t)qb * nq * D);
    float* KV = a.take<float>((size_t)B * nk * 2 * D);
    float* O = a.take<float>((size_t)B * nq * D);
    float* part = a.take<float>(attn_part_floats(B, nq, nk, D, H));
    const size_t ib = mab_img_bytes(D);
    void* img = a.take<uint8_t>(ib);
    if (!a.ok()) return fail(PCA_EWORKSPACE, "MAB: workspace too small (%zu bytes given)", ws_bytes);
    PCA_TRY(launch_linear(Q, m.Wq, m.bq, Qp, (long long)qb * nq, dq, D, 0, st, nullptr, dq <= D ? img : nullptr, ib));
    if (qb == 1 && attn_fold_eligible(B, nq, nk, dk, D, H)) {
        // shared inducing points / seeds against many points: the attention runs on the un-projected keys (attn_tc.cu)
        PCA_TRY(launch_attn_folded(Qp, m.Wkv, m.bkv, K, B, nq, nk, dk, D, H, O, part, st, key_counts));
    } else {
        PCA_TRY(launch_linear(K, m.Wkv, m.bkv, KV, (long long)B * nk, dk, 2 * D, 0, st, nullptr, dk <= D ? img : nullptr, ib));
        PCA_TRY(launch_attn(Qp, qb == 1 ? 0 : (long long)nq * D, KV, B, nq, nk, D, H, O, part, key_counts, st));
    }
    if (ln) PCA_TRY(launch_layernorm(O, (long long)B * nq, D, m.ln0w, m.ln0b, st));
    PCA_TRY(launch_linear(O, m.Wo, m.bo, out, (long long)B * nq, D, D, 2, st, nullptr, img, ib));
    if (ln) PCA_TRY(launch_layernorm(out, (long long)B * nq, D, m.ln1w, m.ln1b, st));
    return 0;
}

// ------------------------------------------------------------------------------------ ISAB / PMA / ST (fp32)
static size_t isab_ws_bytes(int B, int N, int d_in, int D, int H, int M) {
    Arena a(nullptr, 0);
    a.take<float>((size_t)B * M * D);   // H
    size_t w0 = mab_ws_floats(B, 1, M, N, D, H), w1 = mab_ws_floats(B, B, N, M, D, H);
    return a.off + (w0 > w1 ? w0 : w1);
}
static int isab_forward(const float* X, int B, int N, int d_in, int D, int H, int M, int ln,
                        const float* params, float* out, void* ws, size_t ws_bytes, cudaStream_t st,
                        const int* counts = nullptr) {
    const float* I = params;
    const float* p0 = I + (long long)M * D;
    const float* p1 = p0 + mab_count(D, d_in, D, ln);
    Arena a(ws, ws_bytes);
    float* Hb = a.take<float>((size_t)B * M * D);
    if (!a.ok()) return fail(PCA_EWORKSPACE, "ISAB: workspace too small");
    void* sub = (char*)ws + a.off;
    const size_t sub_bytes = ws_bytes - a.off;
    PCA_TRY(mab_forward(I, 1, X, B, M, N, D, d_in, D, H, ln, p0, Hb, sub, sub_bytes, st, counts));   // mab0(I, X): padded points masked
    PCA_TRY(mab_forward(X, B, Hb, B, N, M, d_in, D, D, H, ln, p1, out, sub, sub_bytes, st));    // mab1(X, H)
    return 0;
}
static size_t pma_ws_bytes(int B, int N, int D, int H, int S) { return mab_ws_floats(B, 1, S, N, D, H); }
static int pma_forward(const float* X, int B, int N, int D, int H, int S, int ln, const float* params,
                       float* out, void* ws, size_t ws_bytes, cudaStream_t st, const int* counts = nullptr) {
    const float* Sd = params;
    return mab_forward(Sd, 1, X, B, S, N, D, D, D, H, ln, Sd + (long long)S * D, out, ws, ws_bytes, st, counts);
}

static long long st_count(const pca_st_dims* d) {
    return (long long)d->M * d->D + mab_count(d->D, d->d_in, d->D, d->ln) + mab_count(d->d_in, d->D, d->D, d->ln) +
           (long long)d->M * d->D + mab_count(d->D, d->D, d->D, d->ln) + mab_count(d->D, d->D, d->D, d->ln) +
           (long long)d->S * d->D + mab_count(d->D, d->D, d->D, d->ln) + (long long)d->C * d->D + d->C;
}

static size_t st_f32_ws_bytes(const pca_st_dims* d, int B, int N) {
    Arena a(nullptr, 0);
    a.take<float>((size_t)B * N * d->D);   // Y1
    a.take<float>((size_t)B * N * d->D);   // Y2
    a.take<float>((size_t)B * d->S * d->D);
    size_t w0 = isab_ws_bytes(B, N, d->d_in, d->D, d->H, d->M);
    size_t w1 = isab_ws_bytes(B, N, d->D, d->D, d->H, d->M);
    size_t w2 = pma_ws_bytes(B, N, d->D, d->H, d->S);
    size_t w = w0 > w1 ? w0 : w1;
    w = w > w2 ? w : w2;
    return a.off + w;
}

static int st_f32_forward_chunk(const float* X, const int* counts, int B, int N, const pca_st_dims* d, const float* params,
                                float* logits, void* ws, size_t ws_bytes, cudaStream_t st) {
    const int D = d->D, H = d->H, M = d->M, S = d->S, C = d->C, ln = d->ln;
    const float* p_isab0 = params;
    const float* p_isab1 = p_isab0 + (long long)M * D + mab_count(D, d->d_in, D, ln) + mab_count(d->d_in, D, D, ln);
    const float* p_pma = p_isab1 + (long long)M * D + 2 * mab_count(D, D, D, ln);
    const float* p_lin = p_pma + (long long)S * D + mab_count(D, D, D, ln);
    Arena a(ws, ws_bytes);
    float* Y1 = a.take<float>((size_t)B * N * D);
    float* Y2 = a.take<float>((size_t)B * N * D);
    float* P = a.take<float>((size_t)B * S * D);
    if (!a.ok()) return fail(PCA_EWORKSPACE, "ST: workspace too small");
    void* sub = (char*)ws + a.off;
    const size_t sub_bytes = ws_bytes - a.off;
    PCA_TRY(isab_forward(X, B, N, d->d_in, D, H, M, ln, p_isab0, Y1, sub, sub_bytes, st, counts));
    PCA_TRY(isab_forward(Y1, B, N, D, D, H, M, ln, p_isab1, Y2, sub, sub_bytes, st, counts));
    PCA_TRY(pma_forward(Y2, B, N, D, H, S, ln, p_pma, P, sub, sub_bytes, st, counts));
    PCA_TRY(launch_linear(P, p_lin, p_lin + (long long)C * D, logits, (long long)B * S, D, C, 0, st));
    return 0;
}

static int check_dims(const pca_st_dims* d) {
    if (!d) return fail(PCA_EINVAL, "ST: null dims");
    if (d->d_in <= 0 || d->D <= 0 || d->H <= 0 || d->M <= 0 || d->S <= 0 || d->C <= 0)
        return fail(PCA_EINVAL, "ST: non-positive dimension");
    if (d->D % d->H) return fail(PCA_EINVAL, "ST: dim_hidden %d not divisible by num_heads %d", d->D, d->H);
    if (d->D % 4) return fail(PCA_EUNSUPPORTED, "ST: dim_hidden must be a multiple of 4");
    return 0;
}

// Variable-size sets: every kernel clamps counts[b] to [1, N] (a set needs a key to attend to), so a cloud with NO valid point
// would silently be encoded as the one-point cloud of its first (padding) row.  The reference has no answer for an empty
// set either (it fails on X[:, :0]); the honest answer here is NaN logits for exactly those clouds.
__global__ void nan_empty_sets_kernel(float* __restrict__ logits, const int* __restrict__ counts, int B, int per_cloud) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)B * per_cloud && counts[i / per_cloud] <= 0) logits[i] = __int_as_float(0x7fc00000);
}
static int st_forward_impl(const float* X, int B, int N, const pca_st_dims* d, const float* params, float* logits,
                           void* ws, size_t ws_bytes, int precision, cudaStream_t st, const int* counts);
static int st_forward(const float* X, int B, int N, const pca_st_dims* d, const float* params, float* logits,
                      void* ws, size_t ws_bytes, int precision, cudaStream_t st, const int* counts = nullptr) {
    PCA_TRY(st_forward_impl(X, B, N, d, params, logits, ws, ws_bytes, precision, st, counts));
    if (counts != nullptr && B > 0) {
        const long long n = (long long)B * d->S * d->C;
        nan_empty_sets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(logits, counts, B, d->S * d->C);
        PCA_CHECK_LAUNCH("nan_empty_sets_kernel");
    }
    return 0;
}
static int st_forward_impl(const float* X, int B, int N, const pca_st_dims* d, const float* params, float* logits,
                           void* ws, size_t ws_bytes, int precision, cudaStream_t st, const int* counts) {
    PCA_TRY(check_dims(d));
    if (!X || !params || !logits) return fail(PCA_EINVAL, "ST: null pointer");
    if (B < 0 || N <= 0) return fail(PCA_EINVAL, "ST: bad batch/points (B=%d, N=%d)", B, N);
    if (B == 0) return 0;
    if (precision == PCA_PREC_BF16) {
        if (!st_tc_supported(d, N))
            return fail(PCA_EUNSUPPORTED, "ST: tcgen05 path needs D=64,H=8,M=64,S=1,ln=0,d_in<=4 (got D=%d,H=%d,M=%d,S=%d,ln=%d,d_in=%d)",
                        d->D, d->H, d->M, d->S, d->ln, d->d_in);
        return st_tc_forward(X, counts, B, N, d, params, logits, ws, ws_bytes, st);
    }
    if (precision != PCA_PREC_FP32) return fail(PCA_EINVAL, "ST: unknown precision %d", precision);
    // largest chunk of clouds whose scratch fits the caller's workspace
    const size_t one = st_f32_ws_bytes(d, 1, N);
    if (ws_bytes < one || !ws) return fail(PCA_EWORKSPACE, "ST: workspace %zu B < minimum %zu B", ws_bytes, one);
    int chunk = B;
    while (chunk > 1 && st_f32_ws_bytes(d, chunk, N) > ws_bytes) chunk = (chunk + 1) / 2;
    if (chunk > 32768) chunk = 32768;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int bc = (B - b0) < chunk ? (B - b0) : chunk;
        PCA_TRY(st_f32_forward_chunk(X + (long long)b0 * N * d->d_in, counts ? counts + b0 : nullptr, bc, N, d, params,
                                     logits + (long long)b0 * d->S * d->C, ws, ws_bytes, st));
    }
    return 0;
}

// ------------------------------------------------------------------------------------ DeepSet
static size_t deepset_ws_bytes(int B, int N, int dh) {
    Arena a(nullptr, 0);
    a.take<float>((size_t)B * N * dh);
    a.take<float>((size_t)B * N * dh);
    a.take<float>((size_t)B * dh);
    a.take<float>((size_t)B * dh);
    a.take<uint8_t>(gemm_tc_image_bytes(dh, dh));
    return a.off;
}
static int deepset_forward(const float* X, int B, int N, int d_in, int dh, int out_dim, int pool,
                           const float* p, float* out, void* ws, size_t ws_bytes, cudaStream_t st,
                           const int* counts = nullptr) {
    Arena a(ws, ws_bytes);
    float* t0 = a.take<float>((size_t)B * N * dh);
    float* t1 = a.take<float>((size_t)B * N * dh);
    float* u0 = a.take<float>((size_t)B * dh);
    float* u1 = a.take<float>((size_t)B * dh);
    const size_t ib = gemm_tc_image_bytes(dh, dh);
    void* img = a.take<uint8_t>(ib);
    if (!a.ok() || !ws) return fail(PCA_EWORKSPACE, "DeepSet: workspace too small");
    const long long rows = (long long)B * N;
    const float* W[8]; const float* bb[8];
    for (int i = 0; i < 8; ++i) {
        const int dout = (i == 7) ? out_dim : dh;
        const int di = (i == 0) ? d_in : dh;
        W[i] = p; p += (long long)dout * di;
        bb[i] = p; p += dout;
    }
    PCA_TRY(launch_linear(X, W[0], bb[0], t0, rows, d_in, dh, 1, st));
    PCA_TRY(launch_linear(t0, W[1], bb[1], t1, rows, dh, dh, 1, st, nullptr, img, ib));      // shared MLP over points: tensor cores
    PCA_TRY(launch_linear(t1, W[2], bb[2], t0, rows, dh, dh, 1, st, nullptr, img, ib));
    PCA_TRY(launch_linear(t0, W[3], bb[3], t1, rows, dh, dh, 0, st, nullptr, img, ib));
    PCA_TRY(launch_pool(t1, B, N, dh, pool, u0, counts, st));
    PCA_TRY(launch_linear(u0, W[4], bb[4], u1, B, dh, dh, 1, st));
    PCA_TRY(launch_linear(u1, W[5], bb[5], u0, B, dh, dh, 1, st));
    PCA_TRY(launch_linear(u0, W[6], bb[6], u1, B, dh, dh, 1, st));
    PCA_TRY(launch_linear(u1, W[7], bb[7], out, B, dh, out_dim, 0, st));
    return 0;
}

// ------------------------------------------------------------------------------------ pipeline
struct PipeShape { int nt_all, nt_out, nf, clouds_per_clip, pts_full, pts; int width; };
static int pipe_shape(const pca_pipeline_cfg* c, PipeShape* s) {
    if (!c) return fail(PCA_EINVAL, "pipeline: null cfg");
    if (c->hop <= 0 || c->n_fft <= 0 || c->n_samples <= 0) return fail(PCA_EINVAL, "pipeline: bad STFT config");
    s->nt_all = 1 + c->n_samples / c->hop;
    if (c->mode == 2) {
        s->nf = c->n_fft / 2 + 1;            // FST keeps the Nyquist bin (Code/settransformer.py:40)
        s->nt_out = s->nt_all;
        s->clouds_per_clip = s->nt_all;
        s->pts_full = s->nf;
        s->width = 2;
    } else if (c->mode == 3) {
        if (c->ntemp <= 0) return fail(PCA_EINVAL, "pipeline: mode 3 needs ntemp > 0");
        s->nf = c->n_fft / 2;                // 3ST drops it (Code/settransformertemp.py:52)
        s->clouds_per_clip = s->nt_all / c->ntemp;
        s->nt_out = s->clouds_per_clip * c->ntemp;
        s->pts_full = s->nf * c->ntemp;
        s->width = 3;
    } else {
        return fail(PCA_EINVAL, "pipeline: mode must be 2 (FST) or 3 (3ST)");
    }
    if (c->top_k < 0 || c->top_k > s->pts_full) return fail(PCA_EINVAL, "pipeline: top_k outside [0, %d]", s->pts_full);
    s->pts = c->top_k ? c->top_k : s->pts_full;
    if (c->st.d_in != s->width) return fail(PCA_EINVAL, "pipeline: mode %d clouds are %d-wide but st.d_in=%d", c->mode, s->width, c->st.d_in);
    return 0;
}

static size_t st_any_ws_bytes(const pca_st_dims* d, int B, int N, int precision) {
    if (precision == PCA_PREC_BF16) return st_tc_workspace_bytes(d, B, N);
    return st_f32_ws_bytes(d, B, N);
}

static size_t pipe_ws_bytes(const pca_pipeline_cfg* c, const PipeShape& s, int n_clips, size_t* st_off) {
    Arena a(nullptr, 0);
    const size_t n_clouds = (size_t)n_clips * s.clouds_per_clip;
    a.take<float>((size_t)n_clips * s.nt_out * s.nf);        // log-magnitudes
    a.take<float>(n_clouds * s.pts * s.width);               // clouds
    a.take<int32_t>(n_clouds);                               // kept points per cloud (threshold mode)
    if (st_off) *st_off = a.off;
    // encoder scratch: chunks of <= 256 clouds for the fp32 path (more only helps launch overhead); the
    // tcgen05 path runs one CTA per cloud, so it gets up to 4096 clouds per launch to fill many waves
    const size_t cap = c->precision == PCA_PREC_BF16 ? 4096 : 256;
    int chunk = (int)(n_clouds < cap ? n_clouds : cap);
    if (chunk < 1) chunk = 1;
    return a.off + st_any_ws_bytes(&c->st, chunk, s.pts, c->precision);
}

static int pipeline_run(const pca_pipeline_cfg* c, const float* audio, int n_clips, const float* window,
                        const float* twiddle, const float* farr, const float* tarr, const float* st_params,
                        float* logits, void* ws, size_t ws_bytes, cudaStream_t st) {
    PipeShape s;
    PCA_TRY(pipe_shape(c, &s));
    if (n_clips < 0) return fail(PCA_EINVAL, "pipeline: negative clip count");
    if (n_clips == 0 || s.clouds_per_clip == 0) return 0;
    if (s.width == 3 && !tarr) return fail(PCA_EINVAL, "pipeline: mode 3 needs tarr");
    size_t st_off = 0;
    const size_t need = pipe_ws_bytes(c, s, n_clips, &st_off);
    if (!ws || ws_bytes < need) return fail(PCA_EWORKSPACE, "pipeline: workspace %zu B < %zu B", ws_bytes, need);
    Arena a(ws, ws_bytes);
    const int n_clouds = n_clips * s.clouds_per_clip;
    float* logmag = a.take<float>((size_t)n_clips * s.nt_out * s.nf);
    float* pts = a.take<float>((size_t)n_clouds * s.pts * s.width);
    int32_t* kept = a.take<int32_t>((size_t)n_clouds);
    const bool thr = c->use_threshold != 0;
    // 3-D clouds with a selection step: one fused launch (the log-magnitudes stay in shared memory) when the cloud fits
    // (opt-in: measured slower than the two-kernel route on B200 while the batch's log-magnitudes fit the L2)
    if (c->mode == 3 && (c->top_k || thr) && getenv("PCA_FUSED_FRONTEND") != nullptr) {
        int kpad = 2;
        while (kpad < s.pts) kpad <<= 1;
        const size_t smem = (size_t)kpad * 8 + ((size_t)(c->n_fft / 2) + 8 * (size_t)(c->n_fft / 2 + c->n_fft / 32)) * 8 + (size_t)c->n_fft * 4 +
                            (size_t)s.pts_full * 4;      // 8 padded FFT buffers: launch_fused_frontend
        if (smem <= 227 * 1024 && s.pts <= 16384) {
            PCA_TRY(launch_fused_frontend(audio, n_clips, c->n_samples, c->n_fft, c->hop, window, twiddle, c->scale, 1, c->ntemp,
                                          farr, tarr, s.pts, 1, thr, c->threshold, pts, nullptr, thr ? kept : nullptr, st));
            return st_forward(pts, n_clouds, s.pts, &c->st, st_params, logits, (char*)ws + st_off, ws_bytes - st_off,
                              c->precision, st, thr ? kept : nullptr);
        }
    }
    PCA_TRY(launch_stft_logmag(audio, n_clips, c->n_samples, c->n_fft, c->hop, window, twiddle, c->scale,
                               c->mode == 3, s.nt_out, logmag, st));
    const int nt_cloud = c->mode == 3 ? c->ntemp : 1;
    const float* tarr_use = c->mode == 3 ? tarr : nullptr;
    // threshold mode: points with log-magnitude >= threshold, capped at top_k (or all points), padded; the encoder then
    // runs on variable-size sets (kept[] points per cloud)
    if (c->top_k || thr)
        PCA_TRY(launch_topk(logmag, n_clouds, s.nf, nt_cloud, farr, tarr_use, s.pts, 1, thr, c->threshold, pts, nullptr,
                            thr ? kept : nullptr, st));
    else if (c->precision == PCA_PREC_BF16 && st_tc_supported(&c->st, s.pts) && st_tc_accepts_logmag() && getenv("PCA_BUILD_CLOUDS") == nullptr)
        // front end fused into the encoder: its loader warps synthesise (f, [t,] mag) from the log-magnitudes and the coordinate
        // tables, the ESC_pc / ESC_pc_temp rows are never written (bit-identical to the build_clouds route, tested)
        return st_tc_forward_logmag(logmag, farr, tarr_use, s.nf, n_clouds, s.pts, &c->st, st_params, logits, (char*)ws + st_off,
                                    ws_bytes - st_off, st);
    else
        PCA_TRY(launch_build_clouds(logmag, n_clouds, s.nf, nt_cloud, farr, tarr_use, pts, st));
    return st_forward(pts, n_clouds, s.pts, &c->st, st_params, logits, (char*)ws + st_off, ws_bytes - st_off,
                      c->precision, st, thr ? kept : nullptr);
}

}  // namespace pca

// ====================================================================================== C ABI
using namespace pca;

extern "C" {

int pca_version(void) { return PCA_VERSION; }
const char* pca_last_error(void) { return err_buf(); }
unsigned long long pca_launch_count(void) { return g_launches.load(); }

void pca_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on.store(on ? 1 : 0);
}

int pca_profile_report(char* buf, size_t buf_len) {
    if (!buf || buf_len < 3) return fail(PCA_EINVAL, "profile_report: buffer too small");
    struct Agg { long long n = 0; double ms = 0, flops = 0, bytes = 0; };
    std::map<std::string, Agg> agg;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        for (auto& r : g_prof) {
            float ms = 0.f;
            if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
                Agg& a = agg[r.name];
                a.n += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
            }
            cudaEventDestroy(r.a); cudaEventDestroy(r.b);
        }
        g_prof.clear();
    }
    std::string out = "{";
    bool first = true;
    for (auto& kv : agg) {
        char line[256];
        snprintf(line, sizeof(line), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}",
                 first ? "" : ", ", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops, kv.second.bytes);
        out += line;
        first = false;
    }
    out += "}";
    if (out.size() + 1 > buf_len) return fail(PCA_EINVAL, "profile_report: need %zu bytes", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

int pca_stft_logmag_f32(const float* audio, int n_clips, int n_samples, int n_fft, int hop,
                        const float* window, const float* twiddle, float scale, int drop_nyquist,
                        int nt_out, float* out, void* stream) {
    return launch_stft_logmag(audio, n_clips, n_samples, n_fft, hop, window, twiddle, scale, drop_nyquist,
                              nt_out, out, (cudaStream_t)stream);
}

int pca_build_clouds_f32(const float* logmag, int n_clouds, int nf, int nt, const float* farr,
                         const float* tarr, float* pts, void* stream) {
    return launch_build_clouds(logmag, n_clouds, nf, nt, farr, tarr, pts, (cudaStream_t)stream);
}

int pca_topk_compact_f32(const float* keys, int n_clouds, int nf, int nt, const float* farr,
                         const float* tarr, int K, int sorted_desc, float* pts, int32_t* idx, void* stream) {
    return launch_topk(keys, n_clouds, nf, nt, farr, tarr, K, sorted_desc, 0, 0.f, pts, idx, nullptr, (cudaStream_t)stream);
}
int pca_select_compact_f32(const float* keys, int n_clouds, int nf, int nt, const float* farr, const float* tarr, int K,
                           int sorted_desc, int use_threshold, float threshold, float* pts, int32_t* idx,
                           int32_t* counts, void* stream) {
    return launch_topk(keys, n_clouds, nf, nt, farr, tarr, K, sorted_desc, use_threshold, threshold, pts, idx, counts,
                       (cudaStream_t)stream);
}

int pca_frontend_fused_f32(const float* audio, int n_clips, int n_samples, int n_fft, int hop, const float* window,
                           const float* twiddle, float scale, int drop_nyquist, int ntemp, const float* farr,
                           const float* tarr, int K, int sorted_desc, int use_threshold, float threshold, float* pts,
                           int32_t* idx, int32_t* counts, void* stream) {
    return launch_fused_frontend(audio, n_clips, n_samples, n_fft, hop, window, twiddle, scale, drop_nyquist, ntemp, farr, tarr,
                                 K, sorted_desc, use_threshold, threshold, pts, idx, counts, (cudaStream_t)stream);
}

long long pca_mab_param_count(int dq, int dk, int D, int ln) { return mab_count(dq, dk, D, ln); }
long long pca_isab_param_count(int d_in, int D, int M, int ln) {
    return (long long)M * D + mab_count(D, d_in, D, ln) + mab_count(d_in, D, D, ln);
}
long long pca_pma_param_count(int D, int S, int ln) { return (long long)S * D + mab_count(D, D, D, ln); }
long long pca_st_param_count(const pca_st_dims* dims) { return dims ? st_count(dims) : -1; }

size_t pca_mab_workspace_bytes(int B, int nq, int nk, int dq, int dk, int D, int H) {
    (void)dq; (void)dk;
    return mab_ws_floats(B, B, nq, nk, D, H);
}
int pca_mab_fwd_f32(const float* Q, int q_batch, const float* K, int B, int nq, int nk, int dq, int dk,
                    int D, int H, int ln, const float* params, float* out, void* workspace,
                    size_t workspace_bytes, void* stream) {
    if (!Q || !K || !params || !out) return fail(PCA_EINVAL, "MAB: null pointer");
    if (B == 0) return 0;
    if (D % 4 || D % H) return fail(PCA_EUNSUPPORTED, "MAB: need D %% 4 == 0 and D %% H == 0");
    return mab_forward(Q, q_batch, K, B, nq, nk, dq, dk, D, H, ln, params, out, workspace, workspace_bytes,
                       (cudaStream_t)stream);
}

size_t pca_isab_workspace_bytes(int B, int N, int d_in, int D, int H, int M) { return isab_ws_bytes(B, N, d_in, D, H, M); }
int pca_isab_fwd_f32(const float* X, int B, int N, int d_in, int D, int H, int M, int ln, const float* params,
                     float* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!X || !params || !out) return fail(PCA_EINVAL, "ISAB: null pointer");
    if (B == 0) return 0;
    if (N <= 0) return fail(PCA_EINVAL, "ISAB: empty set");
    if (D % 4 || D % H) return fail(PCA_EUNSUPPORTED, "ISAB: need D %% 4 == 0 and D %% H == 0");
    return isab_forward(X, B, N, d_in, D, H, M, ln, params, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t pca_pma_workspace_bytes(int B, int N, int D, int H, int S) { return pma_ws_bytes(B, N, D, H, S); }
int pca_pma_fwd_f32(const float* X, int B, int N, int D, int H, int S, int ln, const float* params, float* out,
                    void* workspace, size_t workspace_bytes, void* stream) {
    if (!X || !params || !out) return fail(PCA_EINVAL, "PMA: null pointer");
    if (B == 0) return 0;
    if (N <= 0) return fail(PCA_EINVAL, "PMA: empty set");
    if (D % 4 || D % H) return fail(PCA_EUNSUPPORTED, "PMA: need D %% 4 == 0 and D %% H == 0");
    return pma_forward(X, B, N, D, H, S, ln, params, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t pca_st_workspace_bytes(const pca_st_dims* dims, int B, int N, int precision) {
    if (!dims || B <= 0 || N <= 0) return 0;
    return st_any_ws_bytes(dims, B, N, precision);
}
int pca_st_fwd(const float* X, int B, int N, const pca_st_dims* dims, const float* params, float* logits,
               void* workspace, size_t workspace_bytes, int precision, void* stream) {
    return st_forward(X, B, N, dims, params, logits, workspace, workspace_bytes, precision, (cudaStream_t)stream);
}

int pca_st_fwd_masked(const float* X, const int32_t* counts, int B, int N, const pca_st_dims* dims, const float* params,
                      float* logits, void* workspace, size_t workspace_bytes, int precision, void* stream) {
    return st_forward(X, B, N, dims, params, logits, workspace, workspace_bytes, precision, (cudaStream_t)stream, counts);
}

size_t pca_deepset_workspace_bytes(int B, int N, int d_in, int dim_hidden, int out_dim) {
    (void)d_in; (void)out_dim;
    return deepset_ws_bytes(B, N, dim_hidden);
}
int pca_deepset_fwd_f32(const float* X, int B, int N, int d_in, int dim_hidden, int out_dim, int pool,
                        const float* params, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!X || !params || !out) return fail(PCA_EINVAL, "DeepSet: null pointer");
    if (B == 0) return 0;
    if (N <= 0) return fail(PCA_EINVAL, "DeepSet: empty set");
    return deepset_forward(X, B, N, d_in, dim_hidden, out_dim, pool, params, out, workspace, workspace_bytes,
                           (cudaStream_t)stream);
}

int pca_deepset_fwd_masked_f32(const float* X, const int32_t* counts, int B, int N, int d_in, int dim_hidden, int out_dim,
                               int pool, const float* params, float* out, void* workspace, size_t workspace_bytes,
                               void* stream) {
    if (!X || !params || !out) return fail(PCA_EINVAL, "DeepSet: null pointer");
    if (B == 0) return 0;
    if (N <= 0) return fail(PCA_EINVAL, "DeepSet: empty set");
    return deepset_forward(X, B, N, d_in, dim_hidden, out_dim, pool, params, out, workspace, workspace_bytes,
                           (cudaStream_t)stream, counts);
}

int pca_pipeline_clouds_per_clip(const pca_pipeline_cfg* cfg) {
    PipeShape s;
    return pipe_shape(cfg, &s) ? -1 : s.clouds_per_clip;
}
int pca_pipeline_points_per_cloud(const pca_pipeline_cfg* cfg) {
    PipeShape s;
    return pipe_shape(cfg, &s) ? -1 : s.pts;
}
size_t pca_pipeline_workspace_bytes(const pca_pipeline_cfg* cfg, int n_clips) {
    PipeShape s;
    if (pipe_shape(cfg, &s) || n_clips <= 0) return 0;
    return pipe_ws_bytes(cfg, s, n_clips, nullptr);
}
int pca_pipeline_run(const pca_pipeline_cfg* cfg, const float* audio, int n_clips, const float* window,
                     const float* twiddle, const float* farr, const float* tarr, const float* st_params,
                     float* logits, void* workspace, size_t workspace_bytes, void* stream) {
    if (!audio || !window || !twiddle || !farr || !st_params || !logits) return fail(PCA_EINVAL, "pipeline: null pointer");
    return pipeline_run(cfg, audio, n_clips, window, twiddle, farr, tarr, st_params, logits, workspace,
                        workspace_bytes, (cudaStream_t)stream);
}
int pca_pipeline_run_host(const pca_pipeline_cfg* cfg, const float* host_audio, int n_clips, float* dev_audio,
                          const float* window, const float* twiddle, const float* farr, const float* tarr,
                          const float* st_params, float* dev_logits, float* host_logits, void* workspace,
                          size_t workspace_bytes, void* stream) {
    if (!host_audio || !dev_audio || !dev_logits || !host_logits) return fail(PCA_EINVAL, "pipeline: null pointer");
    PipeShape s;
    PCA_TRY(pipe_shape(cfg, &s));
    cudaStream_t st = (cudaStream_t)stream;
    PCA_CHECK_CUDA(cudaMemcpyAsync(dev_audio, host_audio, (size_t)n_clips * cfg->n_samples * sizeof(float),
                                   cudaMemcpyHostToDevice, st));
    PCA_TRY(pca_pipeline_run(cfg, dev_audio, n_clips, window, twiddle, farr, tarr, st_params, dev_logits,
                             workspace, workspace_bytes, stream));
    const size_t out_bytes = (size_t)n_clips * s.clouds_per_clip * cfg->st.S * cfg->st.C * sizeof(float);
    PCA_CHECK_CUDA(cudaMemcpyAsync(host_logits, dev_logits, out_bytes, cudaMemcpyDeviceToHost, st));
    return 0;
}

int pca_pipeline_run_host_chunked(const pca_pipeline_cfg* cfg, const float* host_audio, int n_clips, float* dev_audio,
                                  const float* window, const float* twiddle, const float* farr, const float* tarr,
                                  const float* st_params, float* dev_logits, float* host_logits, void* workspace,
                                  size_t workspace_bytes, int n_chunks, void* copy_stream, void* stream) {
    if (!host_audio || !dev_audio || !dev_logits || !host_logits) return fail(PCA_EINVAL, "pipeline: null pointer");
    if (n_clips <= 0) return 0;
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > 16) n_chunks = 16;
    if (n_chunks > n_clips) n_chunks = n_clips;
    PipeShape s;
    PCA_TRY(pipe_shape(cfg, &s));
    cudaStream_t st = (cudaStream_t)stream, cs = (cudaStream_t)copy_stream;
    if (cs == st || n_chunks == 1)
        return pca_pipeline_run_host(cfg, host_audio, n_clips, dev_audio, window, twiddle, farr, tarr, st_params, dev_logits,
                                     host_logits, workspace, workspace_bytes, stream);
    cudaEvent_t ev[17];
    for (int k = 0; k <= n_chunks; ++k) PCA_CHECK_CUDA(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
    int rc = 0;
    // the staging buffer may still be read by earlier work on `stream`
    cudaError_t e = cudaEventRecord(ev[n_chunks], st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, ev[n_chunks], 0);
    // clip ranges: a short head range (1/8 of the batch) so the kernels start early, the rest split evenly
    size_t bound[17];
    bound[0] = 0;
    bound[1] = (size_t)n_clips / 8 > 0 ? (size_t)n_clips / 8 : 1;
    for (int k = 2; k <= n_chunks; ++k) bound[k] = bound[1] + ((size_t)n_clips - bound[1]) * (k - 1) / (n_chunks - 1);
    const size_t clip_floats = (size_t)cfg->n_samples;
    const size_t out_floats = (size_t)s.clouds_per_clip * cfg->st.S * cfg->st.C;
    for (int k = 0; k < n_chunks && e == cudaSuccess; ++k) {
        const size_t c0 = bound[k], c1 = bound[k + 1];
        if (c1 == c0) { e = cudaEventRecord(ev[k], cs); continue; }
        e = cudaMemcpyAsync(dev_audio + c0 * clip_floats, host_audio + c0 * clip_floats, (c1 - c0) * clip_floats * sizeof(float),
                            cudaMemcpyHostToDevice, cs);
        if (e == cudaSuccess) e = cudaEventRecord(ev[k], cs);
    }
    for (int k = 0; k < n_chunks && e == cudaSuccess && rc == 0; ++k) {
        const size_t c0 = bound[k], c1 = bound[k + 1];
        if (c1 == c0) continue;
        e = cudaStreamWaitEvent(st, ev[k], 0);
        if (e != cudaSuccess) break;
        rc = pca_pipeline_run(cfg, dev_audio + c0 * clip_floats, (int)(c1 - c0), window, twiddle, farr, tarr, st_params,
                              dev_logits + c0 * out_floats, workspace, workspace_bytes, stream);
    }
    if (e == cudaSuccess && rc == 0)
        e = cudaMemcpyAsync(host_logits, dev_logits, (size_t)n_clips * out_floats * sizeof(float), cudaMemcpyDeviceToHost, st);
    for (int k = 0; k <= n_chunks; ++k) cudaEventDestroy(ev[k]);     // release is deferred until the events complete
    if (rc != 0) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "pipeline_run_host_chunked");
    return 0;
}

int pca_debug_st_stages(const float* X, int B, int N, const pca_st_dims* dims, const float* params, float* logits,
                        float* H1, float* Y1, float* H2, float* Y2, float* pooled, void* workspace,
                        size_t workspace_bytes, void* stream) {
    PCA_TRY(check_dims(dims));
    if (!X || !params || !logits) return fail(PCA_EINVAL, "ST stages: null pointer");
    if (B <= 0 || N <= 0) return fail(PCA_EINVAL, "ST stages: bad shape");
    if (!st_tc_supported(dims, N)) return fail(PCA_EUNSUPPORTED, "ST stages: dims not supported by the tcgen05 path");
    return st_tc_forward_stages(X, B, N, dims, params, logits, H1, Y1, H2, Y2, pooled, workspace, workspace_bytes,
                                (cudaStream_t)stream);
}

int pca_linear_fwd_f32(const float* X, long long rows, int din, int dout, const float* params, float* Y, void* stream) {
    if (!X || !params || !Y) return fail(PCA_EINVAL, "linear: null pointer");
    if (rows < 0 || din <= 0 || dout <= 0) return fail(PCA_EINVAL, "linear: bad shape");
    return launch_linear(X, params, params + (long long)dout * din, Y, rows, din, dout, 0, (cudaStream_t)stream);
}

int pca_linear_bwd_f32(const float* dY, const float* X, long long rows, int din, int dout, const float* params, float* dX,
                       float* dparams, void* stream) {
    if (!dY || !X || !params || !dparams) return fail(PCA_EINVAL, "linear backward: null pointer");
    if (rows < 0 || din <= 0 || dout <= 0) return fail(PCA_EINVAL, "linear backward: bad shape");
    return linear_bwd_api(dY, X, rows, din, dout, params, dX, dparams, (cudaStream_t)stream);
}
int pca_dropout_f32(const float* in, float* out, long long n, float p, unsigned long long seed, void* stream) {
    if ((!in || !out) && n > 0) return fail(PCA_EINVAL, "dropout: null pointer");
    if (n < 0) return fail(PCA_EINVAL, "dropout: bad length");
    return dropout_api(in, out, n, p, seed, (cudaStream_t)stream);
}

int pca_random_keys_f32(float* keys, long long n, unsigned long long seed, void* stream) {
    if (!keys && n > 0) return fail(PCA_EINVAL, "random_keys: null pointer");
    return launch_random_keys(keys, n, seed, (cudaStream_t)stream);
}
int pca_gather_points_f32(const float* logmag, int n_clouds, int nf, int nt, const float* farr, const float* tarr,
                          const int32_t* idx, int K, float* pts, void* stream) {
    if (!logmag || !farr || !idx || !pts) return fail(PCA_EINVAL, "gather_points: null pointer");
    if (n_clouds < 0 || nf <= 0 || nt <= 0 || K < 0) return fail(PCA_EINVAL, "gather_points: bad shape");
    return launch_gather_points(logmag, n_clouds, nf, nt, farr, tarr, idx, K, pts, (cudaStream_t)stream);
}
int pca_importance_map_f32(const float* logmag, int n_clouds, int nf, int nt, const float* kf, int wf, const float* kt, int wt,
                           float* heat, float* scratch, void* stream) {
    if (!logmag || !kf || !kt || !heat || !scratch) return fail(PCA_EINVAL, "importance_map: null pointer");
    if (n_clouds < 0 || nf <= 0 || nt <= 0) return fail(PCA_EINVAL, "importance_map: bad shape");
    return launch_importance_map(logmag, n_clouds, nf, nt, kf, wf, kt, wt, heat, scratch, (cudaStream_t)stream);
}
int pca_multinomial_f32(const float* weights, int n_clouds, int n, int K, unsigned long long seed, double* cdf_scratch,
                        int32_t* idx, void* stream) {
    if (!weights || !cdf_scratch || !idx) return fail(PCA_EINVAL, "multinomial: null pointer");
    return launch_multinomial(weights, n_clouds, n, K, seed, cdf_scratch, idx, (cudaStream_t)stream);
}

int pca_resample_f32(const float* x, int n_clips, int n_in, int n_out, double sample_ratio, const double* win, const double* delta,
                     int nwin, int num_table, float out_scale, float* y, void* stream) {
    if (!x || !win || !delta || !y) return fail(PCA_EINVAL, "resample: null pointer");
    return launch_resample(x, n_clips, n_in, n_out, sample_ratio, win, delta, nwin, num_table, out_scale, y, (cudaStream_t)stream);
}

size_t pca_st_train_saved_bytes(const pca_st_dims* dims, int B, int N, float dropout_p) {
    if (!dims || B <= 0 || N <= 0) return 0;
    return st_train_saved_bytes(dims, B, N, dropout_p);
}
size_t pca_st_train_workspace_bytes(const pca_st_dims* dims, int B, int N) {
    if (!dims || B <= 0 || N <= 0) return 0;
    return st_train_ws_bytes(dims, B, N);
}
int pca_st_train_fwd_f32(const float* X, const int32_t* counts, int B, int N, const pca_st_dims* dims, const float* params,
                         float dropout_p, unsigned long long seed, float* logits, void* saved, size_t saved_bytes, void* workspace,
                         size_t workspace_bytes, void* stream) {
    return st_train_forward(X, counts, B, N, dims, params, dropout_p, seed, logits, saved, saved_bytes, workspace, workspace_bytes,
                            (cudaStream_t)stream);
}
int pca_st_train_bwd_f32(const float* X, const int32_t* counts, int B, int N, const pca_st_dims* dims, const float* params,
                         float dropout_p, unsigned long long seed, const float* dlogits, const void* saved, size_t saved_bytes,
                         float* dparams, float* dX, void* workspace, size_t workspace_bytes, void* stream) {
    return st_train_backward(X, counts, B, N, dims, params, dropout_p, seed, dlogits, saved, saved_bytes, dparams, dX, workspace,
                             workspace_bytes, (cudaStream_t)stream);
}
int pca_st_train_bwd_phase_f32(const float* X, const int32_t* counts, int B, int N, const pca_st_dims* dims, const float* params,
                               float dropout_p, unsigned long long seed, const float* dlogits, const void* saved, size_t saved_bytes,
                               float* dparams, float* dX, void* workspace, size_t workspace_bytes, int phase, long long* tail_offset,
                               void* stream) {
    return st_train_backward(X, counts, B, N, dims, params, dropout_p, seed, dlogits, saved, saved_bytes, dparams, dX, workspace,
                             workspace_bytes, (cudaStream_t)stream, phase, tail_offset);
}
size_t pca_mab_train_saved_bytes(int B, int q_batch, int nq, int nk, int D, int H, int ln) {
    return (B > 0 && nq > 0 && nk > 0 && D > 0 && H > 0) ? mab_train_saved_bytes(B, q_batch, nq, nk, D, H, ln) : 0;
}
size_t pca_mab_train_workspace_bytes(int B, int q_batch, int nq, int nk, int D, int H, int ln) {
    return (B > 0 && nq > 0 && nk > 0 && D > 0 && H > 0 && D % H == 0) ? mab_train_ws_bytes(B, q_batch, nq, nk, D, H, ln) : 0;
}
int pca_mab_train_fwd_f32(const float* Q, int q_batch, const float* K, int B, int nq, int nk, int dq, int dk, int D, int H, int ln,
                          const float* params, float* out, void* saved, size_t saved_bytes, void* workspace, size_t workspace_bytes,
                          void* stream) {
    return mab_train_forward_api(Q, q_batch, K, B, nq, nk, dq, dk, D, H, ln, params, out, saved, saved_bytes, workspace, workspace_bytes,
                                 (cudaStream_t)stream);
}
int pca_mab_train_bwd_f32(const float* Q, int q_batch, const float* K, int B, int nq, int nk, int dq, int dk, int D, int H, int ln,
                          const float* params, const float* dout, const void* saved, size_t saved_bytes, float* dparams, float* dQ,
                          float* dK, void* workspace, size_t workspace_bytes, void* stream) {
    return mab_train_backward_api(Q, q_batch, K, B, nq, nk, dq, dk, D, H, ln, params, dout, saved, saved_bytes, dparams, dQ, dK, workspace,
                                  workspace_bytes, (cudaStream_t)stream);
}
size_t pca_deepset_train_saved_bytes(int B, int N, int dim_hidden) {
    return (B > 0 && N > 0 && dim_hidden > 0) ? deepset_train_saved_bytes(B, N, dim_hidden) : 0;
}
size_t pca_deepset_train_workspace_bytes(int B, int N, int dim_hidden) {
    return (B > 0 && N > 0 && dim_hidden > 0) ? deepset_train_ws_bytes(B, N, dim_hidden) : 0;
}
int pca_deepset_train_fwd_f32(const float* X, int B, int N, int d_in, int dim_hidden, int out_dim, int pool, const float* params,
                              float* out, void* saved, size_t saved_bytes, void* workspace, size_t workspace_bytes, void* stream) {
    return deepset_train_forward(X, B, N, d_in, dim_hidden, out_dim, pool, params, out, saved, saved_bytes, workspace, workspace_bytes,
                                 (cudaStream_t)stream);
}
int pca_deepset_train_bwd_f32(const float* X, int B, int N, int d_in, int dim_hidden, int out_dim, int pool, const float* params,
                              const float* dout, const void* saved, size_t saved_bytes, float* dparams, float* dX, void* workspace,
                              size_t workspace_bytes, void* stream) {
    return deepset_train_backward(X, B, N, d_in, dim_hidden, out_dim, pool, params, dout, saved, saved_bytes, dparams, dX, workspace,
                                  workspace_bytes, (cudaStream_t)stream);
}
int pca_cross_entropy_f32(const float* logits, const int64_t* labels, int B, int C, float* loss, int32_t* correct, float* dlogits,
                          void* stream) {
    if (!logits || !labels || !loss) return fail(PCA_EINVAL, "cross entropy: null pointer");
    return launch_cross_entropy(logits, (const long long*)labels, B, C, 1.0f / (float)(B > 0 ? B : 1), loss, correct, dlogits,
                                (cudaStream_t)stream);
}
int pca_adam_step_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq) return fail(PCA_EINVAL, "adam: null pointer");
    return launch_adam(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, (cudaStream_t)stream);
}

void pca_debug_set_gemm_tc(int on) { set_gemm_tc(on); }
void pca_debug_set_attn_tc(int on) { set_attn_tc(on); }
size_t pca_debug_attn_ws_bytes(int B, int nq, int nk, int D, int H) {
    const size_t f = attn_part_floats(B, nq, nk, D, H), b = attn_tc_bwd_floats(B, nq, nk, D, H);
    return ((f > b ? f : b) + 64) * sizeof(float);
}
int pca_debug_attn_tc_eligible(int B, int nq, int nk, int D, int H) { return attn_tc_eligible(B, nq, nk, D, H) ? 1 : 0; }
int pca_debug_attn_fwd(const float* Qp, int q_shared, const float* KV, int B, int nq, int nk, int D, int H, float* O, float* lse,
                       void* ws, size_t ws_bytes, void* stream) {
    if (!Qp || !KV || !O || !ws) return fail(PCA_EINVAL, "attn_fwd: null pointer");
    if (ws_bytes < pca_debug_attn_ws_bytes(B, nq, nk, D, H)) return fail(PCA_EWORKSPACE, "attn_fwd: workspace too small");
    return launch_attn(Qp, q_shared ? 0 : (long long)nq * D, KV, B, nq, nk, D, H, O, (float*)ws, nullptr, (cudaStream_t)stream, lse);
}
int pca_debug_attn_bwd_tc(const float* Qp, int q_shared, const float* KV, const float* dO, const float* lse, const float* delta, int B,
                          int nq, int nk, int D, int H, float* dQp, float* dKV, void* ws, size_t ws_bytes, void* stream) {
    if (!Qp || !KV || !dO || !dQp || !dKV || !ws) return fail(PCA_EINVAL, "attn_bwd_tc: null pointer");
    if (ws_bytes < pca_debug_attn_ws_bytes(B, nq, nk, D, H)) return fail(PCA_EWORKSPACE, "attn_bwd_tc: workspace too small");
    return launch_attn_bwd_tc(Qp, q_shared ? 0 : (long long)nq * D, KV, dO, lse, delta, B, nq, nk, D, H, dQp, dKV, (float*)ws,
                              (cudaStream_t)stream);
}
int pca_debug_linear_tc(const float* X, const float* W, int trans_w, const float* bias, const float* resid, float* Y, float* R,
                        long long rows, int K, int N, int relu, void* image, size_t image_bytes, void* stream) {
    return launch_linear_tc(X, W, trans_w, bias, resid, Y, R, rows, K, N, relu, image, image_bytes, (cudaStream_t)stream);
}
int pca_debug_grad_weight_tc(const float* dY, const float* X, float* dW, long long rows, int M, int N, void* stream) {
    return launch_grad_weight_tc(dY, X, dW, rows, M, N, (cudaStream_t)stream);
}

void pca_debug_set_timeline(long long* device_buffer) { set_timeline(device_buffer); }
void pca_debug_set_tail_max(int tail_max) { set_tail_max(tail_max); }
void pca_debug_set_reduce_variant(int warpgroups) { set_reduce_wg(warpgroups); }
void pca_debug_set_pool_variant(int variant) { set_pool_variant(variant); }
void pca_debug_set_stft_generic(int on) { debug_set_stft_generic(on); }

int pca_debug_umma_probe(const float* A, const float* B, float* D, int N, int K, int a_mode, int b_mode, void* stream) {
    return launch_umma_probe(A, B, D, N, K, a_mode, b_mode, (cudaStream_t)stream);
}

}  // extern "C"
