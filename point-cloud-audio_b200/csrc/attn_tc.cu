// fp32-grade multi-head attention on the tcgen05 tensor cores for MABs with ONE SMALL SIDE (set_transformer-master/modules.py:20-29
// with 16 inducing points or 1 seed against ~1000 points: ISAB mab0 / mab1 and PMA of main_pointcloud.py:62, forward and backward).
//
// With ns <= 16 items on the small side and H heads, the H * ns (head, item) pairs are the COLUMNS of one matrix T (points, HS)
// and every contraction of the attention and of its gradient is one of three GEMM shapes over the points of a cloud:
//   G1  T  = X  Wz^T      X (points, D) activations, Wz (HS, D) a per-cloud BLOCK-STRUCTURED operand: row (h, m) carries the small
//                          side's row m inside head h's feature slice and zeros elsewhere (scores Q K^T, dP = dO V^T)
//   G2  Y  = T  Wz        (points, HS) x (HS, D): P V, dS K, dS^T-free forms of dK / dV when the points are the keys
//   G3  Gz = T^T X        summed over the cloud's points, diagonal (head) blocks extracted: P^T V, dS^T Q, P^T dO
// All three run as 3-term split-bf16 products (a_hi b_hi + a_lo b_hi + a_hi b_lo, fp32 accumulation in TMEM; gemm_tc.cu), so the
// results stay in the 1e-3 fp32 parity class.  The softmax and its backward are fused into the epilogues of G1:
//   points = queries (mab1):  row softmax per head over the ns keys (thread = row holds the whole row)
//   points = keys (mab0/PMA): column softmax over the points of a cloud (col_softmax_kernel between G1 and G3)
// The zero blocks cost 4x redundant MMA work at H = 4 -- irrelevant: these GEMMs are HBM-bound (the tensor pipe is < 25 % busy).
// Kernels: cloud_linear_tc_kernel (G1 / G2: TMA raw-tile ring, converter warps, epilogue variants, optional second K source),
// cloud_gw_tc_kernel (G3: persistent, two accumulators), col_softmax_kernel, cloud_image_kernel (operand images), and the small
// kernels of the FOLDED form of blocks with a shared query set (inducing points / seeds), which runs forward and backward on the
// un-projected points: W_k is folded into the query image and W_v applied to the per-cloud sums P^T X (launch_attn_folded,
// launch_attn_folded_bwd) -- the K | V projection of those blocks and its gradient GEMMs never run.
#include "common.cuh"
#include "tc_prims.cuh"
#include "tma_host.cuh"

namespace pca {
using namespace tc;

constexpr int AT_KC = 32;
constexpr int AT_PSETS = 2;           // producer warp sets: set s stages the K chunks s, s + 2, ... (twice the loads in flight per SM)
constexpr int AT_MMA_WARP = 4 * AT_PSETS;
constexpr int AT_LOAD_WARP = AT_MMA_WARP + 5;
constexpr int AT_THREADS = (4 * AT_PSETS + 6) * 32;   // 8 producer / converter + 1 MMA + 4 epilogue + 1 TMA loader warps
constexpr int CL_STAGES = 3;          // operand stages of the per-cloud linear kernel
constexpr int CL_RAW_MAX = 8;         // raw fp32 tiles (128 rows x 32 floats = 16 KB) the TMA loader may have in flight
constexpr int CL_RAW_BYTES = 128 * AT_KC * 4;

__device__ __forceinline__ void at_warp_arrive(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void at_split8(const float* x, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 hb = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
        const float2 hf = __bfloat1622float2(hb);
        h[j] = *reinterpret_cast<const uint32_t*>(&hb);
        l[j] = pack_bf16(x[2 * j] - hf.x, x[2 * j + 1] - hf.y);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ------------------------------------------------------------------------------------ per-cloud operand images
// kind 0 (G1): B(n = (h, m), k)     = src[m * ld + k] when k / dh == h and m < ns, else 0     (N = HS, K = D)
// kind 1 (G2): B(n = d, k = (h, m)) = src[m * ld + d] when d / dh == h and m < ns, else 0     (N = D,  K = HS)
// Image of cloud b at img + b * img_bstride: per 32-wide K chunk c [hi | lo], each K-major canonical
// (byte = (kk / 8) * N * 16 + n * 16 + (kk % 8) * 2), as weight_image_kernel of gemm_tc.cu with a single column pass.
__global__ void cloud_image_kernel(const float* __restrict__ src, long long s_bstride, int ld, int ns, int nsp, int dh, int D, int HS,
                                   int kind, uint8_t* __restrict__ img, long long img_bstride) {
    const int b = blockIdx.y;
    const int N = (kind == 1 || kind == 3) ? D : HS, K = (kind == 1 || kind == 3) ? HS : D;
    const int total = N * (K / 8);
    const float* s = src + (long long)b * s_bstride;
    uint8_t* ib = img + (long long)b * img_bstride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i % N, k8 = i / N;
        float x[8];
        if (kind == 2) {                         // dense (HS, K) matrix as it stands (the folded query image)
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = __ldg(s + (long long)n * ld + k8 * 8 + j);
        } else if (kind == 3) {                  // the same matrix as a G2 operand: B(n, k) = src[k][n], k < HS
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = __ldg(s + (long long)(k8 * 8 + j) * ld + n);
        } else if (kind == 0) {
            const int h = n / nsp, m = n - h * nsp;
            const bool on = m < ns && (k8 * 8) / dh == h;
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = on ? __ldg(s + (long long)m * ld + k8 * 8 + j) : 0.f;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k8 * 8 + j;
                const int h = k / nsp, m = k - h * nsp;
                x[j] = (m < ns && n / dh == h) ? __ldg(s + (long long)m * ld + n) : 0.f;
            }
        }
        uint4 hi, lo;
        at_split8(x, hi, lo);
        const int c = (k8 * 8) / AT_KC, kk8 = k8 - c * (AT_KC / 8);
        uint8_t* base = ib + ((size_t)c * 2) * N * 64;
        *reinterpret_cast<uint4*>(base + (size_t)kk8 * N * 16 + n * 16) = hi;
        *reinterpret_cast<uint4*>(base + (size_t)N * 64 + (size_t)kk8 * N * 16 + n * 16) = lo;
    }
}

// ------------------------------------------------------------------------------------ G1 / G2: per-cloud linear maps
enum { EPI_STORE = 0, EPI_RESID = 1, EPI_SOFTMAX = 2, EPI_DS_ROW = 3, EPI_P_COL = 4, EPI_DS_COL = 5 };

struct ClinParams {
    const float* X;        // row r of cloud b: X + b * x_bstride + r * ldx  (K columns used)
    long long x_bstride;
    int ldx;
    const uint8_t* img;    // operand image of cloud b at img + b * img_bstride (0: shared by the batch)
    long long img_bstride;
    float* Y;              // (n_rows, N) per cloud, row stride ldy
    long long y_bstride;
    int ldy;
    const float* R;        // EPI_RESID: residual rows; EPI_DS_ROW / EPI_DS_COL: probability rows
    long long r_bstride;
    int ldr;
    float* lse;            // EPI_SOFTMAX: (B, n_rows, H) out, nullable
    const float* vec;      // EPI_P_COL: log-sum-exp, EPI_DS_COL: delta; (B, ns, H)
    int B, n_rows, K, N;
    int H, nsp, ns;
    float scale, scale_log2e;
    float* colsum;         // nullable: colsum[n] += sum over the valid rows of the stored Y[., n] (the bias gradient of the layer
                           // that produced the operand: saves a separate pass over Y)
    int raw_slots;         // raw fp32 tiles in the TMA ring (<= CL_RAW_MAX)
    int x_shared;          // X has no batch dimension (x_bstride == 0)
    // optional second source: the K chunks from k_split on come from X2 (its own tensor map) against the image img2 -- one
    // launch then computes Y = X W + X2 W2 (the folded backward's dX = dS Gq + P dZ')
    const float* X2;
    long long x2_bstride;
    int ldx2;
    const uint8_t* img2;
    long long img2_bstride;
    int k_split;           // K chunks of the first source (= K / 32 when there is no second one)
};

// The activation images are written by the producers with a PADDED chunk stride (the descriptor's leading byte offset is free):
// K-major: 16-byte chunk c of row r at c * A_LBO + r * 16 with A_LBO = 2048 + 16; MN-major (G3): 8-feature group g of row k at
// g * G_SBO + k * 16 with G_SBO = 512 + 16 -- so that the four chunks / groups a quarter warp stores to fall into distinct banks.
struct AtSmem {
    static constexpr int A_LBO = 128 * 16 + 16;
    static constexpr int A_BYTES = 8320;                               // >= 4 * A_LBO, 128-byte multiple
    static constexpr int STAGE = 2 * A_BYTES + 2 * 256 * AT_KC * 2;   // A hi | A lo | B hi | B lo (B sized for N = 256)
    static constexpr int G_SBO = AT_KC * 16 + 16;
    static constexpr int GA_BYTES = 16 * G_SBO;                        // 128 features
    static constexpr int GB_BYTES = 32 * G_SBO;                        // 256 features
    static constexpr int G_STAGE = 2 * GA_BYTES + 2 * GB_BYTES;
    static constexpr int CL_FIXED = 4 * 32 * 33 * 4 + 256 + 1024;       // transpose tiles + barriers + per-CTA column sums
    static constexpr int MAX_BYTES = 227 * 1024;
    // G3 kernel: raw fp32 ring (32-row chunks of both operands) | 2 operand stages | transpose tiles | barriers
    static constexpr int G_OP_STAGES = 2;
    static constexpr int G_RAW_MAX = 4;
    __host__ __device__ static int g_raw_bytes(int M, int N) { return AT_KC * (M + N) * 4; }
    __host__ __device__ static int g_b_bytes(int N) { return (N / 8) * G_SBO; }                    // one hi or lo image of X's slice
    __host__ __device__ static int g_stage(int N) { return 2 * GA_BYTES + 2 * g_b_bytes(N); }
    __host__ static int g_raw_slots(int M, int N) {
        int r = (MAX_BYTES - CL_FIXED - G_OP_STAGES * g_stage(N)) / g_raw_bytes(M, N);
        r = r > G_RAW_MAX ? G_RAW_MAX : r;
        return r & ~1;                                                  // even: see cl_raw_slots
    }
    __host__ __device__ static int g_total(int M, int N, int raw_slots) { return raw_slots * g_raw_bytes(M, N) + G_OP_STAGES * g_stage(N) + CL_FIXED; }
    // per-cloud linear kernel: raw fp32 ring | operand stages (B region sized for the launch's N) | transpose tiles | barriers
    __host__ __device__ static int cl_stage(int N) { return 2 * A_BYTES + 2 * N * 64; }
    __host__ static int cl_raw_slots(int N) {
        // an EVEN count: a slot is then always consumed by the same converter set, in order -- with an odd count the two sets
        // alternate on a slot and a set could probe a slot's barrier two phases ahead (TMA boxes may land out of order), which
        // the parity test cannot tell from "complete"
        int r = (MAX_BYTES - CL_FIXED - CL_STAGES * cl_stage(N)) / CL_RAW_BYTES;
        r = r > CL_RAW_MAX ? CL_RAW_MAX : r;
        return r & ~1;
    }
    __host__ __device__ static int cl_total(int N, int raw_slots) { return raw_slots * CL_RAW_BYTES + CL_STAGES * cl_stage(N) + CL_FIXED; }
};

// lanes 2j / 2j + 1 hold adjacent float4 pieces (a: of row A, b: of row B = A + 4); after the exchange the even lane owns the 8
// consecutive floats of row A and the odd lane those of row B (one 16-byte bf16 chunk each)
__device__ __forceinline__ void at_pair_exchange(const float4 a, const float4 b, bool odd, float* x) {
    const float4 send = odd ? a : b;
    float4 recv;
    recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
    recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
    recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1);
    recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
    const float4 lo4 = odd ? recv : a, hi4 = odd ? b : recv;
    x[0] = lo4.x; x[1] = lo4.y; x[2] = lo4.z; x[3] = lo4.w;
    x[4] = hi4.x; x[5] = hi4.y; x[6] = hi4.z; x[7] = hi4.w;
}

// Persistent, warp-specialised.  ONE loader thread streams the fp32 rows of the 128-row tiles with TMA (a 3-D tensor map
// (K, rows of a cloud, clouds): rows past the end of a cloud arrive as zeros) into a ring of raw 16 KB tiles -- up to eight in
// flight per SM, independent of any warp's registers or scoreboard (with register-staged loads the kernels sat at 2.5 TB/s:
// one chunk in flight per warp set, and the proxy fence after the operand stores waits for outstanding loads).  8 converter
// warps (two sets on alternate K chunks) read the raw rows (quarter warp = the 128 bytes of one row), split them into hi / lo
// bf16 and store the canonical K-major operand images; the cloud's B image arrives by one bulk copy per K chunk; one warp
// issues the MMAs; 4 epilogue warps read the accumulator (thread = row), apply the epilogue and write 128-byte coalesced rows
// through a padded transpose tile.  Tiles never straddle clouds (tile t = (cloud t / tpc, rows 128 (t % tpc) ...)).
template <int EPI, int NSP>
__global__ void __launch_bounds__(AT_THREADS, 1) cloud_linear_tc_kernel(const ClinParams P, const __grid_constant__ CUtensorMap tmx,
                                                                        const __grid_constant__ CUtensorMap tmx2) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int stage_bytes = AtSmem::cl_stage(P.N);
    uint8_t* ops = smem + P.raw_slots * CL_RAW_BYTES;
    uint8_t* trans = ops + CL_STAGES * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(trans + 4 * 32 * 33 * 4);
    uint64_t* full = bars;                       // [3] count 5: the 4 converter warps of a set + the expect_tx arrival of the B image
    uint64_t* empty = bars + CL_STAGES;          // [3] count 1 (MMA commit)
    uint64_t* acc_full = bars + 2 * CL_STAGES;   // [2] count 1
    uint64_t* acc_empty = acc_full + 2;          // [2] count 4 (epilogue warps)
    uint64_t* raw_full = acc_empty + 2;          // [8] count 1 (expect_tx of the loader) + 16 KB
    uint64_t* raw_empty = raw_full + CL_RAW_MAX; // [8] count 4 (converter warps of the set that read the slot)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + CL_RAW_MAX);
    float* cs_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);     // [256] column sums of this CTA's tiles
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkc = P.K / AT_KC;
    const int tpc = (P.n_rows + 127) / 128;
    if (threadIdx.x < 256) cs_s[threadIdx.x] = 0.f;
    const long long ntiles = (long long)P.B * tpc;
    const int my_tiles = blockIdx.x < ntiles ? (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    const uint32_t items = (uint32_t)my_tiles * nkc;       // (tile of this CTA, K chunk) in issue order

    if (warp == AT_MMA_WARP) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 0) {
        for (int i = 0; i < CL_STAGES; ++i) { mbar_init(&full[i], 5); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        for (int i = 0; i < CL_RAW_MAX; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 4); }
        fence_barrier_init();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;
    const uint32_t b_img_bytes = (uint32_t)P.N * 64;           // one hi or lo operand image of a chunk

    if (warp == AT_LOAD_WARP) {
        // ================================================================= TMA loader (one thread)
        if (lane == 0) {
            tma_prefetch_desc(&tmx);
            tma_prefetch_desc(&tmx2);
            for (uint32_t g = 0; g < items; ++g) {
                const int slot = g % P.raw_slots;
                const uint32_t lt = g / nkc;
                const int kc = (int)(g - lt * nkc);
                const int tile = (int)(blockIdx.x + lt * gridDim.x);
                const int b = tile / tpc, j = tile - b * tpc;
                if (g >= (uint32_t)P.raw_slots) mbar_wait(&raw_empty[slot], ((g / P.raw_slots) - 1) & 1);
                mbar_arrive_expect_tx(&raw_full[slot], CL_RAW_BYTES);
                if (kc < P.k_split) tma_load_3d(smem + slot * CL_RAW_BYTES, &tmx, kc * AT_KC, j * 128, P.x_shared ? 0 : b, &raw_full[slot]);
                else tma_load_3d(smem + slot * CL_RAW_BYTES, &tmx2, (kc - P.k_split) * AT_KC, j * 128, b, &raw_full[slot]);
            }
        }
    } else if (warp < AT_MMA_WARP) {
        // ================================================================= converters
        // Warp set g % AT_PSETS converts item g; warp pw of the set the rows [32 pw, 32 pw + 32) of the tile.  Quarter warp q
        // reads the 128 bytes (the whole K chunk) of row 8 it + q and of row 8 it + 4 + q from the raw tile; a pair exchange
        // gives every lane the 8 consecutive floats of one 16-byte bf16 chunk.
        const int set = warp >> 2, pw = warp & 3;
        const int q = lane >> 3, piece = lane & 7;
        const bool odd = piece & 1;
        for (uint32_t g = set; g < items; g += AT_PSETS) {
            const int slot = g % P.raw_slots;
            const int stage = g % CL_STAGES;
            mbar_wait(&raw_full[slot], (g / P.raw_slots) & 1);
            const float4* raw = reinterpret_cast<const float4*>(smem + slot * CL_RAW_BYTES) + piece;
            float4 xv[8];
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int rA = 32 * pw + 8 * it + q;
                xv[2 * it] = raw[rA * 8];
                xv[2 * it + 1] = raw[(rA + 4) * 8];
            }
            if (g >= CL_STAGES) mbar_wait(&empty[stage], ((g / CL_STAGES) - 1) & 1);
            uint8_t* st = ops + stage * stage_bytes;
            if (pw == 0 && lane == 0) {
                const uint32_t lt = g / nkc;
                const int kc = (int)(g - lt * nkc);
                const long long b = (int)(blockIdx.x + lt * gridDim.x) / tpc;
                mbar_arrive_expect_tx(&full[stage], 2 * b_img_bytes);
                const uint8_t* isrc = kc < P.k_split ? P.img + b * P.img_bstride + ((size_t)kc * 2) * b_img_bytes
                                                     : P.img2 + b * P.img2_bstride + ((size_t)(kc - P.k_split) * 2) * b_img_bytes;
                bulk_copy_g2s(st + 2 * AtSmem::A_BYTES, isrc, 2 * b_img_bytes, &full[stage]);
            }
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                float x[8];
                at_pair_exchange(xv[2 * it], xv[2 * it + 1], odd, x);
                uint4 hi, lo;
                at_split8(x, hi, lo);
                const int row = 32 * pw + 8 * it + q + (odd ? 4 : 0);
                const int off = (piece >> 1) * AtSmem::A_LBO + row * 16;
                *reinterpret_cast<uint4*>(st + off) = hi;
                *reinterpret_cast<uint4*>(st + AtSmem::A_BYTES + off) = lo;
            }
            at_warp_arrive(&raw_empty[slot]);              // the raw tile is in registers / converted: the loader may refill the slot
            fence_async_smem();
            at_warp_arrive(&full[stage]);
        }
    } else if (warp == AT_MMA_WARP) {
        // ================================================================= MMA issue (warp-uniform, elected lane)
        const uint32_t idesc = idesc_bf16(128, P.N, 0, 0);
        uint32_t gt = 0, tt = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++tt) {
            const uint32_t buf = tt & 1;
            if (tt >= 2) mbar_wait(&acc_empty[buf], ((tt >> 1) - 1) & 1);
            fence_after_sync();
            const uint32_t acc = tmem_addr(tb, 0, 256 * buf);
            for (int kc = 0; kc < nkc; ++kc, ++gt) {
                const int stage = gt % CL_STAGES;
                mbar_wait(&full[stage], (gt / CL_STAGES) & 1);
                fence_after_sync();
                if (elect_one()) {
                    const uint32_t a_hi = smem_u32(ops + stage * stage_bytes);
                    const uint32_t a_lo = a_hi + AtSmem::A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * AtSmem::A_BYTES;
                    const uint32_t b_lo = b_hi + b_img_bytes;
                    const uint32_t b_lbo = (uint32_t)P.N * 16;
#pragma unroll
                    for (int ks = 0; ks < AT_KC / 16; ++ks) {
                        const uint64_t dah = smem_desc(a_hi + ks * 2 * AtSmem::A_LBO, AtSmem::A_LBO, 128);
                        const uint64_t dal = smem_desc(a_lo + ks * 2 * AtSmem::A_LBO, AtSmem::A_LBO, 128);
                        const uint64_t dbh = smem_desc(b_hi + ks * 2 * b_lbo, b_lbo, 128), dbl = smem_desc(b_lo + ks * 2 * b_lbo, b_lbo, 128);
                        mma_ss(acc, dah, dbh, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
                        mma_ss(acc, dal, dbh, idesc, 1u);
                        mma_ss(acc, dah, dbl, idesc, 1u);
                    }
                    mma_commit(&empty[stage]);
                    if (kc == nkc - 1) mma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp < AT_LOAD_WARP) {
        // ================================================================= epilogue
        const int quad = warp & 3;
        float* T = reinterpret_cast<float*>(trans) + quad * (32 * 33);
        const int rr = lane >> 3, cq = lane & 7;
        uint32_t tt = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++tt) {
            const uint32_t buf = tt & 1;
            const int b = (int)(t / tpc);
            const int i0 = (int)(t - (long long)b * tpc) * 128 + 32 * quad;     // first row (inside the cloud) of this warp
            float* yb = P.Y + (long long)b * P.y_bstride;
            const float* rb = (EPI == EPI_RESID || EPI == EPI_DS_ROW || EPI == EPI_DS_COL) ? P.R + (long long)b * P.r_bstride : nullptr;
            // the R rows (residual / probabilities) of a 32-column block are fetched one block ahead: 4 rows x 128 contiguous
            // bytes per instruction, in flight while the previous block is processed (and, for the first, under the MMAs)
            constexpr bool kUsesR = EPI == EPI_RESID || EPI == EPI_DS_ROW || EPI == EPI_DS_COL;
            float4 rvn[8];
            auto load_r = [&](int c0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = i0 + 4 * i + rr;
                    rvn[i] = r < P.n_rows ? *reinterpret_cast<const float4*>(rb + (long long)r * P.ldr + c0 + 4 * cq) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            if (kUsesR) load_r(0);
            mbar_wait(&acc_full[buf], (tt >> 1) & 1);
            fence_after_sync();
            for (int c0 = 0; c0 < P.N; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_addr(tb, 32 * quad, 256 * buf + c0), v);
                float4 rv8[8];
                if (kUsesR) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) rv8[i] = rvn[i];
                    if (c0 + 32 < P.N) load_r(c0 + 32);
                }
                tmem_ld_wait32(v);
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (EPI == EPI_DS_ROW) {
                    // the row's probabilities: coalesced block load, transposed through the tile
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int lr = 4 * i + rr;
                        const float4 pv = rv8[i];
                        T[lr * 33 + 4 * cq] = pv.x; T[lr * 33 + 4 * cq + 1] = pv.y; T[lr * 33 + 4 * cq + 2] = pv.z; T[lr * 33 + 4 * cq + 3] = pv.w;
                    }
                    __syncwarp();
                    float p[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) p[j] = T[lane * 33 + j];
                    __syncwarp();
#pragma unroll
                    for (int h = 0; h < 32 / NSP; ++h) {
                        float delta = 0.f;                      // sum_m P dP of this (row, head)
#pragma unroll
                        for (int m = 0; m < NSP; ++m) delta = fmaf(p[h * NSP + m], f[h * NSP + m], delta);
#pragma unroll
                        for (int m = 0; m < NSP; ++m) f[h * NSP + m] = p[h * NSP + m] * (f[h * NSP + m] - delta) * P.scale;
                    }
                }
                if (EPI == EPI_SOFTMAX) {
#pragma unroll
                    for (int h = 0; h < 32 / NSP; ++h) {
                        float mx = -INFINITY;
#pragma unroll
                        for (int m = 0; m < NSP; ++m) {
                            f[h * NSP + m] *= P.scale_log2e;
                            if (m < P.ns) mx = fmaxf(mx, f[h * NSP + m]);
                        }
                        float l = 0.f;
#pragma unroll
                        for (int m = 0; m < NSP; ++m) {
                            f[h * NSP + m] = m < P.ns ? ex2(f[h * NSP + m] - mx) : 0.f;
                            l += f[h * NSP + m];
                        }
                        const float inv = 1.f / l;
#pragma unroll
                        for (int m = 0; m < NSP; ++m) f[h * NSP + m] *= inv;
                        const int r = i0 + lane;
                        if (P.lse && r < P.n_rows) P.lse[((long long)b * P.n_rows + r) * P.H + c0 / NSP + h] = mx + log2f(l);
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) T[lane * 33 + j] = f[j];
                __syncwarp();
                const int n = c0 + 4 * cq;
                float cv[4] = {0.f, 0.f, 0.f, 0.f};
                bool cok[4] = {true, true, true, true};
                if (EPI == EPI_P_COL || EPI == EPI_DS_COL) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int h = (n + u) / P.nsp, m = (n + u) - h * P.nsp;
                        cok[u] = m < P.ns;
                        cv[u] = cok[u] ? __ldg(P.vec + ((long long)b * P.ns + m) * P.H + h) : 0.f;
                    }
                }
                float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int lr = 4 * i + rr, r = i0 + lr;
                    float o[4] = {T[lr * 33 + 4 * cq], T[lr * 33 + 4 * cq + 1], T[lr * 33 + 4 * cq + 2], T[lr * 33 + 4 * cq + 3]};
                    if (EPI == EPI_RESID) { o[0] += rv8[i].x; o[1] += rv8[i].y; o[2] += rv8[i].z; o[3] += rv8[i].w; }
                    if (EPI == EPI_P_COL) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) o[u] = cok[u] ? ex2(fmaf(o[u], P.scale_log2e, -cv[u])) : 0.f;
                    }
                    if (EPI == EPI_DS_COL) {
                        const float pr[4] = {rv8[i].x, rv8[i].y, rv8[i].z, rv8[i].w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) o[u] = cok[u] ? pr[u] * (o[u] - cv[u]) * P.scale : 0.f;
                    }
                    if (r < P.n_rows) {
                        *reinterpret_cast<float4*>(yb + (long long)r * P.ldy + n) = make_float4(o[0], o[1], o[2], o[3]);
                        cs[0] += o[0]; cs[1] += o[1]; cs[2] += o[2]; cs[3] += o[3];
                    }
                }
                if (P.colsum) {                              // warp-uniform: 32 rows x 4 columns per lane group -> one 16-byte reduction
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        cs[u] += __shfl_xor_sync(0xffffffffu, cs[u], 8);
                        cs[u] += __shfl_xor_sync(0xffffffffu, cs[u], 16);
                    }
                    if (rr == 0) {                           // per-CTA sums in shared memory (every tile adding to the same 256 global
                                                             // addresses serialised in L2: +45 % on the kernel); flushed once at the end
#pragma unroll
                        for (int u = 0; u < 4; ++u) atomicAdd(&cs_s[n + u], cs[u]);
                    }
                }
                __syncwarp();
            }
            fence_before_sync();
            at_warp_arrive(&acc_empty[buf]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (P.colsum && threadIdx.x < P.N) atomicAdd(P.colsum + threadIdx.x, cs_s[threadIdx.x]);
    if (warp == AT_MMA_WARP) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------ G3: per-cloud T^T X, head blocks extracted
// out[b][m][d] += sum over the rows r of the CTA's range of T[b][r][(d / dh, m)] * X[b][r][d]   (m < ns).  Both operands are
// MN-major straight from the row-major activations (grad_weight_tc_kernel's staging); M = HS <= 128 rows, N = D <= 256 columns.
struct ClGwParams {
    const float* A;        // T: (n_rows, HS) per cloud
    long long a_bstride;
    int lda, Mtot;
    const float* Bm;       // X: (n_rows, N) per cloud, row stride ldb
    long long b_bstride;
    int ldb, N;
    float* out;            // (ns, N) per cloud, row stride ldo; added atomically
    long long o_bstride;
    int ldo;
    int n_rows, rchunk, nsp, ns, dh;
    int raw_slots, a_shared, b_shared;
    int ncta;              // columns of X per CTA (= N)
    int B, nsplit;         // work items = B clouds x nsplit row ranges
    const int* counts;     // nullable: rows of X past counts[b] are padding (possibly NaN) and are read as zeros
    int full_rows;         // 1: out (Mtot, N) per cloud receives every accumulator row (no head-diagonal extraction)
};

__global__ void __launch_bounds__(AT_THREADS, 1) cloud_gw_tc_kernel(const ClGwParams P, const __grid_constant__ CUtensorMap tma,
                                                                    const __grid_constant__ CUtensorMap tmb) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int raw_bytes = AtSmem::g_raw_bytes(P.Mtot, P.ncta);
    const int col0 = 0;
    uint8_t* ops = smem + P.raw_slots * raw_bytes;
    const int op_stage = AtSmem::g_stage(P.ncta), gb_bytes = AtSmem::g_b_bytes(P.ncta);
    uint8_t* trans = ops + AtSmem::G_OP_STAGES * op_stage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(trans + 4 * 32 * 33 * 4);
    uint64_t* full = bars;                                   // [2] count 4 (the converter warps of a set)
    uint64_t* empty = bars + AtSmem::G_OP_STAGES;            // [2] count 1
    uint64_t* acc_full = bars + 2 * AtSmem::G_OP_STAGES;     // [2] count 1
    uint64_t* acc_empty = acc_full + 2;                      // [2] count 4 (epilogue warps)
    uint64_t* raw_full = acc_empty + 2;                      // [4] count 1 (expect_tx of the loader) + the chunk's bytes
    uint64_t* raw_empty = raw_full + AtSmem::G_RAW_MAX;      // [4] count 4
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + AtSmem::G_RAW_MAX);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Persistent: work item = (cloud, row range of rchunk rows -- a multiple of the chunk size, so a chunk never straddles two
    // items); this CTA walks the items blockIdx.x, + gridDim.x, ...; the raw ring / operand stages run across item boundaries
    // (global chunk counter) and two accumulators let the epilogue of an item overlap the chunks of the next one.
    const int n_items = P.B * P.nsplit;
    auto item_rows = [&](int it, int& b, int& r0) -> int {   // chunks of the item
        b = it / P.nsplit;
        r0 = (it - b * P.nsplit) * P.rchunk;
        const int r1 = min(r0 + P.rchunk, P.n_rows);
        return (r1 - r0 + AT_KC - 1) / AT_KC;
    };

    if (warp == AT_MMA_WARP) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 0) {
        for (int i = 0; i < AtSmem::G_OP_STAGES; ++i) { mbar_init(&full[i], 4); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        for (int i = 0; i < AtSmem::G_RAW_MAX; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 4); }
        fence_barrier_init();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;

    if (warp == AT_LOAD_WARP) {
        // TMA loader (one thread): chunk c = rows [r0 + 32 c, + 32) of T (Mtot columns) and of X (N columns); rows past the end
        // of the cloud arrive as zeros
        if (lane == 0) {
            tma_prefetch_desc(&tma);
            tma_prefetch_desc(&tmb);
            int gc = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                int b, r0;
                const int nchunks = item_rows(it, b, r0);
                for (int c = 0; c < nchunks; ++c, ++gc) {
                    const int slot = gc % P.raw_slots;
                    if (gc >= P.raw_slots) mbar_wait(&raw_empty[slot], ((gc / P.raw_slots) - 1) & 1);
                    mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)raw_bytes);
                    uint8_t* dst = smem + slot * raw_bytes;
                    tma_load_3d(dst, &tma, 0, r0 + c * AT_KC, P.a_shared ? 0 : b, &raw_full[slot]);
                    tma_load_3d(dst + AT_KC * P.Mtot * 4, &tmb, col0, r0 + c * AT_KC, P.b_shared ? 0 : b, &raw_full[slot]);
                }
            }
        }
    } else if (warp < AT_MMA_WARP) {
        // Converters: warp set s handles the chunks s, s + 2, ...; warp pw of the set the rows [8 pw, 8 pw + 8) of the chunk for
        // every 32-feature block.  Quarter warp q reads the 128 bytes of row 8 pw + q, then of row 8 pw + 4 + q, from the raw
        // chunk; a pair exchange gives every lane one 16-byte unit (8 consecutive features of one row).
        const int set = warp >> 2, pw = warp & 3;
        const int q = lane >> 3, piece = lane & 7;
        const bool odd = piece & 1;
        const int a_fblocks = (P.Mtot + 31) / 32, b_fblocks = P.ncta / 32;
        int c = 0;                                            // global chunk index: ring slots and stages run across items
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
          int b, r0;
          const int nchunks = item_rows(it, b, r0);
          const int n_valid = P.counts ? max(1, min(P.n_rows, __ldg(P.counts + b))) : P.n_rows;
          for (int ci = 0; ci < nchunks; ++ci, ++c) {
            if ((c & (AT_PSETS - 1)) != set) continue;
            const int rows_valid = n_valid - (r0 + ci * AT_KC);       // rows of this chunk that are real points
            const int slot = c % P.raw_slots;
            const int stage = c % AtSmem::G_OP_STAGES;
            mbar_wait(&raw_full[slot], (c / P.raw_slots) & 1);
            const float* rawA = reinterpret_cast<const float*>(smem + slot * raw_bytes) + 4 * piece;
            const float* rawB = rawA + AT_KC * P.Mtot;
            const int kA = 8 * pw + q;
            float4 va[8], vb[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool on = i < a_fblocks;
                va[2 * i] = on ? *reinterpret_cast<const float4*>(rawA + kA * P.Mtot + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                va[2 * i + 1] = on ? *reinterpret_cast<const float4*>(rawA + (kA + 4) * P.Mtot + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool on = i < b_fblocks;
                vb[2 * i] = on ? *reinterpret_cast<const float4*>(rawB + kA * P.ncta + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                vb[2 * i + 1] = on ? *reinterpret_cast<const float4*>(rawB + (kA + 4) * P.ncta + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (c >= AtSmem::G_OP_STAGES) mbar_wait(&empty[stage], ((c / AtSmem::G_OP_STAGES) - 1) & 1);
            uint8_t* st = ops + stage * op_stage;
            const int k = kA + (odd ? 4 : 0);                         // row of the chunk this lane stores
#pragma unroll
            for (int i = 0; i < 4; ++i) {                             // blocks past Mtot are stored as zeros (accumulator rows nobody reads)
                float x[8];
                at_pair_exchange(va[2 * i], va[2 * i + 1], odd, x);
                uint4 hi, lo;
                at_split8(x, hi, lo);
                const int off = (4 * i + (piece >> 1)) * AtSmem::G_SBO + k * 16;
                *reinterpret_cast<uint4*>(st + off) = hi;
                *reinterpret_cast<uint4*>(st + AtSmem::GA_BYTES + off) = lo;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < b_fblocks) {                                  // warp-uniform
                    float x[8];
                    at_pair_exchange(vb[2 * i], vb[2 * i + 1], odd, x);
                    if (k >= rows_valid) {                            // padding row (0 x NaN would poison the sums)
#pragma unroll
                        for (int j = 0; j < 8; ++j) x[j] = 0.f;
                    }
                    uint4 hi, lo;
                    at_split8(x, hi, lo);
                    const int off = (4 * i + (piece >> 1)) * AtSmem::G_SBO + k * 16;
                    *reinterpret_cast<uint4*>(st + 2 * AtSmem::GA_BYTES + off) = hi;
                    *reinterpret_cast<uint4*>(st + 2 * AtSmem::GA_BYTES + gb_bytes + off) = lo;
                }
            }
            at_warp_arrive(&raw_empty[slot]);
            fence_async_smem();
            at_warp_arrive(&full[stage]);
          }
        }
    } else if (warp == AT_MMA_WARP) {
        const uint32_t idesc = idesc_bf16(128, P.ncta, 1, 1);
        int gc = 0, k = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
            int b, r0;
            const int nchunks = item_rows(it, b, r0);
            const uint32_t buf = k & 1;
            if (k >= 2) mbar_wait(&acc_empty[buf], ((k >> 1) - 1) & 1);
            fence_after_sync();
            const uint32_t acc = tmem_addr(tb, 0, 256 * buf);
            for (int c = 0; c < nchunks; ++c, ++gc) {
                const int stage = gc % AtSmem::G_OP_STAGES;
                mbar_wait(&full[stage], (gc / AtSmem::G_OP_STAGES) & 1);
                fence_after_sync();
                if (elect_one()) {
                    const uint32_t a_hi = smem_u32(ops + stage * op_stage);
                    const uint32_t a_lo = a_hi + AtSmem::GA_BYTES;
                    const uint32_t b_hi = a_hi + 2 * AtSmem::GA_BYTES;
                    const uint32_t b_lo = b_hi + gb_bytes;
#pragma unroll
                    for (int ks = 0; ks < AT_KC / 16; ++ks) {
                        // MN-major: a K step of 16 rows = two 8-row groups = 256 bytes further
                        const uint64_t dah = smem_desc(a_hi + ks * 256, 128, AtSmem::G_SBO), dal = smem_desc(a_lo + ks * 256, 128, AtSmem::G_SBO);
                        const uint64_t dbh = smem_desc(b_hi + ks * 256, 128, AtSmem::G_SBO), dbl = smem_desc(b_lo + ks * 256, 128, AtSmem::G_SBO);
                        mma_ss(acc, dah, dbh, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                        mma_ss(acc, dal, dbh, idesc, 1u);
                        mma_ss(acc, dah, dbl, idesc, 1u);
                    }
                    mma_commit(&empty[stage]);
                    if (c == nchunks - 1) mma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp < AT_LOAD_WARP) {
        const int quad = warp & 3;
        float* T = reinterpret_cast<float*>(trans) + quad * (32 * 33);
        int k = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
            int b, r0;
            item_rows(it, b, r0);
            const uint32_t buf = k & 1;
            mbar_wait(&acc_full[buf], (k >> 1) & 1);
            fence_after_sync();
            if (32 * quad < P.Mtot) {
                const int h_lo = (32 * quad) / P.nsp, h_hi = (32 * quad + 31) / P.nsp;      // heads of this warp's accumulator rows
                float* ob = P.out + (long long)b * P.o_bstride;
                for (int c0 = 0; c0 < P.ncta; c0 += 32) {
                    if (!P.full_rows && ((c0 + 31) / P.dh < h_lo || c0 / P.dh > h_hi)) continue;   // warp-uniform: no diagonal block in here
                    uint32_t v[32];
                    tmem_ld32(tmem_addr(tb, 32 * quad, 256 * buf + c0), v);
                    tmem_ld_wait32(v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) T[lane * 33 + j] = __uint_as_float(v[j]);
                    __syncwarp();
                    const int col = c0 + lane;
                    const int hc = col / P.dh;
#pragma unroll 4
                    for (int lr = 0; lr < 32; ++lr) {
                        const int row = 32 * quad + lr;
                        const int h = row / P.nsp, m = row - h * P.nsp;
                        if (P.full_rows) {
                            if (row < P.Mtot) atomicAdd(ob + (long long)row * P.ldo + col, T[lr * 33 + lane]);
                        } else if (row < P.Mtot && m < P.ns && h == hc) {
                            atomicAdd(ob + (long long)m * P.ldo + col, T[lr * 33 + lane]);
                        }
                    }
                    __syncwarp();
                }
            }
            fence_before_sync();
            at_warp_arrive(&acc_empty[buf]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == AT_MMA_WARP) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------ column softmax (points = keys)
// T (B, n_rows, HS) raw scores -> probabilities over the rows of a cloud, in place: P = 2^(s c - lse), lse = log2 sum_r 2^(s c).
// Block = (32 columns, cloud); lane = column (128-byte coalesced rows), warp = row slice.
// key_counts (nullable; variable-size sets): only the first key_counts[b] rows of cloud b are keys, the rest get probability 0.
__global__ void __launch_bounds__(256) col_softmax_kernel(float* __restrict__ T, int n_all, int HS, int nsp, int ns, int H,
                                                          float scale_log2e, float* __restrict__ lse_out,
                                                          const int* __restrict__ key_counts) {
    __shared__ float red[8][33];
    const int b = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    float* t = T + (long long)b * n_all * HS + c;
    const int n_rows = key_counts ? max(1, min(n_all, __ldg(key_counts + b))) : n_all;
    for (int r = n_rows + w; r < n_all; r += 8) t[(long long)r * HS] = 0.f;
    float m = -INFINITY;
#pragma unroll 4
    for (int r = w; r < n_rows; r += 8) m = fmaxf(m, t[(long long)r * HS] * scale_log2e);
    red[w][lane] = m;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) m = fmaxf(m, red[i][lane]);
    __syncthreads();
    float l = 0.f;
#pragma unroll 4
    for (int r = w; r < n_rows; r += 8) l += ex2(fmaf(t[(long long)r * HS], scale_log2e, -m));
    red[w][lane] = l;
    __syncthreads();
    l = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) l += red[i][lane];
    const float lse = m + log2f(l);
#pragma unroll 4
    for (int r = w; r < n_rows; r += 8) t[(long long)r * HS] = ex2(fmaf(t[(long long)r * HS], scale_log2e, -lse));
    if (w == 0 && lse_out) {
        const int h = c / nsp, mm = c - h * nsp;
        if (mm < ns) lse_out[((long long)b * ns + mm) * H + h] = lse;
    }
}

// dst (B, n) = src + b * s_bstride (s_bstride = 0: one row set broadcast over the batch)
__global__ void bcast_rows_kernel(const float* __restrict__ src, long long s_bstride, long long n, long long total, float* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / n;
    dst[i] = __ldg(src + b * s_bstride + (i - b * n));
}

// ------------------------------------------------------------------------------------ host side
static int g_attn_tc = 1;        // 0: attention stays on the CUDA-core kernels (PCA_ATTN_TC=0 / pca_debug_set_attn_tc)
static int g_fold_train = 1;     // 0: training keeps the projected K | V form (pca_debug_set_attn_tc(2) / PCA_ATTN_FOLD_TRAIN=0)
void set_attn_tc(int on) { g_attn_tc = on ? 1 : 0; g_fold_train = (on & 2) ? 0 : 1; }
bool attn_fold_train_on() {
    static int env = -1;
    if (env < 0) { const char* v = getenv("PCA_ATTN_FOLD_TRAIN"); env = (v && v[0] == '0') ? 0 : 1; }
    return env && g_fold_train;
}
static bool attn_tc_on() {
    static int env = -1;
    if (env < 0) { const char* v = getenv("PCA_ATTN_TC"); env = (v && v[0] == '0') ? 0 : 1; }
    return env && g_attn_tc && linear_tc_eligible(512, 32, 32);      // follows the GEMM switch as well
}

struct AtcShape { int type, ns, nsp, HS, big; };     // type 1: keys are the small side (mab1), 2: queries are (mab0 / PMA)

static AtcShape atc_shape(int B, int nq, int nk, int D, int H) {
    AtcShape s{0, 0, 0, 0, 0};
    if (!attn_tc_on() || H <= 0 || D % H) return s;
    const int dh = D / H;
    if (D % 32 || D < 32 || D > 256 || dh % 8) return s;
    int type = 0, ns = 0, big = 0;
    if (nk <= 16 && nq >= 128) { type = 1; ns = nk; big = nq; }
    else if (nq <= 16 && nk >= 128) { type = 2; ns = nq; big = nk; }
    else return s;
    int nsp = 0;
    for (int cand = 8; cand <= 16; cand *= 2)
        if (cand >= ns && (H * cand == 32 || H * cand == 64)) { nsp = cand; break; }
    if (!nsp || (long long)B * big < 512) return s;
    // the (points, HS) score matrix travels through HBM several times: worth it only while it is much narrower than the
    // activations (ModelNet: 64 of 256 columns; the audio models' PMA -- 8 heads x 8 padded columns against D = 64 -- is not)
    if (2 * H * nsp > D) return s;
    s.type = type; s.ns = ns; s.nsp = nsp; s.HS = H * nsp; s.big = big;
    return s;
}

bool attn_tc_eligible(int B, int nq, int nk, int D, int H) { return atc_shape(B, nq, nk, D, H).type != 0; }
// 0: not eligible; 1: the keys are the small side (the backward derives delta itself); 2: the queries are (delta is an input)
int attn_tc_kind(int B, int nq, int nk, int D, int H) { return atc_shape(B, nq, nk, D, H).type; }

static size_t fl(size_t n) { return align_up(n * sizeof(float), 256) / sizeof(float); }

size_t attn_tc_fwd_floats(int B, int nq, int nk, int D, int H) {
    const AtcShape s = atc_shape(B, nq, nk, D, H);
    if (!s.type) return 0;
    const size_t img = (size_t)s.HS * D;             // floats per cloud image (4 N K bytes)
    // (small-query form: + the folded route's query matrix, its image and slack; its per-cloud sums fit the image budget)
    return fl((size_t)B * s.big * s.HS) + (s.type == 1 ? 2 : 1) * fl((size_t)B * img) + (s.type == 2 ? 3 * fl(img) : 0);
}
// floats of the probability matrix a training forward may keep for the backward (0 when the shape is not eligible)
size_t attn_tc_p_floats(int B, int nq, int nk, int D, int H) {
    const AtcShape s = atc_shape(B, nq, nk, D, H);
    return s.type ? (size_t)B * s.big * s.HS : 0;
}
size_t attn_tc_bwd_floats(int B, int nq, int nk, int D, int H) {
    const AtcShape s = atc_shape(B, nq, nk, D, H);
    if (!s.type) return 0;
    const size_t img = (size_t)s.HS * D;
    // (small-query form: + the folded backward's small pieces -- delta, dGq, the G2 image of Gq)
    return 2 * fl((size_t)B * s.big * s.HS) + (s.type == 1 ? 3 : 4) * fl((size_t)B * img) +
           (s.type == 2 ? 3 * fl(img) + fl((size_t)B * s.HS) : 0);
}

static int attn_tc_configure() {
    static std::atomic<unsigned long long> done_mask{0};
    int dev = 0;
    PCA_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && ((done_mask.load(std::memory_order_acquire) >> dev) & 1ull)) return 0;
#define ATC_ATTR(k) PCA_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, AtSmem::MAX_BYTES))
    ATC_ATTR((cloud_linear_tc_kernel<EPI_STORE, 16>));
    ATC_ATTR((cloud_linear_tc_kernel<EPI_RESID, 16>));
    ATC_ATTR((cloud_linear_tc_kernel<EPI_SOFTMAX, 8>));
    ATC_ATTR((cloud_linear_tc_kernel<EPI_SOFTMAX, 16>));
    ATC_ATTR((cloud_linear_tc_kernel<EPI_DS_ROW, 8>));
    ATC_ATTR((cloud_linear_tc_kernel<EPI_DS_ROW, 16>));
    ATC_ATTR((cloud_linear_tc_kernel<EPI_P_COL, 16>));
    ATC_ATTR((cloud_linear_tc_kernel<EPI_DS_COL, 16>));
    ATC_ATTR(cloud_gw_tc_kernel);
#undef ATC_ATTR
    if (dev < 64) done_mask.fetch_or(1ull << dev, std::memory_order_release);
    return 0;
}

static int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

static int launch_cloud_image(const float* src, long long s_bstride, int ld, int nb, const AtcShape& s, int D, int H, int kind, uint8_t* img,
                              cudaStream_t st) {
    const int total = s.HS * D / 8;
    dim3 grid((unsigned)((total + 255) / 256), (unsigned)nb);
    cloud_image_kernel<<<grid, 256, 0, st>>>(src, s_bstride, ld, s.ns, s.nsp, D / H, D, s.HS, kind, img, (long long)s.HS * D * 4);
    PCA_CHECK_LAUNCH("cloud_image_kernel");
    return 0;
}

template <int EPI>
static int launch_cloud_linear(ClinParams p, int nsp, const char* name, cudaStream_t st) {
    const long long ntiles = (long long)p.B * ((p.n_rows + 127) / 128);
    const int sms = sm_count();
    const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
    const double rows = (double)p.B * p.n_rows;
    p.raw_slots = AtSmem::cl_raw_slots(p.N);
    p.x_shared = p.x_bstride == 0 ? 1 : 0;
    const int smem_bytes = AtSmem::cl_total(p.N, p.raw_slots);
    // the activations as (K, rows of a cloud, clouds): a box of 32 floats x 128 rows is one raw tile
    CUtensorMap tmx, tmx2;
    if (!p.X2) p.k_split = p.K / AT_KC;
    const int k1 = p.k_split * AT_KC;
    PCA_TRY(make_tmap_3d_f32(&tmx, p.X, (unsigned long long)k1, (unsigned long long)p.n_rows, p.x_shared ? 1ull : (unsigned long long)p.B,
                             (unsigned long long)p.ldx * 4, (unsigned long long)(p.x_shared ? (long long)p.n_rows * p.ldx : p.x_bstride) * 4,
                             AT_KC, 128));
    if (p.X2)
        PCA_TRY(make_tmap_3d_f32(&tmx2, p.X2, (unsigned long long)(p.K - k1), (unsigned long long)p.n_rows, (unsigned long long)p.B,
                                 (unsigned long long)p.ldx2 * 4, (unsigned long long)p.x2_bstride * 4, AT_KC, 128));
    else
        tmx2 = tmx;
    LaunchTimer lt(name, st, 2.0 * rows * p.K * p.N, 4.0 * rows * (p.K + p.N));
    if (EPI == EPI_SOFTMAX || EPI == EPI_DS_ROW) {
        if (nsp == 8) cloud_linear_tc_kernel<EPI, 8><<<grid, AT_THREADS, smem_bytes, st>>>(p, tmx, tmx2);
        else cloud_linear_tc_kernel<EPI, 16><<<grid, AT_THREADS, smem_bytes, st>>>(p, tmx, tmx2);
    } else {
        cloud_linear_tc_kernel<EPI, 16><<<grid, AT_THREADS, smem_bytes, st>>>(p, tmx, tmx2);
    }
    return 0;
}

static int launch_cloud_gw(const float* T, const float* X, long long x_bstride, int ldx, float* out, long long o_bstride, int ldo, int B,
                           int n_rows, const AtcShape& s, int D, int H, cudaStream_t st, bool one_range, const int* counts = nullptr,
                           int full_rows = 0) {
    const int sms = sm_count();
    // row ranges per cloud: as many as keep an item >= 128 rows and fill the last round of the persistent CTAs best.  one_range
    // (the forward pass): a cloud is ONE item, so every output element receives exactly one addition and the result is bit-
    // reproducible and independent of the batch the cloud sits in (several ranges add in whatever order the CTAs finish).
    const int max_split = one_range ? 1 : (n_rows / 128 > 0 ? n_rows / 128 : 1);
    int nsplit = 1;
    double best = 0.0;
    for (int c = 1; c <= max_split && c <= 8; ++c) {
        const long long items = (long long)B * c;
        const double fill = (double)items / (double)(((items + sms - 1) / sms) * sms);
        if (fill > best + 0.02) { best = fill; nsplit = c; }
    }
    int rchunk = (n_rows + nsplit - 1) / nsplit;
    rchunk = (rchunk + AT_KC - 1) / AT_KC * AT_KC;
    nsplit = (n_rows + rchunk - 1) / rchunk;
    const int ncta = D;
    ClGwParams p{T, (long long)n_rows * s.HS, s.HS, s.HS, X, x_bstride, ldx, D, out, o_bstride, ldo, n_rows, rchunk, s.nsp, s.ns, D / H,
                 AtSmem::g_raw_slots(s.HS, ncta), 0, x_bstride == 0 ? 1 : 0, ncta, B, nsplit, counts, full_rows};
    const long long items = (long long)B * nsplit;
    const unsigned grid = (unsigned)(items < sms ? items : sms);
    // both operands as (columns, rows of a cloud, clouds): one box = the 32 rows of a chunk, all columns
    CUtensorMap tma, tmb;
    PCA_TRY(make_tmap_3d_f32(&tma, T, (unsigned long long)s.HS, (unsigned long long)n_rows, (unsigned long long)B, (unsigned long long)s.HS * 4,
                             (unsigned long long)n_rows * s.HS * 4, (unsigned)s.HS, AT_KC));
    PCA_TRY(make_tmap_3d_f32(&tmb, X, (unsigned long long)D, (unsigned long long)n_rows, p.b_shared ? 1ull : (unsigned long long)B,
                             (unsigned long long)ldx * 4, (unsigned long long)(p.b_shared ? (long long)n_rows * ldx : x_bstride) * 4, (unsigned)ncta,
                             AT_KC));
    {
        LaunchTimer lt("attn_g3_tc_kernel", st, 2.0 * B * (double)n_rows * s.HS * D, 4.0 * B * (double)n_rows * (s.HS + D));
        cloud_gw_tc_kernel<<<grid, AT_THREADS, AtSmem::g_total(s.HS, ncta, p.raw_slots), st>>>(p, tma, tmb);
    }
    PCA_CHECK_LAUNCH("cloud_gw_tc_kernel");
    return 0;
}

// O (B, nq, D) = Qp + softmax_h(Qp K^T / sqrt(D)) V per head; lse (B, nq, H) optional (log2 domain, as attn_f32_kernel writes it).
// scratch: attn_tc_fwd_floats floats.
int launch_attn_tc(const float* Qp, long long q_bstride, const float* KV, int B, int nq, int nk, int D, int H, float* O, float* scratch,
                   cudaStream_t st, float* lse, const int* key_counts, float* p_out) {
    const AtcShape s = atc_shape(B, nq, nk, D, H);
    if (key_counts && s.type != 2) return fail(PCA_EUNSUPPORTED, "attn_tc: key counts need the small-query form");
    if (!s.type) return fail(PCA_EUNSUPPORTED, "attn_tc: shape (B=%d, nq=%d, nk=%d, D=%d, H=%d) not eligible", B, nq, nk, D, H);
    if (!scratch) return fail(PCA_EWORKSPACE, "attn_tc: no scratch");
    PCA_TRY(attn_tc_configure());
    const float scale = 1.0f / sqrtf((float)D), sl2e = scale * 1.4426950408889634f;
    const size_t imgf = (size_t)s.HS * D;
    const long long img_bytes = (long long)imgf * 4;
    float* T = p_out ? p_out : scratch;          // training: the probabilities are kept (the backward does not recompute them)
    uint8_t* img1 = reinterpret_cast<uint8_t*>(scratch + fl((size_t)B * s.big * s.HS));
    if (s.type == 1) {
        uint8_t* img2 = img1 + fl((size_t)B * imgf) * sizeof(float);
        PCA_TRY(launch_cloud_image(KV, (long long)nk * 2 * D, 2 * D, B, s, D, H, 0, img1, st));
        PCA_TRY(launch_cloud_image(KV + D, (long long)nk * 2 * D, 2 * D, B, s, D, H, 1, img2, st));
        ClinParams p{};
        p.X = Qp; p.x_bstride = q_bstride; p.ldx = D;
        p.img = img1; p.img_bstride = img_bytes;
        p.Y = T; p.y_bstride = (long long)nq * s.HS; p.ldy = s.HS;
        p.lse = lse;
        p.B = B; p.n_rows = nq; p.K = D; p.N = s.HS; p.H = H; p.nsp = s.nsp; p.ns = s.ns; p.scale = scale; p.scale_log2e = sl2e;
        PCA_TRY(launch_cloud_linear<EPI_SOFTMAX>(p, s.nsp, "attn_g1_softmax_tc_kernel", st));
        PCA_CHECK_LAUNCH("cloud_linear_tc_kernel<softmax>");
        ClinParams q{};
        q.X = T; q.x_bstride = (long long)nq * s.HS; q.ldx = s.HS;
        q.img = img2; q.img_bstride = img_bytes;
        q.Y = O; q.y_bstride = (long long)nq * D; q.ldy = D;
        q.R = Qp; q.r_bstride = q_bstride; q.ldr = D;
        q.B = B; q.n_rows = nq; q.K = s.HS; q.N = D; q.H = H; q.nsp = s.nsp; q.ns = s.ns; q.scale = scale; q.scale_log2e = sl2e;
        PCA_TRY(launch_cloud_linear<EPI_RESID>(q, s.nsp, "attn_g2_resid_tc_kernel", st));
        PCA_CHECK_LAUNCH("cloud_linear_tc_kernel<resid>");
        return 0;
    }
    const int nb = q_bstride ? B : 1;
    PCA_TRY(launch_cloud_image(Qp, q_bstride, D, nb, s, D, H, 0, img1, st));
    ClinParams p{};
    p.X = KV; p.x_bstride = (long long)nk * 2 * D; p.ldx = 2 * D;
    p.img = img1; p.img_bstride = q_bstride ? img_bytes : 0;
    p.Y = T; p.y_bstride = (long long)nk * s.HS; p.ldy = s.HS;
    p.B = B; p.n_rows = nk; p.K = D; p.N = s.HS; p.H = H; p.nsp = s.nsp; p.ns = s.ns; p.scale = scale; p.scale_log2e = sl2e;
    PCA_TRY(launch_cloud_linear<EPI_STORE>(p, s.nsp, "attn_g1_scores_tc_kernel", st));
    PCA_CHECK_LAUNCH("cloud_linear_tc_kernel<store>");
    {
        LaunchTimer lt("col_softmax_kernel", st, 0.0, 8.0 * B * (double)nk * s.HS);
        col_softmax_kernel<<<dim3((unsigned)(s.HS / 32), (unsigned)B), 256, 0, st>>>(T, nk, s.HS, s.nsp, s.ns, H, sl2e, lse, key_counts);
    }
    PCA_CHECK_LAUNCH("col_softmax_kernel");
    {
        const long long n = (long long)nq * D, total = n * B;
        bcast_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Qp, q_bstride, n, total, O);
        PCA_CHECK_LAUNCH("bcast_rows_kernel");
    }
    return launch_cloud_gw(T, KV + D, (long long)nk * 2 * D, 2 * D, O, (long long)nq * D, D, B, nk, s, D, H, st, true, key_counts);
}

// ------------------------------------------------------------------------------------ inference: W_k / W_v folded away
// Small-query blocks with SHARED queries (ISAB mab0 with I, PMA with S), inference only.  With K = X Wk^T + bk, V = X Wv^T + bv:
//   scores  q_hm . K_n = (Wk_h^T q_hm) . x_n + const(h, m)      -- the constant cancels in the softmax over the points n
//   output  sum_n P_n V_n = Wv (sum_n P_n x_n) + bv              -- the probabilities of a (head, query) sum to one
// so the block runs on the UN-PROJECTED points: G1 against the model-constant image Gq = (per head) Qp_h Wk_h, the column
// softmax, G3 with full rows (Z = P^T X, (HS, dk) per cloud) and a small projection of Z's rows by their head's slice of Wv.
// The (B n, dk) -> 2 D projection of the block -- the largest linear layer of the model -- and the K | V tensor never exist
// (the identity the bf16 kernels of encoder_tc.cu use for mab0 and the pooled attention).
__global__ void fold_query_kernel(const float* __restrict__ Qp, const float* __restrict__ Wk, int ns, int nsp, int dh, int D, int dk,
                                  float* __restrict__ Gq) {
    const int row = blockIdx.x;                      // (h, m)
    const int h = row / nsp, m = row - h * nsp;
    for (int k = threadIdx.x; k < dk; k += blockDim.x) {
        float a = 0.f;
        if (m < ns)
            for (int d = 0; d < dh; ++d) a = fmaf(__ldg(Qp + (long long)m * D + h * dh + d), __ldg(Wk + (long long)(h * dh + d) * dk + k), a);
        Gq[(long long)row * dk + k] = a;
    }
}
// O[b][m][d] = Qp[m][d] + bv[d] + sum_k Z[b][(d / dh, m)][k] Wv[d][k]: per head a (B ns, dk) x (dk, dh) product.  Block = (head,
// FOLD_CB clouds): the head's slice of Wv is staged once per block and 32-wide k tile and reused by all the block's rows (the first
// version, one block per (head, cloud), re-read it 1024 times and took 143 us per launch); thread = (dim of the head, row lane).
constexpr int FOLD_CB = 8;
__global__ void __launch_bounds__(256) fold_proj_kernel(const float* __restrict__ Z, const float* __restrict__ Wv, const float* __restrict__ bv,
                                                        const float* __restrict__ Qp, float* __restrict__ O, int B, int ns, int nsp, int HS,
                                                        int dk, int D, int dh) {
    __shared__ float zs[FOLD_CB * 16 * 33];
    __shared__ float wsm[64 * 33];
    const int h = blockIdx.x, b0 = blockIdx.y * FOLD_CB;
    const int nb = min(FOLD_CB, B - b0), rows = nb * ns;
    const int dl = threadIdx.x % dh, ml = threadIdx.x / dh, mlanes = 256 / dh;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    for (int k0 = 0; k0 < dk; k0 += 32) {
        __syncthreads();
        for (int i = threadIdx.x; i < dh * 32; i += 256) {
            const int dd = i >> 5, kk = i & 31;
            wsm[dd * 33 + kk] = __ldg(Wv + (long long)(h * dh + dd) * dk + k0 + kk);
        }
        for (int i = threadIdx.x; i < rows * 32; i += 256) {
            const int r = i >> 5, kk = i & 31;
            const int b = b0 + r / ns, m = r - (r / ns) * ns;
            zs[r * 33 + kk] = Z[((long long)b * HS + h * nsp + m) * dk + k0 + kk];
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < 32; ++kk) {
            const float w = wsm[dl * 33 + kk];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int r = ml + i * mlanes;
                if (r < rows) acc[i] = fmaf(zs[r * 33 + kk], w, acc[i]);
            }
        }
    }
    const int d = h * dh + dl;
    const float bias = __ldg(bv + d);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int r = ml + i * mlanes;
        if (r < rows) {
            const int b = b0 + r / ns, m = r - (r / ns) * ns;
            O[((long long)b * ns + m) * D + d] = __ldg(Qp + (long long)m * D + d) + bias + acc[i];
        }
    }
}

bool attn_fold_eligible(int B, int nq, int nk, int dk, int D, int H) {
    const AtcShape s = atc_shape(B, nq, nk, D, H);
    if (s.type != 2) return false;
    const int dh = D / H;
    return dk % 32 == 0 && dk >= 64 && dk <= 256 && dk <= D && 2 * s.HS <= dk && (dh == 8 || dh == 16 || dh == 32 || dh == 64) && s.ns <= 16;
}

// O (B, nq, D) = Qp + softmax_h(Qp K^T / sqrt(D)) V with K | V = X Wkv^T + bkv never formed.  Qp (nq, D) shared by the batch;
// Wkv (2 D, dk) = Wk rows then Wv rows, bkv (2 D); X (B, nk, dk).  scratch: attn_tc_fwd_floats floats.
int launch_attn_folded(const float* Qp, const float* Wkv, const float* bkv, const float* X, int B, int nq, int nk, int dk, int D, int H,
                       float* O, float* scratch, cudaStream_t st, const int* key_counts, float* p_out, float* z_out, float* gq_out) {
    const AtcShape s = atc_shape(B, nq, nk, D, H);
    if (!attn_fold_eligible(B, nq, nk, dk, D, H)) return fail(PCA_EUNSUPPORTED, "attn_folded: shape not eligible");
    if (!scratch) return fail(PCA_EWORKSPACE, "attn_folded: no scratch");
    PCA_TRY(attn_tc_configure());
    const float scale = 1.0f / sqrtf((float)D), sl2e = scale * 1.4426950408889634f;
    const size_t imgf = (size_t)s.HS * D;
    // training keeps P, Z and Gq for the backward (p_out / z_out / gq_out); inference uses the scratch
    float* T = p_out ? p_out : scratch;
    float* Z = z_out ? z_out : scratch + fl((size_t)B * nk * s.HS);   // (B, HS, dk) <= the per-cloud image budget (dk <= D)
    float* Gq = gq_out ? gq_out : scratch + fl((size_t)B * nk * s.HS) + fl((size_t)B * imgf);
    uint8_t* img = reinterpret_cast<uint8_t*>(scratch + fl((size_t)B * nk * s.HS) + fl((size_t)B * imgf) + fl(imgf));
    const int dh = D / H;
    fold_query_kernel<<<s.HS, 256, 0, st>>>(Qp, Wkv, s.ns, s.nsp, dh, D, dk, Gq);
    PCA_CHECK_LAUNCH("fold_query_kernel");
    {
        const int total = s.HS * dk / 8;
        cloud_image_kernel<<<dim3((unsigned)((total + 255) / 256), 1), 256, 0, st>>>(Gq, 0, dk, s.ns, s.nsp, dh, dk, s.HS, 2, img, 0);
        PCA_CHECK_LAUNCH("cloud_image_kernel");
    }
    ClinParams p{};
    p.X = X; p.x_bstride = (long long)nk * dk; p.ldx = dk;
    p.img = img; p.img_bstride = 0;
    p.Y = T; p.y_bstride = (long long)nk * s.HS; p.ldy = s.HS;
    p.B = B; p.n_rows = nk; p.K = dk; p.N = s.HS; p.H = H; p.nsp = s.nsp; p.ns = s.ns; p.scale = scale; p.scale_log2e = sl2e;
    PCA_TRY(launch_cloud_linear<EPI_STORE>(p, s.nsp, "attn_g1_scores_tc_kernel", st));
    PCA_CHECK_LAUNCH("cloud_linear_tc_kernel<store>");
    {
        LaunchTimer lt("col_softmax_kernel", st, 0.0, 8.0 * B * (double)nk * s.HS);
        col_softmax_kernel<<<dim3((unsigned)(s.HS / 32), (unsigned)B), 256, 0, st>>>(T, nk, s.HS, s.nsp, s.ns, H, sl2e, nullptr, key_counts);
    }
    PCA_CHECK_LAUNCH("col_softmax_kernel");
    PCA_CHECK_CUDA(cudaMemsetAsync(Z, 0, (size_t)B * s.HS * dk * sizeof(float), st));
    PCA_TRY(launch_cloud_gw(T, X, (long long)nk * dk, dk, Z, (long long)s.HS * dk, dk, B, nk, s, dk, H, st, true, key_counts, 1));
    fold_proj_kernel<<<dim3((unsigned)H, (unsigned)((B + FOLD_CB - 1) / FOLD_CB)), 256, 0, st>>>(Z, Wkv + (long long)D * dk, bkv + D, Qp, O, B,
                                                                                                   s.ns, s.nsp, s.HS, dk, D, dh);
    PCA_CHECK_LAUNCH("fold_proj_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------ training: the folded block's backward
// With S = X Gq^T (Gq = per head Qp_h Wk_h), P = column softmax, Z = P^T X, O_hm = Qp_hm + bv_h + Wv_h Z_hm and dO given:
//   dZ'_hm = Wv_h^T dO_hm                      (fold_dz_kernel; per cloud (HS, dk))
//   a = X dZ'^T,  dS = P o (a - Z_hm . dZ'_hm) / sqrt(D)     (G1 <ds_col>: the row dots are the softmax backward's delta; dO . bv
//                                                 is constant over the points and cancels)
//   dX  = dS Gq + P dZ'                        (two G2 launches; the second accumulates)
//   dGq = sum_b dS^T X                         (G3, full rows, one matrix for the batch)  ->  dWk_h += Qp_h^T dGq_h, dQp_h = dGq_h Wk_h^T
//   dWv_h += sum_b dO_h^T Z_h, dbv += sum dO   (by the caller: a weight-gradient GEMM on the block-expanded dO, fold_dox_kernel)
// The K | V projection, its 512-wide weight-gradient GEMM and its input-gradient GEMM do not exist in this form.
// Per head a (B ns, dh) x (dh, dk) product.  Block = (head, FOLD_CB clouds, 256 columns k): the thread keeps its column of the head's
// Wv slice in registers, the block's dO rows sit in shared memory (broadcast reads); the padded rows (m >= ns) are written as zeros.
template <int DH>
__global__ void __launch_bounds__(256) fold_dz_kernel(const float* __restrict__ dO, const float* __restrict__ Wv, int B, int ns, int nsp, int D,
                                                      int dk, int HS, float* __restrict__ dZ) {
    __shared__ __align__(16) float gs[FOLD_CB * 16 * DH];
    const int h = blockIdx.x, b0 = blockIdx.y * FOLD_CB;
    const int nb = min(FOLD_CB, B - b0), rows = nb * ns;
    const int k = blockIdx.z * 256 + threadIdx.x;
    for (int i = threadIdx.x; i < rows * DH; i += 256) {
        const int r = i / DH, d = i - r * DH;
        const int b = b0 + r / ns, m = r - (r / ns) * ns;
        gs[i] = __ldg(dO + ((long long)b * ns + m) * D + h * DH + d);
    }
    float w[DH];
    if (k < dk) {
#pragma unroll
        for (int d = 0; d < DH; ++d) w[d] = __ldg(Wv + (long long)(h * DH + d) * dk + k);
    }
    __syncthreads();
    if (k >= dk) return;
    for (int r = 0; r < rows; ++r) {
        const float4* g4 = reinterpret_cast<const float4*>(gs + r * DH);
        float a = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < DH / 4; ++d4) {
            const float4 g = g4[d4];
            a = fmaf(g.x, w[4 * d4], a); a = fmaf(g.y, w[4 * d4 + 1], a); a = fmaf(g.z, w[4 * d4 + 2], a); a = fmaf(g.w, w[4 * d4 + 3], a);
        }
        const int b = b0 + r / ns, m = r - (r / ns) * ns;
        dZ[((long long)b * HS + h * nsp + m) * dk + k] = a;
    }
    for (int bb = 0; bb < nb; ++bb)
        for (int m = ns; m < nsp; ++m) dZ[((long long)(b0 + bb) * HS + h * nsp + m) * dk + k] = 0.f;
}
// delta[(b ns + m) H + h] = Z[b, (h, m), :] . dZ'[b, (h, m), :]; one warp per row
__global__ void fold_delta_kernel(const float* __restrict__ Z, const float* __restrict__ dZ, int B, int ns, int nsp, int H, int HS, int dk,
                                  float* __restrict__ delta) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= B * HS) return;
    const int b = w / HS, row = w - b * HS;
    const int h = row / nsp, m = row - h * nsp;
    if (m >= ns) return;
    float a = 0.f;
    for (int k = lane; k < dk; k += 32) a = fmaf(Z[(long long)w * dk + k], dZ[(long long)w * dk + k], a);
    a = warp_sum(a);
    if (lane == 0) delta[((long long)b * ns + m) * H + h] = a;
}
// dWk[d, k] += sum_m Qp[m, d] dGq[(d / dh, m), k]
__global__ void fold_dwk_kernel(const float* __restrict__ Qp, const float* __restrict__ dGq, int ns, int nsp, int dh, int D, int dk,
                                float* __restrict__ dWk) {
    const int d = blockIdx.x, h = d / dh;
    for (int k = threadIdx.x; k < dk; k += blockDim.x) {
        float a = 0.f;
        for (int m = 0; m < ns; ++m) a = fmaf(__ldg(Qp + (long long)m * D + d), dGq[(long long)(h * nsp + m) * dk + k], a);
        dWk[(long long)d * dk + k] += a;
    }
}
// dQ[m, d] = sum_k dGq[(d / dh, m), k] Wk[d, k]; one warp per (m, d)
__global__ void fold_dq_kernel(const float* __restrict__ dGq, const float* __restrict__ Wk, int ns, int nsp, int dh, int D, int dk,
                               float* __restrict__ dQ) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= ns * D) return;
    const int m = w / D, d = w - m * D, h = d / dh;
    float a = 0.f;
    for (int k = lane; k < dk; k += 32) a = fmaf(dGq[(long long)(h * nsp + m) * dk + k], __ldg(Wk + (long long)d * dk + k), a);
    a = warp_sum(a);
    if (lane == 0) dQ[w] = a;
}
// dOx[(b, h', m), d] = dO[b, m, d] when d / dh == h' and m < ns, else 0: (B HS, D), the rows of Z's layout
__global__ void fold_dox_kernel(const float* __restrict__ dO, int ns, int nsp, int dh, int D, int HS, long long total, float* __restrict__ dOx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int d = (int)(i % D);
    const long long r = i / D;
    const int row = (int)(r % HS);
    const long long b = r / HS;
    const int h = row / nsp, m = row - h * nsp;
    dOx[i] = (m < ns && d / dh == h) ? __ldg(dO + (b * ns + m) * D + d) : 0.f;
}

size_t attn_fold_z_floats(int B, int nq, int nk, int D, int H) {          // saved (B, HS, dk <= D) sums of a folded training block
    const AtcShape s = atc_shape(B, nq, nk, D, H);
    return s.type == 2 ? (size_t)B * s.HS * D : 0;
}
size_t attn_fold_gq_floats(int B, int nq, int nk, int D, int H) {
    const AtcShape s = atc_shape(B, nq, nk, D, H);
    return s.type == 2 ? (size_t)s.HS * D : 0;
}
int attn_fold_rows(int B, int nq, int nk, int D, int H) { return atc_shape(B, nq, nk, D, H).HS; }

// dX (B, nk, dk): OVERWRITTEN (x_acc = 0) or accumulated onto (x_acc = 1), nullable; dWk (D, dk) accumulated; dQ (nq, D): the
// attention part of the shared queries' gradient (OVERWRITTEN; the caller adds the residual part sum_b dO); *dox_out / Z: the
// operands of the caller's Wv weight-gradient GEMM ((B HS, D) and (B HS, dk)).  scratch: attn_tc_bwd_floats floats.
int launch_attn_folded_bwd(const float* Qp, const float* Wkv, const float* X, const float* dO, const float* P, const float* Z,
                           const float* Gq, int B, int nq, int nk, int dk, int D, int H, float* dX, int x_acc, float* dWk, float* dQ,
                           float** dox_out, float* scratch, cudaStream_t st) {
    const AtcShape s = atc_shape(B, nq, nk, D, H);
    if (!attn_fold_eligible(B, nq, nk, dk, D, H)) return fail(PCA_EUNSUPPORTED, "attn_folded_bwd: shape not eligible");
    if (!scratch) return fail(PCA_EWORKSPACE, "attn_folded_bwd: no scratch");
    PCA_TRY(attn_tc_configure());
    const float scale = 1.0f / sqrtf((float)D), sl2e = scale * 1.4426950408889634f;
    const size_t imgf = (size_t)s.HS * D;
    const size_t tf = fl((size_t)B * nk * s.HS), imf = fl((size_t)B * imgf);
    const int dh = D / H;
    float* dS = scratch;                                   // (B nk, HS)   [the second score-sized region stays unused here]
    float* big = scratch + 2 * tf;
    float* small = big + 4 * imf;
    float* delta = small;                                  // (B, ns, H)
    float* dGq = delta + fl((size_t)B * s.HS);             // (HS, dk)
    uint8_t* img_gq2 = reinterpret_cast<uint8_t*>(dGq + fl(imgf));       // Gq as the G2 operand
    float* dZ = big;                                       // (B, HS, dk)
    uint8_t* img_dz1 = reinterpret_cast<uint8_t*>(big + imf);            // dZ' as the G1 operand, per cloud
    uint8_t* img_dz2 = reinterpret_cast<uint8_t*>(big + 2 * imf);        // dZ' as the G2 operand, per cloud
    float* dOx = big + 3 * imf;                            // (B HS, D)
    const long long img_bytes = (long long)s.HS * dk * 4;
    const float* Wk = Wkv;
    const float* Wv = Wkv + (long long)D * dk;

    {
        const dim3 gz((unsigned)H, (unsigned)((B + FOLD_CB - 1) / FOLD_CB), (unsigned)((dk + 255) / 256));
        switch (dh) {
            case 8: fold_dz_kernel<8><<<gz, 256, 0, st>>>(dO, Wv, B, s.ns, s.nsp, D, dk, s.HS, dZ); break;
            case 16: fold_dz_kernel<16><<<gz, 256, 0, st>>>(dO, Wv, B, s.ns, s.nsp, D, dk, s.HS, dZ); break;
            case 32: fold_dz_kernel<32><<<gz, 256, 0, st>>>(dO, Wv, B, s.ns, s.nsp, D, dk, s.HS, dZ); break;
            case 64: fold_dz_kernel<64><<<gz, 256, 0, st>>>(dO, Wv, B, s.ns, s.nsp, D, dk, s.HS, dZ); break;
            default: return fail(PCA_EUNSUPPORTED, "attn_folded_bwd: head dim %d not in {8, 16, 32, 64}", dh);
        }
        PCA_CHECK_LAUNCH("fold_dz_kernel");
    }
    fold_delta_kernel<<<(unsigned)((B * s.HS + 7) / 8), 256, 0, st>>>(Z, dZ, B, s.ns, s.nsp, H, s.HS, dk, delta);
    PCA_CHECK_LAUNCH("fold_delta_kernel");
    {
        const int total = s.HS * dk / 8;
        const dim3 gb((unsigned)((total + 255) / 256), (unsigned)B), g1((unsigned)((total + 255) / 256), 1);
        cloud_image_kernel<<<gb, 256, 0, st>>>(dZ, (long long)s.HS * dk, dk, s.ns, s.nsp, dh, dk, s.HS, 2, img_dz1, img_bytes);
        PCA_CHECK_LAUNCH("cloud_image_kernel");
        if (dX) {
            cloud_image_kernel<<<gb, 256, 0, st>>>(dZ, (long long)s.HS * dk, dk, s.ns, s.nsp, dh, dk, s.HS, 3, img_dz2, img_bytes);
            PCA_CHECK_LAUNCH("cloud_image_kernel");
            cloud_image_kernel<<<g1, 256, 0, st>>>(Gq, 0, dk, s.ns, s.nsp, dh, dk, s.HS, 3, img_gq2, 0);
            PCA_CHECK_LAUNCH("cloud_image_kernel");
        }
    }
    const long long t_bs = (long long)nk * s.HS, x_bs = (long long)nk * dk;
    auto base = [&](int K, int N) {
        ClinParams p{};
        p.B = B; p.n_rows = nk; p.K = K; p.N = N; p.H = H; p.nsp = s.nsp; p.ns = s.ns; p.scale = scale; p.scale_log2e = sl2e;
        return p;
    };
    ClinParams q = base(dk, s.HS);                          // dS = P o (X dZ'^T - delta) / sqrt(D)
    q.X = X; q.x_bstride = x_bs; q.ldx = dk;
    q.img = img_dz1; q.img_bstride = img_bytes;
    q.Y = dS; q.y_bstride = t_bs; q.ldy = s.HS;
    q.R = P; q.r_bstride = t_bs; q.ldr = s.HS;
    q.vec = delta;
    PCA_TRY(launch_cloud_linear<EPI_DS_COL>(q, s.nsp, "attn_g1_dscol_tc_kernel", st));
    PCA_CHECK_LAUNCH("cloud_linear_tc_kernel<ds_col>");
    if (dX) {
        ClinParams g = base(2 * s.HS, dk);                  // dX (+)= dS Gq + P dZ': ONE launch, the K chunks of two sources
        g.X = dS; g.x_bstride = t_bs; g.ldx = s.HS;
        g.img = img_gq2; g.img_bstride = 0;
        g.X2 = P; g.x2_bstride = t_bs; g.ldx2 = s.HS;
        g.img2 = img_dz2; g.img2_bstride = img_bytes;
        g.k_split = s.HS / AT_KC;
        g.Y = dX; g.y_bstride = x_bs; g.ldy = dk;
        if (x_acc) {
            g.R = dX; g.r_bstride = x_bs; g.ldr = dk;
            PCA_TRY(launch_cloud_linear<EPI_RESID>(g, s.nsp, "attn_g2_resid_tc_kernel", st));
        } else {
            PCA_TRY(launch_cloud_linear<EPI_STORE>(g, s.nsp, "attn_g2_store_tc_kernel", st));
        }
        PCA_CHECK_LAUNCH("cloud_linear_tc_kernel<g2>");
    }
    PCA_CHECK_CUDA(cudaMemsetAsync(dGq, 0, (size_t)s.HS * dk * sizeof(float), st));
    PCA_TRY(launch_cloud_gw(dS, X, x_bs, dk, dGq, 0, dk, B, nk, s, dk, H, st, false, nullptr, 1));      // dGq = sum_b dS^T X
    fold_dwk_kernel<<<(unsigned)D, 256, 0, st>>>(Qp, dGq, s.ns, s.nsp, dh, D, dk, dWk);
    PCA_CHECK_LAUNCH("fold_dwk_kernel");
    fold_dq_kernel<<<(unsigned)((s.ns * D + 7) / 8), 256, 0, st>>>(dGq, Wk, s.ns, s.nsp, dh, D, dk, dQ);
    PCA_CHECK_LAUNCH("fold_dq_kernel");
    {
        const long long total = (long long)B * s.HS * D;
        fold_dox_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dO, s.ns, s.nsp, dh, D, s.HS, total, dOx);
        PCA_CHECK_LAUNCH("fold_dox_kernel");
    }
    *dox_out = dOx;
    return 0;
}

// Gradients of the attention above.  dQp (B, nq, D) is OVERWRITTEN with dO (the residual path) + the attention part; dKV
// (B, nk, 2D) is OVERWRITTEN.  delta (B, nq, H) = sum_d dO (O - Qp) per head is needed only when the queries are the small side.
// scratch: attn_tc_bwd_floats floats.
int launch_attn_bwd_tc(const float* Qp, long long q_bstride, const float* KV, const float* dO, const float* lse, const float* delta,
                       int B, int nq, int nk, int D, int H, float* dQp, float* dKV, float* scratch, cudaStream_t st, const float* p_saved,
                       float* db_q, float* db_kv) {
    const AtcShape s = atc_shape(B, nq, nk, D, H);
    if (!s.type) return fail(PCA_EUNSUPPORTED, "attn_bwd_tc: shape (B=%d, nq=%d, nk=%d, D=%d, H=%d) not eligible", B, nq, nk, D, H);
    if (!scratch) return fail(PCA_EWORKSPACE, "attn_bwd_tc: no scratch");
    PCA_TRY(attn_tc_configure());
    const float scale = 1.0f / sqrtf((float)D), sl2e = scale * 1.4426950408889634f;
    const size_t imgf = (size_t)s.HS * D;
    const long long img_bytes = (long long)imgf * 4;
    const size_t tf = fl((size_t)B * s.big * s.HS), imf = fl((size_t)B * imgf);
    const float* Pm = p_saved ? p_saved : scratch;        // the forward's probabilities when it kept them, else recomputed below
    float* Pw = scratch;
    float* dS = scratch + tf;
    uint8_t* imgs = reinterpret_cast<uint8_t*>(scratch + 2 * tf);
    auto img_at = [&](int i) { return imgs + (size_t)i * imf * sizeof(float); };
    const long long t_bstride = (long long)s.big * s.HS;
    auto base = [&](int n_rows, int K, int N) {
        ClinParams p{};
        p.B = B; p.n_rows = n_rows; p.K = K; p.N = N; p.H = H; p.nsp = s.nsp; p.ns = s.ns; p.scale = scale; p.scale_log2e = sl2e;
        return p;
    };
    if (s.type == 1) {
        const long long kv_bs = (long long)nk * 2 * D;
        if (!p_saved) PCA_TRY(launch_cloud_image(KV, kv_bs, 2 * D, B, s, D, H, 0, img_at(0), st));   // K as the G1 operand
        PCA_TRY(launch_cloud_image(KV + D, kv_bs, 2 * D, B, s, D, H, 0, img_at(1), st));      // V as the G1 operand
        PCA_TRY(launch_cloud_image(KV, kv_bs, 2 * D, B, s, D, H, 1, img_at(2), st));          // K as the G2 operand
        if (!p_saved) {
            ClinParams p = base(nq, D, s.HS);                                                 // P = softmax(Qp K^T)
            p.X = Qp; p.x_bstride = q_bstride; p.ldx = D;
            p.img = img_at(0); p.img_bstride = img_bytes;
            p.Y = Pw; p.y_bstride = t_bstride; p.ldy = s.HS;
            PCA_TRY(launch_cloud_linear<EPI_SOFTMAX>(p, s.nsp, "attn_g1_softmax_tc_kernel", st));
            PCA_CHECK_LAUNCH("cloud_linear_tc_kernel<softmax>");
        }
        ClinParams q = base(nq, D, s.HS);                                                     // dS = P o (dO V^T - delta) scale
        q.X = dO; q.x_bstride = (long long)nq * D; q.ldx = D;
        q.img = img_at(1); q.img_bstride = img_bytes;
        q.Y = dS; q.y_bstride = t_bstride; q.ldy = s.HS;
        q.R = Pm; q.r_bstride = t_bstride; q.ldr = s.HS;
        PCA_TRY(launch_cloud_linear<EPI_DS_ROW>(q, s.nsp, "attn_g1_dsrow_tc_kernel", st));
        PCA_CHECK_LAUNCH("cloud_linear_tc_kernel<ds_row>");
        ClinParams g = base(nq, s.HS, D);                                                     // dQp = dO + dS K
        g.X = dS; g.x_bstride = t_bstride; g.ldx = s.HS;
        g.img = img_at(2); g.img_bstride = img_bytes;
        g.Y = dQp; g.y_bstride = (long long)nq * D; g.ldy = D;
        g.R = dO; g.r_bstride = (long long)nq * D; g.ldr = D;
        g.colsum = db_q;                                                                      // (+= column sums of dQp: fc_q's bias gradient)
        PCA_TRY(launch_cloud_linear<EPI_RESID>(g, s.nsp, "attn_g2_resid_tc_kernel", st));
        PCA_CHECK_LAUNCH("cloud_linear_tc_kernel<resid>");
        PCA_CHECK_CUDA(cudaMemsetAsync(dKV, 0, (size_t)B * nk * 2 * D * sizeof(float), st));
        PCA_TRY(launch_cloud_gw(dS, Qp, q_bstride, D, dKV, kv_bs, 2 * D, B, nq, s, D, H, st, false));         // dK = dS^T Qp
        return launch_cloud_gw(Pm, dO, (long long)nq * D, D, dKV + D, kv_bs, 2 * D, B, nq, s, D, H, st, false);   // dV = P^T dO
    }
    if (!delta || (!lse && !p_saved)) return fail(PCA_EINVAL, "attn_bwd_tc: the small-query form needs delta and lse (or the saved probabilities)");
    const int nb = q_bstride ? B : 1;
    const long long kv_bs = (long long)nk * 2 * D, o_bs = (long long)nq * D;
    if (!p_saved) PCA_TRY(launch_cloud_image(Qp, q_bstride, D, nb, s, D, H, 0, img_at(0), st));   // Qp as the G1 operand
    PCA_TRY(launch_cloud_image(Qp, q_bstride, D, nb, s, D, H, 1, img_at(1), st));             // Qp as the G2 operand
    PCA_TRY(launch_cloud_image(dO, o_bs, D, B, s, D, H, 0, img_at(2), st));                   // dO as the G1 operand
    PCA_TRY(launch_cloud_image(dO, o_bs, D, B, s, D, H, 1, img_at(3), st));                   // dO as the G2 operand
    if (!p_saved) {
        ClinParams p = base(nk, D, s.HS);                                                     // P = 2^(K Qp^T c - lse)
        p.X = KV; p.x_bstride = kv_bs; p.ldx = 2 * D;
        p.img = img_at(0); p.img_bstride = q_bstride ? img_bytes : 0;
        p.Y = Pw; p.y_bstride = t_bstride; p.ldy = s.HS;
        p.vec = lse;
        PCA_TRY(launch_cloud_linear<EPI_P_COL>(p, s.nsp, "attn_g1_pcol_tc_kernel", st));
        PCA_CHECK_LAUNCH("cloud_linear_tc_kernel<p_col>");
    }
    ClinParams q = base(nk, D, s.HS);                                                         // dS = P o (V dO^T - delta) scale
    q.X = KV + D; q.x_bstride = kv_bs; q.ldx = 2 * D;
    q.img = img_at(2); q.img_bstride = img_bytes;
    q.Y = dS; q.y_bstride = t_bstride; q.ldy = s.HS;
    q.R = Pm; q.r_bstride = t_bstride; q.ldr = s.HS;
    q.vec = delta;
    PCA_TRY(launch_cloud_linear<EPI_DS_COL>(q, s.nsp, "attn_g1_dscol_tc_kernel", st));
    PCA_CHECK_LAUNCH("cloud_linear_tc_kernel<ds_col>");
    ClinParams gk = base(nk, s.HS, D);                                                        // dK = dS Qp
    gk.X = dS; gk.x_bstride = t_bstride; gk.ldx = s.HS;
    gk.img = img_at(1); gk.img_bstride = q_bstride ? img_bytes : 0;
    gk.Y = dKV; gk.y_bstride = kv_bs; gk.ldy = 2 * D;
    gk.colsum = db_kv;                                                                        // (+= fc_k | fc_v bias gradients)
    PCA_TRY(launch_cloud_linear<EPI_STORE>(gk, s.nsp, "attn_g2_store_tc_kernel", st));
    PCA_CHECK_LAUNCH("cloud_linear_tc_kernel<store>");
    ClinParams gv = base(nk, s.HS, D);                                                        // dV = P dO
    gv.X = Pm; gv.x_bstride = t_bstride; gv.ldx = s.HS;
    gv.img = img_at(3); gv.img_bstride = img_bytes;
    gv.Y = dKV + D; gv.y_bstride = kv_bs; gv.ldy = 2 * D;
    gv.colsum = db_kv ? db_kv + D : nullptr;
    PCA_TRY(launch_cloud_linear<EPI_STORE>(gv, s.nsp, "attn_g2_store_tc_kernel", st));
    PCA_CHECK_LAUNCH("cloud_linear_tc_kernel<store>");
    PCA_CHECK_CUDA(cudaMemcpyAsync(dQp, dO, (size_t)B * nq * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return launch_cloud_gw(dS, KV, kv_bs, 2 * D, dQp, o_bs, D, B, nk, s, D, H, st, false);    // dQp = dO + dS^T K
}

}  // namespace pca
