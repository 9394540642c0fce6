// fp32-grade GEMMs on the tcgen05 tensor cores for the shared-MLP / projection layers over points
// (Y = act(X W^T + b), dX = dY W, dW = dY^T X) of the dims the fused bf16 encoder does not cover (ModelNet D=256, DeepSet,
// training).  Operands are split into bf16 hi + lo parts on the fly (x = hi + lo to 2^-17 relative) and every product is
// evaluated as THREE bf16 MMAs accumulating in fp32 TMEM (a_hi b_hi + a_lo b_hi + a_hi b_lo; the dropped lo*lo term is
// 2^-18 relative), so the results stay inside the 1e-3 fp32 parity budget while the arithmetic runs on the tensor pipe.
//
//   linear_tc_kernel  : persistent, warp-specialised, 14 warps.  ONE loader thread streams the fp32 rows with TMA into a ring of
//                       raw 16 KB tiles (128 rows x 32 floats; up to 8 in flight per SM), 8 converter warps (two sets on alternate
//                       K chunks: raw rows -> hi/lo bf16 -> canonical no-swizzle smem with a padded chunk stride), the weight
//                       operand arrives as a pre-built hi|lo image by ONE bulk async copy (cp.async.bulk + mbarrier
//                       complete_tx) per K chunk, 1 MMA warp (elected lane), 4 epilogue warps (TMEM -> bias / ReLU / residual ->
//                       global).  Two TMEM accumulators of up to 256 columns: the epilogue of tile t overlaps the staging and
//                       MMAs of tile t+1.
//   weight_image_kernel: W (fp32, optionally transposed) -> per (column pass, K chunk) [hi image | lo image].
//   grad_weight_tc_kernel: dW (M, N) += A^T B (both operands MN-major from row-major activations, TMA raw-chunk ring +
//                       converter warps), persistent CTAs over (row range, 128-row block of dW) items with two accumulators,
//                       partial results added atomically.
#include "common.cuh"
#include "tc_prims.cuh"
#include "tma_host.cuh"

namespace pca {
using namespace tc;

constexpr int GT_KC = 32;            // K elements per pipeline stage
constexpr int GT_PSETS = 2;           // producer warp sets: set s stages the work items s, s + 2, ... (twice the loads in flight per SM)
constexpr int GT_MMA_WARP = 4 * GT_PSETS;
constexpr int GT_LOAD_WARP = GT_MMA_WARP + 5;
constexpr int LT_THREADS = (4 * GT_PSETS + 6) * 32;   // linear kernel: 8 converter + 1 MMA + 4 epilogue + 1 TMA loader warps
constexpr int LT_STAGES = 3;          // operand stages of the linear kernel
constexpr int LT_RAW_MAX = 8;         // raw fp32 tiles (128 rows x 32 floats = 16 KB) the TMA loader may have in flight
constexpr int LT_RAW_BYTES = 128 * GT_KC * 4;

__device__ __forceinline__ void gt_warp_arrive(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
// 8 fp32 -> 8 bf16 hi (one 16-byte chunk) and 8 bf16 lo
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
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

// ------------------------------------------------------------------------------------ weight image
// B(n, k) = trans ? W[k * ldw + n] : W[n * ldw + k], n < N, k < K.  Image (pass p, chunk c) at ((p * nkc + c) * 2) * nt * 64 bytes:
// hi image (K-major canonical: byte = (kk / 8) * nt * 16 + nn * 16 + (kk % 8) * 2, nn < nt, kk < 32) then lo image.
__global__ void weight_image_kernel(const float* __restrict__ W, int N, int K, int ldw, int trans, int nt, uint8_t* __restrict__ img) {
    const int nkc = K / GT_KC;
    const long long total = (long long)N * (K / 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i % N), k8 = (int)(i / N);
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k8 * 8 + j;
            x[j] = trans ? __ldg(W + (long long)k * ldw + n) : __ldg(W + (long long)n * ldw + k);
        }
        uint4 hi, lo;
        split8(x, hi, lo);
        const int p = n / nt, nn = n - p * nt;
        const int c = (k8 * 8) / GT_KC, kk8 = k8 - c * (GT_KC / 8);
        uint8_t* base = img + ((size_t)(p * nkc + c) * 2) * nt * 64;
        *reinterpret_cast<uint4*>(base + (size_t)kk8 * nt * 16 + nn * 16) = hi;
        *reinterpret_cast<uint4*>(base + (size_t)nt * 64 + (size_t)kk8 * nt * 16 + nn * 16) = lo;
    }
}

// ------------------------------------------------------------------------------------ linear (forward / grad-input)
struct LinTcParams {
    const float* X;          // (rows, K) row-major
    const uint8_t* img;      // weight image
    const float* bias;       // (N) nullable
    const float* resid;      // (rows, N) nullable: added after the activation (may alias Y)
    float* Y;                // (rows, N)
    float* R;                // (rows, N) nullable: relu output
    long long rows;
    int K, N, nt, relu;
    int raw_slots;           // raw fp32 tiles in the TMA ring (<= LT_RAW_MAX)
};

// lanes 2j / 2j + 1 hold adjacent float4 pieces (a: of row A, b: of row B = A + 4); after the exchange the even lane owns the 8
// consecutive floats of row A and the odd lane those of row B (one 16-byte bf16 chunk each)
__device__ __forceinline__ void pair_exchange(const float4 a, const float4 b, bool odd, float* x) {
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

// The activation tile is staged with a PADDED chunk stride (the descriptor's leading byte offset is free): 16-byte chunk c of
// row r at c * A_LBO + r * 16, A_LBO = 2048 + 16, so that the four chunks a quarter warp stores fall into distinct banks.
struct LinTcSmem {
    static constexpr int A_LBO = 128 * 16 + 16;
    static constexpr int A_BYTES = 8320;                           // one hi or lo image of the activation tile (>= 4 * A_LBO)
    // raw fp32 ring | operand stages: A hi | A lo | B hi | B lo (B sized for the launch's column pass) | transpose tiles | barriers
    static constexpr int FIXED = 4 * 32 * 33 * 4 + 256;
    static constexpr int MAX_BYTES = 227 * 1024;
    __host__ __device__ static int stage(int nt) { return 2 * A_BYTES + 2 * nt * 64; }
    __host__ static int raw_slots(int nt) {
        // an EVEN count: a slot is then always consumed by the same converter set, in order -- with an odd count the two sets
        // alternate on a slot and a set could probe a slot's barrier two phases ahead (TMA boxes may land out of order), which
        // the parity test cannot tell from "complete"
        int r = (MAX_BYTES - FIXED - LT_STAGES * stage(nt)) / LT_RAW_BYTES;
        r = r > LT_RAW_MAX ? LT_RAW_MAX : r;
        return r & ~1;
    }
    __host__ __device__ static int total(int nt, int raw_slots) { return raw_slots * LT_RAW_BYTES + LT_STAGES * stage(nt) + FIXED; }
};

__global__ void __launch_bounds__(LT_THREADS, 1) linear_tc_kernel(const LinTcParams P, const __grid_constant__ CUtensorMap tmx) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int stage_bytes = LinTcSmem::stage(P.nt);
    uint8_t* ops = smem + P.raw_slots * LT_RAW_BYTES;
    uint8_t* trans = ops + LT_STAGES * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(trans + 4 * 32 * 33 * 4);
    uint64_t* full = bars;                       // [3] count 5: the 4 converter warps of a set + the expect_tx arrival of the weight image
    uint64_t* empty = bars + LT_STAGES;          // [3] count 1 (MMA commit)
    uint64_t* acc_full = bars + 2 * LT_STAGES;   // [2] count 1
    uint64_t* acc_empty = acc_full + 2;          // [2] count 4 (epilogue warps)
    uint64_t* raw_full = acc_empty + 2;          // [8] count 1 (expect_tx of the loader) + 16 KB
    uint64_t* raw_empty = raw_full + LT_RAW_MAX; // [8] count 4 (converter warps of the set that read the slot)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + LT_RAW_MAX);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkc = P.K / GT_KC;
    const int npass = P.N / P.nt;
    const long long mtiles = (P.rows + 127) / 128;
    const long long ntiles = mtiles * npass;
    const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t items = (uint32_t)(my_tiles * nkc);      // (tile of this CTA, K chunk) in issue order

    if (warp == GT_MMA_WARP) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 0) {
        for (int i = 0; i < LT_STAGES; ++i) { mbar_init(&full[i], 5); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        for (int i = 0; i < LT_RAW_MAX; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 4); }
        fence_barrier_init();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;
    const uint32_t b_img_bytes = (uint32_t)P.nt * 64;           // one hi or lo weight image of a chunk

    if (warp == GT_LOAD_WARP) {
        // ================================================================= TMA loader (one thread)
        // streams the fp32 rows of the tiles into a ring of raw 16 KB tiles (128 rows x 32 floats; rows past the end arrive as
        // zeros): up to eight tiles in flight per SM whatever the converter warps are doing (register-staged loads kept one
        // chunk per warp set in flight and the kernel at 2.5-2.9 TB/s)
        if (lane == 0) {
            tma_prefetch_desc(&tmx);
            for (uint32_t g = 0; g < items; ++g) {
                const int slot = g % P.raw_slots;
                const uint32_t lt = g / nkc;
                const int kc = (int)(g - lt * nkc);
                const long long mt = (blockIdx.x + (long long)lt * gridDim.x) / npass;
                if (g >= (uint32_t)P.raw_slots) mbar_wait(&raw_empty[slot], ((g / P.raw_slots) - 1) & 1);
                mbar_arrive_expect_tx(&raw_full[slot], LT_RAW_BYTES);
                tma_load_3d(smem + slot * LT_RAW_BYTES, &tmx, kc * GT_KC, (int)(mt * 128), 0, &raw_full[slot]);
            }
        }
    } else if (warp < GT_MMA_WARP) {
        // ================================================================= converters
        // Warp set g % GT_PSETS converts item g; warp pw of the set the rows [32 pw, 32 pw + 32) of the tile.  Quarter warp q
        // reads the 128 bytes (the whole K chunk) of row 8 it + q and of row 8 it + 4 + q from the raw tile; a pair exchange
        // gives every lane the 8 consecutive floats of one 16-byte bf16 chunk; the chunk goes to the canonical K-major image
        // with a padded chunk stride (conflict-free quarter-warp stores).
        const int set = warp >> 2, pw = warp & 3;
        const int q = lane >> 3, piece = lane & 7;
        const bool odd = piece & 1;
        for (uint32_t g = set; g < items; g += GT_PSETS) {
            const int slot = g % P.raw_slots;
            const int stage = g % LT_STAGES;
            mbar_wait(&raw_full[slot], (g / P.raw_slots) & 1);
            const float4* raw = reinterpret_cast<const float4*>(smem + slot * LT_RAW_BYTES) + piece;
            float4 xv[8];
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int rA = 32 * pw + 8 * it + q;
                xv[2 * it] = raw[rA * 8];
                xv[2 * it + 1] = raw[(rA + 4) * 8];
            }
            if (g >= LT_STAGES) mbar_wait(&empty[stage], ((g / LT_STAGES) - 1) & 1);
            uint8_t* st = ops + stage * stage_bytes;
            if (pw == 0 && lane == 0) {
                const uint32_t lt = g / nkc;
                const int kc = (int)(g - lt * nkc);
                const int pass = (int)((blockIdx.x + (long long)lt * gridDim.x) % npass);
                mbar_arrive_expect_tx(&full[stage], 2 * b_img_bytes);
                bulk_copy_g2s(st + 2 * LinTcSmem::A_BYTES, P.img + ((size_t)(pass * nkc + kc) * 2) * b_img_bytes, 2 * b_img_bytes, &full[stage]);
            }
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                float x[8];
                pair_exchange(xv[2 * it], xv[2 * it + 1], odd, x);
                uint4 hi, lo;
                split8(x, hi, lo);
                const int row = 32 * pw + 8 * it + q + (odd ? 4 : 0);
                const int off = (piece >> 1) * LinTcSmem::A_LBO + row * 16;
                *reinterpret_cast<uint4*>(st + off) = hi;
                *reinterpret_cast<uint4*>(st + LinTcSmem::A_BYTES + off) = lo;
            }
            gt_warp_arrive(&raw_empty[slot]);              // the raw tile is converted: the loader may refill the slot
            fence_async_smem();
            gt_warp_arrive(&full[stage]);
        }
    } else if (warp == GT_MMA_WARP) {
        // ================================================================= MMA issue (warp-uniform, elected lane)
        const uint32_t idesc = idesc_bf16(128, P.nt, 0, 0);
        uint32_t gt = 0, tt = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++tt) {
            const uint32_t buf = tt & 1;
            if (tt >= 2) mbar_wait(&acc_empty[buf], ((tt >> 1) - 1) & 1);
            fence_after_sync();
            const uint32_t acc = tmem_addr(tb, 0, 256 * buf);
            for (int kc = 0; kc < nkc; ++kc, ++gt) {
                const int stage = gt % LT_STAGES;
                mbar_wait(&full[stage], (gt / LT_STAGES) & 1);
                fence_after_sync();
                if (elect_one()) {
                    const uint32_t a_hi = smem_u32(ops + stage * stage_bytes);
                    const uint32_t a_lo = a_hi + LinTcSmem::A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * LinTcSmem::A_BYTES;
                    const uint32_t b_lo = b_hi + b_img_bytes;
                    const uint32_t b_lbo = (uint32_t)P.nt * 16;
#pragma unroll
                    for (int ks = 0; ks < GT_KC / 16; ++ks) {
                        const uint64_t dah = smem_desc(a_hi + ks * 2 * LinTcSmem::A_LBO, LinTcSmem::A_LBO, 128);
                        const uint64_t dal = smem_desc(a_lo + ks * 2 * LinTcSmem::A_LBO, LinTcSmem::A_LBO, 128);
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
    } else if (warp < GT_LOAD_WARP) {
        // ================================================================= epilogue
        // TMEM rows arrive one per thread; every 32 x 32 block is transposed through a padded per-warp smem tile so that
        // global traffic is coalesced: lane = (row % 4, float4 of columns) -> 4 rows x 128 contiguous bytes per instruction.
        const int quad = warp & 3;                     // TMEM lane quadrant this warp may read
        float* T = reinterpret_cast<float*>(trans) + quad * (32 * 33);
        const int rr = lane >> 3, cq = lane & 7;
        uint32_t tt = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++tt) {
            const uint32_t buf = tt & 1;
            const long long mt = t / npass;
            const int pass = (int)(t - mt * npass);
            const long long r0 = mt * 128 + 32 * quad;
            mbar_wait(&acc_full[buf], (tt >> 1) & 1);
            fence_after_sync();
            for (int c0 = 0; c0 < P.nt; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_addr(tb, 32 * quad, 256 * buf + c0), v);
                tmem_ld_wait32(v);
#pragma unroll
                for (int j = 0; j < 32; ++j) T[lane * 33 + j] = __uint_as_float(v[j]);
                __syncwarp();
                const int n = pass * P.nt + c0 + 4 * cq;
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (P.bias) bv = __ldg(reinterpret_cast<const float4*>(P.bias + n));
                float4 rv8[8];                    // the block's residual rows: all eight loads in flight at once
                if (P.resid) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const long long r = r0 + 4 * i + rr;
                        rv8[i] = r < P.rows ? *reinterpret_cast<const float4*>(P.resid + r * P.N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int lr = 4 * i + rr;
                    const long long r = r0 + lr;
                    float4 o = make_float4(T[lr * 33 + 4 * cq] + bv.x, T[lr * 33 + 4 * cq + 1] + bv.y, T[lr * 33 + 4 * cq + 2] + bv.z,
                                           T[lr * 33 + 4 * cq + 3] + bv.w);
                    if (P.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    if (r < P.rows) {
                        if (P.R) *reinterpret_cast<float4*>(P.R + r * P.N + n) = o;
                        if (P.resid) { o.x += rv8[i].x; o.y += rv8[i].y; o.z += rv8[i].z; o.w += rv8[i].w; }
                        *reinterpret_cast<float4*>(P.Y + r * P.N + n) = o;
                    }
                }
                __syncwarp();
            }
            fence_before_sync();
            gt_warp_arrive(&acc_empty[buf]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == GT_MMA_WARP) tmem_dealloc(tb, 512);
}

// function attributes belong to the current device's context: configure once per device ordinal (idempotent; see tc_configure)
static int gemm_tc_configure();

static int g_gemm_tc = 1;       // 0: every GEMM stays on the CUDA-core kernels (PCA_GEMM_TC=0 / pca_debug_set_gemm_tc)
void set_gemm_tc(int on) { g_gemm_tc = on ? 1 : 0; }
static bool gemm_tc_on() {
    static int env = -1;
    if (env < 0) { const char* v = getenv("PCA_GEMM_TC"); env = (v && v[0] == '0') ? 0 : 1; }
    return env && g_gemm_tc;
}

static int pick_nt(int N) {
    if (N <= 256) return N;
    for (int nt = 256; nt >= 32; nt -= 32)
        if (N % nt == 0) return nt;
    return 0;
}

size_t gemm_tc_image_bytes(int N, int K) { return (size_t)4 * N * K; }

bool linear_tc_eligible(long long rows, int K, int N) {
    return gemm_tc_on() && rows >= 512 && K % GT_KC == 0 && K >= GT_KC && N % 32 == 0 && N >= 32 && pick_nt(N) >= 32;
}

// Y (rows, N) = act(X (rows, K) B^T + bias) [+ resid];  B(n, k) = trans_w ? W[k * N + n] : W[n * K + k]
int launch_linear_tc(const float* X, const float* W, int trans_w, const float* bias, const float* resid, float* Y, float* R,
                     long long rows, int K, int N, int relu, void* img, size_t img_bytes, cudaStream_t st) {
    if (rows == 0) return 0;
    if (!linear_tc_eligible(rows, K, N)) return fail(PCA_EUNSUPPORTED, "linear_tc: shape (%lld, %d, %d) not eligible", rows, K, N);
    if (!img || img_bytes < gemm_tc_image_bytes(N, K)) return fail(PCA_EWORKSPACE, "linear_tc: weight image buffer too small");
    PCA_TRY(gemm_tc_configure());
    int nt = pick_nt(N);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // few row tiles (the 16-row-per-cloud tensors of the inducing points at small batches): narrower column passes spread the
    // work over more CTAs (every CTA re-stages its 128-row activation tile, which then comes from L2)
    while (((rows + 127) / 128) * (N / nt) * 2 <= sms && nt >= 64 && (nt / 2) % 32 == 0 && N % (nt / 2) == 0) nt /= 2;
    {
        const long long total = (long long)N * (K / 8);
        weight_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(W, N, K, trans_w ? N : K, trans_w, nt, (uint8_t*)img);
        PCA_CHECK_LAUNCH("weight_image_kernel");
    }
    LinTcParams p{X, (const uint8_t*)img, bias, resid, Y, R, rows, K, N, nt, relu, LinTcSmem::raw_slots(nt)};
    CUtensorMap tmx;                  // X as (K, rows): a box of 32 floats x 128 rows is one raw tile
    PCA_TRY(make_tmap_3d_f32(&tmx, X, (unsigned long long)K, (unsigned long long)rows, 1ull, (unsigned long long)K * 4,
                             (unsigned long long)rows * K * 4, GT_KC, 128));
    const long long ntiles = ((rows + 127) / 128) * (N / nt);
    const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
    {
        LaunchTimer lt("linear_tc_kernel", st, 2.0 * rows * K * N, 4.0 * rows * (K + N));
        linear_tc_kernel<<<grid, LT_THREADS, LinTcSmem::total(nt, p.raw_slots), st>>>(p, tmx);
    }
    PCA_CHECK_LAUNCH("linear_tc_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------ weight gradient
// dW (M, N) += A (rows, M)^T B (rows, N) over the rows [r0, r1) of this CTA.  Both operands are MN-major: a 16-byte unit
// is 8 consecutive features of one row; units of one 8-row group are contiguous (128 B), 8-feature groups 'SBO' apart:
//   byte(mn, k) = (mn / 8) * (KC * 16) + k * 16 + (mn % 8) * 2,   k < KC = 32 rows per stage   (LBO = 128, SBO = KC * 16).
// 4 producer warps: thread = (row k of the chunk, quarter of the features).  M = 128 per CTA (blockIdx.y), N <= 256.
struct GwTcParams {
    const float* A;          // (rows, lda): columns [m0, m0 + 128) of dY
    const float* B;          // (rows, N)
    float* dW;               // (Mtot, N)
    long long rows, rchunk;
    int lda, Mtot, N;
    int mbox, raw_slots;     // columns of dY per raw chunk (min(128, Mtot)); raw chunks in the TMA ring
    int nsplit, mtiles;      // work items = nsplit row ranges x mtiles 128-row blocks of dW
};
// 8-feature group g of chunk row k at g * G_SBO + k * 16 with G_SBO = 512 + 16 (padded: the four groups a quarter warp stores
// fall into distinct banks)
struct GwTcSmem {
    static constexpr int G_SBO = GT_KC * 16 + 16;
    static constexpr int A_BYTES = 16 * G_SBO;
    static constexpr int B_BYTES = 32 * G_SBO;
    static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;
    // raw fp32 ring (32-row chunks of both operands) | 2 operand stages | transpose tiles | barriers
    static constexpr int OP_STAGES = 2;
    static constexpr int RAW_MAX = 4;
    static constexpr int FIXED = 4 * 32 * 33 * 4 + 256;
    static constexpr int MAX_BYTES = 227 * 1024;
    __host__ __device__ static int raw_bytes(int mbox, int N) { return GT_KC * (mbox + N) * 4; }
    __host__ static int raw_slots(int mbox, int N) {
        int r = (MAX_BYTES - FIXED - OP_STAGES * STAGE) / raw_bytes(mbox, N);      // even, as in the linear kernel
        r = r > RAW_MAX ? RAW_MAX : r;
        return r & ~1;
    }
    __host__ __device__ static int total(int mbox, int N, int slots) { return slots * raw_bytes(mbox, N) + OP_STAGES * STAGE + FIXED; }
};

__global__ void __launch_bounds__(LT_THREADS, 1) grad_weight_tc_kernel(const GwTcParams P, const __grid_constant__ CUtensorMap tma,
                                                                       const __grid_constant__ CUtensorMap tmb) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int raw_bytes = GwTcSmem::raw_bytes(P.mbox, P.N);
    uint8_t* ops = smem + P.raw_slots * raw_bytes;
    uint8_t* trans = ops + GwTcSmem::OP_STAGES * GwTcSmem::STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(trans + 4 * 32 * 33 * 4);
    uint64_t* full = bars;                                   // [2] count 4 (the converter warps of a set)
    uint64_t* empty = bars + GwTcSmem::OP_STAGES;            // [2] count 1
    uint64_t* acc_full = bars + 2 * GwTcSmem::OP_STAGES;     // [2] count 1
    uint64_t* acc_empty = acc_full + 2;                      // [2] count 4 (epilogue warps)
    uint64_t* raw_full = acc_empty + 2;                      // [4] count 1 (expect_tx of the loader) + the chunk's bytes
    uint64_t* raw_empty = raw_full + GwTcSmem::RAW_MAX;      // [4] count 4
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + GwTcSmem::RAW_MAX);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Persistent: work item = (row range of rchunk rows -- a multiple of the chunk size --, 128-row block of dW); this CTA walks
    // the items blockIdx.x, + gridDim.x, ...; ring slots and operand stages run across item boundaries (global chunk counter)
    // and two accumulators let the reductions of an item overlap the chunks of the next one.
    const int n_items = P.nsplit * P.mtiles;
    auto item_rows = [&](int it, int& m0, long long& r0) -> int {      // chunks of the item (items of one row range are adjacent)
        const int sp = it / P.mtiles;
        m0 = (it - sp * P.mtiles) * 128;
        r0 = (long long)sp * P.rchunk;
        const long long r1 = (r0 + P.rchunk < P.rows) ? r0 + P.rchunk : P.rows;
        return (int)((r1 - r0 + GT_KC - 1) / GT_KC);
    };

    if (warp == GT_MMA_WARP) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 0) {
        for (int i = 0; i < GwTcSmem::OP_STAGES; ++i) { mbar_init(&full[i], 4); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        for (int i = 0; i < GwTcSmem::RAW_MAX; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 4); }
        fence_barrier_init();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;

    if (warp == GT_LOAD_WARP) {
        // TMA loader (one thread): chunk c = rows [r0 + 32 c, + 32) of dY (columns [m0, m0 + mbox)) and of X (N columns); rows /
        // columns outside the tensors arrive as zeros
        if (lane == 0) {
            tma_prefetch_desc(&tma);
            tma_prefetch_desc(&tmb);
            int gc = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                int m0;
                long long r0;
                const int nchunks = item_rows(it, m0, r0);
                for (int c = 0; c < nchunks; ++c, ++gc) {
                    const int slot = gc % P.raw_slots;
                    if (gc >= P.raw_slots) mbar_wait(&raw_empty[slot], ((gc / P.raw_slots) - 1) & 1);
                    mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)raw_bytes);
                    uint8_t* dst = smem + slot * raw_bytes;
                    const int row = (int)(r0 + (long long)c * GT_KC);
                    tma_load_3d(dst, &tma, m0, row, 0, &raw_full[slot]);
                    tma_load_3d(dst + GT_KC * P.mbox * 4, &tmb, 0, row, 0, &raw_full[slot]);
                }
            }
        }
    } else if (warp < GT_MMA_WARP) {
        // Converters: warp set s handles the chunks s, s + 2, ...; warp pw of the set the rows [8 pw, 8 pw + 8) of the chunk for
        // every 32-feature block.  Quarter warp q reads the 128 bytes of row 8 pw + q, then of row 8 pw + 4 + q, from the raw
        // chunk; a pair exchange gives every lane one 16-byte unit (8 consecutive features of one row), stored MN-major with a
        // padded group stride (conflict-free quarter-warp stores).
        const int set = warp >> 2, pw = warp & 3;
        const int q = lane >> 3, piece = lane & 7;
        const bool odd = piece & 1;
        const int a_fblocks = P.mbox / 32, b_fblocks = P.N / 32;
        int total_chunks = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            int m0;
            long long r0;
            total_chunks += item_rows(it, m0, r0);
        }
        for (int c = set; c < total_chunks; c += GT_PSETS) {  // global chunk index: the conversion does not depend on the item
            const int slot = c % P.raw_slots;
            const int stage = c % GwTcSmem::OP_STAGES;
            mbar_wait(&raw_full[slot], (c / P.raw_slots) & 1);
            const float* rawA = reinterpret_cast<const float*>(smem + slot * raw_bytes) + 4 * piece;
            const float* rawB = rawA + GT_KC * P.mbox;
            const int kA = 8 * pw + q;
            float4 va[8], vb[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool on = i < a_fblocks;
                va[2 * i] = on ? *reinterpret_cast<const float4*>(rawA + kA * P.mbox + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                va[2 * i + 1] = on ? *reinterpret_cast<const float4*>(rawA + (kA + 4) * P.mbox + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool on = i < b_fblocks;
                vb[2 * i] = on ? *reinterpret_cast<const float4*>(rawB + kA * P.N + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                vb[2 * i + 1] = on ? *reinterpret_cast<const float4*>(rawB + (kA + 4) * P.N + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (c >= GwTcSmem::OP_STAGES) mbar_wait(&empty[stage], ((c / GwTcSmem::OP_STAGES) - 1) & 1);
            uint8_t* st = ops + stage * GwTcSmem::STAGE;
            const int k = kA + (odd ? 4 : 0);                         // row of the chunk this lane stores
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float x[8];
                pair_exchange(va[2 * i], va[2 * i + 1], odd, x);
                uint4 hi, lo;
                split8(x, hi, lo);
                const int off = (4 * i + (piece >> 1)) * GwTcSmem::G_SBO + k * 16;
                *reinterpret_cast<uint4*>(st + off) = hi;
                *reinterpret_cast<uint4*>(st + GwTcSmem::A_BYTES + off) = lo;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < b_fblocks) {                                  // warp-uniform
                    float x[8];
                    pair_exchange(vb[2 * i], vb[2 * i + 1], odd, x);
                    uint4 hi, lo;
                    split8(x, hi, lo);
                    const int off = (4 * i + (piece >> 1)) * GwTcSmem::G_SBO + k * 16;
                    *reinterpret_cast<uint4*>(st + 2 * GwTcSmem::A_BYTES + off) = hi;
                    *reinterpret_cast<uint4*>(st + 2 * GwTcSmem::A_BYTES + GwTcSmem::B_BYTES + off) = lo;
                }
            }
            gt_warp_arrive(&raw_empty[slot]);
            fence_async_smem();
            gt_warp_arrive(&full[stage]);
        }
    } else if (warp == GT_MMA_WARP) {
        const uint32_t idesc = idesc_bf16(128, P.N, 1, 1);
        int gc = 0, k = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
            int m0;
            long long r0;
            const int nchunks = item_rows(it, m0, r0);
            const uint32_t buf = k & 1;
            if (k >= 2) mbar_wait(&acc_empty[buf], ((k >> 1) - 1) & 1);
            fence_after_sync();
            const uint32_t acc = tmem_addr(tb, 0, 256 * buf);
            for (int c = 0; c < nchunks; ++c, ++gc) {
                const int stage = gc % GwTcSmem::OP_STAGES;
                mbar_wait(&full[stage], (gc / GwTcSmem::OP_STAGES) & 1);
                fence_after_sync();
                if (elect_one()) {
                    const uint32_t a_hi = smem_u32(ops + stage * GwTcSmem::STAGE);
                    const uint32_t a_lo = a_hi + GwTcSmem::A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * GwTcSmem::A_BYTES;
                    const uint32_t b_lo = b_hi + GwTcSmem::B_BYTES;
#pragma unroll
                    for (int ks = 0; ks < GT_KC / 16; ++ks) {
                        // MN-major: a K step of 16 rows = two 8-row groups = 256 bytes further
                        const uint64_t dah = smem_desc(a_hi + ks * 256, 128, GwTcSmem::G_SBO), dal = smem_desc(a_lo + ks * 256, 128, GwTcSmem::G_SBO);
                        const uint64_t dbh = smem_desc(b_hi + ks * 256, 128, GwTcSmem::G_SBO), dbl = smem_desc(b_lo + ks * 256, 128, GwTcSmem::G_SBO);
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
    } else if (warp < GT_LOAD_WARP) {
        const int quad = warp & 3;
        float* T = reinterpret_cast<float*>(trans) + quad * (32 * 33);
        int k = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
            int m0;
            long long r0;
            item_rows(it, m0, r0);
            const uint32_t buf = k & 1;
            mbar_wait(&acc_full[buf], (k >> 1) & 1);
            fence_after_sync();
            for (int c0 = 0; c0 < P.N; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_addr(tb, 32 * quad, 256 * buf + c0), v);
                tmem_ld_wait32(v);
#pragma unroll
                for (int j = 0; j < 32; ++j) T[lane * 33 + j] = __uint_as_float(v[j]);
                __syncwarp();
                // lane = column: one 128-byte reduction per row and instruction
#pragma unroll 4
                for (int lr = 0; lr < 32; ++lr) {
                    const int m = m0 + 32 * quad + lr;
                    if (m < P.Mtot) atomicAdd(P.dW + (long long)m * P.N + c0 + lane, T[lr * 33 + lane]);
                }
                __syncwarp();
            }
            fence_before_sync();
            gt_warp_arrive(&acc_empty[buf]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == GT_MMA_WARP) tmem_dealloc(tb, 512);
}

static int gemm_tc_configure() {
    static std::atomic<unsigned long long> done_mask{0};
    int dev = 0;
    PCA_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && ((done_mask.load(std::memory_order_acquire) >> dev) & 1ull)) return 0;
    PCA_CHECK_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LinTcSmem::MAX_BYTES));
    PCA_CHECK_CUDA(cudaFuncSetAttribute(grad_weight_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GwTcSmem::MAX_BYTES));
    if (dev < 64) done_mask.fetch_or(1ull << dev, std::memory_order_release);
    return 0;
}

bool grad_weight_tc_eligible(long long rows, int M, int N) {
    return gemm_tc_on() && rows >= 2048 && M % 32 == 0 && M >= 64 && N % 32 == 0 && N >= 32 && N <= 256;
}

// dW (M, N) += dY (rows, M)^T X (rows, N)
int launch_grad_weight_tc(const float* dY, const float* X, float* dW, long long rows, int M, int N, cudaStream_t st) {
    if (rows == 0) return 0;
    if (!grad_weight_tc_eligible(rows, M, N)) return fail(PCA_EUNSUPPORTED, "grad_weight_tc: shape (%lld, %d, %d) not eligible", rows, M, N);
    PCA_TRY(gemm_tc_configure());
    const int mtiles = (M + 127) / 128;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // persistent CTAs (one per SM) over (row range, 128-row block of dW) items: the number of row ranges that fills the last
    // round best while an item keeps >= 256 rows (every item ends in 128 x N atomic reductions)
    const long long max_split = (rows + 8 * GT_KC - 1) / (8 * GT_KC);
    long long nsplit = 1;
    double best = 0.0;
    for (long long c = 1; c <= max_split && c * mtiles <= 8LL * sms; ++c) {
        const long long items = c * mtiles;
        const double fill = (double)items / (double)(((items + sms - 1) / sms) * sms);
        if (fill > best + 0.02) { best = fill; nsplit = c; }
    }
    long long rchunk = (rows + nsplit - 1) / nsplit;
    rchunk = (rchunk + GT_KC - 1) / GT_KC * GT_KC;
    nsplit = (rows + rchunk - 1) / rchunk;
    const int mbox = M < 128 ? M : 128;
    GwTcParams p{dY, X, dW, rows, rchunk, M, M, N, mbox, GwTcSmem::raw_slots(mbox, N), (int)nsplit, mtiles};
    const long long n_items = nsplit * mtiles;
    const unsigned grid = (unsigned)(n_items < sms ? n_items : sms);
    CUtensorMap tma, tmb;             // dY as (M, rows), X as (N, rows): one box = the 32 rows of a chunk
    PCA_TRY(make_tmap_3d_f32(&tma, dY, (unsigned long long)M, (unsigned long long)rows, 1ull, (unsigned long long)M * 4,
                             (unsigned long long)rows * M * 4, (unsigned)mbox, GT_KC));
    PCA_TRY(make_tmap_3d_f32(&tmb, X, (unsigned long long)N, (unsigned long long)rows, 1ull, (unsigned long long)N * 4,
                             (unsigned long long)rows * N * 4, (unsigned)N, GT_KC));
    {
        LaunchTimer lt("grad_weight_tc_kernel", st, 2.0 * rows * M * N, 4.0 * rows * (M + N));
        grad_weight_tc_kernel<<<grid, LT_THREADS, GwTcSmem::total(mbox, N, p.raw_slots), st>>>(p, tma, tmb);
    }
    PCA_CHECK_LAUNCH("grad_weight_tc_kernel");
    return 0;
}

}  // namespace pca
