// Training path of the set encoder: fp32 forward that keeps the activations the
// backward pass needs, the backward pass of ST / SetTransformer (modules.py:6-63 differentiated by hand), dropout,
// cross-entropy and a fused Adam step on the flat parameter blob.  CUDA-core fp32 kernels, every MAB shape.
//
//   MAB forward (modules.py:18-33, ln=False)      backward (given dOut)
//     Qp = Qin Wq^T + bq                            dZ  = dOut o [R > 0]
//     KV = Kin [Wk;Wv]^T + [bk;bv]                  dWo += dZ^T O,  dbo += colsum dZ
//     O  = Qp + softmax(Qp_h Kp_h^T / sqrt(D)) Vp_h  dO  = dOut + dZ Wo
//     R  = relu(O Wo^T + bo)                        delta_h = rowsum_h(dO o (O - Qp))
//     Out = O + R                                   dS = P o (dO_h Vp_h^T - delta_h) / sqrt(D)   (P recomputed from lse)
//                                                   dQp = dO + dS Kp_h,  dKp_h = dS^T Qp_h,  dVp_h = P^T dO_h
//                                                   dWq += dQp^T Qin, dQin = dQp Wq, dWkv += dKV^T Kin, dKin = dKV Wkv
#include "common.cuh"
#include <math.h>

namespace pca {

// ------------------------------------------------------------------------------------ generic fp32 GEMM
// C (M, N) [+]= opA (M, K) opB (K, N) [+ Cadd];  opA(m,k) = TA ? A[k*lda+m] : A[m*lda+k];  opB(k,n) = TB ? B[n*ldb+k] : B[k*ldb+n].
// blockIdx.z splits K (partial products are added atomically: gradient accumulation over rows).
constexpr int GBM = 128, GBN = 64, GBK = 16, GTHREADS = 256;

template <bool TA, bool TB>
__global__ void __launch_bounds__(GTHREADS)
gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C, const float* __restrict__ Cadd,
                long long M, int N, long long K, long long lda, long long ldb, long long ldc, long long kchunk, int atomic) {
    __shared__ __align__(16) float As[GBK][GBM + 4];
    __shared__ __align__(16) float Bs[GBK][GBN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * GBM;
    const int n0 = blockIdx.y * GBN;
    const long long kbeg = (long long)blockIdx.z * kchunk;
    const long long kend = (kbeg + kchunk < K) ? kbeg + kchunk : K;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long long k0 = kbeg; k0 < kend; k0 += GBK) {
        if (TA) {
            const int lm = tid & 127, lk = tid >> 7;
            const long long m = m0 + lm;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long k = k0 + lk + 2 * j;
                As[lk + 2 * j][lm] = (m < M && k < kend) ? __ldg(A + k * lda + m) : 0.f;
            }
        } else {
            const int lk = tid & 15, lr = tid >> 4;
            const long long k = k0 + lk;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long m = m0 + lr + 16 * j;
                As[lk][lr + 16 * j] = (m < M && k < kend) ? __ldg(A + m * lda + k) : 0.f;
            }
        }
        if (TB) {
            const int lk = tid & 15, lr = tid >> 4;
            const long long k = k0 + lk;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + lr + 16 * j;
                Bs[lk][lr + 16 * j] = (n < N && k < kend) ? __ldg(Bm + (long long)n * ldb + k) : 0.f;
            }
        } else {
            const int ln = tid & 63, lk = tid >> 6;
            const int n = n0 + ln;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long k = k0 + lk + 4 * j;
                Bs[lk + 4 * j][ln] = (n < N && k < kend) ? __ldg(Bm + k * ldb + n) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (Cadd) v += Cadd[m * ldc + n];
            if (atomic) atomicAdd(C + m * ldc + n, v);
            else C[m * ldc + n] = v;
        }
    }
}

// dX = dY W [+ add] for up to 32 rows (gradients of the inducing points / seeds / small heads; the tile kernel would run as
// 1 x din/64 blocks: measured 40 us for 16 x 256 x 256).  A block owns eight input columns: dY (row stride dout + 1) and its
// 8-column slice of W are staged in shared memory by one round of loads; lane = row, warp = column accumulate from there.
__global__ void __launch_bounds__(256)
grad_input_skinny_kernel(const float* __restrict__ dY, const float* __restrict__ W, float* dX, const float* add, int rows,
                         int din, int dout) {
    extern __shared__ float skinny_s[];
    float* Gs = skinny_s;                          // rows x (dout + 1)
    float* Ws = skinny_s + rows * (dout + 1);      // dout x 8
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i0 = blockIdx.x * 8, i = i0 + w;
    for (int t = threadIdx.x; t < rows * dout; t += 256) {
        const int r = t / dout, o = t - r * dout;
        Gs[r * (dout + 1) + o] = __ldg(dY + t);
    }
    for (int t = threadIdx.x; t < dout * 8; t += 256) {
        const int o = t >> 3, ii = t & 7;
        Ws[t] = (i0 + ii < din) ? __ldg(W + (long long)o * din + i0 + ii) : 0.f;
    }
    __syncthreads();
    if (i >= din || lane >= rows) return;
    const float* g = Gs + lane * (dout + 1);
    float a0 = 0.f, a1 = 0.f;
    int o = 0;
    for (; o + 1 < dout; o += 2) { a0 = fmaf(g[o], Ws[o * 8 + w], a0); a1 = fmaf(g[o + 1], Ws[(o + 1) * 8 + w], a1); }
    if (o < dout) a0 = fmaf(g[o], Ws[o * 8 + w], a0);
    float v = a0 + a1;
    if (add) v += add[(long long)lane * din + i];
    dX[(long long)lane * din + i] = v;
}

// dX (rows, din) = dY (rows, dout) W (dout, din) [+ add (rows, din)]; add may alias dX (accumulation)
static int launch_grad_input(const float* dY, const float* W, float* dX, const float* add, long long rows, int din, int dout,
                             cudaStream_t st, void* img = nullptr, size_t img_bytes = 0) {
    if (rows == 0) return 0;
    if (img && img_bytes >= gemm_tc_image_bytes(din, dout) && linear_tc_eligible(rows, dout, din))      // K = dout, N = din
        return launch_linear_tc(dY, W, 1, nullptr, add, dX, nullptr, rows, dout, din, 0, img, img_bytes, st);
    const size_t skinny_smem = ((size_t)rows * (dout + 1) + 8 * (size_t)dout) * sizeof(float);
    if (rows <= 32 && skinny_smem <= 48 * 1024) {
        LaunchTimer lt("grad_input_skinny_kernel", st, 2.0 * rows * din * dout, 4.0 * ((double)rows * (din + dout) + (double)din * dout));
        grad_input_skinny_kernel<<<(unsigned)((din + 7) / 8), 256, skinny_smem, st>>>(dY, W, dX, add, (int)rows, din, dout);
        PCA_CHECK_LAUNCH("grad_input_skinny_kernel");
        return 0;
    }
    dim3 grid((unsigned)((rows + GBM - 1) / GBM), (din + GBN - 1) / GBN, 1);
    {
        LaunchTimer lt("gemm_f32_kernel", st, 2.0 * rows * din * dout, 4.0 * rows * (din + dout));
        gemm_f32_kernel<false, false><<<grid, GTHREADS, 0, st>>>(dY, W, dX, add, rows, din, dout, dout, din, din, dout, 0);
    }
    PCA_CHECK_LAUNCH("gemm_f32_kernel<grad_input>");
    return 0;
}

// dW (dout, din) += dY (rows, dout)^T X (rows, din): the row range is split across blockIdx.z, partial sums added atomically
static int launch_grad_weight(const float* dY, const float* X, float* dW, long long rows, int din, int dout, cudaStream_t st) {
    if (rows == 0) return 0;
    if (grad_weight_tc_eligible(rows, dout, din)) return launch_grad_weight_tc(dY, X, dW, rows, dout, din, st);
    const int tiles = ((dout + GBM - 1) / GBM) * ((din + GBN - 1) / GBN);
    long long nsplit = (148 * 4 + tiles - 1) / tiles;
    const long long max_split = (rows + 4 * GBK - 1) / (4 * GBK);
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > 65535) nsplit = 65535;
    long long kchunk = (rows + nsplit - 1) / nsplit;
    kchunk = (kchunk + GBK - 1) / GBK * GBK;
    nsplit = (rows + kchunk - 1) / kchunk;
    dim3 grid((dout + GBM - 1) / GBM, (din + GBN - 1) / GBN, (unsigned)nsplit);
    {
        LaunchTimer lt("gemm_f32_kernel", st, 2.0 * rows * din * dout, 4.0 * rows * (din + dout));
        gemm_f32_kernel<true, false><<<grid, GTHREADS, 0, st>>>(dY, X, dW, nullptr, dout, din, rows, dout, din, din, kchunk, 1);
    }
    PCA_CHECK_LAUNCH("gemm_f32_kernel<grad_weight>");
    return 0;
}

// out (cols) += sum over rows of A (rows, cols)
__global__ void colsum_kernel(const float* __restrict__ A, long long rows, int cols, long long rchunk, float* __restrict__ out) {
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rchunk;
    const long long r1 = (r0 + rchunk < rows) ? r0 + rchunk : rows;
    float v = 0.f;
    if (c < cols)
        for (long long r = r0 + threadIdx.y; r < r1; r += 8) v += __ldg(A + r * cols + c);
    red[threadIdx.y][threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        for (int i = 1; i < 8; ++i) v += red[i][threadIdx.x];
        atomicAdd(out + c, v);
    }
}
static int launch_colsum(const float* A, long long rows, int cols, float* out, cudaStream_t st) {
    if (rows == 0 || cols == 0) return 0;
    const int ctiles = (cols + 31) / 32;
    long long nsplit = (148 * 8 + ctiles - 1) / ctiles;
    const long long max_split = (rows + 63) / 64;
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit > 65535) nsplit = 65535;
    if (nsplit < 1) nsplit = 1;
    const long long rchunk = (rows + nsplit - 1) / nsplit;
    nsplit = (rows + rchunk - 1) / rchunk;
    dim3 grid(ctiles, (unsigned)nsplit), block(32, 8);
    {
        LaunchTimer lt("colsum_kernel", st, 0.0, 4.0 * rows * cols);
        colsum_kernel<<<grid, block, 0, st>>>(A, rows, cols, rchunk, out);
    }
    PCA_CHECK_LAUNCH("colsum_kernel");
    return 0;
}

// First layers (d_in = 2 / 3 point coordinates): bias and weight gradient of a Linear with din <= 4 in ONE pass over dY:
// db[c] += sum_r dY[r, c];  dW[c, j] += sum_r dY[r, c] X[r, j].  (A 128 x 64 GEMM tile would be 95 % padding here.)
template <int DIN>
__global__ void colsum_x_kernel(const float* __restrict__ A, const float* __restrict__ X, long long rows, int cols, long long rchunk,
                                float* __restrict__ db, float* __restrict__ dW) {
    __shared__ float red[8][DIN + 1][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rchunk;
    const long long r1 = (r0 + rchunk < rows) ? r0 + rchunk : rows;
    float acc[DIN + 1];
#pragma unroll
    for (int j = 0; j <= DIN; ++j) acc[j] = 0.f;
    if (c < cols) {
        for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
            const float a = __ldg(A + r * cols + c);
            acc[DIN] += a;
#pragma unroll
            for (int j = 0; j < DIN; ++j) acc[j] = fmaf(a, __ldg(X + r * DIN + j), acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j <= DIN; ++j) red[threadIdx.y][j][threadIdx.x] = acc[j];
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
#pragma unroll
        for (int j = 0; j <= DIN; ++j) {
            float v = acc[j];
            for (int i = 1; i < 8; ++i) v += red[i][j][threadIdx.x];
            if (j == DIN) atomicAdd(db + c, v);
            else atomicAdd(dW + (long long)c * DIN + j, v);
        }
    }
}
static int launch_colsum_x(const float* dY, const float* X, long long rows, int cols, int din, float* dW, float* db, cudaStream_t st) {
    if (rows == 0 || cols == 0) return 0;
    const int ctiles = (cols + 31) / 32;
    long long nsplit = (148 * 8 + ctiles - 1) / ctiles;
    const long long max_split = (rows + 63) / 64;
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit > 65535) nsplit = 65535;
    if (nsplit < 1) nsplit = 1;
    const long long rchunk = (rows + nsplit - 1) / nsplit;
    nsplit = (rows + rchunk - 1) / rchunk;
    dim3 grid(ctiles, (unsigned)nsplit), block(32, 8);
    {
        LaunchTimer lt("colsum_x_kernel", st, 2.0 * rows * cols * din, 4.0 * rows * (cols + din));
        switch (din) {
            case 1: colsum_x_kernel<1><<<grid, block, 0, st>>>(dY, X, rows, cols, rchunk, db, dW); break;
            case 2: colsum_x_kernel<2><<<grid, block, 0, st>>>(dY, X, rows, cols, rchunk, db, dW); break;
            case 3: colsum_x_kernel<3><<<grid, block, 0, st>>>(dY, X, rows, cols, rchunk, db, dW); break;
            case 4: colsum_x_kernel<4><<<grid, block, 0, st>>>(dY, X, rows, cols, rchunk, db, dW); break;
            default: return fail(PCA_EINVAL, "colsum_x: din %d not in 1..4", din);
        }
    }
    PCA_CHECK_LAUNCH("colsum_x_kernel");
    return 0;
}
// weight + bias gradient of a Linear: one fused pass for the skinny first layers, GEMM + column sum otherwise
static int launch_grad_weight_bias(const float* dY, const float* X, float* dW, float* db, long long rows, int din, int dout,
                                   cudaStream_t st) {
    if (din <= 4 && rows >= 256) return launch_colsum_x(dY, X, rows, dout, din, dW, db, st);
    PCA_TRY(launch_grad_weight(dY, X, dW, rows, din, dout, st));
    return launch_colsum(dY, rows, dout, db, st);
}

// ------------------------------------------------------------------------------------ elementwise pieces
__device__ __forceinline__ void red_add4_fwd(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__global__ void add_kernel(float* __restrict__ a, const float* __restrict__ b, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += b[i];
}
__global__ void relu_bwd_kernel(const float* __restrict__ dOut, const float* __restrict__ R, float* __restrict__ dZ, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dZ[i] = R[i] > 0.f ? dOut[i] : 0.f;
}
// The same with the bias gradient folded in: dZ = [R > 0] dOut and db[c] += sum_r dZ[r, c] in ONE pass (the separate column-sum
// kernel read dZ a second time).  Block = 256-row slab; thread = (float4 column group, row lane); D % 4 == 0, 256 % (D / 4) == 0.
__global__ void __launch_bounds__(256) relu_bwd_colsum_kernel(const float* __restrict__ dOut, const float* __restrict__ R, float* __restrict__ dZ,
                                                              long long rows, int D, float* __restrict__ db) {
    __shared__ float4 red[256];
    const int cq = D >> 2;
    const int c4 = threadIdx.x % cq, rsub = threadIdx.x / cq, rstep = 256 / cq;
    const long long r_end = min(rows, (long long)(blockIdx.x + 1) * 256);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long r = (long long)blockIdx.x * 256 + rsub; r < r_end; r += rstep) {
        const float4 g = *reinterpret_cast<const float4*>(dOut + r * D + 4 * c4);
        const float4 a = *reinterpret_cast<const float4*>(R + r * D + 4 * c4);
        const float4 z = make_float4(a.x > 0.f ? g.x : 0.f, a.y > 0.f ? g.y : 0.f, a.z > 0.f ? g.z : 0.f, a.w > 0.f ? g.w : 0.f);
        *reinterpret_cast<float4*>(dZ + r * D + 4 * c4) = z;
        acc.x += z.x; acc.y += z.y; acc.z += z.z; acc.w += z.w;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (rsub == 0) {
        for (int j = 1; j < rstep; ++j) {
            const float4 o = red[j * cq + c4];
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        red_add4_fwd(db + 4 * c4, acc.x, acc.y, acc.z, acc.w);
    }
}

// delta (B, nq, H) = sum over the head's dims of dO o (O - Qp)     (= sum_k P dP of the softmax backward)
__global__ void attn_delta_kernel(const float* __restrict__ dO, const float* __restrict__ O, const float* __restrict__ Qp,
                                  long long q_bstride, int nq, int D, int DH, long long total, float* __restrict__ delta) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over (b, q, h)
    if (i >= total) return;
    const int H = D / DH;
    const int h = (int)(i % H);
    const long long bq = i / H;
    const int q = (int)(bq % nq);
    const long long b = bq / nq;
    const float* o = O + bq * D + h * DH;
    const float* g = dO + bq * D + h * DH;
    const float* qp = Qp + b * q_bstride + (long long)q * D + h * DH;
    float s = 0.f;
    for (int j = 0; j < DH; ++j) s = fmaf(g[j], o[j] - __ldg(qp + j), s);
    delta[i] = s;
}

// Same, one warp per (b, q) row with the lanes along D (coalesced): needs D % 32 == 0 and 32 % H == 0, so that a head is a
// group of 32 / H adjacent lanes.
__global__ void attn_delta_warp_kernel(const float* __restrict__ dO, const float* __restrict__ O, const float* __restrict__ Qp,
                                       long long q_bstride, int nq, int D, int H, long long rows, float* __restrict__ delta) {
    const long long bq = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (bq >= rows) return;
    const int lane = threadIdx.x & 31;
    const int per = D >> 5;                         // consecutive elements per lane
    const int q = (int)(bq % nq);
    const long long b = bq / nq;
    const float* o = O + bq * D + lane * per;
    const float* g = dO + bq * D + lane * per;
    const float* qp = Qp + b * q_bstride + (long long)q * D + lane * per;
    float s = 0.f;
    for (int j = 0; j < per; ++j) s = fmaf(g[j], o[j] - __ldg(qp + j), s);
    const int lph = 32 / H;                         // lanes per head
    for (int off = 1; off < lph; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((lane & (lph - 1)) == 0) delta[bq * H + lane / lph] = s;
}

// out = in * keep / (1 - p); the same (seed, index) mask is regenerated by the backward pass (in-place allowed).  One 64-bit hash
// decides two elements (its two 32-bit halves against the threshold); a thread handles four consecutive elements with 16-byte
// accesses (the scalar one-hash-per-element kernel ran at 2.6 TB/s, bound by its integer multiplies).
__device__ __forceinline__ unsigned long long dropout_bits64(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);        // splitmix64 finaliser as a counter-based generator
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long i, uint32_t thresh) {
    const unsigned long long z = dropout_bits64(seed, i >> 1);
    return (uint32_t)((i & 1) ? (z >> 32) : z) >= thresh;
}
__global__ void dropout_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, unsigned long long seed,
                               uint32_t thresh, float inv_keep) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        const float4 v = *reinterpret_cast<const float4*>(in + i);
        const unsigned long long z0 = dropout_bits64(seed, (unsigned long long)i >> 1), z1 = dropout_bits64(seed, ((unsigned long long)i >> 1) + 1);
        float4 o;
        o.x = (uint32_t)z0 >= thresh ? v.x * inv_keep : 0.f;
        o.y = (uint32_t)(z0 >> 32) >= thresh ? v.y * inv_keep : 0.f;
        o.z = (uint32_t)z1 >= thresh ? v.z * inv_keep : 0.f;
        o.w = (uint32_t)(z1 >> 32) >= thresh ? v.w * inv_keep : 0.f;
        *reinterpret_cast<float4*>(out + i) = o;
    } else {
        for (long long j = i; j < n && j < i + 4; ++j) out[j] = dropout_keep(seed, (unsigned long long)j, thresh) ? in[j] * inv_keep : 0.f;
    }
}
static int launch_dropout(const float* in, float* out, long long n, float p, unsigned long long seed, cudaStream_t st) {
    if (n == 0) return 0;
    const double t = (double)p * 4294967296.0;
    const uint32_t thresh = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
    dropout_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, st>>>(in, out, n, seed, thresh, 1.0f / (1.0f - p));
    PCA_CHECK_LAUNCH("dropout_kernel");
    return 0;
}

// 16-byte vector reduction (REDG.ADD.F32x4): one L2 operation for four adjacent floats
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ------------------------------------------------------------------------------------ attention backward
// Warp = head.  Lane = (dim slice g of G, own item, loop slice): an "own" item (MODE 0: a query -> dQp; MODE 1: a key -> dKp,
// dVp) keeps its DH = G*DL head dims in the registers of G adjacent lanes; the "loop" items (MODE 0: keys; MODE 1: queries)
// are staged through shared memory and read as broadcasts.  Loop ranges are cut across blockIdx.y; results are added atomically.
template <int DL, int G, int MODE, int R>
__global__ void attn_bwd_kernel(const float* __restrict__ Qp, long long q_bstride, const float* __restrict__ KV,
                                const float* __restrict__ dO, const float* __restrict__ lse, const float* __restrict__ delta,
                                int nq, int nk, int D, int town_log, int tl, int chunk, int own_tiles, float scale, float scale_log2e,
                                float* __restrict__ dQp, float* __restrict__ dKV, const int* __restrict__ key_counts) {
    extern __shared__ __align__(16) float rows_s[];
    constexpr int DH = DL * G;
    const int H = blockDim.x >> 5;
    const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TOWN = 1 << town_log;
    const int LS = 32 / (G * TOWN);
    const int g = lane & (G - 1);
    const int own_l = (lane / G) & (TOWN - 1);
    const int ls = lane / (G * TOWN);
    const int b = blockIdx.z;
    // variable-size sets: only the first key_counts[b] keys of the padded set exist (their gradients stay zero otherwise)
    const int nk_b = key_counts ? max(1, min(nk, __ldg(key_counts + b))) : nk;
    const int n_own = MODE == 0 ? nq : nk_b;
    const int n_loop = MODE == 0 ? nk_b : nq;
    const int H2 = (2 * H + 3) & ~3;
    const int rs = MODE == 0 ? 2 * D + 4 : 2 * D + H2 + 4;        // smem row stride (floats)
    const int hoff = h * DH + g * DL;
    // Register blocking: a thread group owns R items (own_l, own_l + TOWN, ...), so every streamed row element read from
    // shared memory feeds R of them (ncu at R = 1: l1tex 94-96 % of peak, the broadcast reads are the binding unit).
    // own_tiles > 1 only when the loop set is a single smem tile (staged once, reused by every own tile of the block)
  for (int ot = 0; ot < own_tiles; ++ot) {
    const int own0 = (blockIdx.x * own_tiles + ot) * TOWN * R;
    if (own0 >= n_own) break;                                      // block-uniform
    int own[R];
    bool ovalid[R];
    float a0[R][DL], a1[R][DL], acc0[R][DL], acc1[R][DL];
    float lse_o[R], delta_o[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        own[r] = own0 + r * TOWN + own_l;
        ovalid[r] = own[r] < n_own;
        const int oc = ovalid[r] ? own[r] : 0;
        lse_o[r] = 0.f;
        delta_o[r] = 0.f;
        if (MODE == 0) {
            const float* qp = Qp + (long long)b * q_bstride + (long long)oc * D + hoff;
            const float* gp = dO + ((long long)b * nq + oc) * D + hoff;
#pragma unroll
            for (int j = 0; j < DL; j += 4) {           // per-thread rows: 16-byte accesses
                const float4 t0 = __ldg(reinterpret_cast<const float4*>(qp + j)), t1 = __ldg(reinterpret_cast<const float4*>(gp + j));
                a0[r][j] = t0.x * scale_log2e; a0[r][j + 1] = t0.y * scale_log2e; a0[r][j + 2] = t0.z * scale_log2e; a0[r][j + 3] = t0.w * scale_log2e;
                a1[r][j] = t1.x; a1[r][j + 1] = t1.y; a1[r][j + 2] = t1.z; a1[r][j + 3] = t1.w;
            }
            lse_o[r] = __ldg(lse + ((long long)b * nq + oc) * H + h);
            delta_o[r] = __ldg(delta + ((long long)b * nq + oc) * H + h);
        } else {
            const float* kp = KV + ((long long)b * nk + oc) * 2 * D + hoff;
#pragma unroll
            for (int j = 0; j < DL; j += 4) {
                const float4 t0 = __ldg(reinterpret_cast<const float4*>(kp + j)), t1 = __ldg(reinterpret_cast<const float4*>(kp + D + j));
                a0[r][j] = t0.x; a0[r][j + 1] = t0.y; a0[r][j + 2] = t0.z; a0[r][j + 3] = t0.w;
                a1[r][j] = t1.x; a1[r][j + 1] = t1.y; a1[r][j + 2] = t1.z; a1[r][j + 3] = t1.w;
            }
        }
#pragma unroll
        for (int j = 0; j < DL; ++j) { acc0[r][j] = 0.f; acc1[r][j] = 0.f; }
    }

    const int l_begin = blockIdx.y * chunk;
    const int l_end = min(n_loop, l_begin + chunk);
    for (int lt = l_begin; lt < l_end; lt += tl) {
        const int tn = min(tl, l_end - lt);
        if (ot > 0) {
            // the single loop tile is already staged
        } else if (MODE == 0) {
            const int vec_per_row = (2 * D) >> 2;
            const float* kvb = KV + ((long long)b * nk + lt) * 2 * D;
            for (int i = threadIdx.x; i < tn * vec_per_row; i += blockDim.x) {
                const int r = i / vec_per_row, c = i - r * vec_per_row;
                *reinterpret_cast<float4*>(rows_s + r * rs + c * 4) = __ldg(reinterpret_cast<const float4*>(kvb + (long long)r * 2 * D) + c);
            }
        } else {
            const int vq = D >> 2;
            for (int i = threadIdx.x; i < tn * vq; i += blockDim.x) {
                const int r = i / vq, c = i - r * vq;
                *reinterpret_cast<float4*>(rows_s + r * rs + c * 4) =
                    __ldg(reinterpret_cast<const float4*>(Qp + (long long)b * q_bstride + (long long)(lt + r) * D) + c);
                *reinterpret_cast<float4*>(rows_s + r * rs + D + c * 4) =
                    __ldg(reinterpret_cast<const float4*>(dO + ((long long)b * nq + lt + r) * D) + c);
            }
            for (int i = threadIdx.x; i < tn * H; i += blockDim.x) {
                const int r = i / H, c = i - r * H;
                rows_s[r * rs + 2 * D + c] = __ldg(lse + ((long long)b * nq + lt + r) * H + c);
                rows_s[r * rs + 2 * D + H + c] = __ldg(delta + ((long long)b * nq + lt + r) * H + c);
            }
        }
        __syncthreads();
        // warp-uniform trip count (the dim-slice reductions are full-warp shuffles); out-of-range items contribute zero
        for (int it = 0; it < tn; it += LS) {
            const int li = it + ls;
            const bool lv = li < tn;
            const float* row = rows_s + (lv ? li : 0) * rs;
            // the streamed row's slice, read once for all R own items
            float x0[DL], x1[DL];
#pragma unroll
            for (int j = 0; j < DL; j += 4) {
                const float4 t0 = *reinterpret_cast<const float4*>(row + hoff + j), t1 = *reinterpret_cast<const float4*>(row + D + hoff + j);
                x0[j] = t0.x; x0[j + 1] = t0.y; x0[j + 2] = t0.z; x0[j + 3] = t0.w;
                x1[j] = t1.x; x1[j + 1] = t1.y; x1[j + 2] = t1.z; x1[j + 3] = t1.w;
            }
            float lse_l = 0.f, delta_l = 0.f;
            if (MODE == 1) { lse_l = row[2 * D + h]; delta_l = row[2 * D + H + h]; }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float s = 0.f, dp = 0.f;
#pragma unroll
                for (int j = 0; j < DL; ++j) { s = fmaf(a0[r][j], x0[j], s); dp = fmaf(a1[r][j], x1[j], dp); }
#pragma unroll
                for (int o = 1; o < G; o <<= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); dp += __shfl_xor_sync(0xffffffffu, dp, o); }
                if (MODE == 0) {
                    // x0 = k, x1 = v; a0 = q (scaled), a1 = dO
                    const float p = lv ? exp2f(s - lse_o[r]) : 0.f;
                    const float ds = p * (dp - delta_o[r]) * scale;
#pragma unroll
                    for (int j = 0; j < DL; ++j) acc0[r][j] = fmaf(ds, x0[j], acc0[r][j]);
                } else {
                    // x0 = q, x1 = dO; a0 = k, a1 = v
                    const float p = lv ? exp2f(s * scale_log2e - lse_l) : 0.f;
                    const float ds = p * (dp - delta_l) * scale;
#pragma unroll
                    for (int j = 0; j < DL; ++j) { acc0[r][j] = fmaf(ds, x0[j], acc0[r][j]); acc1[r][j] = fmaf(p, x1[j], acc1[r][j]); }
                }
            }
        }
        if (own_tiles == 1) __syncthreads();
    }
    // merge the loop slices of one own item across lanes
#pragma unroll
    for (int r = 0; r < R; ++r) {
        for (int off = G * TOWN; off < 32; off <<= 1) {
#pragma unroll
            for (int j = 0; j < DL; ++j) {
                acc0[r][j] += __shfl_xor_sync(0xffffffffu, acc0[r][j], off);
                if (MODE == 1) acc1[r][j] += __shfl_xor_sync(0xffffffffu, acc1[r][j], off);
            }
        }
        if (ovalid[r] && ls == 0) {
            if (MODE == 0) {
                float* o = dQp + ((long long)b * nq + own[r]) * D + hoff;
#pragma unroll
                for (int j = 0; j < DL; j += 4) red_add4(o + j, acc0[r][j], acc0[r][j + 1], acc0[r][j + 2], acc0[r][j + 3]);
            } else {
                float* o = dKV + ((long long)b * nk + own[r]) * 2 * D + hoff;
#pragma unroll
                for (int j = 0; j < DL; j += 4) {
                    red_add4(o + j, acc0[r][j], acc0[r][j + 1], acc0[r][j + 2], acc0[r][j + 3]);
                    red_add4(o + D + j, acc1[r][j], acc1[r][j + 1], acc1[r][j + 2], acc1[r][j + 3]);
                }
            }
        }
    }
  }
}

template <int DL, int G, int MODE, int R>
static int launch_attn_bwd_r(const float* Qp, long long q_bstride, const float* KV, const float* dO, const float* lse,
                             const float* delta, int B, int nq, int nk, int D, int H, float* dQp, float* dKV, cudaStream_t st,
                             const int* key_counts) {
    const int n_own = MODE == 0 ? nq : nk, n_loop = MODE == 0 ? nk : nq;
    int town = 32 / G;
    while (town > 1 && (town >> 1) >= n_own) town >>= 1;
    int town_log = 0;
    while ((1 << town_log) < town) ++town_log;
    const int H2 = (2 * H + 3) & ~3;
    const int rs = MODE == 0 ? 2 * D + 4 : 2 * D + H2 + 4;
    int tl = (40 * 1024) / (rs * 4);
    tl = tl > 128 ? 128 : tl;
    if (tl < 1) tl = 1;
    if (tl > n_loop) tl = n_loop;
    const size_t smem = (size_t)tl * rs * 4;
    const long long base_blocks = (long long)B * ((n_own + town * R - 1) / (town * R));
    int nsplit = 1;
    const long long target = 148LL * 4;
    if (base_blocks < target) {
        nsplit = (int)((target + base_blocks - 1) / base_blocks);
        const int max_split = (n_loop + 2 * tl - 1) / (2 * tl);
        if (nsplit > max_split) nsplit = max_split;
        if (nsplit < 1) nsplit = 1;
    }
    int chunk = (n_loop + nsplit - 1) / nsplit;
    chunk = (chunk + tl - 1) / tl * tl;
    nsplit = (n_loop + chunk - 1) / chunk;
    // the whole loop set in one smem tile: a block stages it once and walks several own tiles (keeps >= ~8 blocks per SM)
    const int own_blocks = (n_own + town * R - 1) / (town * R);
    int own_tiles = 1;
    if (nsplit == 1 && n_loop <= tl) {
        long long ot = ((long long)own_blocks * B) / (148LL * 8);
        own_tiles = (int)(ot < 1 ? 1 : (ot > 16 ? 16 : ot));
    }
    dim3 grid((own_blocks + own_tiles - 1) / own_tiles, nsplit, B);
    const float scale = 1.0f / sqrtf((float)D);
    if (smem > 48 * 1024)
        PCA_CHECK_CUDA((cudaFuncSetAttribute(attn_bwd_kernel<DL, G, MODE, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    {
        LaunchTimer lt(MODE == 0 ? "attn_bwd_dq_kernel" : "attn_bwd_dkv_kernel", st, (MODE == 0 ? 6.0 : 8.0) * B * nq * (double)nk * D,
                       4.0 * ((double)B * nk * 2 * D + 3.0 * B * nq * D));
        attn_bwd_kernel<DL, G, MODE, R><<<grid, 32 * H, smem, st>>>(Qp, q_bstride, KV, dO, lse, delta, nq, nk, D, town_log, tl, chunk,
                                                                own_tiles, scale, scale * 1.4426950408889634f, dQp, dKV, key_counts);
    }
    PCA_CHECK_LAUNCH("attn_bwd_kernel");
    return 0;
}

// two own items per thread group when the own set is large enough to keep every lane busy and the grid full
template <int DL, int G, int MODE>
static int launch_attn_bwd_t(const float* Qp, long long q_bstride, const float* KV, const float* dO, const float* lse,
                             const float* delta, int B, int nq, int nk, int D, int H, float* dQp, float* dKV, cudaStream_t st,
                             const int* key_counts) {
    const int n_own = MODE == 0 ? nq : nk;
    const int per_warp = 2 * (32 / G);
    if (n_own >= 2 * per_warp && (long long)B * (n_own / per_warp) >= 148LL * 4)
        return launch_attn_bwd_r<DL, G, MODE, 2>(Qp, q_bstride, KV, dO, lse, delta, B, nq, nk, D, H, dQp, dKV, st, key_counts);
    return launch_attn_bwd_r<DL, G, MODE, 1>(Qp, q_bstride, KV, dO, lse, delta, B, nq, nk, D, H, dQp, dKV, st, key_counts);
}

template <int MODE>
static int launch_attn_bwd(const float* Qp, long long q_bstride, const float* KV, const float* dO, const float* lse,
                           const float* delta, int B, int nq, int nk, int D, int H, float* dQp, float* dKV, cudaStream_t st,
                           const int* key_counts) {
    if (B > 65535) return fail(PCA_EUNSUPPORTED, "attention backward: batch %d exceeds the grid limit", B);
    switch (D / H) {
        case 4: return launch_attn_bwd_t<4, 1, MODE>(Qp, q_bstride, KV, dO, lse, delta, B, nq, nk, D, H, dQp, dKV, st, key_counts);
        case 8: return launch_attn_bwd_t<8, 1, MODE>(Qp, q_bstride, KV, dO, lse, delta, B, nq, nk, D, H, dQp, dKV, st, key_counts);
        case 16: return launch_attn_bwd_t<16, 1, MODE>(Qp, q_bstride, KV, dO, lse, delta, B, nq, nk, D, H, dQp, dKV, st, key_counts);
        case 32: return launch_attn_bwd_t<16, 2, MODE>(Qp, q_bstride, KV, dO, lse, delta, B, nq, nk, D, H, dQp, dKV, st, key_counts);
        case 64: return launch_attn_bwd_t<16, 4, MODE>(Qp, q_bstride, KV, dO, lse, delta, B, nq, nk, D, H, dQp, dKV, st, key_counts);
        default: return fail(PCA_EUNSUPPORTED, "attention backward: head dim %d not in {4,8,16,32,64}", D / H);
    }
}

// ------------------------------------------------------------------------------------ stand-alone Dropout / Linear backward
int dropout_api(const float* in, float* out, long long n, float p, unsigned long long seed, cudaStream_t st) {
    if (!(p >= 0.f && p < 1.f)) return fail(PCA_EINVAL, "dropout: probability %f outside [0, 1)", p);
    if (p == 0.f) {
        if (in != out && n > 0) PCA_CHECK_CUDA(cudaMemcpyAsync(out, in, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    return launch_dropout(in, out, n, p, seed, st);
}
// dparams = dW (dout, din) | db (dout), overwritten; dX optional
int linear_bwd_api(const float* dY, const float* X, long long rows, int din, int dout, const float* W, float* dX, float* dparams,
                   cudaStream_t st) {
    PCA_CHECK_CUDA(cudaMemsetAsync(dparams, 0, ((size_t)dout * din + dout) * sizeof(float), st));
    PCA_TRY(launch_grad_weight_bias(dY, X, dparams, dparams + (size_t)dout * din, rows, din, dout, st));
    if (dX != nullptr) PCA_TRY(launch_grad_input(dY, W, dX, nullptr, rows, din, dout, st));
    return 0;
}

// ------------------------------------------------------------------------------------ MAB forward (saving) / backward
// Opre / pre1: the inputs of ln0 / ln1 (LayerNorm branches only; O and out then hold the normalised tensors)
struct MabSaved { float *Qp, *KV, *O, *R, *lse, *out, *Opre, *pre1, *P, *Z, *Gq; };   // P (probabilities), Z, Gq: tensor-core attention (attn_tc.cu)

// ------------------------------------------------------------------------------------ LayerNorm backward
// y = (x - mean) rstd gamma + beta per row.  dx = rstd (g gamma - mean(g gamma) - xhat mean(g gamma xhat)); dgamma += sum g xhat,
// dbeta += sum g.  One warp per row, 8 rows per block; the parameter gradients are reduced in shared memory per block, then
// added atomically.  g and dx may alias (every lane reads its elements before it writes them).
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* g, const float* __restrict__ x,
                                                            const float* __restrict__ gamma, long long rows, int D,
                                                            float* dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta) {
    extern __shared__ float ln_s[];                 // 2 * D: dgamma | dbeta partials of the block
    for (int j = threadIdx.x; j < 2 * D; j += blockDim.x) ln_s[j] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long r = (long long)blockIdx.x * 8 + w;
    if (r < rows) {
        const float* xr = x + r * D;
        const float* gr = g + r * D;
        float sum = 0.f;
        for (int j = lane; j < D; j += 32) sum += xr[j];
        const float mean = warp_sum(sum) / D;
        float var = 0.f;
        for (int j = lane; j < D; j += 32) { const float d = xr[j] - mean; var += d * d; }
        const float rstd = rsqrtf(warp_sum(var) / D + 1e-5f);
        float s1 = 0.f, s2 = 0.f;
        for (int j = lane; j < D; j += 32) {
            const float xh = (xr[j] - mean) * rstd, gg = gr[j] * __ldg(gamma + j);
            s1 += gg;
            s2 += gg * xh;
        }
        s1 = warp_sum(s1) / D;
        s2 = warp_sum(s2) / D;
        for (int j = lane; j < D; j += 32) {
            const float xh = (xr[j] - mean) * rstd, gv = gr[j];
            atomicAdd(&ln_s[j], gv * xh);
            atomicAdd(&ln_s[D + j], gv);
            dx[r * D + j] = rstd * (gv * __ldg(gamma + j) - s1 - xh * s2);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        atomicAdd(dgamma + j, ln_s[j]);
        atomicAdd(dbeta + j, ln_s[D + j]);
    }
}
static int launch_layernorm_bwd(const float* g, const float* x, const float* gamma, long long rows, int D, float* dx, float* dgamma,
                                float* dbeta, cudaStream_t st) {
    if (rows == 0) return 0;
    layernorm_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 2 * D * sizeof(float), st>>>(g, x, gamma, rows, D, dx, dgamma, dbeta);
    PCA_CHECK_LAUNCH("layernorm_bwd_kernel");
    return 0;
}

static MabSaved mab_saved_take(Arena& a, int B, int qb, int nq, int nk, int D, int H, int ln = 0) {
    MabSaved s;
    s.Qp = a.take<float>((size_t)qb * nq * D);
    s.KV = a.take<float>((size_t)B * nk * 2 * D);
    s.O = a.take<float>((size_t)B * nq * D);
    s.R = a.take<float>((size_t)B * nq * D);
    s.lse = a.take<float>((size_t)B * nq * H);
    s.out = a.take<float>((size_t)B * nq * D);
    s.Opre = ln ? a.take<float>((size_t)B * nq * D) : nullptr;
    s.pre1 = ln ? a.take<float>((size_t)B * nq * D) : nullptr;
    const size_t pf = attn_tc_p_floats(B, nq, nk, D, H);       // (points, heads x small side) probabilities: the backward reuses them
    s.P = pf ? a.take<float>(pf) : nullptr;
    // shared small query set (I / S against the points): room for the folded form's per-cloud sums and query matrix
    const size_t zf = qb == 1 ? attn_fold_z_floats(B, nq, nk, D, H) : 0;
    s.Z = zf ? a.take<float>(zf) : nullptr;
    s.Gq = zf ? a.take<float>(attn_fold_gq_floats(B, nq, nk, D, H)) : nullptr;
    return s;
}

static size_t train_img_bytes(int D) { return gemm_tc_image_bytes(2 * D, D); }

// the folded form of a block (attn_tc.cu): shared queries, no key counts, eligible dims, switch on
static bool mab_folded(int qb, int B, int nq, int nk, int dk, int D, int H, const int* key_counts) {
    return qb == 1 && !key_counts && attn_fold_train_on() && attn_fold_eligible(B, nq, nk, dk, D, H);
}

static int mab_train_forward(const MabSaved& s, const float* Qin, int qb, const float* Kin, int B, int nq, int nk, int dq, int dk,
                             int D, int H, const float* params, float* part, void* img, cudaStream_t st,
                             const int* key_counts = nullptr, int ln = 0) {
    const MabParams m = mab_slice(params, dq, dk, D, ln);
    const size_t ib = train_img_bytes(D);
    PCA_TRY(launch_linear(Qin, m.Wq, m.bq, s.Qp, (long long)qb * nq, dq, D, 0, st, nullptr, dq <= D ? img : nullptr, ib));
    if (mab_folded(qb, B, nq, nk, dk, D, H, key_counts) && s.Z) {
        // shared inducing points / seeds against many points: the block runs on the un-projected keys, forward and backward
        // (attn_tc.cu: the K | V projection and its gradient GEMMs never run); P, Z and Gq are kept for the backward
        PCA_TRY(launch_attn_folded(s.Qp, m.Wkv, m.bkv, Kin, B, nq, nk, dk, D, H, s.O, part, st, nullptr, s.P, s.Z, s.Gq));
    } else {
        PCA_TRY(launch_linear(Kin, m.Wkv, m.bkv, s.KV, (long long)B * nk, dk, 2 * D, 0, st, nullptr, dk <= D ? img : nullptr, ib));
        PCA_TRY(launch_attn(s.Qp, qb == 1 ? 0 : (long long)nq * D, s.KV, B, nq, nk, D, H, s.O, part, key_counts, st, s.lse, s.P));
    }
    const size_t nbytes = (size_t)B * nq * D * sizeof(float);
    if (ln) {                                                   // O = ln0(Qp + A V); the pre-LN tensor is kept for the backward
        PCA_CHECK_CUDA(cudaMemcpyAsync(s.Opre, s.O, nbytes, cudaMemcpyDeviceToDevice, st));
        PCA_TRY(launch_layernorm(s.O, (long long)B * nq, D, m.ln0w, m.ln0b, st));
    }
    PCA_TRY(launch_linear(s.O, m.Wo, m.bo, s.out, (long long)B * nq, D, D, 3, st, s.R, img, ib));
    if (ln) {                                                   // out = ln1(O + relu(fc_o(O)))
        PCA_CHECK_CUDA(cudaMemcpyAsync(s.pre1, s.out, nbytes, cudaMemcpyDeviceToDevice, st));
        PCA_TRY(launch_layernorm(s.out, (long long)B * nq, D, m.ln1w, m.ln1b, st));
    }
    return 0;
}

// scratch of one MAB backward (floats)
static size_t mab_bwd_ws_floats(int B, int qb, int nq, int nk, int D, int H, int ln = 0) {
    Arena a(nullptr, 0);
    if (ln) a.take<float>((size_t)B * nq * D);   // gradient at the input of ln1
    a.take<float>((size_t)B * nq * D);        // dZ, later dQp
    a.take<float>((size_t)B * nq * D);        // dO
    a.take<float>((size_t)B * nq * H);        // delta
    a.take<float>((size_t)B * nk * 2 * D);    // dKV
    if (qb == 1) a.take<float>((size_t)nq * D);
    a.take<uint8_t>(train_img_bytes(D));
    a.take<float>(attn_tc_bwd_floats(B, nq, nk, D, H));      // tensor-core attention backward (0 when the shape is not eligible)
    return a.off;
}

// dQin / dKin may be null (not needed); acc_* != 0 adds to what the buffer already holds.
static int mab_backward(const MabSaved& s, const float* Qin, int qb, const float* Kin, int B, int nq, int nk, int dq, int dk, int D,
                        int H, const float* params, float* dparams, const float* dOut, float* dQin, int acc_q, float* dKin,
                        int acc_k, void* ws, size_t ws_bytes, cudaStream_t st, const int* key_counts = nullptr, int ln = 0) {
    const MabParams m = mab_slice(params, dq, dk, D, ln);
    const MabParams g = mab_slice(dparams, dq, dk, D, ln);
    Arena a(ws, ws_bytes);
    float* g1 = ln ? a.take<float>((size_t)B * nq * D) : nullptr;
    float* dZ = a.take<float>((size_t)B * nq * D);
    float* dO = a.take<float>((size_t)B * nq * D);
    float* delta = a.take<float>((size_t)B * nq * H);
    float* dKV = a.take<float>((size_t)B * nk * 2 * D);
    float* dQ1 = qb == 1 ? a.take<float>((size_t)nq * D) : nullptr;
    const size_t ib = train_img_bytes(D);
    void* img = a.take<uint8_t>(ib);
    void* img_q = dq <= D ? img : nullptr;
    void* img_k = dk <= D ? img : nullptr;
    const bool attn_tc = !key_counts && attn_tc_eligible(B, nq, nk, D, H);
    float* attn_scratch = a.take<float>(attn_tc_bwd_floats(B, nq, nk, D, H));
    if (!a.ok()) return fail(PCA_EWORKSPACE, "MAB backward: workspace too small");
    const long long rq = (long long)B * nq, rk = (long long)B * nk;
    const long long q_bstride = qb == 1 ? 0 : (long long)nq * D;
    const long long n = rq * D;
    if (ln) {                                                   // through ln1: gradient at O + relu(fc_o(O))
        PCA_TRY(launch_layernorm_bwd(dOut, s.pre1, m.ln1w, rq, D, g1, (float*)g.ln1w, (float*)g.ln1b, st));
        dOut = g1;
    }
    if (D % 4 == 0 && D / 4 <= 256 && 256 % (D / 4) == 0 && rq >= 256) {      // ReLU mask and the bias gradient in one pass over dOut / R
        LaunchTimer lt("relu_bwd_colsum_kernel", st, 0.0, 12.0 * rq * D);
        relu_bwd_colsum_kernel<<<(unsigned)((rq + 255) / 256), 256, 0, st>>>(dOut, s.R, dZ, rq, D, (float*)g.bo);
        PCA_CHECK_LAUNCH("relu_bwd_colsum_kernel");
    } else {
        relu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dOut, s.R, dZ, n);
        PCA_CHECK_LAUNCH("relu_bwd_kernel");
        PCA_TRY(launch_colsum(dZ, rq, D, (float*)g.bo, st));
    }
    PCA_TRY(launch_grad_weight(dZ, s.O, (float*)g.Wo, rq, D, D, st));
    PCA_TRY(launch_grad_input(dZ, m.Wo, dO, dOut, rq, D, D, st, img, ib));
    const float* Oatt = s.O;                                    // Qp + A V
    if (ln) {                                                   // through ln0 (in place)
        PCA_TRY(launch_layernorm_bwd(dO, s.Opre, m.ln0w, rq, D, dO, (float*)g.ln0w, (float*)g.ln0b, st));
        Oatt = s.Opre;
    }
    const bool folded = mab_folded(qb, B, nq, nk, dk, D, H, key_counts) && s.Z;
    if (!folded && !(attn_tc && attn_tc_kind(B, nq, nk, D, H) == 1)) {      // (the small-key tensor-core backward computes sum_m P dP in its epilogue; the folded one its own row dots)
        const long long total = rq * H;
        if (D % 32 == 0 && 32 % H == 0)
            attn_delta_warp_kernel<<<(unsigned)((rq + 7) / 8), 256, 0, st>>>(dO, Oatt, s.Qp, q_bstride, nq, D, H, rq, delta);
        else
            attn_delta_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dO, Oatt, s.Qp, q_bstride, nq, D, D / H, total, delta);
        PCA_CHECK_LAUNCH("attn_delta_kernel");
    }
    bool fused_bq = false, fused_bkv = false;
    float* dQp = dZ;                                            // dQp = dO (residual) + attention part
    float* dQfold = nullptr;
    if (folded) {
        // the block on the un-projected keys: dKin, dWk, the attention part of the shared queries' gradient and the operands of
        // the Wv gradient come from attn_tc.cu; the key bias has an exactly zero gradient (softmax shift invariance)
        PCA_CHECK_CUDA(cudaMemcpyAsync(dQp, dO, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st));   // residual part
        float* dOx = nullptr;
        dQfold = dKV;                                           // (nq, D): dKV's buffer is free in this form
        PCA_TRY(launch_attn_folded_bwd(s.Qp, m.Wkv, Kin, dO, s.P, s.Z, s.Gq, B, nq, nk, dk, D, H, dKin, acc_k, (float*)g.Wkv, dQfold, &dOx,
                                       attn_scratch, st));
        const int HS = attn_fold_rows(B, nq, nk, D, H);
        PCA_TRY(launch_grad_weight(dOx, s.Z, (float*)g.Wkv + (long long)D * dk, (long long)B * HS, dk, D, st));     // dWv
        PCA_TRY(launch_colsum(dO, rq, D, (float*)g.bkv + D, st));                                                    // dbv
    } else if (attn_tc) {                                       // one small side: every contraction as a split-bf16 tensor-core GEMM
        // the bias gradients of fc_q (large-query form) / fc_k | fc_v (small-query form) fall out of the epilogues that write
        // dQp / dKV -- no separate column-sum pass (the first layers, d_in <= 4, keep their one-pass weight + bias kernel)
        const int kind = attn_tc_kind(B, nq, nk, D, H);
        fused_bq = kind == 1 && qb != 1 && dq > 4;
        fused_bkv = kind == 2 && dk > 4;
        PCA_TRY(launch_attn_bwd_tc(s.Qp, q_bstride, s.KV, dO, s.lse, delta, B, nq, nk, D, H, dQp, dKV, attn_scratch, st, s.P,
                                   fused_bq ? (float*)g.bq : nullptr, fused_bkv ? (float*)g.bkv : nullptr));
    } else {
        PCA_CHECK_CUDA(cudaMemsetAsync(dKV, 0, (size_t)rk * 2 * D * sizeof(float), st));
        PCA_TRY(launch_attn_bwd<1>(s.Qp, q_bstride, s.KV, dO, s.lse, delta, B, nq, nk, D, H, nullptr, dKV, st, key_counts));
        PCA_CHECK_CUDA(cudaMemcpyAsync(dQp, dO, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st));   // added to atomically
        PCA_TRY(launch_attn_bwd<0>(s.Qp, q_bstride, s.KV, dO, s.lse, delta, B, nq, nk, D, H, dQp, nullptr, st, key_counts));
    }
    long long rows_q = rq;
    if (qb == 1) {                                              // shared queries (I / S): sum the per-cloud gradients
        PCA_CHECK_CUDA(cudaMemsetAsync(dQ1, 0, (size_t)nq * D * sizeof(float), st));
        PCA_TRY(launch_colsum(dQp, B, nq * D, dQ1, st));
        if (dQfold) {                                           // + the attention part of the folded form (one matrix for the batch)
            const long long nn = (long long)nq * D;
            add_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(dQ1, dQfold, nn);
            PCA_CHECK_LAUNCH("add_kernel");
        }
        dQp = dQ1;
        rows_q = nq;
    }
    if (fused_bq) PCA_TRY(launch_grad_weight(dQp, Qin, (float*)g.Wq, rows_q, dq, D, st));
    else PCA_TRY(launch_grad_weight_bias(dQp, Qin, (float*)g.Wq, (float*)g.bq, rows_q, dq, D, st));
    if (dQin) PCA_TRY(launch_grad_input(dQp, m.Wq, dQin, acc_q ? dQin : nullptr, rows_q, dq, D, st, img_q, ib));
    if (folded) return 0;                                       // (dWkv, dbv and dKin are done)
    if (fused_bkv) PCA_TRY(launch_grad_weight(dKV, Kin, (float*)g.Wkv, rk, dk, 2 * D, st));
    else PCA_TRY(launch_grad_weight_bias(dKV, Kin, (float*)g.Wkv, (float*)g.bkv, rk, dk, 2 * D, st));
    if (dKin) PCA_TRY(launch_grad_input(dKV, m.Wkv, dKin, acc_k ? dKin : nullptr, rk, dk, 2 * D, st, img_k, ib));
    return 0;
}

// ------------------------------------------------------------------------------------ stand-alone MAB training
// MAB.forward / backward for user models built from the blocks (modules.py:6-33): SAB = MAB(X, X), ISAB = mab1(X, mab0(I, X)),
// PMA = MAB(S, X) compose on the host through autograd.  Q (qb, nq, dq) with qb in {1, B}; K (B, nk, dk).
size_t mab_train_saved_bytes(int B, int qb, int nq, int nk, int D, int H, int ln) {
    Arena a(nullptr, 0);
    mab_saved_take(a, B, qb, nq, nk, D, H, ln);
    return a.off;
}
size_t mab_train_ws_bytes(int B, int qb, int nq, int nk, int D, int H, int ln) {
    const size_t fwd = align_up(attn_part_floats(B, nq, nk, D, H) * sizeof(float), 256) + align_up(train_img_bytes(D), 256);
    const size_t bwd = mab_bwd_ws_floats(B, qb, nq, nk, D, H, ln);
    return fwd > bwd ? fwd : bwd;
}
static int mab_train_check(int qb, int B, int nq, int nk, int dq, int dk, int D, int H) {
    if (qb != 1 && qb != B) return fail(PCA_EINVAL, "MAB training: query batch must be 1 or B");
    if (B <= 0 || nq <= 0 || nk <= 0 || dq <= 0 || dk <= 0 || D <= 0 || H <= 0) return fail(PCA_EINVAL, "MAB training: bad shape");
    if (D % H || D % 4) return fail(PCA_EINVAL, "MAB training: dim_V must be a multiple of 4 and of num_heads");
    if (B > 65535) return fail(PCA_EUNSUPPORTED, "MAB training: batch %d exceeds the grid limit", B);
    return 0;
}
int mab_train_forward_api(const float* Q, int qb, const float* K, int B, int nq, int nk, int dq, int dk, int D, int H, int ln,
                          const float* params, float* out, void* saved, size_t saved_bytes, void* ws, size_t ws_bytes, cudaStream_t st) {
    PCA_TRY(mab_train_check(qb, B, nq, nk, dq, dk, D, H));
    if (!Q || !K || !params || !out || !saved || !ws) return fail(PCA_EINVAL, "MAB training forward: null pointer");
    Arena sa(saved, saved_bytes);
    const MabSaved s = mab_saved_take(sa, B, qb, nq, nk, D, H, ln);
    if (!sa.ok() || ws_bytes < mab_train_ws_bytes(B, qb, nq, nk, D, H, ln)) return fail(PCA_EWORKSPACE, "MAB training forward: buffers too small");
    Arena wa(ws, ws_bytes);
    float* part = wa.take<float>(attn_part_floats(B, nq, nk, D, H));
    void* img = wa.take<uint8_t>(train_img_bytes(D));
    PCA_TRY(mab_train_forward(s, Q, qb, K, B, nq, nk, dq, dk, D, H, params, part, img, st, nullptr, ln));
    PCA_CHECK_CUDA(cudaMemcpyAsync(out, s.out, (size_t)B * nq * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}
int mab_train_backward_api(const float* Q, int qb, const float* K, int B, int nq, int nk, int dq, int dk, int D, int H, int ln,
                           const float* params, const float* dout, const void* saved, size_t saved_bytes, float* dparams, float* dQ,
                           float* dK, void* ws, size_t ws_bytes, cudaStream_t st) {
    PCA_TRY(mab_train_check(qb, B, nq, nk, dq, dk, D, H));
    if (!Q || !K || !params || !dout || !saved || !dparams || !ws) return fail(PCA_EINVAL, "MAB training backward: null pointer");
    Arena sa(const_cast<void*>(saved), saved_bytes);
    const MabSaved s = mab_saved_take(sa, B, qb, nq, nk, D, H, ln);
    if (!sa.ok() || ws_bytes < mab_train_ws_bytes(B, qb, nq, nk, D, H, ln)) return fail(PCA_EWORKSPACE, "MAB training backward: buffers too small");
    PCA_CHECK_CUDA(cudaMemsetAsync(dparams, 0, (size_t)mab_count(dq, dk, D, ln) * sizeof(float), st));
    return mab_backward(s, Q, qb, K, B, nq, nk, dq, dk, D, H, params, dparams, dout, dQ, 0, dK, 0, ws, ws_bytes, st, nullptr, ln);
}

// ------------------------------------------------------------------------------------ ST / SetTransformer
struct StSaved {
    MabSaved i0m0, i0m1, i1m0, i1m1, pm;
    float *Y2d, *Pd;       // dropout outputs (alias the un-dropped tensors when dropout_p == 0)
};

static StSaved st_saved_take(Arena& a, const pca_st_dims* d, int B, int N, float dropout_p) {
    StSaved s;
    const int D = d->D, H = d->H, M = d->M, S = d->S;
    s.i0m0 = mab_saved_take(a, B, 1, M, N, D, H);
    s.i0m1 = mab_saved_take(a, B, B, N, M, D, H);
    s.i1m0 = mab_saved_take(a, B, 1, M, N, D, H);
    s.i1m1 = mab_saved_take(a, B, B, N, M, D, H);
    s.Y2d = dropout_p > 0.f ? a.take<float>((size_t)B * N * D) : s.i1m1.out;
    s.pm = mab_saved_take(a, B, 1, S, N, D, H);
    s.Pd = dropout_p > 0.f ? a.take<float>((size_t)B * S * D) : s.pm.out;
    return s;
}

struct StParamOffsets { long long I0, i0m0, i0m1, I1, i1m0, i1m1, S, pm, Wl, bl, total; };
static StParamOffsets st_offsets(const pca_st_dims* d) {
    StParamOffsets o;
    const int D = d->D;
    long long p = 0;
    o.I0 = p; p += (long long)d->M * D;
    o.i0m0 = p; p += mab_count(D, d->d_in, D, 0);
    o.i0m1 = p; p += mab_count(d->d_in, D, D, 0);
    o.I1 = p; p += (long long)d->M * D;
    o.i1m0 = p; p += mab_count(D, D, D, 0);
    o.i1m1 = p; p += mab_count(D, D, D, 0);
    o.S = p; p += (long long)d->S * D;
    o.pm = p; p += mab_count(D, D, D, 0);
    o.Wl = p; p += (long long)d->C * D;
    o.bl = p; p += d->C;
    o.total = p;
    return o;
}

static int train_check(const pca_st_dims* d, int B, int N, float dropout_p) {
    if (!d) return fail(PCA_EINVAL, "ST training: null dims");
    if (d->d_in <= 0 || d->D <= 0 || d->H <= 0 || d->M <= 0 || d->S <= 0 || d->C <= 0) return fail(PCA_EINVAL, "ST training: non-positive dimension");
    if (d->D % d->H || d->D % 4) return fail(PCA_EINVAL, "ST training: dim_hidden must be a multiple of 4 and of num_heads");
    if (d->ln) return fail(PCA_EUNSUPPORTED, "ST training: ln=True is not implemented (no reference configuration trains with it)");
    if (B <= 0 || N <= 0) return fail(PCA_EINVAL, "ST training: bad batch/points (B=%d, N=%d)", B, N);
    if (B > 65535) return fail(PCA_EUNSUPPORTED, "ST training: batch %d exceeds the grid limit", B);
    if (!(dropout_p >= 0.f && dropout_p < 1.f)) return fail(PCA_EINVAL, "ST training: dropout probability %f outside [0, 1)", dropout_p);
    return 0;
}

size_t st_train_saved_bytes(const pca_st_dims* d, int B, int N, float dropout_p) {
    Arena a(nullptr, 0);
    st_saved_take(a, d, B, N, dropout_p);
    return a.off;
}

size_t st_train_ws_bytes(const pca_st_dims* d, int B, int N) {
    const int D = d->D, H = d->H, M = d->M, S = d->S;
    Arena a(nullptr, 0);
    a.take<float>((size_t)B * N * D);       // gradient of a (B, N, D) tensor (ping)
    a.take<float>((size_t)B * N * D);       // (pong)
    a.take<float>((size_t)B * M * D);       // gradient of the inducing-point summaries
    a.take<float>((size_t)B * S * D);       // gradient of the pooled vectors
    a.take<float>((size_t)B * S * D);
    size_t w = mab_bwd_ws_floats(B, 1, M, N, D, H);
    size_t w1 = mab_bwd_ws_floats(B, B, N, M, D, H);
    size_t w2 = mab_bwd_ws_floats(B, 1, S, N, D, H);
    w = w > w1 ? w : w1;
    w = w > w2 ? w : w2;
    size_t part = attn_part_floats(B, M, N, D, H), p2 = attn_part_floats(B, S, N, D, H), p3 = attn_part_floats(B, N, M, D, H);
    part = part > p2 ? part : p2;
    part = part > p3 ? part : p3;
    const size_t fwd = align_up(part * sizeof(float), 256) + align_up(train_img_bytes(D), 256);
    return a.off + (w > fwd ? w : fwd);
}

int st_train_forward(const float* X, const int* counts, int B, int N, const pca_st_dims* d, const float* params, float dropout_p,
                     unsigned long long seed, float* logits, void* saved, size_t saved_bytes, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
    PCA_TRY(train_check(d, B, N, dropout_p));
    if (!X || !params || !logits || !saved || !ws) return fail(PCA_EINVAL, "ST training forward: null pointer");
    if (ws_bytes < st_train_ws_bytes(d, B, N)) return fail(PCA_EWORKSPACE, "ST training forward: workspace too small");
    Arena sa(saved, saved_bytes);
    const StSaved s = st_saved_take(sa, d, B, N, dropout_p);
    if (!sa.ok()) return fail(PCA_EWORKSPACE, "ST training forward: activation buffer %zu B too small", saved_bytes);
    const StParamOffsets o = st_offsets(d);
    const int D = d->D, H = d->H, M = d->M, S = d->S, C = d->C, din = d->d_in;
    size_t part_floats = attn_part_floats(B, M, N, D, H);
    {
        const size_t p2 = attn_part_floats(B, S, N, D, H), p3 = attn_part_floats(B, N, M, D, H);
        part_floats = part_floats > p2 ? part_floats : p2;
        part_floats = part_floats > p3 ? part_floats : p3;
    }
    Arena wa(ws, ws_bytes);
    float* part = wa.take<float>(part_floats);
    void* img = wa.take<uint8_t>(train_img_bytes(D));
    if (!wa.ok()) return fail(PCA_EWORKSPACE, "ST training forward: workspace too small");
    PCA_TRY(mab_train_forward(s.i0m0, params + o.I0, 1, X, B, M, N, D, din, D, H, params + o.i0m0, part, img, st, counts));
    PCA_TRY(mab_train_forward(s.i0m1, X, B, s.i0m0.out, B, N, M, din, D, D, H, params + o.i0m1, part, img, st));
    PCA_TRY(mab_train_forward(s.i1m0, params + o.I1, 1, s.i0m1.out, B, M, N, D, D, D, H, params + o.i1m0, part, img, st, counts));
    PCA_TRY(mab_train_forward(s.i1m1, s.i0m1.out, B, s.i1m0.out, B, N, M, D, D, D, H, params + o.i1m1, part, img, st));
    if (dropout_p > 0.f) PCA_TRY(launch_dropout(s.i1m1.out, s.Y2d, (long long)B * N * D, dropout_p, seed, st));
    PCA_TRY(mab_train_forward(s.pm, params + o.S, 1, s.Y2d, B, S, N, D, D, D, H, params + o.pm, part, img, st, counts));
    if (dropout_p > 0.f) PCA_TRY(launch_dropout(s.pm.out, s.Pd, (long long)B * S * D, dropout_p, seed ^ 0xD1B54A32D192ED03ull, st));
    PCA_TRY(launch_linear(s.Pd, params + o.Wl, params + o.bl, logits, (long long)B * S, D, C, 0, st));
    return 0;
}

// phase 0: the whole backward.  phase 1: final Linear, PMA and ISAB 1 -- everything whose parameter gradients sit in the TAIL
// of the flat blob, dparams[*tail_offset ..); phase 2: ISAB 0 (the head of the blob), continuing from the workspace phase 1
// left behind.  The split lets the caller start the gradient all-reduce of the tail while ISAB 0 is still differentiating.
int st_train_backward(const float* X, const int* counts, int B, int N, const pca_st_dims* d, const float* params, float dropout_p,
                      unsigned long long seed, const float* dlogits, const void* saved, size_t saved_bytes, float* dparams,
                      float* dX, void* ws, size_t ws_bytes, cudaStream_t st, int phase = 0, long long* tail_offset = nullptr) {
    PCA_TRY(train_check(d, B, N, dropout_p));
    if (!X || !params || !dlogits || !saved || !dparams || !ws) return fail(PCA_EINVAL, "ST training backward: null pointer");
    Arena sa(const_cast<void*>(saved), saved_bytes);
    const StSaved s = st_saved_take(sa, d, B, N, dropout_p);
    if (!sa.ok()) return fail(PCA_EWORKSPACE, "ST training backward: activation buffer too small");
    const StParamOffsets o = st_offsets(d);
    const int D = d->D, H = d->H, M = d->M, S = d->S, C = d->C, din = d->d_in;
    Arena a(ws, ws_bytes);
    float* gA = a.take<float>((size_t)B * N * D);
    float* gB = a.take<float>((size_t)B * N * D);
    float* gH = a.take<float>((size_t)B * M * D);
    float* gP = a.take<float>((size_t)B * S * D);
    float* gP2 = a.take<float>((size_t)B * S * D);
    if (!a.ok() || ws_bytes < st_train_ws_bytes(d, B, N)) return fail(PCA_EWORKSPACE, "ST training backward: workspace too small");
    void* sub = (char*)ws + a.off;
    const size_t sub_bytes = ws_bytes - a.off;
    if (tail_offset != nullptr) *tail_offset = o.I1;
    if (phase < 0 || phase > 2) return fail(PCA_EINVAL, "ST training backward: phase %d outside {0, 1, 2}", phase);
    const long long rp = (long long)B * S;
    if (phase != 2) {
    PCA_CHECK_CUDA(cudaMemsetAsync(dparams, 0, (size_t)o.total * sizeof(float), st));
    // final Linear
    PCA_TRY(launch_grad_weight(dlogits, s.Pd, dparams + o.Wl, rp, D, C, st));
    PCA_TRY(launch_colsum(dlogits, rp, C, dparams + o.bl, st));
    PCA_TRY(launch_grad_input(dlogits, params + o.Wl, gP, nullptr, rp, D, C, st));
    if (dropout_p > 0.f) PCA_TRY(launch_dropout(gP, gP, rp * D, dropout_p, seed ^ 0xD1B54A32D192ED03ull, st));
    // PMA: mab(S, Y2d)
    PCA_TRY(mab_backward(s.pm, params + o.S, 1, s.Y2d, B, S, N, D, D, D, H, params + o.pm, dparams + o.pm, gP, dparams + o.S, 0, gA, 0,
                         sub, sub_bytes, st, counts));
    if (dropout_p > 0.f) PCA_TRY(launch_dropout(gA, gA, (long long)B * N * D, dropout_p, seed, st));
    (void)gP2;
    // ISAB 1: Y2 = mab1(Y1, H2), H2 = mab0(I1, Y1)
    PCA_TRY(mab_backward(s.i1m1, s.i0m1.out, B, s.i1m0.out, B, N, M, D, D, D, H, params + o.i1m1, dparams + o.i1m1, gA, gB, 0, gH, 0,
                         sub, sub_bytes, st));
    PCA_TRY(mab_backward(s.i1m0, params + o.I1, 1, s.i0m1.out, B, M, N, D, D, D, H, params + o.i1m0, dparams + o.i1m0, gH,
                         dparams + o.I1, 0, gB, 1, sub, sub_bytes, st, counts));
    }
    if (phase == 1) return 0;
    // ISAB 0: Y1 = mab1(X, H1), H1 = mab0(I0, X)
    PCA_TRY(mab_backward(s.i0m1, X, B, s.i0m0.out, B, N, M, din, D, D, H, params + o.i0m1, dparams + o.i0m1, gB, dX, 0, gH, 0,
                         sub, sub_bytes, st));
    PCA_TRY(mab_backward(s.i0m0, params + o.I0, 1, X, B, M, N, D, din, D, H, params + o.i0m0, dparams + o.i0m0, gH, dparams + o.I0, 0,
                         dX, 1, sub, sub_bytes, st, counts));
    return 0;
}

// ------------------------------------------------------------------------------------ DeepSet training
// set_transformer-master/models.py:3-28 (mean pool; max / sum of max_regression_demo.ipynb:41-48): enc = 4 x Linear (ReLU after
// the first three) over points, pool over points, dec = 4 x Linear (ReLU after the first three).  Saved: the four encoder
// activations, the pooled vectors, the three decoder activations and (max pool) the arg-max point of every (cloud, feature).
struct DsSaved { float *t[4], *pooled, *u[3]; int* arg; };
static DsSaved ds_saved_take(Arena& a, int B, int N, int dh) {
    DsSaved s;
    for (int i = 0; i < 4; ++i) s.t[i] = a.take<float>((size_t)B * N * dh);
    s.pooled = a.take<float>((size_t)B * dh);
    for (int i = 0; i < 3; ++i) s.u[i] = a.take<float>((size_t)B * dh);
    s.arg = a.take<int>((size_t)B * dh);
    return s;
}
size_t deepset_train_saved_bytes(int B, int N, int dh) {
    Arena a(nullptr, 0);
    ds_saved_take(a, B, N, dh);
    return a.off;
}
size_t deepset_train_ws_bytes(int B, int N, int dh) {
    Arena a(nullptr, 0);
    a.take<float>((size_t)B * N * dh);
    a.take<float>((size_t)B * N * dh);
    a.take<float>((size_t)B * dh);
    a.take<float>((size_t)B * dh);
    a.take<uint8_t>(gemm_tc_image_bytes(dh, dh));
    return a.off;
}

// pooled (B, D) over the N points of X (B, N, D): 0 mean, 1 max (arg-max kept: first maximal point), 2 sum
__global__ void pool_train_kernel(const float* __restrict__ X, int N, int D, int pool, float* __restrict__ out, int* __restrict__ arg) {
    const int b = blockIdx.y, d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const float* x = X + (long long)b * N * D + d;
    float v = pool == 1 ? -INFINITY : 0.f;
    int am = 0;
    for (int p = 0; p < N; ++p) {
        const float t = __ldg(x + (long long)p * D);
        if (pool == 1) { if (t > v) { v = t; am = p; } }
        else v += t;
    }
    if (pool == 0) v /= (float)N;
    out[(long long)b * D + d] = v;
    if (arg) arg[(long long)b * D + d] = am;
}
// dT (B, N, D) from dPooled (B, D), optionally masked by the ReLU... (the last encoder layer has no ReLU)
__global__ void pool_bwd_kernel(const float* __restrict__ dP, const int* __restrict__ arg, int N, int D, int pool, long long total,
                                float* __restrict__ dT) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int d = (int)(i % D);
    const long long bp = i / D;
    const int p = (int)(bp % N);
    const long long b = bp / N;
    const float g = __ldg(dP + b * D + d);
    dT[i] = pool == 0 ? g / (float)N : (pool == 2 ? g : (__ldg(arg + b * D + d) == p ? g : 0.f));
}

static void ds_params(const float* p, int d_in, int dh, int out_dim, const float** W, const float** bb) {
    for (int i = 0; i < 8; ++i) {
        const int dout = (i == 7) ? out_dim : dh;
        const int di = (i == 0) ? d_in : dh;
        W[i] = p; p += (long long)dout * di;
        bb[i] = p; p += dout;
    }
}

int deepset_train_forward(const float* X, int B, int N, int d_in, int dh, int out_dim, int pool, const float* params, float* out,
                          void* saved, size_t saved_bytes, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!X || !params || !out || !saved || !ws) return fail(PCA_EINVAL, "DeepSet training forward: null pointer");
    if (B <= 0 || N <= 0 || d_in <= 0 || dh <= 0 || out_dim <= 0 || B > 65535) return fail(PCA_EINVAL, "DeepSet training: bad shape");
    if (pool < 0 || pool > 2) return fail(PCA_EINVAL, "DeepSet training: pool %d not in {0 mean, 1 max, 2 sum}", pool);
    Arena sa(saved, saved_bytes);
    const DsSaved s = ds_saved_take(sa, B, N, dh);
    if (!sa.ok() || ws_bytes < deepset_train_ws_bytes(B, N, dh)) return fail(PCA_EWORKSPACE, "DeepSet training forward: buffers too small");
    Arena wa(ws, ws_bytes);
    wa.take<float>((size_t)B * N * dh); wa.take<float>((size_t)B * N * dh); wa.take<float>((size_t)B * dh); wa.take<float>((size_t)B * dh);
    const size_t ib = gemm_tc_image_bytes(dh, dh);
    void* img = wa.take<uint8_t>(ib);
    const float* W[8]; const float* bb[8];
    ds_params(params, d_in, dh, out_dim, W, bb);
    const long long rows = (long long)B * N;
    PCA_TRY(launch_linear(X, W[0], bb[0], s.t[0], rows, d_in, dh, 1, st));
    PCA_TRY(launch_linear(s.t[0], W[1], bb[1], s.t[1], rows, dh, dh, 1, st, nullptr, img, ib));
    PCA_TRY(launch_linear(s.t[1], W[2], bb[2], s.t[2], rows, dh, dh, 1, st, nullptr, img, ib));
    PCA_TRY(launch_linear(s.t[2], W[3], bb[3], s.t[3], rows, dh, dh, 0, st, nullptr, img, ib));
    {
        dim3 grid((dh + 127) / 128, B);
        pool_train_kernel<<<grid, 128, 0, st>>>(s.t[3], N, dh, pool, s.pooled, pool == 1 ? s.arg : nullptr);
        PCA_CHECK_LAUNCH("pool_train_kernel");
    }
    PCA_TRY(launch_linear(s.pooled, W[4], bb[4], s.u[0], B, dh, dh, 1, st));
    PCA_TRY(launch_linear(s.u[0], W[5], bb[5], s.u[1], B, dh, dh, 1, st));
    PCA_TRY(launch_linear(s.u[1], W[6], bb[6], s.u[2], B, dh, dh, 1, st));
    PCA_TRY(launch_linear(s.u[2], W[7], bb[7], out, B, dh, out_dim, 0, st));
    return 0;
}

int deepset_train_backward(const float* X, int B, int N, int d_in, int dh, int out_dim, int pool, const float* params,
                           const float* dout, const void* saved, size_t saved_bytes, float* dparams, float* dX, void* ws,
                           size_t ws_bytes, cudaStream_t st) {
    if (!X || !params || !dout || !saved || !dparams || !ws) return fail(PCA_EINVAL, "DeepSet training backward: null pointer");
    if (B <= 0 || N <= 0 || B > 65535) return fail(PCA_EINVAL, "DeepSet training: bad shape");
    Arena sa(const_cast<void*>(saved), saved_bytes);
    const DsSaved s = ds_saved_take(sa, B, N, dh);
    if (!sa.ok() || ws_bytes < deepset_train_ws_bytes(B, N, dh)) return fail(PCA_EWORKSPACE, "DeepSet training backward: buffers too small");
    Arena wa(ws, ws_bytes);
    float* gA = wa.take<float>((size_t)B * N * dh);
    float* gB = wa.take<float>((size_t)B * N * dh);
    float* hA = wa.take<float>((size_t)B * dh);
    float* hB = wa.take<float>((size_t)B * dh);
    const size_t ib = gemm_tc_image_bytes(dh, dh);
    void* img = wa.take<uint8_t>(ib);
    const float* W[8]; const float* bb[8];
    ds_params(params, d_in, dh, out_dim, W, bb);
    long long total = 0;
    for (int i = 0; i < 8; ++i) total += (long long)((i == 7) ? out_dim : dh) * ((i == 0) ? d_in : dh) + ((i == 7) ? out_dim : dh);
    PCA_CHECK_CUDA(cudaMemsetAsync(dparams, 0, (size_t)total * sizeof(float), st));
    float* dW[8]; float* db[8];
    {
        float* q = dparams;
        for (int i = 0; i < 8; ++i) {
            const int do_ = (i == 7) ? out_dim : dh, di = (i == 0) ? d_in : dh;
            dW[i] = q; q += (long long)do_ * di;
            db[i] = q; q += do_;
        }
    }
    auto relu_mask = [&](const float* g, const float* act, float* o, long long n) -> int {
        relu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, act, o, n);
        PCA_CHECK_LAUNCH("relu_bwd_kernel");
        return 0;
    };
    // ---- decoder (rows = B)
    PCA_TRY(launch_grad_weight(dout, s.u[2], dW[7], B, dh, out_dim, st));
    PCA_TRY(launch_colsum(dout, B, out_dim, db[7], st));
    PCA_TRY(launch_grad_input(dout, W[7], hA, nullptr, B, dh, out_dim, st));
    const float* acts[3] = {s.u[2], s.u[1], s.u[0]};
    const float* ins[3] = {s.u[1], s.u[0], s.pooled};
    float* cur = hA; float* oth = hB;
    for (int l = 0; l < 3; ++l) {                     // layers 6, 5, 4
        const int li = 6 - l;
        PCA_TRY(relu_mask(cur, acts[l], cur, (long long)B * dh));
        PCA_TRY(launch_grad_weight(cur, ins[l], dW[li], B, dh, dh, st));
        PCA_TRY(launch_colsum(cur, B, dh, db[li], st));
        PCA_TRY(launch_grad_input(cur, W[li], oth, nullptr, B, dh, dh, st));
        float* t = cur; cur = oth; oth = t;
    }
    // ---- pool: cur = d pooled (B, dh) -> gA = d t3 (B, N, dh)
    const long long rows = (long long)B * N, n_el = rows * dh;
    pool_bwd_kernel<<<(unsigned)((n_el + 255) / 256), 256, 0, st>>>(cur, s.arg, N, dh, pool, n_el, gA);
    PCA_CHECK_LAUNCH("pool_bwd_kernel");
    // ---- encoder layer 3 (no ReLU), then 2, 1 (ReLU), then 0
    PCA_TRY(launch_grad_weight(gA, s.t[2], dW[3], rows, dh, dh, st));
    PCA_TRY(launch_colsum(gA, rows, dh, db[3], st));
    PCA_TRY(launch_grad_input(gA, W[3], gB, nullptr, rows, dh, dh, st, img, ib));
    float* g = gB; float* go = gA;
    for (int li = 2; li >= 1; --li) {
        PCA_TRY(relu_mask(g, s.t[li], g, n_el));
        PCA_TRY(launch_grad_weight(g, s.t[li - 1], dW[li], rows, dh, dh, st));
        PCA_TRY(launch_colsum(g, rows, dh, db[li], st));
        PCA_TRY(launch_grad_input(g, W[li], go, nullptr, rows, dh, dh, st, img, ib));
        float* t = g; g = go; go = t;
    }
    PCA_TRY(relu_mask(g, s.t[0], g, n_el));
    PCA_TRY(launch_grad_weight_bias(g, X, dW[0], db[0], rows, d_in, dh, st));
    if (dX) PCA_TRY(launch_grad_input(g, W[0], dX, nullptr, rows, d_in, dh, st));
    return 0;
}

// ------------------------------------------------------------------------------------ loss and optimizer
// nn.CrossEntropyLoss (mean reduction; Code/settransformer.py:89, main_pointcloud.py:63): one warp per row.
// loss_sum[0] += sum_b (lse_b - z_b[label_b]) * inv_batch; dlogits = (softmax - onehot) * inv_batch; correct[0] += [argmax == label]
__global__ void cross_entropy_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int B, int C,
                                     float inv_batch, float* __restrict__ loss_sum, int* __restrict__ correct,
                                     float* __restrict__ dlogits) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= B) return;
    const int lane = threadIdx.x & 31;
    const float* z = logits + (long long)r * C;
    float mx = -INFINITY;
    int arg = 0;
    for (int j = lane; j < C; j += 32) {
        const float v = z[j];
        if (v > mx) { mx = v; arg = j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float s = 0.f;
    for (int j = lane; j < C; j += 32) s += expf(z[j] - mx);
    s = warp_sum(s);
    const long long lbl64 = labels[r];
    const float lse = mx + logf(s);
    if (lbl64 < 0 || lbl64 >= C) {
        // A label outside [0, C) (nn.CrossEntropyLoss's ignore_index = -100 included: it is NOT supported) must not index the
        // logits.  torch raises a device-side assert here; this kernel poisons the loss and the row's gradient with NaN so the
        // mistake is loud on the very first step instead of training on garbage.
        const float qnan = __int_as_float(0x7fc00000);
        if (dlogits)
            for (int j = lane; j < C; j += 32) dlogits[(long long)r * C + j] = qnan;
        if (lane == 0) atomicAdd(loss_sum, qnan);
        return;
    }
    const int lbl = (int)lbl64;
    if (dlogits)
        for (int j = lane; j < C; j += 32) dlogits[(long long)r * C + j] = (expf(z[j] - lse) - (j == lbl ? 1.f : 0.f)) * inv_batch;
    if (lane == 0) {
        atomicAdd(loss_sum, (lse - z[lbl]) * inv_batch);
        if (correct && arg == lbl) atomicAdd(correct, 1);
    }
}

int launch_cross_entropy(const float* logits, const long long* labels, int B, int C, float inv_batch, float* loss_sum, int* correct,
                         float* dlogits, cudaStream_t st) {
    if (B <= 0 || C <= 0) return fail(PCA_EINVAL, "cross entropy: empty batch");
    cross_entropy_kernel<<<(B + 7) / 8, 256, 0, st>>>(logits, labels, B, C, inv_batch, loss_sum, correct, dlogits);
    PCA_CHECK_LAUNCH("cross_entropy_kernel");
    return 0;
}

// torch.optim.Adam semantics (L2 weight decay folded into the gradient, bias-corrected moments), one launch over the flat blob.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                            float grad_scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = g[i] * grad_scale;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
}

int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd,
                int step, float grad_scale, cudaStream_t st) {
    if (n <= 0) return 0;
    if (step < 1) return fail(PCA_EINVAL, "adam: step counts from 1");
    const float bc1 = 1.f - powf(b1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(b2, (float)step));
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, wd, bc1, bc2_sqrt, grad_scale);
    PCA_CHECK_LAUNCH("adam_kernel");
    return 0;
}

}  // namespace pca
