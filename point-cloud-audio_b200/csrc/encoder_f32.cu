// fp32 CUDA-core set-encoder kernels (generic over every MAB shape of the reference):
//   linear_f32_kernel : Y = act(X W^T + b) / Y = X + relu(X W^T + b)   (modules.py:20-21,31)
//   attn_f32_kernel   : O = Qp + softmax(Qp_h Kp_h^T / sqrt(D)) Vp_h    (modules.py:23-29)
//                       flash-style online softmax, split over keys when the query set is small
//                       (ISAB mab0 / PMA), one warp per head with broadcast K/V reads from smem.
//   attn_merge_kernel : merges the key splits.
//   layernorm_kernel  : ln0 / ln1 branches (modules.py:30,32).
//   pool_kernel       : DeepSet mean / max / sum over points (set_transformer-master/models.py:26).
// This is the 1e-3 (fp32) parity path and the only path for dims the tcgen05 kernels do not cover.
#include "common.cuh"
#include <float.h>
#include <math.h>

namespace pca {

// ------------------------------------------------------------------------------------ linear
constexpr int LBM = 128, LBN = 64, LBK = 16, LTHREADS = 256;

template <int MODE>   // 0 plain, 1 relu, 2 residual: Y = X + relu(XW^T + b) (din == dout), 3 = 2 that also stores relu(.) in R
__global__ void __launch_bounds__(LTHREADS)
linear_f32_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
                  float* __restrict__ Y, long long rows, int din, int dout, float* __restrict__ R) {
    __shared__ __align__(16) float Xs[LBK][LBM + 4];
    __shared__ __align__(16) float Ws[LBK][LBN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.x * LBM;
    const int col0 = blockIdx.y * LBN;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < din; k0 += LBK) {
        // X tile: 128 rows x 16 k  (8 elements / thread), W tile: 64 rows x 16 k (4 / thread)
        const int lk = tid & 15, lr = tid >> 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int m = lr + 16 * j;
            const long long r = row0 + m;
            const int k = k0 + lk;
            Xs[lk][m] = (r < rows && k < din) ? __ldg(X + r * din + k) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = lr + 16 * j;
            const int c = col0 + n;
            const int k = k0 + lk;
            Ws[lk][n] = (c < dout && k < din) ? __ldg(W + (long long)c * din + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < LBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&Xs[k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&Xs[k][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
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
        const long long r = row0 + ty * 8 + i;
        if (r >= rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = col0 + tx * 4 + j;
            if (c >= dout) continue;
            float v = acc[i][j] + __ldg(bias + c);
            if (MODE == 1) v = fmaxf(v, 0.f);
            if (MODE == 2) v = __ldg(X + r * din + c) + fmaxf(v, 0.f);
            if (MODE == 3) {
                const float rl = fmaxf(v, 0.f);
                R[r * dout + c] = rl;
                v = __ldg(X + r * din + c) + rl;
            }
            Y[r * dout + c] = v;
        }
    }
}

// Up to 32 rows (projections of the inducing points / seeds, heads on small batches): the 128-row tile kernel would run as
// 1 x dout/64 blocks walking din serially (measured 24 us for 16 x 256 x 256).  Here a block owns eight output columns: it
// stages the X rows (row stride din + 1: conflict-free for lane = row) and its eight W rows in shared memory with one round
// of coalesced loads, then lane = row, warp = column accumulates from shared memory (the W element is a broadcast).
template <int MODE>
__global__ void __launch_bounds__(256)
linear_skinny_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
                     float* __restrict__ Y, int rows, int din, int dout, float* __restrict__ R) {
    extern __shared__ float skinny_s[];
    float* Xs = skinny_s;                          // rows x (din + 1)
    float* Ws = skinny_s + rows * (din + 1);       // 8 x din
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 8, c = c0 + w;
    for (int i = threadIdx.x; i < rows * din; i += 256) {
        const int r = i / din, k = i - r * din;
        Xs[r * (din + 1) + k] = __ldg(X + i);
    }
    for (int i = threadIdx.x; i < 8 * din; i += 256) {
        const int cc = i / din, k = i - cc * din;
        Ws[i] = (c0 + cc < dout) ? __ldg(W + (long long)(c0 + cc) * din + k) : 0.f;
    }
    __syncthreads();
    if (c >= dout || lane >= rows) return;
    const float* xr = Xs + lane * (din + 1);
    const float* wc = Ws + w * din;
    float a0 = 0.f, a1 = 0.f;
    int k = 0;
    for (; k + 1 < din; k += 2) { a0 = fmaf(xr[k], wc[k], a0); a1 = fmaf(xr[k + 1], wc[k + 1], a1); }
    if (k < din) a0 = fmaf(xr[k], wc[k], a0);
    const int r = lane;
    float v = (a0 + a1) + __ldg(bias + c);
    if (MODE == 1) v = fmaxf(v, 0.f);
    if (MODE == 2) v = xr[c] + fmaxf(v, 0.f);
    if (MODE == 3) {
        const float rl = fmaxf(v, 0.f);
        R[(long long)r * dout + c] = rl;
        v = xr[c] + rl;
    }
    Y[(long long)r * dout + c] = v;
}

// Y (rows, dout) = X (rows, din) W^T + b for the first layers over raw points (din <= 4: (f, mag) / (f, t, mag) / xyz): the layer is
// bound by writing Y, so a thread produces four consecutive columns of one row from registers (its W rows and bias) and the row's
// din coordinates (a broadcast load), and a warp writes 512 contiguous bytes.  The 128x64 register-tiled GEMM spent a 16-deep K
// loop on 3 terms (0.13 ms per 256 000 x 3 -> 256 launch; this kernel is at the write bandwidth).
template <int DIN>
__global__ void __launch_bounds__(256) linear_tinyk_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ b,
                                                           float* __restrict__ Y, long long rows, int dout) {
    const int cq = dout >> 2;                                   // float4 column groups per row
    const int c4 = (blockIdx.y * blockDim.x + threadIdx.x) % cq;
    const int rsub = (blockIdx.y * blockDim.x + threadIdx.x) / cq;
    const int rstep = (gridDim.y * blockDim.x) / cq;
    if (rstep == 0 || rsub >= rstep) return;
    float w[4][DIN], bb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        bb[u] = __ldg(b + 4 * c4 + u);
#pragma unroll
        for (int k = 0; k < DIN; ++k) w[u][k] = __ldg(W + (long long)(4 * c4 + u) * DIN + k);
    }
    const long long r_end = min(rows, (long long)(blockIdx.x + 1) * 256);
    for (long long r = (long long)blockIdx.x * 256 + rsub; r < r_end; r += rstep) {
        float x[DIN];
#pragma unroll
        for (int k = 0; k < DIN; ++k) x[k] = __ldg(X + r * DIN + k);
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < DIN; ++k) a = fmaf(x[k], w[u][k], a);
            o[u] = a + bb[u];
        }
        *reinterpret_cast<float4*>(Y + r * dout + 4 * c4) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

int launch_linear(const float* X, const float* W, const float* b, float* Y, long long rows, int din,
                  int dout, int mode, cudaStream_t st, float* R, void* img, size_t img_bytes) {
    if (rows == 0) return 0;
    if (mode == 3 && !R) return fail(PCA_EINVAL, "linear: mode 3 needs the relu output buffer");
    if (mode >= 2 && din != dout) return fail(PCA_EINVAL, "linear: residual mode needs din == dout");
    if (img && img_bytes >= gemm_tc_image_bytes(dout, din) && linear_tc_eligible(rows, din, dout))
        return launch_linear_tc(X, W, 0, b, mode >= 2 ? X : nullptr, Y, mode == 3 ? R : nullptr, rows, din, dout, mode >= 1, img,
                                img_bytes, st);
    const size_t skinny_smem = ((size_t)rows * (din + 1) + 8 * (size_t)din) * sizeof(float);
    if (rows <= 32 && Y != X && skinny_smem <= 48 * 1024) {
        const unsigned g = (unsigned)((dout + 7) / 8);
        LaunchTimer lt("linear_skinny_kernel", st, 2.0 * rows * din * dout, 4.0 * ((double)rows * (din + dout) + (double)din * dout));
        if (mode == 0) linear_skinny_kernel<0><<<g, 256, skinny_smem, st>>>(X, W, b, Y, (int)rows, din, dout, nullptr);
        else if (mode == 1) linear_skinny_kernel<1><<<g, 256, skinny_smem, st>>>(X, W, b, Y, (int)rows, din, dout, nullptr);
        else if (mode == 2) linear_skinny_kernel<2><<<g, 256, skinny_smem, st>>>(X, W, b, Y, (int)rows, din, dout, nullptr);
        else linear_skinny_kernel<3><<<g, 256, skinny_smem, st>>>(X, W, b, Y, (int)rows, din, dout, R);
        PCA_CHECK_LAUNCH("linear_skinny_kernel");
        return 0;
    }
    if (mode == 0 && din >= 1 && din <= 4 && dout % 4 == 0 && rows >= 1024 && (dout / 4) <= 256 && 256 % (dout / 4) == 0) {
        // grid.x: 256-row slabs; one block covers (256 / (dout / 4)) rows per step with all column groups
        dim3 grid((unsigned)((rows + 255) / 256), 1);
        LaunchTimer lt("linear_tinyk_kernel", st, 2.0 * rows * din * dout, 4.0 * rows * (din + dout));
        switch (din) {
            case 1: linear_tinyk_kernel<1><<<grid, 256, 0, st>>>(X, W, b, Y, rows, dout); break;
            case 2: linear_tinyk_kernel<2><<<grid, 256, 0, st>>>(X, W, b, Y, rows, dout); break;
            case 3: linear_tinyk_kernel<3><<<grid, 256, 0, st>>>(X, W, b, Y, rows, dout); break;
            default: linear_tinyk_kernel<4><<<grid, 256, 0, st>>>(X, W, b, Y, rows, dout); break;
        }
        PCA_CHECK_LAUNCH("linear_tinyk_kernel");
        return 0;
    }
    dim3 grid((unsigned)((rows + LBM - 1) / LBM), (dout + LBN - 1) / LBN);
    {
        LaunchTimer lt("linear_f32_kernel", st, 2.0 * rows * din * dout, 4.0 * rows * (din + dout));
        if (mode == 0) linear_f32_kernel<0><<<grid, LTHREADS, 0, st>>>(X, W, b, Y, rows, din, dout, nullptr);
        else if (mode == 1) linear_f32_kernel<1><<<grid, LTHREADS, 0, st>>>(X, W, b, Y, rows, din, dout, nullptr);
        else if (mode == 2) linear_f32_kernel<2><<<grid, LTHREADS, 0, st>>>(X, W, b, Y, rows, din, dout, nullptr);
        else linear_f32_kernel<3><<<grid, LTHREADS, 0, st>>>(X, W, b, Y, rows, din, dout, R);
    }
    PCA_CHECK_LAUNCH("linear_f32_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------ attention
// Block = H warps (warp w <-> head w).  Lane <-> (query ql, key slice ks): TQ queries per block,
// KS = 32/TQ key slices.  K|V rows ([Kp (D) | Vp (D)] per key) are staged in smem tiles; all lanes
// of a warp with the same key slice read the same address (broadcast).
template <int DH, int R>
__global__ void attn_f32_kernel(const float* __restrict__ Qp, long long q_bstride,
                                const float* __restrict__ KV, int nq, int nk, int D, int tq_log,
                                int tk, int nsplit, int chunk, int q_tiles, float scale_log2e,
                                float* __restrict__ O, float* __restrict__ part,
                                const int* __restrict__ key_counts, float* __restrict__ lse) {
    extern __shared__ __align__(16) float kv_s[];     // tk rows x (2D + 4)
    const int H = blockDim.x >> 5;
    const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TQ = 1 << tq_log, KS = 32 >> tq_log;
    const int ql = lane & (TQ - 1), ks = lane >> tq_log;
    const int b = blockIdx.z, split = blockIdx.y;
    const int rs = 2 * D + 4;                          // padded smem row stride (floats)
    // Register blocking: a lane owns R queries (ql, ql + TQ, ...), so every K / V element read from shared memory feeds R
    // dot products -- these kernels are bound by the shared-memory broadcast reads (ncu: l1tex 84-96 % of peak at R = 1).
    // q_tiles > 1 only when the key set is a single smem tile: it is staged once and reused by every query tile of the block
  for (int qt = 0; qt < q_tiles; ++qt) {
    const int q0 = (blockIdx.x * q_tiles + qt) * TQ * R;
    if (q0 >= nq) break;                               // block-uniform
    int q[R];
    bool qvalid[R];
    const float* qptr[R];
    float qv[R][DH], acc[R][DH], m[R], l[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        q[r] = q0 + r * TQ + ql;
        qvalid[r] = q[r] < nq;
        qptr[r] = Qp + (long long)b * q_bstride + (long long)(qvalid[r] ? q[r] : 0) * D + h * DH;
        // per-thread rows: 16-byte accesses (a warp touches 32 different rows; scalar accesses would cost 4x the LSU wavefronts)
#pragma unroll
        for (int j = 0; j < DH; j += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(qptr[r] + j));
            qv[r][j] = t.x * scale_log2e; qv[r][j + 1] = t.y * scale_log2e; qv[r][j + 2] = t.z * scale_log2e; qv[r][j + 3] = t.w * scale_log2e;
            acc[r][j] = acc[r][j + 1] = acc[r][j + 2] = acc[r][j + 3] = 0.f;
        }
        m[r] = -INFINITY;
        l[r] = 0.f;
    }

    // variable-size sets: only the first key_counts[b] keys of the padded set take part
    const int nk_b = key_counts ? max(1, min(nk, __ldg(key_counts + b))) : nk;
    const int k_begin = split * chunk;
    const int k_end = min(nk_b, k_begin + chunk);
    const float* kvb = KV + (long long)b * nk * 2 * D;

    for (int kt = k_begin; kt < k_end; kt += tk) {
        const int tn = min(tk, k_end - kt);
        // cooperative, coalesced float4 load of tn rows x 2D floats
        const int vec_per_row = (2 * D) >> 2;
        if (qt == 0)
        for (int i = threadIdx.x; i < tn * vec_per_row; i += blockDim.x) {
            const int r = i / vec_per_row, c = i - r * vec_per_row;
            const float4 v = __ldg(reinterpret_cast<const float4*>(kvb + (long long)(kt + r) * 2 * D) + c);
            *reinterpret_cast<float4*>(kv_s + r * rs + c * 4) = v;
        }
        __syncthreads();
        for (int kk = ks; kk < tn; kk += 4 * KS) {
            // up to 4 keys of this lane's slice per softmax update
            float s[R][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int key = kk + u * KS;
                if (key < tn) {
                    const float* kr = kv_s + key * rs + h * DH;
                    float d[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) d[r] = 0.f;
#pragma unroll
                    for (int j = 0; j < DH; j += 4) {
                        const float4 kq = *reinterpret_cast<const float4*>(kr + j);
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            d[r] = fmaf(qv[r][j], kq.x, d[r]); d[r] = fmaf(qv[r][j + 1], kq.y, d[r]);
                            d[r] = fmaf(qv[r][j + 2], kq.z, d[r]); d[r] = fmaf(qv[r][j + 3], kq.w, d[r]);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) s[r][u] = d[r];
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) s[r][u] = -INFINITY;
                }
            }
            float p[R][4];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float mn = fmaxf(fmaxf(m[r], fmaxf(s[r][0], s[r][1])), fmaxf(s[r][2], s[r][3]));
                const float alpha = exp2f(m[r] - mn);      // m = -inf on the first update -> 0
                m[r] = mn;
                l[r] *= alpha;
#pragma unroll
                for (int j = 0; j < DH; ++j) acc[r][j] *= alpha;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    p[r][u] = (kk + u * KS < tn) ? exp2f(s[r][u] - mn) : 0.f;
                    l[r] += p[r][u];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int key = kk + u * KS;
                if (key < tn) {
                    const float* vr = kv_s + key * rs + D + h * DH;
#pragma unroll
                    for (int j = 0; j < DH; j += 4) {
                        const float4 vv = *reinterpret_cast<const float4*>(vr + j);
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            acc[r][j] = fmaf(p[r][u], vv.x, acc[r][j]); acc[r][j + 1] = fmaf(p[r][u], vv.y, acc[r][j + 1]);
                            acc[r][j + 2] = fmaf(p[r][u], vv.z, acc[r][j + 2]); acc[r][j + 3] = fmaf(p[r][u], vv.w, acc[r][j + 3]);
                        }
                    }
                }
            }
        }
        if (q_tiles == 1) __syncthreads();
    }

    // merge the key slices of one query across lanes (xor over the slice bits)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        for (int off = TQ; off < 32; off <<= 1) {
            const float mo = __shfl_xor_sync(0xffffffffu, m[r], off);
            const float lo = __shfl_xor_sync(0xffffffffu, l[r], off);
            const float mn = fmaxf(m[r], mo);
            const float a0 = (m[r] == -INFINITY) ? 0.f : exp2f(m[r] - mn);
            const float a1 = (mo == -INFINITY) ? 0.f : exp2f(mo - mn);
            l[r] = l[r] * a0 + lo * a1;
#pragma unroll
            for (int j = 0; j < DH; ++j) {
                const float ao = __shfl_xor_sync(0xffffffffu, acc[r][j], off);
                acc[r][j] = acc[r][j] * a0 + ao * a1;
            }
            m[r] = mn;
        }
    }
    if (ks != 0) continue;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (!qvalid[r]) continue;
        if (nsplit == 1) {
            const float inv = 1.f / l[r];
            float* o = O + ((long long)b * nq + q[r]) * D + h * DH;
#pragma unroll
            for (int j = 0; j < DH; j += 4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(qptr[r] + j));
                *reinterpret_cast<float4*>(o + j) = make_float4(t.x + acc[r][j] * inv, t.y + acc[r][j + 1] * inv,
                                                                t.z + acc[r][j + 2] * inv, t.w + acc[r][j + 3] * inv);
            }
            if (lse) lse[((long long)b * nq + q[r]) * H + h] = m[r] + log2f(l[r]);      // log2 of sum_k 2^(s_k), scaled log2 domain
        } else {
            float* pp = part + ((((long long)b * nsplit + split) * nq + q[r]) * H + h) * (DH + 2);
            pp[0] = m[r]; pp[1] = l[r];
#pragma unroll
            for (int j = 0; j < DH; ++j) pp[2 + j] = acc[r][j];
        }
    }
  }
}

// One thread per (batch, query, head, head dim): the DH threads of one (b, q, h) read its partial rows side by side
// (coalesced), each re-deriving the row's maximum and normaliser from the nsplit (m, l) pairs (broadcast loads).  Same
// operations in the same order as a per-row loop, so the result does not depend on the thread layout.
template <int DH>
__global__ void attn_merge_kernel(const float* __restrict__ part, const float* __restrict__ Qp,
                                  long long q_bstride, int B, int nq, int H, int nsplit,
                                  float* __restrict__ O, float* __restrict__ lse) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * nq * H;
    const long long i = t / DH;
    const int j = (int)(t - i * DH);
    if (i >= total) return;
    const int h = (int)(i % H);
    const int q = (int)((i / H) % nq);
    const int b = (int)(i / ((long long)H * nq));
    const long long sstride = (long long)nq * H * (DH + 2);
    const float* p0 = part + (((long long)b * nsplit * nq + q) * H + h) * (DH + 2);
    float m = -INFINITY;
    for (int s = 0; s < nsplit; ++s) m = fmaxf(m, __ldg(p0 + s * sstride));
    float l = 0.f, acc = 0.f;
    for (int s = 0; s < nsplit; ++s) {
        const float* pp = p0 + s * sstride;
        const float pm = __ldg(pp);
        const float a = (pm == -INFINITY) ? 0.f : exp2f(pm - m);
        l += __ldg(pp + 1) * a;
        acc = fmaf(__ldg(pp + 2 + j), a, acc);
    }
    const float inv = 1.f / l;
    const int D = H * DH;
    O[((long long)b * nq + q) * D + h * DH + j] = __ldg(Qp + (long long)b * q_bstride + (long long)q * D + h * DH + j) + acc * inv;
    if (lse && j == 0) lse[i] = m + log2f(l);
}

struct AttnPlan { int tq_log, tk, nsplit, chunk; size_t smem; size_t part_floats; };

static AttnPlan plan_attn(int B, int nq, int nk, int D, int H) {
    AttnPlan p;
    int tq = 32;
    while (tq > 1 && (tq >> 1) >= nq) tq >>= 1;
    p.tq_log = 0;
    while ((1 << p.tq_log) < tq) ++p.tq_log;
    const int rs = 2 * D + 4;
    int tk = (40 * 1024) / (rs * 4);
    tk = tk > 128 ? 128 : tk;
    tk &= ~3;
    if (tk < 4) tk = 4;
    if (tk > nk) tk = (nk + 3) & ~3;
    p.tk = tk;
    p.smem = (size_t)tk * rs * 4;
    const long long qtiles = (nq + tq - 1) / tq;
    const long long base_blocks = (long long)B * qtiles;
    int nsplit = 1;
    const long long target = 148LL * 4;
    if (base_blocks < target) {
        nsplit = (int)((target + base_blocks - 1) / base_blocks);
        const int max_split = (nk + 4 * tk - 1) / (4 * tk);      // keep >= 4 tiles per split
        if (nsplit > max_split) nsplit = max_split;
        if (nsplit < 1) nsplit = 1;
    }
    int chunk = (nk + nsplit - 1) / nsplit;
    chunk = (chunk + tk - 1) / tk * tk;
    nsplit = (nk + chunk - 1) / chunk;
    p.nsplit = nsplit;
    p.chunk = chunk;
    const int dh = D / H;
    p.part_floats = nsplit > 1 ? (size_t)B * nsplit * nq * H * (dh + 2) : 0;
    return p;
}

// scratch of launch_attn: the split partials of the CUDA-core kernel or the score matrix + operand images of the tensor-core route
size_t attn_part_floats(int B, int nq, int nk, int D, int H) {
    const size_t a = plan_attn(B, nq, nk, D, H).part_floats, b = attn_tc_fwd_floats(B, nq, nk, D, H);
    return a > b ? a : b;
}

template <int DH, int R>
static int launch_attn_r(const float* Qp, long long q_bstride, const float* KV, int B, int nq, int nk,
                         int D, int H, float* O, float* part, const int* key_counts, cudaStream_t st, float* lse) {
    const AttnPlan p = plan_attn(B, nq, nk, D, H);
    const int tq = 1 << p.tq_log;
    // key set in one smem tile: stage it once per block and walk several query tiles (keeps >= ~8 blocks per SM)
    const int q_blocks = (nq + tq * R - 1) / (tq * R);
    int q_tiles = 1;
    if (p.nsplit == 1 && nk <= p.tk && key_counts == nullptr) {
        const long long qt = ((long long)q_blocks * B) / (148LL * 8);
        q_tiles = (int)(qt < 1 ? 1 : (qt > 16 ? 16 : qt));
    }
    dim3 grid((q_blocks + q_tiles - 1) / q_tiles, p.nsplit, B);
    const float scale_log2e = (1.0f / sqrtf((float)D)) * 1.4426950408889634f;
    if (p.smem > 48 * 1024)
        PCA_CHECK_CUDA((cudaFuncSetAttribute(attn_f32_kernel<DH, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem)));
    {
        LaunchTimer lt("attn_f32_kernel", st, 4.0 * B * nq * (double)nk * D,
                       4.0 * ((double)B * nk * 2 * D + 2.0 * B * nq * D));
        attn_f32_kernel<DH, R><<<grid, 32 * H, p.smem, st>>>(Qp, q_bstride, KV, nq, nk, D, p.tq_log, p.tk,
                                                             p.nsplit, p.chunk, q_tiles, scale_log2e, O, part, key_counts, lse);
    }
    PCA_CHECK_LAUNCH("attn_f32_kernel");
    if (p.nsplit > 1) {
        const long long total = (long long)B * nq * H;
        {
            LaunchTimer lt("attn_merge_kernel", st, 0.0, 4.0 * (double)p.part_floats);
            attn_merge_kernel<DH><<<(unsigned)((total * DH + 255) / 256), 256, 0, st>>>(part, Qp, q_bstride, B, nq, H, p.nsplit, O, lse);
        }
        PCA_CHECK_LAUNCH("attn_merge_kernel");
    }
    return 0;
}

// RMAX queries per lane when the query set is large enough to keep every lane busy (and the grid full), else one
template <int DH, int RMAX>
static int launch_attn_t(const float* Qp, long long q_bstride, const float* KV, int B, int nq, int nk,
                         int D, int H, float* O, float* part, const int* key_counts, cudaStream_t st, float* lse) {
    if (RMAX > 1 && nq >= 32 * RMAX && (long long)B * (nq / (32 * RMAX)) >= 148LL * 4)
        return launch_attn_r<DH, RMAX>(Qp, q_bstride, KV, B, nq, nk, D, H, O, part, key_counts, st, lse);
    return launch_attn_r<DH, 1>(Qp, q_bstride, KV, B, nq, nk, D, H, O, part, key_counts, st, lse);
}

int launch_attn(const float* Qp, long long q_bstride, const float* KV, int B, int nq, int nk, int D,
                int H, float* O, float* part, const int* key_counts, cudaStream_t st, float* lse, float* p_out) {
    if (B == 0 || nq == 0) return 0;
    if (H < 1 || H > 32 || D % H) return fail(PCA_EUNSUPPORTED, "attention: need 1 <= H <= 32 and D %% H == 0 (D=%d, H=%d)", D, H);
    if (B > 65535) return fail(PCA_EUNSUPPORTED, "attention: batch chunk %d exceeds the grid limit", B);
    // one small side (inducing points / seeds against a large set): Q K^T and P V as split-bf16 tensor-core GEMMs
    // (variable-size sets: the key counts only arise where the points are the keys, i.e. in the small-query form)
    if (attn_tc_eligible(B, nq, nk, D, H) && (!key_counts || attn_tc_kind(B, nq, nk, D, H) == 2))
        return launch_attn_tc(Qp, q_bstride, KV, B, nq, nk, D, H, O, part, st, lse, key_counts, p_out);
    switch (D / H) {
        case 4: return launch_attn_t<4, 2>(Qp, q_bstride, KV, B, nq, nk, D, H, O, part, key_counts, st, lse);
        case 8: return launch_attn_t<8, 2>(Qp, q_bstride, KV, B, nq, nk, D, H, O, part, key_counts, st, lse);
        case 16: return launch_attn_t<16, 1>(Qp, q_bstride, KV, B, nq, nk, D, H, O, part, key_counts, st, lse);
        case 32: return launch_attn_t<32, 1>(Qp, q_bstride, KV, B, nq, nk, D, H, O, part, key_counts, st, lse);
        case 64: return launch_attn_t<64, 1>(Qp, q_bstride, KV, B, nq, nk, D, H, O, part, key_counts, st, lse);
        default: return fail(PCA_EUNSUPPORTED, "attention: head dim %d not in {4,8,16,32,64}", D / H);
    }
}

// ------------------------------------------------------------------------------------ layernorm
__global__ void layernorm_kernel(float* __restrict__ X, long long rows, int D, const float* __restrict__ g,
                                 const float* __restrict__ bta) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    float* x = X + r * D;
    float s = 0.f;
    for (int j = lane; j < D; j += 32) s += x[j];
    const float mean = warp_sum(s) / D;
    float v = 0.f;
    for (int j = lane; j < D; j += 32) { const float d = x[j] - mean; v += d * d; }
    const float rstd = rsqrtf(warp_sum(v) / D + 1e-5f);
    for (int j = lane; j < D; j += 32) x[j] = (x[j] - mean) * rstd * __ldg(g + j) + __ldg(bta + j);
}

int launch_layernorm(float* X, long long rows, int D, const float* g, const float* b, cudaStream_t st) {
    if (rows == 0) return 0;
    LaunchTimer lt("layernorm_kernel", st, 0.0, 8.0 * rows * D);
    layernorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(X, rows, D, g, b);
    PCA_CHECK_LAUNCH("layernorm_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------ pooling
// X (B, N, D) -> out (B, D); pool 0 mean, 1 max, 2 sum
__global__ void pool_kernel(const float* __restrict__ X, int N, int D, int pool, float* __restrict__ out,
                            const int* __restrict__ counts) {
    __shared__ float red[8][33];
    const int b = blockIdx.y;
    const int Nb = counts ? max(1, min(N, __ldg(counts + b))) : N;      // masked pooling over the first counts[b] points
    const int d = blockIdx.x * 32 + threadIdx.x;
    const int ty = threadIdx.y;
    float v = pool == 1 ? -INFINITY : 0.f;
    if (d < D) {
        const float* x = X + (long long)b * N * D + d;
        for (int p = ty; p < Nb; p += 8) {
            const float t = __ldg(x + (long long)p * D);
            v = pool == 1 ? fmaxf(v, t) : v + t;
        }
    }
    red[ty][threadIdx.x] = v;
    __syncthreads();
    if (ty == 0 && d < D) {
        for (int i = 1; i < 8; ++i) v = pool == 1 ? fmaxf(v, red[i][threadIdx.x]) : v + red[i][threadIdx.x];
        if (pool == 0) v /= (float)Nb;
        out[(long long)b * D + d] = v;
    }
}

int launch_pool(const float* X, int B, int N, int D, int pool, float* out, const int* counts, cudaStream_t st) {
    if (B == 0) return 0;
    if (pool < 0 || pool > 2) return fail(PCA_EINVAL, "pool: mode %d not in {0 mean, 1 max, 2 sum}", pool);
    if (B > 65535) return fail(PCA_EUNSUPPORTED, "pool: batch too large");
    dim3 grid((D + 31) / 32, B), block(32, 8);
    LaunchTimer lt("pool_kernel", st, 0.0, 4.0 * (double)B * N * D);
    pool_kernel<<<grid, block, 0, st>>>(X, N, D, pool, out, counts);
    PCA_CHECK_LAUNCH("pool_kernel");
    return 0;
}

}  // namespace pca
