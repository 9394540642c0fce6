// tcgen05 / TMEM set-encoder path for the audio Set Transformer dims (D=64, H=8 -> head dim 8, M=64
// inducing points, one PMA seed, no LayerNorm; ST of Code/models.py:13-44 with the hyper-parameters of
// Code/settransformer.py:81-85).  bf16 operands, fp32 accumulation in TMEM, fp32 softmax statistics.
//
//   prep_kernel             : batch-independent work hoisted out of the per-cloud path: fc_q(I), fc_q(S) (the I.repeat /
//                             S.repeat of modules.py:52,63 is never materialised), the bf16 UMMA operand images of the
//                             weights (plain, split hi/lo, and the pooled-attention query image).
//   mab_reduce5_tc_kernel   : MAB(Q = inducing points, K = points) (ISAB mab0).  Rows of the 128-row MMA tile are
//                             (head-of-pair, query); two heads are stacked per MMA so the K=16 bf16 instruction depth is
//                             fully used by 2 x head-dim 8.  Softmax over the points against a fixed per-row reference
//                             exponent, outputs accumulating in TMEM.
//   finalize_isab_tc_kernel : merges the point splits, O = Qp + A V, H = O + relu(fc_o(O)), then the K/V projections of H
//                             for mab1 (split-bf16 MMAs), emitted as block-diagonal bf16 operand images.
//   mab_apply3_tc_kernel    : MAB(Q = points, K = H) (ISAB mab1): per 128-point tile, Q projection MMA, QK^T (one head per
//                             MMA via the block-diagonal K image), softmax over the 64 keys in registers, P V accumulating
//                             onto the Q projection, fc_o + ReLU residual by dedicated epilogue warps.
//   pma_pool_tc_kernel      : PMA with one seed on the un-projected points; finalize_pool_kernel applies fc_v, the MAB
//                             tail and the final Linear -> logits.
#include "common.cuh"
#include "tc_prims.cuh"
#include "tma_host.cuh"
#include <stdlib.h>

namespace pca {
using namespace tc;

constexpr int TD = 64, TH = 8, TM = 64;                   // dims this path is specialised for
constexpr float kScaleLog2e = 1.4426950408889634f / 8.0f;  // log2(e) / sqrt(dim_V)
// mab1: the scale is folded into the per-cloud K image written by finalize_isab_tc_kernel (keys are 64 per cloud, queries are
// all points), so the query operand is the plain bf16 projection
constexpr float kKImageScale = kScaleLog2e;

struct TcConsts {
    float Qp0[TM * TD];        // isab0.mab0 fc_q(I)
    float Qp1[TM * TD];        // isab1.mab0 fc_q(I)
    float QpS[TD];             // pma fc_q(S)
    float pad[64];
    uint8_t Aq0[16384], Aq1[16384], AqP[16384];   // stacked-pair query operands [4 pairs][2 chunks][128 rows][16 B]
    uint8_t Wkv1[16384], WkvP[16384];             // [Wk;Wv] as B operand (N=128, K=64): [8 chunks][128 rows][16 B]
    uint8_t Wq1[8192], Wo0[8192], Wo1[8192];      // (N=64, K=64) B operands: [8 chunks][64 rows][16 B]
    // split-bf16 (hi | lo) B operands of the finalize GEMMs, per ISAB: mab0.fc_o (N=64, K=64) and mab1 [Wk;Wv] (N=128, K=64)
    uint8_t WoS[2][2][8192];
    uint8_t WkvS[2][2][16384];
    // k-major fp32 copies for the tail paths (leftover points of a cloud, N mod 128 <= TC_TAIL_MAX, handled on CUDA cores):
    float Wkv0T[2][64 * 128];          // per ISAB: mab0 [Wk;Wv]^T (dk <= 64 rows used, 128 columns)
    float Wq1T[2][64 * 64];            // per ISAB: mab1 fc_q^T (dq rows used, 64 columns)
    float Wo1T[2][64 * 64];            // per ISAB: mab1 fc_o^T
    float WqkPool[TH * TD];            // PMA: scale * Wk_h^T fc_q(S)_h (fp32 twin of AqPool)
    float WvT_P[64 * 64], WoT_P[64 * 64];          // pma.mab fc_v / fc_o transposed (k, f): coalesced reads in finalize_pool_kernel
    uint8_t AqPool[16384];                        // PMA: row r = scale * Wk_h^T fc_q(S)_h, h = r / 16 (A operand, 128 x 64)
    uint8_t WqS0[2048];                           // isab0.mab1.fc_q (64, d_in <= 4) as a split-bf16 K=16 B operand
    // apply4: the same B operands with the bias as an extra K step (rows k = 64, 65 = bias hi, lo; against the "ones" A operand)
    uint8_t Wq1e[10240], Wo0e[10240], Wo1e[10240];
    uint8_t WqS0e[2048];                          // WqS0 with the bias (hi, lo) in the spare columns k = 12, 13
    // reduce6: mab0 query operands with W_k folded in (rows = (head of pair, query), K = input features), softmax scale included
    uint8_t Gq1[65536];                           // isab1: [4 pairs][8 chunks][128 rows][16 B]
    uint8_t Gq0s[16384];                          // isab0 (d_in <= 4): split-bf16 K = 16 image [4 pairs][2 chunks][128 rows][16 B]
    uint8_t Wv1e[10240];                          // isab1.mab0 fc_v as (N = 64, K = 80) B operand with the bias step
    // k-major fp32 copies for finalize_isab: mab0.fc_o^T (64 x 64) and mab1 [Wk;Wv]^T (64 x 128), per ISAB
    float WoT[2][64 * 64];
    float WkvT[2][64 * 128];
};

__device__ void transpose_weight(const float* __restrict__ W, int n_rows, float* __restrict__ out) {
    // W (n_rows, 64) -> out (64, n_rows)
    for (int i = threadIdx.x; i < n_rows * 64; i += blockDim.x) {
        const int n = i / 64, k = i % 64;
        out[k * n_rows + n] = W[i];
    }
}

__device__ void transpose_generic(const float* __restrict__ W, int n_rows, int n_cols, float* __restrict__ out) {
    // W (n_rows, n_cols) -> out (n_cols, n_rows)
    for (int i = threadIdx.x; i < n_rows * n_cols; i += blockDim.x) {
        const int n = i / n_cols, k = i % n_cols;
        out[k * n_rows + n] = W[i];
    }
}

// ------------------------------------------------------------------------------------ prep
__device__ void pack_b_operand(const float* __restrict__ W, int n_rows, uint8_t* __restrict__ out) {
    // W (n_rows, 64) row-major fp32 -> bf16 [k/8][n][k%8]
    for (int i = threadIdx.x; i < n_rows * 64; i += blockDim.x) {
        const int n = i / 64, k = i % 64;
        *reinterpret_cast<__nv_bfloat16*>(out + (k / 8) * (n_rows * 16) + n * 16 + (k % 8) * 2) = __float2bfloat16(W[i]);
    }
}

// Split-bf16 image of a narrow weight W (n_rows, d <= 4) as a K = 16 B operand [2 chunks][n_rows][16 B].  With the
// matching A-side columns (split_x16) one K=16 MMA yields x.w to ~2^-17 relative (j = k % 4 is the input column):
//   k in 0..3 : x_hi * w_hi     k in 4..7 : x_lo * w_hi     k in 8..11 : x_hi * w_lo     k in 12..15 : 0
__device__ void pack_split_b_operand(const float* __restrict__ W, int n_rows, int d, uint8_t* __restrict__ out) {
    for (int i = threadIdx.x; i < n_rows * 16; i += blockDim.x) {
        const int n = i / 16, k = i % 16, j = k & 3;
        float v = 0.f;
        if (k < 12 && j < d) {
            const float w = W[n * d + j];
            const float hi = __bfloat162float(__float2bfloat16(w));
            v = (k < 8) ? hi : (w - hi);
        }
        *reinterpret_cast<__nv_bfloat16*>(out + (k / 8) * (n_rows * 16) + n * 16 + (k % 8) * 2) = __float2bfloat16(v);
    }
}
// A-side columns of the split product for one row x[0..4) (entries past d_in are zero)
__device__ __forceinline__ void split_x16(const float* x, float* cols) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float hi = __bfloat162float(__float2bfloat16(x[j]));
        cols[j] = hi;
        cols[4 + j] = x[j] - hi;
        cols[8 + j] = hi;
        cols[12 + j] = 0.f;
    }
}

// (N = 64, K = 80) B operand: the 64 x 64 weight followed by one K step whose rows k = 64, 65 hold the bias as bf16 hi, lo
__device__ void pack_b_operand_bias(const float* __restrict__ W, const float* __restrict__ b, uint8_t* __restrict__ out) {
    pack_b_operand(W, 64, out);
    for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) {
        const int n = i / 16, k = i % 16;
        const float hi = __bfloat162float(__float2bfloat16(b[n]));
        const float v = k == 0 ? hi : (k == 1 ? b[n] - hi : 0.f);
        *reinterpret_cast<__nv_bfloat16*>(out + 8192 + (k / 8) * 1024 + n * 16 + (k % 8) * 2) = __float2bfloat16(v);
    }
}
// pack_split_b_operand plus the bias (hi, lo) in columns k = 12, 13
__device__ void pack_split_b_operand_bias(const float* __restrict__ W, const float* __restrict__ b, int n_rows, int d, uint8_t* __restrict__ out) {
    pack_split_b_operand(W, n_rows, d, out);
    __syncthreads();
    for (int n = threadIdx.x; n < n_rows; n += blockDim.x) {
        const float hi = __bfloat162float(__float2bfloat16(b[n]));
        *reinterpret_cast<__nv_bfloat16*>(out + (n_rows * 16) + n * 16 + 4 * 2) = __float2bfloat16(hi);          // k = 12
        *reinterpret_cast<__nv_bfloat16*>(out + (n_rows * 16) + n * 16 + 5 * 2) = __float2bfloat16(b[n] - hi);   // k = 13
    }
}

// W (n_rows, 64) -> two bf16 B-operand images: hi = bf16(W), lo = bf16(W - hi)
__device__ void pack_b_operand_split(const float* __restrict__ W, int n_rows, uint8_t* __restrict__ hi, uint8_t* __restrict__ lo) {
    for (int i = threadIdx.x; i < n_rows * 64; i += blockDim.x) {
        const int n = i / 64, k = i % 64;
        const size_t off = (size_t)(k / 8) * (n_rows * 16) + n * 16 + (k % 8) * 2;
        const __nv_bfloat16 h = __float2bfloat16(W[i]);
        *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(lo + off) = __float2bfloat16(W[i] - __bfloat162float(h));
    }
}

// Qp = fc_q(Qin) for nq (64 or 1) queries, and the stacked-pair A operand image
__device__ void prep_queries(const float* __restrict__ Qin, int nq, const float* __restrict__ Wq,
                             const float* __restrict__ bq, float* __restrict__ Qp_out, uint8_t* __restrict__ Aq,
                             float* sq /* smem 64*64 */, const float* __restrict__ Wk = nullptr,
                             uint8_t* __restrict__ AqPool = nullptr, float* __restrict__ WqkPool = nullptr,
                             uint8_t* __restrict__ Gq = nullptr, uint8_t* __restrict__ GqSplit = nullptr, int dk = TD) {
    {   // Qp = Qin Wq^T + bq from shared-memory copies (Wq rows padded to 65 floats: thread <-> output feature is conflict free)
        __shared__ float sW[TD * 65];
        float* sI = sq;                                  // the inputs are staged in sq, then replaced by the outputs
        for (int i = threadIdx.x; i < nq * TD; i += blockDim.x) sI[i] = Qin[i];
        for (int i = threadIdx.x; i < TD * TD; i += blockDim.x) sW[(i / TD) * 65 + (i % TD)] = Wq[i];
        __syncthreads();
        float outv[(TM * TD) / 256];
#pragma unroll
        for (int j = 0; j < (TM * TD) / 256; ++j) {
            const int i = threadIdx.x + j * 256;
            if (i < nq * TD) {
                const int m = i / TD, f = i % TD;
                float a = bq[f];
#pragma unroll 8
                for (int k = 0; k < TD; ++k) a = fmaf(sI[m * TD + k], sW[f * 65 + k], a);
                outv[j] = a;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < (TM * TD) / 256; ++j) {
            const int i = threadIdx.x + j * 256;
            if (i < nq * TD) { sq[i] = outv[j]; Qp_out[i] = outv[j]; }
        }
    }
    __syncthreads();
    if (nq == 1 && AqPool != nullptr) {
        // Pooling by one seed (modules.py:62-63) folded through fc_k: the score of head h against point n is
        //   q_h . (Wk_h y_n + bk_h) = (Wk_h^T q_h) . y_n + const(h)   and the constant cancels in the softmax over n,
        // so the scores are a plain product of the un-projected points with one 64-vector per head.
        // sq[64..64+512) = wqk[h][f] = scale * sum_d q[h*8+d] Wk[h*8+d][f]
        for (int i = threadIdx.x; i < TH * TD; i += blockDim.x) {
            const int h = i / TD, f = i % TD;
            float a = 0.f;
            for (int d = 0; d < 8; ++d) a = fmaf(sq[h * 8 + d], Wk[(h * 8 + d) * TD + f], a);
            sq[64 + i] = a * kScaleLog2e;
            if (WqkPool != nullptr) WqkPool[i] = a * kScaleLog2e;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 8 * 128 * 8; i += blockDim.x) {
            const int d = i & 7, r = (i >> 3) & 127, c = i >> 10;
            *reinterpret_cast<__nv_bfloat16*>(AqPool + c * 2048 + r * 16 + d * 2) = __float2bfloat16(sq[64 + (r >> 4) * TD + c * 8 + d]);
        }
    }
    if (nq == 1) {
        // PMA all-heads image (128 x 64, [8 chunks][128 rows][16 B]): row r carries head r / 16 in chunk r / 16
        for (int i = threadIdx.x; i < 8 * 128 * 8; i += blockDim.x) {
            const int d = i & 7, r = (i >> 3) & 127, c = i >> 10;
            const float v = ((r >> 4) == c) ? sq[c * 8 + d] * kScaleLog2e : 0.f;
            *reinterpret_cast<__nv_bfloat16*>(Aq + c * 2048 + r * 16 + d * 2) = __float2bfloat16(v);
        }
        return;
    }
    // reduce6: W_k folded into the queries, G[(h, m)][f] = scale * sum_d Qp[m][8h + d] Wk[8h + d][f]   (Wk: (64, dk))
    if (Gq != nullptr) {
        for (int i = threadIdx.x; i < 4 * 8 * 128 * 8; i += blockDim.x) {
            const int d = i & 7, r = (i >> 3) & 127, c = (i >> 10) & 7, p = i >> 13;
            const int h = 2 * p + (r >> 6), m = r & 63, f = c * 8 + d;
            float a = 0.f;
#pragma unroll
            for (int dd = 0; dd < 8; ++dd) a = fmaf(sq[m * TD + h * 8 + dd], Wk[(h * 8 + dd) * TD + f], a);
            *reinterpret_cast<__nv_bfloat16*>(Gq + p * 16384 + c * 2048 + r * 16 + d * 2) = __float2bfloat16(a * kScaleLog2e);
        }
    }
    if (GqSplit != nullptr) {
        // K = 16 split image against split_x16 columns: k 0..3 G_hi (x_hi), 4..7 G_hi (x_lo), 8..11 G_lo (x_hi), 12..15 zero
        for (int i = threadIdx.x; i < 4 * 2 * 128 * 8; i += blockDim.x) {
            const int d = i & 7, r = (i >> 3) & 127, c = (i >> 10) & 1, p = i >> 11;
            const int h = 2 * p + (r >> 6), m = r & 63, k = c * 8 + d, j = k & 3;
            float v = 0.f;
            if (k < 12 && j < dk) {
                float a = 0.f;
#pragma unroll
                for (int dd = 0; dd < 8; ++dd) a = fmaf(sq[m * TD + h * 8 + dd], Wk[(h * 8 + dd) * dk + j], a);
                a *= kScaleLog2e;
                const float hi = __bfloat162float(__float2bfloat16(a));
                v = (k < 8) ? hi : (a - hi);
            }
            *reinterpret_cast<__nv_bfloat16*>(GqSplit + p * 4096 + c * 2048 + r * 16 + d * 2) = __float2bfloat16(v);
        }
    }
    // Aq[p][c][r][d]: rows 0-63 carry head 2p in chunk 0, rows 64-127 carry head 2p+1 in chunk 1
    for (int i = threadIdx.x; i < 4 * 2 * 128 * 8; i += blockDim.x) {
        const int d = i & 7, r = (i >> 3) & 127, c = (i >> 10) & 1, p = i >> 11;
        const float v = ((r >> 6) == c) ? sq[(r & 63) * TD + (2 * p + c) * 8 + d] * kScaleLog2e : 0.f;
        *reinterpret_cast<__nv_bfloat16*>(Aq + p * 4096 + c * 2048 + r * 16 + d * 2) = __float2bfloat16(v);
    }
}

__global__ void prep_kernel(const float* __restrict__ params, int d_in, TcConsts* __restrict__ c) {
    __shared__ float sq[TM * TD];
    const long long n_isab0 = (long long)TM * TD + mab_count(TD, d_in, TD, 0) + mab_count(d_in, TD, TD, 0);
    const float* p_isab0 = params;
    const float* p_isab1 = p_isab0 + n_isab0;
    const float* p_pma = p_isab1 + (long long)TM * TD + 2 * mab_count(TD, TD, TD, 0);
    const MabParams m00 = mab_slice(p_isab0 + TM * TD, TD, d_in, TD, 0);
    const MabParams m01 = mab_slice(p_isab0 + TM * TD + mab_count(TD, d_in, TD, 0), d_in, TD, TD, 0);
    const MabParams m10 = mab_slice(p_isab1 + TM * TD, TD, TD, TD, 0);
    const MabParams m11 = mab_slice(p_isab1 + TM * TD + mab_count(TD, TD, TD, 0), TD, TD, TD, 0);
    const MabParams mp = mab_slice(p_pma + TD, TD, TD, TD, 0);
    switch (blockIdx.x) {
        case 0: prep_queries(p_isab0, TM, m00.Wq, m00.bq, c->Qp0, c->Aq0, sq, m00.Wkv, nullptr, nullptr, nullptr, c->Gq0s, d_in); break;
        case 1: prep_queries(p_isab1, TM, m10.Wq, m10.bq, c->Qp1, c->Aq1, sq, m10.Wkv, nullptr, nullptr, c->Gq1, nullptr, TD); break;
        case 2: prep_queries(p_pma, 1, mp.Wq, mp.bq, c->QpS, c->AqP, sq, mp.Wkv, c->AqPool, c->WqkPool); break;
        case 3: pack_b_operand(m10.Wkv, 128, c->Wkv1); break;
        case 4: pack_b_operand(mp.Wkv, 128, c->WkvP); break;
        case 5: pack_b_operand(m11.Wq, 64, c->Wq1); break;
        case 6: pack_b_operand(m01.Wo, 64, c->Wo0); break;
        case 7: pack_b_operand(m11.Wo, 64, c->Wo1); break;
        case 8: transpose_weight(m00.Wo, 64, c->WoT[0]); break;
        case 9: transpose_weight(m01.Wkv, 128, c->WkvT[0]); break;
        case 10: transpose_weight(m10.Wo, 64, c->WoT[1]); break;
        case 11: transpose_weight(m11.Wkv, 128, c->WkvT[1]); break;
        case 12: pack_split_b_operand(m01.Wq, 64, d_in, c->WqS0); break;
        case 13: pack_b_operand_split(m00.Wo, 64, c->WoS[0][0], c->WoS[0][1]); break;
        case 14: pack_b_operand_split(m01.Wkv, 128, c->WkvS[0][0], c->WkvS[0][1]); break;
        case 15: pack_b_operand_split(m10.Wo, 64, c->WoS[1][0], c->WoS[1][1]); break;
        case 16: pack_b_operand_split(m11.Wkv, 128, c->WkvS[1][0], c->WkvS[1][1]); break;
        case 19: transpose_generic(m00.Wkv, 128, d_in, c->Wkv0T[0]); break;
        case 20: transpose_generic(m10.Wkv, 128, TD, c->Wkv0T[1]); break;
        case 21: transpose_generic(m01.Wq, 64, d_in, c->Wq1T[0]); break;
        case 22: transpose_generic(m11.Wq, 64, TD, c->Wq1T[1]); break;
        case 23: transpose_generic(m01.Wo, 64, TD, c->Wo1T[0]); break;
        case 24: transpose_generic(m11.Wo, 64, TD, c->Wo1T[1]); break;
        case 25: pack_b_operand_bias(m11.Wq, m11.bq, c->Wq1e); break;
        case 26: pack_b_operand_bias(m01.Wo, m01.bo, c->Wo0e); break;
        case 27: pack_b_operand_bias(m11.Wo, m11.bo, c->Wo1e); break;
        case 28: pack_split_b_operand_bias(m01.Wq, m01.bq, 64, d_in, c->WqS0e); break;
        case 29: pack_b_operand_bias(m10.Wkv + TD * TD, m10.bkv + TD, c->Wv1e); break;
        case 17: transpose_weight(mp.Wkv + TD * TD, 64, c->WvT_P); break;
        case 18: transpose_weight(mp.Wo, 64, c->WoT_P); break;
        default: break;
    }
}

// ------------------------------------------------------------------------------------ shared helpers
__device__ __forceinline__ void copy_to_smem(uint8_t* dst, const uint8_t* __restrict__ src, int bytes) {
    for (int i = threadIdx.x * 16; i < bytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4*>(dst + i) = __ldg(reinterpret_cast<const uint4*>(src + i));
}
__device__ __forceinline__ void st_shared_8bf16(uint8_t* dst, const float* v) {
    uint4 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
    u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst) = u;
}

// ------------------------------------------------------------------------------------ where the points come from
// The d_in <= 4 points of a cloud are either rows of an explicit (B, N, d_in) tensor (X32) or -- on the whole-path call without
// a selection step -- read straight from the front end's output: point n = t * nf + f of a cloud is (farr[f], [tarr[t],]
// logmag[cloud][n]) (Code/dataset.py:50-54, 160-166), so the ESC_pc / ESC_pc_temp rows are never materialised.  The values
// are the ones build_clouds_kernel would have written: results are bit-identical.
struct PointSrc {
    const float* logmag;          // (B, N) log-magnitudes, frequency fastest; nullptr = use X32
    const float* farr;            // (nf)
    const float* tarr;            // (N / nf) or nullptr for 2-wide clouds
    int nf;
};
__device__ __forceinline__ void load_point(const float* __restrict__ X32, const PointSrc& src, int N, int d_in, size_t cloud, int n,
                                           float* x) {
    if (src.logmag != nullptr) {
        const int t = n / src.nf, f = n - t * src.nf;
        const float mag = __ldg(src.logmag + cloud * (size_t)N + n);
        x[0] = __ldg(src.farr + f);
        if (src.tarr != nullptr) { x[1] = __ldg(src.tarr + t); x[2] = mag; }
        else x[1] = mag;
    } else {
        const float* xp = X32 + (cloud * (size_t)N + n) * d_in;
        for (int k = 0; k < d_in; ++k) x[k] = __ldg(xp + k);
    }
}

// ------------------------------------------------------------------------------------ tail rule
// A cloud whose point count leaves 1..tail_max points past a multiple of 128 (the FST cloud: 1025 = 8 x 128 + 1) would
// pay a whole pipeline pass of the tensor-core kernels for them (measured: the 9th tile of the FST cloud costs 77 % of a
// full tile).  Those leftover points are handled exactly (fp32, CUDA cores) where the partial results are merged anyway:
// as an extra softmax slot in the finalize kernels (mab0 / PMA, where points are keys) and by a one-block-per-point
// kernel for mab1 (where points are queries).  main_points() is the prefix the tensor-core kernels process.
constexpr int TC_TAIL_MAX = 4;
__host__ __device__ __forceinline__ int main_points(int nb, int tail_max) {
    const int r = nb & 127;
    return (tail_max > 0 && nb >= 128 && r != 0 && r <= tail_max) ? (nb - r) : nb;
}

// ------------------------------------------------------------------------------------ reduce kernel
struct RParams {
    const float* X32;             // (B, N, d_in) fp32      [DIN64 == false]
    const __nv_bfloat16* Y16;     // (B, N, 64) bf16        [DIN64 == true]
    int N, d_in, tiles_total, tiles_per_split, nsplit;
    int n_work;                   // work items (cloud, split) for the persistent kernels
    const int* counts;            // nullable (B): valid points per cloud (variable-size sets); rows past it are padding
    int tail_max;                 // leftover points (N mod 128 <= tail_max) are left to the tail paths
    const uint8_t* Aq;            // 16 KB query operand (stacked pairs, or all-heads image in PMA mode)
    const float* Wkv32;           // (128, d_in) fp32       [DIN64 == false]
    const float* bkv;             // (128)
    const uint8_t* Wkv16;         // 16 KB B operand        [DIN64 == true]
    long long* timeline;          // debug: per-phase clock64 stamps of CTA 0 / softmax warp 0 (nullptr = off)
    int* redo;                    // (n_work) flags: the streaming variant (NWG == 4) marks work items in which a row outgrew its
                                  // reference exponent; they are redone by the NWG == 2 variant launched with redo_only = 1
    int redo_only;
    float* part;                  // (B, slots, 8, 10, 64): per (head, query) row: m (log2 domain), l, acc[8];
                                  // query index fastest so that a warp's 32 rows store/load 128 contiguous bytes
};

// One 32-column chunk of the online softmax: p = 2^(s - m) as bf16 pairs, partial row sum (packed fp32x2 math).
// pairs of each 32-column chunk whose exponentials run on the FMA pipe (polynomial) instead of MUFU: 6 of 16
#ifndef PCA_POLY5
#define PCA_POLY5 0x0101u       // 2 of 16 pairs: -3 % on the reduce kernel, -5 % on the pooled kernel (4 of 16: -2 %)
#endif
constexpr uint32_t kPolyPairs = PCA_POLY5;
__device__ __forceinline__ void exp_chunk32(const uint32_t* v, const float2 neg_m2, float2& sum2, uint32_t* pk) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        const float2 x = add2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), neg_m2);
        const float2 pr = ((kPolyPairs >> (j >> 1)) & 1u) ? ex2_poly2(x) : make_float2(ex2(x.x), ex2(x.y));
        sum2 = add2(sum2, pr);
        pk[j >> 1] = pack_bf16(pr.x, pr.y);
    }
}
__device__ __forceinline__ float max_chunk32(const uint32_t* v, float mx) {
    float m0 = mx, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        m0 = max3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
        m1 = max3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
    return fmaxf(m0, m1);
}

constexpr int TC_THREADS16 = 16 * 32;

// ------------------------------------------------------------------------------------ finalize (ISAB) on tensor cores
// Same arithmetic as finalize_isab_kernel, two clouds per CTA iteration (MMA rows = (cloud of the pair, inducing point)).
// Both GEMMs run as 3-term split-bf16 products (a_hi w_hi + a_lo w_hi + a_hi w_lo, fp32 accumulate), i.e. to ~2^-17
// relative -- the inducing-point summaries stay fp32-grade and only the K / V images are rounded to bf16, as before.
struct F2Params {
    const float* part; int nslots; int B;
    const float* Qp;              // (64, 64) hoisted fc_q(I)
    const uint8_t* WoS;           // hi 8192 | lo 8192
    const float* bo;
    const uint8_t* WkvS;          // hi 16384 | lo 16384
    const float* bkv;
    uint8_t* KVblk;               // per cloud 32768 B
    float* H_debug;               // nullable (B, 64, 64)
    // leftover points (tail rule): the mab0 keys / values of those points are formed here in fp32 and merged as one more
    // softmax slot
    const float* X32;             // (B, N, dk) fp32 points       [dk <= 4]
    const __nv_bfloat16* Y16;     // (B, N, 64) bf16 points       [dk == 64]
    int N, dk;
    const int* counts;            // nullable (B)
    int tail_max;
    const float* Wkv0T;           // mab0 [Wk;Wv]^T, k-major (dk, 128)
    const float* bkv0;            // mab0 bk | bv (128)
    PointSrc src;                 // [dk <= 4] alternative to X32
};
constexpr uint32_t F2_A = 0, F2_F = 64, F2_KV = 128;       // hi 32 | lo 32 | F 64 | KV 128  (256 columns)
struct F2Smem {
    static constexpr int WO = 0;
    static constexpr int WKV = 16384;
    static constexpr int BIAS = WKV + 32768;      // bo (64) | bkv (128)
    static constexpr int BARS = BIAS + 192 * 4;
    static constexpr int TOTAL = BARS + 32;
};

__device__ __forceinline__ void st_split_a(uint32_t tb, uint32_t lane_base, const float* o) {
    // A operand (this thread's row, K = 64) as bf16 pairs: hi in columns [0, 32), lo in [32, 64)
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(o[c0 + j], o[c0 + j + 1]);
            hi[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
            lo[j >> 1] = pack_bf16(o[c0 + j] - __bfloat162float(h.x), o[c0 + j + 1] - __bfloat162float(h.y));
        }
        tmem_st16(tmem_addr(tb, lane_base, F2_A + (c0 >> 1)), hi);
        tmem_st16(tmem_addr(tb, lane_base, F2_A + 32 + (c0 >> 1)), lo);
    }
}

__global__ void __launch_bounds__(128) finalize_isab_tc_kernel(const F2Params P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sWo = smem + F2Smem::WO;
    uint8_t* sWkv = smem + F2Smem::WKV;
    float* sBo = reinterpret_cast<float*>(smem + F2Smem::BIAS);
    float* sBkv = sBo + 64;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + F2Smem::BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    __shared__ float sKt[2][TC_TAIL_MAX][TD], sVt[2][TC_TAIL_MAX][TD];      // leftover points: K / V rows
    const int tid = threadIdx.x, warp = tid >> 5;
    copy_to_smem(sWo, P.WoS, 16384);
    copy_to_smem(sWkv, P.WkvS, 32768);
    if (tid < 64) sBo[tid] = P.bo[tid];
    sBkv[tid] = P.bkv[tid];
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;
    const uint32_t lane_base = 32 * warp;
    const int cc = tid >> 6, m = tid & 63;
    const uint32_t wo = smem_u32(sWo), wkv = smem_u32(sWkv);
    uint32_t ph = 0;
    for (int pair = blockIdx.x; 2 * pair < P.B; pair += gridDim.x) {
        const int cloud = 2 * pair + cc;
        const bool valid = cloud < P.B;
        float o[64];
        // ---- leftover points of this cloud (tail rule): K|V rows in fp32; thread m forms K[.][m] and V[.][m]
        int r_tail = 0;
        if (valid && P.tail_max > 0) {
            const int nbt = P.counts ? max(1, min(P.N, __ldg(P.counts + cloud))) : P.N;
            const int n0 = main_points(nbt, P.tail_max);
            r_tail = nbt - n0;
            for (int j = 0; j < r_tail; ++j) {
                float kacc = __ldg(P.bkv0 + m), vacc = __ldg(P.bkv0 + 64 + m);
                if (P.Y16 == nullptr) {
                    float xr[4] = {0.f, 0.f, 0.f, 0.f};
                    load_point(P.X32, P.src, P.N, P.dk, (size_t)cloud, n0 + j, xr);
                    for (int k = 0; k < P.dk; ++k) {
                        const float x = xr[k];
                        kacc = fmaf(x, __ldg(P.Wkv0T + k * 128 + m), kacc);
                        vacc = fmaf(x, __ldg(P.Wkv0T + k * 128 + 64 + m), vacc);
                    }
                } else {
                    const uint4* yp = reinterpret_cast<const uint4*>(P.Y16 + ((size_t)cloud * P.N + n0 + j) * 64);
#pragma unroll 2
                    for (int c8 = 0; c8 < 8; ++c8) {
                        const uint4 u = __ldg(yp + c8);
                        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 xy = __bfloat1622float2(h2[q]);
                            const int k = c8 * 8 + 2 * q;
                            kacc = fmaf(xy.x, __ldg(P.Wkv0T + k * 128 + m), kacc);
                            vacc = fmaf(xy.x, __ldg(P.Wkv0T + k * 128 + 64 + m), vacc);
                            kacc = fmaf(xy.y, __ldg(P.Wkv0T + (k + 1) * 128 + m), kacc);
                            vacc = fmaf(xy.y, __ldg(P.Wkv0T + (k + 1) * 128 + 64 + m), vacc);
                        }
                    }
                }
                sKt[cc][j][m] = kacc;
                sVt[cc][j][m] = vacc;
            }
        }
        __syncthreads();
        // ---- merge the point splits / column halves (+ the tail slot): O = Qp + A V
        if (valid) {
            const size_t sstride = (size_t)TH * 10 * TM;
            const bool two = P.nslots == 2, one = P.nslots == 1;
#pragma unroll
            for (int hb = 0; hb < TH; hb += 4) {
            // two-slot case (the usual one): the partials of four heads are fetched together, 80 loads in flight per
            // thread instead of 20 -- the kernel is bound by the latency of these rounds, not by bandwidth
            float w0[4][10], w1[4][10];
            if (two) {
#pragma unroll
                for (int hh = 0; hh < 4; ++hh) {
                    const float* pq = P.part + (((size_t)cloud * 2) * TH + hb + hh) * 10 * TM + m;
#pragma unroll
                    for (int j = 0; j < 10; ++j) { w0[hh][j] = __ldg(pq + j * TM); w1[hh][j] = __ldg(pq + sstride + j * TM); }
                }
            } else if (one) {      // reduce6: one partial per cloud (clouds of up to 2048 points), nothing to merge but the tail slot
#pragma unroll
                for (int hh = 0; hh < 4; ++hh) {
                    const float* pq = P.part + ((size_t)cloud * TH + hb + hh) * 10 * TM + m;
#pragma unroll
                    for (int j = 0; j < 10; ++j) w0[hh][j] = __ldg(pq + j * TM);
                }
            }
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) {
                const int h = hb + hh;
                const float* pp = P.part + (((size_t)cloud * P.nslots) * TH + h) * 10 * TM + m;
                float a[8], l, mmax;
                if (two) {
                    const float* v0 = w0[hh];
                    const float* v1 = w1[hh];
                    mmax = fmaxf(v0[0], v1[0]);
                    const float w0 = exp2f(v0[0] - mmax), w1 = exp2f(v1[0] - mmax);
                    l = fmaf(v0[1], w0, v1[1] * w1);
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = fmaf(v0[2 + j], w0, v1[2 + j] * w1);
                } else if (one) {
                    const float* v0 = w0[hh];
                    mmax = v0[0];
                    l = v0[1];
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = v0[2 + j];
                } else {
                    mmax = -INFINITY;
                    for (int sl = 0; sl < P.nslots; ++sl) mmax = fmaxf(mmax, __ldg(pp + sl * sstride));
                    l = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = 0.f;
                    for (int sl = 0; sl < P.nslots; ++sl) {
                        const float* ps = pp + sl * sstride;
                        const float wgt = exp2f(__ldg(ps) - mmax);
                        l = fmaf(__ldg(ps + TM), wgt, l);
#pragma unroll
                        for (int j = 0; j < 8; ++j) a[j] = fmaf(__ldg(ps + (2 + j) * TM), wgt, a[j]);
                    }
                }
                const float4 q0 = __ldg(reinterpret_cast<const float4*>(P.Qp + m * TD + h * 8));
                const float4 q1 = __ldg(reinterpret_cast<const float4*>(P.Qp + m * TD + h * 8) + 1);
                const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                for (int jt = 0; jt < r_tail; ++jt) {         // leftover points: one more key each, exact fp32 scores
                    float sc = 0.f;
#pragma unroll
                    for (int d = 0; d < 8; ++d) sc = fmaf(qv[d], sKt[cc][jt][h * 8 + d], sc);
                    sc *= kScaleLog2e;
                    const float mnew = fmaxf(mmax, sc);
                    const float w_old = exp2f(mmax - mnew), w_new = exp2f(sc - mnew);
                    l = fmaf(l, w_old, w_new);
#pragma unroll
                    for (int d = 0; d < 8; ++d) a[d] = fmaf(a[d], w_old, w_new * sVt[cc][jt][h * 8 + d]);
                    mmax = mnew;
                }
                const float inv = 1.f / l;
#pragma unroll
                for (int j = 0; j < 8; ++j) o[h * 8 + j] = fmaf(a[j], inv, qv[j]);
            }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 64; ++j) o[j] = 0.f;
        }
        {   // pull the next pair's partials towards L2 while this pair's GEMMs and image stores run
            const int ncloud = 2 * (pair + (int)gridDim.x) + cc;
            if (ncloud < P.B) {
                const char* nb = reinterpret_cast<const char*>(P.part + ((size_t)ncloud * P.nslots) * TH * 10 * TM);
                const int nbytes = P.nslots * TH * 10 * TM * 4;
                for (int off = m * 128; off < nbytes; off += 64 * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + off));
            }
        }
        // ---- F = O Wo^T (split product), H = O + relu(F + bo)
        st_split_a(tb, lane_base, o);
        tmem_st_wait();
        fence_before_sync();
        __syncthreads();
        if (warp == 0 && elect_one()) {
            fence_after_sync();
            const uint32_t idesc = idesc_bf16(128, 64, 0, 0);
#pragma unroll
            for (int term = 0; term < 3; ++term)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma_ts(tmem_addr(tb, 0, F2_F), tmem_addr(tb, 0, F2_A + (term == 1 ? 32 : 0) + 8 * ks),
                           smem_desc(wo + (term == 2 ? 8192 : 0) + ks * 2048, 1024, 128), idesc, (term | ks) != 0);
            mma_commit(bar);
        }
        mbar_wait(bar, ph);
        ph ^= 1;
        fence_after_sync();
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t f[32];
            tmem_ld32(tmem_addr(tb, lane_base, F2_F + c0), f);
            tmem_ld_wait32(f);
#pragma unroll
            for (int j = 0; j < 32; ++j) o[c0 + j] += fmaxf(__uint_as_float(f[j]) + sBo[c0 + j], 0.f);
        }
        if (P.H_debug && valid) {
            float4* hd = reinterpret_cast<float4*>(P.H_debug + ((size_t)cloud * TM + m) * TD);
#pragma unroll
            for (int j = 0; j < 16; ++j) hd[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        }
        // ---- [Kp | Vp] = H [Wk;Wv]^T + b (split product) -> block-diagonal bf16 operand images
        st_split_a(tb, lane_base, o);
        tmem_st_wait();
        fence_before_sync();
        __syncthreads();
        if (warp == 0 && elect_one()) {
            fence_after_sync();
            const uint32_t idesc = idesc_bf16(128, 128, 0, 0);
#pragma unroll
            for (int term = 0; term < 3; ++term)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma_ts(tmem_addr(tb, 0, F2_KV), tmem_addr(tb, 0, F2_A + (term == 1 ? 32 : 0) + 8 * ks),
                           smem_desc(wkv + (term == 2 ? 16384 : 0) + ks * 4096, 2048, 128), idesc, (term | ks) != 0);
            mma_commit(bar);
        }
        mbar_wait(bar, ph);
        ph ^= 1;
        fence_after_sync();
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_addr(tb, lane_base, F2_KV + c0), v);
            tmem_ld_wait32(v);
            if (valid) {
                uint8_t* img = P.KVblk + (size_t)cloud * 32768 + (c0 >= 64 ? 16384 : 0);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int h = ((c0 & 63) >> 3) + q, pr = h >> 1, ch = h & 1;
                    uint4 u;
                    const float ksc = c0 < 64 ? kKImageScale : 1.f;      // K columns carry the softmax scale
                    u.x = pack_bf16((__uint_as_float(v[8 * q + 0]) + sBkv[c0 + 8 * q + 0]) * ksc, (__uint_as_float(v[8 * q + 1]) + sBkv[c0 + 8 * q + 1]) * ksc);
                    u.y = pack_bf16((__uint_as_float(v[8 * q + 2]) + sBkv[c0 + 8 * q + 2]) * ksc, (__uint_as_float(v[8 * q + 3]) + sBkv[c0 + 8 * q + 3]) * ksc);
                    u.z = pack_bf16((__uint_as_float(v[8 * q + 4]) + sBkv[c0 + 8 * q + 4]) * ksc, (__uint_as_float(v[8 * q + 5]) + sBkv[c0 + 8 * q + 5]) * ksc);
                    u.w = pack_bf16((__uint_as_float(v[8 * q + 6]) + sBkv[c0 + 8 * q + 6]) * ksc, (__uint_as_float(v[8 * q + 7]) + sBkv[c0 + 8 * q + 7]) * ksc);
                    *reinterpret_cast<uint4*>(img + pr * 4096 + ch * 2048 + (ch * 64 + m) * 16) = u;
                    *reinterpret_cast<uint4*>(img + pr * 4096 + ch * 2048 + ((1 - ch) * 64 + m) * 16) = make_uint4(0, 0, 0, 0);
                }
            }
        }
        fence_before_sync();
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 256);
}

// ------------------------------------------------------------------------------------ apply kernel
struct AParams {
    const float* X32;             // (B, N, d_in)           [DIN64 == false]
    const __nv_bfloat16* Y16in;   // (B, N, 64)             [DIN64 == true]
    int N, d_in, tiles_total, tiles_per_split, nsplit, n_work;
    const int* counts;            // nullable (B): valid points per cloud
    int tail_max;                 // leftover points (N mod 128 <= tail_max) are left to the tail kernel
    const uint8_t* KVblk;         // per cloud: K image 16384 B | V image 16384 B
    const float* Wq32;            // (64, d_in)             [DIN64 == false]
    const float* bq;              // (64)
    const uint8_t* Wq16;          // fc_q B operand: 8 KB (N=64, K=64) [DIN64] or the 2 KB split-bf16 K=16 image
    const uint8_t* Wo16;          // 8 KB B operand
    const float* bo;              // (64)
    __nv_bfloat16* Yout;          // (B, N, 64)
    long long* timeline;          // debug (PCA_TIMELINE builds): clock64 stamps of CTA 0 / softmax warp 0
    CUtensorMap tmapY;            // [DIN64, apply4] Y16in as an (8, B * N, 8 chunks) bf16 tensor, box (8, 128, 8): one TMA per tile
    CUtensorMap tmapYo;           // [apply4] Yout, same geometry (TMA stores of whole tiles)
    PointSrc src;                 // [DIN64 == false, apply4] alternative to X32
};

// ====================================================================================== chain-scheduled kernels
// The 128 score columns of a head pair are produced and consumed as two independent 64-column "chains" per softmax
// warpgroup (4 chains per CTA, each with its own TMEM half-buffer, barriers and MMA-issuing thread), so the
// P V -> next Q K^T round trip of one chain is hidden behind the softmax of the warpgroup's other chain and the two
// warpgroups never wait on each other.  (Earlier generations of these kernels -- fixed MMA order, 24-warp variants,
// epilogue on the softmax warps, CUDA-core finalize -- are in the git history; DESIGN.md records what each step bought.)

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// one mbarrier arrival per warp (barrier counts are in warps): every lane has fenced its own writes, the
// __syncwarp orders them before the elected lane's releasing arrive
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}

// TMEM columns (reduce): 4 x 64 score/probability half-buffers | 8 x 16 outputs (pair, half) | projection
constexpr uint32_t R2_S = 0, R2_O = 256, R2_PROJ = 384;

struct R2Smem {
    static constexpr int AQ = 0;
    static constexpr int KV = 16384;
    static constexpr int W = KV + 65536;
    static constexpr int Y = W + 16384;
    static constexpr int SMALL = Y + 16384;
    static constexpr int BARS = SMALL + (128 * 4 + 128) * 4;
    static constexpr int TOTAL = BARS + 32 * 8 + 16;
};

// 64-column softmax step on registers: returns the chunk max
__device__ __forceinline__ float max64(const uint32_t* va, const uint32_t* vb) { return max_chunk32(vb, max_chunk32(va, -INFINITY)); }


// Rare path of mab_reduce5_tc_kernel for one 64-column item of a row (ragged last tile, or a row whose scores outgrew
// its reference exponent): everything goes through TMEM in 16-column steps so that no register state of the caller is
// needed.  Called warp-uniformly.  Returns the row sum of the item; P (bf16) is written over columns [0, 32) of the
// score buffer; the row's accumulator (8 columns at oaddr) and running sum are rescaled if the reference moves.
__device__ __noinline__ float reduce5_item_slow(uint32_t sbase, uint32_t oaddr, int nv, bool first, float& m_used, float& l_run) {
    float mx = -INFINITY;
#pragma unroll 1
    for (int c0 = 0; c0 < nv; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(sbase + c0, v);
        tmem_ld_wait16(v);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (c0 + j < nv) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    if (first) {                         // the chain's first P V of the work item overwrites the accumulator
        m_used = mx;
        l_run = 0.f;
    } else {
        const bool move = mx >= m_used + 53.f;
        if (__any_sync(0xffffffffu, move)) {
            const float f = move ? ex2(m_used - mx) : 1.f;
            uint32_t o[8];
            tmem_ld8(oaddr, o);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
            tmem_st8(oaddr, o);
            tmem_st_wait();
            l_run *= f;
            if (move) m_used = mx;
        }
    }
    float sum = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16], pk[8];
        if (c0 < nv) {
            tmem_ld16(sbase + c0, v);
            tmem_ld_wait16(v);
        }
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const float p0 = (c0 + j < nv) ? ex2(__uint_as_float(v[j]) - m_used) : 0.f;
            const float p1 = (c0 + j + 1 < nv) ? ex2(__uint_as_float(v[j + 1]) - m_used) : 0.f;
            sum += p0 + p1;
            pk[j >> 1] = pack_bf16(p0, p1);
        }
        tmem_st8(sbase + (c0 >> 1), pk);
    }
    tmem_st_wait();
    return sum;
}

// Fifth generation of the reduce kernel: like mab_reduce2_tc_kernel (persistent, 4 chains), but the outputs ACCUMULATE IN TMEM
// across the tiles of a work item against a per-row reference exponent m_used that is fixed by the first tile (softmax is
// shift invariant; with fp32 sums, bf16 probabilities and fp32 accumulators any reference within 2^+-60 of the true row
// maximum is exact to rounding).  No row maximum is taken after the first tile; a row whose running sum shows that some
// score exceeded the reference by more than 2^54 takes a slow path that re-references the row (rescales its sum and its
// TMEM accumulator).  The softmax warps therefore never read the outputs until the work item ends.
// Persistent: grid = min(#work items, #SMs); CTA k walks the work items (cloud, point-split) k, k + grid, ...
// Barriers, TMEM and the resident operands are set up once; all pipelines (producer -> MMA -> softmax) run
// straight across work-item boundaries, so there is no per-cloud fill/drain bubble.
// NWG softmax warpgroups: 2 (two chains each, 64 score registers per thread) or 4 (one chain each, scores streamed in
// 32-column chunks at 88 registers; experiment, see the host code).
template <bool DIN64, int NWG>
__global__ void __launch_bounds__((4 * NWG + 12) * 32, 1) mab_reduce5_tc_kernel(const RParams P) {
    constexpr int WP = 4 * NWG;          // first producer warp (8 of them: two per row quadrant, K columns / V columns)
    constexpr int WM = WP + 8;           // first MMA warp
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sAq = smem + R2Smem::AQ;
    uint8_t* sKV = smem + R2Smem::KV;
    uint8_t* sW = smem + R2Smem::W;
    uint8_t* sY = smem + R2Smem::Y;
    float* sWsm = reinterpret_cast<float*>(smem + R2Smem::SMALL);
    float* sBias = sWsm + 128 * 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + R2Smem::BARS);
    uint64_t* kv_full = bars;          // [2] count 256 (every lane of the 8 producer warps)
    uint64_t* kv_empty = bars + 2;     // [2] count 4 (chains)
    uint64_t* s_full = bars + 4;       // [4] count 1
    uint64_t* p_ready = bars + 8;      // [4] count 128 (every lane of the owning warpgroup)
    uint64_t* o_done = bars + 12;      // [4] count 1   (chain: all P V of the work item complete)
    uint64_t* y_full = bars + 20;      // count 8
    uint64_t* proj_done = bars + 21;   // count 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_work = P.n_work, wstep = gridDim.x;
#ifdef PCA_TIMELINE
    long long* tl2 = nullptr;
    int tl2_n = 0;
    if (P.timeline != nullptr && blockIdx.x == 0 && lane == 0) {
        if (warp == WP) tl2 = P.timeline + 4000;
        else if (warp == WM) tl2 = P.timeline + 6000;
    }
    auto stamp2 = [&](int tag) {
        if (tl2 != nullptr && tl2_n < 1000) { tl2[2 * tl2_n] = tag; tl2[2 * tl2_n + 1] = clock64(); ++tl2_n; }
    };
#else
    auto stamp2 = [&](int) {};
#endif
    // tiles of work item w; nb = valid points of its cloud (variable-size sets: rows past nb are padding)
    auto work_tiles = [&](int w, int& cloud, int& split, int& tile0, int& nb) {
        cloud = w / P.nsplit;
        split = w - cloud * P.nsplit;
        tile0 = split * P.tiles_per_split;
        nb = main_points(P.counts ? max(1, min(P.N, __ldg(P.counts + cloud))) : P.N, P.tail_max);
        return max(0, min((nb + 127) >> 7, tile0 + P.tiles_per_split) - tile0);
    };

    // redo mode: only the work items flagged by the streaming variant are processed (normally none: leave at once)
    auto skipped = [&](int w) { return P.redo_only != 0 && __ldg(P.redo + w) == 0; };
    if (P.redo_only != 0) {
        // the CTA's flags are read by all of its threads at once (one dependent load each, not a serial scan per thread)
        int any = 0;
        for (int w = blockIdx.x + (int)threadIdx.x * wstep; w < n_work; w += (int)blockDim.x * wstep) any |= (__ldg(P.redo + w) != 0);
        if (!__syncthreads_or(any)) return;
    }
    copy_to_smem(sAq, P.Aq, 16384);
    if (DIN64) copy_to_smem(sW, P.Wkv16, 16384);
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
        sBias[i] = P.bkv[i];
        if (!DIN64) {
#pragma unroll
            for (int k = 0; k < 4; ++k) sWsm[i * 4 + k] = (k < P.d_in) ? P.Wkv32[i * P.d_in + k] : 0.f;
        }
    }
    if (warp == WM) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 256); mbar_init(&kv_empty[i], 4); }
        for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 128); }
        for (int i = 0; i < 4; ++i) mbar_init(&o_done[i], 1);
        mbar_init(y_full, 8);
        mbar_init(proj_done, 1);
        fence_barrier_init();
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;

    if (warp >= WM) {
        reg_dec<40>();
        {
            // =================================================================== one MMA-issuing warp per chain
            // chain c = warp - 12: strictly serial  Q K^T -> (warpgroup softmax) -> P V -> next Q K^T  on its own
            // half-buffer; the four chains never wait on each other.  The whole warp runs the loop (uniform control flow
            // keeps the descriptors in uniform registers); one elected lane issues the tcgen05 instructions.
            const int c = warp - WM, half = c & 1;
            const bool leader = elect_one();
            const uint32_t idesc_s = idesc_bf16(128, 64, 0, 0);
            const uint32_t idesc_pv = idesc_bf16(128, 16, 0, 1);
            const uint32_t aq = smem_u32(sAq), kvb = smem_u32(sKV);
            int gt = 0;                                    // tiles processed by this CTA so far
            uint32_t ph_p = 0;                             // phase of this chain's p_ready barrier
            for (int w = blockIdx.x; w < n_work; w += wstep) {
                if (skipped(w)) continue;
                int cloud, split, tile0, nb;
                const int ntiles = work_tiles(w, cloud, split, tile0, nb);
                for (int it = 0; it < ntiles; ++it, ++gt) {
                    const uint32_t kbase = kvb + (gt & 1) * 32768, vbase = kbase + 16384;
                    stamp2(60);
                    mbar_wait(&kv_full[gt & 1], (gt >> 1) & 1);
                    fence_after_sync();
                    stamp2(61);
                    // a half without valid points (ragged last tile) is skipped by the chain and by its warpgroup alike
                    const bool empty_half = nb - (tile0 + it) * 128 <= 64 * half;
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp) {
                        if (empty_half) break;
                        const int p = (c >> 1) + 2 * pp;
                        if (leader) {
                            mma_ss(tmem_addr(tb, 0, R2_S + 64 * c), smem_desc(aq + p * 4096, 2048, 128),
                                   smem_desc(kbase + 2 * p * 2048 + half * 1024, 2048, 128), idesc_s, 0);
                            mma_commit(&s_full[c]);
                        }
                        __syncwarp();
                        stamp2(62);
                        mbar_wait(&p_ready[c], ph_p);
                        ph_p ^= 1;
                        fence_after_sync();
                        stamp2(63);
                        if (leader) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                // NWG == 4: P occupies columns 32..63 of the buffer, keys 32..63 first, then keys 0..31
                                const uint32_t pcol = NWG == 4 ? 32 + ks * 8 : ks * 8;
                                const int vrow = NWG == 4 ? ((ks + 2) & 3) : ks;
                                mma_ts(tmem_addr(tb, 0, R2_O + 16 * (2 * p + half)), tmem_addr(tb, 0, R2_S + 64 * c + pcol),
                                       smem_desc(vbase + 2 * p * 2048 + half * 1024 + vrow * 256, 128, 2048), idesc_pv,
                                       (it > 0 || ks > 0) ? 1u : 0u);
                            }
                        }
                        __syncwarp();
                    }
                    if (leader) mma_commit(&kv_empty[gt & 1]);            // 4 chains x 1 arrival free the K|V stage
                    __syncwarp();
                }
                if (leader) mma_commit(&o_done[c]);                       // the work item's accumulators are final
                __syncwarp();
            }
        }
    } else if (warp >= WP) {
        // launch allocation: 96 registers (20 warps) / 72 (28 warps)
        reg_dec<56>();
        // =================================================================== producers: K|V tiles
        // 8 warps: warp pw stages / converts row quadrant pw & 3; pw < 4 produces the K columns, pw >= 4 the V columns.
        // The global loads of tile t+1 are issued before tile t is converted (they are the longest latency on this path).
        const int pw = warp - WP, quad = pw & 3, colhalf = pw >> 2;
        const int row = 32 * quad + lane;
        int gt = 0;
        uint4 yv[4];                              // DIN64: this thread's half row (4 x 8 bf16) of the NEXT tile
        float xn[4] = {0.f, 0.f, 0.f, 0.f};       // !DIN64: the next tile's point
        bool vnext = false;
        auto prefetch = [&](int cloud, int n, int nb) {
            vnext = n < nb;
            if (DIN64) {
                const uint4* src = reinterpret_cast<const uint4*>(P.Y16 + ((size_t)cloud * P.N + (vnext ? n : 0)) * 64) + 4 * colhalf;
#pragma unroll
                for (int c = 0; c < 4; ++c) yv[c] = vnext ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
            } else {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) xn[kk] = 0.f;
                if (vnext) {
                    const float* xp = P.X32 + ((size_t)cloud * P.N + n) * P.d_in;
                    for (int kk = 0; kk < P.d_in; ++kk) xn[kk] = __ldg(xp + kk);
                }
            }
        };
        // (cloud, tile) iterator one step ahead of the processing loop
        int w_n = blockIdx.x, it_n = 0, cloud_n = 0, split_n = 0, tile0_n = 0, nb_n = 0, ntiles_n = 0;
        auto advance = [&]() {           // move (w_n, it_n) to the next existing tile; returns false at the end
            while (w_n < n_work) {
                if (it_n == 0) ntiles_n = skipped(w_n) ? 0 : work_tiles(w_n, cloud_n, split_n, tile0_n, nb_n);
                if (it_n < ntiles_n) return true;
                w_n += wstep;
                it_n = 0;
            }
            return false;
        };
        bool have_next = advance();
        if (have_next) prefetch(cloud_n, (tile0_n + it_n) * 128 + row, nb_n);
        while (have_next) {
            const bool valid = vnext;
            const int stage = gt & 1;
            stamp2(50);
            uint8_t* sK = sKV + stage * 32768;
            uint8_t* sV = sK + 16384;
            float x[4] = {xn[0], xn[1], xn[2], xn[3]};
            if (DIN64) {
#pragma unroll
                for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(sY + (4 * colhalf + c) * 2048 + row * 16) = yv[c];
                fence_async_smem();
                fence_before_sync();
                warp_arrive(y_full);
                if (warp == WP) {
                    // one producer warp issues the K|V projection MMA (elected lane) once all 128 rows of Y are staged
                    mbar_wait(y_full, gt & 1);
                    fence_after_sync();
                    if (elect_one()) {
                        const uint32_t yb = smem_u32(sY), wb = smem_u32(sW);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            mma_ss(tmem_addr(tb, 0, R2_PROJ), smem_desc(yb + ks * 4096, 2048, 128), smem_desc(wb + ks * 4096, 2048, 128),
                                   idesc_bf16(128, 128, 0, 0), ks > 0);
                        mma_commit(proj_done);
                    }
                    __syncwarp();
                }
            }
            // ---- next tile's loads in flight while this one is converted
            ++it_n;
            have_next = advance();
            if (have_next) prefetch(cloud_n, (tile0_n + it_n) * 128 + row, nb_n);
            stamp2(51);
            if (DIN64) {
                mbar_wait(proj_done, gt & 1);
                fence_after_sync();
            }
            stamp2(52);
            if (gt >= 2) mbar_wait(&kv_empty[stage], ((gt >> 1) - 1) & 1);
            stamp2(53);
            uint8_t* dstbase = colhalf == 0 ? sK : sV;
            if (!DIN64) {
#pragma unroll 4
                for (int c = 0; c < 8; ++c) {
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int f = 64 * colhalf + c * 8 + j;
                        const float4 wv = *reinterpret_cast<const float4*>(sWsm + f * 4);
                        o[j] = valid ? fmaf(wv.w, x[3], fmaf(wv.z, x[2], fmaf(wv.y, x[1], fmaf(wv.x, x[0], sBias[f])))) : 0.f;
                    }
                    st_shared_8bf16(dstbase + c * 2048 + row * 16, o);
                }
            } else {
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_addr(tb, 32 * quad, R2_PROJ + 64 * colhalf + c0), v);
                    tmem_ld_wait32(v);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = valid ? __uint_as_float(v[8 * q + j]) + sBias[64 * colhalf + c0 + 8 * q + j] : 0.f;
                        st_shared_8bf16(dstbase + (c0 / 8 + q) * 2048 + row * 16, o);
                    }
                }
            }
            stamp2(54);
            fence_async_smem();
            fence_before_sync();
            mbar_arrive(&kv_full[stage]);
            ++gt;
        }
    } else if constexpr (NWG == 2) {
        reg_inc<160>();
        // =================================================================== softmax warpgroups (2 chains each)
        const int g = warp >> 2, quad = warp & 3;
        const int row = 32 * quad + lane;
        const uint32_t lane_base = 32 * quad;
        const uint32_t ocol_off = (row >= 64) ? 8u : 0u;
        float m_used[2][2], l_run[2][2];          // reference exponent and running sum per (half, pp)
        uint32_t ph_s[2] = {0, 0}, ph_done = 0;
        constexpr float kOverflow = 1.152921504606847e18f;      // 2^60
#ifdef PCA_TIMELINE
        long long* tl = (P.timeline != nullptr && blockIdx.x == 0 && warp == 0 && lane == 0) ? P.timeline : nullptr;
        int tl_n = 0;
        auto stamp = [&](int tag) {
            if (tl != nullptr && tl_n < 1000) { tl[2 * tl_n] = tag; tl[2 * tl_n + 1] = clock64(); ++tl_n; }
        };
#else
        auto stamp = [&](int) {};
#endif

        uint32_t va[32], vb[32];           // scores of the item being processed; refilled early with the next item's
        bool have = false;                 // va / vb already hold (in-flight) loads of the item about to be processed
        // `prefetch`: the next item of this warpgroup exists and lives on the OTHER chain (its scores were issued a whole
        // item ago and never wait for this item's probabilities), so its barrier wait and TMEM loads are issued before
        // this item's store-wait / fence / arrive instead of after them.
        auto softmax_item = [&](const int half, const int pp, const int nv, const bool first, const bool prefetch) {
            const int c = 2 * g + half, p = g + 2 * pp;
            const uint32_t sbase = tmem_addr(tb, lane_base, R2_S + 64 * c);
            const uint32_t oaddr = tmem_addr(tb, lane_base, R2_O + 16 * (2 * p + half) + ocol_off);
            if (nv == 0) {                     // empty half of a ragged tile: skipped (the chain thread skips it as well)
                if (first) { m_used[half][pp] = -INFINITY; l_run[half][pp] = 0.f; }
                return;
            }
            stamp(20);
            uint32_t pk0[16], pk1[16];
            if (!have) {
                mbar_wait(&s_full[c], ph_s[half]);
                ph_s[half] ^= 1;
                fence_after_sync();
                tmem_ld32(sbase, va);
                tmem_ld32(sbase + 32, vb);
            }
            have = false;
            stamp(24);
            tmem_ld_wait64(va, vb);
            stamp(25);
            float sum;
            if (nv == 64) {
                if (first) { m_used[half][pp] = max64(va, vb); l_run[half][pp] = 0.f; }
#ifdef PCA_R5_NOEXP
                // EXPERIMENT ONLY: no softmax arithmetic (what does the rest of the pipeline cost?)
#pragma unroll
                for (int q = 0; q < 16; ++q) { pk0[q] = pack_bf16(__uint_as_float(va[2 * q]), __uint_as_float(va[2 * q + 1])); pk1[q] = pack_bf16(__uint_as_float(vb[2 * q]), __uint_as_float(vb[2 * q + 1])); }
                sum = 1.f;
#else
                const float2 neg2 = make_float2(-m_used[half][pp], -m_used[half][pp]);
                float2 sum2 = make_float2(0.f, 0.f);
                exp_chunk32(va, neg2, sum2, pk0);
                exp_chunk32(vb, neg2, sum2, pk1);
                sum = sum2.x + sum2.y;
#endif
            } else {
                // ragged tail: columns >= nv are padding
                if (first) {
                    float mx = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j < nv) mx = fmaxf(mx, __uint_as_float(va[j]));
                        if (32 + j < nv) mx = fmaxf(mx, __uint_as_float(vb[j]));
                    }
                    m_used[half][pp] = mx;
                    l_run[half][pp] = 0.f;
                }
                const float m = m_used[half][pp];
                sum = 0.f;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float p0 = (j < nv) ? ex2(__uint_as_float(va[j]) - m) : 0.f;
                    const float p1 = (j + 1 < nv) ? ex2(__uint_as_float(va[j + 1]) - m) : 0.f;
                    const float p2 = (32 + j < nv) ? ex2(__uint_as_float(vb[j]) - m) : 0.f;
                    const float p3 = (33 + j < nv) ? ex2(__uint_as_float(vb[j + 1]) - m) : 0.f;
                    sum += (p0 + p1) + (p2 + p3);
                    pk0[j >> 1] = pack_bf16(p0, p1);
                    pk1[j >> 1] = pack_bf16(p2, p3);
                }
            }
            // Some score exceeded the reference by more than ~2^54 (or the sum overflowed): re-reference those rows.  All
            // earlier P V of this chain have completed (the chain thread's commit for THIS item's scores covers them), so
            // the slow path may rescale the accumulator in place.
            if (__any_sync(0xffffffffu, !(sum < kOverflow))) {
                l_run[half][pp] += reduce5_item_slow(sbase, oaddr, nv, false, m_used[half][pp], l_run[half][pp]);
            } else {
                l_run[half][pp] += sum;
                tmem_st16(sbase, pk0);
                tmem_st16(sbase + 16, pk1);
                stamp(26);
                if (prefetch) {
                    const int cn = 2 * g + (half ^ 1);
                    mbar_wait(&s_full[cn], ph_s[half ^ 1]);
                    ph_s[half ^ 1] ^= 1;
                    fence_after_sync();
                    const uint32_t snext = tmem_addr(tb, lane_base, R2_S + 64 * cn);
                    tmem_ld32(snext, va);
                    tmem_ld32(snext + 32, vb);
                    have = true;
                }
                tmem_st_wait();
                stamp(27);
            }
            fence_before_sync();
            mbar_arrive(&p_ready[c]);        // every lane arrives (count 128): no warp sync / divergent branch
        };
        for (int w = blockIdx.x; w < n_work; w += wstep) {
            if (skipped(w)) continue;
            int cloud, split, tile0, nb;
            const int ntiles = work_tiles(w, cloud, split, tile0, nb);
#pragma unroll
            for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2) { m_used[a2][b2] = -INFINITY; l_run[a2][b2] = 0.f; }     // a split may hold no tile of a short cloud
            for (int it = 0; it < ntiles; ++it) {
                const int n_valid = min(128, nb - (tile0 + it) * 128);
                const int nv0 = min(64, n_valid), nv1 = max(0, n_valid - 64);
                const bool first = it == 0;
                // Prefetching the next item's scores ahead of this item's store-wait / arrive is measured SLOWER on B200
                // (1.58 -> 1.69 ms/step), as it was for the apply kernel; kept behind PCA_R5_PREFETCH for experiments.
#ifdef PCA_R5_PREFETCH
                const bool both = nv1 > 0;                          // the tile has points in both column halves
#else
                const bool both = false;
#endif
                const bool next_tile = it + 1 < ntiles;             // ... and is followed by another tile of this work item
                softmax_item(0, 0, nv0, first, both);
                softmax_item(1, 0, nv1, first, both);
                softmax_item(0, 1, nv0, first, both);
                softmax_item(1, 1, nv1, first, both && next_tile);  // next: half 0 of the next tile (never empty)
            }
            // ---- the work item's accumulators are final once both chains have drained
            mbar_wait(&o_done[2 * g], ph_done);
            mbar_wait(&o_done[2 * g + 1], ph_done);
            ph_done ^= 1;
            fence_after_sync();
#pragma unroll
            for (int half = 0; half < 2; ++half)
#pragma unroll
                for (int pp = 0; pp < 2; ++pp) {
                    const int p = g + 2 * pp;
                    uint32_t o[8];
                    tmem_ld8(tmem_addr(tb, lane_base, R2_O + 16 * (2 * p + half) + ocol_off), o);
                    tmem_ld_wait();
                    const int h = 2 * p + (row >> 6);
                    float* dst = P.part + (((size_t)cloud * (2 * P.nsplit) + 2 * split + half) * TH + h) * 10 * TM + (row & 63);
                    dst[0] = m_used[half][pp];
                    dst[TM] = l_run[half][pp];
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[(2 + j) * TM] = (l_run[half][pp] > 0.f) ? __uint_as_float(o[j]) : 0.f;
                }
            fence_before_sync();
        }
        } else {
        reg_inc<88>();
        // =================================================================== 4 softmax warpgroups, one chain each
        // Scores are streamed in 32-column chunks (keys 32..63 first): after the work item's first tile no row maximum
        // is taken, so nothing but one chunk has to be live.  P goes to columns 32..63 of the chain's buffer.
        const int c = warp >> 2, quad = warp & 3;          // chain c: half = c & 1 of the pairs (c >> 1) + 2 pp
        const int half = c & 1, g = c >> 1;
        const int row = 32 * quad + lane;
        const uint32_t lane_base = 32 * quad;
        const uint32_t ocol_off = (row >= 64) ? 8u : 0u;
        const uint32_t sbase = tmem_addr(tb, lane_base, R2_S + 64 * c);
        float m_used[2], l_run[2];
        int cur_w = 0;
        uint32_t ph_s = 0, ph_done = 0;
        constexpr float kOverflow = 1.152921504606847e18f;      // 2^60
        auto softmax_item = [&](const int pp, const int nv, const bool first) {
            if (nv == 0) {
                if (first) { m_used[pp] = -INFINITY; l_run[pp] = 0.f; }
                return;
            }
            mbar_wait(&s_full[c], ph_s);
            ph_s ^= 1;
            fence_after_sync();
            uint32_t v[32], pk[16];
            float sum;
            if (nv == 64 && !first) {
                const float2 neg2 = make_float2(-m_used[pp], -m_used[pp]);
                float2 sum2 = make_float2(0.f, 0.f);
                tmem_ld32(sbase + 32, v);
                tmem_ld_wait32(v);
                exp_chunk32(v, neg2, sum2, pk);
                tmem_st16(sbase + 32, pk);
                tmem_ld32(sbase, v);
                tmem_ld_wait32(v);
                exp_chunk32(v, neg2, sum2, pk);
                tmem_st16(sbase + 48, pk);
                sum = sum2.x + sum2.y;
            } else {
                // first tile of the work item (reference exponent = row maximum) and ragged tails: masked two-pass
                float mx = -INFINITY;
                tmem_ld32(sbase + 32, v);
                tmem_ld_wait32(v);
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (32 + j < nv) mx = fmaxf(mx, __uint_as_float(v[j]));
                uint32_t v0[32];
                tmem_ld32(sbase, v0);
                tmem_ld_wait32(v0);
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < nv) mx = fmaxf(mx, __uint_as_float(v0[j]));
                if (first) { m_used[pp] = mx; l_run[pp] = 0.f; }
                const float m = m_used[pp];
                sum = 0.f;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float p0 = (32 + j < nv) ? ex2(__uint_as_float(v[j]) - m) : 0.f;
                    const float p1 = (33 + j < nv) ? ex2(__uint_as_float(v[j + 1]) - m) : 0.f;
                    sum += p0 + p1;
                    pk[j >> 1] = pack_bf16(p0, p1);
                }
                tmem_st16(sbase + 32, pk);
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float p0 = (j < nv) ? ex2(__uint_as_float(v0[j]) - m) : 0.f;
                    const float p1 = (j + 1 < nv) ? ex2(__uint_as_float(v0[j + 1]) - m) : 0.f;
                    sum += p0 + p1;
                    pk[j >> 1] = pack_bf16(p0, p1);
                }
                tmem_st16(sbase + 48, pk);
            }
            // a score outgrew the row's reference exponent by more than ~2^54 (or the sum overflowed): this variant cannot
            // re-reference a row (its scores are gone once P is written) -- the work item is redone by the exact variant
            if (!(sum < kOverflow)) P.redo[cur_w] = 1;
            l_run[pp] += sum;
            tmem_st_wait();
            fence_before_sync();
            mbar_arrive(&p_ready[c]);        // every lane arrives (count 128): no warp sync / divergent branch
        };
        for (int w = blockIdx.x; w < n_work; w += wstep) {
            if (skipped(w)) continue;
            cur_w = w;
            int cloud, split, tile0, nb;
            const int ntiles = work_tiles(w, cloud, split, tile0, nb);
            m_used[0] = m_used[1] = -INFINITY;
            l_run[0] = l_run[1] = 0.f;
            for (int it = 0; it < ntiles; ++it) {
                const int n_valid = min(128, nb - (tile0 + it) * 128);
                const int nv = half == 0 ? min(64, n_valid) : max(0, n_valid - 64);
                softmax_item(0, nv, it == 0);
                softmax_item(1, nv, it == 0);
            }
            mbar_wait(&o_done[c], ph_done);
            ph_done ^= 1;
            fence_after_sync();
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                const int p = g + 2 * pp;
                uint32_t o[8];
                tmem_ld8(tmem_addr(tb, lane_base, R2_O + 16 * (2 * p + half) + ocol_off), o);
                tmem_ld_wait();
                const int h = 2 * p + (row >> 6);
                float* dst = P.part + (((size_t)cloud * (2 * P.nsplit) + 2 * split + half) * TH + h) * 10 * TM + (row & 63);
                dst[0] = m_used[pp];
                dst[TM] = l_run[pp];
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[(2 + j) * TM] = (l_run[pp] > 0.f) ? __uint_as_float(o[j]) : 0.f;
            }
            fence_before_sync();
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == WM) tmem_dealloc(tb, 512);
}

// ====================================================================================== reduce kernel, sixth generation
// mab0 of an ISAB (Q = 64 inducing points, K = points) with the lessons of mab_apply4_tc_kernel (per-role clock stamps in
// profiles/r02_timeline_*.log): what bounds these kernels is not the MUFU pipe but the serial latency of every softmax item
// (barrier hand-offs, TMEM loads, the MMA round trip) and of the operand producers.
//   * W_k is folded into the query operand: s[(h, m), n] = q_{h,m} . (W_k,h y_n + b_k,h) = (W_k,h^T q_{h,m}) . y_n + const(h, m), and
//     the constant cancels in the softmax over n.  The A operand of Q K^T is therefore a MODEL constant (prep_kernel: "Gq",
//     128 rows x 64 features per head pair, softmax scale included) and its B operand is the raw input tile exactly as the
//     TMA engine delivers it -- no K projection, no TMEM -> register -> shared-memory round trip for the keys.  Only V is
//     projected (one N = 64 MMA per tile + bias step) and staged as the bf16 MN-major operand of P V.
//     (d_in <= 4: the split-bf16 point columns against a split Gq image, one K = 16 step; V on CUDA cores.)
//   * One chain per head PAIR (4 chains = 4 softmax warpgroups = 16 warps, four per scheduler, streaming the scores in
//     32-column chunks at 88 registers): both point halves of a tile go through the chain in turn, so a row has ONE reference
//     exponent, ONE running sum and ONE accumulator (64 TMEM columns in all instead of 128).
//   * Scores and probabilities live in separate TMEM regions (4 x 64 | 4 x 32): the score buffer is released as soon as the
//     second chunk is in registers, so Q K^T of the chain's next item runs under this item's arithmetic; the packed
//     probabilities are kept in registers (2 x 16) and written once s_full of the next item has been seen, which (same
//     issuing thread, in-order tensor pipe) implies that P V of the previous item has drained the probability buffer.
//   * Thread-side mbarrier arrivals by every lane (count 128); a loader warp feeds a three-deep tile ring by TMA.
// 28 warps: 0-15 softmax (warpgroup g = pair g), 16-19 V producers, 20-23 MMA chains, 24 loader, 25-27 idle.
// Partials: ONE slot per (cloud, split) -- part (B, nsplit, 8, 10, 64); the finalize kernel is told the slot count.
// EXACT = true is the redo pass over the work items the streaming pass flagged (a row outgrew its reference exponent; normally
// none): same operands and arithmetic, but every item keeps its scores in TMEM until the row has been re-referenced if needed
// (accumulator and running sum rescaled in place -- all earlier P V of the chain have completed by then), so no early
// release and no look-ahead.  Both passes therefore round identically, which keeps near-tie decisions order independent.
constexpr int R6_THREADS = 28 * 32;
constexpr int R6_YST = 3;                                 // input tile ring
constexpr uint32_t R6_S = 0, R6_P = 256, R6_O = 384, R6_PV = 448;      // 4 x 64 scores | 4 x 32 probabilities | 4 x 16 outputs | 64 V projection
struct R6Smem {
    static constexpr int GQ = 0;                          // 4 pairs x 16384 (64-wide) or 4 x 4096 (split K = 16 image)
    static constexpr int WV = 65536;                      // V projection B operand (N = 64, K = 80)            [DIN64]
    static constexpr int Y = WV + 10240;                  // R6_YST x 16384 input tiles (d_in <= 4: 4096 used per stage)
    static constexpr int V = Y + R6_YST * 16384;          // 2 x 16384 bf16 V tiles (MN-major B operand of P V)
    static constexpr int ONES = V + 32768;                // constant A operand of the bias K step
    static constexpr int SMALL = ONES + 4096;             // d_in <= 4: Wv (64 x 4 fp32) | bv (64)
    static constexpr int BARS = SMALL + (64 * 4 + 64) * 4;
    static constexpr int TOTAL = BARS + 48 * 8 + 16;
};
struct R6Params {
    const float* X32;             // (B, N, d_in) fp32      [DIN64 == false]
    int N, d_in, tiles_total, tiles_per_split, nsplit, n_work;
    const int* counts;            // nullable (B)
    int tail_max;
    const uint8_t* Gq;            // query operand image (see prep_kernel)
    const uint8_t* Wv16;          // V projection B operand with bias step  [DIN64]
    const float* Wv32;            // (64, d_in) fp32                        [DIN64 == false]
    const float* bv;              // (64)                                   [DIN64 == false]
    long long* timeline;
    int* redo;                    // (n_work) flags: set by the streaming pass, consumed by the EXACT pass
    float* part;                  // (B, nsplit, 8, 10, 64)
    CUtensorMap tmapY;            // [DIN64] the input as an (8, B * N, 8 chunks) bf16 tensor, box (8, 128, 8): one TMA per tile
    PointSrc src;                 // [DIN64 == false] alternative to X32
};

// One 64-column item of a row, exact variant: everything goes through TMEM in 16-column steps so that no register state of
// the caller is needed.  Called warp-uniformly, after P V of the previous item has completed.  Returns the row sum of the item;
// P (bf16) goes to the probability buffer; the row's accumulator (8 columns at oaddr) and running sum are rescaled if the
// reference exponent has to move.
__device__ __noinline__ float reduce6_item_exact(uint32_t sbase, uint32_t pbase, uint32_t oaddr, int nv, bool first, float& m_used,
                                                 float& l_run) {
    float mx = -INFINITY;
#pragma unroll 1
    for (int c0 = 0; c0 < nv; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(sbase + c0, v);
        tmem_ld_wait16(v);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (c0 + j < nv) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    if (first) {                         // the chain's first P V of the work item overwrites the accumulator
        m_used = mx;
        l_run = 0.f;
    } else {
        const bool move = mx >= m_used + 53.f;
        if (__any_sync(0xffffffffu, move)) {
            const float f = move ? ex2(m_used - mx) : 1.f;
            uint32_t o[8];
            tmem_ld8(oaddr, o);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
            tmem_st8(oaddr, o);
            tmem_st_wait();
            l_run *= f;
            if (move) m_used = mx;
        }
    }
    float sum = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16], pk[8];
        if (c0 < nv) {
            tmem_ld16(sbase + c0, v);
            tmem_ld_wait16(v);
        }
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const float p0 = (c0 + j < nv) ? ex2(__uint_as_float(v[j]) - m_used) : 0.f;
            const float p1 = (c0 + j + 1 < nv) ? ex2(__uint_as_float(v[j + 1]) - m_used) : 0.f;
            sum += p0 + p1;
            pk[j >> 1] = pack_bf16(p0, p1);
        }
        tmem_st8(pbase + (c0 >> 1), pk);
    }
    tmem_st_wait();
    return sum;
}

template <bool DIN64, bool EXACT>
__global__ void __launch_bounds__(R6_THREADS, 1) mab_reduce6_tc_kernel(const __grid_constant__ R6Params P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sGq = smem + R6Smem::GQ;
    uint8_t* sWv = smem + R6Smem::WV;
    uint8_t* sY = smem + R6Smem::Y;
    uint8_t* sV = smem + R6Smem::V;
    uint8_t* sOnes = smem + R6Smem::ONES;
    float* sWv32 = reinterpret_cast<float*>(smem + R6Smem::SMALL);
    float* sBv = sWv32 + 64 * 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + R6Smem::BARS);
    uint64_t* y_full = bars;           // [3] count 1 (loader; + TMA bytes)
    uint64_t* y_empty = bars + 3;      // [3] count 4 chains (+ 1: the V projection commit, DIN64)
    uint64_t* v_full = bars + 6;       // [2] count 128 (producer lanes)
    uint64_t* v_empty = bars + 8;      // [2] count 4 (chains: their P V of the tile have run)
    uint64_t* s_full = bars + 10;      // [4] count 1
    uint64_t* s_free = bars + 14;      // [4] count 128 (lanes of the warpgroup: scores are in registers)
    uint64_t* p_ready = bars + 18;     // [4] count 128
    uint64_t* p_free = bars + 22;      // [4] count 1 (commit of P V; waited for only where no next item covers it)
    uint64_t* o_done = bars + 26;      // [4] count 1 (chain: all P V of the work item complete)
    uint64_t* vp_done = bars + 30;     // count 1 (V projection MMA)
    uint64_t* vp_free = bars + 31;     // count 128 (producer lanes have read the projection)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 48);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_work = P.n_work, wstep = gridDim.x;
#ifdef PCA_TIMELINE
    long long* tl2 = nullptr;
    int tl2_n = 0;
    if (P.timeline != nullptr && blockIdx.x == 0 && lane == 0) {
        if (warp == 0) tl2 = P.timeline;
        else if (warp == 16) tl2 = P.timeline + 2000;
        else if (warp == 12) tl2 = P.timeline + 4000;
        else if (warp == 20) tl2 = P.timeline + 6000;
    }
    auto stamp2 = [&](int tag) {
        if (tl2 != nullptr && tl2_n < 1000) { tl2[2 * tl2_n] = tag; tl2[2 * tl2_n + 1] = clock64(); ++tl2_n; }
    };
#else
    auto stamp2 = [&](int) {};
#endif
    auto work_tiles = [&](int w, int& cloud, int& split, int& tile0, int& nb) {
        cloud = w / P.nsplit;
        split = w - cloud * P.nsplit;
        tile0 = split * P.tiles_per_split;
        nb = main_points(P.counts ? max(1, min(P.N, __ldg(P.counts + cloud))) : P.N, P.tail_max);
        return max(0, min((nb + 127) >> 7, tile0 + P.tiles_per_split) - tile0);
    };

    // EXACT: only the work items flagged by the streaming pass are processed (normally none: leave at once)
    auto skipped = [&](int w) { return EXACT && __ldg(P.redo + w) == 0; };
    if (EXACT) {
        int any = 0;
        for (int w = blockIdx.x + (int)threadIdx.x * wstep; w < n_work; w += (int)blockDim.x * wstep) any |= (__ldg(P.redo + w) != 0);
        if (!__syncthreads_or(any)) return;
    }
    copy_to_smem(sGq, P.Gq, DIN64 ? 65536 : 16384);
    if (DIN64) copy_to_smem(sWv, P.Wv16, 10240);
    else {
        for (int i = threadIdx.x; i < 64; i += blockDim.x) {
            sBv[i] = P.bv[i];
#pragma unroll
            for (int k = 0; k < 4; ++k) sWv32[i * 4 + k] = (k < P.d_in) ? P.Wv32[i * P.d_in + k] : 0.f;
        }
    }
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        *reinterpret_cast<uint4*>(sOnes + i * 16) = (i < 128) ? make_uint4(0x3F803F80u, 0, 0, 0) : make_uint4(0, 0, 0, 0);
    if (warp == 20) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 0) {
        for (int i = 0; i < R6_YST; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], DIN64 ? 5 : 4); }
        for (int i = 0; i < 2; ++i) { mbar_init(&v_full[i], 128); mbar_init(&v_empty[i], 4); }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 128);
            mbar_init(&p_ready[i], 128); mbar_init(&p_free[i], 1);
            mbar_init(&o_done[i], 1);
        }
        mbar_init(vp_done, 1);
        mbar_init(vp_free, 128);
        fence_barrier_init();
    }
    if (DIN64 && threadIdx.x == 24 * 32) tma_prefetch_desc(&P.tmapY);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;

    if (warp >= 25) {
        reg_dec<40>();
    } else if (warp == 24) {
        reg_dec<40>();
        // =================================================================== loader: input tiles
        int gt = 0;
        for (int w = blockIdx.x; w < n_work; w += wstep) {
            if (skipped(w)) continue;
            int cloud, split, tile0, nb;
            const int ntiles = work_tiles(w, cloud, split, tile0, nb);
            for (int it = 0; it < ntiles; ++it, ++gt) {
                const int stage = gt % R6_YST;
                uint8_t* dst = sY + stage * 16384;
                if (gt >= R6_YST) mbar_wait(&y_empty[stage], ((gt / R6_YST) - 1) & 1);
                if (DIN64) {
                    if (lane == 0) {
                        const long long r0 = (long long)cloud * P.N + (long long)(tile0 + it) * 128;
                        mbar_arrive_expect_tx(&y_full[stage], 16384);
                        tma_load_3d(dst, &P.tmapY, 0, (int)r0, 0, &y_full[stage]);      // one box: (8 elements, 128 rows, 8 chunks)
                    }
                    __syncwarp();
                } else {
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int row = 32 * rr + lane;
                        const int n = (tile0 + it) * 128 + row;
                        float x[4] = {0.f, 0.f, 0.f, 0.f};
                        if (n < nb) load_point(P.X32, P.src, P.N, P.d_in, (size_t)cloud, n, x);
                        float cols[16];
                        split_x16(x, cols);
                        st_shared_8bf16(dst + row * 16, cols);
                        st_shared_8bf16(dst + 2048 + row * 16, cols + 8);
                    }
                    fence_async_smem();
                    fence_before_sync();
                    warp_arrive(&y_full[stage]);
                }
            }
        }
    } else if (warp >= 20) {
        reg_dec<40>();
        // =================================================================== MMA chain g = head pair g
        const int g = warp - 20;
        const bool leader = elect_one();
        const uint32_t idesc_s = idesc_bf16(128, 64, 0, 0);
        const uint32_t idesc_pv = idesc_bf16(128, 16, 0, 1);
        const uint32_t gq = smem_u32(sGq) + g * (DIN64 ? 16384 : 4096), yb = smem_u32(sY), vb = smem_u32(sV);
        const uint32_t s_tm = tmem_addr(tb, 0, R6_S + 64 * g), p_tm = tmem_addr(tb, 0, R6_P + 32 * g), o_tm = tmem_addr(tb, 0, R6_O + 16 * g);
        uint32_t par = 0;
        bool have_prev = false, prev_first = false, prev_tile_last = false;
        uint32_t prev_v = 0;
        int prev_vstage = 0;
        auto pv_prev = [&]() {
            stamp2(62);
            mbar_wait(&p_ready[g], par ^ 1);
            fence_after_sync();
            stamp2(63);
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma_ts(o_tm, p_tm + ks * 8, smem_desc(prev_v + ks * 256, 128, 2048), idesc_pv, (prev_first && ks == 0) ? 0u : 1u);
                mma_commit(&p_free[g]);
                if (prev_tile_last) mma_commit(&v_empty[prev_vstage]);     // this chain is done with the tile's V image
            }
            __syncwarp();
        };
        int gt = 0;
        for (int w = blockIdx.x; w < n_work; w += wstep) {
            if (skipped(w)) continue;
            int cloud, split, tile0, nb;
            const int ntiles = work_tiles(w, cloud, split, tile0, nb);
            bool first_item = true;
            for (int it = 0; it < ntiles; ++it, ++gt) {
                const int stage = gt % R6_YST, vstage = gt & 1;
                stamp2(60);
                mbar_wait(&y_full[stage], (gt / R6_YST) & 1);
                mbar_wait(&v_full[vstage], (gt >> 1) & 1);
                fence_after_sync();
                stamp2(61);
                const int n_valid = min(128, nb - (tile0 + it) * 128);
                const int n_items = n_valid > 64 ? 2 : 1;              // a half without valid points is skipped by chain and warpgroup alike
#pragma unroll 1
                for (int hf = 0; hf < n_items; ++hf) {
                    if (have_prev) {
                        mbar_wait(&s_free[g], par ^ 1);
                        fence_after_sync();
                    }
                    if (leader) {
                        if (DIN64) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                mma_ss(s_tm, smem_desc(gq + ks * 4096, 2048, 128), smem_desc(yb + stage * 16384 + ks * 4096 + hf * 1024, 2048, 128),
                                       idesc_s, ks > 0);
                        } else {
                            mma_ss(s_tm, smem_desc(gq, 2048, 128), smem_desc(yb + stage * 16384 + hf * 1024, 2048, 128), idesc_s, 0u);
                        }
                        mma_commit(&s_full[g]);
                        if (hf == n_items - 1) mma_commit(&y_empty[stage]);       // this chain has consumed the input tile
                    }
                    __syncwarp();
                    if (have_prev) pv_prev();
                    par ^= 1;
                    have_prev = true;
                    prev_first = first_item;
                    first_item = false;
                    prev_tile_last = hf == n_items - 1;
                    prev_vstage = vstage;
                    prev_v = vb + vstage * 16384 + 2 * g * 2048 + hf * 1024;
                }
            }
            // end of the work item: retire the pending P V, publish the accumulators, restart the hand-shake
            if (have_prev) {
                pv_prev();
                have_prev = false;
                mbar_wait(&s_free[g], par ^ 1);
            }
            if (leader) mma_commit(&o_done[g]);
            __syncwarp();
        }
    } else if (warp >= 16) {
        reg_dec<56>();
        // =================================================================== V producers (thread = point row)
        const int quad = warp & 3;
        const int row = 32 * quad + lane;
        int gt = 0;
        for (int w = blockIdx.x; w < n_work; w += wstep) {
            if (skipped(w)) continue;
            int cloud, split, tile0, nb;
            const int ntiles = work_tiles(w, cloud, split, tile0, nb);
            for (int it = 0; it < ntiles; ++it, ++gt) {
                const int stage = gt % R6_YST, vstage = gt & 1;
                const int n = (tile0 + it) * 128 + row;
                const bool valid = n < nb;
                uint8_t* dstV = sV + vstage * 16384;
                stamp2(50);
                if (DIN64) {
                    if (warp == 16) {
                        // V = Y Wv^T + bv into the projection columns, free once all four producer warps have read the previous tile
                        mbar_wait(&y_full[stage], (gt / R6_YST) & 1);
                        if (gt >= 1) mbar_wait(vp_free, (gt - 1) & 1);
                        fence_after_sync();
                        if (elect_one()) {
                            const uint32_t yb = smem_u32(sY) + stage * 16384, wv = smem_u32(sWv), ones = smem_u32(sOnes);
                            const uint32_t d = tmem_addr(tb, 0, R6_PV);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                mma_ss(d, smem_desc(yb + ks * 4096, 2048, 128), smem_desc(wv + ks * 2048, 1024, 128), idesc_bf16(128, 64, 0, 0), ks > 0);
                            mma_ss(d, smem_desc(ones, 2048, 128), smem_desc(wv + 8192, 1024, 128), idesc_bf16(128, 64, 0, 0), 1u);
                            mma_commit(vp_done);
                            mma_commit(&y_empty[stage]);
                        }
                        __syncwarp();
                    }
                    mbar_wait(vp_done, gt & 1);
                    fence_after_sync();
                    stamp2(51);
                    if (gt >= 2) mbar_wait(&v_empty[vstage], ((gt >> 1) - 1) & 1);
                    stamp2(52);
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t v[32];
                        tmem_ld32(tmem_addr(tb, 32 * quad, R6_PV + 32 * hf), v);
                        tmem_ld_wait32(v);
                        if (hf == 1) { fence_before_sync(); mbar_arrive(vp_free); }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 u = make_uint4(0, 0, 0, 0);
                            if (valid) {
                                u.x = pack_bf16(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
                                u.y = pack_bf16(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
                                u.z = pack_bf16(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
                                u.w = pack_bf16(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
                            }
                            *reinterpret_cast<uint4*>(dstV + (4 * hf + q) * 2048 + row * 16) = u;
                        }
                    }
                } else {
                    float x[4] = {0.f, 0.f, 0.f, 0.f};
                    if (valid) load_point(P.X32, P.src, P.N, P.d_in, (size_t)cloud, n, x);
                    if (gt >= 2) mbar_wait(&v_empty[vstage], ((gt >> 1) - 1) & 1);
#pragma unroll 4
                    for (int c = 0; c < 8; ++c) {
                        float o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int f = c * 8 + j;
                            const float4 wv = *reinterpret_cast<const float4*>(sWv32 + f * 4);
                            o[j] = valid ? fmaf(wv.w, x[3], fmaf(wv.z, x[2], fmaf(wv.y, x[1], fmaf(wv.x, x[0], sBv[f])))) : 0.f;
                        }
                        st_shared_8bf16(dstV + c * 2048 + row * 16, o);
                    }
                }
                stamp2(54);
                fence_async_smem();
                fence_before_sync();
                mbar_arrive(&v_full[vstage]);
            }
        }
    } else {
        reg_inc<88>();
        // =================================================================== softmax warpgroup g = head pair g (thread = (head of pair, query) row)
        const int g = warp >> 2, quad = warp & 3;
        const int row = 32 * quad + lane;
        const uint32_t lane_base = 32 * quad;
        const uint32_t sbase = tmem_addr(tb, lane_base, R6_S + 64 * g);
        const uint32_t pbase = tmem_addr(tb, lane_base, R6_P + 32 * g);
        const uint32_t oaddr = tmem_addr(tb, lane_base, R6_O + 16 * g + ((row >= 64) ? 8u : 0u));
        constexpr float kOverflow = 1.152921504606847e18f;      // 2^60
        uint32_t par = 0, ph_done = 0;
        bool have_scores = false, first_ever = true;
        for (int w = blockIdx.x; w < n_work; w += wstep) {
            if (skipped(w)) continue;
            int cloud, split, tile0, nb;
            const int ntiles = work_tiles(w, cloud, split, tile0, nb);
            float m_used = -INFINITY, l_run = 0.f;
            bool flagged = false;
            for (int it = 0; it < ntiles; ++it) {
                const int n_valid = min(128, nb - (tile0 + it) * 128);
                const int n_items = n_valid > 64 ? 2 : 1;
#pragma unroll 1
                for (int hf = 0; hf < n_items; ++hf) {
                    const int nv = hf == 0 ? min(64, n_valid) : n_valid - 64;
                    const bool first = it == 0 && hf == 0;
                    const bool has_next = hf + 1 < n_items || it + 1 < ntiles;       // a later item of this chain in this work item
                    uint32_t v[32], pkA[16], pkB[16];
                    float sum;
                    stamp2(20);
                    if (EXACT) {
                        // scores stay in TMEM; P V of the previous item must have drained the probability buffer AND left
                        // the accumulator quiescent before the row may be re-referenced
                        mbar_wait(&s_full[g], par);
                        fence_after_sync();
                        if (!first_ever) {
                            mbar_wait(&p_free[g], par ^ 1);
                            fence_after_sync();
                        }
                        first_ever = false;
                        l_run += reduce6_item_exact(sbase, pbase, oaddr, nv, first, m_used, l_run);
                        fence_before_sync();
                        mbar_arrive(&s_free[g]);
                        mbar_arrive(&p_ready[g]);
                        par ^= 1;
                        continue;
                    }
                    if (!have_scores) {
                        mbar_wait(&s_full[g], par);
                        fence_after_sync();
                    }
                    stamp2(24);
                    if (nv == 64 && !first) {
                        const float2 neg2 = make_float2(-m_used, -m_used);
                        float2 sum2 = make_float2(0.f, 0.f);
                        tmem_ld32(sbase, v);
                        tmem_ld_wait32(v);
                        exp_chunk32(v, neg2, sum2, pkA);
                        tmem_ld32(sbase + 32, v);
                        tmem_ld_wait32(v);
                        fence_before_sync();
                        mbar_arrive(&s_free[g]);                 // the scores are in registers: the chain's next Q K^T may run
                        stamp2(25);
                        exp_chunk32(v, neg2, sum2, pkB);
                        sum = sum2.x + sum2.y;
                    } else {
                        // first item of the work item (reference exponent := row maximum) and ragged tails: masked, three loads
                        float mx = -INFINITY;
                        tmem_ld32(sbase, v);
                        tmem_ld_wait32(v);
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nv) mx = fmaxf(mx, __uint_as_float(v[j]));
                        tmem_ld32(sbase + 32, v);
                        tmem_ld_wait32(v);
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (32 + j < nv) mx = fmaxf(mx, __uint_as_float(v[j]));
                        if (first) { m_used = mx; l_run = 0.f; }
                        const float m = m_used;
                        sum = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float p0 = (32 + j < nv) ? ex2(__uint_as_float(v[j]) - m) : 0.f;
                            const float p1 = (33 + j < nv) ? ex2(__uint_as_float(v[j + 1]) - m) : 0.f;
                            sum += p0 + p1;
                            pkB[j >> 1] = pack_bf16(p0, p1);
                        }
                        tmem_ld32(sbase, v);
                        tmem_ld_wait32(v);
                        fence_before_sync();
                        mbar_arrive(&s_free[g]);
                        stamp2(25);
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float p0 = (j < nv) ? ex2(__uint_as_float(v[j]) - m) : 0.f;
                            const float p1 = (j + 1 < nv) ? ex2(__uint_as_float(v[j + 1]) - m) : 0.f;
                            sum += p0 + p1;
                            pkA[j >> 1] = pack_bf16(p0, p1);
                        }
                    }
                    // a score outgrew the row's reference exponent by more than ~2^54 (or the sum overflowed): this streaming
                    // kernel cannot re-reference a row -- the work item is redone by the exact variant (flag set once below)
                    if (!(sum < kOverflow)) flagged = true;
                    l_run += sum;
                    stamp2(21);
                    // the probability buffer holds the previous item's P until its P V has run: s_full of the NEXT item covers
                    // it (the chain issues that P V before the next Q K^T); the last item of a work item waits for the commit
                    if (has_next) {
                        mbar_wait(&s_full[g], par ^ 1);
                        fence_after_sync();
                        have_scores = true;
                    } else {
                        if (!first_ever) {
                            mbar_wait(&p_free[g], par ^ 1);
                            fence_after_sync();
                        }
                        have_scores = false;
                    }
                    first_ever = false;
                    stamp2(22);
                    tmem_st16(pbase, pkA);
                    tmem_st16(pbase + 16, pkB);
                    tmem_st_wait();
                    stamp2(27);
                    fence_before_sync();
                    mbar_arrive(&p_ready[g]);
                    par ^= 1;
                }
            }
            if (!EXACT && flagged) P.redo[w] = 1;
            // ---- the work item's accumulators are final once the chain has drained
            mbar_wait(&o_done[g], ph_done);
            ph_done ^= 1;
            fence_after_sync();
            {
                uint32_t o[8];
                if (ntiles > 0) {
                    tmem_ld8(oaddr, o);
                    tmem_ld_wait();
                }
                const int h = 2 * g + (row >> 6);
                float* dst = P.part + (((size_t)cloud * P.nsplit + split) * TH + h) * 10 * TM + (row & 63);
                dst[0] = m_used;
                dst[TM] = l_run;
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[(2 + j) * TM] = (l_run > 0.f) ? __uint_as_float(o[j]) : 0.f;
            }
            fence_before_sync();
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 20) tmem_dealloc(tb, 512);
}

// ====================================================================================== apply kernel, third generation
// 20 warps: 0-7 softmax (two warpgroups), 8-11 producer, 12-15 MMA (one issuing thread per chain), 16-19 epilogue.
// Differences to mab_apply2_tc_kernel (motivated by the per-phase timeline in profiles/: a third of the tile time was
// epilogue work on the softmax warps with the MUFU pipe idle):
//   * the Q projection ALWAYS runs as an MMA into the tile's TMEM accumulator OQ (d_in <= 4: one K=16 split-bf16 step),
//     and the P V products of all heads accumulate ON TOP of it, so O1 = Qp + A V materialises in TMEM without any
//     register work (bias bq is added where OQ is read);
//   * probabilities are normalised before they are written (P = e / sum), so no per-row rescale of the outputs;
//   * the softmax warps do nothing but  wait -> tcgen05.ld -> max -> 2^x -> sum -> scale -> pack -> tcgen05.st -> arrive;
//   * dedicated epilogue warps turn OQ into the bf16 A operand of fc_o (staged in shared memory), and after the fc_o MMA
//     emit Y = O1 + relu(fc_o(O1) + bo); three tile accumulators are in flight (Q projection -> P V -> epilogue), so
//     the projection of tile t+2 never waits for the epilogue of tile t;
//   * every MMA-issuing warp runs in uniform control flow with one elected lane issuing (descriptors stay in uniform
//     registers; a lone `if (lane == 0)` thread costs ~250 cycles per tcgen05.mma in R2UR/ELECT loops and spills).
// A chain is a head PAIR (its two heads run back to back on the chain's score buffer and accumulate into the pair's 16
// output columns from one issuing thread, so no cross-thread accumulate hazard exists).
constexpr int TC_THREADS20 = 20 * 32;
constexpr int A3_NBUF = 3;                        // tile accumulators in flight (Q projection -> P V -> epilogue)
constexpr uint32_t A3_S = 0, A3_OQ = 256, A3_F = 448;      // 4 x 64 scores | 3 x 64 tile accumulators | 64 fc_o  (= 512 columns)

struct A3Smem {
    static constexpr int IMG = 0;                 // 2 x (K image 16384 | V image 16384), one per work item in flight
    static constexpr int WO = 65536;              // fc_o B operand (N=64, K=64)
    static constexpr int WQ = WO + 8192;          // fc_q B operand: (N=64, K=64) or the split-bf16 K=16 image
    static constexpr int AQ = WQ + 8192;          // 2 stages x 16384: scaled bf16 queries (A operand of Q K^T)
    static constexpr int YA = AQ + 32768;         // input tile: Y (128 x 64 bf16) or the split-bf16 X columns (128 x 16)
    static constexpr int O1 = YA + 16384;         // O1 tile as the bf16 A operand of fc_o (written by the epilogue warps)
    static constexpr int SMALL = O1 + 16384;      // bq (64) | bo (64)
    static constexpr int BARS = SMALL + 128 * 4;
    static constexpr int TOTAL = BARS + 40 * 8 + 16;
};

// pairs of each 32-column chunk whose exponentials run on the FMA pipe (degree-3 polynomial) instead of MUFU
#ifndef PCA_POLY3
#define PCA_POLY3 0x1111u      // 4 of every 16 pairs: measured -2 % on the apply kernel (8 of 16: +5 %, the FMA pipe saturates)
#endif
#ifndef PCA_A3_PREFETCH
#define PCA_A3_PREFETCH 0
#endif
#ifndef PCA_A3_STAGGER
#define PCA_A3_STAGGER 0
#endif
constexpr uint32_t kPolyPairs3 = PCA_POLY3;
// e = 2^(s - m) kept in place as fp32, partial row sum
__device__ __forceinline__ void exp_keep32(uint32_t* v, const float2 neg_m2, float2& sum2) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
#ifdef PCA_A3_NOMAX
        const float2 x = make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
        (void)neg_m2;
#else
        const float2 x = add2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), neg_m2);
#endif
        const float2 e = ((kPolyPairs3 >> (j >> 1)) & 1u) ? ex2_poly2(x) : make_float2(ex2(x.x), ex2(x.y));
        sum2 = add2(sum2, e);
        v[j] = __float_as_uint(e.x);
        v[j + 1] = __float_as_uint(e.y);
    }
}
__device__ __forceinline__ void scale_pack32(const uint32_t* v, const float2 inv2, uint32_t* pk) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        const float2 p = mul2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), inv2);
        pk[j >> 1] = pack_bf16(p.x, p.y);
    }
}

template <bool DIN64>
__global__ void __launch_bounds__(TC_THREADS20, 1) mab_apply3_tc_kernel(const AParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sImg = smem + A3Smem::IMG;
    uint8_t* sWo = smem + A3Smem::WO;
    uint8_t* sWq = smem + A3Smem::WQ;
    uint8_t* sAQ = smem + A3Smem::AQ;
    uint8_t* sYA = smem + A3Smem::YA;
    uint8_t* sO1 = smem + A3Smem::O1;
    float* sBq = reinterpret_cast<float*>(smem + A3Smem::SMALL);
    float* sBo = sBq + 64;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A3Smem::BARS);
    uint64_t* aq_full = bars;          // [2] count 4 (producer warps)
    uint64_t* aq_empty = bars + 2;     // [2] count 4 (chain commits)
    uint64_t* s_full = bars + 4;       // [4] count 1
    uint64_t* p_ready = bars + 8;      // [4] count 4 (warps of the owning warpgroup)
    uint64_t* o_full = bars + 12;      // [3] count 4 (chain commits)          -- per OQ buffer
    uint64_t* oq_free = bars + 15;     // [3] count 4 (epilogue warps)         -- per OQ buffer
    uint64_t* ya_full = bars + 18;     // count 4
    uint64_t* qp_done = bars + 19;     // count 1
    uint64_t* o1_ready = bars + 20;    // count 4 (epilogue warps)
    uint64_t* f_full = bars + 21;      // count 1
    uint64_t* img_full = bars + 22;    // [2] count 4 (producer warps)
    uint64_t* img_empty = bars + 24;   // [2] count 4 (chain commits)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_work = P.n_work, wstep = gridDim.x;
#ifdef PCA_TIMELINE
    // debug: role warps 16 (epilogue), 8 (producer), 12 (MMA chain 0) of CTA 0 record into their own 1000-stamp segments
    long long* tl2 = nullptr;
    int tl2_n = 0;
    if (P.timeline != nullptr && blockIdx.x == 0 && lane == 0) {
        if (warp == 16) tl2 = P.timeline + 2000;
        else if (warp == 8) tl2 = P.timeline + 4000;
        else if (warp == 12) tl2 = P.timeline + 6000;
    }
    auto stamp2 = [&](int tag) {
        if (tl2 != nullptr && tl2_n < 1000) { tl2[2 * tl2_n] = tag; tl2[2 * tl2_n + 1] = clock64(); ++tl2_n; }
    };
#else
    auto stamp2 = [&](int) {};
#endif
    // tiles of work item w; nb = valid points of its cloud (variable-size sets: rows past nb are padding)
    auto work_tiles = [&](int w, int& cloud, int& tile0, int& nb) {
        cloud = w / P.nsplit;
        const int split = w - cloud * P.nsplit;
        tile0 = split * P.tiles_per_split;
        nb = main_points(P.counts ? max(1, min(P.N, __ldg(P.counts + cloud))) : P.N, P.tail_max);
        return max(0, min((nb + 127) >> 7, tile0 + P.tiles_per_split) - tile0);
    };

    copy_to_smem(sWo, P.Wo16, 8192);
    copy_to_smem(sWq, P.Wq16, DIN64 ? 8192 : 2048);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) {
        sBq[i] = P.bq[i];
        sBo[i] = P.bo[i];
    }
    if (warp == 12) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&aq_full[i], 4); mbar_init(&aq_empty[i], 4);
            mbar_init(&img_full[i], 4); mbar_init(&img_empty[i], 4);
        }
        for (int i = 0; i < A3_NBUF; ++i) { mbar_init(&o_full[i], 4); mbar_init(&oq_free[i], 4); }
        for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); }
        mbar_init(ya_full, 4);
        mbar_init(qp_done, 1);
        mbar_init(o1_ready, 4);
        mbar_init(f_full, 1);
        fence_barrier_init();
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;

    if (warp >= 16) {
        // register budget: the CTA's pool is 20 warps x 96; softmax 8 x 144 + producer 4 x 56 + MMA 4 x 40 + epilogue 4 x 96 = 1920
        // =================================================================== epilogue warps (thread = point row)
        const int quad = warp & 3;
        const int row = 32 * quad + lane;
        const uint32_t lane_base = 32 * quad;
        int gt = 0;
        for (int w = blockIdx.x; w < n_work; w += wstep) {
            int cloud, tile0, nb;
            const int ntiles = work_tiles(w, cloud, tile0, nb);
            for (int it = 0; it < ntiles; ++it, ++gt) {
                const int buf = gt % A3_NBUF;
                const int n = (tile0 + it) * 128 + row;
                const bool valid = n < nb;
                const bool live = (tile0 + it) * 128 + 32 * quad < nb;
                const uint32_t oq = tmem_addr(tb, lane_base, A3_OQ + 64 * buf);
                // ---- O1 = OQ + bq  ->  bf16 A operand of fc_o, written back to TMEM
                stamp2(40);
                mbar_wait(&o_full[buf], (gt / A3_NBUF) & 1);
                stamp2(41);
                fence_after_sync();
                if (live) {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t v[32];
                        tmem_ld32(oq + 32 * hf, v);
                        tmem_ld_wait32(v);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float o[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(v[8 * q + j]) + sBq[32 * hf + 8 * q + j];
                            st_shared_8bf16(sO1 + (4 * hf + q) * 2048 + row * 16, o);
                        }
                    }
                }
                stamp2(42);
                fence_async_smem();
                fence_before_sync();
                warp_arrive(o1_ready);
                if (warp == 16) {
                    // one epilogue warp issues fc_o (elected lane): F = O1 Wo^T, A operand from TMEM
                    mbar_wait(o1_ready, gt & 1);
                    fence_after_sync();
                    if (elect_one()) {
                        const uint32_t wo = smem_u32(sWo), o1b = smem_u32(sO1);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            mma_ss(tmem_addr(tb, 0, A3_F), smem_desc(o1b + ks * 4096, 2048, 128), smem_desc(wo + ks * 2048, 1024, 128),
                                   idesc_bf16(128, 64, 0, 0), ks > 0);
                        mma_commit(f_full);
                    }
                    __syncwarp();
                }
                // ---- Y = O1 + relu(F + bo)
                stamp2(43);
                mbar_wait(f_full, gt & 1);
                stamp2(44);
                fence_after_sync();
                if (live) {
                    __nv_bfloat16* dst = P.Yout + ((size_t)cloud * P.N + (valid ? n : 0)) * 64;
#pragma unroll
                    for (int qf = 0; qf < 4; ++qf) {
                        uint32_t o[16], f[16];
                        tmem_ld16(oq + 16 * qf, o);
                        tmem_ld16(tmem_addr(tb, lane_base, A3_F + 16 * qf), f);
                        tmem_ld_wait16(o);
                        tmem_ld_wait16(f);
                        uint4 out[2];
                        uint32_t* ow = reinterpret_cast<uint32_t*>(out);
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const int c0 = 16 * qf + j;
                            const float y0 = (__uint_as_float(o[j]) + sBq[c0]) + fmaxf(__uint_as_float(f[j]) + sBo[c0], 0.f);
                            const float y1 = (__uint_as_float(o[j + 1]) + sBq[c0 + 1]) + fmaxf(__uint_as_float(f[j + 1]) + sBo[c0 + 1], 0.f);
                            ow[j >> 1] = pack_bf16(y0, y1);
                        }
                        if (valid) {
                            uint4* d4 = reinterpret_cast<uint4*>(dst + 16 * qf);
                            d4[0] = out[0];
                            d4[1] = out[1];
                        }
                    }
                }
                stamp2(45);
                fence_before_sync();
                warp_arrive(&oq_free[buf]);
            }
        }
    } else if (warp >= 12) {
        reg_dec<40>();
        {
            // =================================================================== one MMA-issuing warp per chain (= head pair)
            // The whole warp runs the loop (uniform control flow keeps the descriptors in uniform registers); one
            // elected lane issues the tcgen05 instructions.
            const int c = warp - 12;                      // pair c, score buffer c, output columns 16c..16c+15
            const bool leader = elect_one();
            const uint32_t idesc_s = idesc_bf16(128, 64, 0, 0);
            const uint32_t idesc_pv = idesc_bf16(128, 16, 0, 1);
            const uint32_t img = smem_u32(sImg), aqb = smem_u32(sAQ);
            int gt = 0, wl = 0;
            for (int w = blockIdx.x; w < n_work; w += wstep, ++wl) {
                int cloud, tile0, nb;
                const int ntiles = work_tiles(w, cloud, tile0, nb);
                const uint32_t kb = img + (wl & 1) * 32768 + c * 4096, vb = kb + 16384;
                mbar_wait(&img_full[wl & 1], (wl >> 1) & 1);
                fence_after_sync();
                for (int it = 0; it < ntiles; ++it, ++gt) {
                    const int buf = gt % A3_NBUF;
                    const uint32_t a_desc_base = aqb + (gt & 1) * 16384 + 2 * c * 2048;
                    const uint32_t d_o = tmem_addr(tb, 0, A3_OQ + 64 * buf + 16 * c);
                    stamp2(60);
                    mbar_wait(&aq_full[gt & 1], (gt >> 1) & 1);
                    fence_after_sync();
                    stamp2(61);
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        if (leader) {
                            mma_ss(tmem_addr(tb, 0, A3_S + 64 * c), smem_desc(a_desc_base, 2048, 128),
                                   smem_desc(kb + hh * 1024, 2048, 128), idesc_s, 0);
                            mma_commit(&s_full[c]);
                        }
                        __syncwarp();
                        stamp2(62);
                        mbar_wait(&p_ready[c], hh);          // two items per tile: parities 0, 1
                        fence_after_sync();
                        stamp2(63);
                        if (leader) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                mma_ts(d_o, tmem_addr(tb, 0, A3_S + 64 * c + ks * 8), smem_desc(vb + hh * 1024 + ks * 256, 128, 2048),
                                       idesc_pv, 1u);
                        }
                        __syncwarp();
                    }
                    if (leader) {
                        mma_commit(&o_full[buf]);               // 4 chains: the tile's O1 is complete
                        mma_commit(&aq_empty[gt & 1]);          // ... and its query stage is free
                    }
                    __syncwarp();
                }
                if (leader) mma_commit(&img_empty[wl & 1]);
                __syncwarp();
            }
        }
    } else if (warp >= 8) {
        reg_dec<56>();
        // =================================================================== producer: images, Q projection, scaled query operand
        const int quad = warp & 3;
        const int row = 32 * quad + lane;
        const int ptid = threadIdx.x - 256;          // 0..127
        int gt = 0, wl = 0;
        for (int w = blockIdx.x; w < n_work; w += wstep, ++wl) {
            int cloud, tile0, nb;
            const int ntiles = work_tiles(w, cloud, tile0, nb);
            {   // stage this cloud's block-diagonal K / V images (32 KB)
                if (wl >= 2) mbar_wait(&img_empty[wl & 1], ((wl >> 1) - 1) & 1);
                const uint4* src = reinterpret_cast<const uint4*>(P.KVblk + (size_t)cloud * 32768);
                uint4* dst = reinterpret_cast<uint4*>(sImg + (wl & 1) * 32768);
#pragma unroll 4
                for (int i = ptid; i < 2048; i += 128) dst[i] = __ldg(src + i);
                fence_async_smem();
                fence_before_sync();
                warp_arrive(&img_full[wl & 1]);
            }
            for (int it = 0; it < ntiles; ++it, ++gt) {
                const int stage = gt & 1;
                const int n = (tile0 + it) * 128 + row;
                const bool valid = n < nb;
                uint8_t* dst = sAQ + stage * 16384;
                stamp2(50);
                // ---- stage the input tile (sYA is free: the previous tile's projection MMA was waited for below)
                if (!DIN64) {
                    float x[4] = {0.f, 0.f, 0.f, 0.f};
                    if (valid) {
                        const float* xp = P.X32 + ((size_t)cloud * P.N + n) * P.d_in;
                        for (int k = 0; k < P.d_in; ++k) x[k] = __ldg(xp + k);
                    }
                    float cols[16];
                    split_x16(x, cols);
                    st_shared_8bf16(sYA + row * 16, cols);
                    st_shared_8bf16(sYA + 2048 + row * 16, cols + 8);
                } else {
                    const uint4* src = reinterpret_cast<const uint4*>(P.Y16in + ((size_t)cloud * P.N + (valid ? n : 0)) * 64);
                    uint4 yv[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) yv[c] = valid ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(sYA + c * 2048 + row * 16) = yv[c];
                }
                fence_async_smem();
                fence_before_sync();
                warp_arrive(ya_full);
                stamp2(51);
                const int buf = gt % A3_NBUF;
                if (warp == 8) {
                    // one producer warp issues the Q projection (elected lane) into the tile's accumulator, which is free
                    // once the epilogue of A3_NBUF tiles ago has read it
                    mbar_wait(ya_full, gt & 1);
                    if (gt >= A3_NBUF) mbar_wait(&oq_free[buf], ((gt / A3_NBUF) - 1) & 1);
                    stamp2(56);
                    fence_after_sync();
                    if (elect_one()) {
                        const uint32_t yab = smem_u32(sYA), wq = smem_u32(sWq);
#pragma unroll
                        for (int ks = 0; ks < (DIN64 ? 4 : 1); ++ks)
                            mma_ss(tmem_addr(tb, 0, A3_OQ + 64 * buf), smem_desc(yab + ks * 4096, 2048, 128),
                                   smem_desc(wq + ks * 2048, 1024, 128), idesc_bf16(128, 64, 0, 0), ks > 0);
                        mma_commit(qp_done);
                    }
                    __syncwarp();
                }
                mbar_wait(qp_done, gt & 1);
                fence_after_sync();
                stamp2(52);
                if (gt >= 2) mbar_wait(&aq_empty[stage], ((gt >> 1) - 1) & 1);
                stamp2(53);
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_addr(tb, 32 * quad, A3_OQ + 64 * buf + c0), v);
                    tmem_ld_wait32(v);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(v[8 * q + j]) + sBq[c0 + 8 * q + j];      // scale: K image
                        st_shared_8bf16(dst + (c0 / 8 + q) * 2048 + row * 16, o);
                    }
                }
                stamp2(54);
                fence_async_smem();
                fence_before_sync();
                warp_arrive(&aq_full[stage]);
            }
        }
    } else {
        reg_inc<144>();
        // =================================================================== softmax warpgroups (chains 2g, 2g+1)
        const int g = warp >> 2, quad = warp & 3;
        const uint32_t lane_base = 32 * quad;
        uint32_t ph_s[2] = {0, 0};
#ifdef PCA_TIMELINE
        long long* tl = (P.timeline != nullptr && blockIdx.x == 0 && warp == 0 && lane == 0) ? P.timeline : nullptr;
        int tl_n = 0;
        auto stamp = [&](int tag) {
            if (tl != nullptr && tl_n < 1000) { tl[2 * tl_n] = tag; tl[2 * tl_n + 1] = clock64(); ++tl_n; }
        };
#else
        auto stamp = [&](int) {};
#endif
        uint32_t va[32], vb[32];          // scores of the current item; refilled with the next item's while it finishes
        // one 64-key item of chain 2g+j.  Its scores are already in flight into va / vb.  The next item (other chain of this
        // warpgroup) was issued by its MMA thread a whole item ago, so its scores are prefetched as soon as va / vb free up.
        auto softmax_item = [&](const int j, const bool live, const bool has_next) {
            const int c = 2 * g + j, cn = 2 * g + (j ^ 1);
            const uint32_t sbase = tmem_addr(tb, lane_base, A3_S + 64 * c);
            const uint32_t snext = tmem_addr(tb, lane_base, A3_S + 64 * cn);
            uint32_t pk[16];
            stamp(20);
            tmem_ld_wait64(va, vb);
            stamp(25);
            if (live) {
                const float mx = max64(va, vb);
                const float2 neg2 = make_float2(-mx, -mx);
                float2 sum2 = make_float2(0.f, 0.f);
                exp_keep32(va, neg2, sum2);
                exp_keep32(vb, neg2, sum2);
                const float inv = __fdividef(1.f, sum2.x + sum2.y);
                const float2 inv2 = make_float2(inv, inv);
                scale_pack32(va, inv2, pk);
                tmem_st16(sbase, pk);
                stamp(26);
                if (has_next) {
                    mbar_wait(&s_full[cn], ph_s[j ^ 1]);
                    ph_s[j ^ 1] ^= 1;
                    fence_after_sync();
                    tmem_ld32(snext, va);
                }
                stamp(24);
                scale_pack32(vb, inv2, pk);
                tmem_st16(sbase + 16, pk);
                if (has_next) tmem_ld32(snext + 32, vb);
                tmem_st_wait();
                stamp(27);
            } else if (has_next) {
                mbar_wait(&s_full[cn], ph_s[j ^ 1]);
                ph_s[j ^ 1] ^= 1;
                fence_after_sync();
                tmem_ld32(snext, va);
                tmem_ld32(snext + 32, vb);
            }
            fence_before_sync();
            warp_arrive(&p_ready[c]);
        };
#if PCA_A3_PREFETCH
        if (blockIdx.x < n_work) {
            mbar_wait(&s_full[2 * g], ph_s[0]);
            ph_s[0] ^= 1;
            fence_after_sync();
            tmem_ld32(tmem_addr(tb, lane_base, A3_S + 64 * (2 * g)), va);
            tmem_ld32(tmem_addr(tb, lane_base, A3_S + 64 * (2 * g) + 32), vb);
        }
        for (int w = blockIdx.x; w < n_work; w += wstep) {
            int cloud, tile0, nb;
            const int ntiles = work_tiles(w, cloud, tile0, nb);
            const bool more_work = w + wstep < n_work;
            for (int it = 0; it < ntiles; ++it) {
                const bool live = (tile0 + it) * 128 + 32 * quad < nb;     // warps whose 32 rows are all padding idle
                softmax_item(0, live, true);
                softmax_item(1, live, true);
                softmax_item(0, live, true);
                softmax_item(1, live, it + 1 < ntiles || more_work);
            }
        }
#else
        (void)softmax_item;
        bool stagger = (PCA_A3_STAGGER != 0) && g == 1;      // experiment: start the two warpgroups in anti-phase
        for (int w = blockIdx.x; w < n_work; w += wstep) {
            int cloud, tile0, nb;
            const int ntiles = work_tiles(w, cloud, tile0, nb);
            for (int it = 0; it < ntiles; ++it) {
                const bool live = (tile0 + it) * 128 + 32 * quad < nb;     // warps whose 32 rows are all padding idle
#pragma unroll
                for (int step = 0; step < 4; ++step) {
                    const int j = step & 1;
                    const int c = 2 * g + j;
                    const uint32_t sbase = tmem_addr(tb, lane_base, A3_S + 64 * c);
                    stamp(20);
                    mbar_wait(&s_full[c], ph_s[j]);
                    ph_s[j] ^= 1;
                    fence_after_sync();
                    if (stagger) { __nanosleep(PCA_A3_STAGGER); stagger = false; }
                    stamp(24);
                    if (live) {
                        uint32_t pk[16];
                        tmem_ld32(sbase, va);
                        tmem_ld32(sbase + 32, vb);
                        tmem_ld_wait64(va, vb);
                        stamp(25);
#ifdef PCA_A3_NOEXP
                        // EXPERIMENT ONLY: no softmax arithmetic at all (what does the rest of the pipeline cost?)
#pragma unroll
                        for (int q = 0; q < 16; ++q) pk[q] = pack_bf16(__uint_as_float(va[2 * q]), __uint_as_float(va[2 * q + 1]));
                        tmem_st16(sbase, pk);
#pragma unroll
                        for (int q = 0; q < 16; ++q) pk[q] = pack_bf16(__uint_as_float(vb[2 * q]), __uint_as_float(vb[2 * q + 1]));
                        tmem_st16(sbase + 16, pk);
                        tmem_st_wait();
                        fence_before_sync();
                        warp_arrive(&p_ready[c]);
                        continue;
#endif
#ifdef PCA_A3_NOMAX
                        const float mx = 0.f;      // EXPERIMENT ONLY (upper bound of folding the shift into the MMA)
#else
                        const float mx = max64(va, vb);
#endif
                        const float2 neg2 = make_float2(-mx, -mx);
                        float2 sum2 = make_float2(0.f, 0.f);
                        exp_keep32(va, neg2, sum2);
                        exp_keep32(vb, neg2, sum2);
                        const float inv = __fdividef(1.f, sum2.x + sum2.y);
                        const float2 inv2 = make_float2(inv, inv);
                        scale_pack32(va, inv2, pk);
                        tmem_st16(sbase, pk);
                        scale_pack32(vb, inv2, pk);
                        tmem_st16(sbase + 16, pk);
                        stamp(26);
                        tmem_st_wait();
                        stamp(27);
                    }
                    fence_before_sync();
                    warp_arrive(&p_ready[c]);
                }
            }
        }
#endif
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tb, 512);
}

// ====================================================================================== apply kernel, fourth generation
// Same arithmetic as mab_apply3_tc_kernel; what changes is how much softmax work is in flight and who moves the data.
// Measured with tools/mb_softmax.cu (the bare per-item instruction stream, no MMA): 8 softmax warps peak at 13.1
// elements/clk/SM, 16 at 14.9 (MUFU limit 16); apply3's two warpgroups with one score buffer per head pair reach 9.4.
//   * THREE softmax warpgroups (12 warps, 3 per scheduler), each with its OWN chain: a 64-column score buffer, a 32-column
//     probability buffer and an MMA-issuing warp.  Head pairs are dealt round-robin (pair q = 4 * tile + p goes to chain
//     q % 3) and every chain walks its own sequence at its own pace -- the schedulers favour the higher warp ids, so a
//     common item order would run all warpgroups at the speed of the lowest one (measured: 8000 cycles per tile).
//   * Scores and probabilities live in SEPARATE TMEM regions.  The score buffer is released the moment the warpgroup has the
//     scores in registers, so Q K^T of the chain's next head runs under this head's arithmetic; the probabilities of head n
//     are written after s_full(n+1) has been seen, which (same issuing thread, in-order tensor pipe) implies P V (n-1) has
//     drained the probability buffer.  One barrier wait per head on the softmax side.
//   * Three tile accumulators; the fc_o output F is written OVER the tile accumulator (the epilogue keeps O1 in registers
//     from the read that builds the bf16 fc_o operand), so no F columns exist and the accumulator is read once.
//   * Biases ride in the MMAs: fc_q / fc_o get a fifth K step against a constant "ones" operand whose weight rows hold the
//     bias as bf16 hi + lo (d_in <= 4: the two spare columns of the split-bf16 K=16 image); the softmax scale is folded into
//     the K image by the finalize kernel.  The producer / epilogue warps therefore do no shared-memory loads at all (they
//     were latency-bound on them: 64 broadcast LDS per phase).
//   * A LOADER warp brings the operands in with the copy engine: the cloud's K / V images by one 32 KB cp.async.bulk, the
//     bf16 input tile by eight 2-D TMA boxes of 128 rows x 16 bytes (cp.async.bulk.tensor), which land directly in the
//     canonical no-swizzle operand layout [chunk][row][16 B].
//   * Registers live per scheduler (16 K each, five warps): 3 softmax x 104 + one producer/epilogue warp x 128 + one light
//     warp x 40 = 480 = 5 x 96.  The producer work (projection -> bf16 query operand) and the epilogue share four warps:
//     iteration j prepares tile j + 1 (a whole tile ahead of its first Q K^T) and then finishes tile j - 1.
// 20 warps: 0-11 softmax (warpgroup g = warps 4g..4g+3), 12-15 producer + epilogue, 17 loader, 16 / 18 / 19 MMA chains 0 / 1 / 2.
constexpr int A4_THREADS = 20 * 32;
// Thread-side arrivals: EVERY lane arrives (barrier counts 128 = four warps) -- measured 3.6 % faster than the
// __syncwarp + elected-lane arrival (no warp sync, no divergent branch on the softmax warps' critical path).
#ifndef PCA_A4_WARPARRIVE
#define A4_SM_ARRIVE(bar) mbar_arrive(bar)
#define A4_SM_COUNT 128
#else
#define A4_SM_ARRIVE(bar) warp_arrive(bar)
#define A4_SM_COUNT 4
#endif
#define A4_FENCE_BEFORE() fence_before_sync()
#define A4_LONG_WAIT(bar, par) mbar_wait(bar, par)
constexpr uint32_t A4_S = 0, A4_P = 192, A4_OQ = 320;   // 3 x 64 scores | 3 x 32 probabilities (+32 spare) | 3 x 64 tile accumulators
struct A4Smem {
    static constexpr int IMG = 0;                 // 2 x (K image 16384 | V image 16384), one per work item in flight
    static constexpr int WO = 65536;              // fc_o B operand (N=64, K=80: weights + bias step)
    static constexpr int WQ = WO + 10240;         // fc_q B operand: (N=64, K=80) or the split-bf16 K=16 image
    static constexpr int AQ = WQ + 10240;         // 2 stages x 16384: bf16 queries (A operand of Q K^T)
    static constexpr int YA = AQ + 32768;         // 2 stages x 16384: input tile, Y (128 x 64 bf16) or the split-bf16 X columns
    static constexpr int O1 = YA + 32768;         // O1 tile as the bf16 A operand of fc_o
    static constexpr int ONES = O1 + 16384;       // constant A operand of the bias K step: columns 0, 1 = 1, the other 14 = 0
    static constexpr int BARS = ONES + 4096;
    static constexpr int TOTAL = BARS + 48 * 8 + 16;
};

template <bool DIN64>
__global__ void __launch_bounds__(A4_THREADS, 1) mab_apply4_tc_kernel(const __grid_constant__ AParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sImg = smem + A4Smem::IMG;
    uint8_t* sWo = smem + A4Smem::WO;
    uint8_t* sWq = smem + A4Smem::WQ;
    uint8_t* sAQ = smem + A4Smem::AQ;
    uint8_t* sYA = smem + A4Smem::YA;
    uint8_t* sO1 = smem + A4Smem::O1;
    uint8_t* sOnes = smem + A4Smem::ONES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A4Smem::BARS);
    uint64_t* aq_full = bars;          // [2] count 4 (producer warps)
    uint64_t* aq_empty = bars + 2;     // [2] count 4 (one commit per head pair: its scores are computed)
    uint64_t* s_full = bars + 4;       // [3] count 1                          -- per chain
    uint64_t* s_free = bars + 7;       // [3] count 4 (warps of the warpgroup: scores are in registers)
    uint64_t* p_ready = bars + 10;     // [3] count 4 (warps of the warpgroup)
    uint64_t* p_free = bars + 14;      // [3] count 1 (commit of P V; only the last head of a chain waits on it)
    uint64_t* o_full = bars + 18;      // [3] count 4 (one commit per head pair) -- per tile accumulator
    uint64_t* oq_free = bars + 21;     // [3] count 4 (epilogue warps)         -- per tile accumulator
    uint64_t* ya_full = bars + 24;     // [2] count 1 (loader; + the TMA byte count)
    uint64_t* ya_empty = bars + 26;    // [2] count 1 (commit of the projection MMA)
    uint64_t* qp_done = bars + 28;     // count 1
    uint64_t* o1_ready = bars + 29;    // count 4 (epilogue warps)
    uint64_t* f_full = bars + 30;      // count 1
    uint64_t* img_full = bars + 32;    // [2] count 1 (loader; + the bulk copy's byte count)
    uint64_t* img_empty = bars + 34;   // [2] count 3 (the chains, as each leaves the work item)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 48);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_work = P.n_work, wstep = gridDim.x;
#ifdef PCA_TIMELINE
    // debug: softmax warp 0, producer/epilogue warp 12, chain warps 16 and 19 of CTA 0 record clock stamps
    long long* tl2 = nullptr;
    int tl2_n = 0;
    if (P.timeline != nullptr && blockIdx.x == 0 && lane == 0) {
        if (warp == 0) tl2 = P.timeline;
        else if (warp == 12) tl2 = P.timeline + 2000;
        else if (warp == 8) tl2 = P.timeline + 4000;
        else if (warp == 16) tl2 = P.timeline + 6000;
    }
    auto stamp2 = [&](int tag) {
        if (tl2 != nullptr && tl2_n < 1000) { tl2[2 * tl2_n] = tag; tl2[2 * tl2_n + 1] = clock64(); ++tl2_n; }
    };
#else
    auto stamp2 = [&](int) {};
#endif
    // tiles of work item w; nb = valid points of its cloud (variable-size sets: rows past nb are padding)
    auto work_tiles = [&](int w, int& cloud, int& tile0, int& nb) {
        cloud = w / P.nsplit;
        const int split = w - cloud * P.nsplit;
        tile0 = split * P.tiles_per_split;
        nb = main_points(P.counts ? max(1, min(P.N, __ldg(P.counts + cloud))) : P.N, P.tail_max);
        return max(0, min((nb + 127) >> 7, tile0 + P.tiles_per_split) - tile0);
    };

    copy_to_smem(sWo, P.Wo16, 10240);
    copy_to_smem(sWq, P.Wq16, DIN64 ? 10240 : 2048);
    for (int i = threadIdx.x; i < 256; i += blockDim.x)          // [2 chunks][128 rows][16 B]: (1, 1, 0, ...) | 0
        *reinterpret_cast<uint4*>(sOnes + i * 16) = (i < 128) ? make_uint4(0x3F803F80u, 0, 0, 0) : make_uint4(0, 0, 0, 0);
    if (warp == 16) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&aq_full[i], A4_SM_COUNT); mbar_init(&aq_empty[i], 4);
            mbar_init(&img_full[i], 1); mbar_init(&img_empty[i], 3);
            mbar_init(&ya_full[i], 1); mbar_init(&ya_empty[i], 1);
        }
        for (int i = 0; i < 3; ++i) {
            mbar_init(&s_full[i], 1); mbar_init(&s_free[i], A4_SM_COUNT);
            mbar_init(&p_ready[i], A4_SM_COUNT); mbar_init(&p_free[i], 1);
            mbar_init(&o_full[i], 4); mbar_init(&oq_free[i], A4_SM_COUNT);
        }
        mbar_init(qp_done, 1);
        mbar_init(o1_ready, A4_SM_COUNT);
        mbar_init(f_full, 1);
        fence_barrier_init();
    }
    if (DIN64 && threadIdx.x == 17 * 32) tma_prefetch_desc(&P.tmapY);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;

    if (warp == 17) {
        reg_dec<40>();
        // =================================================================== loader: K / V images and input tiles by TMA
        int gt = 0, wl = 0;
        for (int w = blockIdx.x; w < n_work; w += wstep, ++wl) {
            int cloud, tile0, nb;
            const int ntiles = work_tiles(w, cloud, tile0, nb);
            if (wl >= 2) A4_LONG_WAIT(&img_empty[wl & 1], ((wl >> 1) - 1) & 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&img_full[wl & 1], 32768);
                bulk_copy_g2s(sImg + (wl & 1) * 32768, P.KVblk + (size_t)cloud * 32768, 32768, &img_full[wl & 1]);
            }
            __syncwarp();
            for (int it = 0; it < ntiles; ++it, ++gt) {
                const int stage = gt & 1;
                uint8_t* dst = sYA + stage * 16384;
                if (gt >= 2) A4_LONG_WAIT(&ya_empty[stage], ((gt >> 1) - 1) & 1);
                if (DIN64) {
                    // rows past the cloud's valid points (padding, or the next cloud's rows) only feed MMA rows whose
                    // results are never stored; rows past the end of the tensor are zero-filled by the copy engine
                    if (lane == 0) {
                        const long long r0 = (long long)cloud * P.N + (long long)(tile0 + it) * 128;
                        mbar_arrive_expect_tx(&ya_full[stage], 16384);
                        tma_load_3d(dst, &P.tmapY, 0, (int)r0, 0, &ya_full[stage]);     // one box: (8 elements, 128 rows, 8 chunks)
                    }
                    __syncwarp();
                } else {
                    // d_in <= 4: the split-bf16 columns of 4 rows per lane (128 x 16 bf16 = 4 KB); columns 12, 13 = 1
                    // meet the bias rows (hi, lo) of the weight image
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int row = 32 * rr + lane;
                        const int n = (tile0 + it) * 128 + row;
                        float x[4] = {0.f, 0.f, 0.f, 0.f};
                        if (n < nb) load_point(P.X32, P.src, P.N, P.d_in, (size_t)cloud, n, x);
                        float cols[16];
                        split_x16(x, cols);
                        cols[12] = 1.f;
                        cols[13] = 1.f;
                        st_shared_8bf16(dst + row * 16, cols);
                        st_shared_8bf16(dst + 2048 + row * 16, cols + 8);
                    }
                    fence_async_smem();
                    fence_before_sync();
                    warp_arrive(&ya_full[stage]);
                }
            }
        }
    } else if (warp >= 16) {
        reg_dec<40>();
        // =================================================================== MMA chain g (uniform control flow, one elected lane issues)
        // Head items of the chain in order: for every tile, the pairs p with (4 * tile + p) % 3 == g, two heads each.
        //   Q K^T (x) is issued as soon as the warpgroup has loaded the scores of the previous head (s_free), i.e. under
        //   that head's softmax arithmetic; P V (prev) follows when its probabilities are written (p_ready).
        const int g = warp == 16 ? 0 : warp - 17;
        const bool leader = elect_one();
        const uint32_t idesc_s = idesc_bf16(128, 64, 0, 0);
        const uint32_t idesc_pv = idesc_bf16(128, 16, 0, 1);
        const uint32_t img = smem_u32(sImg), aqb = smem_u32(sAQ);
        const uint32_t s_tm = tmem_addr(tb, 0, A4_S + 64 * g), p_tm = tmem_addr(tb, 0, A4_P + 32 * g);
        uint32_t par = 0;                    // parity of the chain's barriers for the head being issued (n & 1)
        bool have_prev = false;
        uint32_t prev_v = 0, prev_d = 0;     // P V of the previous head: V operand address, accumulator address
        int prev_buf = 0;
        bool prev_last = false;              // ... it completes its pair
        auto pv_prev = [&]() {
            // probabilities of the previous head (barrier parity: the one before `par`)
            stamp2(62);
            A4_LONG_WAIT(&p_ready[g], par ^ 1);
            fence_after_sync();
            stamp2(63);
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma_ts(prev_d, p_tm + ks * 8, smem_desc(prev_v + ks * 256, 128, 2048), idesc_pv, 1u);
                mma_commit(&p_free[g]);
                if (prev_last) mma_commit(&o_full[prev_buf]);      // one of the tile's four pairs is complete
            }
            __syncwarp();
        };
        int gt = 0, wl = 0, q3 = 0;          // q3 = (4 * gt) % 3: pair p belongs to chain (q3 + p) % 3
        for (int w = blockIdx.x; w < n_work; w += wstep, ++wl) {
            int cloud, tile0, nb;
            const int ntiles = work_tiles(w, cloud, tile0, nb);
            mbar_wait(&img_full[wl & 1], (wl >> 1) & 1);
            fence_after_sync();
            const uint32_t k_base = img + (wl & 1) * 32768, v_base = k_base + 16384;
            for (int it = 0; it < ntiles; ++it, ++gt) {
                stamp2(60);
                mbar_wait(&aq_full[gt & 1], (gt >> 1) & 1);
                fence_after_sync();
                stamp2(61);
                const uint32_t a_base = aqb + (gt & 1) * 16384;
                const int buf = gt % 3;
#pragma unroll 1
                for (int p = 0; p < 4; ++p) {
                    int r = q3 + p;
                    r = r >= 3 ? r - 3 : r;
                    r = r >= 3 ? r - 3 : r;
                    if (r != g) continue;
#pragma unroll 1
                    for (int hh = 0; hh < 2; ++hh) {
                        stamp2(64);
                        if (have_prev) {
                            mbar_wait(&s_free[g], par ^ 1);
                            fence_after_sync();
                        }
                        stamp2(65);
                        if (leader) {
                            mma_ss(s_tm, smem_desc(a_base + 2 * p * 2048, 2048, 128), smem_desc(k_base + p * 4096 + hh * 1024, 2048, 128), idesc_s, 0);
                            mma_commit(&s_full[g]);
                            if (hh == 1) mma_commit(&aq_empty[gt & 1]);      // the pair's query operand has been consumed
                        }
                        __syncwarp();
                        stamp2(66);
                        if (have_prev) pv_prev();
                        par ^= 1;
                        have_prev = true;
                        prev_v = v_base + p * 4096 + hh * 1024;
                        prev_d = tmem_addr(tb, 0, A4_OQ + 64 * buf + 16 * p);
                        prev_buf = buf;
                        prev_last = hh == 1;
                    }
                }
                q3 = q3 == 2 ? 0 : q3 + 1;       // (4 (gt + 1)) % 3 = (q3 + 1) % 3
            }
            // leaving the work item: its last P V may still be pending -- it reads the V image, so it goes first
            if (have_prev) {
                pv_prev();
                have_prev = false;
                // the next head of this chain starts a fresh hand-shake: s_free of the head just retired is consumed here
                mbar_wait(&s_free[g], par ^ 1);
            }
            if (leader) mma_commit(&img_empty[wl & 1]);
            __syncwarp();
        }
    } else if (warp >= 12) {
        reg_inc<128>();
        // =================================================================== producer + epilogue warps (thread = point row)
        // Iteration j prepares tile k = j + 1 and finishes tile e = j - 1.  Both MMA round trips (Q projection of k, fc_o of
        // e) are issued ahead of the register work that does not depend on them:
        //   [0] warp 12 issues the Q projection of k                      (inputs ready since the previous iteration)
        //   [1] O1(e): TMEM -> registers (kept) -> bf16 A operand of fc_o in shared memory
        //   [2] warp 12 issues fc_o(e)
        //   [3] projection(k): TMEM -> bf16 query operand               (under the fc_o MMA)
        //   [4] Y(e) = O1 + relu(F) -> bf16 tile in shared memory -> eight TMA box stores (ragged last tiles: per-row stores)
        const int quad = warp & 3;
        const int row = 32 * quad + lane;
        const uint32_t lane_base = 32 * quad;
        int it_w = blockIdx.x, it_i = 0, it_ntiles = 0, it_cloud = 0, it_tile0 = 0, it_nb = 0;
        if (it_w < n_work) it_ntiles = work_tiles(it_w, it_cloud, it_tile0, it_nb);
        int e_cloud = 0, e_tile = 0, e_nb = 0, m_cloud = 0, m_tile = 0, m_nb = 0, p_cloud = 0, p_tile = 0, p_nb = 0;   // tiles j-1, j, j+1
        bool more = true;
        bool stored = false;                  // a TMA store of the staging tile may still be reading it
        int n_prod = 0;                       // tiles produced so far
        for (int j = -1;; ++j) {
            if (more) {
                while (it_w < n_work && it_i == it_ntiles) {
                    it_w += wstep;
                    it_i = 0;
                    if (it_w < n_work) it_ntiles = work_tiles(it_w, it_cloud, it_tile0, it_nb);
                }
                more = it_w < n_work;
            }
            const bool have_k = more;
            const bool have_e = j >= 1 && j - 1 < n_prod;
            const int k = j + 1, kstage = k & 1, kbuf = k % 3;
            const int e = j - 1, ebuf = (e + 3) % 3;
            if (have_k) {
                p_cloud = it_cloud; p_tile = it_tile0 + it_i; p_nb = it_nb;
                ++it_i;
                n_prod = k + 1;
                stamp2(50);
                if (warp == 12) {
                    // [0] the Q projection (elected lane) into the tile's accumulator, free once the epilogue of three tiles
                    // ago has read its fc_o output on all four warps
                    mbar_wait(&ya_full[kstage], (k >> 1) & 1);
                    if (k >= 3) mbar_wait(&oq_free[kbuf], ((k / 3) - 1) & 1);
                    fence_after_sync();
                    if (elect_one()) {
                        const uint32_t yab = smem_u32(sYA) + kstage * 16384, wq = smem_u32(sWq), ones = smem_u32(sOnes);
                        const uint32_t d = tmem_addr(tb, 0, A4_OQ + 64 * kbuf);
                        if (DIN64) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                mma_ss(d, smem_desc(yab + ks * 4096, 2048, 128), smem_desc(wq + ks * 2048, 1024, 128), idesc_bf16(128, 64, 0, 0), ks > 0);
                            mma_ss(d, smem_desc(ones, 2048, 128), smem_desc(wq + 8192, 1024, 128), idesc_bf16(128, 64, 0, 0), 1u);
                        } else {
                            mma_ss(d, smem_desc(yab, 2048, 128), smem_desc(wq, 1024, 128), idesc_bf16(128, 64, 0, 0), 0u);
                        }
                        mma_commit(qp_done);
                        mma_commit(&ya_empty[kstage]);
                    }
                    __syncwarp();
                }
                stamp2(51);
            }
            uint32_t o1[64];                    // O1(e) (fp32 bits): kept in registers across the fc_o MMA
            const int n = e_tile * 128 + row;
            const bool valid = have_e && n < e_nb;
            const bool live = have_e && e_tile * 128 + 32 * quad < e_nb;
            const bool full_tile = e_tile * 128 + 128 <= e_nb;
            const uint32_t oq = tmem_addr(tb, lane_base, A4_OQ + 64 * ebuf);
            if (have_e) {
                // [1]
                stamp2(40);
                A4_LONG_WAIT(&o_full[ebuf], (e / 3) & 1);
                fence_after_sync();
                stamp2(41);
                if (stored) {
                    // the staging tile is free once the previous tile's TMA store has read it (issued a whole tile ago)
                    if (warp == 12 && lane == 0) bulk_wait_group_read0();
                    named_bar_sync(8, 128);
                    stored = false;
                }
                if (live) {
                    tmem_ld32(oq, o1);
                    tmem_ld32(oq + 32, o1 + 32);
                    tmem_ld_wait64(o1, o1 + 32);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        uint4 u;
                        u.x = pack_bf16(__uint_as_float(o1[8 * q + 0]), __uint_as_float(o1[8 * q + 1]));
                        u.y = pack_bf16(__uint_as_float(o1[8 * q + 2]), __uint_as_float(o1[8 * q + 3]));
                        u.z = pack_bf16(__uint_as_float(o1[8 * q + 4]), __uint_as_float(o1[8 * q + 5]));
                        u.w = pack_bf16(__uint_as_float(o1[8 * q + 6]), __uint_as_float(o1[8 * q + 7]));
                        *reinterpret_cast<uint4*>(sO1 + q * 2048 + row * 16) = u;
                    }
                }
                stamp2(42);
                fence_async_smem();
                fence_before_sync();
                A4_SM_ARRIVE(o1_ready);
                if (warp == 12) {
                    // [2] fc_o (elected lane): F = O1 Wo^T + bo written OVER the tile accumulator, which all four warps have
                    // read by now (their arrivals above follow their tcgen05.wait::ld)
                    mbar_wait(o1_ready, e & 1);
                    fence_after_sync();
                    if (elect_one()) {
                        const uint32_t wo = smem_u32(sWo), o1b = smem_u32(sO1), ones = smem_u32(sOnes);
                        const uint32_t d = tmem_addr(tb, 0, A4_OQ + 64 * ebuf);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            mma_ss(d, smem_desc(o1b + ks * 4096, 2048, 128), smem_desc(wo + ks * 2048, 1024, 128), idesc_bf16(128, 64, 0, 0), ks > 0);
                        mma_ss(d, smem_desc(ones, 2048, 128), smem_desc(wo + 8192, 1024, 128), idesc_bf16(128, 64, 0, 0), 1u);
                        mma_commit(f_full);
                    }
                    __syncwarp();
                }
                stamp2(43);
            }
            if (have_k) {
                // [3] projection of tile k -> bf16 query operand (two 32-column halves: O1(e) stays live)
                uint8_t* dst = sAQ + kstage * 16384;
                mbar_wait(qp_done, k & 1);
                fence_after_sync();
                stamp2(52);
                if (k >= 2) mbar_wait(&aq_empty[kstage], ((k >> 1) - 1) & 1);
                stamp2(53);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t v[32];
                    tmem_ld32(tmem_addr(tb, lane_base, A4_OQ + 64 * kbuf + 32 * hf), v);
                    tmem_ld_wait32(v);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 u;
                        u.x = pack_bf16(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
                        u.y = pack_bf16(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
                        u.z = pack_bf16(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
                        u.w = pack_bf16(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
                        *reinterpret_cast<uint4*>(dst + (4 * hf + q) * 2048 + row * 16) = u;
                    }
                }
                stamp2(54);
                fence_async_smem();
                fence_before_sync();
                A4_SM_ARRIVE(&aq_full[kstage]);
            }
            if (have_e) {
                // [4] Y = O1 + relu(F)      (F already holds the bias)
                mbar_wait(f_full, e & 1);
                fence_after_sync();
                stamp2(44);
                if (live) {
                    __nv_bfloat16* dstp = P.Yout + ((size_t)e_cloud * P.N + (valid ? n : 0)) * 64;
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t f[32];
                        tmem_ld32(oq + 32 * hf, f);
                        tmem_ld_wait32(f);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 u;
                            uint32_t* uw = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
                            for (int x = 0; x < 8; x += 2) {
                                const int c = 32 * hf + 8 * q + x;
                                const float y0 = __uint_as_float(o1[c]) + fmaxf(__uint_as_float(f[8 * q + x]), 0.f);
                                const float y1 = __uint_as_float(o1[c + 1]) + fmaxf(__uint_as_float(f[8 * q + x + 1]), 0.f);
                                uw[x >> 1] = pack_bf16(y0, y1);
                            }
                            if (full_tile) *reinterpret_cast<uint4*>(sO1 + (4 * hf + q) * 2048 + row * 16) = u;
                            else if (valid) *reinterpret_cast<uint4*>(dstp + 32 * hf + 8 * q) = u;
                        }
                    }
                }
                if (full_tile) {
                    // the tile leaves as ONE (8 elements, 128 rows, 8 chunks) box from the canonical layout (fc_o has finished reading it)
                    fence_async_smem();
                    named_bar_sync(8, 128);
                    if (warp == 12 && lane == 0) {
                        const long long r0 = (long long)e_cloud * P.N + (long long)e_tile * 128;
                        tma_store_3d(&P.tmapYo, sO1, 0, (int)r0, 0);       // one box (a TMA instruction costs its thread ~90 cycles)
                        bulk_commit_group();
                    }
                    stored = true;
                }
                stamp2(45);
                fence_before_sync();
                A4_SM_ARRIVE(&oq_free[ebuf]);
            }
            if (!more && (n_prod == 0 || j - 1 >= n_prod - 1)) break;
            e_cloud = m_cloud; e_tile = m_tile; e_nb = m_nb;
            m_cloud = p_cloud; m_tile = p_tile; m_nb = p_nb;
        }
        if (warp == 12 && lane == 0) bulk_wait_group0();       // the last tile's store has been performed before the CTA exits
    } else {
        reg_inc<104>();
        // =================================================================== softmax warpgroup g (chain g)
        const int g = warp >> 2, quad = warp & 3;
        const uint32_t lane_base = 32 * quad;
        const uint32_t sbase = tmem_addr(tb, lane_base, A4_S + 64 * g);
        const uint32_t pbase = tmem_addr(tb, lane_base, A4_P + 32 * g);
        uint32_t par = 0;                     // parity of the chain's barriers for the current head (n & 1)
        bool have_scores = false;             // s_full of the current head has already been waited for (look-ahead below)
        bool first_head = true;               // the chain's very first head finds the probability buffer free
        int q3 = 0;
        for (int w = blockIdx.x; w < n_work; w += wstep) {
            int cloud, tile0, nb;
            const int ntiles = work_tiles(w, cloud, tile0, nb);
            // (the chain restarts its hand-shake at every work item boundary, so the look-ahead never crosses one)
            for (int it = 0; it < ntiles; ++it) {
                const bool live = (tile0 + it) * 128 + 32 * quad < nb;     // warps whose 32 rows are all padding idle
#pragma unroll 1
                for (int p = 0; p < 4; ++p) {
                    int r = q3 + p;
                    r = r >= 3 ? r - 3 : r;
                    r = r >= 3 ? r - 3 : r;
                    if (r != g) continue;
                    // is there a later pair of this chain in this work item?  pairs of the chain are 3 apart in q = 4 * it + p
                    const bool pair_follows = 4 * it + p + 3 < 4 * ntiles;
#pragma unroll 1
                    for (int hh = 0; hh < 2; ++hh) {
                        const bool has_next = hh == 0 || pair_follows;
                        uint32_t va[32], vb[32], pk[16];
                        float inv = 0.f;
                        stamp2(20);
                        if (!have_scores) {
                            mbar_wait(&s_full[g], par);
                            fence_after_sync();
                        }
                        stamp2(24);
                        if (live) {
                            tmem_ld32(sbase, va);
                            tmem_ld32(sbase + 32, vb);
                            tmem_ld_wait64(va, vb);
                        }
                        // the scores are in registers: Q K^T of the chain's next head may overwrite the buffer
                        A4_FENCE_BEFORE();
                        A4_SM_ARRIVE(&s_free[g]);
                        stamp2(25);
                        if (live) {
                            const float mx = max64(va, vb);
                            const float2 neg2 = make_float2(-mx, -mx);
                            float2 sum2 = make_float2(0.f, 0.f);
#ifndef PCA_A4_NOEXP
                            exp_keep32(va, neg2, sum2);
                            exp_keep32(vb, neg2, sum2);
#else
                            sum2 = make_float2(1.f, 1.f);      // EXPERIMENT ONLY: what does the pipeline cost without the exponentials?
#endif
                            inv = __fdividef(1.f, sum2.x + sum2.y);
                        }
                        stamp2(21);
                        // the probability buffer still holds the previous head's P until its P V has run: the chain issues
                        // that P V before the NEXT head's Q K^T, so s_full(next) covers it; the chain's last head of a work
                        // item waits for the P V commit itself
                        if (has_next) {
                            mbar_wait(&s_full[g], par ^ 1);
                            fence_after_sync();
                            have_scores = true;
                        } else {
                            if (!first_head) {
                                mbar_wait(&p_free[g], par ^ 1);
                                fence_after_sync();
                            }
                            have_scores = false;
                        }
                        first_head = false;
                        stamp2(22);
                        if (live) {
                            const float2 inv2 = make_float2(inv, inv);
                            scale_pack32(va, inv2, pk);
                            tmem_st16(pbase, pk);
                            scale_pack32(vb, inv2, pk);
                            tmem_st16(pbase + 16, pk);
                            stamp2(26);
                            tmem_st_wait();
                            stamp2(27);
                        }
                        A4_FENCE_BEFORE();
                        A4_SM_ARRIVE(&p_ready[g]);
                        par ^= 1;
                    }
                }
                q3 = q3 == 2 ? 0 : q3 + 1;
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------ apply kernel: leftover points
// MAB(Q = one leftover point, K = H) in fp32 on CUDA cores (tail rule): one 64-thread block per leftover point, thread f <->
// feature f = head h * 8 + d.  Same arithmetic as mab_apply3_tc_kernel with the cloud's bf16 K / V images as keys.
struct ATailParams {
    const float* X32;             // (B, N, dq) fp32              [dq <= 4]
    const __nv_bfloat16* Y16in;   // (B, N, 64) bf16              [dq == 64]
    int N, dq;
    const int* counts;            // nullable (B)
    int tail_max;
    const uint8_t* KVblk;         // per cloud: K image 16384 B | V image 16384 B
    const float* WqT;             // fc_q^T k-major (dq, 64)
    const float* bq;
    const float* WoT;             // fc_o^T k-major (64, 64)
    const float* bo;
    __nv_bfloat16* Yout;          // (B, N, 64)
    PointSrc src;                 // [dq <= 4] alternative to X32
};
// 16-byte read-only load that does not allocate in L1: the K / V rows are streamed once, the fc_q / fc_o weights that
// every block of the SM re-reads stay resident
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__global__ void __launch_bounds__(64, 16) mab_apply_tail_kernel(const ATailParams P) {
    __shared__ __align__(16) float sX[64], sQ[64], sO1[64];
    const int cloud = blockIdx.x, j = blockIdx.y, f = threadIdx.x;
    const int nb = P.counts ? max(1, min(P.N, __ldg(P.counts + cloud))) : P.N;
    const int n0 = main_points(nb, P.tail_max);
    if (j >= nb - n0) return;
    const size_t rowi = (size_t)cloud * P.N + n0 + j;
    // the K and V rows of this thread's 8 keys (head h, keys 8*part .. 8*part+7) are fetched up front, in the same round as
    // the point itself: the kernel is a chain of dependent global loads otherwise
    const int h = f >> 3, part = f & 7, pr = h >> 1, ch = h & 1;
    const uint8_t* kimg = P.KVblk + (size_t)cloud * 32768 + pr * 4096 + ch * 2048 + (size_t)(ch * 64) * 16;
    const uint8_t* vimg = kimg + 16384;
    uint4 krow[8], vrow[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        krow[i] = ldg_stream16(kimg + (size_t)(8 * part + i) * 16);
        vrow[i] = ldg_stream16(vimg + (size_t)(8 * part + i) * 16);
    }
    if (P.Y16in == nullptr) {
        if (f == 0) {
            float xr[4] = {0.f, 0.f, 0.f, 0.f};
            load_point(P.X32, P.src, P.N, P.dq, (size_t)cloud, n0 + j, xr);
            for (int k = 0; k < P.dq; ++k) sX[k] = xr[k];
        }
    } else sX[f] = __bfloat162float(P.Y16in[rowi * 64 + f]);
    __syncthreads();
    // the kernel is bound by l1tex instructions (weight loads + broadcast reads of the staged row): the row is read as
    // float4, same products in the same order
    float q = __ldg(P.bq + f);
    if (P.dq == 64) {
#pragma unroll 4
        for (int k = 0; k < 64; k += 4) {
            const float4 x4 = *reinterpret_cast<const float4*>(&sX[k]);
            q = fmaf(x4.x, __ldg(P.WqT + k * 64 + f), q);
            q = fmaf(x4.y, __ldg(P.WqT + (k + 1) * 64 + f), q);
            q = fmaf(x4.z, __ldg(P.WqT + (k + 2) * 64 + f), q);
            q = fmaf(x4.w, __ldg(P.WqT + (k + 3) * 64 + f), q);
        }
    } else {
        for (int k = 0; k < P.dq; ++k) q = fmaf(sX[k], __ldg(P.WqT + k * 64 + f), q);
    }
    sQ[f] = q;
    __syncthreads();
    float sc[8], mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint4 u = krow[i];
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
        float d = 0.f;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const float2 kk = __bfloat1622float2(h2[qd]);
            d = fmaf(sQ[h * 8 + 2 * qd], kk.x, d);
            d = fmaf(sQ[h * 8 + 2 * qd + 1], kk.y, d);
        }
        sc[i] = d;                         // the K image carries the softmax scale
        mx = fmaxf(mx, sc[i]);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = exp2f(sc[i] - mx); sum += sc[i]; }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    // this thread's 8 keys against the 8 features of head h, then the sum over the 8 threads of the head
    float o8[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) o8[d] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float pw = sc[i] * inv;
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&vrow[i]);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const float2 vv = __bfloat1622float2(h2[qd]);
            o8[2 * qd] = fmaf(pw, vv.x, o8[2 * qd]);
            o8[2 * qd + 1] = fmaf(pw, vv.y, o8[2 * qd + 1]);
        }
    }
    float o = 0.f;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
#pragma unroll
        for (int x = 1; x < 8; x <<= 1) o8[d] += __shfl_xor_sync(0xffffffffu, o8[d], x);
        if (part == d) o = o8[d];
    }
    const float o1 = sQ[f] + o;
    sO1[f] = o1;
    __syncthreads();
    float fo = __ldg(P.bo + f);
#pragma unroll 4
    for (int k = 0; k < 64; k += 4) {
        const float4 o4 = *reinterpret_cast<const float4*>(&sO1[k]);
        fo = fmaf(o4.x, __ldg(P.WoT + k * 64 + f), fo);
        fo = fmaf(o4.y, __ldg(P.WoT + (k + 1) * 64 + f), fo);
        fo = fmaf(o4.z, __ldg(P.WoT + (k + 2) * 64 + f), fo);
        fo = fmaf(o4.w, __ldg(P.WoT + (k + 3) * 64 + f), fo);
    }
    P.Yout[rowi * 64 + f] = __float2bfloat16(o1 + fmaxf(fo, 0.f));
}

// ====================================================================================== pooled attention (PMA)
// PMA with one seed, computed on the UN-projected points (see prep_queries): per 128-point tile
//   S (128 rows = head x 16 copies, 128 points) = AqPool (128 x 64) . Ytile^T      [4 MMAs, K = 64]
//   online softmax over the points per row, P (bf16) written over the consumed score columns
//   Z (128 rows, 64 features) = P . Ytile                                            [8 MMAs, N = 64, MN-major B]
// so neither fc_k nor fc_v is ever applied per point; fc_v is applied to the 8 pooled 64-vectors in finalize_pool_kernel
// (sum_n p_n (Wv y_n + bv) = Wv (sum_n p_n y_n) + bv).  The kernel is a pure stream over Y (128 B per point):
// persistent CTAs, 3-stage tile ring filled by 4 producer warps, two softmax warpgroups taking alternate tiles.
struct PoolParams {
    const __nv_bfloat16* Y16;     // (B, N, 64)
    int N, tiles_total, tiles_per_split, nsplit, n_work;
    const int* counts;            // nullable (B): valid points per cloud
    int tail_max;                 // leftover points (N mod 128 <= tail_max) are merged in finalize_pool_kernel
    const uint8_t* Aq;            // AqPool image
    float* part;                  // (B, 2 nsplit, 8 heads, 66): m (log2 domain), l, Z[64]
    int* redo;                    // (n_work) flags of the streaming transposed kernel (pma_pool2_tc_kernel<false>), else unused
    long long* timeline;          // debug (PCA_TIMELINE builds, PCA_TL_POOL=1): clock64 stamps of CTA 0, one lane per role
    CUtensorMap tmapY;            // Y16 as an (8, B N, 8 chunks) bf16 tensor, box = (8, 128, 8) (transposed kernel only)
};
constexpr int POOL_STAGES = 3;
struct PoolSmem {
    static constexpr int AQ = 0;
    static constexpr int Y = 16384;                               // POOL_STAGES x 16384
    static constexpr int BARS = Y + POOL_STAGES * 16384;
    static constexpr int TOTAL = BARS + 16 * 8 + 16;
};
constexpr uint32_t PC_S = 0, PC_O = 256;      // 2 x 128 score columns | 2 x 64 output columns

__global__ void __launch_bounds__(TC_THREADS16, 1) pma_pool_tc_kernel(const PoolParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sAq = smem + PoolSmem::AQ;
    uint8_t* sY = smem + PoolSmem::Y;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PoolSmem::BARS);
    uint64_t* y_full = bars;           // [3] count 4 (producer warps)
    uint64_t* y_empty = bars + 3;      // [3] count 1 (commit)
    uint64_t* s_full = bars + 6;       // [2] count 1
    uint64_t* p_ready = bars + 8;      // [2] count 4
    uint64_t* o_full = bars + 10;      // [2] count 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_work = P.n_work, wstep = gridDim.x;
    // tiles of work item w; nb = valid points of its cloud (variable-size sets: rows past nb are padding)
    auto work_tiles = [&](int w, int& cloud, int& split, int& tile0, int& nb) {
        cloud = w / P.nsplit;
        split = w - cloud * P.nsplit;
        tile0 = split * P.tiles_per_split;
        nb = main_points(P.counts ? max(1, min(P.N, __ldg(P.counts + cloud))) : P.N, P.tail_max);
        return max(0, min((nb + 127) >> 7, tile0 + P.tiles_per_split) - tile0);
    };
    copy_to_smem(sAq, P.Aq, 16384);
    if (warp == 12) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 0) {
        for (int i = 0; i < POOL_STAGES; ++i) { mbar_init(&y_full[i], 128); mbar_init(&y_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 128); mbar_init(&o_full[i], 1); }
        fence_barrier_init();
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;

    if (warp >= 12) {
        reg_dec<40>();
        if (warp == 12) {
            // =================================================================== MMA issuer (tiles in order; S runs one tile ahead)
            // whole warp in uniform control flow, one elected lane issues
            const bool leader = elect_one();
            const uint32_t idesc_s = idesc_bf16(128, 128, 0, 0);
            const uint32_t idesc_pv = idesc_bf16(128, 64, 0, 1);
            const uint32_t aq = smem_u32(sAq), yb = smem_u32(sY);
            int total = 0;
            for (int w = blockIdx.x; w < n_work; w += wstep) { int a, b, c, e; total += work_tiles(w, a, b, c, e); }
            auto issue_s = [&](int t) {
                const int stage = t % POOL_STAGES;
                mbar_wait(&y_full[stage], (t / POOL_STAGES) & 1);
                fence_after_sync();
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        mma_ss(tmem_addr(tb, 0, PC_S + 128 * (t & 1)), smem_desc(aq + ks * 4096, 2048, 128),
                               smem_desc(yb + stage * 16384 + ks * 4096, 2048, 128), idesc_s, ks > 0);
                    mma_commit(&s_full[t & 1]);
                }
                __syncwarp();
            };
            if (total > 0) issue_s(0);
            for (int t = 0; t < total; ++t) {
                // S(t+1) overwrites the buffer whose P was consumed by PV(t-1): issued earlier by this warp, in order
                if (t + 1 < total) issue_s(t + 1);
                const int b = t & 1, stage = t % POOL_STAGES;
                mbar_wait(&p_ready[b], (t >> 1) & 1);
                fence_after_sync();
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        mma_ts(tmem_addr(tb, 0, PC_O + 64 * b), tmem_addr(tb, 0, PC_S + 128 * b + ks * 8),
                               smem_desc(yb + stage * 16384 + ks * 256, 128, 2048), idesc_pv, ks > 0);
                    mma_commit(&o_full[b]);
                    mma_commit(&y_empty[stage]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 8) {
        reg_dec<88>();
        // =================================================================== producer: Y tiles -> [chunk][row][16 B]
        const int row = 32 * (warp & 3) + lane;
        int t = 0;
        for (int w = blockIdx.x; w < n_work; w += wstep) {
            int cloud, split, tile0, nb;
            const int ntiles = work_tiles(w, cloud, split, tile0, nb);
            for (int it = 0; it < ntiles; ++it, ++t) {
                const int stage = t % POOL_STAGES;
                const int n = (tile0 + it) * 128 + row;
                const bool valid = n < nb;
                const uint4* src = reinterpret_cast<const uint4*>(P.Y16 + ((size_t)cloud * P.N + (valid ? n : 0)) * 64);
                uint4 yv[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) yv[c] = valid ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
                if (t >= POOL_STAGES) mbar_wait(&y_empty[stage], ((t / POOL_STAGES) - 1) & 1);
                uint8_t* dst = sY + stage * 16384;
#pragma unroll
                for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(dst + c * 2048 + row * 16) = yv[c];
                fence_async_smem();
                fence_before_sync();
                mbar_arrive(&y_full[stage]);
            }
        }
    } else {
        reg_inc<184>();
        // =================================================================== softmax warpgroups: g takes the tiles with t % 2 == g
        const int g = warp >> 2, quad = warp & 3;
        const int row = 32 * quad + lane;
        const uint32_t lane_base = 32 * quad;
        const uint32_t sbase = tmem_addr(tb, lane_base, PC_S + 128 * g);
        float m_run = -INFINITY, l_run = 0.f, alpha = 0.f;
        float acc[64];
        uint32_t ph_s = 0, ph_o = 0;
        int t = 0;
        for (int w = blockIdx.x; w < n_work; w += wstep) {
            int cloud, split, tile0, nb;
            const int ntiles = work_tiles(w, cloud, split, tile0, nb);
            m_run = -INFINITY; l_run = 0.f; alpha = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) acc[j] = 0.f;
            // Within a work item this warpgroup sees either the even or the odd tiles; which of the two depends on how many
            // tiles the CTA has processed before.  The partial is filed under the tile PARITY, so the grouping (and hence
            // every rounding) of a cloud is independent of the batch it is part of.
            const int group = (g ^ t) & 1;
            for (int it = 0; it < ntiles; ++it, ++t) {
                if ((t & 1) != g) continue;
                const int n_valid = min(128, nb - (tile0 + it) * 128);
                mbar_wait(&s_full[g], ph_s);
                ph_s ^= 1;
                fence_after_sync();
                // ---- pass 1: row max over the valid columns
                float mx = -INFINITY;
                {
                    uint32_t v[32];
#pragma unroll 1
                    for (int c0 = 0; c0 < n_valid; c0 += 32) {
                        tmem_ld32(sbase + c0, v);
                        tmem_ld_wait32(v);
                        if (c0 + 32 <= n_valid) mx = max_chunk32(v, mx);
                        else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c0 + j < n_valid) mx = fmaxf(mx, __uint_as_float(v[j]));
                        }
                    }
                }
                const float m_new = fmaxf(m_run, mx);
                alpha = ex2(m_run - m_new);
                const float2 neg2 = make_float2(-m_new, -m_new);
                float2 sum2 = make_float2(0.f, 0.f);
                // ---- pass 2: probabilities; P (bf16) overwrites score columns that were already consumed
#pragma unroll
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    uint32_t v[32], pk[16];
                    if (c0 < n_valid) {
                        tmem_ld32(sbase + c0, v);
                        tmem_ld_wait32(v);
                        if (c0 + 32 <= n_valid) exp_chunk32(v, neg2, sum2, pk);
                        else {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const float p0 = (c0 + j < n_valid) ? ex2(__uint_as_float(v[j]) - m_new) : 0.f;
                                const float p1 = (c0 + j + 1 < n_valid) ? ex2(__uint_as_float(v[j + 1]) - m_new) : 0.f;
                                sum2.x += p0 + p1;
                                pk[j >> 1] = pack_bf16(p0, p1);
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) pk[j] = 0u;
                    }
                    tmem_st16(sbase + (c0 >> 1), pk);
                }
                l_run = l_run * alpha + (sum2.x + sum2.y);
                m_run = m_new;
                tmem_st_wait();
                fence_before_sync();
                mbar_arrive(&p_ready[g]);        // every lane arrives (count 128)
                // ---- Z += P Y (rescaled running sum in registers)
                mbar_wait(&o_full[g], ph_o);
                ph_o ^= 1;
                fence_after_sync();
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t o[32];
                    tmem_ld32(tmem_addr(tb, lane_base, PC_O + 64 * g + c0), o);
                    tmem_ld_wait32(o);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c0 + j] = fmaf(acc[c0 + j], alpha, __uint_as_float(o[j]));
                }
                fence_before_sync();
            }
            if ((row & 15) == 0) {
                const int h = row >> 4;
                float* dst = P.part + ((((size_t)cloud * P.nsplit + split) * 2 + group) * TH + h) * 66;
                dst[0] = m_run;
                dst[1] = l_run;
#pragma unroll
                for (int j = 0; j < 64; ++j) dst[2 + j] = acc[j];
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tb, 512);
}

// ====================================================================================== pooled attention, transposed
// Same arithmetic as pma_pool_tc_kernel with the roles of the MMA dimensions swapped, so that no score is computed twice:
//   S (128 points, 16 columns = 8 heads + 8 zero) = Ytile (128 x 64, the staged tile as the A operand) . AqB^T   [4 MMAs, N = 16]
//   thread = point: 8 exponentials per point and tile (the row-copy formulation evaluates 128 per row);
//   the per-head tile maximum is a warp-shuffle + shared-memory reduction over the 128 points of the tile
//   P (K = points, N = 16 columns = 8 heads + 8 zero) is written to shared memory as a K-major B operand (8 two-byte stores
//   per thread, 4 KB per warpgroup)
//   Z^T (128 rows = 64 features | ones row | 63 don't-care rows, 16 columns) = [Ytile | 1]^T . P    [8 MMAs, N = 16, MN-major A]
//   -- the Y tile is the A operand of BOTH products (K-major for the scores, MN-major here), so the tensor work per tile is
//   128 x 16 x 128 MACs instead of the 128 x 64 x 128 of the P^T-as-A form (whose 120 zero rows were multiplied as well;
//   the timeline showed the issuing thread blocked behind them: 666 cycles per tile in the P V issue alone).
//   Thread f < 64 of a warpgroup holds feature f of the eight heads' running sums; the row sums come out of the ones row.
constexpr int P2_NG = 4;                       // softmax warpgroups = tiles in the softmax stage (the chain S -> softmax -> P V of a
                                               // tile is ~2000 cycles long: with two in flight the kernel ran at 1700 cycles per tile)
constexpr int P2_STAGES = 10;                  // Y tile ring (asynchronous copies keep six tiles in flight)
constexpr int P2_WP = 4 * P2_NG;               // first producer warp
constexpr int P2_WM = P2_WP + 4;               // MMA warp
constexpr int P2_THREADS = (P2_WM + 4) * 32;   // 16 warps
constexpr int P2_YSTAGE = 9 * 2048;            // a Y stage: 8 chunks of 8 features + one constant chunk (1, 0, ..., 0): the "ones" column
struct Pool2Smem {
    static constexpr int AQB = 0;                                 // 16 x 64 K-major B image (2 KB)
    static constexpr int Y = 2048;                                // P2_STAGES x P2_YSTAGE
    static constexpr int PT = Y + P2_STAGES * P2_YSTAGE + 16384;  // (16 KB slack: the A operand of the last stage reads 16 chunks)
                                                                  // P2_NG x 4096: P as a 16-row K-major B image
    static constexpr int RED = PT + P2_NG * 4096;                 // [NG][2 parities][4 warps][8 heads] maxima, then [NG][4][8] sums
    static constexpr int BARS = RED + (P2_NG * 2 * 4 * 8 + P2_NG * 4 * 8) * 4;
    static constexpr int TOTAL = BARS + 48 * 8 + 16;
};
constexpr uint32_t P2_S = 0, P2_O = 32 * P2_NG, P2_OW = 32;         // NG x 32 score columns (16 used) | NG x 32 output columns (16 used)

// EXACT = false (streaming, the first pass): like mab_reduce6_tc_kernel, the heads' sums ACCUMULATE IN TMEM across the tiles of
// a work item against a FIXED reference exponent -- here simply 0: softmax is shift invariant, and with fp32 sums, bf16
// probabilities and fp32 accumulators any reference within 2^+-60 of the true maximum is exact to rounding (log2-domain
// scores of a trained model are O(10)).  The row sums come out of the tensor core as well: every Y stage carries a ninth,
// constant chunk whose first column is 1, so column 64 of Z = P^T [Y | 1] is sum_n p(n).  Per tile a warpgroup then only
// loads 8 scores, evaluates 8 exponentials and writes its P^T column: no maximum, no shuffles, no named barrier, no read-back
// or rescale of Z -- the kernel becomes the stream over Y it should be.  A work item whose sum leaves [2^-60, 2^60] (a score
// far from the reference, overflow, NaN) is flagged in P.redo and redone by the EXACT = true pass, which is the per-tile
// re-referencing algorithm (running maximum, Z read back and rescaled every tile) restricted to the flagged items.
template <bool EXACT>
__global__ void __launch_bounds__(P2_THREADS, 1) pma_pool2_tc_kernel(const __grid_constant__ PoolParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sAqB = smem + Pool2Smem::AQB;
    uint8_t* sY = smem + Pool2Smem::Y;
    uint8_t* sPT = smem + Pool2Smem::PT;
    float* sRedMax = reinterpret_cast<float*>(smem + Pool2Smem::RED);
    float* sRedSum = sRedMax + P2_NG * 2 * 4 * 8;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Pool2Smem::BARS);
    uint64_t* y_full = bars;                        // [P2_STAGES] count 1 + transaction bytes (TMA loads)
    uint64_t* y_empty = bars + P2_STAGES;           // [P2_STAGES] count 1 (commit)
    uint64_t* s_full = bars + 2 * P2_STAGES;        // [NG] count 1
    uint64_t* p_ready = s_full + P2_NG;             // [NG] count 4
    uint64_t* o_full = p_ready + P2_NG;             // [NG] count 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * P2_STAGES + 3 * P2_NG);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_work = P.n_work, wstep = gridDim.x;
    auto work_tiles = [&](int w, int& cloud, int& split, int& tile0, int& nb) {
        cloud = w / P.nsplit;
        split = w - cloud * P.nsplit;
        tile0 = split * P.tiles_per_split;
        nb = main_points(P.counts ? max(1, min(P.N, __ldg(P.counts + cloud))) : P.N, P.tail_max);
        return max(0, min((nb + 127) >> 7, tile0 + P.tiles_per_split) - tile0);
    };
    // EXACT: only the work items flagged by the streaming pass are processed (normally none: leave at once)
    auto skipped = [&](int w) { return EXACT && __ldg(P.redo + w) == 0; };
    if (EXACT) {
        int any = 0;
        for (int w = blockIdx.x + (int)threadIdx.x * wstep; w < n_work; w += (int)blockDim.x * wstep) any |= (__ldg(P.redo + w) != 0);
        if (!__syncthreads_or(any)) return;
    }
    // B image of the pooled queries: row n < 8 = head n (row 16 n of the 128-row A image built by prep_kernel), rows 8..15 zero
    for (int i = threadIdx.x; i < 8 * 16; i += blockDim.x) {
        const int c = i >> 4, n = i & 15;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (n < 8) v = __ldg(reinterpret_cast<const uint4*>(P.Aq + c * 2048 + (16 * n) * 16));
        *reinterpret_cast<uint4*>(sAqB + c * 256 + n * 16) = v;
    }
    for (int i = threadIdx.x * 16; i < P2_NG * 4096; i += blockDim.x * 16) *reinterpret_cast<uint4*>(sPT + i) = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < P2_STAGES * 128; i += blockDim.x)          // the ones chunk of every stage: bf16 (1, 0, 0, 0, 0, 0, 0, 0) per row
        *reinterpret_cast<uint4*>(sY + (i >> 7) * P2_YSTAGE + 8 * 2048 + (i & 127) * 16) = make_uint4(0x00003F80u, 0, 0, 0);
    if (warp == P2_WM) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 0) {
        for (int i = 0; i < P2_STAGES; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 1); }
        for (int i = 0; i < P2_NG; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&o_full[i], 1); }
        fence_barrier_init();
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;
#ifdef PCA_TIMELINE
    // role 0: MMA warp, 1: softmax warpgroup 0 / warp 0, 2: producer warp 0 (2000 stamps each)
    const int tl_role = warp == P2_WM ? 0 : (warp == 0 ? 1 : (warp == P2_WP ? 2 : (warp == P2_WM + 1 ? 3 : -1)));
    long long* tl = (!EXACT && P.timeline != nullptr && blockIdx.x == 0 && tl_role >= 0 && lane == 0) ? P.timeline + 4000 * tl_role : nullptr;
    int tl_n = 0;
    auto stamp = [&](int tag) {
        if (tl != nullptr && tl_n < 2000) { tl[2 * tl_n] = tag; tl[2 * tl_n + 1] = clock64(); ++tl_n; }
    };
#else
    auto stamp = [&](int) {};
#endif

    if (warp >= P2_WM) {
        reg_dec<40>();
        // Two issuing warps (the timeline of the one-warp version: ~75 cycles per tcgen05.mma issue + commits = 1000 of the 2200
        // cycles of a tile in the issuing thread alone): warp WM issues the score products, warp WM + 1 the P V products.
        const uint32_t aqb = smem_u32(sAqB), yb = smem_u32(sY), ptb = smem_u32(sPT);
        int total = 0;
        if (warp <= P2_WM + 1)
            for (int w = blockIdx.x; w < n_work; w += wstep) { int a, b, c, e; if (!skipped(w)) total += work_tiles(w, a, b, c, e); }
        if (warp == P2_WM) {
            // =================================================================== S issuer: at most NG tiles ahead of the softmax
            const bool leader = elect_one();
            const uint32_t idesc_s = idesc_bf16(128, 16, 0, 0);
            for (int t = 0; t < total; ++t) {
                const int stage = t % P2_STAGES, b = t % P2_NG;
                // the score buffer b was last read by its warpgroup before it arrived on p_ready for tile t - NG
                if (t >= P2_NG) mbar_spin(&p_ready[b], ((t / P2_NG) - 1) & 1);
                stamp(1);
                mbar_spin(&y_full[stage], (t / P2_STAGES) & 1);
                stamp(2);
                fence_after_sync();
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        // (even / odd K steps into two accumulators: accumulating MMAs on ONE accumulator serialise at ~80 cycles each)
                        mma_ss(tmem_addr(tb, 0, P2_S + 32 * b + 16 * (ks & 1)), smem_desc(yb + stage * P2_YSTAGE + ks * 4096, 2048, 128),
                               smem_desc(aqb + ks * 512, 256, 128), idesc_s, ks > 1);
                    mma_commit(&s_full[b]);
                }
                __syncwarp();
                stamp(3);
            }
        } else if (warp == P2_WM + 1) {
            // =================================================================== P V issuer
            const bool leader = elect_one();
            // A = [Y | 1]^T, MN-major: rows 0..63 features, row 64 the ones column, rows 65..127 whatever follows the stage in
            // shared memory (they only reach output rows that are never read)
            const uint32_t idesc_pv = idesc_bf16(128, 16, 1, 0);
            // position of the P V stream inside its work item (streaming pass: the first tile of a warpgroup in an item starts
            // the TMEM accumulator, the others add to it)
            int pw = blockIdx.x - wstep, pit = 0, pnt = 0;
            for (int t = 0; t < total; ++t) {
                const int b = t % P2_NG, stage = t % P2_STAGES;
                while (pit >= pnt) {
                    pw += wstep;
                    pit = 0;
                    int a, bb, c, e;
                    pnt = skipped(pw) ? 0 : work_tiles(pw, a, bb, c, e);
                }
                const uint32_t acc_in = (!EXACT && pit >= P2_NG) ? 1u : 0u;
                ++pit;
                mbar_spin(&p_ready[b], (t / P2_NG) & 1);
                stamp(4);
                fence_after_sync();
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        mma_ss(tmem_addr(tb, 0, P2_O + P2_OW * b + 16 * (ks & 1)), smem_desc(yb + stage * P2_YSTAGE + ks * 256, 128, 2048),
                               smem_desc(ptb + b * 4096 + ks * 512, 256, 128), idesc_pv, (ks > 1) ? 1u : acc_in);
                    mma_commit(&o_full[b]);
                    mma_commit(&y_empty[stage]);
                }
                __syncwarp();
                stamp(5);
            }
        }
    } else if (warp >= P2_WP) {
        reg_dec<56>();
        // =================================================================== loader: Y tiles by TMA, 8 boxes of 128 rows x 8
        // columns land as [chunk][row][16 B].  (The four-warp cp.async producer it replaces spent ~1000 cycles per tile in its
        // own bookkeeping, fences and arrivals: the timeline showed it as the critical path.)  Rows past the cloud's points in
        // a ragged tile hold whatever follows in memory; the softmax warpgroup zero-fills them before the P V product.
        if (warp == P2_WP) {
            if (lane == 0) tma_prefetch_desc(&P.tmapY);
            int gt = 0;
            for (int w = blockIdx.x; w < n_work; w += wstep) {
                if (skipped(w)) continue;
                int cloud, split, tile0, nb;
                const int ntiles = work_tiles(w, cloud, split, tile0, nb);
                for (int it = 0; it < ntiles; ++it, ++gt) {
                    const int stage = gt % P2_STAGES;
                    stamp(20);
                    if (gt >= P2_STAGES) mbar_wait(&y_empty[stage], ((gt / P2_STAGES) - 1) & 1);
                    stamp(21);
                    if (lane == 0) {
                        uint8_t* dst = sY + stage * P2_YSTAGE;
                        const long long r0 = (long long)cloud * P.N + (long long)(tile0 + it) * 128;
                        mbar_arrive_expect_tx(&y_full[stage], 16384);
                        tma_load_3d(dst, &P.tmapY, 0, (int)r0, 0, &y_full[stage]);      // one box: (8 elements, 128 rows, 8 chunks)
                    }
                    __syncwarp();
                    stamp(22);
                }
            }
        }
    } else {
        reg_inc<88>();        // per scheduler: 6 warps x 80 at launch >= softmax 4 x 88 + producer 56 + MMA 40
        // =================================================================== softmax warpgroups: g takes the tiles with t % NG == g
        const int g = warp >> 2, quad = warp & 3;
        const int row = 32 * quad + lane;                 // point of the tile
        const uint32_t lane_base = 32 * quad;
        uint8_t* pt = sPT + g * 4096;
        int t = 0, own = 0;                               // own: tiles this warpgroup has processed (phases, reduction buffer parity)
        for (int w = blockIdx.x; w < n_work; w += wstep) {
            if (skipped(w)) continue;
            int cloud, split, tile0, nb;
            const int ntiles = work_tiles(w, cloud, split, tile0, nb);
            if (!EXACT) {
                // ---------------- streaming pass: reference exponent 0, Z and the row sums accumulate in TMEM
                const int group = (g - t) & (P2_NG - 1);
                int mine = 0;                              // tiles of this warpgroup in this work item
                for (int it = 0; it < ntiles; ++it, ++t) {
                    if ((t & (P2_NG - 1)) != g) continue;
                    const bool valid = row < min(128, nb - (tile0 + it) * 128);
                    stamp(10);
                    mbar_spin(&s_full[g], own & 1);
                    stamp(11);
                    fence_after_sync();
                    uint32_t sv[8], sw[8];
                    tmem_ld8(tmem_addr(tb, lane_base, P2_S + 32 * g), sv);
                    tmem_ld8(tmem_addr(tb, lane_base, P2_S + 32 * g + 16), sw);
                    tmem_ld_wait();
#pragma unroll
                    for (int h = 0; h < 8; ++h) sv[h] = __float_as_uint(__uint_as_float(sv[h]) + __uint_as_float(sw[h]));
                    stamp(12);
                    if (own > 0) mbar_spin(&o_full[g], (own - 1) & 1);      // P V of the previous tile has read this P^T buffer
                    stamp(13);
                    if (!valid) {          // ragged tile: this row of the stage is not a point of the cloud (0 x NaN would poison Z)
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            *reinterpret_cast<uint4*>(sY + (t % P2_STAGES) * P2_YSTAGE + c * 2048 + row * 16) = make_uint4(0, 0, 0, 0);
                    }
#pragma unroll
                    for (int h = 0; h < 8; ++h)
                        *reinterpret_cast<__nv_bfloat16*>(pt + (row >> 3) * 256 + h * 16 + (row & 7) * 2) =
                            __float2bfloat16(valid ? ex2(__uint_as_float(sv[h])) : 0.f);
                    fence_async_smem();
                    fence_before_sync();
                    warp_arrive(&p_ready[g]);
                    stamp(15);
                    ++own;
                    ++mine;
                }
                stamp(16);
                {
                    // the item's Z^T and row sums: one read of the TMEM accumulator once its last P V has completed.  Thread f < 64
                    // holds feature f of the eight heads, thread 64 the eight row sums.  (The first P V of the next item needs
                    // every warp's arrival, so the accumulator is not overwritten before it has been read.)
                    float* dst = P.part + (((size_t)cloud * P.nsplit + split) * P2_NG + group) * TH * 66;
                    uint32_t o[8];
#pragma unroll
                    for (int h = 0; h < 8; ++h) o[h] = 0u;
                    if (mine > 0 && quad < 3) {
                        mbar_spin(&o_full[g], (own - 1) & 1);
                        fence_after_sync();
                        uint32_t o1[8];
                        tmem_ld8(tmem_addr(tb, lane_base, P2_O + P2_OW * g), o);
                        tmem_ld8(tmem_addr(tb, lane_base, P2_O + P2_OW * g + 16), o1);
                        tmem_ld_wait();
#pragma unroll
                        for (int h = 0; h < 8; ++h) o[h] = __float_as_uint(__uint_as_float(o[h]) + __uint_as_float(o1[h]));
                        fence_before_sync();
                    }
                    if (row < 64) {
#pragma unroll
                        for (int h = 0; h < 8; ++h) dst[h * 66 + 2 + row] = __uint_as_float(o[h]);
                    } else if (row == 64) {
                        bool bad = false;
#pragma unroll
                        for (int h = 0; h < 8; ++h) {
                            const float l_sum = __uint_as_float(o[h]);
                            dst[h * 66] = mine > 0 ? 0.f : -INFINITY;
                            dst[h * 66 + 1] = l_sum;
                            // sum outside [2^-60, 2^60] (a score far from the reference, overflow, NaN): redo the item exactly
                            bad |= mine > 0 && !(l_sum > 8.673617379884035e-19f && l_sum < 1.152921504606847e18f);
                        }
                        if (bad) P.redo[w] = 1;
                    }
                }
                stamp(17);
                continue;
            }
            float m_run[8], l_part[8];
#pragma unroll
            for (int h = 0; h < 8; ++h) { m_run[h] = -INFINITY; l_part[h] = 0.f; }
            float accz[8];                                 // feature `row` (< 64) of the eight heads' running sums
#pragma unroll
            for (int h = 0; h < 8; ++h) accz[h] = 0.f;
            // the partial is filed under (tile index within the work item) mod NG, so the grouping (and hence every rounding)
            // of a cloud is independent of what the CTA processed before
            const int group = (g - t) & (P2_NG - 1);
            for (int it = 0; it < ntiles; ++it, ++t) {
                if ((t & (P2_NG - 1)) != g) continue;
                const int n_valid = min(128, nb - (tile0 + it) * 128);
                const bool valid = row < n_valid;
                mbar_spin(&s_full[g], own & 1);
                fence_after_sync();
                uint32_t sv[8], sw[8];
                tmem_ld8(tmem_addr(tb, lane_base, P2_S + 32 * g), sv);
                tmem_ld8(tmem_addr(tb, lane_base, P2_S + 32 * g + 16), sw);
                tmem_ld_wait();
                float s[8], mx[8];
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    s[h] = __uint_as_float(sv[h]) + __uint_as_float(sw[h]);
                    mx[h] = valid ? s[h] : -INFINITY;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], o));
                }
                float* red = sRedMax + ((g * 2 + (own & 1)) * 4) * 8;
                if (lane < 8) {
                    float v = mx[0];
#pragma unroll
                    for (int h = 1; h < 8; ++h) v = lane == h ? mx[h] : v;
                    red[quad * 8 + lane] = v;
                }
                named_bar_sync(1 + g, 128);
                if (!valid) {              // ragged tile: see the streaming pass
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<uint4*>(sY + (t % P2_STAGES) * P2_YSTAGE + c * 2048 + row * 16) = make_uint4(0, 0, 0, 0);
                }
                float alpha_h[8];
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    const float m_tile = fmaxf(fmaxf(red[h], red[8 + h]), fmaxf(red[16 + h], red[24 + h]));
                    const float m_new = fmaxf(m_run[h], m_tile);
                    const float alpha = (m_new == -INFINITY) ? 1.f : ex2(m_run[h] - m_new);
                    const float p = valid ? ex2(s[h] - m_new) : 0.f;
                    l_part[h] = fmaf(l_part[h], alpha, p);
                    m_run[h] = m_new;
                    alpha_h[h] = alpha;
                    *reinterpret_cast<__nv_bfloat16*>(pt + (row >> 3) * 256 + h * 16 + (row & 7) * 2) = __float2bfloat16(p);
                }
                fence_async_smem();
                fence_before_sync();
                warp_arrive(&p_ready[g]);
                // ---- Z^T += [Y | 1]^T P: thread f < 64 rescales and adds feature f of the eight heads
                mbar_spin(&o_full[g], own & 1);
                fence_after_sync();
                if (quad < 2) {
                    uint32_t o[8], o1[8];
                    tmem_ld8(tmem_addr(tb, lane_base, P2_O + P2_OW * g), o);
                    tmem_ld8(tmem_addr(tb, lane_base, P2_O + P2_OW * g + 16), o1);
                    tmem_ld_wait();
#pragma unroll
                    for (int h = 0; h < 8; ++h) accz[h] = fmaf(accz[h], alpha_h[h], __uint_as_float(o[h]) + __uint_as_float(o1[h]));
                }
                fence_before_sync();
                ++own;
            }
            // ---- row sums: reduce the per-point partials over the 128 threads
            float l_tot[8];
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                float v = l_part[h];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                l_tot[h] = v;
            }
            float* rs = sRedSum + (g * 4) * 8;
            named_bar_sync(1 + g, 128);                    // the previous work item's sums have been read
            if (lane < 8) {
                float v = l_tot[0];
#pragma unroll
                for (int h = 1; h < 8; ++h) v = lane == h ? l_tot[h] : v;
                rs[quad * 8 + lane] = v;
            }
            named_bar_sync(1 + g, 128);
            float* dst = P.part + (((size_t)cloud * P.nsplit + split) * P2_NG + group) * TH * 66;
            if (row < 64) {
#pragma unroll
                for (int h = 0; h < 8; ++h) dst[h * 66 + 2 + row] = accz[h];
            } else if (row < 72) {
                const int h = row - 64;
                float m_l = m_run[0];
#pragma unroll
                for (int hh = 1; hh < 8; ++hh) m_l = h == hh ? m_run[hh] : m_l;
                dst[h * 66] = m_l;
                dst[h * 66 + 1] = rs[h] + rs[8 + h] + rs[16 + h] + rs[24 + h];
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == P2_WM) tmem_dealloc(tb, 512);
}


// merge the pooled partials, apply fc_v to the 8 pooled vectors, then the MAB tail and the final Linear.
// 4 clouds per 256-thread block (64 threads = 64 features per cloud); weights are read through k-major copies so that
// the 64 threads of a cloud read consecutive addresses.
struct PoolFinParams {
    const float* part; int nslots; int B;    // (B, nslots, 8, 66)
    const float* QpS;                         // (64) fc_q(S)
    const float* WvT; const float* bv;        // pma.mab.fc_v transposed (k, f), bias (64)
    const float* WoT; const float* bo;        // pma.mab.fc_o transposed (k, f), bias
    const float* Wl; const float* bl; int C;  // final Linear (C, 64)
    float* logits;                            // (B, C)
    float* pooled_debug;                      // nullable (B, 64)
    // leftover points (tail rule): merged here as one more softmax slot each
    const __nv_bfloat16* Y16;                 // (B, N, 64) the pooled points
    int N; const int* counts; int tail_max;
    const float* Wqk;                         // (8, 64) scale * Wk_h^T fc_q(S)_h
};
__global__ void __launch_bounds__(256) finalize_pool_kernel(const PoolFinParams P) {
    __shared__ __align__(16) float sZ[4][TH][TD + 4], sO[4][64], sO1[4][64], sY[4][TC_TAIL_MAX][64], sS[4][TC_TAIL_MAX][TH];
    const int sub = threadIdx.x >> 6, f = threadIdx.x & 63;
    const int cloud = blockIdx.x * 4 + sub;
    const bool valid = cloud < P.B;
    // ---- leftover points: rows of Y and their 8 seed scores (thread f: head f / 8, 8-feature slice f % 8)
    int r_tail = 0;
    if (valid && P.tail_max > 0) {
        const int nbt = P.counts ? max(1, min(P.N, __ldg(P.counts + cloud))) : P.N;
        const int n0 = main_points(nbt, P.tail_max);
        r_tail = nbt - n0;
        for (int j = 0; j < r_tail; ++j) sY[sub][j][f] = __bfloat162float(P.Y16[((size_t)cloud * P.N + n0 + j) * 64 + f]);
    }
    __syncthreads();
    for (int j = 0; j < r_tail; ++j) {
        const int h = f >> 3, part = f & 7;
        float d = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) d = fmaf(__ldg(P.Wqk + h * TD + 8 * part + k), sY[sub][j][8 * part + k], d);
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (part == 0) sS[sub][j][h] = d;
    }
    __syncthreads();
    if (valid) {
        // thread f merges feature f of every head
        if (P.nslots == 2) {
            // usual case: the 48 values of all heads are fetched in one round (the kernel is latency-bound); same
            // arithmetic, in the same order, as the general loop below
            float pm[2][TH], pl[2][TH], pz[2][TH];
            const float* pc = P.part + ((size_t)cloud * 2) * TH * 66;
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int h = 0; h < TH; ++h) {
                    const float* pp = pc + ((size_t)s * TH + h) * 66;
                    pm[s][h] = __ldg(pp); pl[s][h] = __ldg(pp + 1); pz[s][h] = __ldg(pp + 2 + f);
                }
#pragma unroll
            for (int h = 0; h < TH; ++h) {
                float mmax = fmaxf(fmaxf(-INFINITY, pm[0][h]), pm[1][h]);
                for (int j = 0; j < r_tail; ++j) mmax = fmaxf(mmax, sS[sub][j][h]);
                float l = 0.f, z = 0.f;
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const float wgt = (pm[s][h] == -INFINITY) ? 0.f : exp2f(pm[s][h] - mmax);
                    l = fmaf(pl[s][h], wgt, l);
                    z = fmaf(pz[s][h], wgt, z);
                }
                for (int j = 0; j < r_tail; ++j) {
                    const float wgt = exp2f(sS[sub][j][h] - mmax);
                    l += wgt;
                    z = fmaf(sY[sub][j][f], wgt, z);
                }
                sZ[sub][h][f] = z / l;
            }
        } else
#pragma unroll
        for (int h = 0; h < TH; ++h) {
            const float* p0 = P.part + (((size_t)cloud * P.nslots) * TH + h) * 66;
            float mmax = -INFINITY;
            for (int s = 0; s < P.nslots; ++s) mmax = fmaxf(mmax, __ldg(p0 + (size_t)s * TH * 66));
            for (int j = 0; j < r_tail; ++j) mmax = fmaxf(mmax, sS[sub][j][h]);
            float l = 0.f, z = 0.f;
            for (int s = 0; s < P.nslots; ++s) {
                const float* pp = p0 + (size_t)s * TH * 66;
                const float m = __ldg(pp);
                const float wgt = (m == -INFINITY) ? 0.f : exp2f(m - mmax);      // a warpgroup that saw no tile
                l = fmaf(__ldg(pp + 1), wgt, l);
                z = fmaf(__ldg(pp + 2 + f), wgt, z);
            }
            for (int j = 0; j < r_tail; ++j) {
                const float wgt = exp2f(sS[sub][j][h] - mmax);
                l += wgt;
                z = fmaf(sY[sub][j][f], wgt, z);
            }
            sZ[sub][h][f] = z / l;
        }
    }
    __syncthreads();
    if (valid) {
        const int h = f >> 3;
        float a = __ldg(P.bv + f);
#pragma unroll 4
        for (int k = 0; k < 64; k += 4) {           // staged rows as float4: same products, same order, 4x fewer LDS
            const float4 z4 = *reinterpret_cast<const float4*>(&sZ[sub][h][k]);
            a = fmaf(z4.x, __ldg(P.WvT + k * 64 + f), a);
            a = fmaf(z4.y, __ldg(P.WvT + (k + 1) * 64 + f), a);
            a = fmaf(z4.z, __ldg(P.WvT + (k + 2) * 64 + f), a);
            a = fmaf(z4.w, __ldg(P.WvT + (k + 3) * 64 + f), a);
        }
        sO[sub][f] = __ldg(P.QpS + f) + a;
    }
    __syncthreads();
    if (valid) {
        float acc = __ldg(P.bo + f);
#pragma unroll 4
        for (int k = 0; k < 64; k += 4) {
            const float4 o4 = *reinterpret_cast<const float4*>(&sO[sub][k]);
            acc = fmaf(o4.x, __ldg(P.WoT + k * 64 + f), acc);
            acc = fmaf(o4.y, __ldg(P.WoT + (k + 1) * 64 + f), acc);
            acc = fmaf(o4.z, __ldg(P.WoT + (k + 2) * 64 + f), acc);
            acc = fmaf(o4.w, __ldg(P.WoT + (k + 3) * 64 + f), acc);
        }
        const float o1 = sO[sub][f] + fmaxf(acc, 0.f);
        sO1[sub][f] = o1;
        if (P.pooled_debug) P.pooled_debug[(size_t)cloud * 64 + f] = o1;
    }
    __syncthreads();
    if (valid) {
        for (int c = f; c < P.C; c += 64) {
            float z = __ldg(P.bl + c);
            for (int k = 0; k < 64; ++k) z = fmaf(sO1[sub][k], __ldg(P.Wl + c * 64 + k), z);
            P.logits[(size_t)cloud * P.C + c] = z;
        }
    }
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __bfloat162float(in[i]);
}

// ------------------------------------------------------------------------------------ host orchestration
static int g_num_sms = 148;
static int g_pool_variant = 2;             // pooled attention: 2 = transposed, streaming + exact redo (default), 3 = transposed, exact pass only,
                                           // 1 = row copies (PCA_TC_POOL / pca_debug_set_pool_variant)
static int g_reduce_wg = 6;               // reduce kernel: 6 = sixth generation (+ exact redo of flagged items), 4 = streaming fifth generation (+ redo), 2 = exact only
static int g_apply_variant = 4;           // apply kernel: 4 = three softmax warpgroups, TMA loader; 3 = previous generation (PCA_TC_APPLY=3)
static int g_tail_max = TC_TAIL_MAX;       // tail rule (PCA_TC_TAIL=0 disables it: every point goes through the tensor-core kernels)
static long long* g_timeline = nullptr;      // set through pca_debug_set_timeline
void set_timeline(long long* p) { g_timeline = p; }
void set_tail_max(int t) { g_tail_max = t < 0 ? 0 : (t > TC_TAIL_MAX ? TC_TAIL_MAX : t); }
void set_reduce_wg(int n) { g_reduce_wg = (n == 4 || n == 6) ? n : 2; }
void set_pool_variant(int v) { g_pool_variant = (v == 1 || v == 3) ? v : 2; }
void set_apply_variant(int v) { g_apply_variant = (v == 3) ? 3 : 4; }
struct TcSplit { int tiles_total, tiles_per_split, nsplit; };
// The point range of a cloud is cut into fixed spans of 16 tiles (2048 points).  The cut depends on N only, never
// on the batch size, so a cloud's logits are bit-identical however the batch is sharded across calls / GPUs.
static TcSplit plan_split(int B, int N) {
    (void)B;
    TcSplit s;
    s.tiles_total = (N + 127) / 128;
    s.tiles_per_split = 16;
    s.nsplit = (s.tiles_total + s.tiles_per_split - 1) / s.tiles_per_split;
    return s;
}

struct TcLayout { size_t consts, part, kvblk, redo, y1, y2, total; };
static TcLayout tc_layout(int B, int N) {
    const TcSplit s = plan_split(B, N);
    Arena a(nullptr, 0);
    TcLayout L;
    L.consts = a.off; a.take<uint8_t>(sizeof(TcConsts));
    L.part = a.off; a.take<float>((size_t)B * 2 * s.nsplit * TH * TM * 10);
    L.kvblk = a.off; a.take<uint8_t>((size_t)B * 32768);
    L.redo = a.off; a.take<int>((size_t)B * s.nsplit);
    L.y1 = a.off; a.take<__nv_bfloat16>((size_t)B * N * 64);
    L.y2 = a.off; a.take<__nv_bfloat16>((size_t)B * N * 64);
    L.total = a.off;
    return L;
}

size_t st_tc_workspace_bytes(const pca_st_dims* d, int B, int N) {
    (void)d;
    return tc_layout(B, N).total;
}

int st_tc_supported(const pca_st_dims* d, int N) {
    return d->D == TD && d->H == TH && d->M == TM && d->S == 1 && d->ln == 0 && d->d_in >= 1 && d->d_in <= 4 && N >= 1;
}

struct TcDebug { float *H1, *Y1, *H2, *Y2, *pooled; };

static int st_tc_chunk(const float* X, const PointSrc& src, const int* counts, int B, int N, const pca_st_dims* d, const float* params,
                       float* logits, uint8_t* ws, const TcConsts* c, const TcDebug* dbg, cudaStream_t st) {
    const TcLayout L = tc_layout(B, N);
    const TcSplit sp = plan_split(B, N);
    float* part = reinterpret_cast<float*>(ws + L.part);
    uint8_t* kvblk = ws + L.kvblk;
    int* redo = reinterpret_cast<int*>(ws + L.redo);
    __nv_bfloat16* Y1 = reinterpret_cast<__nv_bfloat16*>(ws + L.y1);
    __nv_bfloat16* Y2 = reinterpret_cast<__nv_bfloat16*>(ws + L.y2);
    const int d_in = d->d_in;
    const float* p_isab0 = params;
    const long long n_isab0 = (long long)TM * TD + mab_count(TD, d_in, TD, 0) + mab_count(d_in, TD, TD, 0);
    const float* p_isab1 = p_isab0 + n_isab0;
    const float* p_pma = p_isab1 + (long long)TM * TD + 2 * mab_count(TD, TD, TD, 0);
    const float* p_lin = p_pma + TD + mab_count(TD, TD, TD, 0);
    const MabParams m00 = mab_slice(p_isab0 + TM * TD, TD, d_in, TD, 0);
    const MabParams m01 = mab_slice(p_isab0 + TM * TD + mab_count(TD, d_in, TD, 0), d_in, TD, TD, 0);
    const MabParams m10 = mab_slice(p_isab1 + TM * TD, TD, TD, TD, 0);
    const MabParams m11 = mab_slice(p_isab1 + TM * TD + mab_count(TD, TD, TD, 0), TD, TD, TD, 0);
    const MabParams mp = mab_slice(p_pma + TD, TD, TD, TD, 0);

    const int n_work = sp.nsplit * B;                          // persistent kernels: one CTA per SM walks the work items
    const int pgrid = n_work < g_num_sms ? n_work : g_num_sms;
    const int npairs = (B + 1) / 2;
    const int fgrid = npairs < 2 * g_num_sms ? npairs : 2 * g_num_sms;
    const double pts = (double)B * N;
    const bool tl_apply = getenv("PCA_TL_APPLY") != nullptr;   // which kernel a PCA_TIMELINE build records
    const int tm = g_tail_max;
    const dim3 tgrid(B, tm > 0 ? tm : 1);
    // the tail kernel only has work when some cloud can leave 1..tm points past a multiple of 128
    const bool has_tail = tm > 0 && (counts != nullptr ? N >= 129 : main_points(N, tm) != N);

    // ---- ISAB 0
    {
        RParams r{X, nullptr, N, d_in, sp.tiles_total, sp.tiles_per_split, sp.nsplit, n_work, counts, tm, c->Aq0, m00.Wkv, m00.bkv,
                  nullptr, (tl_apply || getenv("PCA_TL_REDUCE64") || getenv("PCA_TL_POOL")) ? nullptr : g_timeline, redo, 0, part};
        LaunchTimer lt("mab_reduce_tc_kernel", st, pts * 2.0 * (2.0 * d_in * TD + 2.0 * TM * TD), pts * 4.0 * d_in);
        if (g_reduce_wg == 6) {
            PCA_CHECK_CUDA(cudaMemsetAsync(redo, 0, (size_t)n_work * sizeof(int), st));
            R6Params r6{X, N, d_in, sp.tiles_total, sp.tiles_per_split, sp.nsplit, n_work, counts, tm, c->Gq0s, nullptr, m00.Wkv + (size_t)TD * d_in,
                        m00.bkv + TD, r.timeline, redo, part, CUtensorMap{}};
            r6.src = src;
            mab_reduce6_tc_kernel<false, false><<<pgrid, R6_THREADS, R6Smem::TOTAL, st>>>(r6);
            r6.timeline = nullptr;
            mab_reduce6_tc_kernel<false, true><<<pgrid, R6_THREADS, R6Smem::TOTAL, st>>>(r6);      // exact redo of flagged work items
            count_launch();
        } else if (g_reduce_wg == 4) {
            // streaming variant, then the exact variant on the (normally empty) list of flagged work items
            PCA_CHECK_CUDA(cudaMemsetAsync(redo, 0, (size_t)n_work * sizeof(int), st));
            mab_reduce5_tc_kernel<false, 4><<<pgrid, 28 * 32, R2Smem::TOTAL, st>>>(r);
            r.redo_only = 1;
            r.timeline = nullptr;
            mab_reduce5_tc_kernel<false, 2><<<pgrid, TC_THREADS20, R2Smem::TOTAL, st>>>(r);
            count_launch();
        } else {
            mab_reduce5_tc_kernel<false, 2><<<pgrid, TC_THREADS20, R2Smem::TOTAL, st>>>(r);
        }
    }
    PCA_CHECK_LAUNCH("mab_reduce_tc_kernel<small>");
    {
        F2Params f{part, (g_reduce_wg == 6 ? 1 : 2) * sp.nsplit, B, c->Qp0, c->WoS[0][0], m00.bo, c->WkvS[0][0], m01.bkv, kvblk, dbg ? dbg->H1 : nullptr,
                   X, nullptr, N, d_in, counts, tm, c->Wkv0T[0], m00.bkv};
        f.src = src;
        
        LaunchTimer lt("finalize_isab_kernel", st, (double)B * 2.0 * (TM * TD * TD + TM * TD * 2 * TD), (double)B * 32768.0);
        finalize_isab_tc_kernel<<<fgrid, 128, F2Smem::TOTAL, st>>>(f);
    }
    PCA_CHECK_LAUNCH("finalize_isab_kernel");
    {
        AParams a{X, nullptr, N, d_in, sp.tiles_total, sp.tiles_per_split, sp.nsplit, n_work, counts, tm, kvblk, m01.Wq, m01.bq,
                  g_apply_variant == 4 ? c->WqS0e : c->WqS0, g_apply_variant == 4 ? c->Wo0e : c->Wo0, m01.bo, Y1, tl_apply ? g_timeline : nullptr};
        a.src = src;
        if (g_apply_variant == 4) PCA_TRY(make_tmap_chunked_bf16(&a.tmapYo, Y1, (unsigned long long)B * N, 128, 128));
        LaunchTimer lt("mab_apply_tc_kernel", st, pts * 2.0 * (1.0 * d_in * TD + 2.0 * TM * TD + TD * TD), pts * (4.0 * d_in + 128.0));
        if (g_apply_variant == 4) mab_apply4_tc_kernel<false><<<pgrid, A4_THREADS, A4Smem::TOTAL, st>>>(a);
        else mab_apply3_tc_kernel<false><<<pgrid, TC_THREADS20, A3Smem::TOTAL, st>>>(a);
    }
    PCA_CHECK_LAUNCH("mab_apply_tc_kernel<small>");
    if (has_tail) {
        ATailParams t{X, nullptr, N, d_in, counts, tm, kvblk, c->Wq1T[0], m01.bq, c->Wo1T[0], m01.bo, Y1};
        t.src = src;
        LaunchTimer lt("mab_apply_tail_kernel", st, 0.0, 0.0);
        mab_apply_tail_kernel<<<tgrid, 64, 0, st>>>(t);
        PCA_CHECK_LAUNCH("mab_apply_tail_kernel");
    }
    // ---- ISAB 1
    {
        RParams r{nullptr, Y1, N, TD, sp.tiles_total, sp.tiles_per_split, sp.nsplit, n_work, counts, tm, c->Aq1, nullptr, m10.bkv,
                  c->Wkv1, getenv("PCA_TL_REDUCE64") ? g_timeline : nullptr, redo, 0, part};
        LaunchTimer lt("mab_reduce_tc_kernel", st, pts * 2.0 * (2.0 * TD * TD + 2.0 * TM * TD), pts * 128.0);
        if (g_reduce_wg == 6) {
            PCA_CHECK_CUDA(cudaMemsetAsync(redo, 0, (size_t)n_work * sizeof(int), st));
            R6Params r6{nullptr, N, TD, sp.tiles_total, sp.tiles_per_split, sp.nsplit, n_work, counts, tm, c->Gq1, c->Wv1e, nullptr, nullptr,
                        r.timeline, redo, part, CUtensorMap{}};
            PCA_TRY(make_tmap_chunked_bf16(&r6.tmapY, Y1, (unsigned long long)B * N, 128, 128));
            mab_reduce6_tc_kernel<true, false><<<pgrid, R6_THREADS, R6Smem::TOTAL, st>>>(r6);
            r6.timeline = nullptr;
            mab_reduce6_tc_kernel<true, true><<<pgrid, R6_THREADS, R6Smem::TOTAL, st>>>(r6);
            count_launch();
        } else if (g_reduce_wg == 4) {
            PCA_CHECK_CUDA(cudaMemsetAsync(redo, 0, (size_t)n_work * sizeof(int), st));
            mab_reduce5_tc_kernel<true, 4><<<pgrid, 28 * 32, R2Smem::TOTAL, st>>>(r);
            r.redo_only = 1;
            r.timeline = nullptr;
            mab_reduce5_tc_kernel<true, 2><<<pgrid, TC_THREADS20, R2Smem::TOTAL, st>>>(r);
            count_launch();
        } else {
            mab_reduce5_tc_kernel<true, 2><<<pgrid, TC_THREADS20, R2Smem::TOTAL, st>>>(r);
        }
    }
    PCA_CHECK_LAUNCH("mab_reduce_tc_kernel<64>");
    {
        F2Params f{part, (g_reduce_wg == 6 ? 1 : 2) * sp.nsplit, B, c->Qp1, c->WoS[1][0], m10.bo, c->WkvS[1][0], m11.bkv, kvblk, dbg ? dbg->H2 : nullptr,
                   nullptr, Y1, N, TD, counts, tm, c->Wkv0T[1], m10.bkv};
        LaunchTimer lt("finalize_isab_kernel", st, (double)B * 2.0 * (TM * TD * TD + TM * TD * 2 * TD), (double)B * 32768.0);
        finalize_isab_tc_kernel<<<fgrid, 128, F2Smem::TOTAL, st>>>(f);
    }
    PCA_CHECK_LAUNCH("finalize_isab_kernel");
    {
        AParams a{nullptr, Y1, N, TD, sp.tiles_total, sp.tiles_per_split, sp.nsplit, n_work, counts, tm, kvblk, nullptr, m11.bq,
                  g_apply_variant == 4 ? c->Wq1e : c->Wq1, g_apply_variant == 4 ? c->Wo1e : c->Wo1, m11.bo, Y2, getenv("PCA_TL_APPLY64") ? g_timeline : nullptr};
        if (g_apply_variant == 4) {
            PCA_TRY(make_tmap_chunked_bf16(&a.tmapY, Y1, (unsigned long long)B * N, 128, 128));
            PCA_TRY(make_tmap_chunked_bf16(&a.tmapYo, Y2, (unsigned long long)B * N, 128, 128));
        }
        LaunchTimer lt("mab_apply_tc_kernel", st, pts * 2.0 * (1.0 * TD * TD + 2.0 * TM * TD + TD * TD), pts * 256.0);
        if (g_apply_variant == 4) mab_apply4_tc_kernel<true><<<pgrid, A4_THREADS, A4Smem::TOTAL, st>>>(a);
        else mab_apply3_tc_kernel<true><<<pgrid, TC_THREADS20, A3Smem::TOTAL, st>>>(a);
    }
    PCA_CHECK_LAUNCH("mab_apply_tc_kernel<64>");
    if (has_tail) {
        ATailParams t{nullptr, Y1, N, TD, counts, tm, kvblk, c->Wq1T[1], m11.bq, c->Wo1T[1], m11.bo, Y2};
        LaunchTimer lt("mab_apply_tail_kernel", st, 0.0, 0.0);
        mab_apply_tail_kernel<<<tgrid, 64, 0, st>>>(t);
        PCA_CHECK_LAUNCH("mab_apply_tail_kernel");
    }
    // ---- PMA + Linear
    {
        PoolParams r{Y2, N, sp.tiles_total, sp.tiles_per_split, sp.nsplit, n_work, counts, tm, c->AqPool, part, redo,
                     getenv("PCA_TL_POOL") ? g_timeline : nullptr, CUtensorMap{}};
        if (g_pool_variant >= 2) PCA_TRY(make_tmap_chunked_bf16(&r.tmapY, Y2, (unsigned long long)B * N, 128, 128));
        if (g_pool_variant == 2) PCA_CHECK_CUDA(cudaMemsetAsync(redo, 0, (size_t)n_work * sizeof(int), st));
        if (g_pool_variant == 3) PCA_CHECK_CUDA(cudaMemsetAsync(redo, 0xff, (size_t)n_work * sizeof(int), st));   // debug: every item exact
        LaunchTimer lt("pma_pool_tc_kernel", st, pts * 2.0 * (2.0 * TH * TD), pts * 128.0);
        if (g_pool_variant == 2) {
            pma_pool2_tc_kernel<false><<<pgrid, P2_THREADS, Pool2Smem::TOTAL, st>>>(r);
            pma_pool2_tc_kernel<true><<<pgrid, P2_THREADS, Pool2Smem::TOTAL, st>>>(r);       // exact redo of flagged work items
        } else if (g_pool_variant == 3) {
            pma_pool2_tc_kernel<true><<<pgrid, P2_THREADS, Pool2Smem::TOTAL, st>>>(r);
        } else pma_pool_tc_kernel<<<pgrid, TC_THREADS16, PoolSmem::TOTAL, st>>>(r);
    }
    PCA_CHECK_LAUNCH("pma_pool_tc_kernel");
    {
        PoolFinParams p{part, (g_pool_variant >= 2 ? P2_NG : 2) * sp.nsplit, B, c->QpS, c->WvT_P, mp.bkv + TD, c->WoT_P, mp.bo, p_lin, p_lin + (long long)d->C * TD,
                        d->C, logits, dbg ? dbg->pooled : nullptr, Y2, N, counts, tm, c->WqkPool};
        LaunchTimer lt("finalize_pool_kernel", st, (double)B * 2.0 * (2.0 * TD * TD + TD * d->C), (double)B * 4.0 * d->C);
        finalize_pool_kernel<<<(B + 3) / 4, 256, 0, st>>>(p);
    }
    PCA_CHECK_LAUNCH("finalize_pool_kernel");
    if (dbg) {
        const long long n = (long long)B * N * 64;
        if (dbg->Y1) { bf16_to_f32_kernel<<<1024, 256, 0, st>>>(Y1, dbg->Y1, n); PCA_CHECK_LAUNCH("bf16_to_f32_kernel"); }
        if (dbg->Y2) { bf16_to_f32_kernel<<<1024, 256, 0, st>>>(Y2, dbg->Y2, n); PCA_CHECK_LAUNCH("bf16_to_f32_kernel"); }
    }
    return 0;
}

static int tc_configure() {
    // function attributes belong to the current device's context: configure once per device ordinal (a single process may
    // drive several GPUs from different threads, as nn.DataParallel does); idempotent, so races are benign
    static std::atomic<unsigned long long> done_mask{0};
    int dev = 0, major = 0;
    PCA_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && ((done_mask.load(std::memory_order_acquire) >> dev) & 1ull)) return 0;
    PCA_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    PCA_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    if (major != 10) return fail(PCA_EDEVICE, "tcgen05 path needs an sm_100 device (found compute capability %d.x)", major);
    if (const char* v = getenv("PCA_TC_TAIL")) { g_tail_max = atoi(v); if (g_tail_max < 0) g_tail_max = 0; if (g_tail_max > TC_TAIL_MAX) g_tail_max = TC_TAIL_MAX; }
    PCA_CHECK_CUDA((cudaFuncSetAttribute(mab_reduce5_tc_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, R2Smem::TOTAL)));
    PCA_CHECK_CUDA((cudaFuncSetAttribute(mab_reduce5_tc_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, R2Smem::TOTAL)));
    PCA_CHECK_CUDA((cudaFuncSetAttribute(mab_reduce5_tc_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, R2Smem::TOTAL)));
    PCA_CHECK_CUDA((cudaFuncSetAttribute(mab_reduce5_tc_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, R2Smem::TOTAL)));
    PCA_CHECK_CUDA((cudaFuncSetAttribute(mab_reduce6_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, R6Smem::TOTAL)));
    PCA_CHECK_CUDA((cudaFuncSetAttribute(mab_reduce6_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, R6Smem::TOTAL)));
    PCA_CHECK_CUDA((cudaFuncSetAttribute(mab_reduce6_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, R6Smem::TOTAL)));
    PCA_CHECK_CUDA((cudaFuncSetAttribute(mab_reduce6_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, R6Smem::TOTAL)));
    if (const char* v = getenv("PCA_TC_REDUCE_WG")) g_reduce_wg = (v[0] == '2') ? 2 : (v[0] == '4' ? 4 : 6);
    PCA_CHECK_CUDA(cudaFuncSetAttribute(mab_apply3_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, A3Smem::TOTAL));
    PCA_CHECK_CUDA(cudaFuncSetAttribute(mab_apply3_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, A3Smem::TOTAL));
    PCA_CHECK_CUDA(cudaFuncSetAttribute(mab_apply4_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, A4Smem::TOTAL));
    PCA_CHECK_CUDA(cudaFuncSetAttribute(mab_apply4_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, A4Smem::TOTAL));
    if (const char* v = getenv("PCA_TC_APPLY")) g_apply_variant = (v[0] == '3') ? 3 : 4;
    PCA_CHECK_CUDA(cudaFuncSetAttribute(finalize_isab_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F2Smem::TOTAL));
    PCA_CHECK_CUDA(cudaFuncSetAttribute(pma_pool_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PoolSmem::TOTAL));
    PCA_CHECK_CUDA(cudaFuncSetAttribute(pma_pool2_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Pool2Smem::TOTAL));
    PCA_CHECK_CUDA(cudaFuncSetAttribute(pma_pool2_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Pool2Smem::TOTAL));
    if (const char* v = getenv("PCA_TC_POOL")) g_pool_variant = (v[0] == '1') ? 1 : (v[0] == '3' ? 3 : 2);
    if (dev < 64) done_mask.fetch_or(1ull << dev, std::memory_order_release);
    return 0;
}

// the kernels that read the clouds straight from the front end's log-magnitudes are the current generations only
int st_tc_accepts_logmag() { return g_apply_variant == 4 && g_reduce_wg == 6; }

int st_tc_forward_dbg(const float* X, const int* counts, int B, int N, const pca_st_dims* d, const float* params, float* logits,
                      void* ws, size_t ws_bytes, const TcDebug* dbg, cudaStream_t st, const PointSrc* psrc = nullptr) {
    PCA_TRY(tc_configure());
    const size_t one = tc_layout(1, N).total;
    if (!ws || ws_bytes < one) return fail(PCA_EWORKSPACE, "ST(bf16): workspace %zu B < minimum %zu B", ws_bytes, one);
    int chunk = B;
    while (chunk > 1 && tc_layout(chunk, N).total > ws_bytes) chunk = (chunk + 1) / 2;
    if (chunk > 65535) chunk = 65535;
    if (dbg && chunk < B) return fail(PCA_EWORKSPACE, "ST(bf16) debug: workspace must hold the whole batch");
    uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
    TcConsts* c = reinterpret_cast<TcConsts*>(w8 + tc_layout(chunk, N).consts);
    prep_kernel<<<30, 256, 0, st>>>(params, d->d_in, c);
    PCA_CHECK_LAUNCH("prep_kernel");
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int bc = (B - b0) < chunk ? (B - b0) : chunk;
        // the layout of a smaller last chunk fits inside the layout of `chunk` (same consts offset 0)
        PointSrc src{nullptr, nullptr, nullptr, 1};
        if (psrc != nullptr) {
            src = *psrc;
            src.logmag += (size_t)b0 * N;
        }
        PCA_TRY(st_tc_chunk(X ? X + (size_t)b0 * N * d->d_in : nullptr, src, counts ? counts + b0 : nullptr, bc, N, d, params,
                            logits + (size_t)b0 * d->C, w8, c, dbg, st));
    }
    return 0;
}

int st_tc_forward(const float* X, const int* counts, int B, int N, const pca_st_dims* d, const float* params, float* logits,
                  void* ws, size_t ws_bytes, cudaStream_t st) {
    return st_tc_forward_dbg(X, counts, B, N, d, params, logits, ws, ws_bytes, nullptr, st);
}

// ST.forward on the clouds of the whole-path call without a selection step: cloud b, point n = t * nf + f is
// (farr[f], [tarr[t],] logmag[b][n]); no (B, N, d_in) tensor exists
int st_tc_forward_logmag(const float* logmag, const float* farr, const float* tarr, int nf, int B, int N, const pca_st_dims* d,
                         const float* params, float* logits, void* ws, size_t ws_bytes, cudaStream_t st) {
    PCA_TRY(tc_configure());
    if (!st_tc_accepts_logmag()) return fail(PCA_EUNSUPPORTED, "ST(bf16) from log-magnitudes: needs the current kernel generations");
    if (nf <= 0 || N % nf != 0) return fail(PCA_EINVAL, "ST(bf16) from log-magnitudes: N=%d is not a multiple of nf=%d", N, nf);
    const PointSrc src{logmag, farr, tarr, nf};
    return st_tc_forward_dbg(nullptr, nullptr, B, N, d, params, logits, ws, ws_bytes, nullptr, st, &src);
}

int st_tc_forward_stages(const float* X, int B, int N, const pca_st_dims* d, const float* params, float* logits,
                         float* H1, float* Y1, float* H2, float* Y2, float* pooled, void* ws, size_t ws_bytes,
                         cudaStream_t st) {
    TcDebug dbg{H1, Y1, H2, Y2, pooled};
    return st_tc_forward_dbg(X, nullptr, B, N, d, params, logits, ws, ws_bytes, &dbg, st);
}

}  // namespace pca
