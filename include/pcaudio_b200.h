/*
 * pcaudio_b200 -- C ABI of the B200 (sm_100a) audio -> point-cloud -> set-encoder hot path.
 *
 * The reference (SubramaniKrishna/point-cloud-audio) is pure Python and has no FFI layer
 * (SURVEY.md 8b); every entry point below therefore cites the reference *Python* interface
 * whose arithmetic it replaces.  The host-side mirror of those interfaces (same class /
 * function names and argument meaning) lives in point-cloud-audio_b200/ and binds these
 * symbols with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with `host_`;
 *   - all tensors are dense, row-major, float32 unless stated; sizes are element counts;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are
 *     asynchronous on that stream and never call cudaDeviceSynchronize;
 *   - the library allocates no device memory: scratch is caller-provided (`workspace`);
 *   - return 0 on success, a positive cudaError_t, or a negative PCA_E* code; the message is
 *     available from pca_last_error() (thread local).  No exceptions cross the boundary.
 */
#ifndef PCAUDIO_B200_H
#define PCAUDIO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCA_VERSION 102 /* 0.1.2 */

enum {
    PCA_OK = 0,
    PCA_EINVAL = -1,      /* bad shape / null pointer / misaligned */
    PCA_EUNSUPPORTED = -2, /* dims outside what the kernels implement */
    PCA_EWORKSPACE = -3,  /* workspace too small */
    PCA_EDEVICE = -4      /* not an sm_100 device */
};

/* precision of the set-encoder kernels */
enum {
    PCA_PREC_FP32 = 0, /* CUDA-core fp32 everywhere (1e-3 parity path) */
    PCA_PREC_BF16 = 2  /* tcgen05 bf16 operands, fp32 accumulate (2e-2 parity path) */
};

int pca_version(void);
const char* pca_last_error(void);

/* ---------------------------------------------------------------- L2: spectral front end
 * Replaces the inline recipe
 *   x = librosa.stft(x, n_fft, win_length, hop_length=int(N*hf), window='hann') / N
 *   [x = x[:-1,:]] ; a = np.log(1.0e-8 + np.abs(x))
 * Code/settransformer.py:49-50, Code/settransformertemp.py:51-53, Code/pceval.py:76-77.
 *
 * audio   (n_clips, n_samples)
 * window  (n_fft)   periodic Hann of win_length centred/zero-padded to n_fft (host-built in
 *                   float64, rounded once)
 * twiddle (n_fft/2, 2) = (cos, -sin)(2 pi k / n_fft), k < n_fft/2
 * out     (n_clips, nt_out, nf_out), frequency fastest; nf_out = n_fft/2+1-drop_nyquist;
 *         frame t is centred on sample t*hop (reflect padding by index mirroring);
 *         nt_out <= 1 + n_samples/hop frames are produced (the 3ST chunking of
 *         Code/settransformertemp.py:54-58 keeps only floor(Nt/Ntemp)*Ntemp of them).
 * scale   multiplies |STFT| before the log (= 1/N). */
int pca_stft_logmag_f32(const float* audio, int n_clips, int n_samples, int n_fft, int hop,
                        const float* window, const float* twiddle, float scale, int drop_nyquist,
                        int nt_out, float* out, void* stream);

/* ---------------------------------------------------------------- L3: point clouds
 * ESC_pc.__getitem__ (Code/dataset.py:50-54) when tarr == NULL: rows (farr[f], x[f]);
 * ESC_pc_temp.__getitem__ (Code/dataset.py:160-166) otherwise: point p = t*nf + f has
 * columns (farr[f], tarr[t], x[f,t]).  logmag is (n_clouds, nt, nf) f-fastest (nt = 1 for
 * 2-D clouds); pts is (n_clouds, nt*nf, 2|3). */
int pca_build_clouds_f32(const float* logmag, int n_clouds, int nf, int nt, const float* farr,
                         const float* tarr, float* pts, void* stream);

/* ESC_pc_temp_maxKSS.__getitem__ (Code/dataset.py:194-202) and utils.pc_maxK
 * (Code/utils.py:25-52): per cloud, the K largest keys.  With sorted_desc != 0 rows are
 * emitted in the order of (-key).argsort(kind='stable')[:K]; otherwise in ascending flat
 * index (scan) order.  Ties at the K boundary keep the lowest flat indices.
 * pts (n_clouds, K, 2|3) may be NULL (indices only); idx (n_clouds, K) int32 flat indices
 * p = t*nf + f may be NULL.  K <= nf*nt; sorted output needs K <= 16384. */
int pca_topk_compact_f32(const float* keys, int n_clouds, int nf, int nt, const float* farr,
                         const float* tarr, int K, int sorted_desc, float* pts, int32_t* idx,
                         void* stream);

/* Threshold / capped selection with padded output (extension of the rule above): per cloud keep the
 * points with key >= threshold (use_threshold != 0), at most K of them chosen by the top-K rule, emitted as above;
 * rows past the number kept are zero-filled (idx -1) and counts[c] (nullable) receives the number kept.  With
 * use_threshold == 0 this is pca_topk_compact_f32 plus counts[c] = K. */
int pca_select_compact_f32(const float* keys, int n_clouds, int nf, int nt, const float* farr,
                           const float* tarr, int K, int sorted_desc, int use_threshold,
                           float threshold, float* pts, int32_t* idx, int32_t* counts, void* stream);

/* Whole front end in one launch (a1 -> a7 of SURVEY.md 8a; the fused call of SURVEY.md 8b): per clip, STFT ->
 * log-magnitude -> non-overlapping ntemp-frame (f, t, mag) clouds (remainder frames dropped, as
 * Code/settransformertemp.py:54-58) -> selection as pca_select_compact_f32, one thread block per cloud with the
 * cloud's log-magnitudes held in shared memory (HBM traffic: the audio once + 16 B per selected point).
 * pts (n_clips * (Nt / ntemp), K, 3), idx, counts as in pca_select_compact_f32 (any may be NULL, not all).
 * Clouds whose keys + sort buffer exceed 227 KB of shared memory return PCA_EUNSUPPORTED: use the unfused calls. */
int pca_frontend_fused_f32(const float* audio, int n_clips, int n_samples, int n_fft, int hop,
                           const float* window, const float* twiddle, float scale, int drop_nyquist,
                           int ntemp, const float* farr, const float* tarr, int K, int sorted_desc,
                           int use_threshold, float threshold, float* pts, int32_t* idx, int32_t* counts,
                           void* stream);

/* ---------------------------------------------------------------- L4: set encoder
 * Weights are passed as ONE packed float32 blob per block, nn.Linear layout (out, in):
 *   MAB  := Wq (D,dq) | bq (D) | Wk (D,dk) | Wv (D,dk) | bk (D) | bv (D) | Wo (D,D) | bo (D)
 *           [| ln0.w (D) | ln0.b (D) | ln1.w (D) | ln1.b (D)   when ln != 0]
 *   ISAB := I (M,D) | MAB(dq=D, dk=d_in)   /mab0/ | MAB(dq=d_in, dk=D) /mab1/
 *   PMA  := S (S,D) | MAB(dq=D, dk=D)
 *   ST   := ISAB(d_in) | ISAB(D) | PMA | Wl (C,D) | bl (C)
 * pca_*_param_count return the blob length so both sides agree. */
typedef struct {
    int d_in;   /* input point width (2 or 3 for audio) */
    int D;      /* dim_hidden */
    int H;      /* num_heads */
    int M;      /* num_inds */
    int S;      /* num_outputs (PMA seeds) */
    int C;      /* dim_output (classes) */
    int ln;     /* LayerNorm branches of MAB (modules.py:14-16,30,32) */
} pca_st_dims;

long long pca_mab_param_count(int dq, int dk, int D, int ln);
long long pca_isab_param_count(int d_in, int D, int M, int ln);
long long pca_pma_param_count(int D, int S, int ln);
long long pca_st_param_count(const pca_st_dims* dims);

/* MAB.forward (set_transformer-master/modules.py:19-33).
 * Q (Bq, nq, dq) with Bq == B, or Bq == 1 to broadcast one query set over the batch (the
 * I.repeat / S.repeat of modules.py:52,63 is never materialised); K (B, nk, dk);
 * out (B, nq, D). */
size_t pca_mab_workspace_bytes(int B, int nq, int nk, int dq, int dk, int D, int H);
int pca_mab_fwd_f32(const float* Q, int q_batch, const float* K, int B, int nq, int nk, int dq,
                    int dk, int D, int H, int ln, const float* params, float* out,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ISAB.forward (modules.py:51-53): X (B,N,d_in) -> out (B,N,D). */
size_t pca_isab_workspace_bytes(int B, int N, int d_in, int D, int H, int M);
int pca_isab_fwd_f32(const float* X, int B, int N, int d_in, int D, int H, int M, int ln,
                     const float* params, float* out, void* workspace, size_t workspace_bytes,
                     void* stream);

/* PMA.forward (modules.py:62-63): X (B,N,D) -> out (B,S,D). */
size_t pca_pma_workspace_bytes(int B, int N, int D, int H, int S);
int pca_pma_fwd_f32(const float* X, int B, int N, int D, int H, int S, int ln,
                    const float* params, float* out, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ST.forward (Code/models.py:43-44) == main_pointcloud.SetTransformer.forward in eval mode
 * (set_transformer-master/main_pointcloud.py:36-37): X (B,N,d_in) -> logits (B,S,C)
 * (the caller applies the reference's .squeeze()).  The batch is processed in chunks that
 * fit `workspace_bytes` (any size >= pca_st_workspace_bytes(dims, 1, N, precision) works;
 * larger is faster).  precision: PCA_PREC_FP32 (all dims) or PCA_PREC_BF16 (tcgen05 path;
 * D=64, H=8, M=64, S=1, ln=0 only -- anything else returns PCA_EUNSUPPORTED). */
size_t pca_st_workspace_bytes(const pca_st_dims* dims, int B, int N, int precision);
int pca_st_fwd(const float* X, int B, int N, const pca_st_dims* dims, const float* params,
               float* logits, void* workspace, size_t workspace_bytes, int precision,
               void* stream);

/* Variable-size sets (extension; the reference batches only equal-size sets): cloud b consists of the first counts[b]
 * (1 <= counts[b] <= N) rows of its padded (N, d_in) slot; padding rows never act as keys of mab0 / PMA, so the logits
 * equal ST.forward on X[b:b+1, :counts[b]].  counts == NULL is pca_st_fwd.  A cloud with counts[b] <= 0 has no answer
 * (the reference fails on an empty set): its logits are NaN; counts above N are clamped to N. */
int pca_st_fwd_masked(const float* X, const int32_t* counts, int B, int N, const pca_st_dims* dims,
                      const float* params, float* logits, void* workspace, size_t workspace_bytes,
                      int precision, void* stream);

/* nn.Linear on rows: Y (rows, dout) = X (rows, din) W^T + b, params := W (dout, din) | b (dout).  The final Linear of the
 * generic SetTransformer (set_transformer-master/models.py:41) after its SAB decoder blocks. */
int pca_linear_fwd_f32(const float* X, long long rows, int din, int dout, const float* params, float* Y, void* stream);
/* Backward of that Linear (autograd of torch.nn.Linear in the reference's training loops, e.g. set_transformer-master/run.py:
 * 97-100 through models.py:41): dparams := dW (dout, din) | db (dout) is OVERWRITTEN with dY^T X and the column sums of dY;
 * dX (rows, din) = dY W is written when dX != NULL. */
int pca_linear_bwd_f32(const float* dY, const float* X, long long rows, int din, int dout, const float* params, float* dX,
                       float* dparams, void* stream);
/* nn.Dropout (set_transformer-master/main_pointcloud.py:30,32) with a counter-based mask: out[i] = keep(seed, i) ? in[i] / (1 - p)
 * : 0.  The mask is a pure function of (seed, i), so the backward pass is the same call on the gradient (no mask tensor);
 * in == out is allowed.  torch's Philox stream cannot be matched: parity is statistical. */
int pca_dropout_f32(const float* in, float* out, long long n, float p, unsigned long long seed, void* stream);

/* DeepSet.forward (set_transformer-master/models.py:25-28) / SmallDeepSet
 * (max_regression_demo.ipynb:41-48): 4 shared Linear (+ReLU) over points, pool over points
 * (0 mean, 1 max, 2 sum), 4 Linear decoder.  params := for enc then dec, 4 x (W (out,in) | b).
 * X (B,N,d_in) -> out (B, num_outputs*dim_output). */
size_t pca_deepset_workspace_bytes(int B, int N, int d_in, int dim_hidden, int out_dim);
int pca_deepset_fwd_f32(const float* X, int B, int N, int d_in, int dim_hidden, int out_dim,
                        int pool, const float* params, float* out, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Masked pooling variant (north_star: PointNet-style shared MLP + masked max-pool): pools over the first counts[b] points. */
int pca_deepset_fwd_masked_f32(const float* X, const int32_t* counts, int B, int N, int d_in, int dim_hidden,
                               int out_dim, int pool, const float* params, float* out, void* workspace,
                               size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- random-K / importance subsampling
 * ESC_pc_temp_randKSS (Code/dataset.py:230-238), pc_randK (Code/utils.py:55-82): a uniformly random K-subset in random
 * order = the K largest of i.i.d. uniform keys -> pca_random_keys_f32 (counter-based generator, keys in (0,1)) followed by
 * pca_topk_compact_f32 (indices only) and pca_gather_points_f32.  numpy's permutation stream cannot be reproduced: parity
 * is distributional.
 * ESC_pc_temp_importancerandKSS (Code/dataset.py:276-290): pca_importance_map_f32 computes
 *   heat (n_clouds, nf, nt) [t fastest = g.view(-1) of the reference] =
 *       conv2d(|d x/d f| + |d x/d t|, kf kt^T, padding='same') + 1e-6,   x(f, t) = logmag[c, t, f], torch.gradient rule,
 * kf (wf) / kt (wt) the two Kaiser windows (wf = 2, wt = winF in the reference); scratch: n_clouds*nf*nt floats.
 * choice 1 = pca_topk_compact_f32 on the heat map; choice 0 = pca_multinomial_f32 (K draws with replacement from
 * weights / sum(weights); cdf_scratch: n_clouds*n doubles).  NOTE the reference indexes the (t-major) cloud with the
 * (f-major) heat-map index; pca_gather_points_f32 takes flat indices in CLOUD order p = t*nf + f, so passing the heat-map
 * indices unchanged reproduces the reference's behaviour.  idx < 0 gathers a zero row. */
int pca_random_keys_f32(float* keys, long long n, unsigned long long seed, void* stream);
int pca_gather_points_f32(const float* logmag, int n_clouds, int nf, int nt, const float* farr, const float* tarr,
                          const int32_t* idx, int K, float* pts, void* stream);
int pca_importance_map_f32(const float* logmag, int n_clouds, int nf, int nt, const float* kf, int wf, const float* kt,
                           int wt, float* heat, float* scratch, void* stream);
int pca_multinomial_f32(const float* weights, int n_clouds, int n, int K, unsigned long long seed, double* cdf_scratch,
                        int32_t* idx, void* stream);

/* ---------------------------------------------------------------- test-time resampling
 * librosa.resample(x, sr_orig, sr_new, res_type='kaiser_fast', scale=True) (Code/pceval.py:75, Code/pc_temp3d_eval.py:74) =
 * resampy 0.2.2 band-limited sinc interpolation.  x (n_clips, n_in) -> y (n_clips, n_out); win / delta: the half window of
 * the interpolation filter and its first differences (float64, nwin entries, num_table entries per zero crossing; already
 * multiplied by the ratio when downsampling, as resampy does); every output is multiplied by out_scale (1/sqrt(ratio) for
 * scale=True).  Outputs past int(n_in * ratio) are the zero padding of librosa's fix_length.  resampy is not available in
 * this image: the table comes from its published recipe and the parity of this entry point is unpinned. */
int pca_resample_f32(const float* x, int n_clips, int n_in, int n_out, double sample_ratio, const double* win,
                     const double* delta, int nwin, int num_table, float out_scale, float* y, void* stream);

/* ---------------------------------------------------------------- training (fp32, every dim; ln = 0)
 * The reference trains these models with loss.backward() + torch.optim.Adam under nn.DataParallel
 * (Code/settransformer.py:89-109, Code/settransformertemp.py:110-128, set_transformer-master/main_pointcloud.py:61-79).
 * Here the training forward keeps the activations the hand-derived backward needs in a caller-owned buffer
 * (`saved`, pca_st_train_saved_bytes), and the backward returns the gradient of EVERY parameter as one flat blob in
 * the layout of `params` -- the unit of the per-step gradient allreduce and of pca_adam_step_f32.
 * dropout_p > 0 is nn.Dropout(p) before and after the PMA (main_pointcloud.py:30-33); the (seed, element index) mask
 * is regenerated by the backward call, which must be given the same seed.  counts (B) int32, nullable: variable-size
 * sets as in pca_st_fwd_masked (extension) -- rows >= counts[b] never act as keys and receive zero input gradient.
 * dlogits (B, S, C); dparams (pca_st_param_count) is overwritten; dX (B, N, d_in) may be NULL. */
size_t pca_st_train_saved_bytes(const pca_st_dims* dims, int B, int N, float dropout_p);
size_t pca_st_train_workspace_bytes(const pca_st_dims* dims, int B, int N);
int pca_st_train_fwd_f32(const float* X, const int32_t* counts, int B, int N, const pca_st_dims* dims, const float* params,
                         float dropout_p, unsigned long long seed, float* logits, void* saved, size_t saved_bytes,
                         void* workspace, size_t workspace_bytes, void* stream);
int pca_st_train_bwd_f32(const float* X, const int32_t* counts, int B, int N, const pca_st_dims* dims, const float* params,
                         float dropout_p, unsigned long long seed, const float* dlogits, const void* saved,
                         size_t saved_bytes, float* dparams, float* dX, void* workspace, size_t workspace_bytes,
                         void* stream);
/* The same backward in two phases, for data-parallel training (replaces the reduce_add of nn.DataParallel,
 * set_transformer-master/main_pointcloud.py:65): phase 1 differentiates the final Linear, the PMA and ISAB 1, after which
 * dparams[*tail_offset ..) is final and its all-reduce can start; phase 2 differentiates ISAB 0 (dparams[.. *tail_offset),
 * dX) from the workspace phase 1 left behind (same workspace, untouched in between).  phase 0 = pca_st_train_bwd_f32.
 * tail_offset (host pointer, may be NULL) receives the float index where the tail begins. */
int pca_st_train_bwd_phase_f32(const float* X, const int32_t* counts, int B, int N, const pca_st_dims* dims, const float* params,
                               float dropout_p, unsigned long long seed, const float* dlogits, const void* saved,
                               size_t saved_bytes, float* dparams, float* dX, void* workspace, size_t workspace_bytes,
                               int phase, long long* tail_offset, void* stream);

/* Stand-alone MAB training (modules.py:6-33, LayerNorm branches included) for models composed from the blocks on the host (SAB = MAB(X, X),
 * ISAB = mab1(X, mab0(I, X)), PMA = MAB(S, X)): Q (q_batch, nq, dq) with q_batch in {1, B} (1 = shared queries: dQ is summed
 * over the batch), K (B, nk, dk), out / dout (B, nq, D).  dparams in the MAB blob layout; dQ / dK may be NULL. */
size_t pca_mab_train_saved_bytes(int B, int q_batch, int nq, int nk, int D, int H, int ln);
size_t pca_mab_train_workspace_bytes(int B, int q_batch, int nq, int nk, int D, int H, int ln);
int pca_mab_train_fwd_f32(const float* Q, int q_batch, const float* K, int B, int nq, int nk, int dq, int dk, int D, int H, int ln,
                          const float* params, float* out, void* saved, size_t saved_bytes, void* workspace,
                          size_t workspace_bytes, void* stream);
int pca_mab_train_bwd_f32(const float* Q, int q_batch, const float* K, int B, int nq, int nk, int dq, int dk, int D, int H, int ln,
                          const float* params, const float* dout, const void* saved, size_t saved_bytes, float* dparams,
                          float* dQ, float* dK, void* workspace, size_t workspace_bytes, void* stream);

/* DeepSet training (set_transformer-master/models.py:3-28; pool 0 mean / 1 max / 2 sum): forward that keeps the activations
 * (and the arg-max points of the max pool), backward returning the flat parameter gradient in the layout of `params`
 * (enc then dec, 4 x (W | b)) and optionally dX.  out / dout (B, out_dim).  Equal-size sets. */
size_t pca_deepset_train_saved_bytes(int B, int N, int dim_hidden);
size_t pca_deepset_train_workspace_bytes(int B, int N, int dim_hidden);
int pca_deepset_train_fwd_f32(const float* X, int B, int N, int d_in, int dim_hidden, int out_dim, int pool, const float* params,
                              float* out, void* saved, size_t saved_bytes, void* workspace, size_t workspace_bytes, void* stream);
int pca_deepset_train_bwd_f32(const float* X, int B, int N, int d_in, int dim_hidden, int out_dim, int pool, const float* params,
                              const float* dout, const void* saved, size_t saved_bytes, float* dparams, float* dX,
                              void* workspace, size_t workspace_bytes, void* stream);

/* nn.CrossEntropyLoss(reduction='mean') on logits (B, C) with int64 labels: loss[0] += mean loss, correct[0] +=
 * number of rows whose arg-max equals the label (both caller-zeroed, either may be NULL except loss),
 * dlogits (B, C) = d loss / d logits (may be NULL).  Labels must lie in [0, C): ignore_index is NOT supported; a row with
 * an out-of-range label turns the loss and that row of dlogits into NaN (torch raises a device-side assert there). */
int pca_cross_entropy_f32(const float* logits, const int64_t* labels, int B, int C, float* loss, int32_t* correct,
                          float* dlogits, void* stream);

/* torch.optim.Adam step (L2 weight_decay folded into the gradient, bias-corrected moments; `step` counts from 1) over
 * a flat fp32 blob in one launch; grads are multiplied by grad_scale first (1/world_size after a summing allreduce). */
int pca_adam_step_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

/* ---------------------------------------------------------------- whole path, one call
 * audio -> STFT/log-magnitude -> [Ntemp chunking] -> [top-K] -> clouds -> ST -> logits.
 * This is what the reference's eval loop does per batch (Code/pc_temp3d_eval.py:126-185,
 * Code/pceval.py:73-99) minus file I/O.  mode 2: one (f,mag) cloud per frame (FST);
 * mode 3: one (f,t,mag) cloud per ntemp-frame chunk (3ST; ntemp = frames per clip gives
 * clip-as-cloud).  top_k == 0 keeps every point. */
typedef struct {
    int n_samples;
    int n_fft;
    int hop;
    float scale;        /* 1/N applied to |STFT| */
    int mode;           /* 2 = FST frame clouds, 3 = 3ST chunk clouds */
    int ntemp;          /* mode 3: frames per cloud */
    int top_k;          /* 0 = all points */
    int precision;      /* PCA_PREC_* for the encoder */
    pca_st_dims st;
    int use_threshold;  /* != 0: keep only points with log-magnitude >= threshold (capped at top_k if top_k > 0), */
    float threshold;    /*       padded to the full width; the encoder masks the padding (variable-size sets)      */
} pca_pipeline_cfg;

/* number of clouds produced per clip and points per cloud for a config */
int pca_pipeline_clouds_per_clip(const pca_pipeline_cfg* cfg);
int pca_pipeline_points_per_cloud(const pca_pipeline_cfg* cfg);
size_t pca_pipeline_workspace_bytes(const pca_pipeline_cfg* cfg, int n_clips);

/* Device-resident variant: audio (n_clips, n_samples) and logits
 * (n_clips*clouds_per_clip, S, C) are device buffers. */
int pca_pipeline_run(const pca_pipeline_cfg* cfg, const float* audio, int n_clips,
                     const float* window, const float* twiddle, const float* farr,
                     const float* tarr, const float* st_params, float* logits, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Host-buffer variant (the end-to-end call): host_audio / host_logits are HOST pointers
 * (pinned for asynchronous copies); dev_audio / dev_logits are caller-owned device staging
 * buffers.  Enqueues H2D copy, the kernels and the D2H copy on `stream`; the caller
 * synchronises the stream before reading host_logits. */
int pca_pipeline_run_host(const pca_pipeline_cfg* cfg, const float* host_audio, int n_clips,
                          float* dev_audio, const float* window, const float* twiddle,
                          const float* farr, const float* tarr, const float* st_params,
                          float* dev_logits, float* host_logits, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Same, with the host->device copy overlapped with compute: the batch is cut into n_chunks (1..16) clip ranges (a short head range of 1/8 of the clips so the kernels start
 * early, the rest split evenly); chunk
 * k+1 is copied on `copy_stream` while chunk k runs on `stream` (ordering by events created inside the call; both
 * streams are the caller's, the library keeps no state).  copy_stream == stream or n_chunks == 1 degrades to
 * pca_pipeline_run_host.  Results are identical to the unchunked call (per-cloud arithmetic never depends on the
 * batch).  The workspace only needs to hold the largest chunk. */
int pca_pipeline_run_host_chunked(const pca_pipeline_cfg* cfg, const float* host_audio, int n_clips,
                                  float* dev_audio, const float* window, const float* twiddle,
                                  const float* farr, const float* tarr, const float* st_params,
                                  float* dev_logits, float* host_logits, void* workspace,
                                  size_t workspace_bytes, int n_chunks, void* copy_stream, void* stream);

/* number of kernels the library has launched in this process (for bench.py gpu_launches) */
unsigned long long pca_launch_count(void);

/* Measurement aid: while enabled, every kernel launch is bracketed by CUDA events on its stream.
 * pca_profile_report synchronises those events and writes a JSON object
 *   {"<kernel>": {"launches": n, "ms": total device ms, "flops": algorithmic flops, "bytes": algorithmic bytes}, ...}
 * into buf, then clears the records.  Enabling/disabling also clears them. */
void pca_profile_enable(int on);
int pca_profile_report(char* buf, size_t buf_len);

/* Debug variant of pca_st_fwd(precision = PCA_PREC_BF16) that also returns the intermediate stages as fp32:
 * H1, H2 (B, M, D) inducing-point summaries of the two ISABs, Y1, Y2 (B, N, D) ISAB outputs, pooled (B, D) PMA
 * output before the final Linear.  Any stage pointer may be NULL.  The workspace must hold the whole batch. */
int pca_debug_st_stages(const float* X, int B, int N, const pca_st_dims* dims, const float* params,
                        float* logits, float* H1, float* Y1, float* H2, float* Y2, float* pooled,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Debug: when set to a device buffer of >= 8000 int64, CTA 0 / warp 0 of the first reduce kernel of every bf16 ST forward
 * records (phase tag, clock64) pairs of its softmax loop there.  NULL switches it off. */
void pca_debug_set_timeline(long long* device_buffer);

/* Debug / experiments: the bf16 path hands clouds with 1..tail_max (<= 4, the default) points past a multiple of 128 to exact
 * fp32 tail paths instead of a ninth, nearly empty tensor-core tile.  0 switches the rule off. */
void pca_debug_set_tail_max(int tail_max);

/* Debug / experiments: variant of the mab0 (reduce) kernel of the bf16 path: 6 (default) = sixth generation (W_k folded into the
 * query operand, TMA-fed tiles, one chain per head pair) and 4 = the fifth generation's four streaming softmax warpgroups, each
 * followed by the exact 2-warpgroup variant on the work items in which a row outgrew its reference exponent (normally none);
 * 2 = the exact 2-warpgroup variant only. */
void pca_debug_set_reduce_variant(int warpgroups);

/* Debug / experiments: pooled-attention (PMA) kernel of the bf16 path: 2 (default) = transposed formulation (thread = point,
 * 8 exponentials per point), streaming pass against a fixed reference exponent with the sums accumulating in TMEM + exact redo
 * of the work items it flags; 3 = the exact (per-tile re-referencing) pass on every work item; 1 = the round-1 kernel whose rows
 * are (head, copy) pairs.  Also PCA_TC_POOL=1|2|3 in the environment. */
void pca_debug_set_pool_variant(int variant);

/* Debug / tests: 1 = run the generic STFT kernel (any power-of-two n_fft, any thread count; the code the fused front end
 * uses) also for n_fft in [512, 4096], where the size-specialised kernel is the default.  Both give bit-identical spectra. */
void pca_debug_set_stft_generic(int on);

/* fp32-grade GEMMs on the tensor cores (split-bf16, three MMAs per product; csrc/gemm_tc.cu), used by the fp32 encoder path,
 * DeepSet's shared MLP and the training path for layers with K % 32 == 0 and N % 32 == 0.  pca_debug_set_gemm_tc(0) keeps
 * every layer on the CUDA-core kernels.  pca_debug_linear_tc: Y (rows, N) = act(X (rows, K) B^T + bias) [+ resid], B(n,k) =
 * trans_w ? W[k*N+n] : W[n*K+k]; R (nullable) receives the activation output before the residual; image: >= 4*N*K bytes of
 * scratch.  pca_debug_grad_weight_tc: dW (M, N) += dY (rows, M)^T X (rows, N). */
void pca_debug_set_gemm_tc(int on);
int pca_debug_linear_tc(const float* X, const float* W, int trans_w, const float* bias, const float* resid, float* Y, float* R,
                        long long rows, int K, int N, int relu, void* image, size_t image_bytes, void* stream);
int pca_debug_grad_weight_tc(const float* dY, const float* X, float* dW, long long rows, int M, int N, void* stream);

/* Attention of a MAB with one small side -- at most 16 queries (ISAB mab0, PMA) or at most 16 keys (ISAB mab1) against >= 128
 * items, dim_V <= 256 a multiple of 32, head dim a multiple of 8, num_heads * 8 or * 16 in {32, 64} (the ModelNet model of
 * set_transformer-master/main_pointcloud.py:62) -- runs Q K^T, P V and every contraction of their gradient as split-bf16
 * tcgen05 GEMMs (csrc/attn_tc.cu; fp32-grade, replaces set_transformer-master/modules.py:28-29 and its autograd).  It is what
 * the fp32 encoder and the training path use for eligible shapes; pca_debug_set_attn_tc(0) keeps them on the CUDA-core kernels,
 * (2) keeps the route but makes training use the projected K | V form for the shared-query blocks (default 1: those blocks run
 * on the un-projected points, forward and backward).
 * pca_debug_attn_fwd: O (B, nq, D) = Qp + softmax_h(Qp K^T / sqrt(D)) V on projected operands -- Qp (B or 1, nq, D), KV
 * (B, nk, 2D) rows [K | V] -- through the same dispatch; lse (B, nq, H) nullable receives log2 sum_k 2^(s_k log2 e).
 * pca_debug_attn_bwd_tc: the tensor-core backward alone: dQp (B, nq, D) = dO + dS K, dKV (B, nk, 2D) = [dS^T Qp | P^T dO];
 * delta (B, nq, H) = sum_d dO (O - Qp) per head is read when the queries are the small side.  PCA_EUNSUPPORTED for
 * ineligible shapes.  Workspace: pca_debug_attn_ws_bytes. */
void pca_debug_set_attn_tc(int on);
int pca_debug_attn_tc_eligible(int B, int nq, int nk, int D, int H);
size_t pca_debug_attn_ws_bytes(int B, int nq, int nk, int D, int H);
int pca_debug_attn_fwd(const float* Qp, int q_shared, const float* KV, int B, int nq, int nk, int D, int H, float* O, float* lse,
                       void* ws, size_t ws_bytes, void* stream);
int pca_debug_attn_bwd_tc(const float* Qp, int q_shared, const float* KV, const float* dO, const float* lse, const float* delta, int B,
                          int nq, int nk, int D, int H, float* dQp, float* dKV, void* ws, size_t ws_bytes, void* stream);

/* Unit probe of the tcgen05 building blocks used by the bf16 encoder path: one CTA computes
 * D (128, N) = A (128, K) * B (K, N), bf16 operands, fp32 accumulation in TMEM.
 * a_mode: 0 A (128,K) via shared memory K-major, 1 A via TMEM, 2 A given as (K,128) via shared memory MN-major;
 * b_mode: 0 B given as (N,K) K-major, 1 B given as (K,N) MN-major.  N, K multiples of 16 in [16,128]. */
int pca_debug_umma_probe(const float* A, const float* B, float* D, int N, int K, int a_mode, int b_mode,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCAUDIO_B200_H */
