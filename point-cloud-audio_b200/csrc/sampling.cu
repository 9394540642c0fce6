// Random-K and importance subsampling of spectral point clouds:
//   ESC_pc_temp_randKSS.__getitem__            (Code/dataset.py:230-238)  np.random.permutation(P)[:K]
//   ESC_pc_temp_importancerandKSS.__getitem__  (Code/dataset.py:276-290)  |grad| heat map, Kaiser smoothing, multinomial / top-K
//   pc_randK                                   (Code/utils.py:55-82)
// A uniformly random K-subset in uniformly random order = the K largest of i.i.d. uniform keys, so random-K reuses the
// radix-select top-K kernel on counter-based random keys; the importance top-K reuses it on the heat map.  The host
// generators of the reference (numpy MT19937 permutation, torch Philox multinomial) cannot be matched bit for bit: parity of
// the random modes is distributional, the heat map and its top-K are deterministic and tested against the reference.
#include "common.cuh"
#include <math.h>

namespace pca {

__device__ __forceinline__ unsigned long long mix64(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);        // splitmix64 finaliser, counter based
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// keys (n) in (0, 1): 24 random bits each, distinct counters -> i.i.d. uniform (ties broken by the top-K rule)
__global__ void random_keys_kernel(float* __restrict__ keys, long long n, unsigned long long seed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ((float)(mix64(seed, (unsigned long long)i) >> 40) + 0.5f) * (1.0f / 16777216.0f);
}

int launch_random_keys(float* keys, long long n, unsigned long long seed, cudaStream_t st) {
    if (n <= 0) return 0;
    random_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys, n, seed);
    PCA_CHECK_LAUNCH("random_keys_kernel");
    return 0;
}

// pts (n_clouds, K, 3|2) rows (farr[f], [tarr[t],] logmag[c, t, f]) for the flat indices p = t*nf + f in idx; idx < 0 -> zero row
__global__ void gather_points_kernel(const float* __restrict__ logmag, int nf, int nt, const float* __restrict__ farr,
                                     const float* __restrict__ tarr, const int32_t* __restrict__ idx, int K, long long total,
                                     float* __restrict__ pts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long c = i / K;
    const int p = idx[i];
    const int w = tarr ? 3 : 2;
    float* o = pts + i * w;
    if (p < 0 || p >= nf * nt) {
        for (int j = 0; j < w; ++j) o[j] = 0.f;
        return;
    }
    const int t = p / nf, f = p - t * nf;
    o[0] = __ldg(farr + f);
    if (tarr) o[1] = __ldg(tarr + t);
    o[w - 1] = __ldg(logmag + (c * nt + t) * nf + f);
}

int launch_gather_points(const float* logmag, int n_clouds, int nf, int nt, const float* farr, const float* tarr,
                         const int32_t* idx, int K, float* pts, cudaStream_t st) {
    const long long total = (long long)n_clouds * K;
    if (total <= 0) return 0;
    gather_points_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(logmag, nf, nt, farr, tarr, idx, K, total, pts);
    PCA_CHECK_LAUNCH("gather_points_kernel");
    return 0;
}

// g (c, f, t) = |d x / d f| + |d x / d t| with torch.gradient's unit-spacing rule (central differences inside, one-sided at
// the edges; Code/dataset.py:280-281).  x(f, t) = logmag[c, t, f].  A dimension of size 1 has no gradient in torch; here it
// contributes 0.
__global__ void grad_abs_kernel(const float* __restrict__ logmag, int nf, int nt, long long total, float* __restrict__ g) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over (c, t, f), f fastest (coalesced reads)
    if (i >= total) return;
    const int f = (int)(i % nf);
    const long long ct = i / nf;
    const int t = (int)(ct % nt);
    const long long c = ct / nt;
    const float* x = logmag + c * (long long)nt * nf;
    auto X = [&](int ff, int tt) { return __ldg(x + (long long)tt * nf + ff); };
    float gf = 0.f, gt = 0.f;
    if (nf > 1) gf = f == 0 ? X(1, t) - X(0, t) : (f == nf - 1 ? X(nf - 1, t) - X(nf - 2, t) : (X(f + 1, t) - X(f - 1, t)) * 0.5f);
    if (nt > 1) gt = t == 0 ? X(f, 1) - X(f, 0) : (t == nt - 1 ? X(f, nt - 1) - X(f, nt - 2) : (X(f, t + 1) - X(f, t - 1)) * 0.5f);
    g[(c * nf + f) * nt + t] = fabsf(gf) + fabsf(gt);
}

// heat (c, f, t) = 1e-6 + sum_{a < wf, b < wt} kf[a] kt[b] g[f + a - (wf-1)/2, t + b - (wt-1)/2]   (zero padding):
// F.conv2d(g, kf kt^T, padding='same') of Code/dataset.py:282-283 (cross-correlation; 'same' pads (k-1)/2 on the left).
__global__ void smooth_kernel(const float* __restrict__ g, int nf, int nt, const float* __restrict__ kf, int wf,
                              const float* __restrict__ kt, int wt, long long total, float* __restrict__ heat) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over (c, f, t), t fastest
    if (i >= total) return;
    const int t = (int)(i % nt);
    const long long cf = i / nt;
    const int f = (int)(cf % nf);
    const long long c = cf / nf;
    const float* gc = g + c * (long long)nf * nt;
    const int pf = (wf - 1) / 2, pt = (wt - 1) / 2;
    float acc = 0.f;
    for (int a = 0; a < wf; ++a) {
        const int ff = f + a - pf;
        if (ff < 0 || ff >= nf) continue;
        const float ka = __ldg(kf + a);
        for (int b = 0; b < wt; ++b) {
            const int tt = t + b - pt;
            if (tt < 0 || tt >= nt) continue;
            acc = fmaf(ka * __ldg(kt + b), gc[(long long)ff * nt + tt], acc);
        }
    }
    heat[i] = acc + 1.0e-6f;
}

int launch_importance_map(const float* logmag, int n_clouds, int nf, int nt, const float* kf, int wf, const float* kt, int wt,
                          float* heat, float* scratch, cudaStream_t st) {
    const long long total = (long long)n_clouds * nf * nt;
    if (total <= 0) return 0;
    if (wf < 1 || wt < 1) return fail(PCA_EINVAL, "importance map: empty smoothing kernel");
    grad_abs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(logmag, nf, nt, total, scratch);
    PCA_CHECK_LAUNCH("grad_abs_kernel");
    smooth_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(scratch, nf, nt, kf, wf, kt, wt, total, heat);
    PCA_CHECK_LAUNCH("smooth_kernel");
    return 0;
}

// torch.multinomial(w, K, replacement=True) (Code/dataset.py:285): K independent draws from the categorical distribution
// w / sum(w), by inversion of the cumulative sums.  One block per cloud: per-thread partial sums over contiguous ranges,
// block scan, cumulative sums written to `cdf` (double), then K binary searches with counter-based uniforms.
constexpr int MN_THREADS = 256;
__global__ void __launch_bounds__(MN_THREADS) multinomial_kernel(const float* __restrict__ w, int n, int K, unsigned long long seed,
                                                                 double* __restrict__ cdf, int32_t* __restrict__ idx) {
    __shared__ double part[MN_THREADS];
    const int c = blockIdx.x, tid = threadIdx.x;
    const float* wc = w + (long long)c * n;
    double* cc = cdf + (long long)c * n;
    const int per = (n + MN_THREADS - 1) / MN_THREADS;
    const int lo = tid * per, hi = min(n, lo + per);
    double s = 0.0;
    for (int i = lo; i < hi; ++i) s += (double)fmaxf(wc[i], 0.f);
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        double run = 0.0;
        for (int i = 0; i < MN_THREADS; ++i) { const double v = part[i]; part[i] = run; run += v; }
    }
    __syncthreads();
    double run = part[tid];
    for (int i = lo; i < hi; ++i) { run += (double)fmaxf(wc[i], 0.f); cc[i] = run; }
    __syncthreads();
    const double total = cc[n - 1];
    for (int k = tid; k < K; k += MN_THREADS) {
        const unsigned long long r = mix64(seed, (unsigned long long)c * (unsigned long long)K + k);
        const double u = ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0) * total;      // (0, total)
        int a = 0, b = n - 1;                     // first index with cdf >= u
        while (a < b) {
            const int m = (a + b) >> 1;
            if (cc[m] >= u) b = m; else a = m + 1;
        }
        idx[(long long)c * K + k] = a;
    }
}

int launch_multinomial(const float* w, int n_clouds, int n, int K, unsigned long long seed, double* cdf, int32_t* idx,
                       cudaStream_t st) {
    if (n_clouds <= 0 || K <= 0) return 0;
    if (n <= 0) return fail(PCA_EINVAL, "multinomial: empty categories");
    multinomial_kernel<<<n_clouds, MN_THREADS, 0, st>>>(w, n, K, seed, cdf, idx);
    PCA_CHECK_LAUNCH("multinomial_kernel");
    return 0;
}


// ------------------------------------------------------------------------------------ test-time resampling
// librosa.resample(x, sr_orig, sr_new, res_type='kaiser_fast', scale=True) of the evaluation sweeps (Code/pceval.py:75,
// Code/pc_temp3d_eval.py:74) = resampy 0.2.2's band-limited sinc interpolation (interpn.resample_f): for output sample t,
// with time_register = t / ratio, n = floor(time_register), the left wing sums win(frac + i step) x[n - i] and the right wing
// win(scale - frac + k step) x[n + 1 + k], the window being linearly interpolated in its table (interp_win + eta * interp_delta).
// One thread per output sample; the (<= 64 K entry) tables stay in L2.  resampy itself is not in the image: the filter table is
// rebuilt from resampy's published recipe on the host and the parity of this row is UNPINNED.
__global__ void resample_kernel(const float* __restrict__ x, int n_in, int n_out, double time_increment, double scale,
                                const double* __restrict__ win, const double* __restrict__ delta, int nwin, int num_table,
                                int index_step, float out_scale, float* __restrict__ y) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (t >= n_out) return;
    const float* xb = x + (long long)b * n_in;
    const double time_register = (double)t * time_increment;
    const int n = (int)time_register;
    double acc = 0.0;
    if (n < n_in) {
        double frac = scale * (time_register - (double)n);
        double index_frac = frac * (double)num_table;
        int offset = (int)index_frac;
        double eta = index_frac - (double)offset;
        const int i_max = min(n + 1, (nwin - offset) / index_step);
        for (int i = 0; i < i_max; ++i) {
            const int j = offset + i * index_step;
            acc += (__ldg(win + j) + eta * __ldg(delta + j)) * (double)__ldg(xb + n - i);
        }
        frac = scale - frac;
        index_frac = frac * (double)num_table;
        offset = (int)index_frac;
        eta = index_frac - (double)offset;
        const int k_max = min(n_in - n - 1, (nwin - offset) / index_step);
        for (int k = 0; k < k_max; ++k) {
            const int j = offset + k * index_step;
            acc += (__ldg(win + j) + eta * __ldg(delta + j)) * (double)__ldg(xb + n + k + 1);
        }
    }
    y[(long long)b * n_out + t] = (float)acc * out_scale;
}

int launch_resample(const float* x, int n_clips, int n_in, int n_out, double sample_ratio, const double* win, const double* delta,
                    int nwin, int num_table, float out_scale, float* y, cudaStream_t st) {
    if (n_clips <= 0 || n_out <= 0) return 0;
    if (n_in <= 0 || !(sample_ratio > 0.0) || nwin <= 0 || num_table <= 0) return fail(PCA_EINVAL, "resample: bad arguments");
    if (n_clips > 65535) return fail(PCA_EUNSUPPORTED, "resample: more than 65535 clips per call");
    const double scale = sample_ratio < 1.0 ? sample_ratio : 1.0;
    const int index_step = (int)(scale * num_table);
    if (index_step < 1) return fail(PCA_EUNSUPPORTED, "resample: ratio %g too small for a table of %d entries per zero crossing", sample_ratio, num_table);
    dim3 grid((n_out + 255) / 256, n_clips);
    {
        LaunchTimer lt("resample_kernel", st, 4.0 * n_clips * (double)n_out * (2.0 * nwin / index_step), 4.0 * n_clips * ((double)n_in + n_out));
        resample_kernel<<<grid, 256, 0, st>>>(x, n_in, n_out, 1.0 / sample_ratio, scale, win, delta, nwin, num_table, index_step, out_scale, y);
    }
    PCA_CHECK_LAUNCH("resample_kernel");
    return 0;
}

}  // namespace pca
