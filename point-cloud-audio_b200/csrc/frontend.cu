// Spectral front end and point selection for sm_100a.
//
//   stft_logmag_kernel : reflect-padded framing (index mirroring, no padded copy) -> periodic
//                        Hann -> real FFT (packed as a half-length complex Stockham FFT in shared
//                        memory: radix-8 stages in registers + one radix-4 / radix-2 stage) -> |.|*scale -> log(1e-8 + .)
//                        replaces librosa.stft + np.log recipe (Code/settransformer.py:49-50,
//                        Code/settransformertemp.py:51-53).
//   build_clouds_kernel: ESC_pc / ESC_pc_temp __getitem__ (Code/dataset.py:50-54,160-166).
//   topk_kernel        : ESC_pc_temp_maxKSS / pc_maxK selection (Code/dataset.py:194-202,
//                        Code/utils.py:42-45): 4-pass 8-bit radix select on order-preserving
//                        keys, warp-ballot/prefix-sum compaction in flat-index order (ties keep
//                        the lowest indices), bitonic sort of the K survivors for the
//                        descending-magnitude emission order.
#include "common.cuh"
#include <math.h>

namespace pca {

// ------------------------------------------------------------------------------------ STFT
// slots of one padded FFT buffer (see stft_frame)
__host__ __device__ static inline int stft_buf_len(int nc) { return nc + (nc >> 4); }

// Every floating-point operation of a frame is spelled with a rounding intrinsic: the compiler then cannot contract a
// product into a neighbouring sum differently in the generic and the size-specialised code (or from one compiler version to
// the next), so all STFT kernels give bit-identical spectra.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(__fmaf_rn(a.x, b.x, -__fmul_rn(a.y, b.y)), __fmaf_rn(a.x, b.y, __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cwin(float2 w, float2 v) { return make_float2(__fmul_rn(w.x, v.x), __fmul_rn(w.y, v.y)); }

// ---- arithmetic shared by the generic and the size-specialised frame code
// 8-point DFT, outputs in natural order: even outputs = DFT4 of (a_j + a_{j+4}), odd = DFT4 of ((a_j - a_{j+4}) w8^j)
__device__ __forceinline__ void dft8(const float2 (&a)[8], float2 (&y)[8]) {
    const float2 b0 = cadd(a[0], a[4]), b1 = cadd(a[1], a[5]), b2 = cadd(a[2], a[6]), b3 = cadd(a[3], a[7]);
    const float2 c0 = csub(a[0], a[4]), d1 = csub(a[1], a[5]), d2 = csub(a[2], a[6]), d3 = csub(a[3], a[7]);
    const float r2 = 0.70710678118654752f;
    const float2 s1 = make_float2(__fadd_rn(d1.x, d1.y), __fsub_rn(d1.y, d1.x));    // c1 = d1 (1 - i)/sqrt2 = s1 r2
    const float2 c2 = make_float2(d2.y, -d2.x);                                    // d2 (-i)
    const float2 c3 = make_float2(__fmul_rn(__fsub_rn(d3.y, d3.x), r2), -__fmul_rn(__fadd_rn(d3.x, d3.y), r2));   // d3 (-1 - i)/sqrt2
    // DFT4(z0..z3): y0 = (z0+z2)+(z1+z3), y2 = (z0+z2)-(z1+z3), y1 = (z0-z2) - i (z1-z3), y3 = (z0-z2) + i (z1-z3)
    const float2 e0 = cadd(b0, b2), e1 = cadd(b1, b3), e2 = csub(b0, b2), e3 = csub(b1, b3);
    const float2 f0 = cadd(c0, c2), f2 = csub(c0, c2);
    const float2 f1 = make_float2(__fmaf_rn(s1.x, r2, c3.x), __fmaf_rn(s1.y, r2, c3.y));       // c1 + c3
    const float2 f3 = make_float2(__fmaf_rn(s1.x, r2, -c3.x), __fmaf_rn(s1.y, r2, -c3.y));     // c1 - c3
    y[0] = cadd(e0, e1);
    y[4] = csub(e0, e1);
    y[2] = make_float2(__fadd_rn(e2.x, e3.y), __fsub_rn(e2.y, e3.x));
    y[6] = make_float2(__fsub_rn(e2.x, e3.y), __fadd_rn(e2.y, e3.x));
    y[1] = cadd(f0, f1);
    y[5] = csub(f0, f1);
    y[3] = make_float2(__fadd_rn(f2.x, f3.y), __fsub_rn(f2.y, f3.x));
    y[7] = make_float2(__fsub_rn(f2.x, f3.y), __fadd_rn(f2.y, f3.x));
}
// 4-point DFT of (a, b, c, d), natural order
__device__ __forceinline__ void dft4(float2 a, float2 b, float2 c, float2 d, float2 (&y)[4]) {
    const float2 apc = cadd(a, c), amc = csub(a, c), bpd = cadd(b, d), bmd = csub(b, d);
    y[0] = cadd(apc, bpd);
    y[1] = make_float2(__fadd_rn(amc.x, bmd.y), __fsub_rn(amc.y, bmd.x));
    y[2] = csub(apc, bpd);
    y[3] = make_float2(__fsub_rn(amc.x, bmd.y), __fadd_rn(amc.y, bmd.x));
}
// exp(-2 pi i k / n_fft) for k < n_fft from the half table tw[0, nc)
__device__ __forceinline__ float2 tw_at(const float2* tw, int k, int nc) {
    if (k < nc) return tw[k];
    const float2 w = tw[k - nc];
    return make_float2(-w.x, -w.y);
}
// Bins k and nc - k of the real transform from the packed transform Z: Xe = (Z[k] + conj Z[nc-k]) / 2,
// Xo = -i (Z[k] - conj Z[nc-k]) / 2, X[k] = Xe + tw[k] Xo, X[nc-k] = conj(Xe - tw[k] Xo); then log(1e-8 + |X| scale).
// log(1e-8 + mag scale) and |z| on the special-function unit: lg2.approx / sqrt.approx (relative error ~2^-22, two orders below
// the parity tolerance on |S|); the argument of the logarithm is >= 1e-8, so no subnormal handling is needed (.ftz)
__device__ __forceinline__ float logmag1(float mag, float scale) {
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(__fmaf_rn(mag, scale, 1.0e-8f)));
    return __fmul_rn(l, 0.69314718055994531f);
}
__device__ __forceinline__ float cabs_fast(float2 z) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(__fmaf_rn(z.x, z.x, __fmul_rn(z.y, z.y))));
    return r;
}
__device__ __forceinline__ void logmag_pair(float2 zk, float2 zc, float2 w, float scale, float& lk, float& lc) {
    const float2 e = make_float2(__fmul_rn(0.5f, __fadd_rn(zk.x, zc.x)), __fmul_rn(0.5f, __fsub_rn(zk.y, zc.y)));
    const float2 od = make_float2(__fmul_rn(0.5f, __fadd_rn(zk.y, zc.y)), __fmul_rn(-0.5f, __fsub_rn(zk.x, zc.x)));
    const float2 r = cmul(w, od);
    const float2 p = cadd(e, r), m = csub(e, r);
    lk = logmag1(cabs_fast(p), scale);
    lc = logmag1(cabs_fast(m), scale);
}
// packed, windowed sample pair i of frame t (centre = True: the frame starts at t*hop - n_fft/2; reflect padding by mirroring)
__device__ __forceinline__ float2 frame_pair(const float* __restrict__ x, int L, int start, int i, float2 w) {
    int j0 = start + 2 * i, j1 = j0 + 1;
    j0 = j0 < 0 ? -j0 : (j0 >= L ? 2 * (L - 1) - j0 : j0);
    j1 = j1 < 0 ? -j1 : (j1 >= L ? 2 * (L - 1) - j1 : j1);
    return cwin(w, make_float2(__ldg(x + j0), __ldg(x + j1)));
}

// One STFT frame by `nthr` cooperating threads (all threads of the block must call: block-wide barriers inside).
// x: the clip (L samples); frame t is centred on sample t*hop (reflect padding by index mirroring); tw / win: shared
// twiddle and window tables; bufa / bufb: this thread group's ping-pong buffers (n_fft/2 complex each);
// o: nf_out log-magnitudes (shared or global).  `active` == false runs the barriers only (ragged frame groups).
// Generic in n_fft and nthr (the fused front end and the sizes without a specialised kernel).
__device__ __forceinline__ void stft_frame(const float* __restrict__ x, int L, int n_fft, int hop, int t,
                                           const float2* tw, const float* win, float2* bufa, float2* bufb,
                                           float scale, int nf_out, float* o, int tid, int nthr, bool active) {
    const int nc = n_fft >> 1;
    const int tw_shift_base = 31 - __clz(n_fft);   // log2(n_fft)
    // buffer index padding: one extra slot per 16, so that the power-of-two strides of the Stockham stores spread over the
    // banks (ncu on the unpadded kernel: 17 M store bank conflicts, l1tex 79 % of peak); buffers hold stft_buf_len(nc) slots
    auto PD = [](int i) { return i + (i >> 4); };
    // ---- load + window, pack even/odd samples into one complex sequence
    if (active) {
        const int start = t * hop - nc;            // centre=True: frame t starts at t*hop - n_fft/2
        const float2* wp = reinterpret_cast<const float2*>(win);
        if (start >= 0 && start + n_fft <= L && ((reinterpret_cast<size_t>(x + start) & 7) == 0)) {
            // interior frame: no mirroring, 8-byte loads of sample pairs
            const float2* xp = reinterpret_cast<const float2*>(x + start);
            for (int i = tid; i < nc; i += nthr) {
                const float2 v = __ldg(xp + i), w = wp[i];
                bufa[PD(i)] = cwin(w, v);
            }
        } else {
            for (int i = tid; i < nc; i += nthr) bufa[PD(i)] = frame_pair(x, L, start, i, wp[i]);
        }
    }
    __syncthreads();

    // ---- Stockham autosort, decimation in frequency: radix-8 stages in registers, then one radix-4 / radix-2 stage for
    // what is left of log2(nc).  Stage of radix R on sub-length n (n1 = n / R, stride S = 2^s_log): butterfly (p, q) reads
    // src[q + (p + r n1) S], r < R, and writes dst[q + (R p + r) S] = w_p^r Y_r, w_p = exp(-2 pi i p / n).
    float2* src = bufa;
    float2* dst = bufb;
    int n = nc, s_log = 0;
    while (n >= 8) {
        const int n1 = n >> 3;
        const int tsh = tw_shift_base - (31 - __clz(n));     // log2(n_fft / n)
        if (active) {
            for (int i = tid; i < (nc >> 3); i += nthr) {
                const int p = i >> s_log, q = i & ((1 << s_log) - 1);
                const int ib = q + (p << s_log);
                const int st = n1 << s_log;
                float2 a[8], y[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) a[r] = src[PD(ib + r * st)];
                dft8(a, y);
                const int ob = q + ((8 * p) << s_log);
                const int so = 1 << s_log;
                dst[PD(ob)] = y[0];
                if (n1 == 1) {            // last radix-8 stage of a power of 8: p = 0, every twiddle is 1
#pragma unroll
                    for (int r = 1; r < 8; ++r) dst[PD(ob + r * so)] = y[r];
                } else {
#pragma unroll
                    for (int r = 1; r < 8; ++r) dst[PD(ob + r * so)] = cmul(tw_at(tw, (r * p) << tsh, nc), y[r]);
                }
            }
        }
        __syncthreads();
        float2* tmp = src; src = dst; dst = tmp;
        n >>= 3;
        s_log += 3;
    }
    if (n == 4) {   // p = 0 only, twiddles = 1
        if (active) {
            const int st = 1 << s_log;               // = nc / 4
            for (int q = tid; q < st; q += nthr) {
                float2 y[4];
                dft4(src[PD(q)], src[PD(q + st)], src[PD(q + 2 * st)], src[PD(q + 3 * st)], y);
                dst[PD(q)] = y[0]; dst[PD(q + st)] = y[1]; dst[PD(q + 2 * st)] = y[2]; dst[PD(q + 3 * st)] = y[3];
            }
        }
        __syncthreads();
        float2* tmp = src; src = dst; dst = tmp;
    } else if (n == 2) {   // p = 0 only, twiddle = 1
        if (active) {
            for (int q = tid; q < (nc >> 1); q += nthr) {
                const float2 a = src[PD(q)];
                const float2 b = src[PD(q + (nc >> 1))];
                dst[PD(q)] = cadd(a, b);
                dst[PD(q + (nc >> 1))] = csub(a, b);
            }
        }
        __syncthreads();
        float2* tmp = src; src = dst; dst = tmp;
    }

    // ---- split the packed spectrum, magnitude, log: one pass gives bins k and nc - k
    if (active) {
        for (int k = tid; k <= (nc >> 1); k += nthr) {
            if (k == 0) {
                const float2 z0 = src[0];      // PD(0) = 0
                o[0] = logmag1(fabsf(__fadd_rn(z0.x, z0.y)), scale);
                if (nc < nf_out) o[nc] = logmag1(fabsf(__fsub_rn(z0.x, z0.y)), scale);
                continue;
            }
            float lk, lc;
            logmag_pair(src[PD(k)], src[PD(nc - k)], tw[k], scale, lk, lc);
            o[k] = lk;
            if (k != nc - k) o[nc - k] = lc;
        }
    }
    __syncthreads();   // bufa/bufb reused by the next frame
}

// ---- size-specialised frames: NC = n_fft / 2 complex points by NT = NC / 8 threads (one radix-8 butterfly per thread
// and stage, every index a compile-time function of the thread id).  The first butterfly takes its inputs straight from
// global memory (the windowed samples never pass through shared memory), the window pairs and first-stage twiddles of a
// thread are the same for every frame and live in registers, and the ping-pong buffers alternate so that one named
// barrier per stage is the only synchronisation.  Same arithmetic as stft_frame, expression for expression.
template <int NC>
struct StftT {
    static constexpr int NT = NC / 8;
    static constexpr int PST = NT + NT / 16;                  // padded distance of a butterfly's inputs
    static constexpr int PLEN = NC + NC / 16;                 // padded buffer length
    static constexpr int LOGNC = (NC == 256) ? 8 : (NC == 512) ? 9 : (NC == 1024) ? 10 : 11;
    static constexpr int N8 = LOGNC / 3;                      // radix-8 stages
    static constexpr int TAIL = 1 << (LOGNC - 3 * N8);        // 1 (none), 2 or 4
    static constexpr int NSTAGE = N8 + (TAIL > 1 ? 1 : 0);

    __device__ static __forceinline__ int PD(int i) { return i + (i >> 4); }
    __device__ static __forceinline__ void gsync(int bar) { asm volatile("bar.sync %0, %1;" :: "r"(bar), "n"(NT) : "memory"); }

    // twiddles of the radix-8 stages S >= 1 as one table per stage, [p][r - 1] = exp(-2 pi i r p / (NC >> 3 S)): 56 B per p, so the
    // few distinct p of a warp fall into distinct banks (in the master table they are a multiple of 128 B apart)
    __host__ __device__ static constexpr int stage_tab_off(int S) { return S <= 1 ? 0 : stage_tab_off(S - 1) + 7 * ((NC >> (3 * (S - 1))) / 8); }
    static constexpr int STAGE_TAB = stage_tab_off(N8);          // float2 entries (stages whose N1 == 1 included: 7 entries, unused)

    template <int S>
    __device__ static __forceinline__ void r8_write(const float2 (&a)[8], const float2* stab, const float2 (&w0)[7],
                                                    float2* dst, int ltid) {
        constexpr int S_LOG = 3 * S, N = NC >> S_LOG, N1 = N / 8, SO = 1 << S_LOG;
        float2 y[8];
        dft8(a, y);
        const int p = ltid >> S_LOG, q = ltid & (SO - 1);
        const int ob = q + ((8 * p) << S_LOG);
        // the last buffer of a frame is written without padding: its reader walks k upwards and nc - k downwards, and a run
        // of 16 consecutive slots is conflict-free only in an unpadded buffer (the Stockham strides that need the padding
        // are those of the earlier stages)
        constexpr bool LAST = (S == NSTAGE - 1);
        float2* d = dst + (LAST ? ob : PD(ob));
        d[0] = y[0];
#pragma unroll
        for (int r = 1; r < 8; ++r) {
            // PD(ob + r SO) - PD(ob): SO >= 16 is a multiple of the padding period; SO = 8: ob = q + 64 p, q < 8; SO = 1: ob = 8 p
            const int off = LAST ? r * SO : SO >= 16 ? r * (SO + SO / 16) : (SO == 8 ? 8 * r + (r >> 1) : r);
            if (N1 == 1) d[off] = y[r];
            else d[off] = cmul(S == 0 ? w0[r - 1] : stab[stage_tab_off(S) + 7 * p + r - 1], y[r]);
        }
    }

    // wv: this thread's window pairs (ltid + r NT), w0: its first-stage twiddles exp(-2 pi i r ltid / NC); cur: buffer the first
    // stage writes (flipped by the caller after every frame when NSTAGE is odd)
    __device__ static __forceinline__ void frame(const float* __restrict__ x, int L, int hop, int t, const float2* tw,
                                                 const float2* stab, const float2 (&wv)[8], const float2 (&w0)[7], float2* buf0, float2* buf1,
                                                 float scale, int nf_out, float* __restrict__ o, int ltid, int bar) {
        constexpr int n_fft = 2 * NC;
        float2 a[8];
        const int start = t * hop - NC;
        if (start >= 0 && start + n_fft <= L && ((reinterpret_cast<size_t>(x + start) & 7) == 0)) {
            const float2* xp = reinterpret_cast<const float2*>(x + start) + ltid;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float2 v = __ldg(xp + r * NT);
                a[r] = cwin(wv[r], v);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) a[r] = frame_pair(x, L, start, ltid + r * NT, wv[r]);
        }
        float2* src = buf1;
        float2* dst = buf0;
        r8_write<0>(a, stab, w0, dst, ltid);
        gsync(bar);
#pragma unroll
        for (int S = 1; S < 4; ++S) {
            if (S >= N8) continue;
            { float2* tmp = src; src = dst; dst = tmp; }
            const float2* sp = src + PD(ltid);
#pragma unroll
            for (int r = 0; r < 8; ++r) a[r] = sp[r * PST];
            if (S == 1) r8_write<1>(a, stab, w0, dst, ltid);
            else if (S == 2) r8_write<2>(a, stab, w0, dst, ltid);
            else r8_write<3>(a, stab, w0, dst, ltid);
            gsync(bar);
        }
        if (TAIL == 4) {
            { float2* tmp = src; src = dst; dst = tmp; }
            constexpr int ST = NC / 4, PSTT = ST + ST / 16;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int pq = PD(ltid + m * NT);
                float2 y[4];
                dft4(src[pq], src[pq + PSTT], src[pq + 2 * PSTT], src[pq + 3 * PSTT], y);
                float2* d = dst + ltid + m * NT;           // (last buffer: unpadded, see r8_write)
                d[0] = y[0]; d[ST] = y[1]; d[2 * ST] = y[2]; d[3 * ST] = y[3];
            }
            gsync(bar);
        } else if (TAIL == 2) {
            { float2* tmp = src; src = dst; dst = tmp; }
            constexpr int ST = NC / 2, PSTT = ST + ST / 16;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int pq = PD(ltid + m * NT);
                const float2 u = src[pq], v = src[pq + PSTT];
                dst[ltid + m * NT] = cadd(u, v);           // (last buffer: unpadded, see r8_write)
                dst[ltid + m * NT + ST] = csub(u, v);
            }
            gsync(bar);
        }
        // the packed spectrum is in dst: bins k = ltid + m NT, m < 4, and nc - k; k = NC / 2 by thread 0
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int k = ltid + m * NT;
            if (m == 0 && ltid == 0) {
                const float2 z0 = dst[0];
                o[0] = logmag1(fabsf(__fadd_rn(z0.x, z0.y)), scale);
                if (NC < nf_out) o[NC] = logmag1(fabsf(__fsub_rn(z0.x, z0.y)), scale);
                float lk, lc;
                logmag_pair(dst[NC / 2], dst[NC / 2], tw[NC / 2], scale, lk, lc);
                o[NC / 2] = lk;
                continue;
            }
            float lk, lc;
            logmag_pair(dst[k], dst[NC - k], tw[k], scale, lk, lc);
            o[k] = lk;
            o[NC - k] = lc;
        }
    }
};

constexpr int STFT_T_THREADS = 256;

template <int NC>
__global__ void __launch_bounds__(STFT_T_THREADS, NC <= 512 ? 4 : 3)
stft_logmag_t_kernel(const float* __restrict__ audio, int L, int hop, const float* __restrict__ window,
                     const float2* __restrict__ twiddle, float scale, int nf_out, int nt_out, int frames_per_group,
                     float* __restrict__ out) {
    using T = StftT<NC>;
    constexpr int G = STFT_T_THREADS / T::NT;          // frames in flight per block
    extern __shared__ float2 smem_f2[];
    float2* tw = smem_f2;                              // NC entries: exp(-2 pi i k / n_fft)
    float2* stab = tw + NC;                            // per-stage twiddle tables (StftT::stage_tab_off)
    float2* bufs = stab + T::STAGE_TAB;                // G x 2 x PLEN
    const int tid = threadIdx.x, g = tid / T::NT, ltid = tid % T::NT;
    const int clip = blockIdx.y;
    const float* x = audio + (size_t)clip * L;

    for (int i = tid; i < NC; i += STFT_T_THREADS) tw[i] = twiddle[i];
#pragma unroll
    for (int S = 1; S < T::N8; ++S) {
        const int n1 = (NC >> (3 * S)) / 8, tsh = 1 + 3 * S;
        for (int i = tid; i < 7 * n1; i += STFT_T_THREADS) {
            const int pp = i / 7, r = i - 7 * pp + 1, k = (r * pp) << tsh;
            const float2 w = __ldg(twiddle + (k < NC ? k : k - NC));
            stab[T::stage_tab_off(S) + i] = k < NC ? w : make_float2(-w.x, -w.y);
        }
    }
    float2 wv[8], w0[7];
#pragma unroll
    for (int r = 0; r < 8; ++r) wv[r] = __ldg(reinterpret_cast<const float2*>(window) + ltid + r * T::NT);
#pragma unroll
    for (int r = 1; r < 8; ++r) {
        const int k = (r * ltid) << 1;
        const float2 w = __ldg(twiddle + (k < NC ? k : k - NC));
        w0[r - 1] = k < NC ? w : make_float2(-w.x, -w.y);
    }
    __syncthreads();

    float2* b0 = bufs + (size_t)(2 * g) * T::PLEN;
    float2* b1 = b0 + T::PLEN;
    const int t0 = blockIdx.x * G * frames_per_group;
    for (int i = 0; i < frames_per_group; ++i) {
        const int t = t0 + i * G + g;
        if (t >= nt_out) break;                        // (whole groups leave together: the barriers are per group)
        T::frame(x, L, hop, t, tw, stab, wv, w0, b0, b1, scale, nf_out, out + ((size_t)clip * nt_out + t) * nf_out, ltid, 1 + g);
        if (T::NSTAGE & 1) { float2* tmp = b0; b0 = b1; b1 = tmp; }
    }
}

__global__ void stft_logmag_kernel(const float* __restrict__ audio, int L, int n_fft, int hop,
                                   const float* __restrict__ window,
                                   const float2* __restrict__ twiddle, float scale, int nf_out,
                                   int nt_out, int frames_per_block, float* __restrict__ out) {
    extern __shared__ float2 smem_f2[];
    const int nc = n_fft >> 1;             // complex FFT length
    float2* tw = smem_f2;                  // nc entries: exp(-2 pi i k / n_fft)
    float2* bufa = tw + nc;
    float2* bufb = bufa + stft_buf_len(nc);
    float* win = reinterpret_cast<float*>(bufb + stft_buf_len(nc));   // n_fft entries

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int clip = blockIdx.y;
    const float* x = audio + (size_t)clip * L;

    for (int i = tid; i < nc; i += nthr) tw[i] = twiddle[i];
    for (int i = tid; i < n_fft; i += nthr) win[i] = window[i];
    __syncthreads();

    const int t0 = blockIdx.x * frames_per_block;
    const int t1 = min(nt_out, t0 + frames_per_block);
    for (int t = t0; t < t1; ++t)
        stft_frame(x, L, n_fft, hop, t, tw, win, bufa, bufb, scale, nf_out, out + ((size_t)clip * nt_out + t) * nf_out, tid, nthr, true);
}

// ------------------------------------------------------------------------------------ clouds
__global__ void build_clouds_kernel(const float* __restrict__ logmag, long long total, int nf, int nt,
                                    const float* __restrict__ farr, const float* __restrict__ tarr,
                                    float* __restrict__ pts) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int npts = nf * nt;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int p = (int)(i % npts);
        const int f = p % nf, t = p / nf;
        const float v = logmag[i];
        if (tarr != nullptr) {
            float* o = pts + i * 3;
            o[0] = __ldg(farr + f); o[1] = __ldg(tarr + t); o[2] = v;
        } else {
            *reinterpret_cast<float2*>(pts + i * 2) = make_float2(__ldg(farr + f), v);
        }
    }
}

// ------------------------------------------------------------------------------------ top-K
__device__ __forceinline__ uint32_t ordered_key(float x) {
    uint32_t u = __float_as_uint(x);
    if (u == 0x80000000u) u = 0u;                       // -0.0 ties with +0.0, as in (-x).argsort()
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // larger float -> larger uint
}

constexpr int TOPK_THREADS = 512;
constexpr int TOPK_CAND = 1024;        // shared candidate list of the register-key kernel's later radix passes

// Selection of one cloud by the whole block (TOPK_THREADS threads).  keys: the cloud's N keys, in global memory
// (SMEM_KEYS == false, read through the read-only path) or already in shared memory (fused front end).
template <bool SMEM_KEYS>
__device__ __forceinline__ void topk_core(const float* keys, const int cloud, int N, int nf, const float* __restrict__ farr,
                                          const float* __restrict__ tarr, int K, int kpad, int sorted, int use_tau, float tau,
                                          float* __restrict__ pts_all, int32_t* __restrict__ idx_all,
                                          int32_t* __restrict__ counts, unsigned long long* sortbuf) {
    __shared__ int hist[256];
    __shared__ int warp_gt[TOPK_THREADS / 32], warp_eq[TOPK_THREADS / 32];
    __shared__ uint32_t s_prefix;
    __shared__ int s_remaining;
    __shared__ int s_gt_base, s_eq_base;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    auto ldkey = [&](int i) { return SMEM_KEYS ? keys[i] : __ldg(keys + i); };
    const int width = tarr != nullptr ? 3 : 2;
    float* pts = pts_all ? pts_all + (size_t)cloud * K * width : nullptr;
    int32_t* idx_out = idx_all ? idx_all + (size_t)cloud * K : nullptr;

    // ---- K-th largest key (skipped when every point is kept)
    uint32_t kth = 0;
    int n_ties_take = 0;
    const bool all = (K >= N);
    if (!all && SMEM_KEYS) {
        // Keys in shared memory: bitwise search for the largest v with #{key >= v} >= K, two bits per round (three
        // candidate thresholds counted in one sweep over the keys; plain compares, no histogram, no atomics).
        __shared__ int s_cnt[2][3][TOPK_THREADS / 32];
        uint32_t prefix = 0;
        int cnt_ge_prefix = N;                       // #{key >= prefix}
        for (int round = 0; round < 16; ++round) {
            const int bit = 30 - 2 * round;
            const uint32_t c1 = prefix | (1u << bit), c2 = prefix | (2u << bit), c3 = prefix | (3u << bit);
            int n1 = 0, n2 = 0, n3 = 0;
            for (int i = tid; i < N; i += TOPK_THREADS) {
                const uint32_t o = ordered_key(keys[i]);
                n1 += o >= c1; n2 += o >= c2; n3 += o >= c3;
            }
            n1 = __reduce_add_sync(0xffffffffu, n1);
            n2 = __reduce_add_sync(0xffffffffu, n2);
            n3 = __reduce_add_sync(0xffffffffu, n3);
            int (*sc)[TOPK_THREADS / 32] = s_cnt[round & 1];
            if (lane == 0) { sc[0][wid] = n1; sc[1][wid] = n2; sc[2][wid] = n3; }
            __syncthreads();
            int t1 = 0, t2 = 0, t3 = 0;
#pragma unroll
            for (int w = 0; w < TOPK_THREADS / 32; ++w) { t1 += sc[0][w]; t2 += sc[1][w]; t3 += sc[2][w]; }
            if (t3 >= K) { prefix = c3; cnt_ge_prefix = t3; }
            else if (t2 >= K) { prefix = c2; cnt_ge_prefix = t2; }
            else if (t1 >= K) { prefix = c1; cnt_ge_prefix = t1; }
            // (the double-buffered counters make one barrier per round sufficient)
        }
        kth = prefix;
        // ties: #{key > kth} by one more sweep
        int ngt = 0;
        for (int i = tid; i < N; i += TOPK_THREADS) ngt += ordered_key(keys[i]) > kth;
        ngt = __reduce_add_sync(0xffffffffu, ngt);
        __syncthreads();
        if (lane == 0) s_cnt[0][0][wid] = ngt;
        __syncthreads();
        int tgt = 0;
#pragma unroll
        for (int w = 0; w < TOPK_THREADS / 32; ++w) tgt += s_cnt[0][0][w];
        n_ties_take = K - tgt;
        (void)cnt_ge_prefix;
        __syncthreads();
    } else if (!all) {
        if (tid == 0) { s_prefix = 0; s_remaining = K; }
        uint32_t mask = 0;
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += TOPK_THREADS) hist[i] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            for (int base = 0; base < N; base += TOPK_THREADS) {
                const int i = base + tid;
                const bool act = i < N;
                uint32_t o = act ? ordered_key(ldkey(i)) : 0u;
                const bool cand = act && ((o & mask) == prefix);
                const uint32_t digit = (o >> shift) & 255u;
                // warp-aggregated shared atomics: log-magnitudes share their top bytes
                const uint32_t cmask = __ballot_sync(0xffffffffu, cand);
                if (cand) {
                    const uint32_t peers = __match_any_sync(cmask, digit);
                    if ((peers & ((1u << lane) - 1)) == 0) atomicAdd(&hist[digit], __popc(peers));
                }
            }
            __syncthreads();
            if (wid == 0) {
                // lane l owns digits [8l, 8l+8); find the digit holding the `remaining`-th largest
                const int remaining = s_remaining;
                int c[8], tot = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; tot += c[j]; }
                int suf = tot;        // inclusive suffix sum over lanes >= lane
#pragma unroll
                for (int o2 = 1; o2 < 32; o2 <<= 1) {
                    const int v = __shfl_down_sync(0xffffffffu, suf, o2);
                    if (lane + o2 < 32) suf += v;
                }
                const int above = suf - tot;   // candidates with a larger digit than this lane's
                if (above < remaining && remaining <= above + tot) {
                    int cum = above;
#pragma unroll
                    for (int j = 7; j >= 0; --j) {
                        if (cum < remaining && remaining <= cum + c[j]) {
                            s_prefix = prefix | ((uint32_t)(lane * 8 + j) << shift);
                            s_remaining = remaining - cum;
                        }
                        cum += c[j];
                    }
                }
            }
            mask |= 255u << shift;
            __syncthreads();
        }
        kth = s_prefix;
        n_ties_take = s_remaining;
    }
    // Threshold mode (extension): keep key >= tau, capped at K by the rule above.  When tau lies above the
    // K-th largest key the threshold alone decides ("greater" class = key >= tau, no tie class); otherwise the cap does.
    const uint32_t thr = use_tau ? ordered_key(tau) : 0u;
    const bool tau_rules = use_tau && (all || thr > kth);
    if (tau_rules) n_ties_take = 0;
    if (tid == 0) { s_gt_base = 0; s_eq_base = 0; }
    if (sorted) {
        for (int i = tid; i < kpad; i += TOPK_THREADS) sortbuf[i] = ~0ull;
    }
    __syncthreads();

    // ---- ordered compaction: position = (#selected before me) in flat-index order
    for (int base = 0; base < N; base += TOPK_THREADS) {
        const int i = base + tid;
        const bool act = i < N;
        float kv = 0.f;
        uint32_t o = 0;
        if (act) { kv = ldkey(i); o = ordered_key(kv); }
        const bool gt = act && (tau_rules ? (o >= thr) : (all || o > kth));
        const bool eq = act && !all && !tau_rules && (o == kth);
        const uint32_t bg = __ballot_sync(0xffffffffu, gt);
        const uint32_t be = __ballot_sync(0xffffffffu, eq);
        const uint32_t lt = (1u << lane) - 1;
        if (lane == 0) { warp_gt[wid] = __popc(bg); warp_eq[wid] = __popc(be); }
        __syncthreads();
        int gt_before = s_gt_base, eq_before = s_eq_base;
        for (int w = 0; w < wid; ++w) { gt_before += warp_gt[w]; eq_before += warp_eq[w]; }
        gt_before += __popc(bg & lt);
        eq_before += __popc(be & lt);
        const bool sel = gt || (eq && eq_before < n_ties_take);
        if (sel) {
            const int pos = gt_before + min(eq_before, n_ties_take);
            if (sorted) {
                sortbuf[pos] = ((unsigned long long)(~o) << 32) | (uint32_t)i;
            } else {
                const int f = i % nf, t = i / nf;
                if (pts) {
                    if (width == 3) { pts[pos * 3] = __ldg(farr + f); pts[pos * 3 + 1] = __ldg(tarr + t); pts[pos * 3 + 2] = kv; }
                    else { pts[pos * 2] = __ldg(farr + f); pts[pos * 2 + 1] = kv; }
                }
                if (idx_out) idx_out[pos] = i;
            }
        }
        __syncthreads();
        if (tid == 0) {
            int g = 0, e = 0;
            for (int w = 0; w < TOPK_THREADS / 32; ++w) { g += warp_gt[w]; e += warp_eq[w]; }
            s_gt_base += g; s_eq_base += e;
        }
        __syncthreads();
    }
    // rows past the number of kept points are padding: zeros, index -1
    const int n_kept = s_gt_base + min(s_eq_base, n_ties_take);
    if (counts != nullptr && tid == 0) counts[cloud] = n_kept;
    if (!sorted) {
        for (int r = n_kept + tid; r < K; r += TOPK_THREADS) {
            if (pts) { for (int j = 0; j < width; ++j) pts[r * width + j] = 0.f; }
            if (idx_out) idx_out[r] = -1;
        }
        return;
    }

    // ---- bitonic sort of the survivors: ascending (~key, index) == descending key, stable
    for (int k2 = 2; k2 <= kpad; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (kpad >> 1); t += TOPK_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int ixj = i | j;
                const unsigned long long a = sortbuf[i], b = sortbuf[ixj];
                const bool asc = (i & k2) == 0;
                if ((a > b) == asc) { sortbuf[i] = b; sortbuf[ixj] = a; }
            }
            __syncthreads();
        }
    }
    for (int r = tid; r < K; r += TOPK_THREADS) {
        if (r >= n_kept) {
            if (pts) { for (int j = 0; j < width; ++j) pts[r * width + j] = 0.f; }
            if (idx_out) idx_out[r] = -1;
            continue;
        }
        const int i = (int)(uint32_t)sortbuf[r];
        const float kv = ldkey(i);
        const int f = i % nf, t = i / nf;
        if (pts) {
            if (width == 3) { pts[r * 3] = __ldg(farr + f); pts[r * 3 + 1] = __ldg(tarr + t); pts[r * 3 + 2] = kv; }
            else { pts[r * 2] = __ldg(farr + f); pts[r * 2 + 1] = kv; }
        }
        if (idx_out) idx_out[r] = i;
    }
}


__global__ void __launch_bounds__(TOPK_THREADS)
topk_kernel(const float* __restrict__ keys_all, int N, int nf, const float* __restrict__ farr,
            const float* __restrict__ tarr, int K, int kpad, int sorted, int use_tau, float tau,
            float* __restrict__ pts_all, int32_t* __restrict__ idx_all, int32_t* __restrict__ counts) {
    extern __shared__ unsigned long long sortbuf[];      // kpad entries when sorted
    topk_core<false>(keys_all + (size_t)blockIdx.x * N, blockIdx.x, N, nf, farr, tarr, K, kpad, sorted, use_tau, tau, pts_all,
                     idx_all, counts, sortbuf);
}

// ------------------------------------------------------------------------------------ top-K, keys in registers
// Clouds of up to 512 * KPT points: thread t owns the KPT CONSECUTIVE flat indices [t KPT, (t+1) KPT), as order-preserving
// uint keys in registers.  Same selection rule as topk_core (bit-identical results), organised so that the per-key work is
// a handful of register instructions (ncu of the previous version: 95 instructions per key, issue-bound):
//   * the select runs on u = (o - omin) << clz(omax - omin), an order- and tie-preserving map of the cloud's key range onto
//     the full 32 bits: log-magnitudes share sign and exponent, the normalised top byte follows the value distribution;
//   * ONE histogram pass over the registers (warp-private shared histograms) fixes the first digit D of the K-th key;
//     keys above D are selected outright, the few keys IN D go to a shared candidate list on which the remaining three
//     digits are resolved (register passes remain as the fallback when more than TOPK_CAND keys share D);
//   * selection flags as two 32-bit masks per thread, ONE block-wide exclusive scan of the per-thread counts gives every
//     selected point its position in flat-index order (ties at the K boundary keep the lowest indices);
//   * the K survivors are sorted on (~key, index) for the descending emission order: bitonic network with the elements in
//     registers, partner exchange by shuffles (distance < 32), shared memory (< 512) or in-thread (>= 512).
__device__ __forceinline__ uint32_t ordered_key_fast(float x) {
    const uint32_t u = __float_as_uint(x + 0.0f);       // -0.0 + 0.0 = +0.0 (ties with +0.0); every other value unchanged
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}

constexpr int TOPK_SORT_E = 4;         // elements per thread of the register sort: K <= 2048

// Bitonic sort of kp = NE * 512 (or, for NE == 1, kp <= 512) 64-bit elements held NE per thread: element x = tid + 512 e
// lives in v[e].  Partner exchange by warp shuffles (distance < 32), through shared memory (< 512; barrier 1 counts the kp
// participating threads -- for kp < 512 the other warps have left) or inside the thread (>= 512).  Emits element r < K
// through emit(r, low 32 bits).
template <int NE, class Emit>
__device__ __forceinline__ void sort_regs(unsigned long long* sortbuf, unsigned long long* sortbuf2, int kpad, int kp, int K,
                                          int tid, Emit emit) {
    const int nthr = min(kp, TOPK_THREADS);
    int xsel = 0;
    uint32_t hi[NE], lo[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int x = tid + TOPK_THREADS * e;
        const unsigned long long v = (x < kpad) ? sortbuf[x] : ~0ull;
        hi[e] = (uint32_t)(v >> 32); lo[e] = (uint32_t)v;
    }
    // keep the smaller (keep_min) or the larger of (hi, lo)[e] and the partner's (ph, pl)
    auto cx = [&](int e, uint32_t ph, uint32_t pl, bool keep_min) {
        const bool p_less = (ph < hi[e]) || (ph == hi[e] && pl < lo[e]);
        const bool take = (p_less == keep_min);
        hi[e] = take ? ph : hi[e];
        lo[e] = take ? pl : lo[e];
    };
    // compare-exchange with the lane at distance jj inside the merge of width k2 (both compile-time in the callers)
    auto warp_stage = [&](int k2, int jj) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const int x = tid + TOPK_THREADS * e;
            const uint32_t ph = __shfl_xor_sync(0xffffffffu, hi[e], jj), pl = __shfl_xor_sync(0xffffffffu, lo[e], jj);
            cx(e, ph, pl, ((x & k2) == 0) == ((x & jj) == 0));
        }
    };
    // merges of width <= 32: all inside a warp
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
        for (int jj = k2 >> 1; jj > 0; jj >>= 1) warp_stage(k2, jj);
    }
    for (int k2 = 64; k2 <= kp; k2 <<= 1) {
        int j = k2 >> 1;
        if (NE > 1) {
            for (; j >= TOPK_THREADS; j >>= 1) {                   // partner in the same thread: e ^ (j / 512)
                const int je = j / TOPK_THREADS;
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    if (e & je) continue;
                    const bool asc = (((tid + TOPK_THREADS * e) & k2) == 0);
#pragma unroll
                    for (int f = 0; f < NE; ++f) {
                        if (f != (e | je) || f == e) continue;     // (static after unrolling: je is 1 or 2)
                        const uint32_t ah = hi[e], al = lo[e];
                        cx(e, hi[f], lo[f], asc);
                        cx(f, ah, al, !asc);
                    }
                }
            }
        }
        for (; j >= 32; j >>= 1) {                                 // partner in another warp
            // two exchange buffers in turn: the barrier of the next exchange orders this one's reads before the writes of
            // the one after it, so one barrier per exchange suffices
            unsigned long long* xb = xsel ? sortbuf2 : sortbuf;
            xsel ^= 1;
#pragma unroll
            for (int e = 0; e < NE; ++e)
                xb[tid + TOPK_THREADS * e] = ((unsigned long long)hi[e] << 32) | lo[e];
            asm volatile("bar.sync 1, %0;" :: "r"(nthr) : "memory");
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                const int x = tid + TOPK_THREADS * e;
                const unsigned long long p = xb[(tid ^ j) + TOPK_THREADS * e];
                cx(e, (uint32_t)(p >> 32), (uint32_t)p, ((x & k2) == 0) == ((x & j) == 0));
            }
        }
#pragma unroll
        for (int jj = 16; jj > 0; jj >>= 1) warp_stage(k2, jj);    // partner in the same warp
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int r = tid + TOPK_THREADS * e;
        if (r < K) emit(r, (int)lo[e]);
    }
}

template <int KPT>
__global__ void __launch_bounds__(TOPK_THREADS, 2)
topk_reg_kernel(const float* __restrict__ keys_all, int N, int nf, const float* __restrict__ farr,
                const float* __restrict__ tarr, int K, int kpad, int sorted, int use_tau, float tau,
                float* __restrict__ pts_all, int32_t* __restrict__ idx_all, int32_t* __restrict__ counts) {
    extern __shared__ unsigned long long sortbuf[];      // kpad entries when sorted
    constexpr int NW = TOPK_THREADS / 32;
    __shared__ __align__(16) int hist[NW][256];          // (reused as the sort's second exchange buffer)
    __shared__ int tot[256];
    __shared__ int ltot[3][256];                         // digit counts of the candidate list, one table per lower digit
    __shared__ int warp_sum[NW];
    __shared__ uint32_t s_prefix, s_lo[NW], s_hi[NW];
    __shared__ int s_remaining, s_bin, s_ncand;
    __shared__ uint32_t cand[TOPK_CAND];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int cloud = blockIdx.x;
    const float* keys = keys_all + (size_t)cloud * N;
    const int width = tarr != nullptr ? 3 : 2;
    float* pts = pts_all ? pts_all + (size_t)cloud * K * width : nullptr;
    int32_t* idx_out = idx_all ? idx_all + (size_t)cloud * K : nullptr;

    const int i0 = tid * KPT;
    const int nval = max(0, min(KPT, N - i0));          // valid keys of this thread
    const uint32_t valid = nval >= 32 ? 0xffffffffu : ((1u << nval) - 1u);
    uint32_t o[KPT];                                    // slots past the cloud repeat its last key (neutral for min / max)
    if (nval == KPT && ((reinterpret_cast<size_t>(keys + i0) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < KPT; j += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(keys + i0 + j));
            o[j] = ordered_key_fast(v.x); o[j + 1] = ordered_key_fast(v.y);
            o[j + 2] = ordered_key_fast(v.z); o[j + 3] = ordered_key_fast(v.w);
        }
    } else {
#pragma unroll
        for (int j = 0; j < KPT; ++j) o[j] = ordered_key_fast(__ldg(keys + min(i0 + j, N - 1)));
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) hist[wid][lane + 32 * b] = 0;
    if (tid == 0) s_ncand = 0;
    if (tid < 256) { ltot[0][tid] = 0; ltot[1][tid] = 0; ltot[2][tid] = 0; }

    // Digit holding the `remaining`-th largest of the candidates counted in t[256], found by warp 0 (lane l owns digits
    // [8l, 8l+8)) and published through shared memory (measured: every warp evaluating it for itself costs more issue slots
    // than the barrier it saves).  Requires 1 <= remaining <= sum(t).  Block-wide call (one barrier inside); updates
    // prefix / remaining, returns the population of the chosen digit.
    auto pick_digit = [&](const int* t, uint32_t& prefix, int& remaining, int shift) {
        if (wid == 0) {
            int c[8], tl = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { c[j] = t[lane * 8 + j]; tl += c[j]; }
            int suf = tl;        // inclusive suffix sum over lanes >= lane
#pragma unroll
            for (int o2 = 1; o2 < 32; o2 <<= 1) {
                const int v = __shfl_down_sync(0xffffffffu, suf, o2);
                if (lane + o2 < 32) suf += v;
            }
            const int above = suf - tl;   // candidates with a larger digit than this lane's
            const bool mine = above < remaining && remaining <= above + tl;
            int dj = 0, rem = 0, bin = 0;
            if (mine) {
                int cum = above;
#pragma unroll
                for (int j = 7; j >= 0; --j) {
                    if (cum < remaining && remaining <= cum + c[j]) { dj = lane * 8 + j; rem = remaining - cum; bin = c[j]; }
                    cum += c[j];
                }
            }
    if (mine) { s_prefix = prefix | ((uint32_t)dj << shift); s_remaining = rem; s_bin = bin; }      // exactly one lane
        }
        __syncthreads();
        prefix = s_prefix;
        remaining = s_remaining;
        return s_bin;
    };

    // ---- K-th largest key (skipped when every point is kept)
    uint32_t kth = 0, kth_u = 0, omin = 0;
    uint32_t gt1 = 0, cm = 0;          // keys above / inside the first digit of the K-th key
    int n_ties_take = 0, lz = 32, cand_at = -1;
    const bool all = (K >= N);
    if (!all) {
        uint32_t lo = o[0], hi = o[0];
#pragma unroll
        for (int j = 1; j < KPT; ++j) { lo = min(lo, o[j]); hi = max(hi, o[j]); }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) { s_lo[wid] = lo; s_hi[wid] = hi; }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < NW; ++w) { lo = min(lo, s_lo[w]); hi = max(hi, s_hi[w]); }
        omin = lo;
        lz = (hi == lo) ? 32 : __clz(hi - lo);
        kth = omin;
        n_ties_take = K;
        cm = valid;                    // lz == 32: all keys equal, the K lowest indices are kept
        if (lz < 32) {
            // first digit: one atomic per key into the warp's histogram
            if (nval == KPT) {
#pragma unroll
                for (int j = 0; j < KPT; ++j) atomicAdd(&hist[wid][((o[j] - omin) << lz) >> 24], 1);
            } else {
#pragma unroll
                for (int j = 0; j < KPT; ++j) if (j < nval) atomicAdd(&hist[wid][((o[j] - omin) << lz) >> 24], 1);
            }
            __syncthreads();
            if (tid < 256) {
                int t = 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) t += hist[w][tid];
                tot[tid] = t;
            }
            __syncthreads();
            uint32_t prefix = 0u, mask = 0xff000000u;
            int remaining = K;
            const int bin = pick_digit(tot, prefix, remaining, 24);
            const uint32_t top = prefix | 0x00ffffffu;
            gt1 = 0; cm = 0;
#pragma unroll
            for (int j = 0; j < KPT; ++j) {
                const uint32_t u = (o[j] - omin) << lz;
                gt1 |= (uint32_t)(u > top) << j;
                cm |= (uint32_t)(u >= prefix) << j;
            }
            gt1 &= valid;
            cm &= valid & ~gt1;
            if (lz < 24) {                     // (digits below bit lz of u are all zero)
                const bool listed = bin <= TOPK_CAND;
                if (listed && cm) {
                    // this thread's candidates, in index order, at cand[cand_at ...] (keys re-read through the read-only
                    // path: a dynamic index into o[] would spill it, a static loop costs 5 instructions per key)
                    cand_at = atomicAdd(&s_ncand, __popc(cm));
                    int at = cand_at;
                    uint32_t m = cm;
                    while (m) {
                        const int j = __ffs(m) - 1;
                        m &= m - 1;
                        cand[at++] = (ordered_key_fast(__ldg(keys + i0 + j)) - omin) << lz;
                    }
                }
                for (int shift = 16; shift >= 0 && shift + 8 > lz; shift -= 8) {
                    const int* t = tot;
                    if (listed) {
                        if (shift == 16) __syncthreads();            // publishes the list
                        int* lt = ltot[shift >> 3];
                        for (int c = tid; c < bin; c += TOPK_THREADS) {
                            const uint32_t u = cand[c];
                            if ((u & mask) == prefix) atomicAdd(&lt[(u >> shift) & 255u], 1);
                        }
                        __syncthreads();
                        t = lt;
                    } else {
#pragma unroll
                        for (int b = 0; b < 8; ++b) hist[wid][lane + 32 * b] = 0;
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < KPT; ++j) {
                            const uint32_t u = (o[j] - omin) << lz;
                            if (((cm >> j) & 1u) && (u & mask) == prefix) atomicAdd(&hist[wid][(u >> shift) & 255u], 1);
                        }
                        __syncthreads();
                        if (tid < 256) {
                            int sum = 0;
#pragma unroll
                            for (int w = 0; w < NW; ++w) sum += hist[w][tid];
                            tot[tid] = sum;
                        }
                        __syncthreads();
                    }
                    pick_digit(t, prefix, remaining, shift);
                    mask |= 255u << shift;
                }
            }
            kth_u = prefix;
            kth = (prefix >> lz) + omin;
            n_ties_take = remaining;
        }
    }
    // Threshold mode: see topk_core
    const uint32_t thr = use_tau ? ordered_key_fast(tau) : 0u;
    const bool tau_rules = use_tau && (all || thr > kth);
    if (tau_rules) n_ties_take = 0;

    // ---- selection flags and ONE block-wide exclusive scan of (greater, equal) counts
    uint32_t gtm = 0, eqm = 0;
    if (all || tau_rules) {
        if (tau_rules) {
#pragma unroll
            for (int j = 0; j < KPT; ++j) gtm |= (uint32_t)(o[j] >= thr) << j;
            gtm &= valid;
        } else {
            gtm = valid;
        }
    } else if (lz >= 32) {
        eqm = valid;
    } else {
        // keys above the K-th key's first digit are in; the candidates inside it are compared in full
        gtm = gt1;
        uint32_t m = cm;
        int at = cand_at;
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t u = (at >= 0) ? cand[at++] : ((ordered_key_fast(__ldg(keys + i0 + j)) - omin) << lz);
            if (u > kth_u) gtm |= 1u << j;
            else if (u == kth_u) eqm |= 1u << j;
        }
    }
    const int mine = __popc(gtm) | (__popc(eqm) << 16);          // both counts fit 15 bits (N <= 16384)
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) warp_sum[wid] = incl;
    if (sorted) {
        for (int i = tid; i < kpad; i += TOPK_THREADS) sortbuf[i] = ~0ull;
    }
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const int v = warp_sum[w];
        if (w < wid) base += v;
        total += v;
    }
    int gt_before = (base + incl - mine) & 0xffff, eq_before = (base + incl - mine) >> 16;
    const int n_kept = (total & 0xffff) + min(total >> 16, n_ties_take);
    if (counts != nullptr && tid == 0) counts[cloud] = n_kept;

    // ---- ordered compaction of this thread's selected keys: walk the set bits only (the key is re-read through the
    // read-only path; a dynamic index into o[] would spill the array)
    uint32_t selm = gtm | eqm;
    while (selm) {
        const int j = __ffs(selm) - 1;
        selm &= selm - 1;
        const bool gt = (gtm >> j) & 1u;
        const bool take = gt || eq_before < n_ties_take;
        if (take) {
            const int pos = gt_before + min(eq_before, n_ties_take);
            const int i = i0 + j;
            const float kv = __ldg(keys + i);
            if (sorted) {
                sortbuf[pos] = ((unsigned long long)(~ordered_key_fast(kv)) << 32) | (uint32_t)i;
            } else {
                const int f = i % nf, t = i / nf;
                if (pts) {
                    if (width == 3) { pts[pos * 3] = __ldg(farr + f); pts[pos * 3 + 1] = __ldg(tarr + t); pts[pos * 3 + 2] = kv; }
                    else { pts[pos * 2] = __ldg(farr + f); pts[pos * 2 + 1] = kv; }
                }
                if (idx_out) idx_out[pos] = i;
            }
        }
        if (gt) ++gt_before; else ++eq_before;
    }
    if (!sorted) {
        for (int r = n_kept + tid; r < K; r += TOPK_THREADS) {
            if (pts) { for (int j = 0; j < width; ++j) pts[r * width + j] = 0.f; }
            if (idx_out) idx_out[r] = -1;
        }
        return;
    }
    __syncthreads();

    auto emit = [&](int r, int i) {
        if (r >= n_kept) {
            if (pts) { for (int j = 0; j < width; ++j) pts[r * width + j] = 0.f; }
            if (idx_out) idx_out[r] = -1;
            return;
        }
        const float kv = __ldg(keys + i);
        const int f = i % nf, t = i / nf;
        if (pts) {
            if (width == 3) { pts[r * 3] = __ldg(farr + f); pts[r * 3 + 1] = __ldg(tarr + t); pts[r * 3 + 2] = kv; }
            else { pts[r * 2] = __ldg(farr + f); pts[r * 2 + 1] = kv; }
        }
        if (idx_out) idx_out[r] = i;
    };

    // ---- bitonic sort of the survivors: ascending (~key, index) == descending key, stable
    if (kpad <= TOPK_THREADS * TOPK_SORT_E) {
        const int kp = max(kpad, 32);                 // sentinels ~0 pad to a whole warp
        if (kp < TOPK_THREADS && tid >= kp) return;   // whole warps without an element leave (the sort's barriers count kp)
        unsigned long long* sortbuf2 = reinterpret_cast<unsigned long long*>(&hist[0][0]);   // (free after the select: 2048 entries)
        if (kp <= TOPK_THREADS) sort_regs<1>(sortbuf, sortbuf2, kpad, kp, K, tid, emit);
        else if (kp == 2 * TOPK_THREADS) sort_regs<2>(sortbuf, sortbuf2, kpad, kp, K, tid, emit);
        else sort_regs<4>(sortbuf, sortbuf2, kpad, kp, K, tid, emit);
        return;
    }
    for (int k2 = 2; k2 <= kpad; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (kpad >> 1); t += TOPK_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int ixj = i | j;
                const unsigned long long a = sortbuf[i], b = sortbuf[ixj];
                const bool asc = (i & k2) == 0;
                if ((a > b) == asc) { sortbuf[i] = b; sortbuf[ixj] = a; }
            }
            __syncthreads();
        }
    }
    for (int r = tid; r < K; r += TOPK_THREADS) emit(r, (int)(uint32_t)sortbuf[r]);
}

// ------------------------------------------------------------------------------------ fused front end
// audio -> STFT -> log-magnitude -> (f, t, mag) cloud -> selection in ONE launch (a1..a7 of SURVEY.md 8a): one block per
// cloud (= ntemp consecutive frames of a clip).  The cloud's log-magnitudes never leave shared memory, so the HBM traffic
// is the audio read (4 L bytes per clip) plus 16 B per selected point.  Four 128-thread groups transform four frames at
// a time; the selection then runs on the shared-memory keys.
constexpr int FUSED_GROUPS = TOPK_THREADS / 128;

__global__ void __launch_bounds__(TOPK_THREADS)
fused_frontend_kernel(const float* __restrict__ audio, int L, int n_fft, int hop, const float* __restrict__ window,
                      const float2* __restrict__ twiddle, float scale, int nf, int nt_cloud, int clouds_per_clip,
                      const float* __restrict__ farr, const float* __restrict__ tarr, int K, int kpad, int sorted,
                      int use_tau, float tau, float* __restrict__ pts_all, int32_t* __restrict__ idx_all,
                      int32_t* __restrict__ counts) {
    extern __shared__ unsigned long long fused_smem[];
    const int nc = n_fft >> 1;
    unsigned long long* sortbuf = fused_smem;                               // kpad entries (8-byte aligned first)
    float2* tw = reinterpret_cast<float2*>(sortbuf + kpad);                 // nc
    float2* bufs = tw + nc;                                                 // FUSED_GROUPS x 2 x stft_buf_len(nc)
    float* win = reinterpret_cast<float*>(bufs + 2 * FUSED_GROUPS * stft_buf_len(nc));    // n_fft
    float* keys = win + n_fft;                                              // nt_cloud * nf

    const int tid = threadIdx.x;
    const int cloud = blockIdx.x, clip = cloud / clouds_per_clip, chunk = cloud - clip * clouds_per_clip;
    const float* x = audio + (size_t)clip * L;
    for (int i = tid; i < nc; i += TOPK_THREADS) tw[i] = twiddle[i];
    for (int i = tid; i < n_fft; i += TOPK_THREADS) win[i] = window[i];
    __syncthreads();
    const int g = tid >> 7, ltid = tid & 127;
    for (int t0 = 0; t0 < nt_cloud; t0 += FUSED_GROUPS) {
        const int t = t0 + g;
        stft_frame(x, L, n_fft, hop, chunk * nt_cloud + t, tw, win, bufs + (2 * g) * stft_buf_len(nc), bufs + (2 * g + 1) * stft_buf_len(nc), scale, nf,
                   keys + (size_t)t * nf, ltid, 128, t < nt_cloud);
    }
    __syncthreads();
    topk_core<true>(keys, cloud, nt_cloud * nf, nf, farr, tarr, K, kpad, sorted, use_tau, tau, pts_all, idx_all, counts, sortbuf);
}

// ------------------------------------------------------------------------------------ host
static bool g_stft_generic = false;     // debug: force the generic STFT kernel (pca_debug_set_stft_generic)
void debug_set_stft_generic(int on) { g_stft_generic = on != 0; }

int launch_stft_logmag(const float* audio, int n_clips, int n_samples, int n_fft, int hop,
                       const float* window, const float* twiddle, float scale, int drop_nyquist,
                       int nt_out, float* out, cudaStream_t st) {
    if (!audio || !window || !twiddle || !out) return fail(PCA_EINVAL, "stft: null pointer");
    if (n_fft < 16 || n_fft > 8192 || (n_fft & (n_fft - 1))) return fail(PCA_EUNSUPPORTED, "stft: n_fft=%d must be a power of two in [16, 8192]", n_fft);
    if (hop <= 0 || n_clips < 0 || n_samples <= n_fft / 2) return fail(PCA_EINVAL, "stft: need hop > 0 and n_samples > n_fft/2 (reflect padding)");
    const int nt_max = 1 + n_samples / hop;
    if (nt_out < 0 || nt_out > nt_max) return fail(PCA_EINVAL, "stft: nt_out=%d exceeds 1 + n_samples/hop = %d", nt_out, nt_max);
    if (n_clips == 0 || nt_out == 0) return 0;
    const int nc = n_fft / 2;
    const int nf_out = nc + 1 - (drop_nyquist ? 1 : 0);
    if (n_fft >= 512 && n_fft <= 4096 && (reinterpret_cast<size_t>(window) & 7) == 0 && !g_stft_generic) {
        // size-specialised kernel: 256 threads = G frames in flight, group g of block b takes frames b G fpg + g + i G
        const int G = STFT_T_THREADS / (nc / 8);
        int fpg = 1;
        while ((long long)n_clips * ((nt_out + G * fpg - 1) / (G * fpg)) > 148LL * 16 && G * fpg < nt_out) fpg *= 2;
        dim3 grid((nt_out + G * fpg - 1) / (G * fpg), n_clips);
        const size_t smem = ((size_t)nc + (size_t)nc / 8 + (size_t)G * 2 * (nc + nc / 16)) * sizeof(float2);   // (stage tables: < nc/8 entries)
        const double frames = (double)n_clips * nt_out;
        LaunchTimer lt("stft_logmag_kernel", st, frames * (2.5 * n_fft * log2((double)n_fft) + n_fft + 6.0 * nf_out),
                       4.0 * n_clips * n_samples + 4.0 * frames * nf_out);
        const float2* tw2 = reinterpret_cast<const float2*>(twiddle);
#define PCA_STFT_T(NC_)                                                                                                  \
        do {                                                                                                             \
            if (smem > 48 * 1024) PCA_CHECK_CUDA(cudaFuncSetAttribute(stft_logmag_t_kernel<NC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            PCA_CHECK_CUDA(cudaFuncSetAttribute(stft_logmag_t_kernel<NC_>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)); \
            stft_logmag_t_kernel<NC_><<<grid, STFT_T_THREADS, smem, st>>>(audio, n_samples, hop, window, tw2, scale, nf_out, nt_out, fpg, out); \
        } while (0)
        if (nc == 256) PCA_STFT_T(256);
        else if (nc == 512) PCA_STFT_T(512);
        else if (nc == 1024) PCA_STFT_T(1024);
        else PCA_STFT_T(2048);
#undef PCA_STFT_T
        PCA_CHECK_LAUNCH("stft_logmag_t_kernel");
        return 0;
    }
    int threads = nc / 8;                  // one radix-8 butterfly per thread and stage
    threads = threads < 64 ? 64 : (threads > 512 ? 512 : threads);
    // enough blocks to fill 148 SMs several times over, while amortising the table loads
    int fpb = 1;
    while ((long long)n_clips * ((nt_out + fpb - 1) / fpb) > 148LL * 32 && fpb < nt_out) fpb *= 2;
    dim3 grid((nt_out + fpb - 1) / fpb, n_clips);
    const size_t smem = ((size_t)nc + 2 * (size_t)stft_buf_len(nc)) * sizeof(float2) + (size_t)n_fft * sizeof(float);
    if (smem > 48 * 1024) PCA_CHECK_CUDA(cudaFuncSetAttribute(stft_logmag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        const double frames = (double)n_clips * nt_out;
        LaunchTimer lt("stft_logmag_kernel", st, frames * (2.5 * n_fft * log2((double)n_fft) + n_fft + 6.0 * nf_out),
                       4.0 * n_clips * n_samples + 4.0 * frames * nf_out);
        stft_logmag_kernel<<<grid, threads, smem, st>>>(audio, n_samples, n_fft, hop, window,
                                                        reinterpret_cast<const float2*>(twiddle), scale,
                                                        nf_out, nt_out, fpb, out);
    }
    PCA_CHECK_LAUNCH("stft_logmag_kernel");
    return 0;
}

int launch_build_clouds(const float* logmag, int n_clouds, int nf, int nt, const float* farr,
                        const float* tarr, float* pts, cudaStream_t st) {
    if (!logmag || !farr || !pts) return fail(PCA_EINVAL, "build_clouds: null pointer");
    if (n_clouds < 0 || nf <= 0 || nt <= 0) return fail(PCA_EINVAL, "build_clouds: bad shape");
    const long long total = (long long)n_clouds * nf * nt;
    if (total == 0) return 0;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    {
        LaunchTimer lt("build_clouds_kernel", st, 0.0, 4.0 * total * (1 + (tarr ? 3 : 2)));
        build_clouds_kernel<<<(int)blocks, 256, 0, st>>>(logmag, total, nf, nt, farr, tarr, pts);
    }
    PCA_CHECK_LAUNCH("build_clouds_kernel");
    return 0;
}

int launch_topk(const float* keys, int n_clouds, int nf, int nt, const float* farr, const float* tarr,
                int K, int sorted_desc, int use_tau, float tau, float* pts, int32_t* idx, int32_t* counts, cudaStream_t st) {
    if (!keys || (!pts && !idx && !counts)) return fail(PCA_EINVAL, "topk: null pointer");
    if (use_tau && tau != tau) return fail(PCA_EINVAL, "topk: threshold is NaN");
    if (pts && !farr) return fail(PCA_EINVAL, "topk: farr required when pts is requested");
    if (n_clouds < 0 || nf <= 0 || nt <= 0) return fail(PCA_EINVAL, "topk: bad shape");
    const long long N = (long long)nf * nt;
    if (N > (1LL << 30)) return fail(PCA_EUNSUPPORTED, "topk: cloud too large");
    if (K < 0 || K > N) return fail(PCA_EINVAL, "topk: K=%d outside [0, %lld]", K, N);
    if (n_clouds == 0 || K == 0) return 0;
    int kpad = 0;
    size_t smem = 0;
    if (sorted_desc) {
        if (K > 16384) return fail(PCA_EUNSUPPORTED, "topk: sorted output supports K <= 16384 (got %d)", K);
        kpad = 2;
        while (kpad < K) kpad <<= 1;
        smem = (size_t)kpad * sizeof(unsigned long long);
        if (smem > 24 * 1024) {      // static (17.5 KB of histograms) + dynamic beyond the 48 KB default
            PCA_CHECK_CUDA(cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            PCA_CHECK_CUDA(cudaFuncSetAttribute(topk_reg_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            PCA_CHECK_CUDA(cudaFuncSetAttribute(topk_reg_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            PCA_CHECK_CUDA(cudaFuncSetAttribute(topk_reg_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
    }
    {
        // algorithmic traffic: keys read once, 16 B per selected point written
        LaunchTimer lt("topk_kernel", st, 0.0, (double)n_clouds * (4.0 * N + 16.0 * K));
        // clouds of up to 16 384 points: keys in registers (one thread owns KPT consecutive points)
        const int kpt = (int)((N + TOPK_THREADS - 1) / TOPK_THREADS);
        if (kpt <= 8) topk_reg_kernel<8><<<n_clouds, TOPK_THREADS, smem, st>>>(keys, (int)N, nf, farr, tarr, K, kpad, sorted_desc, use_tau, tau, pts, idx, counts);
        else if (kpt <= 16) topk_reg_kernel<16><<<n_clouds, TOPK_THREADS, smem, st>>>(keys, (int)N, nf, farr, tarr, K, kpad, sorted_desc, use_tau, tau, pts, idx, counts);
        else if (kpt <= 32) topk_reg_kernel<32><<<n_clouds, TOPK_THREADS, smem, st>>>(keys, (int)N, nf, farr, tarr, K, kpad, sorted_desc, use_tau, tau, pts, idx, counts);
        else topk_kernel<<<n_clouds, TOPK_THREADS, smem, st>>>(keys, (int)N, nf, farr, tarr, K, kpad, sorted_desc, use_tau, tau, pts, idx, counts);
    }
    PCA_CHECK_LAUNCH("topk_kernel");
    return 0;
}

int launch_fused_frontend(const float* audio, int n_clips, int n_samples, int n_fft, int hop, const float* window,
                          const float* twiddle, float scale, int drop_nyquist, int ntemp, const float* farr,
                          const float* tarr, int K, int sorted_desc, int use_tau, float tau, float* pts, int32_t* idx,
                          int32_t* counts, cudaStream_t st) {
    if (!audio || !window || !twiddle || !farr || !tarr) return fail(PCA_EINVAL, "fused front end: null pointer");
    if (!pts && !idx && !counts) return fail(PCA_EINVAL, "fused front end: no output requested");
    if (n_fft < 16 || n_fft > 8192 || (n_fft & (n_fft - 1))) return fail(PCA_EUNSUPPORTED, "fused front end: n_fft=%d must be a power of two in [16, 8192]", n_fft);
    if (hop <= 0 || n_clips < 0 || n_samples <= n_fft / 2) return fail(PCA_EINVAL, "fused front end: need hop > 0 and n_samples > n_fft/2");
    if (use_tau && tau != tau) return fail(PCA_EINVAL, "fused front end: threshold is NaN");
    const int nt_all = 1 + n_samples / hop;
    if (ntemp <= 0 || ntemp > nt_all) return fail(PCA_EINVAL, "fused front end: ntemp=%d outside [1, %d]", ntemp, nt_all);
    const int clouds_per_clip = nt_all / ntemp;
    const int nc = n_fft / 2, nf = nc + 1 - (drop_nyquist ? 1 : 0);
    const long long N = (long long)nf * ntemp;
    if (K <= 0 || K > N) return fail(PCA_EINVAL, "fused front end: K=%d outside [1, %lld]", K, N);
    if (n_clips == 0) return 0;
    int kpad = 0;
    if (sorted_desc) {
        if (K > 16384) return fail(PCA_EUNSUPPORTED, "fused front end: sorted output supports K <= 16384 (got %d)", K);
        kpad = 2;
        while (kpad < K) kpad <<= 1;
    }
    const size_t smem = (size_t)kpad * 8 + ((size_t)nc + 2 * FUSED_GROUPS * (size_t)stft_buf_len(nc)) * 8 + (size_t)n_fft * 4 + (size_t)N * 4;
    if (smem > 227 * 1024)
        return fail(PCA_EUNSUPPORTED, "fused front end: cloud of %lld points with K=%d needs %zu B of shared memory (> 227 KB); use the unfused calls",
                    N, K, smem);
    PCA_CHECK_CUDA(cudaFuncSetAttribute(fused_frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        const double clouds = (double)n_clips * clouds_per_clip;
        LaunchTimer lt("fused_frontend_kernel", st, clouds * ntemp * (2.5 * n_fft * log2((double)n_fft) + n_fft + 6.0 * nf),
                       4.0 * n_clips * (double)n_samples + clouds * 16.0 * K);
        fused_frontend_kernel<<<n_clips * clouds_per_clip, TOPK_THREADS, smem, st>>>(
            audio, n_samples, n_fft, hop, window, reinterpret_cast<const float2*>(twiddle), scale, nf, ntemp, clouds_per_clip, farr,
            tarr, K, kpad, sorted_desc, use_tau, tau, pts, idx, counts);
    }
    PCA_CHECK_LAUNCH("fused_frontend_kernel");
    return 0;
}

}  // namespace pca
