"""Batched spectral front end and point selection (host mirror over the C ABI).

The reference has no function for "audio -> points": the recipe is inline script code
(Code/settransformer.py:45-54, Code/settransformertemp.py:49-61, Code/pc_temp3d_eval.py:126-143).
``stft_logmag`` / ``spectral_point_cloud`` are the batched equivalents; their results equal that
recipe followed by the Dataset ``__getitem__`` of Code/dataset.py applied per clip.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, _runtime as rt
from . import ops as _ops  # noqa: F401  (registers torch.ops.pcaudio.*)


def coord_tables(fs: float, nf: int, n_fft: int, hop_factor: float, ntemp: int | None = None):
    """farr = linspace(0, fs/2, Nf)/fs ; tarr = linspace(0, (hf*Nfft/fs)*Ntemp, Ntemp), float64,
    exactly the expressions of Code/settransformer.py:40 and Code/settransformertemp.py:40-41."""
    farr = np.linspace(0, fs / 2, nf) / fs
    tarr = None if ntemp is None else np.linspace(0, ((hop_factor * n_fft) / fs) * ntemp, ntemp)
    return farr, tarr


def stft_logmag(audio: torch.Tensor, n_fft: int, win_length: int | None = None, hop_factor: float = 0.5,
                drop_nyquist: bool = False, divisor: float | None = None, n_frames: int | None = None) -> torch.Tensor:
    """log(1e-8 + |librosa.stft(x, n_fft, win_length, hop=int(win_length*hf), 'hann')| / divisor).

    audio (B, L) float32 CUDA -> (B, Nt, Nf) float32 CUDA, frequency fastest (the transpose of the
    reference's (Nf, Nt) matrix, i.e. already in point order p = t*Nf + f).  ``divisor`` defaults to
    win_length (the scripts divide by N: Code/settransformer.py:49, Code/pceval.py:76)."""
    rt.require_cuda(audio, "stft_logmag")
    if audio.dim() != 2:
        raise ValueError("audio must be (n_clips, n_samples)")
    audio = rt.f32c(audio)
    win_length = n_fft if win_length is None else int(win_length)
    hop = int(win_length * hop_factor)
    divisor = float(win_length if divisor is None else divisor)
    B, L = audio.shape
    nt_all = 1 + L // hop
    nt = nt_all if n_frames is None else int(n_frames)
    nf = n_fft // 2 + 1 - (1 if drop_nyquist else 0)
    win, tw = rt.stft_tables(n_fft, win_length, audio.device)
    return torch.ops.pcaudio.stft_logmag(audio, win, tw, n_fft, hop, 1.0 / divisor, bool(drop_nyquist), nt)


def build_clouds(logmag: torch.Tensor, farr, tarr=None) -> torch.Tensor:
    """(n_clouds, nt, nf) log-magnitudes -> (n_clouds, nt*nf, 3) clouds (f, t, mag), or
    (n_clouds, nf, 2) clouds (f, mag) when tarr is None.  ESC_pc_temp / ESC_pc __getitem__ batched
    (Code/dataset.py:160-166, 50-54)."""
    rt.require_cuda(logmag, "build_clouds")
    logmag = rt.f32c(logmag)
    if logmag.dim() == 2:
        logmag = logmag.unsqueeze(1)
    n, nt, nf = logmag.shape
    dev = logmag.device
    f_t = farr if isinstance(farr, torch.Tensor) else rt.coord_table(farr, dev)
    t_t = None if tarr is None else (tarr if isinstance(tarr, torch.Tensor) else rt.coord_table(tarr, dev))
    if tarr is None and nt != 1:
        raise ValueError("2-D clouds take one frame per cloud")
    width = 2 if t_t is None else 3
    pts = torch.empty((n, nt * nf, width), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pca_build_clouds_f32(_lib.ptr(logmag), n, nf, nt, _lib.ptr(f_t), _lib.ptr(t_t),
                                                   _lib.ptr(pts), rt.stream_ptr(dev)), "build_clouds")
    return pts


def topk_points(logmag: torch.Tensor, farr, tarr, k: int, sorted_desc: bool = True, want_points: bool = True):
    """Per cloud, the k largest log-magnitudes as (f[, t], mag) rows plus their flat indices
    p = t*nf + f.  Emission order (-mag).argsort(kind='stable')[:k] when sorted_desc
    (ESC_pc_temp_maxKSS, Code/dataset.py:199-200; pc_maxK, Code/utils.py:43-45), else scan order.
    Returns (pts (n,k,2|3) float32 or None, idx (n,k) int32)."""
    rt.require_cuda(logmag, "topk_points")
    logmag = rt.f32c(logmag)
    if logmag.dim() == 2:
        logmag = logmag.unsqueeze(1)
    n, nt, nf = logmag.shape
    dev = logmag.device
    f_t = None if farr is None else (farr if isinstance(farr, torch.Tensor) else rt.coord_table(farr, dev))
    t_t = None if tarr is None else (tarr if isinstance(tarr, torch.Tensor) else rt.coord_table(tarr, dev))
    width = 2 if t_t is None else 3
    pts = torch.empty((n, k, width), dtype=torch.float32, device=dev) if want_points else None
    idx = torch.empty((n, k), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pca_topk_compact_f32(_lib.ptr(logmag), n, nf, nt, _lib.ptr(f_t), _lib.ptr(t_t), int(k),
                                                   int(sorted_desc), _lib.ptr(pts), _lib.ptr(idx),
                                                   rt.stream_ptr(dev)), "topk_points")
    return pts, idx


def select_points(logmag: torch.Tensor, farr, tarr, k: int, threshold: float | None = None, sorted_desc: bool = True):
    """Threshold / capped selection into padded sets (extension of ``topk_points``): per cloud the points with
    log-magnitude >= threshold, at most k of them by the top-K rule; rows past the number kept are zero (index -1).
    Returns (pts (n,k,2|3), idx (n,k) int32, counts (n,) int32)."""
    rt.require_cuda(logmag, "select_points")
    logmag = rt.f32c(logmag)
    if logmag.dim() == 2:
        logmag = logmag.unsqueeze(1)
    n, nt, nf = logmag.shape
    dev = logmag.device
    f_t = farr if isinstance(farr, torch.Tensor) else rt.coord_table(farr, dev)
    t_t = None if tarr is None else (tarr if isinstance(tarr, torch.Tensor) else rt.coord_table(tarr, dev))
    return torch.ops.pcaudio.select_points(logmag, f_t, t_t, int(k), bool(sorted_desc), threshold is not None,
                                           float(threshold if threshold is not None else 0.0))


def fused_frontend_fits(n_fft: int, n_pts: int, k: int) -> bool:
    """Shared-memory budget of pca_frontend_fused_f32: sort buffer + FFT tables/buffers + the cloud's keys <= 227 KB."""
    kpad = 2
    while kpad < k:
        kpad <<= 1
    nc = n_fft // 2
    return k <= 16384 and kpad * 8 + (nc + 8 * (nc + nc // 16)) * 8 + n_fft * 4 + n_pts * 4 <= 227 * 1024


def spectral_point_cloud(audio: torch.Tensor, *, n_fft: int, sr: float, win_length: int | None = None,
                         hop_factor: float = 0.5, drop_nyquist: bool = True, ntemp: int | None = None,
                         top_k: int | None = None, sorted_desc: bool = True, threshold: float | None = None,
                         fused: bool | None = None):
    """NEW batched front end: audio (B, L) -> (points (B', K, 3), counts (B',), indices).

    Per clip: STFT recipe -> log-magnitude -> [drop Nyquist] -> non-overlapping ``ntemp``-frame chunks
    (remainder dropped; ntemp=None keeps the whole clip as one cloud) -> (f, t, mag) clouds -> optional
    top-K and/or magnitude threshold (padded sets, ``counts`` = points kept).  B' = B * chunks_per_clip, clouds of
    one clip are consecutive."""
    win_length = n_fft if win_length is None else int(win_length)
    hop = int(win_length * hop_factor)
    B, L = audio.shape
    nt_all = 1 + L // hop
    ntemp_eff = nt_all if ntemp is None else int(ntemp)
    chunks = nt_all // ntemp_eff
    nf = n_fft // 2 + 1 - (1 if drop_nyquist else 0)
    n_pts = ntemp_eff * nf
    k_sel = n_pts if top_k is None else min(int(top_k), n_pts)
    if fused is None:
        # Measured on B200 (tools/frontend_sweep.py, profiles/): the two-kernel route is faster at every K of the
        # BASELINE sweep because the log-magnitude intermediate of a batch stays in the 126 MB L2, while the fused kernel
        # runs one 512-thread block per SM.  The fused launch is kept for callers that must not touch HBM in between.
        fused = False
    if fused:
        # one launch: the log-magnitudes never leave shared memory
        farr, tarr = coord_tables(sr, nf, n_fft, hop_factor, ntemp_eff)
        dev = audio.device
        rt.require_cuda(audio, "spectral_point_cloud")
        audio = rt.f32c(audio)
        win, tw = rt.stft_tables(n_fft, win_length, dev)
        f_t, t_t = rt.coord_table(farr, dev), rt.coord_table(tarr, dev)
        pts = torch.empty((B * chunks, k_sel, 3), dtype=torch.float32, device=dev)
        idx = torch.empty((B * chunks, k_sel), dtype=torch.int32, device=dev)
        counts = torch.empty((B * chunks,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pca_frontend_fused_f32(
                _lib.ptr(audio), B, L, n_fft, hop, _lib.ptr(win), _lib.ptr(tw), 1.0 / win_length, int(drop_nyquist), ntemp_eff,
                _lib.ptr(f_t), _lib.ptr(t_t), k_sel, int(sorted_desc), int(threshold is not None),
                float(threshold if threshold is not None else 0.0), _lib.ptr(pts), _lib.ptr(idx), _lib.ptr(counts),
                rt.stream_ptr(dev)), "spectral_point_cloud(fused)")
        return pts, counts, idx
    logmag = stft_logmag(audio, n_fft, win_length, hop_factor, drop_nyquist, n_frames=chunks * ntemp_eff)
    logmag = logmag.view(B * chunks, ntemp_eff, nf)
    farr, tarr = coord_tables(sr, nf, n_fft, hop_factor, ntemp_eff)
    if threshold is not None:
        # threshold mode: log-magnitude >= threshold, capped at top_k, zero-padded; counts = points kept per cloud
        pts, idx, counts = select_points(logmag, farr, tarr, n_pts if top_k is None else min(int(top_k), n_pts),
                                         float(threshold), sorted_desc)
        return pts, counts, idx
    if top_k is None or top_k >= n_pts and not sorted_desc:
        pts = build_clouds(logmag, farr, tarr)
        idx = torch.arange(n_pts, dtype=torch.int32, device=audio.device).expand(B * chunks, -1)
    else:
        pts, idx = topk_points(logmag, farr, tarr, min(int(top_k), n_pts), sorted_desc)
    counts = torch.full((B * chunks,), pts.shape[1], dtype=torch.int32, device=audio.device)
    return pts, counts, idx


# ------------------------------------------------------------------------------------ random-K / importance subsampling
def _tables(farr, tarr, dev):
    f_t = farr if isinstance(farr, torch.Tensor) else rt.coord_table(farr, dev)
    t_t = None if tarr is None else (tarr if isinstance(tarr, torch.Tensor) else rt.coord_table(tarr, dev))
    return f_t, t_t


def gather_points(logmag: torch.Tensor, farr, tarr, idx: torch.Tensor) -> torch.Tensor:
    """Rows (f[, t], mag) of the flat cloud indices p = t*nf + f in idx (n, K) int32; idx < 0 gives a zero row."""
    rt.require_cuda(logmag, "gather_points")
    logmag = rt.f32c(logmag)
    if logmag.dim() == 2:
        logmag = logmag.unsqueeze(1)
    n, nt, nf = logmag.shape
    dev = logmag.device
    f_t, t_t = _tables(farr, tarr, dev)
    idx = idx.to(device=dev, dtype=torch.int32).contiguous()
    K = idx.shape[1]
    pts = torch.empty((n, K, 2 if t_t is None else 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pca_gather_points_f32(_lib.ptr(logmag), n, nf, nt, _lib.ptr(f_t), _lib.ptr(t_t), _lib.ptr(idx), K,
                                                    _lib.ptr(pts), rt.stream_ptr(dev)), "gather_points")
    return pts


def random_points(logmag: torch.Tensor, farr, tarr, k: int, seed: int = 0):
    """Random-K subsampling (ESC_pc_temp_randKSS, Code/dataset.py:230-238; pc_randK, Code/utils.py:55-82): per cloud a
    uniformly random k-subset of the points in uniformly random order -- the k largest of i.i.d. uniform keys from a
    counter-based generator (the radix-select kernel does the selection).  numpy's permutation stream is not
    reproduced: parity with the reference is distributional.  Returns (pts (n,k,2|3), idx (n,k) int32)."""
    rt.require_cuda(logmag, "random_points")
    logmag = rt.f32c(logmag)
    if logmag.dim() == 2:
        logmag = logmag.unsqueeze(1)
    n, nt, nf = logmag.shape
    dev = logmag.device
    keys = torch.empty((n, 1, nt * nf), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pca_random_keys_f32(_lib.ptr(keys), keys.numel(), int(seed) & 0xFFFFFFFFFFFFFFFF, rt.stream_ptr(dev)),
                   "random_keys")
    _, idx = topk_points(keys, None, None, min(int(k), nt * nf), sorted_desc=True, want_points=False)
    return gather_points(logmag, farr, tarr, idx), idx


def importance_heat(logmag: torch.Tensor, winF: int) -> torch.Tensor:
    """Heat map of ESC_pc_temp_importancerandKSS (Code/dataset.py:280-283) for a batch: (n, nt, nf) log-magnitudes ->
    (n, nf, nt) float32 = conv2d(|d/df| + |d/dt|, kaiser(2) x kaiser(winF), 'same') + 1e-6 (the reference's g, f-major)."""
    rt.require_cuda(logmag, "importance_heat")
    logmag = rt.f32c(logmag)
    n, nt, nf = logmag.shape
    dev = logmag.device
    key = ("kaiser", int(winF), dev.index)
    hit = rt._table_cache.get(key)
    if hit is None:      # the reference's own window call, evaluated once on the host
        hit = (torch.kaiser_window(window_length=2, periodic=True, beta=5.09).float().to(dev),
               torch.kaiser_window(window_length=int(winF), periodic=True, beta=5.09).float().to(dev))
        rt._table_cache[key] = hit
    kf, kt = hit
    heat = torch.empty((n, nf, nt), dtype=torch.float32, device=dev)
    scratch = torch.empty_like(heat)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pca_importance_map_f32(_lib.ptr(logmag), n, nf, nt, _lib.ptr(kf), 2, _lib.ptr(kt), int(winF),
                                                     _lib.ptr(heat), _lib.ptr(scratch), rt.stream_ptr(dev)), "importance_map")
    return heat


def importance_points(logmag: torch.Tensor, farr, tarr, k: int, winF: int, choice: int = 1, seed: int = 0):
    """Importance subsampling (ESC_pc_temp_importancerandKSS, Code/dataset.py:276-290): choice 1 keeps the k largest
    heat-map entries, choice 0 draws k with replacement from the heat map as a categorical distribution
    (torch.multinomial; distributional parity).  Like the reference, the f-major heat-map index selects the row of the
    t-major cloud unchanged.  Returns (pts (n,k,3), idx (n,k) int32 = the indices used on the cloud rows)."""
    logmag = rt.f32c(logmag)
    n, nt, nf = logmag.shape
    dev = logmag.device
    heat = importance_heat(logmag, winF)
    if choice == 0:
        idx = torch.empty((n, int(k)), dtype=torch.int32, device=dev)
        cdf = torch.empty((n, nt * nf), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pca_multinomial_f32(_lib.ptr(heat), n, nt * nf, int(k), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                                      _lib.ptr(cdf), _lib.ptr(idx), rt.stream_ptr(dev)), "multinomial")
    else:
        _, idx = topk_points(heat.view(n, 1, nt * nf), None, None, min(int(k), nt * nf), sorted_desc=True, want_points=False)
    return gather_points(logmag, farr, tarr, idx), idx


# ------------------------------------------------------------------------------------ test-time resampling
RESAMPY_FILTERS = {     # resampy 0.2.2: zero crossings, Kaiser beta, roll-off (2**9 table entries per zero crossing)
    "kaiser_best": (64, 14.769656459379492, 0.9475937167399596),
    "kaiser_fast": (16, 8.555504641634386, 0.85),
}


def _resampy_tables(res_type: str, ratio: float, device):
    """Half window of resampy's interpolation filter (recipe of resampy.filters.sinc_window, float64) and its first
    differences, scaled by the ratio when downsampling; cached per (filter, ratio, device)."""
    key = ("resampy", res_type, float(ratio), torch.device(device).index)
    hit = rt._table_cache.get(key)
    if hit is None:
        num_zeros, beta, rolloff = RESAMPY_FILTERS[res_type]
        num_table = 2 ** 9
        n = num_table * num_zeros
        win = np.kaiser(2 * n + 1, beta)[n:] * (rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True)))
        if ratio < 1:
            win = win * ratio
        delta = np.zeros_like(win)
        delta[:-1] = np.diff(win)
        hit = (torch.from_numpy(win).to(device), torch.from_numpy(delta).to(device), num_table)
        rt._table_cache[key] = hit
    return hit


def resample(audio: torch.Tensor, orig_sr: float, target_sr: float, res_type: str = "kaiser_fast", fix: bool = True,
             scale: bool = False) -> torch.Tensor:
    """Batched ``librosa.resample(x, orig_sr, target_sr, res_type='kaiser_fast', scale=True)`` of the evaluation sweeps
    (Code/pceval.py:75, Code/pc_temp3d_eval.py:74): audio (B, L) float32 CUDA -> (B, ceil(L * ratio)) (``fix=True``; else
    int(L * ratio)).  resampy's band-limited sinc interpolation runs one thread per output sample.  resampy / librosa are
    not in the image: the filter table is rebuilt from resampy's published recipe and this function's parity with the
    reference is unpinned (it is tested against the CPU restatement and cross-checked against scipy's polyphase resampler)."""
    rt.require_cuda(audio, "resample")
    if audio.dim() != 2:
        raise ValueError("audio must be (n_clips, n_samples)")
    if res_type not in RESAMPY_FILTERS:
        raise ValueError(f"res_type must be one of {sorted(RESAMPY_FILTERS)}")
    if orig_sr == target_sr:
        return audio
    audio = rt.f32c(audio)
    B, L = audio.shape
    ratio = float(target_sr) / float(orig_sr)
    n_res = int(L * ratio)
    n_out = int(np.ceil(L * ratio)) if fix else n_res
    dev = audio.device
    win, delta, num_table = _resampy_tables(res_type, ratio, dev)
    out = torch.zeros((B, n_out), dtype=torch.float32, device=dev)
    if B == 0 or n_res == 0:
        return out
    tmp = out if n_out == n_res else torch.empty((B, n_res), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pca_resample_f32(_lib.ptr(audio), B, L, n_res, ratio, _lib.ptr(win), _lib.ptr(delta), win.numel(),
                                               num_table, float(1.0 / np.sqrt(ratio)) if scale else 1.0, _lib.ptr(tmp),
                                               rt.stream_ptr(dev)), "resample")
    if tmp is not out:
        out[:, :min(n_res, n_out)] = tmp[:, :min(n_res, n_out)]
    return out
