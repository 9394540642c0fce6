"""CPU oracle for the audio -> point-cloud -> set-encoder hot path.

THIS FILE IS TEST INFRASTRUCTURE.  It is a CPU restatement (numpy / torch-CPU) of the
reference algorithm, used only as the checker by ``tests/``, by ``__graft_entry__.smoke()``
and by the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product
package (``point-cloud-audio_b200``) never imports it and has no CPU fallback.

Parity status
-------------
* Point-cloud construction, top-K selection and the set encoder (MAB/ISAB/PMA/ST,
  DeepSet, the ModelNet SetTransformer): PINNED.  ``tests/golden/make_golden.py`` imports
  the unmodified reference from ``/root/reference`` in the authoring container, runs it on
  seeded inputs (and on the shipped FST / 3ST checkpoints) and freezes the outputs under
  ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below against
  those vectors.
* STFT: PARITY UNPINNED against librosa.  The reference calls third-party
  ``librosa.stft`` (librosa==0.8.0, ``environment.yml:82``), which is not vendored in
  ``/root/reference`` and not installed; the reference has no golden vectors for it.
  ``stft_librosa080`` restates librosa 0.8.0 ``core/spectrum.py::stft`` semantics
  (periodic Hann via ``scipy.signal.get_window(fftbins=True)``, ``util.pad_center`` of the
  window, ``np.pad(mode='reflect')`` by ``n_fft//2``, ``util.frame``, float64 ``rfft``,
  complex64 result) and is cross-checked against two independent implementations
  (``torch.stft`` and ``scipy.signal.stft``) in the CPU tests.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math

import numpy as np

try:  # torch is only needed for the encoder restatement
    import torch
except Exception:  # pragma: no cover
    torch = None


# --------------------------------------------------------------------------------------
# L2: spectral front end
# --------------------------------------------------------------------------------------
def hann_periodic(win_length: int) -> np.ndarray:
    """Periodic Hann window, float64.

    librosa 0.8.0 ``filters.get_window('hann', win_length, fftbins=True)`` delegates to
    ``scipy.signal.get_window``; the periodic Hann is 0.5 - 0.5 cos(2 pi n / win_length).
    Called from ``librosa.stft`` as used at Code/settransformer.py:49.
    """
    n = np.arange(win_length, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)


def padded_window(n_fft: int, win_length: int | None = None) -> np.ndarray:
    """Window centred and zero padded to n_fft (librosa ``util.pad_center``).

    Matters for the eval variant with win_length < n_fft (Code/pceval.py:76).
    """
    if win_length is None:
        win_length = n_fft
    w = hann_periodic(win_length)
    lpad = (n_fft - win_length) // 2
    out = np.zeros(n_fft, dtype=np.float64)
    out[lpad:lpad + win_length] = w
    return out


def stft_librosa080(x: np.ndarray, n_fft: int, win_length: int | None = None,
                    hop_length: int | None = None) -> np.ndarray:
    """``librosa.stft(x, n_fft, hop_length, win_length, window='hann', center=True,
    pad_mode='reflect')`` as of librosa 0.8.0 -> complex64 (1 + n_fft//2, n_frames).

    Call sites: Code/settransformer.py:49, Code/settransformertemp.py:51,
    Code/pc_temp3d_eval.py:130, Code/pceval.py:76.
    """
    x = np.asarray(x)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    w = padded_window(n_fft, win_length)
    y = np.pad(x, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(y) - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    frames = y[idx]                                    # (n_fft, n_frames)
    spec = np.fft.rfft(w[:, None] * frames, axis=0)    # float64 FFT
    return spec.astype(np.complex64)


def logmag_recipe(x: np.ndarray, n_fft: int, hop_factor: float = 0.5,
                  win_length: int | None = None, drop_nyquist: bool = False,
                  divisor: float | None = None) -> np.ndarray:
    """Inline recipe of the training / eval scripts -> float32 (Nf, Nt).

    ``x = librosa.stft(...)/Nfft`` (complex64) ; ``[x = x[:-1,:]]`` ;
    ``a = np.log(1.0e-8 + np.abs(x))``.
    Code/settransformer.py:49-50 (FST), Code/settransformertemp.py:51-53 (3ST, drops the
    Nyquist bin).  In the eval variant the divisor is the window length N, not n_fft
    (Code/pceval.py:76).
    """
    if win_length is None:
        win_length = n_fft
    if divisor is None:
        divisor = win_length
    hop = int(win_length * hop_factor)
    s = stft_librosa080(np.asarray(x, dtype=np.float32), n_fft, win_length, hop) / divisor
    s = s.astype(np.complex64)
    if drop_nyquist:
        s = s[:-1, :]
    return np.log(np.float32(1.0e-8) + np.abs(s)).astype(np.float32)


def chunk_frames(a: np.ndarray, ntemp: int) -> np.ndarray:
    """``np.hsplit(a, np.arange(0, T, Ntemp))`` keeping only full-width chunks, then
    ``np.dstack`` -> (Nf, Ntemp, n_chunks).  Code/settransformertemp.py:54-61.
    """
    nt = a.shape[1]
    n_chunks = nt // ntemp
    if n_chunks == 0:
        return np.zeros((a.shape[0], ntemp, 0), dtype=a.dtype)
    return np.stack([a[:, c * ntemp:(c + 1) * ntemp] for c in range(n_chunks)], axis=2)


def coord_tables(fs: float, nf: int, n_fft: int, hop_factor: float, ntemp: int | None):
    """``farr = linspace(0, fs/2, Nf)/fs`` ; ``tarr = linspace(0, (hf*Nfft/fs)*Ntemp, Ntemp)``.

    float64, exactly as Code/settransformer.py:40 and Code/settransformertemp.py:40-41
    (note farr[-1] == 0.5 even when the Nyquist bin was dropped, and tarr is
    endpoint-inclusive).
    """
    farr = np.linspace(0, fs / 2, nf) / fs
    tarr = None
    if ntemp is not None:
        tarr = np.linspace(0, ((hop_factor * n_fft) / fs) * ntemp, ntemp)
    return farr, tarr


# --------------------------------------------------------------------------------------
# L3: point clouds and selection
# --------------------------------------------------------------------------------------
def cloud_2d(x: np.ndarray, farr: np.ndarray, idx: int) -> np.ndarray:
    """``ESC_pc.__getitem__``: rows (farr[f], x[f, idx]) -> float32 (Nf, 2).
    Code/dataset.py:50-54 (built in float64, cast once)."""
    pc = np.stack([farr.astype(np.float64), x[:, idx].astype(np.float64)], axis=1)
    return pc.astype(np.float32)


def cloud_2d_ss(x_ss: np.ndarray, f_ss: np.ndarray, idx: int) -> np.ndarray:
    """``ESC_pc_ss.__getitem__``: rows (f_ss[k, idx], x_ss[k, idx]).  Code/dataset.py:75-79."""
    pc = np.stack([f_ss[:, idx].astype(np.float64), x_ss[:, idx].astype(np.float64)], axis=1)
    return pc.astype(np.float32)


def cloud_3d_f64(x: np.ndarray, farr: np.ndarray, tarr: np.ndarray, idx: int) -> np.ndarray:
    """``ESC_pc_temp.__getitem__`` before the float cast: point p = t*Nf + f has columns
    (farr[f], tarr[t], x[f, t, idx]); float64 (Nf*Nt, 3).  Code/dataset.py:160-164."""
    nf, nt = farr.shape[0], tarr.shape[0]
    xt = x[:, :, idx]
    out = np.empty((nf * nt, 3), dtype=np.float64)
    out[:, 0] = np.tile(farr, nt)
    out[:, 1] = np.repeat(tarr, nf)
    out[:, 2] = xt.T.reshape(-1)          # t-major, f fastest
    return out


def cloud_3d(x, farr, tarr, idx) -> np.ndarray:
    """``ESC_pc_temp.__getitem__`` -> float32 (Nf*Nt, 3).  Code/dataset.py:160-166."""
    return cloud_3d_f64(x, farr, tarr, idx).astype(np.float32)


def topk_order(keys: np.ndarray, k: int) -> np.ndarray:
    """Selection order contract: ``(-keys).argsort(kind='stable')[:k]``.

    The reference calls ``(-pc[:,-1]).argsort()[:K]`` (Code/dataset.py:199,
    Code/utils.py:43) with numpy's default (unstable) sort; on tie-free keys both agree
    exactly, and with ties the contract is "lowest flat index first" (SURVEY.md 8c).
    """
    return np.argsort(-keys, kind="stable")[:k]


def cloud_3d_maxk_f64(x, farr, tarr, idx, k) -> tuple[np.ndarray, np.ndarray]:
    """``ESC_pc_temp_maxKSS.__getitem__``: full cloud, then rows in descending-magnitude
    order, float64 (K, 3).  Code/dataset.py:194-202.  Also returns the flat indices."""
    pc = cloud_3d_f64(x, farr, tarr, idx)
    order = topk_order(pc[:, -1], k)
    return pc[order, :], order


def cloud_3d_randk_f64(x, farr, tarr, idx, k, rng=np.random):
    """``ESC_pc_temp_randKSS.__getitem__``: full cloud, then ``np.random.permutation(P)[:K]`` rows (a uniformly random
    K-subset in random order), float64 (K, 3).  Code/dataset.py:230-238.  ``rng`` is numpy's global generator in the
    reference; pass a RandomState to reproduce a seeded run.  Also returns the flat indices."""
    pc = cloud_3d_f64(x, farr, tarr, idx)
    order = rng.permutation(pc.shape[0])[:k]
    return pc[order, :], order


def pc_randk(x: np.ndarray, farr: np.ndarray, kmax: int, rng=np.random):
    """``utils.pc_randK``: per frame a random K-subset of the spectrum -> (mags (K,T), freqs (K,T)).  Code/utils.py:55-82."""
    xs, fs_ = [], []
    for t in range(x.shape[1]):
        order = rng.permutation(x.shape[0])[:kmax]
        xs.append(x[order, t])
        fs_.append(farr[order])
    return np.stack(xs, axis=1), np.stack(fs_, axis=1)


def importance_heat(xt: np.ndarray, winf: int) -> np.ndarray:
    """Heat map of ``ESC_pc_temp_importancerandKSS`` (Code/dataset.py:280-283): |d/df| + |d/dt| of the (Nf, Nt)
    log-magnitudes by ``torch.gradient`` (unit spacing, one-sided at the edges), smoothed by the 2 x winF outer product
    of periodic Kaiser windows (beta 5.09) with ``conv2d(padding='same')``, plus 1e-6.  float32 (Nf, Nt)."""
    t = torch.as_tensor(np.ascontiguousarray(xt))
    g = torch.gradient(t)
    g = g[0].abs() + g[1].abs()
    k = (torch.kaiser_window(window_length=2, periodic=True, beta=5.09)[:, None]
         @ torch.kaiser_window(window_length=winf, periodic=True, beta=5.09)[None, :])
    return (torch.nn.functional.conv2d(g[None, None, ...], k[None, None], padding="same")[0, 0] + 1.0e-6).numpy()


def cloud_3d_importance_f64(x, farr, tarr, idx, k, winf, choice=1, generator=None):
    """``ESC_pc_temp_importancerandKSS.__getitem__`` (Code/dataset.py:276-290): choice 1 keeps the K largest heat-map
    entries (``(-g.view(-1)).argsort()[:K]``), choice 0 draws K of them with ``torch.multinomial(..., replacement=True)``.
    As in the reference the heat map is flattened f-major (index f*Nt + t) while the cloud rows are t-major
    (p = t*Nf + f), and the heat-map index is used on the cloud rows unchanged.  Returns (rows (K,3) float64, indices)."""
    pc = cloud_3d_f64(x, farr, tarr, idx)
    g = importance_heat(x[:, :, idx], winf)
    if choice == 0:
        order = torch.multinomial(torch.from_numpy(g).view(-1), k, replacement=True, generator=generator).numpy()
    else:
        order = topk_order(g.reshape(-1), k)
    return pc[order, :], order


def pc_maxk(x: np.ndarray, farr: np.ndarray, kmax: int):
    """``utils.pc_maxK``: per-frame top-K of the spectrum -> (mags (K,T), freqs (K,T)).
    Code/utils.py:25-52 (keys are the float32 spectrum column)."""
    nt = x.shape[1]
    k = min(kmax, x.shape[0])
    xs = np.empty((k, nt), dtype=x.dtype)
    fs_ = np.empty((k, nt), dtype=farr.dtype)
    for t in range(nt):
        order = topk_order(x[:, t], kmax)
        xs[:, t] = x[order, t]
        fs_[:, t] = farr[order]
    return xs, fs_


# --------------------------------------------------------------------------------------
# L1.5: test-time resampling (PARITY UNPINNED: resampy 0.2.2 / librosa 0.8.0 are not in the image)
# --------------------------------------------------------------------------------------
RESAMPY_FILTERS = {     # resampy 0.2.2 filters.py docstring: zero crossings, Kaiser beta, roll-off; 2**9 table entries per crossing
    "kaiser_best": (64, 14.769656459379492, 0.9475937167399596),
    "kaiser_fast": (16, 8.555504641634386, 0.85),
}


def resampy_filter(name: str = "kaiser_fast", precision: int = 9):
    """Half window of resampy's interpolation filter, rebuilt from its published recipe ``filters.sinc_window``:
    ``rolloff * sinc(rolloff * linspace(0, num_zeros, n+1))`` tapered by the right half of a symmetric Kaiser window of
    2n+1 points, n = num_zeros * 2**precision.  Returns (half_window float64 (n+1,), num_table = 2**precision).
    The shipped ``kaiser_*.npz`` tables cannot be read here, so this table -- and everything resampled with it -- is unpinned."""
    num_zeros, beta, rolloff = RESAMPY_FILTERS[name]
    num_table = 2 ** precision
    n = num_table * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = np.kaiser(2 * n + 1, beta)[n:]
    return taper * sinc_win, num_table


def resample_librosa080(x: np.ndarray, orig_sr: float, target_sr: float, res_type: str = "kaiser_fast", fix: bool = True,
                        scale: bool = False) -> np.ndarray:
    """``librosa.resample`` 0.8.0 with a resampy 0.2.2 filter (call sites: Code/pceval.py:75, Code/pc_temp3d_eval.py:74):
    resampy.interpn.resample_f restated per output sample (left wing over x[n-i], right wing over x[n+1+k], window linearly
    interpolated in its table), output length int(n*ratio), then ``fix_length`` to ceil(n*ratio) and ``/= sqrt(ratio)`` for
    scale=True.  float64 accumulation, result in x.dtype.  UNPINNED (see ``resampy_filter``)."""
    x = np.asarray(x)
    if orig_sr == target_sr:
        return x
    ratio = float(target_sr) / orig_sr
    interp_win, num_table = resampy_filter(res_type)
    interp_win = interp_win.copy()
    if ratio < 1:
        interp_win *= ratio
    interp_delta = np.zeros_like(interp_win)
    interp_delta[:-1] = np.diff(interp_win)
    n_in = x.shape[-1]
    n_res = int(n_in * ratio)
    scale_f = min(1.0, ratio)
    time_increment = 1.0 / ratio
    index_step = int(scale_f * num_table)
    nwin = interp_win.shape[0]
    y = np.zeros(x.shape[:-1] + (n_res,), dtype=np.float64)
    xd = x.astype(np.float64)
    for t in range(n_res):
        time_register = t * time_increment
        n = int(time_register)
        frac = scale_f * (time_register - n)
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        i_max = min(n + 1, (nwin - offset) // index_step)
        if i_max > 0:
            j = offset + np.arange(i_max) * index_step
            y[..., t] += ((interp_win[j] + eta * interp_delta[j]) * xd[..., n - np.arange(i_max)]).sum(-1)
        frac = scale_f - frac
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        k_max = min(n_in - n - 1, (nwin - offset) // index_step)
        if k_max > 0:
            j = offset + np.arange(k_max) * index_step
            y[..., t] += ((interp_win[j] + eta * interp_delta[j]) * xd[..., n + 1 + np.arange(k_max)]).sum(-1)
    if fix:
        n_fix = int(np.ceil(n_in * ratio))
        if n_fix > n_res:
            y = np.concatenate([y, np.zeros(y.shape[:-1] + (n_fix - n_res,))], axis=-1)
        else:
            y = y[..., :n_fix]
    if scale:
        y = y / np.sqrt(ratio)
    return np.ascontiguousarray(y, dtype=x.dtype)


# --------------------------------------------------------------------------------------
# L4: set encoder (torch CPU restatement; dtype follows the inputs/weights)
# --------------------------------------------------------------------------------------
def _linear(x, w, b):
    return x @ w.transpose(-1, -2) + b


def mab_forward(p: dict, prefix: str, Q, K, num_heads: int):
    """``MAB.forward`` (set_transformer-master/modules.py:19-33).

    p[prefix+'fc_q.weight'] etc. are nn.Linear-layout tensors.  Scale is
    1/sqrt(dim_V) (NOT 1/sqrt(head dim)), the residual adds the PROJECTED Q, and the
    output is O + relu(fc_o(O)); optional LayerNorms when the ln0/ln1 keys exist.
    """
    q = _linear(Q, p[prefix + "fc_q.weight"], p[prefix + "fc_q.bias"])
    k = _linear(K, p[prefix + "fc_k.weight"], p[prefix + "fc_k.bias"])
    v = _linear(K, p[prefix + "fc_v.weight"], p[prefix + "fc_v.bias"])
    B, nq, D = q.shape
    nk = k.shape[1]
    dh = D // num_heads
    qh = q.reshape(B, nq, num_heads, dh).permute(0, 2, 1, 3)
    kh = k.reshape(B, nk, num_heads, dh).permute(0, 2, 1, 3)
    vh = v.reshape(B, nk, num_heads, dh).permute(0, 2, 1, 3)
    a = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(D), dim=-1)
    o = (qh + a @ vh).permute(0, 2, 1, 3).reshape(B, nq, D)
    if prefix + "ln0.weight" in p:
        o = torch.nn.functional.layer_norm(o, (D,), p[prefix + "ln0.weight"], p[prefix + "ln0.bias"])
    o = o + torch.relu(_linear(o, p[prefix + "fc_o.weight"], p[prefix + "fc_o.bias"]))
    if prefix + "ln1.weight" in p:
        o = torch.nn.functional.layer_norm(o, (D,), p[prefix + "ln1.weight"], p[prefix + "ln1.bias"])
    return o


def sab_forward(p, prefix, X, num_heads):
    """``SAB.forward`` = MAB(X, X).  modules.py:40-41."""
    return mab_forward(p, prefix + "mab.", X, X, num_heads)


def isab_forward(p, prefix, X, num_heads):
    """``ISAB.forward``: H = mab0(I, X); mab1(X, H).  modules.py:51-53."""
    I = p[prefix + "I"].expand(X.shape[0], -1, -1)
    H = mab_forward(p, prefix + "mab0.", I, X, num_heads)
    return mab_forward(p, prefix + "mab1.", X, H, num_heads)


def pma_forward(p, prefix, X, num_heads):
    """``PMA.forward``: mab(S, X).  modules.py:62-63."""
    S = p[prefix + "S"].expand(X.shape[0], -1, -1)
    return mab_forward(p, prefix + "mab.", S, X, num_heads)


def strip_module_prefix(state: dict) -> dict:
    """Shipped checkpoints were saved from nn.DataParallel -> keys start with 'module.'."""
    return {(k[7:] if k.startswith("module.") else k): v for k, v in state.items()}


def st_forward(p: dict, X, num_heads: int):
    """``ST.forward`` (Code/models.py:33-44): enc = ISAB, ISAB; dec = PMA, Linear; then
    ``.squeeze()`` ((B,1,C)->(B,C); (C,) when B == 1).  Identical layer keys/semantics for
    ``main_pointcloud.SetTransformer`` in eval mode, where dec = Dropout, PMA, Dropout,
    Linear (set_transformer-master/main_pointcloud.py:24-37) -> pass pma_key/lin_key."""
    return set_encoder_forward(p, X, num_heads, "dec.0.", "dec.1.")


def set_encoder_forward(p, X, num_heads, pma_key="dec.0.", lin_key="dec.1."):
    y = isab_forward(p, "enc.0.", X, num_heads)
    y = isab_forward(p, "enc.1.", y, num_heads)
    y = pma_forward(p, pma_key, y, num_heads)
    y = _linear(y, p[lin_key + "weight"], p[lin_key + "bias"])
    return y.squeeze()


def modelnet_forward(p, X, num_heads):
    """``main_pointcloud.SetTransformer.forward`` in eval mode (Dropout = identity):
    dec.1 is the PMA and dec.3 the Linear.  main_pointcloud.py:24-37."""
    return set_encoder_forward(p, X, num_heads, "dec.1.", "dec.3.")


def deepset_forward(p, X, num_outputs: int, dim_output: int, pool: str = "mean"):
    """``DeepSet.forward`` (set_transformer-master/models.py:25-28): 4x shared Linear
    (+ReLU between) over points, pool over points, 4x Linear decoder, reshape.
    pool='max'/'sum' follow ``SmallDeepSet`` (max_regression_demo.ipynb:41-48)."""
    h = X
    for i in (0, 2, 4, 6):
        h = _linear(h, p[f"enc.{i}.weight"], p[f"enc.{i}.bias"])
        if i != 6:
            h = torch.relu(h)
    if pool == "mean":
        h = h.mean(-2)
    elif pool == "sum":
        h = h.sum(-2)
    elif pool == "max":
        h = h.max(dim=-2)[0]
    else:
        raise ValueError(pool)
    for i in (0, 2, 4, 6):
        h = _linear(h, p[f"dec.{i}.weight"], p[f"dec.{i}.bias"])
        if i != 6:
            h = torch.relu(h)
    return h.reshape(-1, num_outputs, dim_output)


# --------------------------------------------------------------------------------------
# Whole path (used by the CPU baseline and the integration parity tests)
# --------------------------------------------------------------------------------------
def synth_audio(n_clips: int, n_samples: int, fs: float, seed: int) -> np.ndarray:
    """Seeded synthetic audio: 0.05*randn + 4 random sinusoids, clamped
    to [-1, 1], float32 (n_clips, n_samples).  Deterministic across numpy versions
    (legacy RandomState stream)."""
    rs = np.random.RandomState(seed)
    n = np.arange(n_samples, dtype=np.float64)
    out = np.empty((n_clips, n_samples), dtype=np.float32)
    for c in range(n_clips):
        x = 0.05 * rs.randn(n_samples)
        for _ in range(4):
            a = rs.uniform(0.05, 0.4)
            f = rs.uniform(50.0, fs / 2 - 50.0)
            ph = rs.uniform(0.0, 2 * np.pi)
            x = x + a * np.sin(2 * np.pi * f * n / fs + ph)
        out[c] = np.clip(x, -1.0, 1.0).astype(np.float32)
    return out


def clip_frame_clouds(audio: np.ndarray, fs: float, n_fft: int, hop_factor: float = 0.5):
    """FST path for one clip: recipe -> one 2-D cloud (Nf, 2) per frame.
    Code/settransformer.py:45-54,70 + Code/dataset.py:50-54."""
    a = logmag_recipe(audio, n_fft, hop_factor)
    farr, _ = coord_tables(fs, a.shape[0], n_fft, hop_factor, None)
    return np.stack([cloud_2d(a, farr, t) for t in range(a.shape[1])], axis=0)


def clip_chunk_clouds(audio: np.ndarray, fs: float, n_fft: int, hop_factor: float, ntemp: int,
                      top_k: int | None = None):
    """3ST path for one clip: recipe (Nyquist dropped) -> Ntemp-frame chunks -> 3-D clouds
    (n_chunks, Nf*Ntemp | K, 3) float32.  Code/settransformertemp.py:51-61,
    Code/dataset.py:160-166,194-202."""
    a = logmag_recipe(audio, n_fft, hop_factor, drop_nyquist=True)
    x = chunk_frames(a, ntemp)
    farr, tarr = coord_tables(fs, x.shape[0], n_fft, hop_factor, ntemp)
    out = []
    for c in range(x.shape[2]):
        if top_k is None:
            out.append(cloud_3d(x, farr, tarr, c))
        else:
            out.append(cloud_3d_maxk_f64(x, farr, tarr, c, top_k)[0].astype(np.float32))
    return np.stack(out, axis=0) if out else np.zeros((0, 0, 3), np.float32)
