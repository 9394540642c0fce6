"""torch custom-op layer over the C ABI (``torch.ops.pcaudio.*``).

The host mirrors call the library through these ops, so the hot path is visible to the PyTorch dispatcher
(``torch.library``: schema, fake/meta shape functions for tracing, CUDA-only implementations) while the arithmetic stays
behind ``include/pcaudio_b200.h``.  There is no CPU implementation registered on purpose: calling an op with CPU tensors
fails in the dispatcher.  The training ops (``st_train_fwd`` / ``st_train_bwd`` ...) are plain forward ops too: gradients are
wired by the ``torch.autograd.Function`` classes of ``training.py`` around the hand-written backward kernels, not by
``register_autograd`` formulas."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, _runtime as rt

ST_WORKSPACE_BYTES = 8 << 30      # scratch budget for the encoder (180 GB of HBM per GPU): larger batches are processed in chunks that fit


def _dims(d_in, D, H, M, S, Cc, ln):
    return _lib.StDims(d_in=d_in, D=D, H=H, M=M, S=S, C=Cc, ln=ln)


@torch.library.custom_op("pcaudio::st_fwd", mutates_args=(), device_types="cuda")
def st_fwd(X: torch.Tensor, counts: torch.Tensor | None, params: torch.Tensor, d_in: int, D: int, H: int, M: int, S: int,
           n_out: int, ln: int, precision: int) -> torch.Tensor:
    """ST.forward before the .squeeze(): X (B, N, d_in) fp32, packed weights -> logits (B, S, n_out).
    counts (B,) int32 or None (variable-size sets)."""
    B, N, _ = X.shape
    dims = _dims(d_in, D, H, M, S, n_out, ln)
    L = _lib.lib()
    out = torch.empty((B, S, n_out), dtype=torch.float32, device=X.device)
    if B == 0:
        return out
    need1 = L.pca_st_workspace_bytes(C.byref(dims), 1, N, precision)
    needB = L.pca_st_workspace_bytes(C.byref(dims), B, N, precision)
    ws = rt.workspace(X.device, max(need1, min(needB, ST_WORKSPACE_BYTES)))
    with torch.cuda.device(X.device):
        _lib.check(L.pca_st_fwd_masked(_lib.ptr(X), _lib.ptr(counts), B, N, C.byref(dims), _lib.ptr(params), _lib.ptr(out),
                                       _lib.ptr(ws), ws.numel(), precision, rt.stream_ptr(X.device)), "st_fwd")
    return out


@st_fwd.register_fake
def _(X, counts, params, d_in, D, H, M, S, n_out, ln, precision):
    return X.new_empty((X.shape[0], S, n_out), dtype=torch.float32)


@torch.library.custom_op("pcaudio::stft_logmag", mutates_args=(), device_types="cuda")
def stft_logmag_op(audio: torch.Tensor, window: torch.Tensor, twiddle: torch.Tensor, n_fft: int, hop: int, scale: float,
                   drop_nyquist: bool, n_frames: int) -> torch.Tensor:
    """log(1e-8 + |STFT| * scale): audio (B, L) -> (B, n_frames, n_fft/2 + 1 - drop_nyquist), frequency fastest."""
    B, Ls = audio.shape
    nf = n_fft // 2 + 1 - (1 if drop_nyquist else 0)
    out = torch.empty((B, n_frames, nf), dtype=torch.float32, device=audio.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(audio.device):
        _lib.check(_lib.lib().pca_stft_logmag_f32(_lib.ptr(audio), B, Ls, n_fft, hop, _lib.ptr(window), _lib.ptr(twiddle),
                                                  scale, int(drop_nyquist), n_frames, _lib.ptr(out),
                                                  rt.stream_ptr(audio.device)), "stft_logmag")
    return out


@stft_logmag_op.register_fake
def _(audio, window, twiddle, n_fft, hop, scale, drop_nyquist, n_frames):
    nf = n_fft // 2 + 1 - (1 if drop_nyquist else 0)
    return audio.new_empty((audio.shape[0], n_frames, nf), dtype=torch.float32)


@torch.library.custom_op("pcaudio::select_points", mutates_args=(), device_types="cuda")
def select_points_op(logmag: torch.Tensor, farr: torch.Tensor, tarr: torch.Tensor | None, k: int, sorted_desc: bool,
                     use_threshold: bool, threshold: float) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Top-K / threshold selection into padded sets: logmag (n, nt, nf) -> (pts (n,k,2|3), idx (n,k) i32, counts (n,) i32)."""
    n, nt, nf = logmag.shape
    dev = logmag.device
    width = 2 if tarr is None else 3
    pts = torch.empty((n, k, width), dtype=torch.float32, device=dev)
    idx = torch.empty((n, k), dtype=torch.int32, device=dev)
    counts = torch.empty((n,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pca_select_compact_f32(_lib.ptr(logmag), n, nf, nt, _lib.ptr(farr), _lib.ptr(tarr), int(k),
                                                     int(sorted_desc), int(use_threshold), float(threshold), _lib.ptr(pts),
                                                     _lib.ptr(idx), _lib.ptr(counts), rt.stream_ptr(dev)), "select_points")
    return pts, idx, counts


@select_points_op.register_fake
def _(logmag, farr, tarr, k, sorted_desc, use_threshold, threshold):
    n = logmag.shape[0]
    width = 2 if tarr is None else 3
    return (logmag.new_empty((n, k, width), dtype=torch.float32), logmag.new_empty((n, k), dtype=torch.int32),
            logmag.new_empty((n,), dtype=torch.int32))


# ------------------------------------------------------------------------------------ training (fp32)
def _train_ws(dev, dims, B, N):
    return rt.workspace(dev, _lib.lib().pca_st_train_workspace_bytes(C.byref(dims), B, N))


@torch.library.custom_op("pcaudio::st_train_fwd", mutates_args=(), device_types="cuda")
def st_train_fwd(X: torch.Tensor, counts: torch.Tensor | None, params: torch.Tensor, d_in: int, D: int, H: int, M: int, S: int,
                 n_out: int, ln: int, dropout_p: float, seed: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Training forward of ST / SetTransformer: (logits (B, S, n_out), saved activations (bytes) for st_train_bwd)."""
    B, N, _ = X.shape
    dims = _dims(d_in, D, H, M, S, n_out, ln)
    L = _lib.lib()
    logits = torch.empty((B, S, n_out), dtype=torch.float32, device=X.device)
    saved = torch.empty(max(1, L.pca_st_train_saved_bytes(C.byref(dims), B, N, dropout_p)), dtype=torch.uint8, device=X.device)
    ws = _train_ws(X.device, dims, B, N)
    with torch.cuda.device(X.device):
        _lib.check(L.pca_st_train_fwd_f32(_lib.ptr(X), _lib.ptr(counts), B, N, C.byref(dims), _lib.ptr(params), dropout_p, seed, _lib.ptr(logits),
                                          _lib.ptr(saved), saved.numel(), _lib.ptr(ws), ws.numel(), rt.stream_ptr(X.device)),
                   "st_train_fwd")
    return logits, saved


@st_train_fwd.register_fake
def _(X, counts, params, d_in, D, H, M, S, n_out, ln, dropout_p, seed):
    return X.new_empty((X.shape[0], S, n_out), dtype=torch.float32), X.new_empty((1,), dtype=torch.uint8)


@torch.library.custom_op("pcaudio::st_train_bwd", mutates_args=(), device_types="cuda")
def st_train_bwd(X: torch.Tensor, counts: torch.Tensor | None, params: torch.Tensor, d_in: int, D: int, H: int, M: int, S: int,
                 n_out: int, ln: int, dropout_p: float, seed: int, dlogits: torch.Tensor, saved: torch.Tensor,
                 need_dx: bool) -> tuple[torch.Tensor, torch.Tensor]:
    """Backward of st_train_fwd: (flat gradient of every parameter in the layout of `params`, dX or an empty tensor)."""
    B, N, _ = X.shape
    dims = _dims(d_in, D, H, M, S, n_out, ln)
    L = _lib.lib()
    dparams = torch.empty_like(params)
    dX = torch.empty_like(X) if need_dx else X.new_empty((0,))
    ws = _train_ws(X.device, dims, B, N)
    with torch.cuda.device(X.device):
        _lib.check(L.pca_st_train_bwd_f32(_lib.ptr(X), _lib.ptr(counts), B, N, C.byref(dims), _lib.ptr(params), dropout_p, seed, _lib.ptr(dlogits),
                                          _lib.ptr(saved), saved.numel(), _lib.ptr(dparams), _lib.ptr(dX) if need_dx else None,
                                          _lib.ptr(ws), ws.numel(), rt.stream_ptr(X.device)), "st_train_bwd")
    return dparams, dX


@st_train_bwd.register_fake
def _(X, counts, params, d_in, D, H, M, S, n_out, ln, dropout_p, seed, dlogits, saved, need_dx):
    return torch.empty_like(params), (torch.empty_like(X) if need_dx else X.new_empty((0,)))
