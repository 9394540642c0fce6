"""Device-side plumbing shared by the host mirrors: stream handles, cached scratch buffers and
constant tables.  PyTorch is used only for memory, streams and (elsewhere) torch.distributed."""
from __future__ import annotations

import numpy as np
import torch

_ws_cache: dict = {}
_table_cache: dict = {}


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"pcaudio_b200.{what}: expected a CUDA tensor (this path has no CPU fallback)")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def workspace(device, nbytes: int) -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream); kernels on one stream are ordered, so reuse
    across calls is safe."""
    key = (torch.device(device).index, stream_ptr(device))
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _ws_cache.pop(key, None)
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def stft_tables(n_fft: int, win_length: int, device):
    """(window, twiddle) float32 device tables, built once in float64 on the host.

    window: periodic Hann of win_length, centred / zero padded to n_fft (librosa 0.8.0
    get_window(fftbins=True) + util.pad_center, used by Code/settransformer.py:49 and
    Code/pceval.py:76); twiddle[k] = (cos, -sin)(2 pi k / n_fft)."""
    key = ("stft", n_fft, win_length, torch.device(device).index)
    hit = _table_cache.get(key)
    if hit is None:
        n = np.arange(win_length, dtype=np.float64)
        w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)
        win = np.zeros(n_fft, dtype=np.float64)
        lpad = (n_fft - win_length) // 2
        win[lpad:lpad + win_length] = w
        k = np.arange(n_fft // 2, dtype=np.float64)
        tw = np.stack([np.cos(2 * np.pi * k / n_fft), -np.sin(2 * np.pi * k / n_fft)], axis=1)
        hit = (torch.from_numpy(win.astype(np.float32)).to(device),
               torch.from_numpy(tw.astype(np.float32)).contiguous().to(device))
        _table_cache[key] = hit
    return hit


def coord_table(arr, device) -> torch.Tensor:
    """float64 coordinate vector (farr / tarr) -> float32 device table, rounded exactly once as the
    reference's .float() cast does (Code/dataset.py:54,166)."""
    a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64))
    return torch.from_numpy(a.astype(np.float32)).to(device)
