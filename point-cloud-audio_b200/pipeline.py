"""Whole hot path in one call: raw audio -> STFT/log-magnitude -> point clouds [-> top-K] -> ST logits.

Mirrors what the reference's eval loops do per batch (Code/pceval.py:73-99 for FST,
Code/pc_temp3d_eval.py:126-185 for 3ST) without the file I/O, behind ``pca_pipeline_run`` /
``pca_pipeline_run_host`` of the C ABI."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib, _runtime as rt
from .frontend import coord_tables


@dataclass
class AudioConfig:
    """Mirrors the keys of the reference's *_config.json (Code/settransformer.py:134-151)."""
    sampling_rate: float = 16000.0
    window_size: int = 2048          # Nfft
    hop_factor: float = 0.5
    n_samples: int = 16000
    mode: int = 2                    # 2: FST frame clouds (f, mag); 3: 3ST chunk clouds (f, t, mag)
    Ntemp: int = 10                  # mode 3: frames per cloud
    top_k: int = 0                   # 0 = keep all points
    precision: str = "fp32"          # encoder precision: 'fp32' | 'bf16'
    threshold: float | None = None   # keep only points with log-magnitude >= threshold (padded, masked sets)


class AudioSetPipeline:
    """Device-resident pipeline bound to a set model (``ST``).  ``__call__(audio_cuda)`` returns logits
    (n_clips * clouds_per_clip, C) on the device; ``run_host(pinned_audio)`` is the end-to-end call with
    host buffers (H2D copy, kernels, D2H copy on the current stream)."""

    def __init__(self, model, cfg: AudioConfig, device="cuda"):
        self.model = model
        self.cfg = cfg
        self.device = torch.device(device)
        dims = model._dims()
        hop = int(cfg.window_size * cfg.hop_factor)
        self.c = _lib.PipelineCfg(n_samples=cfg.n_samples, n_fft=cfg.window_size, hop=hop,
                                  scale=1.0 / cfg.window_size, mode=cfg.mode, ntemp=cfg.Ntemp, top_k=cfg.top_k,
                                  precision={"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}[cfg.precision], st=dims,
                                  use_threshold=int(cfg.threshold is not None),
                                  threshold=float(cfg.threshold if cfg.threshold is not None else 0.0))
        L = _lib.lib()
        self.clouds_per_clip = L.pca_pipeline_clouds_per_clip(C.byref(self.c))
        self.points_per_cloud = L.pca_pipeline_points_per_cloud(C.byref(self.c))
        if self.clouds_per_clip < 0:
            _lib.check(-1, "AudioSetPipeline")
        nf = cfg.window_size // 2 + (1 if cfg.mode == 2 else 0)
        farr, tarr = coord_tables(cfg.sampling_rate, nf, cfg.window_size, cfg.hop_factor,
                                  cfg.Ntemp if cfg.mode == 3 else None)
        self.farr = rt.coord_table(farr, self.device)
        self.tarr = None if tarr is None else rt.coord_table(tarr, self.device)
        self.window, self.twiddle = rt.stft_tables(cfg.window_size, cfg.window_size, self.device)
        self._staging = {}
        self._copy_stream = None
        self._pipe = None

    def _ws(self, n_clips):
        need = _lib.lib().pca_pipeline_workspace_bytes(C.byref(self.c), n_clips)
        return rt.workspace(self.device, need)

    def __call__(self, audio: torch.Tensor) -> torch.Tensor:
        rt.require_cuda(audio, "AudioSetPipeline")
        audio = rt.f32c(audio)
        n_clips, L_ = audio.shape
        if L_ != self.cfg.n_samples:
            raise ValueError(f"clips must have {self.cfg.n_samples} samples, got {L_}")
        st = self.c.st
        out = torch.empty((n_clips * self.clouds_per_clip, st.S, st.C), dtype=torch.float32, device=self.device)
        ws = self._ws(n_clips)
        blob = self.model._blob()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().pca_pipeline_run(
                C.byref(self.c), _lib.ptr(audio), n_clips, _lib.ptr(self.window), _lib.ptr(self.twiddle),
                _lib.ptr(self.farr), _lib.ptr(self.tarr), _lib.ptr(blob), _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                rt.stream_ptr(self.device)), "pipeline_run")
        return out.squeeze(1) if st.S == 1 else out

    def run_host(self, host_audio: torch.Tensor, host_logits: torch.Tensor | None = None, chunks: int | None = None) -> torch.Tensor:
        """host_audio: (n_clips, n_samples) float32 CPU tensor (pin it for async copies).  Returns the
        host logits tensor; the work is enqueued on the current stream -- synchronise before reading.
        With chunks > 1 the host->device copy of clip range k+1 overlaps the kernels of range k (a private copy
        stream of this pipeline object); default 1: measured on B200, the extra launches
        of a second range cost as much as the overlap saves at the bench shape (tools/e2e_chunks.py)."""
        if host_audio.is_cuda or host_audio.dtype != torch.float32 or not host_audio.is_contiguous():
            raise ValueError("host_audio must be a contiguous float32 CPU tensor")
        n_clips = host_audio.shape[0]
        st = self.c.st
        n_out = n_clips * self.clouds_per_clip
        key = n_clips
        if key not in self._staging:
            self._staging[key] = (torch.empty((n_clips, self.cfg.n_samples), dtype=torch.float32, device=self.device),
                                  torch.empty((n_out, st.S, st.C), dtype=torch.float32, device=self.device))
        dev_audio, dev_logits = self._staging[key]
        if host_logits is None:
            host_logits = torch.empty((n_out, st.S, st.C), dtype=torch.float32).pin_memory()
        ws = self._ws(n_clips)
        blob = self.model._blob()
        if chunks is None:
            chunks = 1
        with torch.cuda.device(self.device):
            if chunks > 1:
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(device=self.device)
                _lib.check(_lib.lib().pca_pipeline_run_host_chunked(
                    C.byref(self.c), _lib.ptr(host_audio), n_clips, _lib.ptr(dev_audio), _lib.ptr(self.window),
                    _lib.ptr(self.twiddle), _lib.ptr(self.farr), _lib.ptr(self.tarr), _lib.ptr(blob),
                    _lib.ptr(dev_logits), _lib.ptr(host_logits), _lib.ptr(ws), ws.numel(), int(chunks),
                    C.c_void_p(self._copy_stream.cuda_stream), rt.stream_ptr(self.device)), "pipeline_run_host_chunked")
            else:
                _lib.check(_lib.lib().pca_pipeline_run_host(
                    C.byref(self.c), _lib.ptr(host_audio), n_clips, _lib.ptr(dev_audio), _lib.ptr(self.window),
                    _lib.ptr(self.twiddle), _lib.ptr(self.farr), _lib.ptr(self.tarr), _lib.ptr(blob),
                    _lib.ptr(dev_logits), _lib.ptr(host_logits), _lib.ptr(ws), ws.numel(),
                    rt.stream_ptr(self.device)), "pipeline_run_host")
        return host_logits

    # ---- pipelined host-buffer interface: the H2D copy of batch k+1 overlaps the kernels of batch k ----------------
    def submit_host(self, host_audio: torch.Tensor, host_logits: torch.Tensor):
        """Enqueue one whole-path pass over a pinned host batch and return a ticket; ``wait_host(ticket)`` blocks until
        ``host_logits`` (pinned, (n_clips * clouds_per_clip, S, C)) holds its result.  Two staging slots: the copy of
        batch k+1 runs on a private copy stream while batch k computes on the current stream, so a loop
        ``t1 = submit(b1); wait(t0); t2 = submit(b2); wait(t1); ...`` streams batches at the speed of the slower of
        {copy, compute}.  Every batch is still copied host->device and its logits device->host."""
        if host_audio.is_cuda or host_audio.dtype != torch.float32 or not host_audio.is_contiguous():
            raise ValueError("host_audio must be a contiguous float32 CPU tensor")
        n_clips = host_audio.shape[0]
        st = self.c.st
        n_out = n_clips * self.clouds_per_clip
        if self._pipe is None or self._pipe["n_clips"] != n_clips:
            self._pipe = {"n_clips": n_clips, "k": 0,
                          "audio": [torch.empty((n_clips, self.cfg.n_samples), dtype=torch.float32, device=self.device) for _ in range(2)],
                          "logits": [torch.empty((n_out, st.S, st.C), dtype=torch.float32, device=self.device) for _ in range(2)],
                          "copied": [torch.cuda.Event() for _ in range(2)], "done": [None, None]}
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
        P = self._pipe
        slot = P["k"] & 1
        P["k"] += 1
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            with torch.cuda.stream(self._copy_stream):
                if P["done"][slot] is not None:
                    self._copy_stream.wait_event(P["done"][slot])        # the slot's previous batch has been consumed
                P["audio"][slot].copy_(host_audio, non_blocking=True)
                P["copied"][slot].record(self._copy_stream)
            cur.wait_event(P["copied"][slot])
            ws = self._ws(n_clips)
            blob = self.model._blob()
            _lib.check(_lib.lib().pca_pipeline_run(
                C.byref(self.c), _lib.ptr(P["audio"][slot]), n_clips, _lib.ptr(self.window), _lib.ptr(self.twiddle),
                _lib.ptr(self.farr), _lib.ptr(self.tarr), _lib.ptr(blob), _lib.ptr(P["logits"][slot]), _lib.ptr(ws), ws.numel(),
                rt.stream_ptr(self.device)), "pipeline_run")
            host_logits.copy_(P["logits"][slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cur)
            P["done"][slot] = ev
        return ev

    @staticmethod
    def wait_host(ticket) -> None:
        ticket.synchronize()

    @property
    def h2d_bytes_per_clip(self) -> int:
        return self.cfg.n_samples * 4

    @property
    def d2h_bytes_per_clip(self) -> int:
        return self.clouds_per_clip * self.c.st.S * self.c.st.C * 4
