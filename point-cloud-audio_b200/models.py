"""Drop-in set models: ``ST`` (Code/models.py:13-44), the ModelNet ``SetTransformer``
(set_transformer-master/main_pointcloud.py:13-37) and ``DeepSet``
(set_transformer-master/models.py:3-28).  Same constructor signatures, module trees and state-dict
keys as the reference (shipped checkpoints load directly, with or without the DataParallel
``module.`` prefix); ``forward`` is ONE call into the fused CUDA path."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib, _runtime as rt
from . import ops as _ops  # noqa: F401  (registers torch.ops.pcaudio.*)
from .modules import ISAB, PMA, SAB, _PackedParams, _guard, _mab_tensors



def strip_module_prefix(state_dict):
    """Checkpoints saved from nn.DataParallel carry a 'module.' prefix (Code/settransformer.py:94,160)."""
    return {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}


class _SetEncoderBase(nn.Module):
    """ISAB, ISAB -> PMA -> Linear, run through pca_st_fwd."""

    precision = _lib.PREC_FP32

    def _parts(self):
        raise NotImplementedError

    def set_precision(self, precision: str):
        """Inference precision: 'fp32' (1e-3 parity path) or 'bf16' (tcgen05 tensor-core tiles, 2e-2 parity).  Calls made
        with gradients enabled always take the fp32 training path, whatever is set here."""
        self.precision = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}[precision]
        return self

    def _dims(self):
        isab0, isab1, pma, lin = self._parts()
        return _lib.StDims(d_in=isab0.mab0.fc_k.in_features, D=isab0.mab0.dim_V, H=isab0.mab0.num_heads,
                           M=isab0.I.shape[1], S=pma.S.shape[1], C=lin.out_features, ln=isab0.mab0._ln)

    def _blob(self):
        ts = self._param_tensors()
        flat = getattr(self, "_flat", None)
        if flat is not None:
            # parameters are views of the flat buffer (flatten_parameters): use it directly while that still holds
            off, ok = flat.data_ptr(), True
            for t in ts:
                ok = ok and t.data_ptr() == off and t.dtype == torch.float32
                off += 4 * t.numel()
            if ok:
                return flat
            object.__setattr__(self, "_flat", None)
        if not hasattr(self, "_packed"):
            object.__setattr__(self, "_packed", _PackedParams())
        return self._packed.get(ts)

    def load_state_dict(self, state_dict, *args, **kwargs):
        return super().load_state_dict(strip_module_prefix(state_dict), *args, **kwargs)

    def _param_tensors(self):
        isab0, isab1, pma, lin = self._parts()
        return isab0._tensors() + isab1._tensors() + pma._tensors() + [lin.weight, lin.bias]

    def _dropout_p(self) -> float:
        return 0.0

    def flatten_parameters(self) -> torch.Tensor:
        """Re-point every parameter at a view of ONE flat fp32 buffer in the blob order of include/pcaudio_b200.h, so
        that the packed weights, the flat gradient (allreduce bucket) and the fused optimizer all address the same
        memory and nothing is re-packed per step.  Returns the flat buffer (state-dict keys are unchanged)."""
        ts = self._param_tensors()
        with torch.no_grad():
            flat = torch.cat([t.detach().reshape(-1).float() for t in ts]).contiguous()
            off = 0
            for t in ts:
                n = t.numel()
                t.data = flat[off:off + n].view(t.shape)
                off += n
        object.__setattr__(self, "_flat", flat)
        return flat

    def encode(self, X: torch.Tensor, counts: torch.Tensor | None = None) -> torch.Tensor:
        """(B, N, d_in) CUDA -> logits (B, S, C) (before the reference's .squeeze()).
        ``counts`` (B,) int32 CUDA, optional extension: cloud b consists of its first counts[b] rows, the rest is
        padding (variable-size sets); the result equals the reference module applied to X[b:b+1, :counts[b]].  A cloud
        with counts[b] <= 0 has no defined encoding (the reference fails on an empty set): its logits are NaN."""
        rt.require_cuda(X, type(self).__name__ + ".forward")
        X = rt.f32c(X)
        B, N, d_in = X.shape
        if counts is not None:
            rt.require_cuda(counts, type(self).__name__ + ".forward(counts)")
            if counts.shape != (B,):
                raise ValueError(f"counts must have shape ({B},), got {tuple(counts.shape)}")
            counts = counts.to(torch.int32).contiguous()
        dims = self._dims()
        if d_in != dims.d_in:
            raise ValueError(f"expected clouds of width {dims.d_in}, got {d_in}")
        blob = self._blob()
        assert blob.numel() == _lib.lib().pca_st_param_count(C.byref(dims))
        ps = self._param_tensors()
        if B > 0 and torch.is_grad_enabled() and (X.requires_grad or any(p.requires_grad for p in ps)):
            # training: fp32 forward that keeps activations + hand-written backward (pcaudio_b200/training.py)
            if dims.ln:
                # LayerNorm variant: composed from the block training kernels (the fused whole-model path covers ln=False)
                if counts is not None:
                    raise NotImplementedError("pcaudio_b200: training with LayerNorm and variable-size sets is not implemented")
                from .training import LinearFunction, dropout
                isab0, isab1, pma, lin = self._parts()
                p = self._dropout_p()
                Y = pma(dropout(isab1(isab0(X)), p))
                return LinearFunction.apply(dropout(Y, p), lin.weight, lin.bias)
            from .training import STTrainFunction
            p = self._dropout_p()
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0
            return STTrainFunction.apply(X, counts, blob, dims, p, seed, *ps)
        if self._dropout_p() > 0:
            raise RuntimeError("SetTransformer: train-mode (Dropout) forward without gradients; call .eval() for inference")
        # one dispatcher-visible custom op (pcaudio_b200/ops.py) -> pca_st_fwd_masked of the C ABI
        out = torch.ops.pcaudio.st_fwd(X, counts, blob, dims.d_in, dims.D, dims.H, dims.M, dims.S, dims.C, dims.ln, self.precision)
        return _guard(out, self)


class ST(_SetEncoderBase):
    """Set Transformer for 2-D / 3-D spectral point clouds (Code/models.py:13-44)."""

    def __init__(self, dim_input=2, num_outputs=1, dim_output=10, num_inds=4, dim_hidden=4, num_heads=2, ln=False):
        super().__init__()
        self.enc = nn.Sequential(
            ISAB(dim_input, dim_hidden, num_heads, num_inds, ln=ln),
            ISAB(dim_hidden, dim_hidden, num_heads, num_inds, ln=ln),
        )
        self.dec = nn.Sequential(
            PMA(dim_hidden, num_heads, num_outputs, ln=ln),
            nn.Linear(dim_hidden, dim_output),
        )

    def _parts(self):
        return self.enc[0], self.enc[1], self.dec[0], self.dec[1]

    def forward(self, X, counts=None):
        return self.encode(X, counts).squeeze()       # (B,1,C)->(B,C); (C,) when B == 1 (Code/models.py:44)


class SetTransformer(_SetEncoderBase):
    """ModelNet40 classifier of set_transformer-master/main_pointcloud.py:13-37: ST with Dropout
    around the PMA.  Eval-mode forward runs the fused inference path; with gradients enabled the
    training path runs (dropout masks from a counter-based generator, regenerated in the backward pass)."""

    def __init__(self, dim_input=3, num_outputs=1, dim_output=40, num_inds=32, dim_hidden=128, num_heads=4, ln=False):
        super().__init__()
        self.enc = nn.Sequential(
            ISAB(dim_input, dim_hidden, num_heads, num_inds, ln=ln),
            ISAB(dim_hidden, dim_hidden, num_heads, num_inds, ln=ln),
        )
        self.dec = nn.Sequential(
            nn.Dropout(),
            PMA(dim_hidden, num_heads, num_outputs, ln=ln),
            nn.Dropout(),
            nn.Linear(dim_hidden, dim_output),
        )

    def _parts(self):
        return self.enc[0], self.enc[1], self.dec[1], self.dec[3]

    def _dropout_p(self) -> float:
        return float(self.dec[0].p) if self.training else 0.0

    def forward(self, X, counts=None):
        return self.encode(X, counts).squeeze()


class DeepSet(nn.Module):
    """Shared 4-layer MLP over points -> pool -> 4-layer decoder (set_transformer-master/models.py:3-28).
    ``pool`` extends the reference's mean with SmallDeepSet's max / sum (max_regression_demo.ipynb:41-48)."""

    def __init__(self, dim_input, num_outputs, dim_output, dim_hidden=128, pool="mean"):
        super().__init__()
        self.num_outputs = num_outputs
        self.dim_output = dim_output
        self.pool = pool
        self.enc = nn.Sequential(
            nn.Linear(dim_input, dim_hidden), nn.ReLU(),
            nn.Linear(dim_hidden, dim_hidden), nn.ReLU(),
            nn.Linear(dim_hidden, dim_hidden), nn.ReLU(),
            nn.Linear(dim_hidden, dim_hidden))
        self.dec = nn.Sequential(
            nn.Linear(dim_hidden, dim_hidden), nn.ReLU(),
            nn.Linear(dim_hidden, dim_hidden), nn.ReLU(),
            nn.Linear(dim_hidden, dim_hidden), nn.ReLU(),
            nn.Linear(dim_hidden, num_outputs * dim_output))
        self._packed = _PackedParams()

    def forward(self, X, counts=None):
        """``counts`` (B,) int32 CUDA, optional: pool only over the first counts[b] points (masked mean / max / sum)."""
        rt.require_cuda(X, "DeepSet.forward")
        X = rt.f32c(X)
        B, N, d_in = X.shape
        if counts is not None:
            rt.require_cuda(counts, "DeepSet.forward(counts)")
            if counts.shape != (B,):
                raise ValueError(f"counts must have shape ({B},), got {tuple(counts.shape)}")
            counts = counts.to(torch.int32).contiguous()
        dh = self.enc[0].out_features
        out_dim = self.num_outputs * self.dim_output
        ts = []
        for seq in (self.enc, self.dec):
            for i in (0, 2, 4, 6):
                ts += [seq[i].weight, seq[i].bias]
        blob = self._packed.get(ts)
        pool = {"mean": 0, "max": 1, "sum": 2}[self.pool]
        if B > 0 and counts is None and torch.is_grad_enabled() and (X.requires_grad or any(p.requires_grad for p in ts)):
            # training: fp32 forward that keeps activations + hand-written backward (pcaudio_b200/training.py)
            from .training import DeepSetTrainFunction
            out = DeepSetTrainFunction.apply(X, blob, (d_in, dh, out_dim, pool), *ts)
            return out.reshape(-1, self.num_outputs, self.dim_output)
        out = torch.empty((B, out_dim), dtype=torch.float32, device=X.device)
        L = _lib.lib()
        ws = rt.workspace(X.device, L.pca_deepset_workspace_bytes(B, N, d_in, dh, out_dim))
        with torch.cuda.device(X.device):
            _lib.check(L.pca_deepset_fwd_masked_f32(_lib.ptr(X), _lib.ptr(counts), B, N, d_in, dh, out_dim, pool, _lib.ptr(blob),
                                                    _lib.ptr(out), _lib.ptr(ws), ws.numel(), rt.stream_ptr(X.device)),
                       "DeepSet.forward")
        return _guard(out.reshape(-1, self.num_outputs, self.dim_output), self)


class SetTransformerSAB(nn.Module):
    """The generic ``SetTransformer`` of set_transformer-master/models.py:30-44 (the clustering model): ISAB, ISAB ->
    PMA(num_outputs seeds) -> SAB, SAB -> Linear, optional LayerNorm.  Same constructor signature, module tree and
    state-dict keys as the reference class (which shares its name with main_pointcloud.SetTransformer, hence the suffix
    here; ``pcaudio_b200.st_models.SetTransformer`` is the same class under the reference's name).  Forward runs the fp32
    per-block CUDA kernels (every dim, ``ln`` included); output (B, num_outputs, dim_output), no squeeze."""

    def __init__(self, dim_input, num_outputs, dim_output, num_inds=32, dim_hidden=128, num_heads=4, ln=False):
        super().__init__()
        self.enc = nn.Sequential(
            ISAB(dim_input, dim_hidden, num_heads, num_inds, ln=ln),
            ISAB(dim_hidden, dim_hidden, num_heads, num_inds, ln=ln))
        self.dec = nn.Sequential(
            PMA(dim_hidden, num_heads, num_outputs, ln=ln),
            SAB(dim_hidden, dim_hidden, num_heads, ln=ln),
            SAB(dim_hidden, dim_hidden, num_heads, ln=ln),
            nn.Linear(dim_hidden, dim_output))
        self._packed = _PackedParams()

    def load_state_dict(self, state_dict, *args, **kwargs):
        return super().load_state_dict(strip_module_prefix(state_dict), *args, **kwargs)

    def forward(self, X):
        rt.require_cuda(X, "SetTransformerSAB.forward")
        Y = self.dec[2](self.dec[1](self.dec[0](self.enc(rt.f32c(X)))))             # ISAB, ISAB, PMA, SAB, SAB
        lin = self.dec[3]
        B, S, D = Y.shape
        if Y.requires_grad:
            # training: the blocks above ran their training kernels and autograd composes them; the last layer runs the
            # repo's own Linear forward / backward kernels as well
            from .training import LinearFunction
            return LinearFunction.apply(Y, lin.weight, lin.bias)
        out = torch.empty((B, S, lin.out_features), dtype=torch.float32, device=Y.device)
        blob = self._packed.get([lin.weight, lin.bias])
        with torch.cuda.device(Y.device):
            _lib.check(_lib.lib().pca_linear_fwd_f32(_lib.ptr(Y), B * S, D, lin.out_features, _lib.ptr(blob), _lib.ptr(out),
                                                    rt.stream_ptr(Y.device)), "SetTransformerSAB.forward(Linear)")
        return _guard(out, self)
