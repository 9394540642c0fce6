"""Drop-in attention blocks with the constructor signatures, parameter names/shapes and forward
contracts of set_transformer-master/modules.py (MAB :6-33, SAB :35-41, ISAB :43-53, PMA :55-63).
State dicts are interchangeable with the reference.  The forward pass runs the CUDA kernels through
the C ABI; there is no PyTorch-op fallback (CPU tensors raise)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib, _runtime as rt


def _wants_grad(module, *inputs, tensors=None):
    """Training path?  Decided from the parameter TENSORS the forward is about to use (``tensors``), not from
    ``module.parameters()``: nn.DataParallel replicas hold plain attribute tensors and an empty parameter list."""
    if not torch.is_grad_enabled():
        return False
    if any(isinstance(t, torch.Tensor) and t.requires_grad for t in inputs):
        return True
    ps = tensors if tensors is not None else list(module.parameters())
    return any(p.requires_grad for p in ps)


class _NoBackward(torch.autograd.Function):
    """Marks a fused forward output for which no backward kernel exists so that a backward pass fails loudly instead of
    silently training with zero gradients (every block and model has training kernels; this guards future gaps)."""

    @staticmethod
    def forward(ctx, out, *params):
        return out.view_as(out)

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError("pcaudio_b200: no backward kernel for this configuration; "
                                  "run inference under torch.no_grad() or detach the output")


def _guard(out, module):
    if torch.is_grad_enabled():
        ps = [p for p in module.parameters() if p.requires_grad]
        if ps:
            return _NoBackward.apply(out, *ps)
    return out


class _PackedParams:
    """Flattens a module's parameters into the canonical blob of include/pcaudio_b200.h and caches it per device until a
    parameter changes (in-place update through autograd-visible ops, load_state_dict, .to()).

    nn.DataParallel replicas share this object (replicate() shallow-copies ``__dict__``) and call ``get`` from one thread
    per GPU: the cache is therefore keyed by device, every call works on local variables only, and a lost race merely
    rebuilds a blob.  Edits that bypass the version counter (``p.data.mul_()``, ``p.data.clamp_()``) are NOT seen -- call
    ``invalidate()`` (or ``module.invalidate_packed()``) after such an edit."""

    def __init__(self):
        self._cache = {}

    def invalidate(self):
        self._cache = {}

    def get(self, tensors):
        dev = tensors[0].device
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        hit = self._cache.get(dev)
        if hit is not None and hit[0] == key:
            return hit[1]
        with torch.no_grad():
            blob = torch.cat([t.detach().reshape(-1).float() for t in tensors]).contiguous()
        self._cache[dev] = (key, blob)
        return blob


def invalidate_packed(module: nn.Module) -> None:
    """Drop every cached parameter blob below ``module`` (needed only after edits through ``.data`` that do not bump the
    tensors' version counters)."""
    for m in module.modules():
        pk = getattr(m, "_packed", None)
        if isinstance(pk, _PackedParams):
            pk.invalidate()


def _mab_tensors(m: "MAB"):
    ts = [m.fc_q.weight, m.fc_q.bias, m.fc_k.weight, m.fc_v.weight, m.fc_k.bias, m.fc_v.bias,
          m.fc_o.weight, m.fc_o.bias]
    if getattr(m, "ln0", None) is not None:
        ts += [m.ln0.weight, m.ln0.bias, m.ln1.weight, m.ln1.bias]
    return ts


class MAB(nn.Module):
    def __init__(self, dim_Q, dim_K, dim_V, num_heads, ln=False):
        super().__init__()
        self.dim_V = dim_V
        self.num_heads = num_heads
        self.fc_q = nn.Linear(dim_Q, dim_V)
        self.fc_k = nn.Linear(dim_K, dim_V)
        self.fc_v = nn.Linear(dim_K, dim_V)
        if ln:
            self.ln0 = nn.LayerNorm(dim_V)
            self.ln1 = nn.LayerNorm(dim_V)
        self.fc_o = nn.Linear(dim_V, dim_V)
        self._packed = _PackedParams()

    @property
    def _ln(self):
        return int(getattr(self, "ln0", None) is not None)

    def forward(self, Q, K):
        rt.require_cuda(Q, "MAB.forward")
        rt.require_cuda(K, "MAB.forward")
        Q, K = rt.f32c(Q), rt.f32c(K)
        B, nk, dk = K.shape
        qb, nq, dq = Q.shape
        D, H = self.dim_V, self.num_heads
        blob = self._packed.get(_mab_tensors(self))
        if B > 0 and _wants_grad(self, Q, K, tensors=_mab_tensors(self)):
            # training: forward that keeps activations + hand-written backward (pcaudio_b200/training.py)
            from .training import MABTrainFunction
            return MABTrainFunction.apply(Q, K, blob, (D, H, self._ln), *_mab_tensors(self))
        out = torch.empty((B, nq, D), dtype=torch.float32, device=K.device)
        L = _lib.lib()
        ws = rt.workspace(K.device, L.pca_mab_workspace_bytes(B, nq, nk, dq, dk, D, H))
        with torch.cuda.device(K.device):
            _lib.check(L.pca_mab_fwd_f32(_lib.ptr(Q), qb, _lib.ptr(K), B, nq, nk, dq, dk, D, H, self._ln,
                                         _lib.ptr(blob), _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                         rt.stream_ptr(K.device)), "MAB.forward")
        return _guard(out, self)


class SAB(nn.Module):
    def __init__(self, dim_in, dim_out, num_heads, ln=False):
        super().__init__()
        self.mab = MAB(dim_in, dim_in, dim_out, num_heads, ln=ln)

    def forward(self, X):
        return self.mab(X, X)


class ISAB(nn.Module):
    def __init__(self, dim_in, dim_out, num_heads, num_inds, ln=False):
        super().__init__()
        self.I = nn.Parameter(torch.Tensor(1, num_inds, dim_out))
        nn.init.xavier_uniform_(self.I)
        self.mab0 = MAB(dim_out, dim_in, dim_out, num_heads, ln=ln)
        self.mab1 = MAB(dim_in, dim_out, dim_out, num_heads, ln=ln)
        self._packed = _PackedParams()

    def _tensors(self):
        return [self.I] + _mab_tensors(self.mab0) + _mab_tensors(self.mab1)

    def forward(self, X):
        rt.require_cuda(X, "ISAB.forward")
        X = rt.f32c(X)
        B, N, d_in = X.shape
        D, H, M = self.mab0.dim_V, self.mab0.num_heads, self.I.shape[1]
        if B > 0 and _wants_grad(self, X, tensors=self._tensors()):
            # training: H = mab0(I, X); mab1(X, H) through the MAB training kernels (I is a shared query set: no repeat)
            return self.mab1(X, self.mab0(self.I, X))
        blob = self._packed.get(self._tensors())
        out = torch.empty((B, N, D), dtype=torch.float32, device=X.device)
        L = _lib.lib()
        ws = rt.workspace(X.device, L.pca_isab_workspace_bytes(B, N, d_in, D, H, M))
        with torch.cuda.device(X.device):
            _lib.check(L.pca_isab_fwd_f32(_lib.ptr(X), B, N, d_in, D, H, M, self.mab0._ln, _lib.ptr(blob),
                                          _lib.ptr(out), _lib.ptr(ws), ws.numel(), rt.stream_ptr(X.device)),
                       "ISAB.forward")
        return _guard(out, self)


class PMA(nn.Module):
    def __init__(self, dim, num_heads, num_seeds, ln=False):
        super().__init__()
        self.S = nn.Parameter(torch.Tensor(1, num_seeds, dim))
        nn.init.xavier_uniform_(self.S)
        self.mab = MAB(dim, dim, dim, num_heads, ln=ln)
        self._packed = _PackedParams()

    def _tensors(self):
        return [self.S] + _mab_tensors(self.mab)

    def forward(self, X):
        rt.require_cuda(X, "PMA.forward")
        X = rt.f32c(X)
        B, N, D = X.shape
        H, S = self.mab.num_heads, self.S.shape[1]
        if B > 0 and _wants_grad(self, X, tensors=self._tensors()):
            return self.mab(self.S, X)                     # training: MAB(S, X), shared seeds
        blob = self._packed.get(self._tensors())
        out = torch.empty((B, S, D), dtype=torch.float32, device=X.device)
        L = _lib.lib()
        ws = rt.workspace(X.device, L.pca_pma_workspace_bytes(B, N, D, H, S))
        with torch.cuda.device(X.device):
            _lib.check(L.pca_pma_fwd_f32(_lib.ptr(X), B, N, D, H, S, self.mab._ln, _lib.ptr(blob), _lib.ptr(out),
                                         _lib.ptr(ws), ws.numel(), rt.stream_ptr(X.device)), "PMA.forward")
        return _guard(out, self)
