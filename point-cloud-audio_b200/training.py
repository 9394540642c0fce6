"""Training of the set encoders (config 5 of BASELINE.json).

Two ways in, both on the same hand-written fp32 CUDA forward/backward kernels (csrc/encoder_train.cu):

* the reference's own loop works unchanged -- ``preds = model(X); loss = criterion(preds, y); loss.backward();
  optimizer.step()`` (Code/settransformer.py:101-109, main_pointcloud.py:71-79): with gradients enabled, ``ST`` /
  ``SetTransformer`` run through ``STTrainFunction``, which hands every parameter its gradient;
* ``SetTrainer.step(X, labels)`` is the fused B200-first step: forward, cross-entropy, backward, ONE flat-bucket
  gradient allreduce over NCCL (replacing nn.DataParallel's reduce_add + broadcast) and one fused Adam launch over the
  flat parameter blob, with no autograd graph and no host synchronisation.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib, _runtime as rt
from . import ops as _ops  # noqa: F401
from .parallel import reduce_flat_gradient_


class STTrainFunction(torch.autograd.Function):
    """logits = f(X, *params): forward keeps the activations, backward returns dX (if needed) and per-parameter views of
    the flat gradient blob."""

    @staticmethod
    def forward(ctx, X, counts, blob, dims, dropout_p, seed, *params):
        d = dims
        logits, saved = torch.ops.pcaudio.st_train_fwd(X, counts, blob, d.d_in, d.D, d.H, d.M, d.S, d.C, d.ln, dropout_p, seed)
        ctx.save_for_backward(X, blob, saved)
        ctx.counts = counts
        ctx.dims, ctx.dropout_p, ctx.seed = d, dropout_p, seed
        ctx.shapes = [tuple(p.shape) for p in params]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        X, blob, saved = ctx.saved_tensors
        d = ctx.dims
        need_dx = bool(ctx.needs_input_grad[0])
        dparams, dX = torch.ops.pcaudio.st_train_bwd(X, ctx.counts, blob, d.d_in, d.D, d.H, d.M, d.S, d.C, d.ln, ctx.dropout_p,
                                                     ctx.seed, rt.f32c(dlogits), saved, need_dx)
        grads, off = [], 0
        for shp in ctx.shapes:
            n = 1
            for s in shp:
                n *= s
            grads.append(dparams[off:off + n].view(shp))
            off += n
        assert off == dparams.numel()
        return (dX if need_dx else None, None, None, None, None, None, *grads)


class LinearFunction(torch.autograd.Function):
    """Y = X W^T + b on rows through pca_linear_fwd_f32 / pca_linear_bwd_f32 (the final nn.Linear of the composed models when
    gradients are enabled -- torch's own Linear would be a cuBLAS call on the product path)."""

    @staticmethod
    def forward(ctx, X, weight, bias):
        X = rt.f32c(X)
        dout, din = weight.shape
        rows = X.numel() // din
        with torch.no_grad():
            blob = torch.cat([weight.detach().reshape(-1).float(), bias.detach().reshape(-1).float()]).contiguous()
        Y = torch.empty(X.shape[:-1] + (dout,), dtype=torch.float32, device=X.device)
        with torch.cuda.device(X.device):
            _lib.check(_lib.lib().pca_linear_fwd_f32(_lib.ptr(X), rows, din, dout, _lib.ptr(blob), _lib.ptr(Y),
                                                    rt.stream_ptr(X.device)), "Linear.forward")
        ctx.save_for_backward(X, blob)
        ctx.dims = (rows, din, dout)
        return Y

    @staticmethod
    def backward(ctx, dY):
        X, blob = ctx.saved_tensors
        rows, din, dout = ctx.dims
        dY = rt.f32c(dY)
        need_dx = bool(ctx.needs_input_grad[0])
        dX = torch.empty_like(X) if need_dx else None
        dparams = torch.empty(dout * din + dout, dtype=torch.float32, device=X.device)
        with torch.cuda.device(X.device):
            _lib.check(_lib.lib().pca_linear_bwd_f32(_lib.ptr(dY), _lib.ptr(X), rows, din, dout, _lib.ptr(blob), _lib.ptr(dX),
                                                    _lib.ptr(dparams), rt.stream_ptr(X.device)), "Linear.backward")
        return dX, dparams[:dout * din].view(dout, din), dparams[dout * din:]


class DropoutFunction(torch.autograd.Function):
    """nn.Dropout (main_pointcloud.py:30,32) through pca_dropout_f32: the mask is a pure function of (seed, element index), so
    backward is the same call on the gradient and no mask tensor exists."""

    @staticmethod
    def forward(ctx, X, p, seed):
        X = rt.f32c(X)
        out = torch.empty_like(X)
        with torch.cuda.device(X.device):
            _lib.check(_lib.lib().pca_dropout_f32(_lib.ptr(X), _lib.ptr(out), X.numel(), float(p), int(seed), rt.stream_ptr(X.device)),
                       "Dropout.forward")
        ctx.p, ctx.seed = float(p), int(seed)
        return out

    @staticmethod
    def backward(ctx, g):
        g = rt.f32c(g)
        out = torch.empty_like(g)
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib().pca_dropout_f32(_lib.ptr(g), _lib.ptr(out), g.numel(), ctx.p, ctx.seed, rt.stream_ptr(g.device)),
                       "Dropout.backward")
        return out, None, None


def dropout(X, p):
    """Train-mode dropout with a fresh seed drawn from torch's global generator (so torch.manual_seed governs it)."""
    if p <= 0:
        return X
    return DropoutFunction.apply(X, p, int(torch.randint(0, 2 ** 62, (1,)).item()))


class MABTrainFunction(torch.autograd.Function):
    """out = MAB(Q, K; params) for models composed from the blocks (modules.py:6-33, LayerNorm branches included): Q (1 | B, nq, dq), K (B, nk, dk).
    A query batch of 1 is the shared-query case (ISAB's I, PMA's S): its gradient is summed over the batch in the kernel."""

    @staticmethod
    def forward(ctx, Q, K, blob, cfg, *params):
        D, H, ln = cfg
        qb, nq, dq = Q.shape
        B, nk, dk = K.shape
        L = _lib.lib()
        dev = K.device
        out = torch.empty((B, nq, D), dtype=torch.float32, device=dev)
        saved = torch.empty(max(1, L.pca_mab_train_saved_bytes(B, qb, nq, nk, D, H, ln)), dtype=torch.uint8, device=dev)
        ws = rt.workspace(dev, L.pca_mab_train_workspace_bytes(B, qb, nq, nk, D, H, ln))
        with torch.cuda.device(dev):
            _lib.check(L.pca_mab_train_fwd_f32(_lib.ptr(Q), qb, _lib.ptr(K), B, nq, nk, dq, dk, D, H, ln, _lib.ptr(blob), _lib.ptr(out),
                                               _lib.ptr(saved), saved.numel(), _lib.ptr(ws), ws.numel(), rt.stream_ptr(dev)),
                       "mab_train_fwd")
        ctx.save_for_backward(Q, K, blob, saved)
        ctx.cfg = cfg
        ctx.shapes = [tuple(p.shape) for p in params]
        return out

    @staticmethod
    def backward(ctx, dout):
        Q, K, blob, saved = ctx.saved_tensors
        D, H, ln = ctx.cfg
        qb, nq, dq = Q.shape
        B, nk, dk = K.shape
        L = _lib.lib()
        dev = K.device
        dparams = torch.empty_like(blob)
        dQ = torch.empty_like(Q) if ctx.needs_input_grad[0] else None
        dK = torch.empty_like(K) if ctx.needs_input_grad[1] else None
        ws = rt.workspace(dev, L.pca_mab_train_workspace_bytes(B, qb, nq, nk, D, H, ln))
        dout = rt.f32c(dout)
        with torch.cuda.device(dev):
            _lib.check(L.pca_mab_train_bwd_f32(_lib.ptr(Q), qb, _lib.ptr(K), B, nq, nk, dq, dk, D, H, ln, _lib.ptr(blob), _lib.ptr(dout),
                                               _lib.ptr(saved), saved.numel(), _lib.ptr(dparams), _lib.ptr(dQ), _lib.ptr(dK),
                                               _lib.ptr(ws), ws.numel(), rt.stream_ptr(dev)), "mab_train_bwd")
        grads, off = [], 0
        for shp in ctx.shapes:
            n = 1
            for s_ in shp:
                n *= s_
            grads.append(dparams[off:off + n].view(shp))
            off += n
        return (dQ, dK, None, None, *grads)


class DeepSetTrainFunction(torch.autograd.Function):
    """out = DeepSet(X; params): training forward / backward of set_transformer-master/models.py:3-28 (pool mean / max / sum)."""

    @staticmethod
    def forward(ctx, X, blob, cfg, *params):
        d_in, dh, out_dim, pool = cfg
        B, N, _ = X.shape
        L = _lib.lib()
        dev = X.device
        out = torch.empty((B, out_dim), dtype=torch.float32, device=dev)
        saved = torch.empty(max(1, L.pca_deepset_train_saved_bytes(B, N, dh)), dtype=torch.uint8, device=dev)
        ws = rt.workspace(dev, L.pca_deepset_train_workspace_bytes(B, N, dh))
        with torch.cuda.device(dev):
            _lib.check(L.pca_deepset_train_fwd_f32(_lib.ptr(X), B, N, d_in, dh, out_dim, pool, _lib.ptr(blob), _lib.ptr(out),
                                                   _lib.ptr(saved), saved.numel(), _lib.ptr(ws), ws.numel(), rt.stream_ptr(dev)),
                       "deepset_train_fwd")
        ctx.save_for_backward(X, blob, saved)
        ctx.cfg = cfg
        ctx.shapes = [tuple(p.shape) for p in params]
        return out

    @staticmethod
    def backward(ctx, dout):
        X, blob, saved = ctx.saved_tensors
        d_in, dh, out_dim, pool = ctx.cfg
        B, N, _ = X.shape
        L = _lib.lib()
        dev = X.device
        need_dx = bool(ctx.needs_input_grad[0])
        dparams = torch.empty_like(blob)
        dX = torch.empty_like(X) if need_dx else None
        ws = rt.workspace(dev, L.pca_deepset_train_workspace_bytes(B, N, dh))
        dout = rt.f32c(dout)
        with torch.cuda.device(dev):
            _lib.check(L.pca_deepset_train_bwd_f32(_lib.ptr(X), B, N, d_in, dh, out_dim, pool, _lib.ptr(blob), _lib.ptr(dout),
                                                   _lib.ptr(saved), saved.numel(), _lib.ptr(dparams), _lib.ptr(dX), _lib.ptr(ws),
                                                   ws.numel(), rt.stream_ptr(dev)), "deepset_train_bwd")
        grads, off = [], 0
        for shp in ctx.shapes:
            n = 1
            for s_ in shp:
                n *= s_
            grads.append(dparams[off:off + n].view(shp))
            off += n
        return (dX, None, None, *grads)


class SetTrainer:
    """Fused data-parallel training step for ``ST`` / ``SetTransformer`` (one process per GPU).

    ``step(X, labels)`` enqueues forward + cross-entropy + backward + (world > 1) ONE NCCL all-reduce of the flat fp32
    gradient buffer + fused Adam on the caller's current stream and returns ``(loss, correct)`` as device tensors.
    ``overlap_allreduce=True`` runs the backward in two phases and reduces the tail of the blob (ISAB 1, PMA, Linear) on a
    side stream while ISAB 0 is still differentiated; measured on 8 B200 at the config-5 shape it does not pay (3.35 vs
    3.32 ms per step: two latency-bound 2 MB reductions cost more than the one they replace), so it is opt-in.  Returns
    ``(loss, correct)`` as device tensors
    (mean loss of the local batch, number of correct arg-max predictions) without synchronising.
    Hyper-parameters follow torch.optim.Adam as the reference uses it (lr 1e-3; weight_decay 1e-3 for the audio models,
    Code/settransformer.py:90-91)."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None, seed=0,
                 overlap_allreduce=None):
        self.model = model
        if overlap_allreduce is None:                         # PCA_TRAIN_OVERLAP=1 switches the two-phase overlap on
            import os
            overlap_allreduce = os.environ.get("PCA_TRAIN_OVERLAP", "0") == "1"
        self.overlap_allreduce = bool(overlap_allreduce)      # world > 1: all-reduce the blob's tail under ISAB 0's backward
        self._comm_stream = None
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.flat = model.flatten_parameters()
        self.grads = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.t = 0
        self.seed = int(seed)
        self._saved = None
        if self.world > 1:                      # replicas start from rank 0's weights, as DataParallel's broadcast does
            dist.broadcast(self.flat, src=0, group=process_group)

    def _buffers(self, dims, B, N, p, dev):
        L = _lib.lib()
        need = max(1, L.pca_st_train_saved_bytes(C.byref(dims), B, N, p))
        if self._saved is None or self._saved.numel() < need:
            self._saved = torch.empty(need, dtype=torch.uint8, device=dev)
        ws = rt.workspace(dev, L.pca_st_train_workspace_bytes(C.byref(dims), B, N))
        return self._saved, ws

    def step(self, X, labels, counts=None):
        """``counts`` (B,) int32 CUDA, optional: variable-size sets (cloud b = its first counts[b] rows)."""
        m = self.model
        rt.require_cuda(X, "SetTrainer.step")
        X = rt.f32c(X)
        labels = labels.to(device=X.device, dtype=torch.int64).contiguous()
        if counts is not None:
            counts = counts.to(device=X.device, dtype=torch.int32).contiguous()
        B, N, _ = X.shape
        dev = X.device
        dims = m._dims()
        if dims.S != 1:
            raise ValueError("SetTrainer: classification needs num_outputs == 1")
        p = m._dropout_p()
        self.t += 1
        seed = (self.seed * 0x9E3779B1 + self.t) & 0xFFFFFFFFFFFFFFFF
        saved, ws = self._buffers(dims, B, N, p, dev)
        L = _lib.lib()
        st = rt.stream_ptr(dev)
        logits = torch.empty((B, dims.C), dtype=torch.float32, device=dev)
        dlogits = torch.empty_like(logits)
        stats = torch.zeros(2, dtype=torch.float32, device=dev)      # loss | correct (int32 bits)
        with torch.cuda.device(dev):
            _lib.check(L.pca_st_train_fwd_f32(_lib.ptr(X), _lib.ptr(counts), B, N, C.byref(dims), _lib.ptr(self.flat), p, seed, _lib.ptr(logits),
                                              _lib.ptr(saved), saved.numel(), _lib.ptr(ws), ws.numel(), st), "st_train_fwd")
            _lib.check(L.pca_cross_entropy_f32(_lib.ptr(logits), _lib.ptr(labels), B, dims.C, C.c_void_p(stats.data_ptr()),
                                               C.c_void_p(stats.data_ptr() + 4), _lib.ptr(dlogits), st), "cross_entropy")
            world = dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1
            if world > 1 and self.overlap_allreduce:
                # backward in two phases: the all-reduce of the blob's tail (ISAB 1, PMA, Linear gradients: final after phase
                # 1) runs on a side stream while ISAB 0 is differentiated; the head follows on the main stream
                off = C.c_longlong(0)
                bwd = lambda phase: _lib.check(L.pca_st_train_bwd_phase_f32(
                    _lib.ptr(X), _lib.ptr(counts), B, N, C.byref(dims), _lib.ptr(self.flat), p, seed, _lib.ptr(dlogits), _lib.ptr(saved),
                    saved.numel(), _lib.ptr(self.grads), None, _lib.ptr(ws), ws.numel(), phase, C.byref(off), st), "st_train_bwd")
                bwd(1)
                main = torch.cuda.current_stream(dev)
                if self._comm_stream is None:
                    self._comm_stream = torch.cuda.Stream(device=dev)
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(self._comm_stream):
                    self._comm_stream.wait_event(ev)
                    dist.all_reduce(self.grads[off.value:], op=dist.ReduceOp.SUM, group=self.group)
                bwd(2)
                dist.all_reduce(self.grads[:off.value], op=dist.ReduceOp.SUM, group=self.group)
                main.wait_stream(self._comm_stream)
                grad_scale = 1.0 / world
            else:
                _lib.check(L.pca_st_train_bwd_f32(_lib.ptr(X), _lib.ptr(counts), B, N, C.byref(dims), _lib.ptr(self.flat), p, seed,
                                                  _lib.ptr(dlogits), _lib.ptr(saved), saved.numel(), _lib.ptr(self.grads), None, _lib.ptr(ws),
                                                  ws.numel(), st), "st_train_bwd")
                grad_scale = reduce_flat_gradient_(self.grads, self.group)
            _lib.check(L.pca_adam_step_f32(_lib.ptr(self.flat), _lib.ptr(self.grads), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                           self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                           self.t, grad_scale, st), "adam_step")
        self.logits = logits
        return stats[0], stats[1:2].view(torch.int32)[0]
