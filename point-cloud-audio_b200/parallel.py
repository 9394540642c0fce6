"""Multi-GPU plumbing: one process per GPU (torchrun), clips sharded contiguously, no data-path
collective for inference.  Replaces the reference's single-process nn.DataParallel
(Code/settransformer.py:94, Code/pc_temp3d_eval.py:53)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_distributed(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, world, local_rank).
    Single-process runs return (0, 1, 0) without creating a process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_items for this rank; sizes differ by at most one and the
    concatenation over ranks is the identity (per-clip outputs are bit-identical to a 1-GPU run)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local: torch.Tensor, n_total: int, rank: int, world: int) -> torch.Tensor:
    """Optional epilogue: assemble the per-rank logits (rows of the shard) into the full (n_total, ...)
    tensor on every rank.  40 B per clip -- not on the timed data path."""
    if world == 1:
        return local
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def allreduce_mean_(flat_grad: torch.Tensor, world: int) -> torch.Tensor:
    """One flat-bucket gradient allreduce (sum, then / world) per training step -- the replacement for
    DataParallel's reduce_add + broadcast."""
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        flat_grad.div_(world)
    return flat_grad


def reduce_flat_gradient_(flat_grad: torch.Tensor, group=None) -> float:
    """The per-step gradient exchange of data-parallel training (SetTrainer): ONE summing allreduce of the flat fp32
    gradient buffer, in place.  Returns the factor the optimizer applies to the summed gradient (1 / world size; the
    fused Adam kernel folds it in, so no separate division pass runs).  No-op (factor 1.0) without a process group."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world
