"""Worker of tests/test_gpu_multi.py (launched with torch.distributed.run, one rank per GPU): data-parallel SetTrainer steps
over NCCL must reproduce the single-process large-batch steps (flat-gradient all-reduce = mean over ranks)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import pcaudio_b200 as pca
from pcaudio_b200 import parallel


def make(dev):
    torch.manual_seed(7)
    return pca.SetTransformer(dim_input=3, num_outputs=1, dim_output=5, num_inds=8, dim_hidden=32, num_heads=4).to(dev).eval()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    g = torch.Generator().manual_seed(3)
    X = torch.randn(8, 200, 3, generator=g).to(dev)
    y = torch.randint(0, 5, (8,), generator=g).to(dev)
    # single-process reference on the full batch (no process group yet).  Gradients are compared after ONE step from identical
    # weights: Adam divides by sqrt(v), so weights after several steps amplify last-bit gradient differences arbitrarily.
    ref = pca.SetTrainer(make(dev), lr=1e-2)
    ref.step(X, y)
    ref_grads = ref.grads.clone()
    parallel.init_distributed("nccl")
    lo, hi = parallel.shard_range(8, rank, world)
    outs = {}
    for overlap in (True, False):
        tr = pca.SetTrainer(make(dev), lr=1e-2, overlap_allreduce=overlap)
        tr.step(X[lo:hi], y[lo:hi])
        torch.cuda.synchronize(dev)
        g_mean = tr.grads / world                       # the buffer holds the SUM over ranks; Adam folds the 1 / world in
        err = (g_mean - ref_grads).abs().max().item() / ref_grads.abs().max().item()
        assert err < 1e-4, f"rank {rank} overlap={overlap}: all-reduced gradient differs from the large-batch gradient by {err:.3e}"
        for _ in range(2):
            tr.step(X[lo:hi], y[lo:hi])
        torch.cuda.synchronize(dev)
        outs[overlap] = tr.flat.clone()
    # (the backward kernels add split partial sums atomically, so two runs agree to rounding, not bitwise)
    for k in (True, False):                             # every rank applied the same reduced gradient: replicas stay identical
        chk = outs[k].clone()
        dist.all_reduce(chk, op=dist.ReduceOp.MAX)
        assert torch.equal(chk, outs[k]), "ranks diverged"
    if rank == 0:
        print("DDP_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
