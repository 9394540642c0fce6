#!/usr/bin/env python
"""Training-step throughput of the set encoders (BASELINE.json config 5 and the FST training shape).

    python tools/train_bench.py [--config modelnet|fst] [--steps K] [--warmup W] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py ...

One step = SetTrainer.step: forward + cross-entropy + backward + (N > 1) ONE NCCL allreduce of the flat fp32 gradient
buffer + fused Adam.  --batch is the GLOBAL batch (config 5: 256), sharded evenly over the ranks (strong scaling of one
optimisation step, as nn.DataParallel does in the reference).  Timed with CUDA events, max over ranks; prints one JSON
line (clouds/s, ms/step, per-kernel device time of one profiled step)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="modelnet", choices=["modelnet", "fst"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    import pcaudio_b200 as pca
    from pcaudio_b200 import _lib, parallel

    rank, world, local = parallel.init_distributed()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(505)
    if args.config == "modelnet":      # main_pointcloud.py defaults: --dim 256 --n_heads 4 --n_anc 16, 1000 points, 40 classes
        model = pca.SetTransformer(dim_hidden=256, num_heads=4, num_inds=16).to(dev).train()
        N, d_in, Cc, wd = 1000, 3, 40, 0.0
        desc = "main_pointcloud.SetTransformer(dim_hidden=256,num_heads=4,num_inds=16), 1000-point clouds, 40 classes, Dropout(0.5), Adam lr 1e-3"
        flops_fwd = 2 * (N * (3 * 3 * 256 + 8 * 16 * 256 + 7 * 256 * 256 + 2 * 256) + 8 * 16 * 256 * 256 + 2 * 256 * 256 + 256 * 40)
    else:                              # Code/settransformer.py: FST, batch of frame clouds, Adam lr 1e-3 wd 1e-3
        model = pca.ST(dim_input=2, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev).train()
        N, d_in, Cc, wd = 1025, 2, 10, 1e-3
        desc = "Code/settransformer.py FST ST(D=64,H=8,M=64), 1025-point frame clouds, 10 classes, Adam lr 1e-3 wd 1e-3"
        flops_fwd = 2 * (N * (3 * 2 * 64 + 8 * 64 * 64 + 7 * 64 * 64 + 2 * 64) + 8 * 64 * 64 * 64 + 2 * 64 * 64 + 64 * 10)
    lo, hi = parallel.shard_range(args.batch, rank, world)
    Bl = hi - lo
    gen = torch.Generator().manual_seed(505)
    pool = []
    for i in range(4):                 # distinct synthetic batches, standardised per cloud (data_modelnet40.standardize :29-34)
        X = torch.randn(args.batch, N, d_in, generator=gen)
        X = (X - X.mean(dim=1, keepdim=True)) / X.std(dim=1, keepdim=True).clamp_min(1e-6)
        y = torch.randint(0, Cc, (args.batch,), generator=gen)
        pool.append((X[lo:hi].to(dev), y[lo:hi].to(dev)))
    tr = pca.SetTrainer(model, lr=1e-3, weight_decay=wd)

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    first_loss = None
    for i in range(args.warmup):
        loss, _ = tr.step(*pool[i % len(pool)])
        if first_loss is None:
            first_loss = loss.item()
    sync()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss, correct = tr.step(*pool[i % len(pool)])
    e1.record()
    sync()
    launches = _lib.launch_count() - l0
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    _lib.profile_enable(True)
    tr.step(*pool[0])
    torch.cuda.synchronize(dev)
    rep = _lib.profile_report()
    _lib.profile_enable(False)
    if rank == 0:
        kernels = {k: round(v["ms"], 4) for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])}
        out = {"metric": "clouds/sec training step (fwd+bwd+allreduce+Adam)", "value": args.batch / ms * 1e3, "unit": "clouds/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "scaling": "strong",
               "dtype": "f32", "data": "synthetic", "config": {"workload": desc, "global_batch": args.batch, "points": N,
                                                              "local_batch": Bl, "parallelism": f"dp{world} (flat-gradient allreduce)"},
               "gpu_launches": int(launches), "achieved_tflops": 3 * flops_fwd * args.batch / ms / 1e9,
               "loss_first": first_loss, "loss_last": loss.item(), "kernels_ms_one_step": kernels}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
