#!/usr/bin/env python
"""Per-launch device times of the tensor-core attention route (csrc/attn_tc.cu) at the ModelNet training shape
(main_pointcloud.py:62: dim 256, 4 heads, 16 inducing points, 1000 points), forward and backward of the two MAB kinds:

    python tools/attn_tc_bench.py [--batch 256] [--reps 5] [--once]

Prints one JSON line: per (case, kernel) mean milliseconds per launch and the algorithmic TB/s the launcher states for it.
--once runs every case exactly one time without the profile brackets (the form to put under ncu)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--once", action="store_true")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build()
    from pcaudio_b200 import _lib as L
    dev = torch.device("cuda:0")
    B, N, D, H = args.batch, 1000, 256, 4
    cases = [("mab1 (1000 queries x 16 keys)", N, 16, 0), ("mab0 (16 shared queries x 1000 keys)", 16, N, 1),
             ("pma (1 shared seed x 1000 keys)", 1, N, 1)]
    out = {}
    for name, nq, nk, shared in cases:
        gen = torch.Generator().manual_seed(nq + nk)
        Qp = torch.randn(1 if shared else B, nq, D, generator=gen).to(dev)
        KV = torch.randn(B, nk, 2 * D, generator=gen).to(dev)
        dO = torch.randn(B, nq, D, generator=gen).to(dev)
        O = torch.empty(B, nq, D, device=dev)
        lse = torch.empty(B, nq, H, device=dev)
        delta = torch.zeros(B, nq, H, device=dev)
        dQp = torch.empty(B, nq, D, device=dev)
        dKV = torch.empty(B, nk, 2 * D, device=dev)
        ws = torch.empty(L.lib().pca_debug_attn_ws_bytes(B, nq, nk, D, H), dtype=torch.uint8, device=dev)

        def run():
            L.check(L.lib().pca_debug_attn_fwd(L.ptr(Qp), shared, L.ptr(KV), B, nq, nk, D, H, L.ptr(O), L.ptr(lse), L.ptr(ws), ws.numel(),
                                               None), "attn_fwd")
            L.check(L.lib().pca_debug_attn_bwd_tc(L.ptr(Qp), shared, L.ptr(KV), L.ptr(dO), L.ptr(lse), L.ptr(delta), B, nq, nk, D, H,
                                                  L.ptr(dQp), L.ptr(dKV), L.ptr(ws), ws.numel(), None), "attn_bwd")
        if args.once:
            run()
            torch.cuda.synchronize()
            continue
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        L.profile_enable(True)
        for _ in range(args.reps):
            run()
        torch.cuda.synchronize()
        rep = L.profile_report()
        L.profile_enable(False)
        out[name] = {k: {"ms_per_launch": round(v["ms"] / v["launches"], 4), "launches_per_pass": v["launches"] // args.reps,
                         "algorithmic_TBps": round(v["bytes"] / v["launches"] / (v["ms"] / v["launches"] * 1e-3) / 1e12, 2)
                         if v.get("bytes") else None}
                     for k, v in rep.items()}
    if not args.once:
        print(json.dumps({"batch": B, "cases": out}))


if __name__ == "__main__":
    main()
