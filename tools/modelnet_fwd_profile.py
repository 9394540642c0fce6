#!/usr/bin/env python
"""Per-kernel device times of one inference forward of main_pointcloud.SetTransformer(256, 4 heads, 16 inducing points) on
256 x 1000-point clouds (fp32 parity class: TMA-fed split-bf16 tcgen05 GEMMs, attention of csrc/attn_tc.cu)."""
import torch, sys, json
sys.path.insert(0, "/root/repo")
import __graft_entry__ as g
import pcaudio_b200 as pca
from pcaudio_b200 import _lib as L
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = pca.SetTransformer(dim_hidden=256, num_heads=4, num_inds=16).to(dev).eval()
X = [torch.randn(256, 1000, 3, device=dev) for _ in range(4)]
with torch.no_grad():
    for i in range(3): m(X[i % 4])
    torch.cuda.synchronize()
    L.profile_enable(True)
    m(X[0]); torch.cuda.synchronize()
    rep = L.profile_report(); L.profile_enable(False)
tot = 0
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
    print(f'{k:32s} {v["launches"]:3d} {v["ms"]:.4f}'); tot += v["ms"]
print("sum", tot)
