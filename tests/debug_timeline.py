"""Per-phase clock stamps of one softmax warp of the reduce kernel (debug aid).  python tests/debug_timeline.py"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as g

g.build()
import pcaudio_b200 as pca
from pcaudio_b200 import _lib

dev = torch.device("cuda:0")
st = pca.ST(dim_input=2, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev).set_precision("bf16")
X = torch.rand(4096, 1025, 2, device=dev)
buf = torch.zeros(8000, dtype=torch.int64, device=dev)
with torch.no_grad():
    st(X)
    torch.cuda.synchronize()
    _lib.lib().pca_debug_set_timeline(_lib.ptr(buf))
    st(X)
    torch.cuda.synchronize()
    _lib.lib().pca_debug_set_timeline(None)
t = buf.cpu().numpy().reshape(-1, 2)
t = t[t[:, 1] > 0]
tags, clk = t[:, 0], t[:, 1]
names = {0: "item start", 1: "ld wait done", 2: "exp chunk a + st", 3: "mid wait + prefetch issue", 4: "exp chunk b + st",
         5: "st wait", 6: "fence+arrive", 7: "o_full wait", 8: "consume"}
d = np.diff(clk)
agg = collections.defaultdict(list)
for i in range(1, len(tags)):
    agg[(int(tags[i - 1]), int(tags[i]))].append(int(d[i - 1]))
print("stamps", len(tags), "span cycles", int(clk[-1] - clk[0]), "items", int((tags == 0).sum()))
for k, v in sorted(agg.items()):
    v = np.array(v[8:]) if len(v) > 16 else np.array(v)
    print(f"{names[k[0]]:28s} -> {names[k[1]]:28s} n={len(v):4d} median={int(np.median(v)):6d} mean={int(v.mean()):6d} p90={int(np.percentile(v, 90)):6d}")
per_item = np.diff(clk[tags == 0])
print("cycles per item: median", int(np.median(per_item)), "mean", int(per_item.mean()))
