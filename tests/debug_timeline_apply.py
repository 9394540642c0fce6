"""Per-phase clock stamps of one softmax warp of the apply kernel (debug aid; builds an instrumented copy of the library
with -DPCA_TIMELINE next to the product one).  python tests/debug_timeline_apply.py [--build-only]"""
import collections
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

TL_LIB = os.path.join(g.CSRC, "libpcaudio_b200_tl.so")


def build_tl():
    srcs = [os.path.join(g.CSRC, s) for s in g.SOURCES]
    if os.path.exists(TL_LIB) and all(os.path.getmtime(TL_LIB) > os.path.getmtime(s) for s in srcs):
        return
    subprocess.run(["/usr/local/cuda/bin/nvcc"] + g.NVCC_FLAGS + ["-DPCA_TIMELINE", "-o", TL_LIB] + srcs, check=True, cwd=g.CSRC)


build_tl()
if "--build-only" in sys.argv:
    sys.exit(0)

import numpy as np
import torch

import pcaudio_b200 as pca
from pcaudio_b200 import _lib

_lib.LIB_PATH = TL_LIB
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
names = {0: "item: before ld wait", 1: "ld wait done", 2: "max + exp a + st", 3: "s_full(next) wait + ld issue", 4: "exp b + st",
         5: "st wait", 6: "fence + arrive p_ready", 7: "o_full[0] wait", 8: "O1 pair 0", 9: "o_full[1] wait", 10: "O1 pair 1",
         11: "arrive o1_ready", 12: "f_epilogue start", 13: "f_full wait", 14: "f_epilogue done", 15: "tile start",
         16: "issue_loads(0) done", 20: "item0 start", 21: "item1 start", 22: "item2 start", 23: "item3 start", 24: "s_full wait done",
         25: "ld wait done", 26: "max/exp/scale/pack/st issued", 27: "st wait done"}
d = np.diff(clk)
agg = collections.defaultdict(list)
for i in range(1, len(tags)):
    agg[(int(tags[i - 1]), int(tags[i]))].append(int(d[i - 1]))
ntile = int((tags == 15).sum()) or int((tags == 20).sum())
print("stamps", len(tags), "span cycles", int(clk[-1] - clk[0]), "tiles", ntile)
tot = 0
for k, v in sorted(agg.items()):
    v = np.array(v[4:]) if len(v) > 16 else np.array(v)
    print(f"{names[k[0]]:30s} -> {names[k[1]]:30s} n={len(v):4d} median={int(np.median(v)):6d} mean={int(v.mean()):6d} "
          f"p90={int(np.percentile(v, 90)):6d}  per-tile {v.sum() / max(ntile, 1):8.0f}")
    tot += v.sum()
per_tile = np.diff(clk[tags == (15 if (tags == 15).any() else 20)])
print("cycles per tile: median", int(np.median(per_tile)), "mean", int(per_tile.mean()))
