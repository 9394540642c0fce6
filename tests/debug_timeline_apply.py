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
X = torch.rand(4096, 1024, 2, device=dev)
buf = torch.zeros(8000, dtype=torch.int64, device=dev)
with torch.no_grad():
    st(X)
    torch.cuda.synchronize()
    _lib.lib().pca_debug_set_timeline(_lib.ptr(buf))
    st(X)
    torch.cuda.synchronize()
    _lib.lib().pca_debug_set_timeline(None)
raw = buf.cpu().numpy()
names = {20: "item start", 24: "s_full wait done", 25: "ld wait done", 26: "softmax math + st issued", 27: "st wait done",
         40: "epi: tile start", 41: "epi: o_full wait done", 42: "epi: O1 -> bf16 TMEM done", 43: "epi: arrived o1_ready (+MMA issue)",
         44: "epi: f_full wait done", 45: "epi: Y stored", 50: "prod: tile start", 51: "prod: input staged + arrived (reduce: loaded)",
         52: "prod: qp_done wait done", 53: "prod: aq_empty wait done", 54: "prod: AQ converted", 56: "prod0: oq_free wait done",
         60: "mma: tile start", 61: "mma: aq_full wait done", 62: "mma: S issued", 63: "mma: p_ready wait done"}
for role, off in (("softmax warp 0", 0), ("epilogue warp 16", 2000), ("producer warp 8", 4000), ("MMA chain 0", 6000)):
    t = raw[off:off + 2000].reshape(-1, 2)
    t = t[t[:, 1] > 0]
    if len(t) < 10:
        continue
    tags, clk = t[:, 0], t[:, 1]
    d = np.diff(clk)
    agg = collections.defaultdict(list)
    for i in range(1, len(tags)):
        agg[(int(tags[i - 1]), int(tags[i]))].append(int(d[i - 1]))
    first = int(tags[0])
    per = np.diff(clk[tags == first])
    print(f"== {role}: {len(tags)} stamps; cycles per period (tag {first}): median {int(np.median(per))} mean {int(per.mean())}")
    for k, v in sorted(agg.items()):
        v = np.array(v[2:]) if len(v) > 8 else np.array(v)
        print(f"   {names.get(k[0], k[0]):36s} -> {names.get(k[1], k[1]):36s} n={len(v):4d} median={int(np.median(v)):6d} mean={int(v.mean()):6d} p90={int(np.percentile(v, 90)):6d}")
