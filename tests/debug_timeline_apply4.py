"""Per-role clock stamps of the fourth-generation apply kernel (debug aid; builds an instrumented copy of the library with
-DPCA_TIMELINE next to the product one).  PCA_TL_APPLY=1 records the d_in<=4 launch, PCA_TL_APPLY64=1 the 64-wide one.
python tests/debug_timeline_apply4.py [--build-only]"""
import collections
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

TL_LIB = os.path.join(g.CSRC, os.environ.get("PCA_TL_LIBNAME", "libpcaudio_b200_tl.so"))


def build_tl():
    srcs = [os.path.join(g.CSRC, s) for s in g.SOURCES]
    if os.path.exists(TL_LIB) and all(os.path.getmtime(TL_LIB) > os.path.getmtime(s) for s in srcs):
        return
    subprocess.run(["/usr/local/cuda/bin/nvcc"] + g.NVCC_FLAGS + ["-DPCA_TIMELINE"] + os.environ.get("PCA_TL_DEFS", "").split() + ["--shared", "-o", TL_LIB] + srcs, check=True, cwd=g.CSRC)


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
if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
    np.save(os.path.join(ROOT, "gpurun_out", "tl_raw.npy"), raw)
# overlap of the exponential phases (stamp 25 -> 21) of the two recorded softmax warps (same scheduler)
def _phases(off):
    t = raw[off:off + 2000].reshape(-1, 2)
    t = t[t[:, 1] > 0]
    out, start = [], None
    for tag, clk in t:
        if tag == 25:
            start = clk
        elif tag == 21 and start is not None:
            out.append((start, clk))
            start = None
    return out
pa, pb = _phases(0), _phases(4000)
if pa and pb:
    tot = ov = 0
    for (a0, a1) in pa[5:-5]:
        tot += a1 - a0
        for (b0, b1) in pb:
            ov += max(0, min(a1, b1) - max(a0, b0))
    print(f"exp-phase overlap of softmax warp 0 with warp 8: {ov / max(tot, 1):.2f} of warp 0's exp time; warp 0 exp phases: {len(pa)}, mean {tot / max(len(pa) - 10, 1):.0f} cycles")
    print("first phases warp0:", [(int(a - pa[0][0]), int(b - pa[0][0])) for a, b in pa[10:16]])
    print("first phases warp8:", [(int(a - pa[0][0]), int(b - pa[0][0])) for a, b in pb[10:16]])
names = {20: "sm: item start", 24: "sm: s_full wait done", 25: "sm: ld wait done", 26: "sm: math + st issued", 27: "sm: st wait done", 21: "sm: exps done", 22: "sm: next s_full / p_free seen", 64: "mma: head start", 65: "mma: s_free wait done", 66: "mma: QK issued",
         40: "epi: start", 41: "epi: o_full wait done", 42: "epi: O1 staged", 43: "epi: arrived (+fc_o issue)",
         44: "epi: f_full wait done", 45: "epi: Y stored", 50: "prod: start", 51: "prod0: ya_full wait done",
         52: "prod: qp_done wait done", 53: "prod: aq_empty wait done", 54: "prod: AQ converted", 56: "prod0: oq_free wait done",
         60: "mma: tile start", 61: "mma: aq_full wait done", 62: "mma: pv step start", 63: "mma: p_ready wait done",
         70: "ld: tile start", 71: "ld: ya_empty wait done"}
for role, off in (("softmax warp 0", 0), ("producer/epilogue warp 12", 2000), ("softmax warp 8", 4000), ("MMA warp", 6000)):
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
        print(f"   {names.get(k[0], k[0]):30s} -> {names.get(k[1], k[1]):30s} n={len(v):4d} median={int(np.median(v)):6d} mean={int(v.mean()):6d} p90={int(np.percentile(v, 90)):6d}")
