"""Per-role clock stamps of the streaming transposed pooled-attention kernel (debug aid; builds an instrumented copy of the
library with -DPCA_TIMELINE).  python tests/debug_timeline_pool.py"""
import collections
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

TL_LIB = os.path.join(g.CSRC, "libpcaudio_b200_tl.so")
srcs = [os.path.join(g.CSRC, s) for s in g.SOURCES]
if not (os.path.exists(TL_LIB) and all(os.path.getmtime(TL_LIB) > os.path.getmtime(s) for s in srcs)):
    subprocess.run(["/usr/local/cuda/bin/nvcc"] + g.NVCC_FLAGS + ["-DPCA_TIMELINE"] + os.environ.get("PCA_TL_DEFS", "").split() + ["--shared", "-o", TL_LIB] + srcs, check=True, cwd=g.CSRC)
if "--build-only" in sys.argv:
    sys.exit(0)
os.environ["PCA_TL_POOL"] = "1"
os.environ["PCA_TC_POOL"] = "2"
import numpy as np
import torch
import pcaudio_b200 as pca
from pcaudio_b200 import _lib
_lib.LIB_PATH = TL_LIB
dev = torch.device("cuda:0")
st = pca.ST(dim_input=2, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev).set_precision("bf16")
X = torch.rand(4096, 1025, 2, device=dev)
buf = torch.zeros(16000, dtype=torch.int64, device=dev)
with torch.no_grad():
    st(X)
    torch.cuda.synchronize()
    _lib.lib().pca_debug_set_timeline(_lib.ptr(buf))
    st(X)
    torch.cuda.synchronize()
    _lib.lib().pca_debug_set_timeline(None)
raw = buf.cpu().numpy()
for role, name in ((0, "S issuer"), (3, "P V issuer"), (1, "softmax wg0 warp0"), (2, "loader")):
    t = raw[4000 * role:4000 * (role + 1)].reshape(-1, 2)
    t = t[t[:, 1] > 0]
    if len(t) < 10:
        print(name, "no stamps"); continue
    span = t[-1, 1] - t[0, 1]
    d = collections.defaultdict(list)
    for (a, ta), (b, tb) in zip(t[:-1], t[1:]):
        d[(int(a), int(b))].append(int(tb - ta))
    print(f"{name}: {len(t)} stamps over {span} cycles")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        v = np.array(v[len(v) // 5:]) if len(v) > 20 else np.array(v)
        print(f"   {k[0]:3d} -> {k[1]:3d}: n={len(v):4d} mean {v.mean():8.0f} median {np.median(v):8.0f} max {v.max():8d}  share {100 * sum(d[k]) / span:5.1f}%")
