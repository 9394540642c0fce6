"""Per-kernel time of the bf16 encoder as a function of N (full tiles vs ragged tail)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
import pcaudio_b200 as pca
from pcaudio_b200 import _lib
dev = torch.device("cuda:0")
st = pca.ST(dim_input=2, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev).set_precision("bf16")
for B, N in ((4096, 1024), (4096, 1025), (4096, 1056), (4096, 1152), (4144, 1024), (4096, 2048)):
    X = torch.rand(B, N, 2, device=dev)
    with torch.no_grad():
        for _ in range(3): st(X)
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        for _ in range(5): st(X)
        torch.cuda.synchronize()
        rep = _lib.profile_report()
        _lib.profile_enable(False)
    print(B, N, {k: round(v["ms"] / 5, 3) for k, v in rep.items()})
