"""Stage-by-stage comparison of the tcgen05 (bf16) encoder against the CPU oracle.
Usage: python tests/debug_tc_stages.py [d_in] [B] [N]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as g

g.build()
import pcaudio_b200 as pca
from oracle import pcaudio_oracle as orc
from pcaudio_b200 import _lib


def stage_oracle(p, X, heads=8):
    I0 = p["enc.0.I"].expand(X.shape[0], -1, -1)
    H1 = orc.mab_forward(p, "enc.0.mab0.", I0, X, heads)
    Y1 = orc.mab_forward(p, "enc.0.mab1.", X, H1, heads)
    I1 = p["enc.1.I"].expand(X.shape[0], -1, -1)
    H2 = orc.mab_forward(p, "enc.1.mab0.", I1, Y1, heads)
    Y2 = orc.mab_forward(p, "enc.1.mab1.", Y1, H2, heads)
    pooled = orc.pma_forward(p, "dec.0.", Y2, heads)
    logits = pooled @ p["dec.1.weight"].T + p["dec.1.bias"]
    return H1, Y1, H2, Y2, pooled.squeeze(1), logits.squeeze(1)


def stages(model, Xd):
    """Every stage output of the bf16 path for the device batch Xd (B, N, d_in), as fp32 device tensors."""
    B, N, _ = Xd.shape
    dev = Xd.device
    dims = model._dims()
    blob = model._blob()
    L = _lib.lib()
    need = L.pca_st_workspace_bytes(C.byref(dims), B, N, _lib.PREC_BF16)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    outs = {"logits": torch.zeros(B, 10, device=dev), "H1": torch.zeros(B, 64, 64, device=dev),
            "Y1": torch.zeros(B, N, 64, device=dev), "H2": torch.zeros(B, 64, 64, device=dev),
            "Y2": torch.zeros(B, N, 64, device=dev), "pooled": torch.zeros(B, 64, device=dev)}
    _lib.check(L.pca_debug_st_stages(_lib.ptr(Xd), B, N, C.byref(dims), _lib.ptr(blob), _lib.ptr(outs["logits"]),
                                     _lib.ptr(outs["H1"]), _lib.ptr(outs["Y1"]), _lib.ptr(outs["H2"]), _lib.ptr(outs["Y2"]),
                                     _lib.ptr(outs["pooled"]), _lib.ptr(ws), ws.numel(),
                                     torch.cuda.current_stream().cuda_stream), "debug_st_stages")
    torch.cuda.synchronize()
    return outs


def run(d_in=2, B=3, N=300, seed=0, ckpt=True):
    dev = torch.device("cuda:0")
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    model = pca.ST(dim_input=d_in, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev)
    if ckpt:
        tag = "fst" if d_in == 2 else "3st"
        model.load_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(gdir, f"{tag}_weights.npz")).items()})
    rs = np.random.RandomState(seed)
    X = np.empty((B, N, d_in), dtype=np.float32)
    X[:, :, 0] = rs.uniform(0, 0.5, (B, N))
    if d_in == 3:
        X[:, :, 1] = rs.uniform(0, 0.12, (B, N))
    X[:, :, -1] = rs.uniform(-18, -1, (B, N))
    outs = stages(model, torch.from_numpy(X).to(dev))
    p = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref = dict(zip(["H1", "Y1", "H2", "Y2", "pooled", "logits"], stage_oracle(p, torch.from_numpy(X))))
    errs = {}
    for k in ["H1", "Y1", "H2", "Y2", "pooled", "logits"]:
        a, b = outs[k].cpu().numpy().astype(np.float64), ref[k].numpy().astype(np.float64)
        errs[k] = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
    return errs


if __name__ == "__main__":
    d_in = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    for k, v in run(d_in, B, N).items():
        print(f"d_in={d_in} B={B} N={N} stage {k}: rel err {v:.3e}", flush=True)
