"""Order sensitivity of the bf16 encoder at extreme input gains: per-stage errors against the CPU oracle for the
big-last / big-first orderings, and the fp32 CUDA path for comparison.  Usage: python tools/debug_gain.py [gain] [N]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import debug_tc_stages as dts
import pcaudio_b200 as pca
from pcaudio_b200 import _lib


def stages(model, Xd):
    B, N, _ = Xd.shape
    dev = Xd.device
    dims, blob, L = model._dims(), model._blob(), _lib.lib()
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
    return {k: v.cpu() for k, v in outs.items()}


def main():
    gain = float(sys.argv[1]) if len(sys.argv) > 1 else 1e5
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    st = pca.ST(dim_input=3, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev)
    g = torch.Generator().manual_seed(N)
    X = torch.randn(4, N, 3, generator=g) * 0.05
    X[:, N - max(1, N // 7):] *= gain
    Xf = torch.flip(X, dims=[1]).contiguous()
    p = {k: v.detach().cpu() for k, v in st.state_dict().items()}
    names = ["H1", "Y1", "H2", "Y2", "pooled", "logits"]
    ref_l = dict(zip(names, dts.stage_oracle(p, X)))
    ref_f = dict(zip(names, dts.stage_oracle(p, Xf)))
    tl, tf = stages(st, X.to(dev)), stages(st, Xf.to(dev))

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / max(b.double().abs().max(), 1e-30))
    for k in names:
        fl = torch.flip(tf[k], dims=[1]) if k in ("Y1", "Y2") else tf[k]
        rfl = torch.flip(ref_f[k], dims=[1]) if k in ("Y1", "Y2") else ref_f[k]
        print(f"gain={gain:g} N={N} {k:7s} |ref|max {ref_l[k].abs().max():.3e}  bf16(last) vs oracle {rel(tl[k], ref_l[k]):.3e}  "
              f"bf16(first) vs oracle {rel(fl, ref_l[k]):.3e}  bf16 last-vs-first {rel(tl[k], fl):.3e}  oracle last-vs-first {rel(ref_l[k], rfl):.3e}",
              flush=True)
    with torch.no_grad():
        st.set_precision("fp32")
        a, b = st(X.to(dev)).cpu(), st(Xf.to(dev)).cpu()
    print(f"fp32 CUDA path last-vs-first {rel(a, b):.3e}  vs oracle {rel(a, ref_l['logits']):.3e}")


if __name__ == "__main__":
    main()
