import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g; g.build()
import pcaudio_b200 as pca
from oracle import pcaudio_oracle as orc
dev = torch.device("cuda:0")
for pool, dh, N, B in [("mean", 256, 2100, 3), ("mean", 256, 2100, 1), ("mean", 256, 1000, 1), ("mean", 256, 600, 1), ("mean", 64, 600, 1)]:
    torch.manual_seed(N)
    ds = pca.DeepSet(3, 2, 5, dim_hidden=dh, pool=pool).to(dev)
    X = torch.randn(B, N, 3, device=dev, requires_grad=True)
    G = torch.randn(B, 2, 5)
    out = ds(X); (out * G.to(dev)).sum().backward()
    p = {k: v.detach().cpu().double().requires_grad_(True) for k, v in ds.state_dict().items()}
    Xc = X.detach().cpu().double().requires_grad_(True)
    ref = orc.deepset_forward(p, Xc, 2, 5, pool); (ref * G.double()).sum().backward()
    d = (X.grad.cpu().double() - Xc.grad).abs().reshape(B * N, 3)
    bad = (d.max(dim=1).values > 1e-3 * Xc.grad.abs().max()).nonzero().flatten()
    print(pool, dh, N, B, "max err %.3e" % (d.max() / Xc.grad.abs().max()).item(), "bad rows:", bad.numel(), bad[:8].tolist(), bad[-4:].tolist(), flush=True)
