"""Small end-to-end run for compute-sanitizer (memcheck): every kernel family once at small shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pcaudio_b200 as pca
dev = torch.device("cuda:0")
torch.manual_seed(0)
audio = (0.1 * torch.randn(3, 16000, device=dev)).clamp_(-1, 1)
for prec in ("bf16", "fp32"):
    st2 = pca.ST(dim_input=2, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev).set_precision(prec)
    st3 = pca.ST(dim_input=3, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev).set_precision(prec)
    p2 = pca.AudioSetPipeline(st2, pca.AudioConfig(window_size=2048, n_samples=16000, mode=2, precision=prec), dev)
    p3 = pca.AudioSetPipeline(st3, pca.AudioConfig(window_size=1024, n_samples=16000, mode=3, Ntemp=10, top_k=300, precision=prec, threshold=-6.0), dev)
    a = p2(audio); b = p3(audio)
    X = torch.randn(5, 700, 3, device=dev)
    c = st3(X, counts=torch.tensor([700, 1, 129, 513, 640], dtype=torch.int32, device=dev))
    torch.cuda.synchronize()
    assert torch.isfinite(a).all() and torch.isfinite(b).all() and torch.isfinite(c).all()
pts, counts, idx = pca.spectral_point_cloud(audio, n_fft=1024, sr=16000.0, top_k=256, fused=True)
ds = pca.DeepSet(3, 1, 10, dim_hidden=64, pool="max").to(dev)
d = ds(torch.randn(4, 100, 3, device=dev), counts=torch.tensor([100, 5, 64, 99], dtype=torch.int32, device=dev))
torch.cuda.synchronize()
print("sanitize_small ok")
