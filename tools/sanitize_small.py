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
# ---- inference everything above ran under grad mode for the ST calls? no: keep them inference-only
with torch.no_grad():
    mn = pca.SetTransformer(dim_hidden=256, num_heads=4, num_inds=16).to(dev).eval()
    e = mn(torch.randn(3, 700, 3, device=dev))                       # split-bf16 tensor-core GEMMs, ragged last tile
    lm = pca.stft_logmag(audio, 1024, drop_nyquist=True)[:, :10].contiguous()
    farr, tarr = pca.coord_tables(16000.0, 512, 1024, 0.5, 10)
    r1 = pca.random_points(lm, farr, tarr, 100, seed=3)
    r2 = pca.importance_points(lm, farr, tarr, 100, 5, choice=0, seed=3)
    r3 = pca.importance_points(lm, farr, tarr, 100, 4, choice=1)
    rs = pca.resample(audio, 16000, 11025, scale=True)
    sab = pca.SetTransformerSAB(3, 4, 6, num_inds=8, dim_hidden=32, num_heads=4, ln=True).to(dev)
    f = sab(torch.randn(2, 90, 3, device=dev))
torch.cuda.synchronize()
# ---- training paths (forward that keeps activations + backward), small odd shapes
for model, X, kw in ((pca.ST(dim_input=3, num_outputs=1, dim_output=5, num_inds=8, dim_hidden=16, num_heads=4).to(dev), torch.randn(3, 131, 3, device=dev), {}),
                     (pca.ST(dim_input=2, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev), torch.randn(4, 1025, 2, device=dev),
                      {"counts": torch.tensor([1025, 1, 513, 640], dtype=torch.int32, device=dev)}),
                     (pca.SetTransformer(dim_hidden=256, num_heads=4, num_inds=16).to(dev).train(), torch.randn(5, 333, 3, device=dev), {}),
                     (pca.SetTransformerSAB(3, 3, 4, num_inds=4, dim_hidden=32, num_heads=2, ln=True).to(dev), torch.randn(2, 77, 3, device=dev), {}),
                     (pca.DeepSet(3, 1, 7, dim_hidden=64, pool="max").to(dev), torch.randn(9, 300, 3, device=dev), {})):
    X.requires_grad_(True)
    out = model(X, **kw)
    out.sum().backward()
    torch.cuda.synchronize()
    assert torch.isfinite(out).all() and torch.isfinite(X.grad).all()
tr = pca.SetTrainer(pca.ST(dim_input=2, num_outputs=1, dim_output=10, num_inds=16, dim_hidden=32, num_heads=8).to(dev), lr=1e-3, weight_decay=1e-3)
loss, correct = tr.step(torch.randn(6, 200, 2, device=dev), torch.randint(0, 10, (6,), device=dev))
torch.cuda.synchronize()
assert torch.isfinite(loss)
print("sanitize_small ok")
