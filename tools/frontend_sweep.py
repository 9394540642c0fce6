"""Front-end timing for the BASELINE config-4 shapes: 1 s clips -> 16 384-point clip clouds -> top-K, fused vs unfused."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
import pcaudio_b200 as pca
dev = torch.device("cuda:0")
B = 2048
audio = (0.1 * torch.randn(B, 16000, device=dev)).clamp_(-1, 1)
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
rows = []
for K in (256, 512, 1024, 2048, 4096, 8192):
    tf = timeit(lambda: pca.spectral_point_cloud(audio, n_fft=1024, sr=16000.0, top_k=K, fused=True))
    tu = timeit(lambda: pca.spectral_point_cloud(audio, n_fft=1024, sr=16000.0, top_k=K, fused=False))
    alg = B * (64000 + 16 * K)          # audio read + 16 B per selected point
    from pcaudio_b200 import _lib
    _lib.profile_enable(True)
    pca.spectral_point_cloud(audio, n_fft=1024, sr=16000.0, top_k=K, fused=False)
    torch.cuda.synchronize()
    rep = _lib.profile_report()
    _lib.profile_enable(False)
    kern = {k: {"ms": v["ms"], "algorithmic_GBps": v["bytes"] / v["ms"] / 1e6, "frac_of_hbm_peak": v["bytes"] / v["ms"] / 1e6 / peaks["hbm_gbs"]} for k, v in rep.items()}
    rows.append({"K": K, "unfused_kernels": kern, "clips": B, "fused_ms": tf, "unfused_ms": tu, "fused_clips_per_s": B / tf * 1e3,
                 "fused_algorithmic_GBps": alg / tf / 1e6, "frac_of_hbm_peak": alg / tf / 1e6 / peaks["hbm_gbs"]})
    print(json.dumps(rows[-1]))
