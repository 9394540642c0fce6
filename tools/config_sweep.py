"""Throughput of the other BASELINE.json configs through the public pipeline call (device-resident audio, CUDA events,
one GPU).  Not the bench line -- context numbers for profiles/.   python tools/config_sweep.py [--quick]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import pcaudio_b200 as pca
dev = torch.device("cuda:0")
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

def model(tag, d_in, precision):
    w = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, f"{tag}_weights.npz")).items()}
    st = pca.ST(dim_input=d_in, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev)
    st.load_state_dict(w)
    return st.set_precision(precision)

def timeit(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def flops_per_cloud(n, d_in):
    return 2 * (n * (3 * d_in * 64 + 8 * 64 * 64 + 7 * 64 * 64 + 2 * 64) + 8 * 64 * 64 * 64 + 2 * 64 * 64 + 64 * 10)

rows = []
def run(name, tag, d_in, n_clips, n_samples, n_fft, mode, ntemp, top_k, precision, iters=5):
    st = model(tag, d_in, precision)
    cfg = pca.AudioConfig(sampling_rate=16000, window_size=n_fft, n_samples=n_samples, mode=mode, Ntemp=ntemp, top_k=top_k, precision=precision)
    pipe = pca.AudioSetPipeline(st, cfg, dev)
    audio = (0.1 * torch.randn(n_clips, n_samples, device=dev)).clamp_(-1, 1)
    ms = timeit(lambda: pipe(audio), iters)
    clouds = n_clips * pipe.clouds_per_clip
    r = {"config": name, "precision": precision, "clips": n_clips, "clouds_per_clip": pipe.clouds_per_clip, "points_per_cloud": pipe.points_per_cloud,
         "ms": ms, "clips_per_s": n_clips / ms * 1e3, "clouds_per_s": clouds / ms * 1e3,
         "encoder_tflops": clouds * flops_per_cloud(pipe.points_per_cloud, d_in) / ms / 1e9}
    rows.append(r); print(json.dumps(r), flush=True)

quick = "--quick" in sys.argv
for prec in ("bf16", "fp32"):
    # config 1: 3ST, batch 16 x 1 s, clip-as-cloud (16 384 points) and the reference's 3 chunk clouds of 5120 points
    run("1: 3ST B=16 1 s clip-as-cloud N=16384", "3st", 3, 16, 16000, 1024, 3, 32, 0, prec, 3 if prec == "fp32" else 10)
    run("1: 3ST B=16 1 s chunk clouds 3 x 5120", "3st", 3, 16, 16000, 1024, 3, 10, 0, prec, 3 if prec == "fp32" else 10)
    if prec == "fp32" and quick: continue
    # config 3: 4 s clips: 12 chunk clouds x 5120 and clip-as-cloud N = 64512 (one GPU's shard of the dp8 batch)
    run("3: 3ST 4 s clips, 12 chunk clouds x 5120 (B=64)", "3st", 3, 64, 64000, 1024, 3, 10, 0, prec, 3)
    run("3: 3ST 4 s clips, clip-as-cloud N=64512 (B=32)", "3st", 3, 32, 64000, 1024, 3, 126, 0, prec, 2)
if not quick:
    # config 4: top-K sweep on the 16 384-point clip cloud (a7 semantics), 10k clips
    for K in (256, 512, 1024, 2048, 4096, 8192):
        run(f"4: sweep 10k clips top-K={K}", "3st", 3, 10000, 16000, 1024, 3, 32, K, "bf16", 3)
# config 5 (forward only; training is the next scope row): main_pointcloud.SetTransformer D=256,H=4,M=16,C=40 on
# ModelNet40-shaped clouds, fp32 CUDA-core path (no tcgen05 kernels for these dims yet)
torch.manual_seed(0)
mn = pca.SetTransformer(dim_input=3, num_outputs=1, dim_output=40, num_inds=16, dim_hidden=256, num_heads=4).to(dev).eval()
Xm = torch.randn(256, 1000, 3, device=dev)
Xm = (Xm - Xm.mean(1, keepdim=True)) / Xm.std(1, keepdim=True)
with torch.no_grad():
    ms = timeit(lambda: mn(Xm), 5)
r = {"config": "5: ModelNet SetTransformer forward B=256 N=1000 (fp32 path)", "precision": "fp32", "clouds": 256, "ms": ms,
     "clouds_per_s": 256 / ms * 1e3, "encoder_tflops": 256 * 1.006e9 / ms / 1e9}
rows.append(r); print(json.dumps(r), flush=True)
ds = pca.DeepSet(3, 1, 40, dim_hidden=256, pool="max").to(dev)
with torch.no_grad():
    ms = timeit(lambda: ds(Xm), 5)
r = {"config": "DeepSet (shared MLP + max-pool) B=256 N=1000 dim_hidden=256 (fp32 path)", "precision": "fp32", "clouds": 256, "ms": ms,
     "clouds_per_s": 256 / ms * 1e3}
rows.append(r); print(json.dumps(r), flush=True)
json.dump(rows, open(os.path.join("gpurun_out", "config_sweep.json"), "w"), indent=1)
