"""End-to-end (host buffers) step time of the bench workload for different chunk counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import pcaudio_b200 as pca
dev = torch.device("cuda:0")
w = {k: torch.from_numpy(v) for k, v in np.load("tests/golden/fst_weights.npz").items()}
model = pca.ST(dim_input=2, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev)
model.load_state_dict(w)
cfg = pca.AudioConfig(sampling_rate=16000, window_size=2048, n_samples=16000, mode=2, precision="bf16")
pipe = pca.AudioSetPipeline(model, cfg, dev)
host = [torch.randn(256, 16000).clamp_(-1, 1).pin_memory() for _ in range(4)]
out = torch.empty((4096, 1, 10)).pin_memory()
ref = None
for chunks in (1, 2, 3, 4, 6):
    for i in range(3):
        pipe.run_host(host[i % 4], out, chunks=chunks); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        pipe.run_host(host[i % 4], out, chunks=chunks); torch.cuda.current_stream().synchronize()
    e1.record(); torch.cuda.synchronize()
    pipe.run_host(host[0], out, chunks=chunks); torch.cuda.synchronize()
    if ref is None: ref = out.clone()
    print(f"chunks={chunks}: {e0.elapsed_time(e1) / 10:.3f} ms/step, identical to unchunked: {bool((out == ref).all())}")
