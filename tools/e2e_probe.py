"""Where does the end-to-end (host-buffer) step differ from the device-resident one?  Times AudioSetPipeline.submit_host /
wait_host on the bench workload with and without the host->device copy, and the device-resident call, over many steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import pcaudio_b200 as pca
dev = torch.device("cuda:0")
w = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(g.ROOT, "tests", "golden", "fst_weights.npz")).items()}
st = pca.ST(dim_input=2, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev); st.load_state_dict(w)
pipe = pca.AudioSetPipeline(st, pca.AudioConfig(sampling_rate=16000, window_size=2048, n_samples=16000, mode=2, precision="bf16"), dev)
pool = [(0.1 * torch.randn(256, 16000, device=dev)).clamp_(-1, 1) for _ in range(10)]
host = [p.cpu().pin_memory() for p in pool[:4]]
outs = [torch.empty((4096, 1, 10)).pin_memory() for _ in range(2)]
def timed(fn, n, tail=None):
    for i in range(5): fn(i)
    if tail: tail()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    if tail: tail()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n
N = 100
print("device resident           : %.3f ms/step (events) %.3f (wall)" % timed(lambda i: pipe(pool[i % 10]), N))
pend = []
def step_host(i):
    pend.append(pipe.submit_host(host[i % 4], outs[i & 1]))
    if len(pend) > 1: pipe.wait_host(pend.pop(0))
def drain():
    while pend: pipe.wait_host(pend.pop(0))
print("submit_host / wait_host   : %.3f ms/step (events) %.3f (wall)" % timed(step_host, N, drain))
# enqueue cost alone: submit without waiting
t0 = time.perf_counter()
for i in range(50): pend.append(pipe.submit_host(host[i % 4], outs[i & 1]))
enq = (time.perf_counter() - t0) * 1e3 / 50
drain()
print("CPU enqueue time per submit_host: %.3f ms" % enq)
t0 = time.perf_counter()
for i in range(50): pipe(pool[i % 10])
enq = (time.perf_counter() - t0) * 1e3 / 50
torch.cuda.synchronize()
print("CPU enqueue time per device call: %.3f ms" % enq)
