"""One stand-alone launch of the selection kernel on config-4 shaped clouds (for ncu captures): 2048 clouds x 16 384 points."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
import pcaudio_b200 as pca
dev = torch.device("cuda:0")
K = int(os.environ.get("TOPK_K", "256"))
audio = (0.1 * torch.randn(2048, 16000, device=dev)).clamp_(-1, 1)
lm = pca.stft_logmag(audio, 1024, drop_nyquist=True, n_frames=32)
import numpy as np
farr, tarr = np.arange(512, dtype=np.float64), np.arange(32, dtype=np.float64)
for _ in range(3):
    pts, idx = pca.topk_points(lm.reshape(2048, 32, 512), farr, tarr, K, sorted_desc=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    pca.topk_points(lm.reshape(2048, 32, 512), farr, tarr, K, sorted_desc=True)
e1.record(); torch.cuda.synchronize()
from pcaudio_b200 import _lib
_lib.profile_enable(True)
pca.topk_points(lm.reshape(2048, 32, 512), farr, tarr, K, sorted_desc=True)
torch.cuda.synchronize()
rep = _lib.profile_report()
_lib.profile_enable(False)
print("topk K=%d: %.4f ms per call (incl. host wrapper); kernel %.4f ms" % (K, e0.elapsed_time(e1) / 10, rep["topk_kernel"]["ms"]))
