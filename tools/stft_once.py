"""One stand-alone launch series of the STFT kernel on config-4 shaped input (for ncu captures): 2048 clips x 1 s, n_fft 1024."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
import pcaudio_b200 as pca
from pcaudio_b200 import _lib
dev = torch.device("cuda:0")
n_fft = int(os.environ.get("STFT_NFFT", "1024"))
clips = int(os.environ.get("STFT_CLIPS", "2048"))
audio = (0.1 * torch.randn(clips, 16000, device=dev)).clamp_(-1, 1)
for _ in range(3):
    lm = pca.stft_logmag(audio, n_fft, drop_nyquist=True)
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(5):
    pca.stft_logmag(audio, n_fft, drop_nyquist=True)
torch.cuda.synchronize()
rep = _lib.profile_report()
_lib.profile_enable(False)
v = rep["stft_logmag_kernel"]
print("stft n_fft=%d clips=%d: kernel %.4f ms, %.0f GB/s algorithmic" % (n_fft, clips, v["ms"], v["bytes"] / v["ms"] / 1e6))
