"""Stand-alone runner for the tcgen05 probe: one process per operand-mode pair so that a faulting
variant cannot poison the others.  Usage: python tests/probe_tc.py A_MODE B_MODE"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as g

g.build()
from pcaudio_b200 import _lib

a_mode, b_mode = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
for N, K in [(128, 16), (16, 16), (16, 128), (64, 64), (128, 128)]:
    gen = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=gen)
    B = torch.randn(K, N, generator=gen)
    ref = A.bfloat16().float() @ B.bfloat16().float()
    A_in = (A.t().contiguous() if a_mode == 2 else A).to(dev)
    B_in = (B.t().contiguous() if b_mode == 0 else B).to(dev)
    D = torch.full((128, N), float("nan"), device=dev)
    _lib.check(_lib.lib().pca_debug_umma_probe(_lib.ptr(A_in), _lib.ptr(B_in), _lib.ptr(D), N, K, a_mode, b_mode,
                                               torch.cuda.current_stream().cuda_stream), "umma_probe")
    torch.cuda.synchronize()
    err = (D.cpu() - ref).abs().max().item()
    print(f"a_mode={a_mode} b_mode={b_mode} N={N} K={K}: max abs err {err:.3e} {'OK' if err < 1e-3 * K ** 0.5 else 'MISMATCH'}", flush=True)
