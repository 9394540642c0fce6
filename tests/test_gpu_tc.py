"""GPU tests of the tcgen05 (bf16 tensor-core) path: building-block probe first, then the encoder
kernels against the fp32 path / the oracle at the bf16 tolerance (2e-2 relative, north_star)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BF16_REL_TOL = 2e-2


@pytest.fixture(scope="module")
def pca():
    import __graft_entry__ as g
    g.build()
    import pcaudio_b200
    return pcaudio_b200


@pytest.mark.parametrize("a_mode,b_mode", [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1)])
@pytest.mark.parametrize("N,K", [(128, 16), (16, 128), (64, 64), (128, 128), (16, 16)])
def test_umma_probe(pca, a_mode, b_mode, N, K):
    from pcaudio_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N * 1000 + K + 10 * a_mode + b_mode)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(K, N, generator=g)
    ref = A.bfloat16().float() @ B.bfloat16().float()
    A_in = (A.t().contiguous() if a_mode == 2 else A).to(dev)
    B_in = (B.t().contiguous() if b_mode == 0 else B).to(dev)
    D = torch.full((128, N), float("nan"), device=dev)
    _lib.check(_lib.lib().pca_debug_umma_probe(_lib.ptr(A_in), _lib.ptr(B_in), _lib.ptr(D), N, K, a_mode, b_mode,
                                               torch.cuda.current_stream().cuda_stream), "umma_probe")
    torch.cuda.synchronize()
    err = (D.cpu() - ref).abs().max().item()
    assert err < 1e-3 * K ** 0.5, f"a_mode={a_mode} b_mode={b_mode} N={N} K={K}: max abs err {err}"
