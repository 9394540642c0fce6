"""Split-bf16 tcgen05 GEMMs (csrc/gemm_tc.cu) against float64 matmul: fp32-grade results from three bf16 MMAs per product."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu
SPLIT_BF16_REL_TOL = 1e-4      # max |y - y_ref| / max |y_ref|; measured ~3e-6 (the dropped lo*lo term is 2^-18 relative)


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as g
    g.build()
    from pcaudio_b200 import _lib
    return _lib


@pytest.mark.parametrize("rows,K,N,trans_w,relu,bias,resid,want_r", [
    (1000, 64, 64, 0, 0, True, False, False),        # audio dims, ragged last tile
    (4096, 256, 256, 0, 1, True, True, True),        # ModelNet fc_o: residual + ReLU output kept (training forward)
    (3000, 256, 512, 0, 0, True, False, False),      # K|V projection: two column passes of 256
    (2048, 512, 256, 1, 0, False, True, False),      # grad-input: dX = dKV Wkv + accumulate
    (777, 128, 96, 0, 1, True, False, False),        # N not a power of two
    (130000, 32, 32, 0, 0, True, False, False),      # many tiles per CTA (persistent loop, accumulator ping-pong)
])
def test_linear_tc(L, rows, K, N, trans_w, relu, bias, resid, want_r):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(rows + K + N)
    X = torch.randn(rows, K, generator=g).to(dev)
    W = (torch.randn(K, N, generator=g) if trans_w else torch.randn(N, K, generator=g)).to(dev) / K ** 0.5
    b = torch.randn(N, generator=g).to(dev) if bias else None
    Rs = torch.randn(rows, N, generator=g).to(dev) if resid else None
    Y = torch.full((rows, N), float("nan"), device=dev)
    R = torch.full((rows, N), float("nan"), device=dev) if want_r else None
    img = torch.empty(4 * N * K, dtype=torch.uint8, device=dev)
    L.check(L.lib().pca_debug_linear_tc(L.ptr(X), L.ptr(W), trans_w, L.ptr(b), L.ptr(Rs), L.ptr(Y), L.ptr(R), rows, K, N, relu,
                                        L.ptr(img), img.numel(), None), "linear_tc")
    torch.cuda.synchronize()
    ref = X.double() @ (W.double() if trans_w else W.double().T)
    if bias:
        ref = ref + b.double()
    if relu:
        ref = ref.clamp_min(0)
    if want_r:
        assert ((R.double() - ref).abs().max() / ref.abs().max()).item() < SPLIT_BF16_REL_TOL
    if resid:
        ref = ref + Rs.double()
    err = ((Y.double() - ref).abs().max() / ref.abs().max()).item()
    assert err < SPLIT_BF16_REL_TOL, f"rel err {err:.3e}"


@pytest.mark.parametrize("rows,M,N", [(5000, 64, 64), (4096, 256, 256), (10000, 512, 256), (2500, 128, 32), (2049, 64, 128)])
def test_grad_weight_tc(L, rows, M, N):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(rows + M + N)
    dY = torch.randn(rows, M, generator=g).to(dev)
    X = torch.randn(rows, N, generator=g).to(dev)
    dW = torch.ones(M, N, device=dev)                 # accumulates on top of what is there
    L.check(L.lib().pca_debug_grad_weight_tc(L.ptr(dY), L.ptr(X), L.ptr(dW), rows, M, N, None), "grad_weight_tc")
    torch.cuda.synchronize()
    ref = dY.double().T @ X.double() + 1.0
    err = ((dW.double() - ref).abs().max() / ref.abs().max()).item()
    assert err < SPLIT_BF16_REL_TOL, f"rel err {err:.3e}"


def test_tensor_core_and_cuda_core_paths_agree(L):
    """Same fp32 model, GEMMs on the tensor cores vs on the CUDA cores (pca_debug_set_gemm_tc): ModelNet dims."""
    import pcaudio_b200 as pca
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = pca.SetTransformer(dim_hidden=256, num_heads=4, num_inds=16).to(dev).eval()
    X = torch.randn(8, 1000, 3, device=dev)
    try:
        with torch.no_grad():
            a = model(X).clone()
            L.lib().pca_debug_set_gemm_tc(0)
            b = model(X).clone()
    finally:
        L.lib().pca_debug_set_gemm_tc(1)
    assert not torch.equal(a, b)                      # the two paths really are different kernels
    assert ((a - b).abs().max() / b.abs().max()).item() < 1e-4
