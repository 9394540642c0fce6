#!/usr/bin/env python
"""Stand-alone timing of the split-bf16 tcgen05 GEMMs (csrc/gemm_tc.cu) at the ModelNet / audio training shapes.
    python tools/gemm_tc_bench.py [rows] [K] [N] [mode]     mode: 0 plain, 3 relu + residual + R"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g

g.build()
from pcaudio_b200 import _lib as L


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 256000
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    dev = torch.device("cuda:0")
    X = torch.randn(rows, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    Y = torch.empty(rows, N, device=dev)
    R = torch.empty(rows, N, device=dev) if mode == 3 else None
    resid = torch.randn(rows, N, device=dev) if mode == 3 else None
    img = torch.empty(4 * N * K, dtype=torch.uint8, device=dev)
    dW = torch.zeros(N, K, device=dev)
    lib = L.lib()

    def lin():
        L.check(lib.pca_debug_linear_tc(L.ptr(X), L.ptr(W), 0, L.ptr(b), L.ptr(resid), L.ptr(Y), L.ptr(R), rows, K, N, int(mode == 3),
                                        L.ptr(img), img.numel(), None), "linear_tc")

    def gw():
        L.check(lib.pca_debug_grad_weight_tc(L.ptr(Y), L.ptr(X), L.ptr(dW), rows, N, K, None), "grad_weight_tc")

    for name, fn, bytes_ in (("linear_tc", lin, 4.0 * rows * (K + N * (3 if mode == 3 else 1))), ("grad_weight_tc", gw, 4.0 * rows * (K + N))):
        if name == "grad_weight_tc" and K > 256:
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name}: rows={rows} K={K} N={N} mode={mode}: {ms:.3f} ms  {2.0 * rows * K * N / ms / 1e9:.1f} TFLOP/s  "
              f"{bytes_ / ms / 1e6:.0f} GB/s algorithmic", flush=True)


if __name__ == "__main__":
    main()
