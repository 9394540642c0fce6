"""Host mirror of Code/utils.py for the hot path: ``pc_maxK`` with the selection done on the GPU."""
from __future__ import annotations

import numpy as np
import torch

from .frontend import random_points, topk_points


def pc_maxK(x, farr, Kmax, device="cuda"):
    """Per-frame top-K of the spectrum (Code/utils.py:25-52).

    x: array (N, T) spectral frames; farr: (N,) frequency coordinates; returns
    (subsampled_x (K, T), subsampled_x_fs (K, T)) with each column in descending-magnitude order,
    same dtypes as the inputs.  The K largest bins per frame and their order come from the CUDA
    radix-select/compaction kernel; values are gathered from the caller's arrays by index, so they
    are bit-identical to the reference's fancy indexing."""
    x = np.asarray(x)
    farr = np.asarray(farr)
    n, t = x.shape
    k = min(int(Kmax), n)
    keys = torch.from_numpy(np.ascontiguousarray(x.T, dtype=np.float32)).to(device)     # (T, N)
    _, idx = topk_points(keys, None, None, k, sorted_desc=True, want_points=False)
    idx = idx.cpu().numpy().astype(np.int64)                                              # (T, K)
    if x.dtype == np.float64:
        # float32 device keys: re-rank the float32 candidates by the original float64 values (see refine_topk_float64)
        for c in range(t):
            idx[c] = refine_topk_float64(x[:, c], idx[c], k)
    cols = np.arange(t)[:, None]
    return x[idx, cols].T.copy(), farr[idx].T.copy()


def refine_topk_float64(vals64, idx_f32, k):
    """Exact ``(-vals64).argsort(kind='stable')[:k]`` from the device's float32 selection ``idx_f32`` (descending float32 keys,
    lowest index first among equal keys): every element whose float32 key is not below the k-th selected key is a candidate
    (float32 rounding is monotone, so the exact top-k is among them); the candidates are ordered by the float64 values."""
    vals64 = np.asarray(vals64, dtype=np.float64)
    k32 = vals64.astype(np.float32)
    cand = np.nonzero(k32 >= k32[int(idx_f32[k - 1])])[0]
    order = np.lexsort((cand, -vals64[cand]))          # primary key: descending value; secondary: ascending index
    return cand[order][:k].astype(np.int64)


def pc_randK(x, farr, Kmax, device="cuda", seed=0):
    """Per-frame random-K subsampling of the spectrum (Code/utils.py:55-82): (subsampled_x (K, T), subsampled_x_fs (K, T)),
    each column a uniformly random K-subset of the bins in random order.  The draw comes from the CUDA path
    (``frontend.random_points``); numpy's generator is not involved (distributional parity)."""
    x = np.asarray(x)
    farr = np.asarray(farr)
    n, t = x.shape
    k = min(int(Kmax), n)
    keys = torch.from_numpy(np.ascontiguousarray(x.T, dtype=np.float32)).to(device)     # (T, N)
    _, idx = random_points(keys, np.zeros(n), None, k, seed=seed)
    idx = idx.cpu().numpy().astype(np.int64)
    cols = np.arange(t)[:, None]
    return x[idx, cols].T.copy(), farr[idx].T.copy()
