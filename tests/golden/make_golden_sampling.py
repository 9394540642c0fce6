"""Golden vectors for the random-K / importance subsampling datasets, from the UNMODIFIED reference
(Code/dataset.py:205-290, Code/utils.py:55-82).  Authoring container only:

    python tests/golden/make_golden_sampling.py

Deterministic outputs (importance top-K, choice=1) are stored as is; the random modes are stored together with the
numpy / torch seeds that produced them so that the CPU oracle (same generators) can be pinned exactly."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    ref_dataset, ref_utils, *_ = import_reference()
    rs = np.random.RandomState(4321)
    nf, nt, T = 40, 12, 4
    x3 = rs.randn(nf, nt, T).astype(np.float32)
    farr = np.linspace(0, 16000 / 2, nf) / 16000
    tarr = np.linspace(0, ((0.5 * 80) / 16000) * nt, nt)
    y = np.arange(T)
    out = {"x3": x3, "farr": farr, "tarr": tarr}
    for K, winF in ((16, 3), (100, 4), (nf * nt, 7)):
        ds = ref_dataset.ESC_pc_temp_importancerandKSS(x3, y, farr, tarr, K, 1, winF)
        out[f"imp_top_K{K}_w{winF}"] = np.stack([ds[i][0].numpy() for i in range(T)])
    K, winF = 64, 5
    ds = ref_dataset.ESC_pc_temp_importancerandKSS(x3, y, farr, tarr, K, 0, winF)
    torch.manual_seed(77)
    out["imp_multinomial_K64_w5_seed77"] = np.stack([ds[i][0].numpy() for i in range(T)])
    ds = ref_dataset.ESC_pc_temp_randKSS(x3, y, farr, tarr, 50)
    np.random.seed(99)
    out["randk_K50_seed99"] = np.stack([ds[i][0].numpy() for i in range(T)])
    np.random.seed(98)
    xs, fs_ = ref_utils.pc_randK(x3[:, :, 0], farr, 10)
    out["pc_randK_x_seed98"], out["pc_randK_f_seed98"] = xs, fs_
    np.savez_compressed(os.path.join(HERE, "sampling_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
