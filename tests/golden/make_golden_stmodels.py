"""Golden vectors for the generic set_transformer-master/models.py SetTransformer (ISAB, ISAB -> PMA -> SAB, SAB -> Linear;
:30-44), with and without LayerNorm, from the UNMODIFIED reference.  Authoring container only:
    python tests/golden/make_golden_stmodels.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    *_, st_models = import_reference()
    out = {}
    for tag, ln, k in (("noln", False, 4), ("ln", True, 3)):
        torch.manual_seed(11 + int(ln))
        m = st_models.SetTransformer(2, k, 6, num_inds=8, dim_hidden=32, num_heads=4, ln=ln).eval()
        X = torch.randn(3, 70, 2)
        with torch.no_grad():
            Y = m(X)
        out[f"{tag}_X"], out[f"{tag}_Y"] = X.numpy(), Y.numpy()
        for key, v in m.state_dict().items():
            out[f"{tag}_w_{key}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "stmodels_golden.npz"), **out)
    print({k: v.shape for k, v in out.items() if not "_w_" in k})


if __name__ == "__main__":
    main()
