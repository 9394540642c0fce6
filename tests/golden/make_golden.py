"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference modules in place (Code/dataset.py, Code/utils.py, Code/models.py,
set_transformer-master/modules.py, set_transformer-master/models.py), feeds them seeded
inputs and writes small .npz fixtures next to this script.  Nothing at test/bench run time
reads /root/reference; the GPU box only sees the fixtures.

The shipped FST / 3ST checkpoints are converted to .npz weight fixtures (data, not source)
so that the integration parity test can run the exact published weights.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.modules.setdefault("prettytable", types.SimpleNamespace(PrettyTable=object))
    sys.path.insert(0, os.path.join(REF, "set_transformer-master"))
    sys.path.insert(0, os.path.join(REF, "Code"))
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "Code"))
    import dataset as ref_dataset      # noqa
    import utils as ref_utils          # noqa
    import models as ref_models        # Code/models.py (ST)
    import modules as ref_modules      # set_transformer-master/modules.py
    os.chdir(cwd)
    # set_transformer-master/models.py collides with Code/models.py by name: load by path
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "st_models", os.path.join(REF, "set_transformer-master", "models.py"))
    st_models = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(st_models)
    return ref_dataset, ref_utils, ref_models, ref_modules, st_models


def sd_to_np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def main():
    ref_dataset, ref_utils, ref_models, ref_modules, st_models = import_reference()
    torch.manual_seed(0)
    rs = np.random.RandomState(1234)

    # ---------------------------------------------------------------- datasets / selection
    nf, nt, T = 48, 6, 5
    x3 = rs.randn(nf, nt, T).astype(np.float32)
    farr = np.linspace(0, 16000 / 2, nf) / 16000
    tarr = np.linspace(0, ((0.5 * 96) / 16000) * nt, nt)
    y = np.arange(T)
    out = {"x3": x3, "farr": farr, "tarr": tarr}
    ds = ref_dataset.ESC_pc_temp(x3, y, farr, tarr)
    out["pc_temp"] = np.stack([ds[i][0].numpy() for i in range(T)])
    for K in (1, 17, 64, nf * nt):
        dsk = ref_dataset.ESC_pc_temp_maxKSS(x3, y, farr, tarr, K)
        out[f"pc_temp_maxk_{K}"] = np.stack([dsk[i][0].numpy() for i in range(T)])  # float64
    x2 = rs.randn(nf, 9).astype(np.float32)
    ds2 = ref_dataset.ESC_pc(x2, np.arange(9), farr)
    out["x2"] = x2
    out["pc_2d"] = np.stack([ds2[i][0].numpy() for i in range(9)])
    for K in (1, 10, nf):
        xs, fs_ = ref_utils.pc_maxK(x2, farr, K)
        out[f"pc_maxK_x_{K}"] = xs
        out[f"pc_maxK_f_{K}"] = fs_
        dss = ref_dataset.ESC_pc_ss(xs, np.arange(9), fs_)
        out[f"pc_ss_{K}"] = np.stack([dss[i][0].numpy() for i in range(9)])
    np.savez_compressed(os.path.join(HERE, "pointcloud_golden.npz"), **out)

    # ---------------------------------------------------------------- encoder blocks, random init
    enc = {}
    with torch.no_grad():
        mab = ref_modules.MAB(5, 7, 16, 4)
        Q = torch.randn(3, 6, 5)
        K_ = torch.randn(3, 11, 7)
        enc.update({"mab." + k: v for k, v in sd_to_np(mab.state_dict()).items()})
        enc["mab_Q"], enc["mab_K"], enc["mab_out"] = Q.numpy(), K_.numpy(), mab(Q, K_).numpy()

        mabln = ref_modules.MAB(5, 7, 16, 4, ln=True)
        enc.update({"mabln." + k: v for k, v in sd_to_np(mabln.state_dict()).items()})
        enc["mabln_out"] = mabln(Q, K_).numpy()

        isab = ref_modules.ISAB(3, 16, 4, 8)
        X = torch.randn(2, 37, 3)
        enc.update({"isab." + k: v for k, v in sd_to_np(isab.state_dict()).items()})
        enc["isab_X"], enc["isab_out"] = X.numpy(), isab(X).numpy()

        pma = ref_modules.PMA(16, 4, 2)
        Xp = torch.randn(2, 19, 16)
        enc.update({"pma." + k: v for k, v in sd_to_np(pma.state_dict()).items()})
        enc["pma_X"], enc["pma_out"] = Xp.numpy(), pma(Xp).numpy()

        sab = ref_modules.SAB(6, 16, 2)
        Xs = torch.randn(2, 13, 6)
        enc.update({"sab." + k: v for k, v in sd_to_np(sab.state_dict()).items()})
        enc["sab_X"], enc["sab_out"] = Xs.numpy(), sab(Xs).numpy()

        # ST with the audio hyper-parameters, random init, both input widths
        for d_in in (2, 3):
            st = ref_models.ST(dim_input=d_in, num_outputs=1, dim_output=10, num_inds=64,
                               dim_hidden=64, num_heads=8)
            Xst = torch.randn(3, 200, d_in)
            enc.update({f"st{d_in}." + k: v for k, v in sd_to_np(st.state_dict()).items()})
            enc[f"st{d_in}_X"], enc[f"st{d_in}_out"] = Xst.numpy(), st(Xst).numpy()
            enc[f"st{d_in}_out_b1"] = st(Xst[:1]).numpy()          # squeeze quirk: (C,)

        # ModelNet SetTransformer (main_pointcloud.py:13-37), eval mode; the file cannot be
        # imported (argparse + h5py at import), so build the identical module tree.
        class SetTransformerMN(torch.nn.Module):
            def __init__(self, dim_input=3, num_outputs=1, dim_output=40, num_inds=32,
                         dim_hidden=128, num_heads=4, ln=False):
                super().__init__()
                self.enc = torch.nn.Sequential(
                    ref_modules.ISAB(dim_input, dim_hidden, num_heads, num_inds, ln=ln),
                    ref_modules.ISAB(dim_hidden, dim_hidden, num_heads, num_inds, ln=ln))
                self.dec = torch.nn.Sequential(
                    torch.nn.Dropout(), ref_modules.PMA(dim_hidden, num_heads, num_outputs, ln=ln),
                    torch.nn.Dropout(), torch.nn.Linear(dim_hidden, dim_output))

            def forward(self, X):
                return self.dec(self.enc(X)).squeeze()

        # small hidden size keeps the fixture small; the config-5 dims (256/4/16) are
        # checked GPU-vs-oracle, with the oracle pinned structurally by this vector
        mn = SetTransformerMN(dim_hidden=64, num_heads=4, num_inds=16).eval()
        Xmn = torch.randn(2, 100, 3)
        enc.update({"mn." + k: v for k, v in sd_to_np(mn.state_dict()).items()})
        enc["mn_X"], enc["mn_out"] = Xmn.numpy(), mn(Xmn).numpy()

        dsn = st_models.DeepSet(3, 2, 5, dim_hidden=32)
        Xd = torch.randn(3, 41, 3)
        enc.update({"ds." + k: v for k, v in sd_to_np(dsn.state_dict()).items()})
        enc["ds_X"], enc["ds_out"] = Xd.numpy(), dsn(Xd).numpy()
    np.savez_compressed(os.path.join(HERE, "encoder_golden.npz"), **enc)

    # ---------------------------------------------------------------- shipped checkpoints
    saves = os.path.join(REF, "Code", "model_saves")
    ck = {}
    for tag, fname, d_in, n_pts in (
            ("fst", "FST(2021-04-26 21_49_40.977943)_net.pth", 2, 1025),
            ("3st", "3ST(2021-04-27 05_14_06.922134)_net.pth", 3, 5120)):
        sd = torch.load(os.path.join(saves, fname), map_location="cpu")
        model = torch.nn.DataParallel(ref_models.ST(dim_input=d_in, num_outputs=1, dim_output=10,
                                                    num_inds=64, dim_hidden=64, num_heads=8))
        model.load_state_dict(sd)
        model = model.module.eval()
        np.savez_compressed(os.path.join(HERE, f"{tag}_weights.npz"), **sd_to_np(sd))
        # realistic cloud: coordinates in [0, .5] / [0, .12], log-magnitudes in [-18, 0]
        g = np.random.RandomState(77 + d_in)
        B = 4
        Xc = np.empty((B, n_pts, d_in), dtype=np.float32)
        Xc[:, :, 0] = g.uniform(0, 0.5, (B, n_pts))
        if d_in == 3:
            Xc[:, :, 1] = g.uniform(0, 0.12, (B, n_pts))
        Xc[:, :, -1] = g.uniform(-18.0, -1.0, (B, n_pts))
        with torch.no_grad():
            ck[f"{tag}_X"] = Xc
            ck[f"{tag}_out"] = model(torch.from_numpy(Xc)).numpy()
    np.savez_compressed(os.path.join(HERE, "checkpoint_golden.npz"), **ck)
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        print(f"  {f}: {os.path.getsize(os.path.join(HERE, f))} B")


if __name__ == "__main__":
    main()
