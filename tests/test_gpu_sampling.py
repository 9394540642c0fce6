"""Random-K and importance subsampling on the device against the CPU oracle, which is pinned to the
reference's datasets in tests/test_oracle_golden.py.  The deterministic parts (heat map, its top-K, the reference's index
quirk) are compared directly; the random modes cannot share numpy's / torch's generators and are checked by distribution."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pcaudio_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HEAT_REL_TOL = 1e-5


@pytest.fixture(scope="module")
def pca():
    import __graft_entry__ as g
    g.build()
    import pcaudio_b200
    return pcaudio_b200


def _logmag(x3, dev):      # (Nf, Nt, T) -> (T, Nt, Nf)
    return torch.from_numpy(np.ascontiguousarray(x3.transpose(2, 1, 0))).to(dev)


@pytest.mark.parametrize("nf,nt,winF", [(40, 12, 3), (40, 12, 4), (512, 10, 7), (33, 1, 5), (1, 9, 2)])
def test_heat_map_vs_oracle(pca, nf, nt, winF):
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(nf + nt + winF)
    x3 = rs.randn(nf, nt, 3).astype(np.float32)
    heat = pca.importance_heat(_logmag(x3, dev), winF).cpu().numpy()
    for i in range(3):
        if nf == 1 or nt == 1:
            continue        # torch.gradient refuses size-1 dimensions; the kernel defines that gradient as 0 (finite output)
        ref = orc.importance_heat(x3[:, :, i], winF)
        assert np.abs(heat[i] - ref).max() <= HEAT_REL_TOL * np.abs(ref).max()
    assert np.isfinite(heat).all() and (heat > 0).all()


@pytest.mark.parametrize("K,winF", [(16, 3), (100, 4), (480, 7)])
def test_importance_topk_dataset_vs_reference_golden(pca, K, winF):
    dev = torch.device("cuda:0")
    g = dict(np.load(os.path.join(G, "sampling_golden.npz")))
    x3, farr, tarr = g["x3"], g["farr"], g["tarr"]
    ds = pca.ESC_pc_temp_importancerandKSS(x3, np.arange(x3.shape[2]), farr, tarr, K, 1, winF, device=dev)
    heat = pca.importance_heat(_logmag(x3, dev), winF).cpu().numpy()
    same = 0
    for i in range(x3.shape[2]):
        # selection is bit-exact given the keys: the indices equal the stable arg-sort of OUR heat map ...
        assert np.array_equal(ds.indices(i), orc.topk_order(heat[i].reshape(-1), K))
        item, label = ds[i]
        assert item.dtype == torch.float64 and item.shape == (K, 3) and int(label) == i
        same += int(np.array_equal(item.numpy(), g[f"imp_top_K{K}_w{winF}"][i]))
        # ... and the rows are the reference's rows wherever the fp32 heat maps order the points identically
        ref_rows, ref_idx = orc.cloud_3d_importance_f64(x3, farr, tarr, i, K, winF, choice=1)
        agree = ds.indices(i) == ref_idx
        assert agree.mean() > 0.9
        assert np.array_equal(item.numpy()[agree], ref_rows[agree])
    assert same >= 1        # at least one whole item identical to the reference's (all of them unless near-ties reorder)


def test_random_k_is_a_uniform_subset_in_uniform_order(pca):
    dev = torch.device("cuda:0")
    n, nt, nf, K = 4000, 6, 20, 30
    P = nt * nf
    logmag = torch.randn(n, nt, nf, device=dev)
    farr, tarr = np.linspace(0, 0.5, nf), np.linspace(0, 0.1, nt)
    pts, idx = pca.random_points(logmag, farr, tarr, K, seed=5)
    idx_h = idx.cpu().numpy()
    assert idx_h.min() >= 0 and idx_h.max() < P
    assert all(len(set(r)) == K for r in idx_h[:200])                      # without replacement
    # rows are the cloud rows of those indices
    t, f = idx_h // nf, idx_h % nf
    ref = np.stack([farr.astype(np.float32)[f], tarr.astype(np.float32)[t],
                    logmag.cpu().numpy()[np.arange(n)[:, None], t, f]], axis=2)
    assert np.array_equal(pts.cpu().numpy(), ref)
    # every point is kept with probability K/P: chi-square over the P cells (dof P-1 = 119; mean 119, sd 15.4)
    counts = np.bincount(idx_h.reshape(-1), minlength=P).astype(np.float64)
    expect = n * K / P
    chi2 = ((counts - expect) ** 2 / (expect * (1 - K / P))).sum()
    assert chi2 < 119 + 6 * 15.4, chi2
    # the order is uniform too: the index at every output position has the mean of a uniform draw from 0..P-1
    pos_mean = idx_h.mean(axis=0)
    assert np.abs(pos_mean - (P - 1) / 2).max() < 5 * (P / np.sqrt(12)) / np.sqrt(n)
    # different seeds / clouds give different draws
    _, idx2 = pca.random_points(logmag, farr, tarr, K, seed=6)
    assert not np.array_equal(idx_h, idx2.cpu().numpy())
    assert not np.array_equal(idx_h[0], idx_h[1])


def test_multinomial_follows_the_heat_map(pca):
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(3)
    nf, nt, winF, K = 24, 8, 3, 200000
    x3 = rs.randn(nf, nt, 2).astype(np.float32)
    lm = _logmag(x3, dev)
    farr, tarr = np.linspace(0, 0.5, nf), np.linspace(0, 0.1, nt)
    pts, idx = pca.importance_points(lm, farr, tarr, K, winF, choice=0, seed=11)
    heat = pca.importance_heat(lm, winF).cpu().numpy().reshape(2, -1).astype(np.float64)
    idx_h = idx.cpu().numpy()
    for c in range(2):
        p = heat[c] / heat[c].sum()
        counts = np.bincount(idx_h[c], minlength=nf * nt)
        chi2 = ((counts - K * p) ** 2 / (K * p)).sum()
        dof = nf * nt - 1
        assert chi2 < dof + 6 * np.sqrt(2 * dof), (c, chi2)
    assert pts.shape == (2, K, 3)


def test_pc_randK_host_function(pca):
    rs = np.random.RandomState(8)
    x = rs.randn(64, 7).astype(np.float32)
    farr = np.linspace(0, 0.5, 64)
    xs, fs_ = pca.pc_randK(x, farr, 10, seed=3)
    assert xs.shape == (10, 7) and fs_.shape == (10, 7) and xs.dtype == x.dtype and fs_.dtype == farr.dtype
    for t in range(7):
        bins = np.round(fs_[:, t] / (farr[1] - farr[0])).astype(int)
        assert len(set(bins)) == 10
        assert np.array_equal(xs[:, t], x[bins, t])


def test_random_k_dataset_items(pca):
    dev = torch.device("cuda:0")
    g = dict(np.load(os.path.join(G, "sampling_golden.npz")))
    x3, farr, tarr = g["x3"], g["farr"], g["tarr"]
    ds = pca.ESC_pc_temp_randKSS(x3, np.arange(x3.shape[2]), farr, tarr, 50, device=dev)
    full = orc.cloud_3d_f64(x3, farr, tarr, 2)
    item, label = ds[2]
    assert item.dtype == torch.float64 and item.shape == (50, 3) and int(label) == 2
    assert np.array_equal(item.numpy(), full[ds.indices(2)])                 # rows of the reference's full cloud
    b3 = ds.cuda_batch([3]).cpu().numpy()[0]                                  # first access of item 3: same draw
    assert np.allclose(b3, orc.cloud_3d_f64(x3, farr, tarr, 3)[ds.indices(3)].astype(np.float32))
    before = ds.indices(2).copy()
    ds.resample()
    assert not np.array_equal(before, ds.indices(2))


# ------------------------------------------------------------------------------------ test-time resampling (parity unpinned)
@pytest.mark.parametrize("orig_sr,target_sr,L,res_type,scale", [(44100, 16000, 6000, "kaiser_fast", True),
                                                                (44100, 32000, 5001, "kaiser_fast", True),
                                                                (16000, 8000, 4000, "kaiser_fast", False),
                                                                (16000, 22050, 3000, "kaiser_best", True)])
def test_resample_matches_cpu_restatement(pca, orig_sr, target_sr, L, res_type, scale):
    """GPU resampler against the CPU restatement of librosa.resample / resampy.resample_f (oracle; both unpinned against the
    real resampy, which is not in the image): identical filter table, float64 accumulation on both sides."""
    dev = torch.device("cuda:0")
    x = orc.synth_audio(3, L, orig_sr, seed=7)
    got = pca.resample(torch.from_numpy(x).to(dev), orig_sr, target_sr, res_type=res_type, scale=scale).cpu().numpy()
    ref = orc.resample_librosa080(x, orig_sr, target_sr, res_type=res_type, fix=True, scale=scale)
    assert got.shape == ref.shape == (3, int(np.ceil(L * target_sr / orig_sr)))
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    same = pca.resample(torch.from_numpy(x).to(dev), orig_sr, orig_sr)
    assert same.shape == (3, L)


def test_random_k_dataset_redraws_every_pass(pca):
    """the reference draws a fresh subset on every __getitem__; here the batched draw is renewed whenever an item is
    requested again (a new pass), and instances without an explicit seed follow numpy's global generator."""
    dev = torch.device("cuda:0")
    g = dict(np.load(os.path.join(G, "sampling_golden.npz")))
    x3, farr, tarr = g["x3"], g["farr"], g["tarr"]
    n = x3.shape[2]
    np.random.seed(5)
    ds = pca.ESC_pc_temp_randKSS(x3, np.arange(n), farr, tarr, 50, device=dev)
    first = [ds[i][0].numpy().copy() for i in range(n)]             # pass 1: one draw serves every item
    for i in range(n):                                              # rows of the reference's full cloud, current draw
        assert np.array_equal(first[i], orc.cloud_3d_f64(x3, farr, tarr, i)[ds.indices(i)])
    second = [ds[i][0].numpy().copy() for i in range(n)]            # pass 2: item 0 is requested again -> new draw
    assert any(not np.array_equal(a, b) for a, b in zip(first, second))
    np.random.seed(5)
    ds2 = pca.ESC_pc_temp_randKSS(x3, np.arange(n), farr, tarr, 50, device=dev)
    assert np.array_equal(ds2[0][0].numpy(), first[0])              # np.random.seed governs the sequence
    ds3 = pca.ESC_pc_temp_randKSS(x3, np.arange(n), farr, tarr, 50, device=dev)
    assert not np.array_equal(ds3[0][0].numpy(), first[0])          # a second instance gets its own draw


def test_topk_float64_inputs_follow_the_float64_order(pca):
    """float64 magnitudes that collapse into float32 ties must still come out in the reference's float64
    (-x).argsort() order (lowest index first among exact ties)."""
    rng = np.random.default_rng(0)
    base = rng.standard_normal((40, 3)).astype(np.float32).astype(np.float64)
    x = np.repeat(base, 4, axis=0)                                    # 160 bins, groups of 4 equal in float32 ...
    x += 1e-11 * rng.standard_normal(x.shape)                         # ... but distinct in float64
    farr = np.linspace(0, 0.5, x.shape[0])
    for K in (1, 7, 33, 160):
        xs, fs_ = pca.pc_maxK(x, farr, K)
        for t in range(x.shape[1]):
            order = (-x[:, t]).argsort(kind="stable")[:K]
            assert np.array_equal(xs[:, t], x[order, t]) and np.array_equal(fs_[:, t], farr[order])
    # 3-D dataset class
    x3 = np.stack([x[:, :2] for _ in range(3)], axis=2) + 1e-12 * rng.standard_normal((160, 2, 3))
    tarr = np.array([0.0, 0.1])
    ds = pca.ESC_pc_temp_maxKSS(x3, np.arange(3), farr, tarr, 25, device=torch.device("cuda:0"))
    for i in range(3):
        flat = x3[:, :, i].T.reshape(-1)
        assert np.array_equal(ds.indices(i), (-flat).argsort(kind="stable")[:25])
        assert np.array_equal(ds[i][0].numpy()[:, 2], flat[ds.indices(i)])
