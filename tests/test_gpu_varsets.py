"""GPU tests of the variable-size-set extensions (SURVEY.md 8c "oracle for extensions"): threshold / capped selection
into padded sets, and the masked encoders.  Oracle construction as the survey prescribes: for each sample b the
reference arithmetic (oracle port, pinned against the reference) is run on X[b:b+1, :count[b]] and the results are
stacked; all-valid masks must reproduce the unmasked path bit for bit."""
import os

import numpy as np
import pytest
import torch

from oracle import pcaudio_oracle as orc

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def pca():
    import __graft_entry__ as g
    g.build()
    import pcaudio_b200
    return pcaudio_b200


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# ------------------------------------------------------------------------------------ selection
@pytest.mark.parametrize("nf,nt,K,q", [(512, 10, 5120, 0.5), (512, 10, 256, 0.9), (512, 10, 256, 0.99), (513, 1, 513, 0.3),
                                       (512, 32, 8192, 0.7), (64, 3, 100, 1.5), (64, 3, 100, -1.0)])
@pytest.mark.parametrize("sorted_desc", [True, False])
def test_threshold_selection_bit_exact(pca, nf, nt, K, q, sorted_desc):
    """keep key >= tau, capped at K by the (-key).argsort(kind='stable')[:K] rule, zero padding, counts."""
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(nf * 31 + nt * 7 + K)
    n = 5
    keys = rng.normal(size=(n, nt, nf)).astype(np.float32)
    keys[1, 0, :40] = keys[1, 0, 40]                 # a run of ties
    if q > 1.0:
        tau = float(keys.max()) + 1.0               # nothing passes
    elif q < 0.0:
        tau = float(keys.min()) - 1.0               # everything passes: pure top-K
    else:
        tau = float(np.quantile(keys, q))
    farr, tarr = orc.coord_tables(16000.0, nf, 2 * nf, 0.5, nt)
    tarr_use = tarr if nt > 1 else None
    pts, idx, counts = pca.select_points(torch.from_numpy(keys).to(dev), farr, tarr_use, K, threshold=tau, sorted_desc=sorted_desc)
    pts, idx, counts = pts.cpu().numpy(), idx.cpu().numpy(), counts.cpu().numpy()
    width = 3 if nt > 1 else 2
    for c in range(n):
        flat = keys[c].reshape(-1)
        order = orc.topk_order(flat, min(K, flat.size))
        keep = order[flat[order] >= np.float32(tau)]
        if not sorted_desc:
            keep = np.sort(keep)
        assert counts[c] == keep.size, (c, counts[c], keep.size)
        assert np.array_equal(idx[c, :keep.size], keep.astype(np.int32))
        assert (idx[c, keep.size:] == -1).all() and (pts[c, keep.size:] == 0).all()
        f = (farr[keep % nf]).astype(np.float32)
        assert np.array_equal(pts[c, :keep.size, 0], f)
        assert np.array_equal(pts[c, :keep.size, width - 1], flat[keep])
        if width == 3:
            assert np.array_equal(pts[c, :keep.size, 1], tarr[keep // nf].astype(np.float32))


def test_select_without_threshold_equals_topk(pca):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    keys = torch.randn(4, 10, 512, generator=g).to(dev)
    farr, tarr = orc.coord_tables(16000.0, 512, 1024, 0.5, 10)
    p0, i0 = pca.topk_points(keys, farr, tarr, 300)
    p1, i1, c1 = pca.select_points(keys, farr, tarr, 300, threshold=None)
    assert torch.equal(p0, p1) and torch.equal(i0, i1) and (c1 == 300).all()


# ------------------------------------------------------------------------------------ masked encoders
def _st_pair(pca, d_in, precision):
    dev = torch.device("cuda:0")
    torch.manual_seed(17 + d_in)
    st = pca.ST(dim_input=d_in, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev)
    st.set_precision(precision)
    params = {k: v.detach().cpu() for k, v in st.state_dict().items()}
    return st, params, dev


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 2e-2)])
@pytest.mark.parametrize("d_in,N,counts", [(2, 300, [300, 1, 129, 64, 257, 128]), (3, 2500, [2500, 2049, 2048, 17]),
                                           (2, 1025, [1025, 1024, 513, 5])])
def test_masked_st_matches_per_sample_reference(pca, precision, tol, d_in, N, counts):
    st, params, dev = _st_pair(pca, d_in, precision)
    B = len(counts)
    g = torch.Generator().manual_seed(N + d_in)
    X = torch.randn(B, N, d_in, generator=g)
    Xpad = X.clone()
    for b, c in enumerate(counts):
        Xpad[b, c:] = float("nan")          # padding must never be read as a key (NaN would poison the logits)
    cnt = torch.tensor(counts, dtype=torch.int32, device=dev)
    with torch.no_grad():
        out = st(Xpad.to(dev), counts=cnt).cpu().numpy().reshape(B, 10)
    assert np.isfinite(out).all()
    ref = np.stack([orc.st_forward(params, X[b:b + 1, :c], 8).reshape(10).numpy() for b, c in enumerate(counts)])
    assert rel_err(out, ref) < tol, rel_err(out, ref)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_all_valid_mask_is_bit_identical(pca, precision):
    st, _, dev = _st_pair(pca, 2, precision)
    X = torch.randn(6, 700, 2, generator=torch.Generator().manual_seed(1)).to(dev)
    with torch.no_grad():
        a = st(X)
        b = st(X, counts=torch.full((6,), 700, dtype=torch.int32, device=dev))
    assert torch.equal(a, b)


@pytest.mark.parametrize("pool", ["mean", "max", "sum"])
def test_masked_deepset_pool(pca, pool):
    """north_star's PointNet-style shared MLP + masked pool (DeepSet / SmallDeepSet stand-in)."""
    dev = torch.device("cuda:0")
    torch.manual_seed(4)
    ds = pca.DeepSet(3, 1, 10, dim_hidden=64, pool=pool).to(dev)
    params = {k: v.detach().cpu() for k, v in ds.state_dict().items()}
    counts = [400, 1, 77, 399]
    X = torch.randn(4, 400, 3, generator=torch.Generator().manual_seed(8))
    Xpad = X.clone()
    for b, c in enumerate(counts):
        Xpad[b, c:] = 1e30
    with torch.no_grad():
        out = ds(Xpad.to(dev), counts=torch.tensor(counts, dtype=torch.int32, device=dev)).cpu().numpy().reshape(4, 10)
    ref = np.stack([orc.deepset_forward(params, X[b:b + 1, :c], 1, 10, pool).reshape(10).numpy() for b, c in enumerate(counts)])
    assert rel_err(out, ref) < 1e-3


# ------------------------------------------------------------------------------------ whole path with a threshold
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 2e-2)])
def test_pipeline_threshold_mode(pca, precision, tol):
    """audio -> log-magnitudes -> points above a magnitude threshold (capped at top_k) -> masked ST, against the oracle
    recipe applied per cloud on exactly the points the GPU selection kept (selection itself is tested bit-exactly above)."""
    dev = torch.device("cuda:0")
    w = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "3st_weights.npz")).items()}
    st = pca.ST(dim_input=3, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev)
    st.load_state_dict(w)
    params = orc.strip_module_prefix(w)
    audio = torch.from_numpy(orc.synth_audio(3, 16000, 16000.0, seed=77)).to(dev)
    tau, K = -6.0, 1500
    cfg = pca.AudioConfig(sampling_rate=16000, window_size=1024, n_samples=16000, mode=3, Ntemp=10, top_k=K,
                          precision=precision, threshold=tau)
    pipe = pca.AudioSetPipeline(st, cfg, dev)
    logits = pipe(audio).cpu().numpy()
    pts, counts, _ = pca.spectral_point_cloud(audio, n_fft=1024, sr=16000, ntemp=10, top_k=K, threshold=tau)
    pts, counts = pts.cpu(), counts.cpu().numpy()
    assert logits.shape == (9, 10) and pts.shape == (9, K, 3)
    assert (counts >= 1).all() and (counts < K).any(), counts          # the threshold actually bites
    ref = np.stack([orc.st_forward(params, pts[b:b + 1, :counts[b]], 8).reshape(10).numpy() for b in range(9)])
    assert rel_err(logits, ref) < tol, rel_err(logits, ref)


# ------------------------------------------------------------------------------------ fused front end
@pytest.mark.parametrize("n_fft,L,ntemp,K,tau", [(1024, 16000, None, 1024, None), (1024, 16000, None, 8192, None),
                                                 (1024, 16000, 10, 5120, None), (1024, 16000, 10, 300, -5.0),
                                                 (256, 4000, 7, 64, None), (2048, 16000, 5, 1000, -7.5),
                                                 (1024, 16000, None, 8192, -6.0)])
def test_fused_frontend_equals_unfused(pca, n_fft, L, ntemp, K, tau):
    """audio -> points in one launch (log-magnitudes held in shared memory) must equal STFT kernel + selection kernel
    bit for bit: same per-frame arithmetic, same selection rule."""
    dev = torch.device("cuda:0")
    audio = torch.from_numpy(orc.synth_audio(5, L, 16000.0, seed=n_fft + K)).to(dev)
    kw = dict(n_fft=n_fft, sr=16000.0, ntemp=ntemp, top_k=K, threshold=tau)
    p1, c1, i1 = pca.spectral_point_cloud(audio, fused=True, **kw)
    p0, c0, i0 = pca.spectral_point_cloud(audio, fused=False, **kw)
    assert p1.shape == p0.shape and torch.equal(i1, i0) and torch.equal(p1, p0) and torch.equal(c1.cpu(), c0.cpu())


def test_fused_frontend_rejects_oversized_clouds(pca):
    dev = torch.device("cuda:0")
    audio = torch.zeros(1, 64000, device=dev)
    with pytest.raises(RuntimeError, match="shared memory"):
        pca.spectral_point_cloud(audio, n_fft=1024, sr=16000.0, ntemp=None, top_k=8192, fused=True)     # 64512-point cloud


def test_empty_sets_give_nan_logits(pca):
    """counts[b] == 0 must not be encoded silently as a one-point cloud; the other clouds are unaffected."""
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    model = pca.ST(dim_input=3, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev).eval()
    X = torch.randn(5, 300, 3, device=dev)
    counts = torch.tensor([300, 0, 17, 0, 129], dtype=torch.int32, device=dev)
    for prec in ("fp32", "bf16"):
        model.set_precision(prec)
        with torch.no_grad():
            out = model(X, counts)
            ref = model(X, counts.clamp(min=1))
        assert torch.isnan(out[[1, 3]]).all()
        assert torch.equal(out[[0, 2, 4]], ref[[0, 2, 4]]) and torch.isfinite(ref).all()
