"""GPU parity tests (run on the B200 box): every call goes through the C ABI (ctypes) and is compared
with the CPU oracle / the golden vectors produced by the unmodified reference.

Bars (BASELINE.json north_star): selection / compaction indices bit-exact when fed the reference's
magnitudes; STFT magnitudes 1e-5 relative (to the per-clip max |S|); encoder outputs
1e-3 relative in fp32."""
import os

import numpy as np
import pytest
import torch

from oracle import pcaudio_oracle as orc

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

STFT_REL_TOL = 1e-5      # |S| vs float64 restatement, relative to max |S| of the clip
ENC_REL_TOL = 1e-3       # fp32 encoder outputs / logits, relative to max |ref|


@pytest.fixture(scope="module")
def pca():
    import __graft_entry__ as g
    g.build()
    import pcaudio_b200
    assert torch.cuda.is_available()
    return pcaudio_b200


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def sub(d, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in d.items() if k.startswith(prefix)}


# ------------------------------------------------------------------------------------ STFT
@pytest.mark.parametrize("n_fft,win,L,drop", [(1024, 1024, 16000, True), (2048, 2048, 16000, False),
                                              (1024, 1024, 64000, True), (2048, 1434, 9000, False),
                                              (512, 512, 4000, False), (256, 204, 3000, True),
                                              (4096, 4096, 16000, False)])
def test_stft_logmag_matches_oracle(pca, dev, n_fft, win, L, drop):
    audio = orc.synth_audio(3, L, 16000, seed=101)
    got = pca.stft_logmag(torch.from_numpy(audio).to(dev), n_fft, win, 0.5, drop_nyquist=drop).cpu().numpy()
    for c in range(audio.shape[0]):
        ref = orc.logmag_recipe(audio[c], n_fft, 0.5, win_length=win, drop_nyquist=drop)      # (Nf, Nt)
        assert got[c].shape == ref.T.shape
        mag_ref = np.exp(ref.astype(np.float64)) - 1e-8
        mag_got = np.exp(got[c].T.astype(np.float64)) - 1e-8
        assert np.abs(mag_got - mag_ref).max() / mag_ref.max() < STFT_REL_TOL
        strong = mag_ref > 1e-3 * mag_ref.max()          # log-magnitudes agree where the bin is not empty
        assert np.abs(got[c].T[strong] - ref[strong]).max() < 1e-3


@pytest.mark.parametrize("n_fft,L,frames", [(512, 7000, None), (1024, 16000, None), (1024, 16000, 30), (2048, 9000, None),
                                            (4096, 16000, 5), (2048, 16000, 1)])
def test_stft_specialised_kernel_equals_generic(pca, dev, n_fft, L, frames):
    """n_fft in [512, 4096] runs a size-specialised kernel (first butterfly fed from global memory, window and first-stage
    twiddles in registers, G frames in flight per block); it must reproduce the generic kernel -- the code the fused front end
    runs -- bit for bit, edge frames (reflect padding) and odd clip offsets included."""
    from pcaudio_b200 import _lib
    audio = torch.from_numpy(orc.synth_audio(3, L + 1, 16000, seed=n_fft)).to(dev)
    for a in (audio[:, :L], audio[:, 1:]):               # second view: clips start at odd float offsets (no 8-byte loads)
        a = a if a.is_contiguous() else a.contiguous()
        fast = pca.stft_logmag(a, n_fft, drop_nyquist=False, n_frames=frames)
        _lib.lib().pca_debug_set_stft_generic(1)
        try:
            slow = pca.stft_logmag(a, n_fft, drop_nyquist=False, n_frames=frames)
        finally:
            _lib.lib().pca_debug_set_stft_generic(0)
        assert torch.isfinite(fast).all() and torch.equal(fast, slow)


def test_stft_frame_limit_and_errors(pca, dev):
    audio = torch.from_numpy(orc.synth_audio(2, 16000, 16000, seed=3)).to(dev)
    full = pca.stft_logmag(audio, 1024, drop_nyquist=True)
    part = pca.stft_logmag(audio, 1024, drop_nyquist=True, n_frames=30)
    assert full.shape == (2, 32, 512) and part.shape == (2, 30, 512)
    assert torch.equal(full[:, :30], part)
    with pytest.raises(RuntimeError, match="power of two"):
        pca.stft_logmag(audio, 1000)
    with pytest.raises(RuntimeError, match="nt_out"):
        pca.stft_logmag(audio, 1024, n_frames=40)
    empty = pca.stft_logmag(audio[:0], 1024)
    assert empty.shape == (0, 32, 513)


# ------------------------------------------------------------------------------------ clouds / selection
def test_dataset_classes_match_reference_golden(pca, dev):
    g = dict(np.load(os.path.join(G, "pointcloud_golden.npz")))
    T = g["x3"].shape[2]
    ds = pca.ESC_pc_temp(g["x3"], np.arange(T), g["farr"], g["tarr"], device=dev)
    assert len(ds) == T
    for i in range(T):
        pc, lbl = ds[i]
        assert pc.dtype == torch.float32 and int(lbl) == i
        np.testing.assert_array_equal(pc.numpy(), g["pc_temp"][i])
    np.testing.assert_array_equal(ds.cuda_batch([0, 3]).cpu().numpy(), g["pc_temp"][[0, 3]])
    for K in (1, 17, 64, 288):
        dk = pca.ESC_pc_temp_maxKSS(g["x3"], np.arange(T), g["farr"], g["tarr"], K, device=dev)
        for i in range(T):
            pc, _ = dk[i]
            assert pc.dtype == torch.float64
            np.testing.assert_array_equal(pc.numpy(), g[f"pc_temp_maxk_{K}"][i])          # bit exact
        np.testing.assert_array_equal(dk.cuda_batch(list(range(T))).cpu().numpy(),
                                      g[f"pc_temp_maxk_{K}"].astype(np.float32))
    d2 = pca.ESC_pc(g["x2"], np.arange(9), g["farr"], device=dev)
    for i in range(9):
        np.testing.assert_array_equal(d2[i][0].numpy(), g["pc_2d"][i])
    for K in (1, 10, 48):
        xs, fs_ = pca.pc_maxK(g["x2"], g["farr"], K, device=dev)
        assert xs.dtype == g[f"pc_maxK_x_{K}"].dtype and fs_.dtype == g[f"pc_maxK_f_{K}"].dtype
        np.testing.assert_array_equal(xs, g[f"pc_maxK_x_{K}"])
        np.testing.assert_array_equal(fs_, g[f"pc_maxK_f_{K}"])
        dss = pca.ESC_pc_ss(xs, np.arange(9), fs_, device=dev)
        for i in range(9):
            np.testing.assert_array_equal(dss[i][0].numpy(), g[f"pc_ss_{K}"][i])


@pytest.mark.parametrize("nf,nt,K", [(512, 10, 256), (512, 10, 5120), (512, 32, 8192), (512, 32, 1),
                                     (512, 126, 8192), (1025, 1, 501), (1025, 1, 1025), (37, 3, 50)])
def test_topk_indices_bit_exact(pca, dev, nf, nt, K):
    rs = np.random.RandomState(nf * 31 + nt)
    n = 5
    keys = (rs.randn(n, nt, nf) * 3 - 8).astype(np.float32)
    farr, tarr = orc.coord_tables(16000, nf, 2 * nf, 0.5, nt)
    pts, idx = pca.topk_points(torch.from_numpy(keys).to(dev), farr, tarr, K, sorted_desc=True)
    idx, pts = idx.cpu().numpy(), pts.cpu().numpy()
    for c in range(n):
        flat = keys[c].reshape(-1)
        order = orc.topk_order(flat, K)
        np.testing.assert_array_equal(idx[c], order)
        np.testing.assert_array_equal(pts[c, :, 2], flat[order])
        np.testing.assert_array_equal(pts[c, :, 0], farr.astype(np.float32)[order % nf])
        np.testing.assert_array_equal(pts[c, :, 1], tarr.astype(np.float32)[order // nf])
    # scan-order emission: same set, ascending flat index
    _, idx_scan = pca.topk_points(torch.from_numpy(keys).to(dev), farr, tarr, K, sorted_desc=False)
    for c in range(n):
        np.testing.assert_array_equal(idx_scan[c].cpu().numpy(), np.sort(orc.topk_order(keys[c].reshape(-1), K)))


def test_topk_ties_keep_lowest_indices(pca, dev):
    rs = np.random.RandomState(5)
    keys = rs.randint(-4, 4, size=(6, 8, 257)).astype(np.float32)      # massive ties, incl. +-0
    keys[0, 0, :5] = -0.0
    farr, tarr = orc.coord_tables(16000, 257, 512, 0.5, 8)
    for K in (1, 100, 1000, 2056):
        _, idx = pca.topk_points(torch.from_numpy(keys).to(dev), farr, tarr, K)
        for c in range(keys.shape[0]):
            np.testing.assert_array_equal(idx[c].cpu().numpy(), orc.topk_order(keys[c].reshape(-1), K))


@pytest.mark.parametrize("case", ["all_equal", "narrow", "two_values", "wide", "heavy_bin", "signs"])
def test_topk_key_range_edge_cases(pca, dev, case):
    """The register-key kernel selects on range-normalised keys ((o - omin) << clz(omax - omin)) and moves the candidates of
    the chosen first digit to a bounded shared list: degenerate ranges, overflow of that list and mixed signs must give the
    oracle's order bit for bit."""
    rs = np.random.RandomState(11)
    nf, nt, n = 512, 32, 4
    if case == "all_equal":
        keys = np.full((n, nt, nf), -3.25, np.float32)
    elif case == "narrow":                       # neighbours in float32: only the low mantissa bits differ
        base = np.float32(-7.5).view(np.uint32)
        keys = (base + rs.randint(0, 37, size=(n, nt, nf)).astype(np.uint32)).view(np.float32)
    elif case == "two_values":
        keys = np.where(rs.rand(n, nt, nf) < 0.5, np.float32(1.0), np.float32(2.0)).astype(np.float32)
    elif case == "wide":
        keys = (rs.randn(n, nt, nf) * 1e3).astype(np.float32)
        keys[:, 0, :3] = [np.float32(3e38), np.float32(-3e38), np.float32(1e-30)]
    elif case == "heavy_bin":                    # > TOPK_CAND keys share the K-th key's first digit: register fallback
        keys = (rs.randn(n, nt, nf) * 1e-3 - 8).astype(np.float32)
        keys[:, :, :4] = rs.randn(n, nt, 4).astype(np.float32) * 50
    else:
        keys = rs.randn(n, nt, nf).astype(np.float32)
        keys[:, 1, :7] = [0.0, -0.0, 1e-45, -1e-45, 0.0, -0.0, 0.0]
    farr, tarr = orc.coord_tables(16000, nf, 2 * nf, 0.5, nt)
    for K in (1, 7, 256, 1500, 8192, nf * nt - 1):
        for srt in (True, False):
            _, idx = pca.topk_points(torch.from_numpy(keys).to(dev), farr, tarr, K, sorted_desc=srt)
            for c in range(n):
                want = orc.topk_order(keys[c].reshape(-1), K)
                np.testing.assert_array_equal(idx[c].cpu().numpy(), want if srt else np.sort(want))


def test_topk_errors(pca, dev):
    keys = torch.zeros(2, 4, 64, device=dev)
    with pytest.raises(RuntimeError, match="outside"):
        pca.topk_points(keys, np.zeros(64), np.zeros(4), 257)
    big = torch.zeros(1, 40, 512, device=dev)
    with pytest.raises(RuntimeError, match="16384"):
        pca.topk_points(big, np.zeros(512), np.zeros(40), 20000)


def test_spectral_point_cloud_matches_recipe(pca, dev):
    fs, n_fft = 16000, 1024
    audio = orc.synth_audio(2, 16000, fs, seed=202)
    # magnitudes come from the GPU STFT; feed THOSE to the oracle's selection for the bit-exact check
    logmag = pca.stft_logmag(torch.from_numpy(audio).to(dev), n_fft, drop_nyquist=True, n_frames=30)
    pts, counts, idx = pca.spectral_point_cloud(torch.from_numpy(audio).to(dev), n_fft=n_fft, sr=fs, ntemp=10, top_k=300)
    assert pts.shape == (6, 300, 3) and counts.tolist() == [300] * 6
    lm = logmag.cpu().numpy().reshape(6, 10 * 512)
    farr, tarr = orc.coord_tables(fs, 512, n_fft, 0.5, 10)
    for c in range(6):
        order = orc.topk_order(lm[c], 300)
        np.testing.assert_array_equal(idx[c].cpu().numpy(), order)
        np.testing.assert_array_equal(pts[c, :, 2].cpu().numpy(), lm[c][order])
    # full clouds against the oracle recipe (magnitudes to STFT tolerance, coordinates exact)
    full, _, _ = pca.spectral_point_cloud(torch.from_numpy(audio).to(dev), n_fft=n_fft, sr=fs, ntemp=10)
    ref = np.concatenate([orc.clip_chunk_clouds(audio[c], fs, n_fft, 0.5, 10) for c in range(2)])
    np.testing.assert_array_equal(full[:, :, :2].cpu().numpy(), ref[:, :, :2])
    mag_ref, mag_got = np.exp(ref[:, :, 2].astype(np.float64)), np.exp(full[:, :, 2].cpu().numpy().astype(np.float64))
    assert np.abs(mag_got - mag_ref).max() / mag_ref.max() < STFT_REL_TOL


# ------------------------------------------------------------------------------------ encoder (fp32)
def test_blocks_match_reference_golden(pca, dev):
    g = dict(np.load(os.path.join(G, "encoder_golden.npz")))
    mab = pca.MAB(5, 7, 16, 4).to(dev)
    mab.load_state_dict(sub(g, "mab."))
    out = mab(torch.from_numpy(g["mab_Q"]).to(dev), torch.from_numpy(g["mab_K"]).to(dev))
    assert rel_err(out.detach().cpu().numpy(), g["mab_out"]) < ENC_REL_TOL
    mabln = pca.MAB(5, 7, 16, 4, ln=True).to(dev)
    mabln.load_state_dict(sub(g, "mabln."))
    out = mabln(torch.from_numpy(g["mab_Q"]).to(dev), torch.from_numpy(g["mab_K"]).to(dev))
    assert rel_err(out.detach().cpu().numpy(), g["mabln_out"]) < ENC_REL_TOL
    isab = pca.ISAB(3, 16, 4, 8).to(dev)
    isab.load_state_dict(sub(g, "isab."))
    assert rel_err(isab(torch.from_numpy(g["isab_X"]).to(dev)).detach().cpu().numpy(), g["isab_out"]) < ENC_REL_TOL
    pma = pca.PMA(16, 4, 2).to(dev)
    pma.load_state_dict(sub(g, "pma."))
    assert rel_err(pma(torch.from_numpy(g["pma_X"]).to(dev)).detach().cpu().numpy(), g["pma_out"]) < ENC_REL_TOL
    sab = pca.SAB(6, 16, 2).to(dev)
    sab.load_state_dict(sub(g, "sab."))
    assert rel_err(sab(torch.from_numpy(g["sab_X"]).to(dev)).detach().cpu().numpy(), g["sab_out"]) < ENC_REL_TOL
    ds = pca.DeepSet(3, 2, 5, dim_hidden=32).to(dev)
    ds.load_state_dict(sub(g, "ds."))
    out = ds(torch.from_numpy(g["ds_X"]).to(dev))
    assert out.shape == (3, 2, 5) and rel_err(out.detach().cpu().numpy(), g["ds_out"]) < ENC_REL_TOL
    mn = pca.SetTransformer(dim_hidden=64, num_heads=4, num_inds=16).to(dev).eval()
    mn.load_state_dict(sub(g, "mn."))
    assert rel_err(mn(torch.from_numpy(g["mn_X"]).to(dev)).detach().cpu().numpy(), g["mn_out"]) < ENC_REL_TOL


@pytest.mark.parametrize("d_in", [2, 3])
def test_st_matches_reference_golden(pca, dev, d_in):
    g = dict(np.load(os.path.join(G, "encoder_golden.npz")))
    st = pca.ST(dim_input=d_in, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev)
    st.load_state_dict(sub(g, f"st{d_in}."))
    X = torch.from_numpy(g[f"st{d_in}_X"]).to(dev)
    with torch.no_grad():
        out = st(X)
        out1 = st(X[:1])
    assert out.shape == (3, 10) and out1.shape == (10,)               # .squeeze() quirk kept
    assert rel_err(out.cpu().numpy(), g[f"st{d_in}_out"]) < ENC_REL_TOL
    assert rel_err(out1.cpu().numpy(), g[f"st{d_in}_out_b1"]) < ENC_REL_TOL
    assert st(X[:0]).shape == (0, 10) or st(X[:0]).numel() == 0       # empty batch


@pytest.mark.parametrize("tag,d_in", [("fst", 2), ("3st", 3)])
def test_shipped_checkpoints_match_reference(pca, dev, tag, d_in):
    w = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, f"{tag}_weights.npz")).items()}
    ck = np.load(os.path.join(G, "checkpoint_golden.npz"))
    st = pca.ST(dim_input=d_in, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev)
    st.load_state_dict(w)
    with torch.no_grad():
        out = st(torch.from_numpy(ck[f"{tag}_X"]).to(dev))
    assert rel_err(out.cpu().numpy(), ck[f"{tag}_out"]) < ENC_REL_TOL


def test_modelnet_config5_dims_vs_oracle(pca, dev):
    torch.manual_seed(3)
    mn = pca.SetTransformer(dim_hidden=256, num_heads=4, num_inds=16).to(dev).eval()
    X = torch.randn(6, 1000, 3)
    X = (X - X.mean(dim=(1, 2), keepdim=True)) / X.std(dim=(1, 2), keepdim=True)   # standardize per cloud
    with torch.no_grad():
        out = mn(X.to(dev)).cpu()
    ref = orc.modelnet_forward({k: v.cpu() for k, v in mn.state_dict().items()}, X, 4)
    assert out.shape == (6, 40) and rel_err(out.numpy(), ref.numpy()) < ENC_REL_TOL


def test_deepset_pools_vs_oracle(pca, dev):
    torch.manual_seed(4)
    X = torch.randn(5, 300, 3)
    for pool in ("mean", "max", "sum"):
        ds = pca.DeepSet(3, 1, 7, dim_hidden=128, pool=pool).to(dev)
        ref = orc.deepset_forward({k: v.cpu() for k, v in ds.state_dict().items()}, X, 1, 7, pool)
        with torch.no_grad():
            out = ds(X.to(dev)).cpu()
        assert rel_err(out.numpy(), ref.numpy()) < ENC_REL_TOL


def test_configurations_without_training_kernels_fail_loudly(pca, dev):
    """Everything trains (tests/test_gpu_train.py) except LayerNorm models on variable-size sets: loud failure, no silent
    zero gradients."""
    st = pca.ST(dim_input=2, num_outputs=1, dim_output=10, num_inds=8, dim_hidden=16, num_heads=4, ln=True).to(dev)
    with torch.enable_grad(), pytest.raises(NotImplementedError):
        st(torch.randn(2, 50, 2, device=dev), counts=torch.tensor([50, 20], dtype=torch.int32, device=dev))


# ------------------------------------------------------------------------------------ whole path
def _load_ckpt(pca, dev, tag, d_in):
    w = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, f"{tag}_weights.npz")).items()}
    st = pca.ST(dim_input=d_in, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev)
    st.load_state_dict(w)
    return st, orc.strip_module_prefix(w)


def test_pipeline_fst_audio_to_logits(pca, dev):
    """BASELINE config 2 shape (FST: 2048/1024, 16 frame clouds of 1025 points per 1 s clip), 3 clips."""
    st, w = _load_ckpt(pca, dev, "fst", 2)
    audio = orc.synth_audio(3, 16000, 16000, seed=202)
    pipe = pca.AudioSetPipeline(st, pca.AudioConfig(window_size=2048, n_samples=16000, mode=2), dev)
    assert pipe.clouds_per_clip == 16 and pipe.points_per_cloud == 1025
    logits = pipe(torch.from_numpy(audio).to(dev)).cpu().numpy()
    ref = np.concatenate([orc.st_forward(w, torch.from_numpy(orc.clip_frame_clouds(audio[c], 16000, 2048)), 8).numpy()
                          for c in range(3)])
    assert logits.shape == (48, 10) and rel_err(logits, ref) < ENC_REL_TOL
    # host-buffer entry point gives the same numbers
    host = pipe.run_host(torch.from_numpy(audio).pin_memory())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(host.numpy().reshape(48, 10), logits)


def test_pipeline_3st_chunks_and_topk(pca, dev):
    """BASELINE config 1/3 shape (3ST: 1024/512, 10-frame chunk clouds of 5120 points)."""
    st, w = _load_ckpt(pca, dev, "3st", 3)
    audio = orc.synth_audio(2, 16000, 16000, seed=101)
    pipe = pca.AudioSetPipeline(st, pca.AudioConfig(window_size=1024, n_samples=16000, mode=3, Ntemp=10), dev)
    assert pipe.clouds_per_clip == 3 and pipe.points_per_cloud == 5120
    logits = pipe(torch.from_numpy(audio).to(dev)).cpu().numpy()
    ref = np.concatenate([orc.st_forward(w, torch.from_numpy(orc.clip_chunk_clouds(audio[c], 16000, 1024, 0.5, 10)), 8).numpy()
                          for c in range(2)])
    assert rel_err(logits, ref) < ENC_REL_TOL
    pipe_k = pca.AudioSetPipeline(st, pca.AudioConfig(window_size=1024, n_samples=16000, mode=3, Ntemp=10, top_k=1024), dev)
    logits_k = pipe_k(torch.from_numpy(audio).to(dev)).cpu().numpy()
    ref_k = np.concatenate([orc.st_forward(w, torch.from_numpy(orc.clip_chunk_clouds(audio[c], 16000, 1024, 0.5, 10, top_k=1024)), 8).numpy()
                            for c in range(2)])
    assert rel_err(logits_k, ref_k) < ENC_REL_TOL


def test_full_size_properties(pca, dev):
    """Size-independent checks at BASELINE sizes: batch of 256 clips (FST, 4096 clouds) -- sharding the
    batch gives bit-identical per-clip logits; permuting the points of a cloud leaves logits unchanged to
    fp32 tolerance; top-K output is sorted and is a subset."""
    st, _ = _load_ckpt(pca, dev, "fst", 2)
    g = torch.Generator().manual_seed(9)
    audio = (0.1 * torch.randn(256, 16000, generator=g)).to(dev)
    pipe = pca.AudioSetPipeline(st, pca.AudioConfig(window_size=2048, n_samples=16000, mode=2), dev)
    full = pipe(audio)
    assert full.shape == (4096, 10) and torch.isfinite(full).all()
    halves = torch.cat([pipe(audio[:128]), pipe(audio[128:])])
    assert torch.equal(full, halves)
    st3, _ = _load_ckpt(pca, dev, "3st", 3)
    pts, _, idx = pca.spectral_point_cloud(audio[:4, :16000], n_fft=1024, sr=16000, ntemp=None, top_k=8192)
    assert pts.shape == (4, 8192, 3)
    mags = pts[:, :, 2]
    assert (mags[:, :-1] >= mags[:, 1:]).all()
    assert all(len(set(r.tolist())) == 8192 for r in idx.cpu())
    with torch.no_grad():
        a = st3(pts)
        b = st3(pts[:, torch.randperm(8192, device=dev)])
    assert rel_err(b.cpu().numpy(), a.cpu().numpy()) < ENC_REL_TOL


# ------------------------------------------------------------------------------------ generic SetTransformer (SAB decoder)
@pytest.mark.parametrize("tag", ["noln", "ln"])
def test_generic_set_transformer_sab_decoder_matches_reference_golden(pca, dev, tag):
    """set_transformer-master/models.py:30-44 (ISAB, ISAB -> PMA(k seeds) -> SAB, SAB -> Linear), with and without
    LayerNorm, against outputs of the unmodified reference (tests/golden/make_golden_stmodels.py)."""
    from pcaudio_b200 import st_models
    g = dict(np.load(os.path.join(G, "stmodels_golden.npz")))
    k = g[f"{tag}_Y"].shape[1]
    m = st_models.SetTransformer(2, k, 6, num_inds=8, dim_hidden=32, num_heads=4, ln=(tag == "ln")).to(dev)
    m.load_state_dict({key[len(tag) + 3:]: torch.from_numpy(v) for key, v in g.items() if key.startswith(f"{tag}_w_")})
    out = m(torch.from_numpy(g[f"{tag}_X"]).to(dev)).cpu().numpy()
    assert out.shape == g[f"{tag}_Y"].shape
    assert rel_err(out, g[f"{tag}_Y"]) < ENC_REL_TOL


# ------------------------------------------------------------------------------------ experiment-1 chain (Code/pceval.py:61-99)
@pytest.mark.parametrize("fs,N", [(32000, 1434), (8000, 1843)])
def test_eval_sweep_chain_resample_window_stft_encoder(pca, dev, fs, N):
    """The per-file body of the window-size / sampling-rate sweep (Code/pceval.py:73-82,90-99): resample to fs (kaiser_fast,
    scale=True) -> STFT with win_length = N zero-padded to the next power of two, hop = N/2, divided by N -> log-magnitude ->
    2-D frame clouds with farr = linspace(0, fs/2, Nf)/fs -> shipped FST checkpoint.  GPU chain against the oracle chain
    (the resampling step is unpinned on both sides)."""
    fsog = 16000
    st, w = _load_ckpt(pca, dev, "fst", 2)
    audio = orc.synth_audio(2, 12000, fsog, seed=41)
    n_fft = int(2 ** np.ceil(np.log2(N)))
    x_gpu = pca.resample(torch.from_numpy(audio).to(dev), fsog, fs, res_type="kaiser_fast", scale=True)
    lm = pca.stft_logmag(x_gpu, n_fft, win_length=N, hop_factor=0.5)                  # (clips, Nt, Nf), divided by N
    farr = np.linspace(0, fs / 2, lm.shape[2]) / fs
    clouds = pca.build_clouds(lm.reshape(-1, 1, lm.shape[2]), farr)                   # one (Nf, 2) cloud per frame
    with torch.no_grad():
        got = st(clouds).cpu().numpy()
    ref = []
    for c in range(audio.shape[0]):
        xr = orc.resample_librosa080(audio[c], fsog, fs, res_type="kaiser_fast", scale=True)
        a = orc.logmag_recipe(xr, n_fft, 0.5, win_length=N)                           # (Nf, Nt)
        pcs = np.stack([orc.cloud_2d(a, farr, t) for t in range(a.shape[1])])
        ref.append(orc.st_forward({k: torch.from_numpy(np.asarray(v)) if not torch.is_tensor(v) else v for k, v in w.items()},
                                  torch.from_numpy(pcs), 8).numpy())
    ref = np.concatenate(ref)
    assert got.shape == ref.shape
    # fp32 STFT vs float64 FFT on near-empty bins moves the log-magnitudes slightly; logits agree to 5e-3 of their range
    assert rel_err(got, ref) < 5e-3
