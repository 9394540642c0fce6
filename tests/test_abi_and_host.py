"""CPU tests: the C-ABI library loads and exports every symbol include/pcaudio_b200.h declares, the
ctypes prototypes cover them, and the host-side logic (param packing, sharding, tables) is right.
No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import pcaudio_b200 as pca
    return pca


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pcaudio_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pca_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from pcaudio_b200 import _lib
    names = header_symbols()
    assert len(names) >= 20
    handle = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototypes out of sync with the header"
    assert _lib.lib().pca_version() == 102


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "point-cloud-audio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"


def test_param_counts_match_reference_models(built):
    pca = built
    from pcaudio_b200 import _lib
    for d_in, expect in ((2, 80202), (3, 80394)):          # model_params of the shipped configs
        m = pca.ST(dim_input=d_in, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8)
        n = sum(p.numel() for p in m.parameters())
        assert n == expect
        dims = m._dims()
        assert _lib.lib().pca_st_param_count(C.byref(dims)) == n
        assert m._blob().numel() == n
    mn = pca.SetTransformer(dim_hidden=256, num_heads=4, num_inds=16)
    assert sum(p.numel() for p in mn.parameters()) == 1140264


def test_state_dict_keys_and_module_prefix(built, golden_dir):
    pca = built
    w = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, "3st_weights.npz")).items()}
    assert all(k.startswith("module.") for k in w)
    m = pca.ST(dim_input=3, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8)
    m.load_state_dict(w)                                            # DataParallel prefix accepted
    assert set(m.state_dict()) == {k[7:] for k in w}
    blob = m._blob()
    # canonical order starts with enc.0.I then mab0.fc_q.weight
    np.testing.assert_array_equal(blob[:64 * 64].numpy(), w["module.enc.0.I"].reshape(-1).numpy())
    np.testing.assert_array_equal(blob[64 * 64:2 * 64 * 64].numpy(), w["module.enc.0.mab0.fc_q.weight"].reshape(-1).numpy())
    # cache invalidation on in-place update
    with torch.no_grad():
        m.enc[0].I.add_(1.0)
    assert torch.equal(m._blob()[:4096], (w["module.enc.0.I"].reshape(-1) + 1.0))


def test_cpu_tensors_fail_loudly(built):
    pca = built
    m = pca.ST(dim_input=2, num_outputs=1, dim_output=10, num_inds=4, dim_hidden=8, num_heads=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 5, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pca.stft_logmag(torch.zeros(1, 4000), 256)


def test_missing_library_fails_loudly(built, monkeypatch):
    from pcaudio_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpcaudio_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_coord_tables_match_oracle(built):
    from oracle import pcaudio_oracle as orc
    f1, t1 = built.coord_tables(16000, 512, 1024, 0.5, 10)
    f2, t2 = orc.coord_tables(16000, 512, 1024, 0.5, 10)
    np.testing.assert_array_equal(f1, f2)
    np.testing.assert_array_equal(t1, t2)


def test_pipeline_shape_queries(built):
    from pcaudio_b200 import _lib
    L = _lib.lib()
    dims = _lib.StDims(d_in=2, D=64, H=8, M=64, S=1, C=10, ln=0)
    cfg = _lib.PipelineCfg(n_samples=16000, n_fft=2048, hop=1024, scale=1 / 2048, mode=2, ntemp=0, top_k=0,
                           precision=0, st=dims)
    assert L.pca_pipeline_clouds_per_clip(C.byref(cfg)) == 16
    assert L.pca_pipeline_points_per_cloud(C.byref(cfg)) == 1025
    dims3 = _lib.StDims(d_in=3, D=64, H=8, M=64, S=1, C=10, ln=0)
    cfg3 = _lib.PipelineCfg(n_samples=64000, n_fft=1024, hop=512, scale=1 / 1024, mode=3, ntemp=10, top_k=0,
                            precision=0, st=dims3)
    assert L.pca_pipeline_clouds_per_clip(C.byref(cfg3)) == 12       # 126 frames -> 12 chunks of 10
    assert L.pca_pipeline_points_per_cloud(C.byref(cfg3)) == 5120
    cfg3.top_k = 256
    assert L.pca_pipeline_points_per_cloud(C.byref(cfg3)) == 256
    cfg3.st.d_in = 2                                                  # width mismatch is an error, not UB
    assert L.pca_pipeline_clouds_per_clip(C.byref(cfg3)) == -1
    assert b"wide" in L.pca_last_error()


def test_shard_range_partitions():
    from pcaudio_b200.parallel import shard_range
    for n in (0, 1, 7, 256, 1000):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pcaudio_b200.parallel import allreduce_mean_, gather_rows, init_distributed, reduce_flat_gradient_, shard_range
    r, w, _ = init_distributed("gloo")
    full = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
    lo, hi = shard_range(7, r, w)
    got = gather_rows(full[lo:hi].clone(), 7, r, w)
    g = torch.full((5,), float(r + 1))
    allreduce_mean_(g, w)
    # the training step's exchange: one summing allreduce of the flat gradient, 1/world folded into the optimizer
    flat = torch.arange(6, dtype=torch.float32) * (r + 1)
    scale = reduce_flat_gradient_(flat)
    q.put((r, torch.equal(got, full), g.tolist(), (flat * scale).tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_and_grad_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for r, ok, g, mean_grad in res:
        assert ok, f"rank {r}: gathered rows differ from the unsharded tensor"
        assert g == [1.5] * 5
        assert mean_grad == [1.5 * i for i in range(6)]          # mean over ranks of i * (rank + 1)


def test_torch_custom_ops_registered_with_fake_impls(built):
    """The host mirrors reach the C ABI through torch.ops.pcaudio.* (dispatcher-visible custom ops); shape inference
    must work without running a kernel (fake tensors), and there is no CPU kernel to fall back to."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    import pcaudio_b200  # noqa: F401
    for name in ("st_fwd", "stft_logmag", "select_points"):
        assert hasattr(torch.ops.pcaudio, name)
    with FakeTensorMode():
        X = torch.empty(5, 300, 2, device="cuda")
        blob = torch.empty(100, device="cuda")
        out = torch.ops.pcaudio.st_fwd(X, None, blob, 2, 64, 8, 64, 1, 10, 0, 0)
        assert out.shape == (5, 1, 10) and out.device.type == "cuda"
        lm = torch.ops.pcaudio.stft_logmag(torch.empty(3, 16000, device="cuda"), torch.empty(1024, device="cuda"),
                                           torch.empty(512, 2, device="cuda"), 1024, 512, 1.0 / 1024, True, 30)
        assert lm.shape == (3, 30, 512)
        pts, idx, cnt = torch.ops.pcaudio.select_points(lm.view(9, 10, 512), torch.empty(512, device="cuda"),
                                                        torch.empty(10, device="cuda"), 256, True, True, -6.0)
        assert pts.shape == (9, 256, 3) and idx.shape == (9, 256) and cnt.shape == (9,)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.pcaudio.st_fwd(torch.zeros(1, 4, 2), None, torch.zeros(10), 2, 64, 8, 64, 1, 10, 0, 0)   # CPU: no kernel


def test_flatten_parameters_packs_views_in_blob_order(built):
    """SetTrainer's memory model: every parameter is a view of ONE flat fp32 buffer in the packed-weight order of the C
    ABI, so the weights blob, the flat gradient and the fused optimizer address the same memory (host logic, no GPU)."""
    import pcaudio_b200 as pca
    torch.manual_seed(0)
    m = pca.SetTransformer(dim_input=3, num_outputs=1, dim_output=5, num_inds=4, dim_hidden=8, num_heads=2)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    packed = m._blob().clone()
    flat = m.flatten_parameters()
    assert flat.numel() == sum(p.numel() for p in m.parameters()) == packed.numel()
    assert torch.equal(flat, packed) and m._blob() is flat
    assert list(m.state_dict().keys()) == list(before.keys())
    assert all(torch.equal(v, before[k]) for k, v in m.state_dict().items())
    off = 0
    for t in m._param_tensors():                           # views, in order, no gaps
        assert t.data_ptr() == flat.data_ptr() + 4 * off
        off += t.numel()
    flat.mul_(2.0)                                         # an in-place optimizer step on the flat buffer ...
    assert all(torch.equal(v, 2.0 * before[k]) for k, v in m.state_dict().items())      # ... is seen by every parameter
    m.load_state_dict(before)                              # copy_ keeps the views
    assert m._blob() is flat and torch.equal(flat, packed)
    m.dec[3].weight.data = m.dec[3].weight.data.clone()    # a parameter re-pointed elsewhere: fall back to re-packing
    assert m._blob() is not flat and torch.equal(m._blob(), packed)


def test_packed_params_cache_is_per_device_and_invalidatable():
    """nn.DataParallel replicas share the cache object: it must key by device, return local values and offer an
    explicit invalidate for edits through .data (which do not bump the version counter)."""
    import torch
    from pcaudio_b200.modules import MAB, _PackedParams, _mab_tensors, invalidate_packed
    m = MAB(4, 4, 8, 2)
    pk = m._packed
    assert isinstance(pk, _PackedParams)
    b0 = pk.get(_mab_tensors(m))
    assert pk.get(_mab_tensors(m)) is b0                      # cached
    with torch.no_grad():
        m.fc_q.weight.mul_(2.0)                              # version bump -> rebuilt
    b1 = pk.get(_mab_tensors(m))
    assert b1 is not b0 and torch.equal(b1[:32], m.fc_q.weight.reshape(-1))
    m.fc_q.weight.data.mul_(0.5)                             # .data edit: invisible to the key ...
    assert pk.get(_mab_tensors(m)) is b1
    invalidate_packed(m)                                     # ... until invalidated
    b2 = pk.get(_mab_tensors(m))
    assert b2 is not b1 and torch.equal(b2[:32], m.fc_q.weight.reshape(-1))
    meta = [t.to("meta") for t in _mab_tensors(m)]           # a second "device": separate slot, the first one survives
    assert set(pk._cache) == {torch.device("cpu")}
    del meta


def test_attention_route_dispatch_is_by_shape(built):
    """Host logic of the tensor-core attention route (csrc/attn_tc.cu): which MAB shapes take it, and that its scratch query
    grows with the batch (no device needed: these entry points only inspect shapes)."""
    from pcaudio_b200 import _lib
    L = _lib.lib()
    # ModelNet model (main_pointcloud.py:62): 16 inducing points / 1 seed against 1000 points, dim 256, 4 heads
    assert L.pca_debug_attn_tc_eligible(256, 1000, 16, 256, 4) == 1      # ISAB mab1: points are the queries
    assert L.pca_debug_attn_tc_eligible(256, 16, 1000, 256, 4) == 1      # ISAB mab0: points are the keys
    assert L.pca_debug_attn_tc_eligible(256, 1, 1000, 256, 4) == 1       # PMA
    # audio models (Code/models.py: dim 64, 8 heads, 64 inducing points): the score matrix would be 8x the activations
    assert L.pca_debug_attn_tc_eligible(4096, 64, 1025, 64, 8) == 0
    assert L.pca_debug_attn_tc_eligible(4096, 1025, 64, 64, 8) == 0
    assert L.pca_debug_attn_tc_eligible(4096, 1, 1025, 64, 8) == 0
    assert L.pca_debug_attn_tc_eligible(2, 100, 16, 256, 4) == 0         # fewer than 128 items on the large side
    small = L.pca_debug_attn_ws_bytes(8, 1000, 16, 256, 4)
    large = L.pca_debug_attn_ws_bytes(64, 1000, 16, 256, 4)
    assert 0 < small < large
    try:
        L.pca_debug_set_attn_tc(0)
        assert L.pca_debug_attn_tc_eligible(256, 1000, 16, 256, 4) == 0  # the switch keeps every shape on the CUDA-core kernels
    finally:
        L.pca_debug_set_attn_tc(1)
