"""Tensor-core attention for MABs with one small side (csrc/attn_tc.cu) against float64 torch: forward (O, log-sum-exp) and the
hand-written backward (dQp, dKV) -- the contraction of set_transformer-master/modules.py:20-29 on projected operands."""
import math
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu
ATTN_TC_REL_TOL = 1e-4       # max |y - y_ref| / max |y_ref| per tensor (split-bf16 products: measured ~1e-5; fp32 parity class 1e-3)


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as g
    g.build()
    from pcaudio_b200 import _lib
    return _lib


def _ref(Qp, KV, D, H):
    """float64 reference on the projected operands: O = Qp + softmax_h(Qp K^T / sqrt(D)) V, lse in the log2 domain."""
    B, nk = KV.shape[0], KV.shape[1]
    Q = Qp.expand(B, -1, -1) if Qp.shape[0] == 1 else Qp
    K, V = KV[..., :D], KV[..., D:]
    dh = D // H
    q = Q.reshape(B, -1, H, dh).transpose(1, 2)
    k = K.reshape(B, nk, H, dh).transpose(1, 2)
    v = V.reshape(B, nk, H, dh).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / math.sqrt(D)
    A = torch.softmax(s, dim=-1)
    O = Q + (A @ v).transpose(1, 2).reshape(B, -1, D)
    lse = torch.logsumexp(s, dim=-1) / math.log(2.0)          # (B, H, nq)
    return O, lse.transpose(1, 2)


SHAPES = [
    # B, nq, nk, D, H, q_shared
    (6, 1000, 16, 256, 4, 0),      # ISAB mab1 of the ModelNet model: points are the queries, 16 inducing keys
    (6, 16, 1000, 256, 4, 1),      # ISAB mab0: 16 shared inducing queries, points are the keys
    (6, 1, 1000, 256, 4, 1),       # PMA: one shared seed
    (5, 333, 11, 128, 4, 0),       # ragged tiles, fewer keys than the padded column group
    (5, 7, 411, 128, 4, 0),        # per-cloud small queries
    (3, 640, 8, 256, 8, 0),        # 8 heads x 8 keys
    (4, 16, 300, 64, 2, 1),        # 2 heads x 16 queries: 32 columns
]


@pytest.mark.parametrize("B,nq,nk,D,H,q_shared", SHAPES)
def test_attn_tc_forward(L, B, nq, nk, D, H, q_shared):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + nq + nk)
    Qp = torch.randn(1 if q_shared else B, nq, D, generator=g).to(dev) * 2.0
    KV = torch.randn(B, nk, 2 * D, generator=g).to(dev)
    assert L.lib().pca_debug_attn_tc_eligible(B, nq, nk, D, H) == 1
    ws = torch.empty(L.lib().pca_debug_attn_ws_bytes(B, nq, nk, D, H), dtype=torch.uint8, device=dev)
    out = {}
    try:
        for on in (1, 0):
            L.lib().pca_debug_set_attn_tc(on)
            O = torch.full((B, nq, D), float("nan"), device=dev)
            lse = torch.full((B, nq, H), float("nan"), device=dev)
            L.check(L.lib().pca_debug_attn_fwd(L.ptr(Qp), q_shared, L.ptr(KV), B, nq, nk, D, H, L.ptr(O), L.ptr(lse), L.ptr(ws), ws.numel(),
                                               None), "attn_fwd")
            torch.cuda.synchronize()
            out[on] = (O, lse)
    finally:
        L.lib().pca_debug_set_attn_tc(1)
    O_ref, lse_ref = _ref(Qp.double(), KV.double(), D, H)
    for on in (1, 0):
        O, lse = out[on]
        err = ((O.double() - O_ref).abs().max() / O_ref.abs().max()).item()
        err_l = (lse.double() - lse_ref).abs().max().item()
        assert err < ATTN_TC_REL_TOL, f"attn_tc={on}: O rel err {err:.3e}"
        assert err_l < 1e-3, f"attn_tc={on}: lse abs err {err_l:.3e}"
    assert not torch.equal(out[0][0], out[1][0])        # the switch really selects a different kernel


@pytest.mark.parametrize("B,nq,nk,D,H,q_shared", SHAPES)
def test_attn_tc_backward(L, B, nq, nk, D, H, q_shared):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(7 + B * 1000 + nq + nk)
    Qp = (torch.randn(1 if q_shared else B, nq, D, generator=g) * 2.0).to(dev)
    KV = torch.randn(B, nk, 2 * D, generator=g).to(dev)
    dO = torch.randn(B, nq, D, generator=g).to(dev)
    # float64 autograd; a shared query set gets the per-cloud gradients here (the caller sums them)
    Qd = (Qp.double().expand(B, -1, -1) if q_shared else Qp.double()).clone().requires_grad_(True)
    KVd = KV.double().clone().requires_grad_(True)
    O_ref, lse_ref = _ref(Qd, KVd, D, H)
    (O_ref * dO.double()).sum().backward()
    delta = ((O_ref.detach() - Qd.detach()) * dO.double()).reshape(B, nq, H, D // H).sum(-1).float().contiguous()
    lse = lse_ref.detach().float().contiguous()
    ws = torch.empty(L.lib().pca_debug_attn_ws_bytes(B, nq, nk, D, H), dtype=torch.uint8, device=dev)
    dQp = torch.full((B, nq, D), float("nan"), device=dev)
    dKV = torch.full((B, nk, 2 * D), float("nan"), device=dev)
    L.check(L.lib().pca_debug_attn_bwd_tc(L.ptr(Qp), q_shared, L.ptr(KV), L.ptr(dO), L.ptr(lse), L.ptr(delta), B, nq, nk, D, H, L.ptr(dQp),
                                          L.ptr(dKV), L.ptr(ws), ws.numel(), None), "attn_bwd_tc")
    torch.cuda.synchronize()
    for name, got, ref in (("dQp", dQp, Qd.grad), ("dK", dKV[..., :D], KVd.grad[..., :D]), ("dV", dKV[..., D:], KVd.grad[..., D:])):
        err = ((got.double() - ref).abs().max() / ref.abs().max()).item()
        assert err < ATTN_TC_REL_TOL, f"{name}: rel err {err:.3e}"


def test_attn_tc_not_eligible_shapes(L):
    lib = L.lib()
    assert lib.pca_debug_attn_tc_eligible(8, 64, 1025, 64, 8) == 0        # audio dims: 8 heads x 64 inducing points = 512 columns
    assert lib.pca_debug_attn_tc_eligible(8, 100, 16, 256, 4) == 0        # large side below one tile
    assert lib.pca_debug_attn_tc_eligible(8, 16, 16, 256, 4) == 0         # both sides small (SAB decoder)
    assert lib.pca_debug_attn_tc_eligible(8, 1000, 16, 100, 4) == 0       # dim_V not a multiple of 32
    assert lib.pca_debug_attn_tc_eligible(64, 1, 1025, 64, 8) == 0        # audio PMA: 8 heads x 8 padded columns are as wide as D = 64


def test_modelnet_model_attention_paths_agree(L):
    """The fp32 ModelNet model with its attention on the tensor cores vs on the CUDA-core kernels, inference and gradients."""
    import pcaudio_b200 as pca
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    model = pca.SetTransformer(dim_hidden=256, num_heads=4, num_inds=16).to(dev)
    X = torch.randn(8, 1000, 3, device=dev)
    y = torch.randint(0, 40, (8,), device=dev)
    res = {}
    try:
        for on in (1, 2, 0):                             # 1: default (training folds the shared-query blocks), 2: projected K | V form
            L.lib().pca_debug_set_attn_tc(on)
            model.eval()
            with torch.no_grad():
                logits = model(X).clone()
            model.train()
            for m in model.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            model.zero_grad()
            with torch.enable_grad():                    # conftest runs the inference tests under no_grad
                loss = torch.nn.functional.cross_entropy(model(X).squeeze(), y)
                loss.backward()
            res[on] = (logits, [p.grad.clone() for p in model.parameters()])
    finally:
        L.lib().pca_debug_set_attn_tc(1)
    b = res[0]
    gmax = max(gb.abs().max().item() for gb in b[1])
    for mode in (1, 2):
        a = res[mode]
        assert not torch.equal(a[0], b[0])
        assert ((a[0] - b[0]).abs().max() / b[0].abs().max()).item() < 1e-4
        for ga, gb in zip(a[1], b[1]):
            # relative to the tensor's own scale, with a floor: the key-bias gradients are zero up to rounding (softmax is shift invariant)
            assert ((ga - gb).abs().max() / gb.abs().max().clamp_min(1e-4 * gmax)).item() < 1e-3, mode
    assert any(not torch.equal(x, y) for x, y in zip(res[1][1], res[2][1]))      # the folded backward really is a different route


def test_modelnet_model_variable_size_sets(L):
    """Variable-size sets through the tensor-core attention (masked column softmax, NaN padding never read as a key): the
    result equals the model applied to X[b:b+1, :counts[b]], and an all-valid mask reproduces the unmasked path bit for bit."""
    import pcaudio_b200 as pca
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    model = pca.SetTransformer(dim_hidden=256, num_heads=4, num_inds=16).to(dev).eval()
    counts = [1000, 999, 513, 128, 700, 1]
    X = torch.randn(len(counts), 1000, 3, device=dev)
    Xpad = X.clone()
    for b, c in enumerate(counts):
        Xpad[b, c:] = float("nan")
    with torch.no_grad():
        out = model(Xpad, counts=torch.tensor(counts, dtype=torch.int32, device=dev))
        assert torch.isfinite(out).all()
        try:
            L.lib().pca_debug_set_attn_tc(0)             # per-sample reference on the CUDA-core kernels (short sets are not eligible anyway)
            ref = torch.stack([model(X[b:b + 1, :c]).reshape(-1) for b, c in enumerate(counts)])
        finally:
            L.lib().pca_debug_set_attn_tc(1)
        assert ((out.reshape(ref.shape) - ref).abs().max() / ref.abs().max()).item() < 1e-4
        full = torch.full((len(counts),), 1000, dtype=torch.int32, device=dev)
        assert torch.equal(model(X, counts=full), model(X))
