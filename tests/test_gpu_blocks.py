"""GPU: building-block kernels through the C ABI against the oracle / torch autograd (fp64 on the host)."""
import numpy as np
import pytest
import torch

from oracle import mvf_oracle as O
from tests import helpers as H
from video_rep_learning_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("F,P,E,SPC,one_hot", [(5, 16, 3, 32, 1), (3, 196, 3, 384, 1), (2, 49, 6, 64, 0), (2, 30, 16, 96, 1),
                                                (4, 50, 4, 128, 1), (3, 9, 1, 256, 0), (2, 784, 2, 384, 1)])
@pytest.mark.parametrize("single_pass", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_xattn_pool_fwd_bwd(F, P, E, SPC, one_hot, dtype, single_pass):
    """single_pass=True hands the fp32 pooled-entity buffer to the kernels, which selects the one-pass bf16 kernels
    (E <= 4); otherwise (and for fp32 / E > 4) the generic kernels run."""
    g = torch.Generator().manual_seed(F * P + E)
    kv = torch.randn(F * P, 2 * SPC, generator=g).to(dtype)
    q_s = torch.randn(E, SPC, generator=g) * 0.3
    q_b = torch.randn(SPC, generator=g) * 0.1
    W = SPC + (E if one_hot else 0)
    ld = (W + 7) // 8 * 8
    # host reference in fp64 on the (possibly bf16-rounded) inputs
    kvd = kv.double().view(F, P, 2 * SPC).requires_grad_(True)
    qs, qb = q_s.double().requires_grad_(True), q_b.double().requires_grad_(True)
    K, V = kvd[..., :SPC], kvd[..., SPC:]
    A = torch.softmax(torch.einsum("fpc,ec->fep", K, qs + qb) / np.sqrt(SPC), -1)
    ent = torch.einsum("fep,fpc->fec", A, V)
    d_ent = torch.randn(F, E, SPC, generator=g).double()
    (ent * d_ent).sum().backward()

    dev = "cuda"
    kv_d, qs_d, qb_d = kv.to(dev), q_s.to(dev), q_b.to(dev)      # keep references: raw pointers cross the C ABI
    attn = torch.empty(F, E, P, device=dev)
    out = torch.full((F * E, ld), float("nan"), dtype=torch.float32, device=dev)   # pooled entities are always fp32
    ent32 = torch.full((F * E, SPC), float("nan"), device=dev) if single_pass else None
    md = L.MVF_BF16 if dtype == torch.bfloat16 else L.MVF_F32
    L.check(L.lib().mvf_xattn_pool_fwd(md, F, P, E, SPC, L.ptr(kv_d), L.ptr(qs_d), L.ptr(qb_d), L.ptr(attn),
                                       L.ptr(out), ld, L.ptr(ent32), one_hot, 0.0, 0, _stream()))
    torch.cuda.synchronize()
    tol = 1e-5
    assert float((attn.cpu().double() - A.detach()).abs().max()) < 1e-5
    got = out.float().cpu().view(F, E, ld)
    assert float((got[..., :SPC].double() - ent.detach()).abs().max()) < tol
    if one_hot:
        assert torch.equal(got[..., SPC:SPC + E], torch.eye(E).expand(F, E, E))
    assert float(got[..., W:].abs().max()) == 0.0 if ld > W else True
    if single_pass and dtype == torch.bfloat16 and E <= 4:
        assert float((ent32.cpu().double().view(F, E, SPC) - ent.detach()).abs().max()) < 1e-5

    d_in = torch.zeros(F * E, ld, dtype=torch.float32)
    d_in[:, :SPC] = d_ent.view(F * E, SPC).float()
    d_in_d = d_in.to(dev)
    d_kv = torch.empty_like(kv_d)
    dqs, dqb = torch.zeros(E, SPC, device=dev), torch.zeros(SPC, device=dev)
    dbk, dbv = torch.zeros(SPC, device=dev), torch.zeros(SPC, device=dev)
    L.check(L.lib().mvf_xattn_pool_bwd(md, F, P, E, SPC, L.ptr(kv_d), L.ptr(qs_d), L.ptr(qb_d), L.ptr(attn),
                                       L.ptr(d_in_d), ld, L.ptr(ent32), one_hot, 0.0, 0, L.ptr(d_kv), L.ptr(dqs), L.ptr(dqb),
                                       L.ptr(dbk), L.ptr(dbv), _stream()))
    torch.cuda.synchronize()
    rt = 2e-5 if dtype == torch.float32 else 2e-2
    assert H.rel_l2(d_kv.float().cpu().view(F, P, 2 * SPC), kvd.grad) < rt
    assert H.rel_l2(dqs.cpu(), qs.grad) < rt and H.rel_l2(dqb.cpu(), qb.grad) < rt
    assert H.rel_l2(dbv.cpu(), kvd.grad[..., SPC:].sum((0, 1))) < rt
    # key bias gradient is analytically zero (softmax shift invariance): only rounding noise may remain
    assert float(dbk.abs().max()) < 1e-3 * float(kvd.grad.abs().max()) * P * F


@pytest.mark.parametrize("B,S,heads,dk,masked", [(2, 24, 4, 8, True), (3, 60, 8, 32, True), (1, 100, 2, 16, False), (2, 70, 2, 64, True)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_temporal_attention_fwd_bwd(B, S, heads, dk, masked, dtype):
    Hd = heads * dk
    g = torch.Generator().manual_seed(B * S + heads)
    qkv = (torch.randn(B * S, 3 * Hd, generator=g) * 0.7).to(dtype)
    km = None
    if masked:
        km = (torch.rand(B, S, generator=g) > 0.25).float()
        km[:, 0] = 1
    x = qkv.double().view(B, S, 3, heads, dk).requires_grad_(True)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    sc = q @ k.transpose(-1, -2) / np.sqrt(dk)
    if km is not None:
        sc = sc.masked_fill(km[:, None, None, :] == 0, -float("inf"))
    ctx_ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * S, Hd)
    d_ctx = torch.randn(B * S, Hd, generator=g).to(dtype)
    (ctx_ref * d_ctx.double()).sum().backward()

    dev = "cuda"
    md = L.MVF_BF16 if dtype == torch.bfloat16 else L.MVF_F32
    qkv_d = qkv.to(dev)
    ctx = torch.empty(B * S, Hd, dtype=dtype, device=dev)
    lse = torch.empty(B, heads, S, device=dev)
    kmd = None if km is None else km.to(dev)
    L.check(L.lib().mvf_attention_fwd(md, B, S, heads, dk, L.ptr(qkv_d), L.ptr(kmd), L.ptr(ctx), L.ptr(lse), None, 0, _stream()))
    torch.cuda.synchronize()
    tol = 2e-6 if dtype == torch.float32 else 1.5e-2
    assert float((ctx.float().cpu().double() - ctx_ref.detach()).abs().max()) < tol * max(1.0, float(ctx_ref.abs().max()))
    d_qkv = torch.full_like(qkv_d, float("nan"))
    delta = torch.empty(B, heads, S, device=dev)
    d_ctx_d = d_ctx.to(dev)
    L.check(L.lib().mvf_attention_bwd(md, B, S, heads, dk, L.ptr(qkv_d), L.ptr(kmd), L.ptr(ctx), L.ptr(lse),
                                      L.ptr(d_ctx_d), L.ptr(d_qkv), L.ptr(delta), None, 0, _stream()))
    torch.cuda.synchronize()
    want = x.grad.reshape(B * S, 3 * Hd)
    assert H.rel_l2(d_qkv.float().cpu(), want) < (2e-5 if dtype == torch.float32 else 3e-2)


def _attention_reference(qkv, km, B, S, heads, dk, d_ctx):
    """fp64 statement of attention() (models/utils.py:11-44) + autograd, one (view, head) at a time to bound memory."""
    Hd = heads * dk
    x = qkv.double().view(B, S, 3, heads, dk).requires_grad_(True)
    ctx_ref = torch.empty(B, S, heads, dk, dtype=torch.float64)
    lse_ref = torch.empty(B, heads, S, dtype=torch.float64)
    g = d_ctx.double().view(B, S, heads, dk)
    grads = torch.zeros_like(x)
    for b in range(B):
        for h in range(heads):
            xb = x[b, :, :, h].detach().clone().requires_grad_(True)          # [S, 3, dk]
            sc = xb[:, 0] @ xb[:, 1].t() / np.sqrt(dk)
            if km is not None:
                sc = sc.masked_fill(km[b][None, :] == 0, -float("inf"))
            o = torch.softmax(sc, -1) @ xb[:, 2]
            (o * g[b, :, h]).sum().backward()
            ctx_ref[b, :, h] = o.detach()
            lse_ref[b, h] = torch.logsumexp(sc.detach(), -1)
            grads[b, :, :, h] = xb.grad
    return ctx_ref.reshape(B * S, Hd), lse_ref, grads.reshape(B * S, 3 * Hd)


def _run_attention_tc(qkv, km, B, S, heads, dk, d_ctx):
    dev = "cuda"
    Hd = heads * dk
    qkv_d = qkv.to(dev)
    ctx = torch.full((B * S, Hd), float("nan"), device=dev)
    lse = torch.full((B, heads, S), float("nan"), device=dev)
    kmd = None if km is None else km.to(dev)
    nb = L.lib().mvf_attention_ws_bytes(B, S, heads, dk)
    assert nb > 0
    ws = torch.empty(nb + 1024, dtype=torch.uint8, device=dev)
    wsp = (ws.data_ptr() + 1023) // 1024 * 1024
    L.check(L.lib().mvf_attention_fwd(L.MVF_F32, B, S, heads, dk, L.ptr(qkv_d), L.ptr(kmd), L.ptr(ctx), L.ptr(lse), wsp, nb, _stream()))
    torch.cuda.synchronize()
    d_qkv = torch.full_like(qkv_d, float("nan"))
    delta = torch.empty(B, heads, S, device=dev)
    L.check(L.lib().mvf_attention_bwd(L.MVF_F32, B, S, heads, dk, L.ptr(qkv_d), L.ptr(kmd), L.ptr(ctx), L.ptr(lse),
                                      L.ptr(d_ctx.to(dev)), L.ptr(d_qkv), L.ptr(delta), wsp, nb, _stream()))
    torch.cuda.synchronize()
    return ctx.cpu(), lse.cpu(), d_qkv.cpu()


def _attention_case(B, S, heads, masked, seed_extra=0):
    dk = 32
    Hd = heads * dk
    g = torch.Generator().manual_seed(B * S + heads + seed_extra)
    qkv = torch.randn(B * S, 3 * Hd, generator=g) * 0.7
    km = None
    if masked:
        km = (torch.rand(B, S, generator=g) > 0.25).float()
        km[:, 0] = 1
        km[0, 1:] = 0                      # view 0: a single valid key
        if B > 1 and S > 70:
            km[1, S // 2:] = 0             # view 1: the whole tail padded (ragged sequence), key tiles with no valid key
    d_ctx = torch.randn(B * S, Hd, generator=g)
    return dk, qkv, km, d_ctx


@pytest.mark.parametrize("B,S,heads,masked", [(3, 60, 8, True), (2, 64, 8, False), (2, 17, 2, True), (5, 33, 4, True), (1, 1, 1, False)])
def test_temporal_attention_tensor_core_path(B, S, heads, masked):
    """attention_tc.cu (S <= 64, d_k = 32; bf16 hi/lo operand splits on mma.sync) -- what the C ABI runs when a split-operand
    workspace is supplied -- against the fp64 formulation, including a view whose keys are all masked except one and a
    partially filled last warp."""
    dk, qkv, km, d_ctx = _attention_case(B, S, heads, masked)
    ctx_ref, lse_ref, want = _attention_reference(qkv, km, B, S, heads, dk, d_ctx)
    ctx, lse, d_qkv = _run_attention_tc(qkv, km, B, S, heads, dk, d_ctx)
    assert float((ctx.double() - ctx_ref).abs().max()) < 2e-5 * max(1.0, float(ctx_ref.abs().max()))
    assert float((lse.double() - lse_ref).abs().max()) < 2e-5
    assert H.rel_l2(d_qkv, want) < 5e-5


@pytest.mark.parametrize("B,S,heads,masked", [(2, 65, 2, True), (2, 128, 8, False), (3, 200, 4, True), (16, 480, 8, True),
                                              (2, 1000, 2, True), (2, 3840, 2, True), (1, 6000, 1, False)])
def test_temporal_attention_tcgen05_long_sequences(B, S, heads, masked):
    """attention_fa.cu: flash-attention forward, dQ and dK/dV on tcgen05 / TMEM / TMA (bf16 hi|lo operand splits) for S > 64:
    cfg4 (S = 480), cfg5 (S = 3840) and whole-video evaluation lengths, ragged key masks, tiles with no valid key, sequence
    lengths that are not multiples of the 64 / 128 tile sizes -- against the fp64 formulation."""
    dk, qkv, km, d_ctx = _attention_case(B, S, heads, masked, seed_extra=7)
    ctx_ref, lse_ref, want = _attention_reference(qkv, km, B, S, heads, dk, d_ctx)
    ctx, lse, d_qkv = _run_attention_tc(qkv, km, B, S, heads, dk, d_ctx)
    assert torch.isfinite(ctx).all() and torch.isfinite(d_qkv).all()
    assert float((ctx.double() - ctx_ref).abs().max()) < 2e-5 * max(1.0, float(ctx_ref.abs().max()))
    assert float((lse.double() - lse_ref).abs().max()) < 2e-5
    err = H.rel_l2(d_qkv, want)
    print(f"tcgen05 attention B {B} S {S} heads {heads}: ctx max err {float((ctx.double() - ctx_ref).abs().max()):.2e} d_qkv rel {err:.2e}")
    assert err < 5e-5


@pytest.mark.parametrize("B,S,heads", [(2, 17, 2), (3, 60, 8), (1, 64, 1)])
def test_temporal_attention_tcgen05_short_sequences_forced(monkeypatch, B, S, heads):
    """MVF_ATTN_FA=2 sends every length through the tcgen05 kernels: a single, partially filled key tile."""
    monkeypatch.setenv("MVF_ATTN_FA", "2")
    dk, qkv, km, d_ctx = _attention_case(B, S, heads, True, seed_extra=3)
    ctx_ref, lse_ref, want = _attention_reference(qkv, km, B, S, heads, dk, d_ctx)
    ctx, lse, d_qkv = _run_attention_tc(qkv, km, B, S, heads, dk, d_ctx)
    assert float((ctx.double() - ctx_ref).abs().max()) < 2e-5 * max(1.0, float(ctx_ref.abs().max()))
    assert float((lse.double() - lse_ref).abs().max()) < 2e-5
    assert H.rel_l2(d_qkv, want) < 5e-5


def test_dropout_mask_is_counter_based_and_unbiased():
    n_r, n_c, p = 1000, 257, 0.1
    a = torch.empty(n_r, n_c, device="cuda")
    b = torch.empty(n_r, n_c, device="cuda")
    L.check(L.lib().mvf_dropout_mask(1234, 3, n_r, n_c, p, L.ptr(a), _stream()))
    L.check(L.lib().mvf_dropout_mask(1234, 3, n_r, n_c, p, L.ptr(b), _stream()))
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    vals = torch.unique(a).cpu()
    assert torch.allclose(vals, torch.tensor([0.0, 1.0 / 0.9]))
    keep = float((a > 0).float().mean())
    assert abs(keep - 0.9) < 0.005
    L.check(L.lib().mvf_dropout_mask(1235, 3, n_r, n_c, p, L.ptr(b), _stream()))
    torch.cuda.synchronize()
    assert 0.7 < float(((a > 0) == (b > 0)).float().mean()) < 0.9       # a different seed decorrelates (~0.82)
    L.check(L.lib().mvf_dropout_mask(1234, 3, n_r, n_c, 0.0, L.ptr(b), _stream()))
    torch.cuda.synchronize()
    assert float(b.min()) == 1.0 and float(b.max()) == 1.0
