"""GPU: the folded entity pooling (pool_fold.cu) through the C ABI against the reference formulation evaluated in
fp64 on the same inputs: K = X Wk^T + bk, V = X Wv^T + bv, A = softmax(Q K^T / sqrt(SPC)), ent = A V
(mvformer.py:352-414, utils.py:11-44) -- forward pieces and, through autograd on that formulation, backward pieces."""
import math

import pytest
import torch

from video_rep_learning_b200 import _lib as L

pytestmark = pytest.mark.gpu

# (F, P, E, C_in, SPC, dtype): bench shape slice, ragged token groups (P % 8 != 0), E > 4 (two entity passes), E = 16,
# more frames than resident CTAs (several frames per CTA), tiny channels (partially filled warps)
SHAPES = [(6, 196, 3, 2304, 384, torch.bfloat16), (3, 196, 3, 1152, 384, torch.bfloat16), (5, 30, 8, 768, 64, torch.bfloat16),
          (4, 16, 11, 256, 32, torch.bfloat16), (300, 40, 3, 2304, 384, torch.bfloat16), (5, 196, 3, 1152, 384, torch.float32), (7, 50, 6, 384, 64, torch.bfloat16),
          (3, 17, 16, 64, 32, torch.float32), (400, 33, 3, 2304, 384, torch.bfloat16), (9, 9, 1, 48, 32, torch.float32),
          (4, 8, 2, 200, 40, torch.bfloat16), (310, 196, 4, 768, 384, torch.float32)]


def _st():
    return torch.cuda.current_stream().cuda_stream


def _mk(F, P, E, C, SPC, dtype, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed + F * 3 + P + C)
    X = torch.randn(F, P, C, generator=g, device="cuda").to(dtype)
    qs = torch.randn(E, SPC, generator=g, device="cuda") * 0.7
    qb = torch.randn(SPC, generator=g, device="cuda") * 0.1
    Wk = (torch.rand(SPC, C, generator=g, device="cuda") * 2 - 1) / math.sqrt(C) * 3.0   # sharp-ish softmax
    Wv = (torch.rand(SPC, C, generator=g, device="cuda") * 2 - 1) / math.sqrt(C)
    bk = torch.randn(SPC, generator=g, device="cuda") * 0.1
    bv = torch.randn(SPC, generator=g, device="cuda") * 0.1
    return X, qs, qb, Wk, Wv, bk, bv


def _reference(X, qs, qb, Wk, Wv, bk, bv):
    """As written in the reference, fp64."""
    Xd = X.double()
    K = Xd @ Wk.double().t() + bk.double()
    V = Xd @ Wv.double().t() + bv.double()
    Q = qs.double() + qb.double()
    A = torch.softmax(torch.einsum("fpc,ec->fep", K, Q) / math.sqrt(qs.shape[1]), -1)
    return A, torch.einsum("fep,fpc->fec", A, V)


def _fold_forward(X, qs, qb, Wk):
    lib = L.lib()
    F, P, C = X.shape
    E, SPC = qs.shape
    dt = L.MVF_BF16 if X.dtype == torch.bfloat16 else L.MVF_F32
    wq = torch.full((E, C), float("nan"), device="cuda")
    attn = torch.full((F, E, P), float("nan"), device="cuda")
    px = torch.full((F * E, C), float("nan"), device="cuda")
    L.check(lib.mvf_pool_fold_prep(L.ptr(qs), L.ptr(qb), L.ptr(Wk), E, SPC, C, L.ptr(wq), _st()))
    L.check(lib.mvf_pool_fold_fwd(dt, F, P, E, C, L.ptr(X), L.ptr(wq), L.ptr(attn), L.ptr(px), _st()))
    torch.cuda.synchronize()
    return wq, attn, px


@pytest.mark.parametrize("F,P,E,C,SPC,dtype", SHAPES)
def test_fold_forward_equals_dense_formulation(F, P, E, C, SPC, dtype):
    X, qs, qb, Wk, Wv, bk, bv = _mk(F, P, E, C, SPC, dtype)
    A_ref, ent_ref = _reference(X, qs, qb, Wk, Wv, bk, bv)
    wq, attn, px = _fold_forward(X, qs, qb, Wk)
    wq_ref = ((qs + qb).double() @ Wk.double()) / math.sqrt(SPC)
    assert float((wq.double() - wq_ref).abs().max() / wq_ref.abs().max()) < 2e-6
    assert float((attn.double() - A_ref).abs().max()) < 3e-6
    assert float((attn.sum(-1) - 1).abs().max()) < 1e-5
    ent = px.double().view(F, E, C) @ Wv.double().t() + bv.double()      # the value projection after the pooling
    assert float((ent - ent_ref).abs().max() / ent_ref.abs().max()) < 5e-6


@pytest.mark.parametrize("F,P,E,C,SPC,dtype", SHAPES)
def test_fold_backward_equals_autograd_of_dense_formulation(F, P, E, C, SPC, dtype):
    X, qs, qb, Wk, Wv, bk, bv = _mk(F, P, E, C, SPC, dtype, seed=5)
    lib = L.lib()
    dt = L.MVF_BF16 if dtype == torch.bfloat16 else L.MVF_F32
    # reference gradients by autograd on the dense formulation (fp64)
    P64 = [t.double().clone().requires_grad_(True) for t in (qs, qb, Wk, Wv, bk, bv)]
    A_ref, ent_ref = _reference(X, *P64)
    g = torch.Generator(device="cuda").manual_seed(11)
    dEnt = torch.randn(F, E, SPC, generator=g, device="cuda")
    (ent_ref * dEnt.double()).sum().backward()
    d_qs, d_qb, d_Wk, d_Wv, d_bk, d_bv = [p.grad for p in P64]
    # folded backward: G = dEnt Wv, streaming pass, finish
    wq, attn, px = _fold_forward(X, qs, qb, Wk)
    G = (dEnt.view(F * E, SPC).double() @ Wv.double()).float().contiguous()
    dwq = torch.zeros(E, C, device="cuda")
    L.check(lib.mvf_pool_fold_bwd(dt, F, P, E, C, L.ptr(X), L.ptr(G), L.ptr(px), L.ptr(attn), L.ptr(dwq), _st()))
    ld = C + 8
    dWk = torch.zeros(SPC, ld, device="cuda")
    dqs = torch.zeros(E, SPC, device="cuda")
    dqb = torch.zeros(SPC, device="cuda")
    L.check(lib.mvf_pool_fold_finish(L.ptr(dwq), L.ptr(qs), L.ptr(qb), L.ptr(Wk), E, SPC, C, L.ptr(dWk), ld, L.ptr(dqs),
                                     L.ptr(dqb), _st()))
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.double() - b).norm() / (b.norm() + 1e-300))
    tol = 2e-5
    assert rel(dWk[:, :C], d_Wk) < tol
    assert float(dWk[:, C:].abs().max()) == 0.0
    assert rel(dqs, d_qs) < tol
    assert rel(dqb, d_qb) < tol
    # the pieces the small GEMMs produce: dWv = dEnt^T px, dbv = colsum(dEnt); d(bk) is analytically zero
    assert rel(dEnt.view(F * E, SPC).double().t() @ px.double(), d_Wv) < tol
    assert rel(dEnt.view(F * E, SPC).double().sum(0), d_bv) < 1e-12
    assert float(d_bk.abs().max()) < 1e-9 * float(d_bv.abs().max())


@pytest.mark.parametrize("F,P,E,C", [(9, 196, 3, 2304), (301, 196, 3, 2304), (7, 50, 6, 384), (5, 30, 11, 768), (6, 9, 3, 48)])
def test_fold_tensor_core_and_cuda_core_kernels_agree(monkeypatch, F, P, E, C):
    """pool_fold_ws.cu (bf16 tokens, C_in % 16 == 0: warp-specialised mma.sync kernels, the default) against the CUDA-core
    kernels of pool_fold.cu (MVF_FOLD_WS=0): forward outputs, and the backward pass with the caller-supplied delta, with the
    in-kernel delta, and on the CUDA-core kernel."""
    SPC = 64
    X, qs, qb, Wk, Wv, bk, bv = _mk(F, P, E, C, SPC, torch.bfloat16, seed=8)
    lib = L.lib()
    _, attn_w, px_w = _fold_forward(X, qs, qb, Wk)
    g = torch.Generator(device="cuda").manual_seed(2)
    G = torch.randn(F * E, C, generator=g, device="cuda") * 0.05
    delta = (G.double() * px_w.double()).sum(-1).float().contiguous()

    def bwd(with_delta):
        dwq = torch.zeros(E, C, device="cuda")
        if with_delta:
            L.check(lib.mvf_pool_fold_bwd_delta(L.MVF_BF16, F, P, E, C, L.ptr(X), L.ptr(G), L.ptr(px_w), L.ptr(attn_w),
                                                L.ptr(delta), L.ptr(dwq), _st()))
        else:
            L.check(lib.mvf_pool_fold_bwd(L.MVF_BF16, F, P, E, C, L.ptr(X), L.ptr(G), L.ptr(px_w), L.ptr(attn_w), L.ptr(dwq), _st()))
        torch.cuda.synchronize()
        return dwq

    d_given, d_inner = bwd(True), bwd(False)
    monkeypatch.setenv("MVF_FOLD_WS", "0")
    _, attn_m, px_m = _fold_forward(X, qs, qb, Wk)
    d_first = bwd(False)
    assert float((attn_w - attn_m).abs().max()) < 2e-6
    assert float((px_w - px_m).abs().max() / px_m.abs().max()) < 1e-5      # bf16 hi/lo tensor-core products vs fp32 FMA
    # fp64 reference of the streaming pass on the same operands
    dA = torch.einsum("fec,fpc->fep", G.view(F, E, C).double(), X.double())
    dS = attn_w.double() * (dA - delta.double().view(F, E, 1))
    ref = torch.einsum("fep,fpc->ec", dS, X.double())
    for got in (d_given, d_inner, d_first):
        assert float((got.double() - ref).norm() / ref.norm()) < 2e-5


def test_fold_rejects_unsupported_shapes():
    lib = L.lib()
    x = torch.zeros(2, 4, 44, device="cuda")
    o = torch.zeros(4, 44, device="cuda")
    st = lib.mvf_pool_fold_fwd(L.MVF_F32, 2, 4, 2, 44, L.ptr(x), L.ptr(o), L.ptr(o), L.ptr(o), _st())
    assert st != 0 and "multiple of 8" in L.last_error()
    st = lib.mvf_pool_fold_fwd(L.MVF_F32, 2, 4, 2, 48, None, L.ptr(o), L.ptr(o), L.ptr(o), _st())
    assert st == 1 and "null pointer" in L.last_error()
