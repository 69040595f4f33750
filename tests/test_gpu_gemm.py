"""GPU: the two GEMM engines through the C ABI.  SIMT fp32 vs torch fp64; tcgen05 (bf16) vs the same product of
the bf16-rounded operands in fp64 -- every operand-major combination, ragged sizes, epilogue flags, split-K."""
import pytest
import torch

from video_rep_learning_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _gemm(backend, A, B, a_k, b_k, M, N, K, c_dtype, bias=None, relu_src=None, flags=0, split_k=1, C=None):
    dev = A.device
    ab = L.MVF_BF16 if A.dtype == torch.bfloat16 else L.MVF_F32
    if C is None:
        C = torch.full((M, N), float("nan"), dtype=c_dtype, device=dev)
    cd = L.MVF_BF16 if c_dtype == torch.bfloat16 else L.MVF_F32
    st = L.lib().mvf_gemm(backend, ab, cd, int(a_k), int(b_k), M, N, K, L.ptr(A), A.stride(0), L.ptr(B), B.stride(0),
                          L.ptr(C), C.stride(0), L.ptr(bias), L.ptr(relu_src), relu_src.stride(0) if relu_src is not None else 0,
                          flags, split_k, torch.cuda.current_stream().cuda_stream)
    L.check(st, "mvf_gemm")
    torch.cuda.synchronize()
    return C


def _operands(M, N, K, a_k, b_k, dtype, pad=0):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((M, K + pad) if a_k else (K, M + pad), generator=g, device="cuda").to(dtype)
    B = torch.randn((N, K + pad) if b_k else (K, N + pad), generator=g, device="cuda").to(dtype)
    Av = A[:, :K] if a_k else A[:, :M]
    Bv = B[:, :K] if b_k else B[:, :N]
    Am = Av.double() if a_k else Av.double().t()
    Bm = Bv.double().t() if b_k else Bv.double()
    return A, B, Am @ Bm


@pytest.mark.parametrize("a_k,b_k", [(1, 1), (1, 0), (0, 0), (0, 1)])
@pytest.mark.parametrize("M,N,K", [(64, 64, 16), (130, 70, 37), (257, 129, 300)])
def test_simt_fp32(a_k, b_k, M, N, K):
    A, B, ref = _operands(M, N, K, a_k, b_k, torch.float32)
    bias = torch.randn(N, device="cuda")
    C = _gemm(L.GEMM_SIMT, A, B, a_k, b_k, M, N, K, torch.float32, bias=bias)
    err = (C.double() - (ref + bias.double())).abs().max() / ref.abs().max()
    assert float(err) < 2e-6
    C2 = _gemm(L.GEMM_SIMT, A, B, a_k, b_k, M, N, K, torch.float32, flags=L.GEMM_RELU)
    assert float((C2.double() - ref.clamp_min(0)).abs().max() / ref.abs().max()) < 2e-6
    base = torch.randn(M, N, device="cuda")
    C3 = _gemm(L.GEMM_SIMT, A, B, a_k, b_k, M, N, K, torch.float32, flags=L.GEMM_ACCUM, C=base.clone())
    assert float((C3.double() - (ref + base.double())).abs().max() / ref.abs().max()) < 2e-6
    mask_src = torch.randn(M, N, device="cuda")
    C4 = _gemm(L.GEMM_SIMT, A, B, a_k, b_k, M, N, K, torch.float32, relu_src=mask_src, flags=L.GEMM_RELUMASK)
    assert float((C4.double() - ref * (mask_src > 0)).abs().max() / ref.abs().max()) < 2e-6


TC_SHAPES = [(128, 64, 64), (128, 256, 128), (256, 128, 192), (200, 72, 100), (384, 768, 320), (1000, 520, 72),
             (3840, 1024, 256), (130, 392, 512)]


@pytest.mark.parametrize("a_k,b_k", [(1, 1), (1, 0), (0, 0), (0, 1)])
@pytest.mark.parametrize("M,N,K", TC_SHAPES)
def test_tcgen05_bf16(a_k, b_k, M, N, K):
    assert L.lib().mvf_has_tcgen05() == 1, "tcgen05 path unavailable on this device"
    # leading dimensions must be multiples of 8 elements (16-byte TMA rows): pad when the logical size is not
    padA = (-(K if a_k else M)) % 8
    padB = (-(K if b_k else N)) % 8
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn((M, K + padA) if a_k else (K, M + padA), generator=g, device="cuda").to(torch.bfloat16)
    B = torch.randn((N, K + padB) if b_k else (K, N + padB), generator=g, device="cuda").to(torch.bfloat16)
    Am = A[:, :K].double() if a_k else A[:, :M].double().t()
    Bm = B[:, :K].double().t() if b_k else B[:, :N].double()
    ref = Am @ Bm
    bias = torch.randn(N, device="cuda")
    C = _gemm(L.GEMM_TCGEN05, A, B, a_k, b_k, M, N, K, torch.float32, bias=bias)
    err = float((C.double() - (ref + bias.double())).abs().max() / ref.abs().max())
    assert err < 1e-5, f"fp32-out error {err}"
    Cb = _gemm(L.GEMM_TCGEN05, A, B, a_k, b_k, M, N, K, torch.bfloat16, flags=L.GEMM_RELU)
    errb = float((Cb.double() - ref.clamp_min(0)).abs().max() / ref.abs().max())
    assert errb < 8e-3, f"bf16-out error {errb}"
    # cross-check against the SIMT engine on identical bf16 operands
    Cs = _gemm(L.GEMM_SIMT, A, B, a_k, b_k, M, N, K, torch.float32, bias=bias)
    assert float((C - Cs).abs().max() / ref.abs().max()) < 1e-5


@pytest.mark.parametrize("a_k,b_k", [(1, 1), (1, 0), (0, 0), (0, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 64, 32), (256, 256, 128), (200, 72, 100), (3840, 512, 392), (130, 392, 516), (512, 1024, 3840)])
def test_tcgen05_tf32(a_k, b_k, M, N, K):
    """fp32 operands on the tensor cores (kind::tf32): compared with the exact product of the tf32-truncated operands
    (tight) and with the exact fp32 product (tf32-level tolerance)."""
    padA = (-(K if a_k else M)) % 4
    padB = (-(K if b_k else N)) % 4
    g = torch.Generator(device="cuda").manual_seed(M + 3 * N + K)
    A = torch.randn((M, K + padA) if a_k else (K, M + padA), generator=g, device="cuda")
    B = torch.randn((N, K + padB) if b_k else (K, N + padB), generator=g, device="cuda")
    Am = A[:, :K].double() if a_k else A[:, :M].double().t()
    Bm = B[:, :K].double().t() if b_k else B[:, :N].double()
    ref = Am @ Bm
    bias = torch.randn(N, device="cuda")
    C = _gemm(L.GEMM_TCGEN05, A, B, a_k, b_k, M, N, K, torch.float32, bias=bias)
    err = float((C.double() - (ref + bias.double())).abs().max() / ref.abs().max())
    assert err < 2e-3, f"tf32 error {err}"
    trunc = lambda t: (t.float().view(torch.int32) & ~0x1FFF).view(torch.float32).double()
    ref_t = trunc(Am) @ trunc(Bm)
    err_t = float((C.double() - (ref_t + bias.double())).abs().max() / ref.abs().max())
    assert err_t < 1e-3      # the hardware may round instead of truncate: only a loose bound is portable
    base = torch.randn(M, N, device="cuda")
    C2 = _gemm(L.GEMM_TCGEN05, A, B, a_k, b_k, M, N, K, torch.float32, flags=L.GEMM_ACCUM, split_k=0, C=base.clone())
    assert float((C2.double() - (ref + base.double())).abs().max() / ref.abs().max()) < 2e-3
    src = torch.randn(M, N, device="cuda")
    C3 = _gemm(L.GEMM_TCGEN05, A, B, a_k, b_k, M, N, K, torch.float32, relu_src=src, flags=L.GEMM_RELUMASK)
    assert float((C3.double() - ref * (src > 0)).abs().max() / ref.abs().max()) < 2e-3


@pytest.mark.parametrize("M,N,K", [(128, 64, 32), (256, 256, 128), (200, 72, 100), (3840, 512, 392), (130, 392, 516),
                                   (3840, 1024, 256), (1280, 128, 1024), (77, 1536, 388)])
def test_tcgen05_bf16x3_split(M, N, K):
    """MVF_GEMM_SPLIT3: fp32 operands split in shared memory into bf16 hi + lo, three kind::f16 MMAs per K slice.
    Error bound: the dropped lo*lo term and the 16-bit mantissa of hi+lo, ~2^-16 relative per product."""
    padK = (-K) % 4
    g = torch.Generator(device="cuda").manual_seed(M + 5 * N + K)
    A = torch.randn(M, K + padK, generator=g, device="cuda")
    B = torch.randn(N, K + padK, generator=g, device="cuda")
    ref = A[:, :K].double() @ B[:, :K].double().t()
    bias = torch.randn(N, device="cuda")
    C = _gemm(L.GEMM_TCGEN05, A, B, 1, 1, M, N, K, torch.float32, bias=bias, flags=L.GEMM_SPLIT3)
    err = float((C.double() - (ref + bias.double())).abs().max() / ref.abs().max())
    assert err < 2e-5, f"bf16x3 error {err}"
    # exact model: product of the (hi + lo) operands minus the lo*lo term, fp32 accumulation
    def split(t):
        hi = t.bfloat16().float()
        lo = (t - hi).bfloat16().float()
        return hi.double(), lo.double()
    ah, al = split(A[:, :K]); bh, bl = split(B[:, :K])
    model = ah @ bh.t() + ah @ bl.t() + al @ bh.t()
    assert float((C.double() - (model + bias.double())).abs().max() / ref.abs().max()) < 1e-5
    Ct = _gemm(L.GEMM_TCGEN05, A, B, 1, 1, M, N, K, torch.float32, bias=bias)
    err_t = float((Ct.double() - (ref + bias.double())).abs().max() / ref.abs().max())
    assert err < err_t / 8, (err, err_t)          # and it really is far tighter than plain tf32
    C2 = _gemm(L.GEMM_TCGEN05, A, B, 1, 1, M, N, K, torch.float32, bias=bias, flags=L.GEMM_SPLIT3 | L.GEMM_RELU)
    assert float((C2.double() - (ref + bias.double()).clamp_min(0)).abs().max() / ref.abs().max()) < 2e-5


def _presplit(W, K):
    """[N, K] fp32 -> container [N, round_up(K, 32)] 'floats' whose 32-float blocks hold [hi(32) | lo(32)] bf16."""
    N = W.shape[0]
    Kp = (K + 31) // 32 * 32
    Wp = torch.zeros(N, Kp, device=W.device)
    Wp[:, :K] = W[:, :K]
    hi = Wp.bfloat16()
    lo = (Wp - hi.float()).bfloat16()
    cont = torch.stack([hi.view(N, Kp // 32, 32), lo.view(N, Kp // 32, 32)], dim=2).reshape(N, Kp * 2).contiguous()
    return cont.view(torch.float32).view(N, Kp)


@pytest.mark.parametrize("split_k", [1, 0])
@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (200, 72, 100), (3840, 512, 392), (3840, 256, 1024), (3840, 384, 2304)])
def test_tcgen05_bf16x3_presplit_weights(M, N, K, split_k):
    """MVF_GEMM_B_PRESPLIT: the weight operand arrives already split (packed once per step); split_k = 0 lets long
    contractions be split along K with the TMA reduce-add epilogue."""
    g = torch.Generator(device="cuda").manual_seed(M + N + 7 * K)
    padK = (-K) % 4
    A = torch.randn(M, K + padK, generator=g, device="cuda")
    W = torch.randn(N, K, generator=g, device="cuda")
    ref = A[:, :K].double() @ W.double().t()
    bias = torch.randn(N, device="cuda")
    Wc = _presplit(W, K)
    C = _gemm(L.GEMM_TCGEN05, A, Wc, 1, 1, M, N, K, torch.float32, bias=bias, flags=L.GEMM_SPLIT3 | L.GEMM_B_PRESPLIT,
              split_k=split_k)
    err = float((C.double() - (ref + bias.double())).abs().max() / ref.abs().max())
    assert err < 2e-5, err
    C1 = _gemm(L.GEMM_TCGEN05, A, W, 1, 1, M, N, K, torch.float32, bias=bias, flags=L.GEMM_SPLIT3) if K % 4 == 0 else None
    if C1 is not None:
        assert float((C - C1).abs().max() / ref.abs().max()) < 2e-5      # same operands, different staging (and K split)


@pytest.mark.parametrize("split", [0, 2, 7])
def test_tcgen05_split_k_weight_gradient_shape(split):
    """dW = dY^T X with a long contraction and few output tiles: both operands MN-major, fp32 atomics."""
    M, N, K = 768, 264, 196 * 40
    g = torch.Generator(device="cuda").manual_seed(3)
    dY = (torch.randn(K, M, generator=g, device="cuda") * 0.1).to(torch.bfloat16)
    X = torch.randn(K, N, generator=g, device="cuda").to(torch.bfloat16)
    ref = dY.double().t() @ X.double()
    base = torch.randn(M, N, device="cuda")
    C = _gemm(L.GEMM_TCGEN05, dY, X, 0, 0, M, N, K, torch.float32, flags=L.GEMM_ACCUM, split_k=split, C=base.clone())
    err = float((C.double() - (ref + base.double())).abs().max() / ref.abs().max())
    assert err < 2e-5
    C0 = _gemm(L.GEMM_TCGEN05, dY, X, 0, 0, M, N, K, torch.float32, split_k=split)
    assert float((C0.double() - ref).abs().max() / ref.abs().max()) < 2e-5


def test_tcgen05_relumask_and_alignment_errors():
    M, N, K = 256, 128, 64
    A, B, ref = _operands(M, N, K, 1, 1, torch.bfloat16)
    src = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    C = _gemm(L.GEMM_TCGEN05, A, B, 1, 1, M, N, K, torch.bfloat16, relu_src=src, flags=L.GEMM_RELUMASK)
    want = ref * (src.double() > 0)
    assert float((C.double() - want).abs().max() / ref.abs().max()) < 8e-3
    A2 = torch.randn(M, K + 4, device="cuda").to(torch.bfloat16)      # row stride 136 B: not a multiple of 16
    with pytest.raises(RuntimeError, match="multiple of 8"):
        _gemm(L.GEMM_TCGEN05, A2, B, 1, 1, M, N, K, torch.float32)
    with pytest.raises(RuntimeError, match="fp32 output"):
        _gemm(L.GEMM_TCGEN05, A.float(), B.float(), 1, 1, M, N, K, torch.bfloat16)
