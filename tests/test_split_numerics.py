"""CPU: the arithmetic model behind the "bf16 hi/lo" tensor-core products (SCL, temporal attention, bf16x3 GEMMs).

An fp32 operand x is carried as hi = bf16(x), lo = bf16(x - hi); a product uses hi*hi + hi*lo + lo*hi with fp32 accumulation
(csrc/scl_mma.cu:mma3, attention_tc.cu, attention_fa.cu, gemm_tc.cu SPLIT3).  This emulates exactly that on the host and pins
the error figures DESIGN.md quotes: ~2^-17 per product, a few 1e-6 (rms; 1.4e-5 on the worst element) on an SCL-sized logit
block after the 1/tau = 10 amplification -- against ~2^-9 for plain bf16 operands, which would break the 1e-5 tolerance by two orders of magnitude."""
import torch


def _split(x: torch.Tensor):
    hi = x.to(torch.bfloat16).to(torch.float32)
    lo = (x - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def _matmul_split3(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    ah, al = _split(a)
    bh, bl = _split(b)
    # three bf16 x bf16 products, exact in fp32, accumulated in fp32 (the tensor core adds in fp32 as well)
    return (al @ bh.t()) + (ah @ bl.t()) + (ah @ bh.t())


def test_split_residue_is_below_2_to_minus_16():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1 << 16, generator=g)
    hi, lo = _split(x)
    rel = ((x - hi - lo).abs() / x.abs().clamp_min(1e-30)).max()
    assert float(rel) <= 2.0 ** -16
    assert float(((x - hi).abs() / x.abs()).max()) <= 2.0 ** -8      # one bf16 alone


def test_scl_sized_logit_block_error():
    """T = 20 frames, D = 128 unit rows, logits / tau with tau = 0.1: error of exp(l / tau) relative to fp64."""
    g = torch.Generator().manual_seed(1)
    e0 = torch.nn.functional.normalize(torch.randn(20, 128, generator=g), dim=-1)
    e1 = torch.nn.functional.normalize(torch.randn(20, 128, generator=g), dim=-1)
    ref = e0.double() @ e1.double().t()
    s3 = _matmul_split3(e0, e1).double()
    s1 = (e0.to(torch.bfloat16).float() @ e1.to(torch.bfloat16).float().t()).double()
    err3 = (s3 - ref) * 10                                # error of the exponent l / tau
    err1 = float(((s1 - ref) * 10).abs().max())
    assert float(err3.abs().max()) < 3e-5, float(err3.abs().max())    # worst element of the block (measured 1.4e-5)
    assert float(err3.pow(2).mean().sqrt()) < 8e-6                    # rms: what the loss / gradient norms see
    assert err1 > 1e-3, err1                              # plain bf16 operands: two orders of magnitude worse
    rel3 = (torch.exp(s3 * 10) / torch.exp(ref * 10) - 1)
    assert float(rel3.abs().max()) < 3e-5 and float(rel3.pow(2).mean().sqrt()) < 8e-6


def test_gradient_product_error_is_a_few_1e_6():
    """dE = C . E_cols with the coefficient tile split the same way (the accumulator fragments re-used as the A operand)."""
    g = torch.Generator().manual_seed(2)
    C = torch.randn(20, 20, generator=g) * 1e-2
    E = torch.nn.functional.normalize(torch.randn(20, 128, generator=g), dim=-1)
    ref = C.double() @ E.double()
    got = _matmul_split3(C, E.t().contiguous()).double()
    rel = float((got - ref).norm() / ref.norm())
    assert rel < 8e-6, rel
