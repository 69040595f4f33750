"""Micro-benchmark of the chain GEMM shapes (cfg2: 3840 entity rows / 3840 encoder rows / 1280 frame rows) in the three
fp32-operand modes of the tcgen05 engine: tf32, bf16x3 (SPLIT3) and -- as the speed reference -- bf16 operands.
CUDA events on the launching stream, 20 warm-up + 100 timed launches per case."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_rep_learning_b200 import _lib as L

lib = L.lib()
st = torch.cuda.current_stream().cuda_stream
SHAPES = [(3840, 512, 392), (3840, 512, 512), (3840, 256, 512), (3840, 768, 256), (3840, 256, 256), (3840, 1024, 256),
          (3840, 256, 1024), (1280, 128, 256), (1280, 128, 128), (3840, 384, 2304)]


def run(M, N, K, ab, flags, c_dtype=torch.float32):
    dt = torch.bfloat16 if ab == L.MVF_BF16 else torch.float32
    A = torch.randn(M, K, device="cuda").to(dt)
    B = torch.randn(N, K, device="cuda").to(dt)
    C = torch.empty(M, N, device="cuda", dtype=c_dtype)
    bias = torch.randn(N, device="cuda")
    cd = L.MVF_BF16 if c_dtype == torch.bfloat16 else L.MVF_F32
    call = lambda: L.check(lib.mvf_gemm(L.GEMM_TCGEN05, ab, cd, 1, 1, M, N, K, L.ptr(A), K, L.ptr(B), K, L.ptr(C), N,
                                        L.ptr(bias), None, 0, flags, 1, st))
    for _ in range(20):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 10.0   # us per launch


print(f"{'M':>6} {'N':>5} {'K':>5} | {'tf32 us':>8} {'bf16x3 us':>9} {'bf16 us':>8} {'bf16->bf16':>10} | GF")
for M, N, K in SHAPES:
    t = run(M, N, K, L.MVF_F32, 0)
    s = run(M, N, K, L.MVF_F32, L.GEMM_SPLIT3)
    b = run(M, N, K, L.MVF_BF16, 0)
    bb = run(M, N, K, L.MVF_BF16, 0, torch.bfloat16)
    print(f"{M:6d} {N:5d} {K:5d} | {t:8.2f} {s:9.2f} {b:8.2f} {bb:10.2f} | {2e-9 * M * N * K:.2f}")

print("\nK sweep at 3840 x 512 (fp32 out): fixed cost vs per-k-block cost")
for K in (32, 64, 128, 256, 512, 1024, 2048):
    print(f"  K={K:5d}  tf32 {run(3840, 512, K, L.MVF_F32, 0):7.2f} us   bf16x3 {run(3840, 512, K, L.MVF_F32, L.GEMM_SPLIT3):7.2f} us   "
          f"bf16 {run(3840, 512, K, L.MVF_BF16, 0):7.2f} us")
print("N sweep at 3840 x N x 256 tf32")
for N in (64, 128, 256, 512, 1024):
    print(f"  N={N:5d}  tf32 {run(3840, N, 256, L.MVF_F32, 0):7.2f} us")
