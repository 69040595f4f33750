"""One chain-sized GEMM (3840 x 512 x 512, fp32 operands) for an `ncu --set full` capture: argv[1] = tf32 | split3 | bf16."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_rep_learning_b200 import _lib as L
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
lib = L.lib(); st = torch.cuda.current_stream().cuda_stream
M, N, K = 3840, 512, 512
dt = torch.bfloat16 if mode == "bf16" else torch.float32
A = torch.randn(M, K, device="cuda").to(dt); B = torch.randn(N, K, device="cuda").to(dt)
C = torch.empty(M, N, device="cuda"); bias = torch.randn(N, device="cuda")
ab = L.MVF_BF16 if mode == "bf16" else L.MVF_F32
fl = L.GEMM_SPLIT3 if mode == "split3" else 0
for _ in range(4):
    L.check(lib.mvf_gemm(L.GEMM_TCGEN05, ab, L.MVF_F32, 1, 1, M, N, K, L.ptr(A), K, L.ptr(B), K, L.ptr(C), N, L.ptr(bias), None, 0, fl, 1, st))
torch.cuda.synchronize()
print("done")
