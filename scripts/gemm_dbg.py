"""Phase timeline (globaltimer stamps of CTA 0) of chain-sized GEMMs: run with MVF_GEMM_DBG=1."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_rep_learning_b200 import _lib as L
lib = L.lib(); st = torch.cuda.current_stream().cuda_stream
def run(M, N, K, a_k, b_k, fl, split_k=1, reps=3):
    A = torch.randn((M, K) if a_k else (K, M), device="cuda"); B = torch.randn((N, K) if b_k else (K, N), device="cuda")
    C = torch.zeros(M, N, device="cuda"); bias = torch.randn(N, device="cuda")
    for _ in range(reps):
        L.check(lib.mvf_gemm(L.GEMM_TCGEN05, 0, 0, a_k, b_k, M, N, K, L.ptr(A), A.stride(0), L.ptr(B), B.stride(0), L.ptr(C), N, L.ptr(bias), None, 0, fl, split_k, st))
print("fwd split3 3840x512x512", file=sys.stderr); run(3840, 512, 512, 1, 1, L.GEMM_SPLIT3)
print("fwd tf32 3840x512x512", file=sys.stderr); run(3840, 512, 512, 1, 1, 0)
print("fwd split3 3840x256x1024", file=sys.stderr); run(3840, 256, 1024, 1, 1, L.GEMM_SPLIT3)
print("fwd split3 3840x384x2304", file=sys.stderr); run(3840, 384, 2304, 1, 1, L.GEMM_SPLIT3)
print("dX tf32 3840x256x1024 (B MN-major)", file=sys.stderr); run(3840, 256, 1024, 1, 0, 0)
print("dW tf32 1024x256 K=3840 split auto (MN-major both, accumulate)", file=sys.stderr); run(1024, 256, 3840, 0, 0, L.GEMM_ACCUM, 0)
print("dW tf32 384x2304 K=3840 split auto", file=sys.stderr); run(384, 2304, 3840, 0, 0, L.GEMM_ACCUM, 0)
