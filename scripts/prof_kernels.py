"""Stand-alone launches of the dominant kernels at the bench (cfg2) shapes, for `ncu --set full` captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_rep_learning_b200 import _lib as L

which = sys.argv[1] if len(sys.argv) > 1 else "kv_fwd"
lib = L.lib()
st = torch.cuda.current_stream().cuda_stream
F, P, C, SPC, E = 1280, 196, 2304, 384, 3
M = F * P
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(M, C, device="cuda", generator=g).to(torch.bfloat16)
if which == "kv_fwd":
    W = (torch.randn(2 * SPC, C, device="cuda", generator=g) / 48).to(torch.bfloat16)
    b = torch.zeros(2 * SPC, device="cuda")
    KV = torch.empty(M, 2 * SPC, dtype=torch.bfloat16, device="cuda")
    for _ in range(3):
        L.check(lib.mvf_gemm(L.GEMM_TCGEN05, 1, 1, 1, 1, M, 2 * SPC, C, L.ptr(X), C, L.ptr(W), C, L.ptr(KV), 2 * SPC, L.ptr(b), None, 0, 0, 1, st))
elif which == "kv_dw":
    dKV = (torch.randn(M, 2 * SPC, device="cuda", generator=g) * 0.01).to(torch.bfloat16)
    dW = torch.zeros(2 * SPC, C, device="cuda")
    for _ in range(3):
        L.check(lib.mvf_gemm(L.GEMM_TCGEN05, 1, 0, 0, 0, 2 * SPC, C, M, L.ptr(dKV), 2 * SPC, L.ptr(X), C, L.ptr(dW), C, None, None, 0, L.GEMM_ACCUM, 0, st))
elif which in ("xattn_fwd", "xattn_bwd"):
    KV = torch.randn(M, 2 * SPC, device="cuda", generator=g).to(torch.bfloat16)
    qs = torch.randn(E, SPC, device="cuda", generator=g) * 0.05
    qb = torch.zeros(SPC, device="cuda")
    attn = torch.empty(F, E, P, device="cuda")
    ent = torch.empty(F * E, 392, dtype=torch.bfloat16, device="cuda")
    ent32 = torch.empty(F * E, SPC, device="cuda")
    for _ in range(3):
        L.check(lib.mvf_xattn_pool_fwd(1, F, P, E, SPC, L.ptr(KV), L.ptr(qs), L.ptr(qb), L.ptr(attn), L.ptr(ent), 392, L.ptr(ent32), 1, 0.1, 7, st))
    if which == "xattn_bwd":
        dkv = torch.empty_like(KV)
        dq, dqb, dbk, dbv = torch.zeros(E, SPC, device="cuda"), torch.zeros(SPC, device="cuda"), torch.zeros(SPC, device="cuda"), torch.zeros(SPC, device="cuda")
        for _ in range(3):
            L.check(lib.mvf_xattn_pool_bwd(1, F, P, E, SPC, L.ptr(KV), L.ptr(qs), L.ptr(qb), L.ptr(attn), L.ptr(ent), 392, L.ptr(ent32), 1, 0.1, 7,
                                           L.ptr(dkv), L.ptr(dq), L.ptr(dqb), L.ptr(dbk), L.ptr(dbv), st))
torch.cuda.synchronize()
print("done", which)
