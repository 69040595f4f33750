"""One SCL shape, a few calls (for ncu): python scripts/scl_one.py Bv T D masked(0/1) [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_rep_learning_b200 import _lib as L

Bv, T, D, masked = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
lib = L.lib(); st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(Bv + T)
e = torch.nn.functional.normalize(torch.randn(Bv, 2, T, D, device="cuda", generator=g), dim=-1).contiguous()
sl = torch.full((Bv, 2), 3 * T, dtype=torch.int64, device="cuda")
steps = torch.sort(torch.randint(0, 3 * T, (Bv, 2, T), device="cuda", generator=g), dim=-1).values
mk = torch.ones(Bv, 2, T, device="cuda")
if masked:
    mk[:, :, -max(1, T // 5):] = 0
nb = lib.mvf_scl_ws_bytes(Bv, T, D)
ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
loss = torch.empty((), device="cuda"); dE = torch.empty_like(e)
for _ in range(reps):
    L.check(lib.mvf_scl_fwd_bwd(L.ptr(e), L.ptr(sl), L.ptr(steps), L.ptr(mk), Bv, T, D, 0.1, 10.0, 0, 1,
                                L.ptr(loss), L.ptr(dE), L.ptr(ws), nb, st))
torch.cuda.synchronize()
print(float(loss))
