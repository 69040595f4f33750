"""Stand-alone timing of the folded pooling passes at the bench shape (cfg2: 1280 frames x 196 tokens x 2304 channels,
bf16, E = 3): host wall time per call (launch overhead) and device time by CUDA events; achieved GB/s."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_rep_learning_b200 import _lib as L

lib = L.lib(); st = torch.cuda.current_stream().cuda_stream
F, P, C, SPC, E = 1280, 196, 2304, 384, 3
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(F * P, C, device="cuda", generator=g).to(torch.bfloat16)
qs = torch.randn(E, SPC, device="cuda", generator=g) * 0.5
qb = torch.zeros(SPC, device="cuda")
Wk = (torch.rand(SPC, C, device="cuda", generator=g) * 2 - 1) / 48
wq = torch.empty(E, C, device="cuda"); attn = torch.empty(F, E, P, device="cuda"); px = torch.empty(F * E, C, device="cuda")
G = torch.randn(F * E, C, device="cuda", generator=g) * 0.01
dwq = torch.zeros(E, C, device="cuda")
L.check(lib.mvf_pool_fold_prep(L.ptr(qs), L.ptr(qb), L.ptr(Wk), E, SPC, C, L.ptr(wq), st))
fwd = lambda: L.check(lib.mvf_pool_fold_fwd(1, F, P, E, C, L.ptr(X), L.ptr(wq), L.ptr(attn), L.ptr(px), st))
fwd()
delta = (G * px).sum(-1).contiguous()      # what the fused head gets from ent_finish_bwd: <dEnt, ent - bv> = <G, px>
bwd = lambda: L.check(lib.mvf_pool_fold_bwd_delta(1, F, P, E, C, L.ptr(X), L.ptr(G), L.ptr(px), L.ptr(attn), L.ptr(delta), L.ptr(dwq), st))
tok = F * P * C * 2
for name, fn, nbytes in (("fwd", fwd, tok + F * E * C * 4 + F * E * P * 4), ("bwd", bwd, tok + 2 * F * E * C * 4 + F * E * P * 4)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"pool_fold_{name}: device {ms*1e3:.1f} us/launch  {nbytes/ms/1e6:.0f} GB/s  host enqueue {t_host/reps*1e6:.1f} us/call")
