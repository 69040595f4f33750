"""SCL kernel (mvf_scl_fwd_bwd: similarity + softmax + Gaussian-KL + gradient, per video pair) at the named shapes
(latency) and on scaled-up synthetic batches (achieved HBM GB/s against the algorithmic bytes of SURVEY.md 8d:
read 2*T*D*4 + steps/masks/seq_lens, write 2*T*D*4 per pair).  All frames valid in the scaled runs (the masked-column
quirk of scl.py:80 couples every masked frame of the local batch to every row, which is quadratic in the batch)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_rep_learning_b200 import _lib as L

lib = L.lib(); st = torch.cuda.current_stream().cuda_stream
peak = 6538.9
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = []
for Bv, T, D, masked in ((32, 20, 128, True), (8, 80, 256, True), (4, 240, 128, True), (4096, 20, 128, False), (65536, 20, 128, False),
                         (16384, 80, 128, False), (2048, 240, 128, False)):
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
    call = lambda: L.check(lib.mvf_scl_fwd_bwd(L.ptr(e), L.ptr(sl), L.ptr(steps), L.ptr(mk), Bv, T, D, 0.1, 10.0, 0, 1,
                                               L.ptr(loss), L.ptr(dE), L.ptr(ws), nb, st))
    for _ in range(3):
        call()
    reps = 20 if Bv <= 4096 else 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = Bv * (2 * T * D * 4 * 2 + 2 * T * 12 + 16)
    rec = dict(pairs=Bv, T=T, D=D, masked_frames=masked, ms=ms, algorithmic_MB=nbytes / 1e6, GBps=nbytes / ms / 1e6,
               frac_of_measured_hbm=nbytes / ms / 1e6 / peak, loss=float(loss))
    out.append(rec)
    print(json.dumps(rec))
