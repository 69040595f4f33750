"""Host-side cost of one training step (enqueue only, no device sync inside): where does the CPU time go?
Times the public-API step loop and, separately, each C-ABI call (monkey-patched ctypes entry points)."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import mvf_oracle as O
from video_rep_learning_b200 import _lib as L, engine
from video_rep_learning_b200.algos import get_algo
from video_rep_learning_b200.config import mvf_cfg
from video_rep_learning_b200.models import build_model

dev = torch.device("cuda", 0)
Bv, T, P, C_in = 32, 20, 196, 2304
cfg = mvf_cfg(c_in=C_in, num_frames=T)
class _NB(torch.nn.Module):
    def forward(self, x): raise RuntimeError
model = build_model(cfg, backbone=_NB()).to(dev); model.train()
hc = O.HeadCfg(c_in=C_in, train_frames=T, drop_p=0.1)
model.load_state_dict(O.init_params(hc, seed=1), strict=False)
algo = get_algo(cfg)
tok = torch.randn(2 * Bv, T, P, C_in, device=dev).to(torch.bfloat16)
_, seq_lens, steps, masks = O.synth_batch(Bv, T, 1, 1, seed=1)
sl, st_, mk = seq_lens.to(dev), steps.to(dev), masks.to(dev)
params = [p for n, p in model.named_parameters() if "backbone" not in n]
lib = L.lib()
acc = collections.defaultdict(float); cnt = collections.defaultdict(int)
def wrap(name):
    fn = getattr(lib, name)
    def w(*a):
        t0 = time.perf_counter(); r = fn(*a); acc[name] += time.perf_counter() - t0; cnt[name] += 1; return r
    setattr(lib, name, w)
for nm in ("mvf_head_forward", "mvf_head_backward", "mvf_proj_forward", "mvf_proj_backward", "mvf_scl_fwd_bwd", "mvf_unpack_grads"):
    wrap(nm)
def step():
    for p in params: p.grad = None
    e = model.forward_tokens(tok, video_masks=mk, project=True)
    loss = algo.compute_sequence_loss(e.view(Bv, 2, T, -1), sl, st_, mk)["loss"]
    loss.backward()
for _ in range(5): step()
torch.cuda.synchronize(); acc.clear(); cnt.clear()
K = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(K): step()
t_enq = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
print(f"host enqueue {t_enq / K * 1e3:.3f} ms/step   device span {e0.elapsed_time(e1) / K:.3f} ms/step")
for k, v in sorted(acc.items(), key=lambda x: -x[1]):
    print(f"  {k:20s} {v / K * 1e3:.3f} ms/step ({cnt[k] // K} calls)")
print(f"  python + autograd remainder {(t_enq - sum(acc.values())) / K * 1e3:.3f} ms/step")
