"""Run the same forward twice (same seed) at the cfg2 shard shape and report, per saved stage, the relative difference
between the two runs: localises any run-to-run nondeterminism (split-K reduction order vs. a real race)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import mvf_oracle as O
from tests import helpers as H
from video_rep_learning_b200 import _lib as L

hc = O.HeadCfg(c_in=2304, train_frames=20)
Pm = O.init_params(hc, seed=3)
g = torch.Generator().manual_seed(5)
Bv, T, P = 8, 20, 196
tokens = torch.randn(2 * Bv, T, P, hc.c_in, generator=g).bfloat16()
_, seq_lens, steps, masks = O.synth_batch(Bv, T, 1, 1, seed=6)
def run():
    r = H.run_cuda(hc, Pm, None, tokens, masks, seq_lens, steps, dtype=torch.bfloat16, drop_p=float(os.environ.get("DROP", "0.1")), seed=77)
    cs, plan = r["cs"], r["cs"].plan
    regs = {}
    for n in ["wq", "px", "attn", "ent32", "h0", "fc0.x", "fc0.a", "fc1.x", "fc1.a", "h3", "z0", "l0.qkv", "l0.ctx", "z1", "l0.f", "z2", "z4", "z6", "y"]:
        try:
            regs[n] = plan.region(cs.head_save, n).float().clone()
        except RuntimeError:
            pass
    regs["e"] = r["e"].cuda()
    regs["grads"] = H.grad_vector(r["grads"], list(Pm.keys())).float().cuda()
    return regs
a, b = run(), run()
for n in a:
    d = float((a[n].double() - b[n].double()).norm() / (a[n].double().norm() + 1e-30))
    print(f"{n:10s} rel diff {d:.3e}   max abs {float((a[n] - b[n]).abs().max()):.3e}")
