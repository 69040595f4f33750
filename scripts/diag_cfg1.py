"""cfg1 shape (BASELINE configs[0]) in fp32: per-tensor gradient error of the folded and dense pooling paths against the
fp64 oracle and against the fp32 reference golden digest."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import mvf_oracle as O
from tests import helpers as H
from video_rep_learning_b200 import _lib as L

m = H.meta()["penn_cfg1"]
kw = dict(m["head_cfg"]); kw["fc_channels"] = tuple(kw["fc_channels"])
hc = O.HeadCfg(**kw)
P = O.init_params(hc, seed=m["seed"])
tokens, seq_lens, steps, masks = O.synth_batch(m["Bv"], m["T"], m["Ptok"], hc.c_in, seed=m["seed"])
torch.set_num_threads(os.cpu_count() or 1)
o = H.run_oracle(hc, P, None, tokens, masks, seq_lens, steps, dtype=torch.float64)
keys = list(P.keys())
for name, pm in (("folded", L.POOL_FOLDED), ("dense", L.POOL_DENSE)):
    r = H.run_cuda(hc, P, None, tokens, masks, seq_lens, steps, dtype=torch.float32, pool_mode=pm)
    print(f"== {name}: e rel {H.rel_l2(r['e'], o['e']):.2e} loss rel {abs(float(r['loss'])-float(o['loss']))/float(o['loss']):.2e} "
          f"grad rel {H.rel_l2(H.grad_vector(r['grads'], keys), H.grad_vector(o['grads'], keys)):.2e}")
    worst = []
    for k in keys:
        a, b = r["grads"][k].double(), o["grads"][k]
        rel = float((a - b).norm() / (b.norm() + 1e-30)); mx = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        dg = m["grad_digest"].get(k)
        dl2 = abs(float(a.norm()) - dg["l2"]) / dg["l2"] if dg and dg["l2"] > 0 else float("nan")
        hd = float((a.reshape(-1)[:6] - torch.tensor(dg["head"], dtype=torch.float64)).abs().max() / dg["absmax"]) if dg and dg["absmax"] > 0 else float("nan")
        worst.append((mx, k, rel, dl2, hd))
    for mx, k, rel, dl2, hd in sorted(worst, reverse=True)[:12]:
        print(f"   {k:60s} vs fp64: rel {rel:.2e} max/absmax {mx:.2e} | vs fp32 golden digest: l2 {dl2:.2e} head {hd:.2e}")
