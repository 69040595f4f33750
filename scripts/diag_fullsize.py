"""Per-tensor gradient error of the fp32 (exact-FMA) CUDA path against the fp64 oracle at a full-size shard shape.
    python scripts/diag_fullsize.py [videos] [T] [entities] [fc]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import mvf_oracle as O
from tests import helpers as H

Bv = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
E = int(sys.argv[3]) if len(sys.argv) > 3 else 3
FC = int(sys.argv[4]) if len(sys.argv) > 4 else 512
hc = O.HeadCfg(c_in=2304, n_entities=E, fc_channels=(FC, FC), train_frames=T)
Pm = O.init_params(hc, seed=3)
g = torch.Generator().manual_seed(5)
tokens = torch.randn(2 * Bv, T, 196, hc.c_in, generator=g)
_, seq_lens, steps, masks = O.synth_batch(Bv, T, 1, 1, seed=6)
torch.set_num_threads(os.cpu_count() or 1)
keys = list(Pm.keys())
o = H.run_oracle(hc, Pm, None, tokens, masks, seq_lens, steps, dtype=torch.float64)
r = H.run_cuda(hc, Pm, None, tokens, masks, seq_lens, steps, dtype=torch.float32)
G = float(H.grad_vector(o["grads"], keys).norm())
print(f"Bv {Bv} T {T} E {E} FC {FC}: e rel {H.rel_l2(r['e'], o['e']):.2e} emb rel {H.rel_l2(r['emb'], o['emb']):.2e} "
      f"loss rel {abs(float(r['loss']) - float(o['loss'])) / float(o['loss']):.2e} "
      f"grad rel {H.rel_l2(H.grad_vector(r['grads'], keys), H.grad_vector(o['grads'], keys)):.2e}")
rows = []
for k in keys:
    a, b = r["grads"][k].double(), o["grads"][k]
    rows.append((float((a - b).norm()) / G, k, float(b.norm()) / G, float((a - b).norm() / (b.norm() + 1e-300))))
for err, k, share, rel in (rows if os.environ.get("DIAG_ALL") else sorted(rows, reverse=True)[:14]):
    print(f"   {k:58s} |g|/|G| {share:.2e}  err/|G| {err:.2e}  rel {rel:.2e}")
