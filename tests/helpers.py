"""Shared test plumbing: golden loading, oracle runs, and driving the CUDA path through the C ABI."""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import numpy as np
import torch

from oracle import mvf_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def meta(fname="meta.json"):
    with open(os.path.join(GOLDEN, fname)) as f:
        return json.load(f)


def load_case(name, meta_file="meta.json"):
    m = meta(meta_file)["cases"][name]
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    kw = dict(m["head_cfg"])
    kw["fc_channels"] = tuple(kw["fc_channels"])
    hc = O.HeadCfg(**kw)
    P = {k[len("param:"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param:")}
    G = {k[len("grad:"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad:")}
    B = {k[len("buf:"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("buf:")}
    return m, hc, z, P, G, B


def spec_from_headcfg(hc: O.HeadCfg, drop_p: Optional[float] = None):
    from video_rep_learning_b200 import engine
    return engine.HeadSpec(c_in=hc.c_in, n_entities=hc.n_entities, pool_channels=hc.pool_channels,
                           fc_channels=tuple(hc.fc_channels), hidden=hc.hidden, d_ff=hc.d_ff, n_heads=hc.n_heads,
                           n_layers=hc.n_layers, emb=hc.emb, proj=hc.proj, one_hot=hc.one_hot, final=hc.final,
                           train_frames=hc.train_frames, drop_p=hc.drop_p if drop_p is None else drop_p,
                           ln_eps=hc.ln_eps, bn_eps=hc.bn_eps, bn_momentum=hc.bn_momentum,
                           pool_kind=getattr(hc, "pool_kind", "lstp"), cls_dim=getattr(hc, "cls_dim", 0))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().reshape(-1).cpu(), b.double().reshape(-1).cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


def grad_vector(g: Dict[str, torch.Tensor], keys):
    return torch.cat([g[k].double().reshape(-1).cpu() for k in keys])


def run_cuda(hc: O.HeadCfg, P: Dict[str, torch.Tensor], buf: Optional[Dict[str, torch.Tensor]], tokens, masks, seq_lens,
             steps, *, dtype=torch.float32, negative_type="single_noself", training=True, drop_p=0.0, seed=0,
             backend=0, quirk=True, project=True, device="cuda", pool_mode=0, cls_emb=None):
    """One step of the CUDA path through engine.ModelFn + engine.SCLFn.  Returns a dict with emb (head output),
    e (normalised projection), loss, grads (by reference state_dict name), new BN buffers, the CallState."""
    from video_rep_learning_b200 import engine
    spec = spec_from_headcfg(hc, drop_p)
    dev = torch.device(device)
    names = list(O.param_shapes(hc).keys())
    params = [P[n].to(dev).float().clone().requires_grad_(True) for n in names]
    bn_names = O.bn_buffer_names(hc)
    buf = buf if buf is not None else O.init_bn_buffers(hc)
    running, tracked = [], []
    for pre in bn_names:
        running += [buf[pre + ".running_mean"].to(dev).float().clone(), buf[pre + ".running_var"].to(dev).float().clone()]
        tracked.append(buf[pre + ".num_batches_tracked"].to(dev).clone())
    opts = engine.RunOptions(gemm_backend=backend, scl_quirk=quirk, pool_mode=pool_mode)
    cs = engine.CallState(spec=spec, opts=opts, training=training, bn_running=running, bn_tracked=tracked,
                          project=1 if project else 0, seed=seed,
                          cls_emb=None if cls_emb is None else cls_emb.to(dev).float())
    tok = tokens.to(dev).to(dtype)
    m = None if masks is None else masks.to(dev)
    out = engine.ModelFn.apply(tok, m, cs, *params)
    res = {"e": out.detach().cpu(), "cs": cs, "params": params, "names": names}
    plan = cs.plan
    res["emb"] = None
    BV, T = tok.shape[0], tok.shape[1]
    res["emb"] = plan.region(cs.proj_save, "proj:emb").float().cpu().view(BV, T, -1) if project else None
    if seq_lens is not None:
        Bv = BV // 2
        loss = engine.SCLFn.apply(out.view(Bv, 2, T, -1), seq_lens.to(dev), steps.to(dev), m, 0.1, 10.0, negative_type, quirk)
        loss.backward()
        res["loss"] = loss.detach().cpu()
        res["grads"] = {n: p.grad.detach().cpu() for n, p in zip(names, params)}
    nb = {}
    for i, pre in enumerate(bn_names):
        nb[pre + ".running_mean"] = running[2 * i].cpu()
        nb[pre + ".running_var"] = running[2 * i + 1].cpu()
        nb[pre + ".num_batches_tracked"] = tracked[i].cpu()
    res["bufs"] = nb
    return res


def run_oracle(hc: O.HeadCfg, P, buf, tokens, masks, seq_lens, steps, *, dtype=torch.float64,
               negative_type="single_noself", drop_masks=None, cls_emb=None):
    Pr = {k: v.clone().to(dtype).requires_grad_(True) for k, v in P.items()}
    b2 = None if buf is None else {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in buf.items()}
    emb, nb, aux = O.head_forward(Pr, b2, tokens.to(dtype), None if masks is None else masks.to(dtype), hc, True,
                                  drop_masks, return_aux=True, cls_emb=None if cls_emb is None else cls_emb.to(dtype))
    e, nb2 = O.proj_forward(Pr, b2, emb, hc, True)
    nb.update(nb2)
    Bv, T = tokens.shape[0] // 2, tokens.shape[1]
    loss = O.scl_loss_dense(e.view(Bv, 2, T, -1), seq_lens, steps, masks.to(dtype), negative_type=negative_type)
    loss.backward()
    return dict(emb=emb.detach(), e=e.detach(), loss=loss.detach(), grads={k: v.grad for k, v in Pr.items()}, bufs=nb,
                aux={k: v.detach() for k, v in aux.items() if v is not None})


def run_oracle_quantized(hc, P, tokens, masks, seq_lens, steps, negative_type="single_noself", kv_bf16=True):
    """The reference algorithm in fp64 on the SAME quantised operands the bf16 path consumes.  Dense pooling
    (kv_bf16=True): bf16 tokens, bf16 W_k|W_v, and K|V rounded to bf16 (straight-through).  Folded pooling
    (kv_bf16=False): only the tokens are bf16 -- K|V never exist and the weights stay fp32.  Separates the error of the
    implementation from the error that "bf16 operands" makes inevitable (which the 1/tau of SCL amplifies ~40x from
    embeddings to gradients)."""
    import torch.nn.functional as F

    if not kv_bf16:
        return run_oracle(hc, P, None, tokens.bfloat16().float(), masks, seq_lens, steps, dtype=torch.float64,
                          negative_type=negative_type)

    def bf(x):
        return x.float().bfloat16().double()

    Pq = dict(P)
    for k in ("embed.pooling.cross_att.linear_K2d.weight", "embed.pooling.cross_att.linear_V2d.weight"):
        Pq[k] = P[k].bfloat16().float()
    orig = O.xattn_pool

    def xattn_q(Pd, tok, cfg):
        pre = "embed.pooling.cross_att."
        K = F.linear(tok, Pd[pre + "linear_K2d.weight"], Pd[pre + "linear_K2d.bias"])
        V = F.linear(tok, Pd[pre + "linear_V2d.weight"], Pd[pre + "linear_V2d.bias"])
        K = K + (bf(K.detach()) - K.detach())
        V = V + (bf(V.detach()) - V.detach())
        Q = Pd[pre + "Q_s"][0] + Pd[pre + "Q_s_b"]
        A = torch.softmax(torch.einsum("btpc,ec->btep", K, Q) / np.sqrt(cfg.pool_channels), -1)
        return torch.einsum("btep,btpc->btec", A, V), A

    O.xattn_pool = xattn_q
    try:
        return run_oracle(hc, Pq, None, tokens.bfloat16().float(), masks, seq_lens, steps, dtype=torch.float64,
                          negative_type=negative_type)
    finally:
        O.xattn_pool = orig
