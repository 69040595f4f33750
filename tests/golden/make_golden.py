"""Generate the committed golden vectors by running the REFERENCE's own modules (build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz, *.json and PINNING.txt

For every case the script (1) draws deterministic parameters / inputs with the oracle's generators,
(2) loads them into the reference's MultiEntityTransformerEmbModel + MLPHead + SCL (oracle/ref_shim.py),
(3) runs forward + backward there, (4) checks the oracle restatement against those outputs (this is the
"pinning" of the oracle; the max deviations are written to PINNING.txt) and (5) stores the reference's
outputs as fixtures.  The GPU box has no /root/reference, so tests only ever read the fixtures.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mvf_oracle as O  # noqa: E402
from oracle import ref_shim as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
LOG = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    LOG.append(s)


TINY = dict(c_in=48, n_entities=3, pool_channels=32, fc_channels=(64, 64), hidden=32, d_ff=64, n_heads=4,
            n_layers=2, emb=16, proj=16, train_frames=8)

CASES = {
    # name: (HeadCfg kwargs, Bv, T, Ptok, seed, negative_type)
    "tiny_penn": (dict(TINY), 2, 8, 16, 5, "single_noself"),
    "tiny_fg_avg": (dict(TINY, n_entities=4, final="avg", emb=24, proj=16, train_frames=12, fc_channels=(96, 96)), 3, 12, 16, 7, "single_noself"),
    "tiny_max_nohot": (dict(TINY, final="max", one_hot="none"), 2, 8, 16, 9, "single_noself"),
    "tiny_lin": (dict(TINY, final="lin", n_entities=2), 2, 8, 16, 11, "single_noself"),
    "tiny_batch_noself": (dict(TINY), 3, 8, 16, 13, "batch_noself"),
    "tiny_e1": (dict(TINY, n_entities=1), 2, 8, 9, 15, "single_noself"),
}


def run_reference(hc, P, buf, tokens, masks, seq_lens, steps, negative_type, Bv, T):
    cfg, head, proj, algo = R.build_reference_modules(hc, P)
    algo.negative_type = negative_type
    head.train()
    proj.train()
    x = R.tokens_to_nchw(tokens)
    emb = head(x, video_masks=masks, cls_emb=None)
    attn_last = head.pooling.cross_att.attn_matrix.detach().clone()          # [T,E,P] of the LAST video
    e = torch.nn.functional.normalize(proj(emb), dim=-1)
    e.retain_grad()
    emb.retain_grad()
    loss = algo.compute_sequence_loss(e.view(Bv, 2, T, -1), seq_lens, steps, masks)["loss"]
    loss.backward()
    grads = {"embed." + k: v.grad.detach().clone() for k, v in head.named_parameters()}
    grads.update({"ssl_projection." + k: v.grad.detach().clone() for k, v in proj.named_parameters()})
    bufs = {"embed." + k: v.detach().clone() for k, v in head.named_buffers()}
    bufs.update({"ssl_projection." + k: v.detach().clone() for k, v in proj.named_buffers()})
    return dict(emb=emb.detach(), e=e.detach(), loss=loss.detach(), grads=grads, bufs=bufs,
                attn_last=attn_last, d_e=e.grad.detach().clone(), d_emb=emb.grad.detach().clone())


def run_oracle(hc, P, buf, tokens, masks, seq_lens, steps, negative_type, Bv, T, dtype=torch.float32):
    Pr = {k: v.clone().to(dtype).requires_grad_(True) for k, v in P.items()}
    b2 = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in buf.items()}
    emb, nb = O.head_forward(Pr, b2, tokens.to(dtype), masks.to(dtype), hc, True)
    e, nb2 = O.proj_forward(Pr, b2, emb, hc, True)
    nb.update(nb2)
    loss = O.scl_loss_dense(e.view(Bv, 2, T, -1), seq_lens, steps, masks.to(dtype), negative_type=negative_type)
    loss.backward()
    return dict(emb=emb.detach(), e=e.detach(), loss=loss.detach(), grads={k: v.grad for k, v in Pr.items()}, bufs=nb)


def gcat(g, keys):
    return torch.cat([g[k].reshape(-1).double() for k in keys])


def make_case(name, spec):
    kw, Bv, T, Ptok, seed, neg = spec
    hc = O.HeadCfg(**kw)
    P = O.init_params(hc, seed=seed)
    buf = O.init_bn_buffers(hc)
    tokens, seq_lens, steps, masks = O.synth_batch(Bv, T, Ptok, hc.c_in, seed=seed)
    ref = run_reference(hc, P, buf, tokens, masks, seq_lens, steps, neg, Bv, T)
    orc = run_oracle(hc, P, buf, tokens, masks, seq_lens, steps, neg, Bv, T)
    keys = list(P.keys())
    gr, go = gcat(ref["grads"], keys), gcat(orc["grads"], keys)
    log(f"[{name}] oracle-vs-reference fp32: emb {float((orc['emb']-ref['emb']).abs().max()):.2e}  "
        f"e {float((orc['e']-ref['e']).abs().max()):.2e}  loss rel {float(abs(orc['loss']-ref['loss'])/abs(ref['loss'])):.2e}  "
        f"grad rel(L2) {float((gr-go).norm()/gr.norm()):.2e}  masks valid {int(masks.sum())}/{masks.numel()}")
    assert float((orc["emb"] - ref["emb"]).abs().max()) < 2e-5
    assert float(abs(orc["loss"] - ref["loss"]) / abs(ref["loss"])) < 1e-5
    assert float((gr - go).norm() / gr.norm()) < 1e-5
    for k, v in orc["bufs"].items():
        assert float((v.double() - ref["bufs"][k].double()).abs().max()) < 1e-5, k
    arrs = {"tokens": tokens.numpy(), "masks": masks.numpy(), "seq_lens": seq_lens.numpy(), "steps": steps.numpy(),
            "ref_emb": ref["emb"].numpy(), "ref_e": ref["e"].numpy(), "ref_loss": ref["loss"].numpy(),
            "ref_attn_last": ref["attn_last"].numpy(), "ref_d_e": ref["d_e"].numpy(), "ref_d_emb": ref["d_emb"].numpy()}
    for k in keys:
        arrs["param:" + k] = P[k].numpy()
        arrs["grad:" + k] = ref["grads"][k].numpy()
    for k, v in ref["bufs"].items():
        arrs["buf:" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    meta = dict(head_cfg={k: (list(v) if isinstance(v, tuple) else v) for k, v in hc.__dict__.items()},
                Bv=Bv, T=T, Ptok=Ptok, seed=seed, negative_type=neg)
    return meta


def make_eval_case():
    """evaluate.py:58-62 call: no masks, project=False, L2 normalise, BN eval stats, S != train length."""
    hc = O.HeadCfg(**dict(TINY))
    P = O.init_params(hc, seed=21)
    buf = O.init_bn_buffers(hc)
    g = torch.Generator().manual_seed(22)
    for k in list(buf):
        if k.endswith("running_mean"):
            buf[k] = 0.2 * torch.randn(buf[k].shape, generator=g)
        if k.endswith("running_var"):
            buf[k] = 0.5 + torch.rand(buf[k].shape, generator=g)
    S = 13                                            # != train_frames (8) -> linspace positions
    tokens = torch.randn(1, S, 16, hc.c_in, generator=g)
    cfg, head, proj, algo = R.build_reference_modules(hc, P)
    sd = {k[len("embed."):]: v for k, v in buf.items() if k.startswith("embed.")}
    head.load_state_dict(sd, strict=False)
    head.eval()
    with torch.no_grad():
        y = head(R.tokens_to_nchw(tokens), video_masks=None, cls_emb=None)
        y = torch.nn.functional.normalize(y, dim=-1)
        o, _ = O.model_forward(P, buf, tokens, None, hc, project=False, training=False)
    log(f"[tiny_eval] oracle-vs-reference fp32: out {float((o-y).abs().max()):.2e}")
    assert float((o - y).abs().max()) < 2e-5
    arrs = {"tokens": tokens.numpy(), "ref_out": y.numpy()}
    for k, v in P.items():
        arrs["param:" + k] = v.numpy()
    for k, v in buf.items():
        arrs["buf:" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "tiny_eval.npz"), **arrs)
    return dict(head_cfg={k: (list(v) if isinstance(v, tuple) else v) for k, v in hc.__dict__.items()}, S=S, Ptok=16)


def make_scl_cases():
    """SCL alone on random unit embeddings, larger T, both negative types, ragged masks (fp64 reference too)."""
    ref = R.load_reference()
    metas = {}
    for name, (Bv, T, D, neg, seed) in {"scl_T40_single": (3, 40, 32, "single_noself", 31),
                                        "scl_T20_batch": (4, 20, 128, "batch_noself", 33),
                                        "scl_T80_single_nopad": (2, 80, 64, "single_noself", 35)}.items():
        g = torch.Generator().manual_seed(seed)
        e = torch.nn.functional.normalize(torch.randn(Bv, 2, T, D, generator=g), dim=-1)
        _, seq_lens, steps, masks = O.synth_batch(Bv, T, 1, 1, seed=seed, with_padding=("nopad" not in name))
        cfg = R.reference_cfg("penn_mvf.yml", NEGATIVE_TYPE=neg)
        algo = ref.SCL(cfg)
        out = {}
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            ee = e.to(dt).clone().requires_grad_(True)
            loss = algo.compute_sequence_loss(ee, seq_lens, steps, masks.to(dt))["loss"]
            loss.backward()
            out[tag] = (loss.detach(), ee.grad.detach())
            eo = e.to(dt).clone().requires_grad_(True)
            lo = O.scl_loss_dense(eo, seq_lens, steps, masks.to(dt), negative_type=neg)
            lo.backward()
            log(f"[{name}/{tag}] oracle dense vs reference: loss rel {float((lo-loss).abs()/loss.abs()):.2e}  "
                f"dE rel {float((eo.grad-ee.grad).norm()/ee.grad.norm()):.2e}")
            assert float(abs(lo - loss) / abs(loss)) < (1e-5 if dt == torch.float32 else 1e-11)
        if neg == "single_noself":
            lp, dEp = O.scl_loss_pairs(e.numpy(), seq_lens.numpy(), steps.numpy(), masks.numpy())
            l64, d64 = out["f64"]
            log(f"[{name}] per-pair closed form vs reference fp64: loss rel {abs(lp-float(l64))/float(l64):.2e}  "
                f"dE rel {np.linalg.norm(dEp-d64.numpy())/np.linalg.norm(d64.numpy()):.2e}")
            assert abs(lp - float(l64)) / float(l64) < 1e-12
            assert np.linalg.norm(dEp - d64.numpy()) / np.linalg.norm(d64.numpy()) < 1e-12
            lq, dEq = O.scl_loss_pairs(e.numpy(), seq_lens.numpy(), steps.numpy(), masks.numpy(), quirk=False)
            log(f"[{name}] closed form WITHOUT the 1e-6 cross terms: loss rel {abs(lq-float(l64))/float(l64):.2e}  "
                f"dE rel {np.linalg.norm(dEq-d64.numpy())/np.linalg.norm(d64.numpy()):.2e}")
        np.savez_compressed(os.path.join(OUT, name + ".npz"), embs=e.numpy(), seq_lens=seq_lens.numpy(),
                            steps=steps.numpy(), masks=masks.numpy(),
                            ref_loss_f32=out["f32"][0].numpy(), ref_dE_f32=out["f32"][1].numpy(),
                            ref_loss_f64=out["f64"][0].numpy(), ref_dE_f64=out["f64"][1].numpy())
        metas[name] = dict(Bv=Bv, T=T, D=D, negative_type=neg, seed=seed)
    return metas


def make_cfg1_digest():
    """BASELINE configs[0] shape with the real penn_mvf.yml head sizes: digest only (params are seeded)."""
    hc = O.HeadCfg(c_in=1152, train_frames=20)
    Bv, T, Ptok, seed = 2, 20, 196, 1
    P = O.init_params(hc, seed=seed)
    buf = O.init_bn_buffers(hc)
    tokens, seq_lens, steps, masks = O.synth_batch(Bv, T, Ptok, hc.c_in, seed=seed)
    ref = run_reference(hc, P, buf, tokens, masks, seq_lens, steps, "single_noself", Bv, T)
    orc = run_oracle(hc, P, buf, tokens, masks, seq_lens, steps, "single_noself", Bv, T)
    keys = list(P.keys())
    gr, go = gcat(ref["grads"], keys), gcat(orc["grads"], keys)
    log(f"[penn_cfg1] oracle-vs-reference fp32: emb {float((orc['emb']-ref['emb']).abs().max()):.2e}  "
        f"loss rel {float(abs(orc['loss']-ref['loss'])/abs(ref['loss'])):.2e}  grad rel(L2) {float((gr-go).norm()/gr.norm()):.2e}")
    assert float((gr - go).norm() / gr.norm()) < 1e-5
    arrs = {"ref_emb": ref["emb"].numpy(), "ref_e": ref["e"].numpy(), "ref_loss": ref["loss"].numpy(),
            "masks": masks.numpy(), "seq_lens": seq_lens.numpy(), "steps": steps.numpy(),
            "tokens_head": tokens.reshape(-1)[:64].numpy(), "ref_attn_last": ref["attn_last"].numpy()[:, :, :8]}
    digest = {}
    for k in keys:
        g = ref["grads"][k].double().reshape(-1)
        digest[k] = dict(l2=float(g.norm()), sum=float(g.sum()), absmax=float(g.abs().max()), head=[float(v) for v in g[:6]])
        if g.numel() <= 1024:
            arrs["grad:" + k] = ref["grads"][k].numpy()
    np.savez_compressed(os.path.join(OUT, "penn_cfg1.npz"), **arrs)
    return dict(head_cfg={k: (list(v) if isinstance(v, tuple) else v) for k, v in hc.__dict__.items()},
                Bv=Bv, T=T, Ptok=Ptok, seed=seed, grad_digest=digest, grad_l2_total=float(gr.norm()))


def make_sampler_golden():
    """a13: run the reference's PennAction / FineGym sample_frames unbound, record int outputs."""
    import random
    import types
    ref = R.load_reference()
    out = {}
    if not hasattr(ref, "PennAction"):
        log("[sampler] reference sampler import failed:", getattr(ref, "sampler_import_error", "?"))
        return out
    cfg = R.reference_cfg("penn_mvf.yml")
    for variant, cls in (("penn_action", ref.PennAction), ("finegym", ref.FineGym),
                         ("pouring", ref.Pouring), ("pouring_fix", ref.Pouring)):
        recs = []
        for seed, seq_len, T in ((1, 100, 20), (2, 37, 20), (3, 12, 20), (4, 400, 80), (5, 250, 240), (6, 20, 20), (7, 1000, 32)):
            random.seed(seed)
            np.random.seed(seed)
            torch.manual_seed(seed)
            self = types.SimpleNamespace(cfg=cfg, num_contexts=1, sample_fix=(variant == "pouring_fix"))
            s0, c0, m0 = cls.sample_frames(self, seq_len, T)
            s1, c1, m1 = cls.sample_frames(self, seq_len, T, pre_steps=c0)
            recs.append(dict(seed=seed, seq_len=seq_len, T=T, steps0=s0.tolist(), chosen0=c0.tolist(), mask0=m0.tolist(),
                             steps1=s1.tolist(), chosen1=c1.tolist(), mask1=m1.tolist()))
            np.random.seed(seed)
            torch.manual_seed(seed)
            o0 = O.sample_frames_oracle(seq_len, T, None, variant=variant)
            o1 = O.sample_frames_oracle(seq_len, T, o0[1], variant=variant)
            assert o0[0].tolist() == s0.tolist() and o0[2].tolist() == m0.tolist(), (variant, seed)
            assert o1[0].tolist() == s1.tolist() and o1[2].tolist() == m1.tolist(), (variant, seed)
        out[variant] = recs
        log(f"[sampler/{variant}] {len(recs)} two-view draws recorded; oracle restatement bit-exact")
    return out


def main():
    # Single-threaded on purpose: with 8 intra-op threads torch 2.11's CPU backward of the reference modules
    # returned a d(emb) that is 4.6 % off its own fp64 result at the cfg1 shape (1 thread: 2.6e-6).  The
    # reference *algorithm* is what is pinned here, so the fixtures come from the deterministic 1-thread run.
    torch.set_num_threads(1)
    meta = {"cases": {}, "torch": torch.__version__, "numpy": np.__version__}
    for name, spec in CASES.items():
        meta["cases"][name] = make_case(name, spec)
    meta["cases"]["tiny_eval"] = make_eval_case()
    meta["scl"] = make_scl_cases()
    meta["penn_cfg1"] = make_cfg1_digest()
    samp = make_sampler_golden()
    with open(os.path.join(OUT, "sampler.json"), "w") as f:
        json.dump(samp, f)
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    with open(os.path.join(OUT, "PINNING.txt"), "w") as f:
        f.write("Oracle pinned against the reference's own PyTorch modules (see make_golden.py)\n")
        f.write("\n".join(LOG) + "\n")


if __name__ == "__main__":
    main()
