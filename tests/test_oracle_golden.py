"""CPU: the oracle restatement against the committed reference-generated golden vectors."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import mvf_oracle as O
from oracle import ref_shim as R
from tests import helpers as H

CASES = ["tiny_penn", "tiny_fg_avg", "tiny_max_nohot", "tiny_lin", "tiny_batch_noself", "tiny_e1"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    m, hc, z, P, G, B = H.load_case(name)
    tokens, masks = torch.from_numpy(z["tokens"]), torch.from_numpy(z["masks"])
    seq_lens, steps = torch.from_numpy(z["seq_lens"]), torch.from_numpy(z["steps"])
    o = H.run_oracle(hc, P, O.init_bn_buffers(hc), tokens, masks, seq_lens, steps, dtype=torch.float32,
                     negative_type=m["negative_type"])
    assert float((o["emb"] - torch.from_numpy(z["ref_emb"])).abs().max()) < 2e-5
    assert float((o["e"] - torch.from_numpy(z["ref_e"])).abs().max()) < 2e-5
    assert abs(float(o["loss"]) - float(z["ref_loss"])) / float(z["ref_loss"]) < 1e-5
    keys = list(P.keys())
    assert H.rel_l2(H.grad_vector(o["grads"], keys), H.grad_vector(G, keys)) < 1e-5
    for k, v in o["bufs"].items():
        assert float((v.double() - B[k].double()).abs().max()) < 1e-5, k
    # attention side channel: last video-view, [T, E, P] (mvformer.py:408-411)
    BV = tokens.shape[0]
    assert float((o["aux"]["attn"][BV - 1] - torch.from_numpy(z["ref_attn_last"])).abs().max()) < 1e-6


def test_oracle_eval_path_golden():
    m = H.meta()["cases"]["tiny_eval"]
    z = np.load(os.path.join(H.GOLDEN, "tiny_eval.npz"))
    kw = dict(m["head_cfg"])
    kw["fc_channels"] = tuple(kw["fc_channels"])
    hc = O.HeadCfg(**kw)
    P = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param:")}
    B = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("buf:")}
    out, _ = O.model_forward(P, B, torch.from_numpy(z["tokens"]), None, hc, project=False, training=False)
    assert float((out - torch.from_numpy(z["ref_out"])).abs().max()) < 2e-5


@pytest.mark.parametrize("name", ["scl_T40_single", "scl_T20_batch", "scl_T80_single_nopad"])
def test_scl_oracle_golden(name):
    z = np.load(os.path.join(H.GOLDEN, name + ".npz"))
    neg = H.meta()["scl"][name]["negative_type"]
    e = torch.from_numpy(z["embs"]).double().requires_grad_(True)
    loss = O.scl_loss_dense(e, torch.from_numpy(z["seq_lens"]), torch.from_numpy(z["steps"]),
                            torch.from_numpy(z["masks"]).double(), negative_type=neg)
    loss.backward()
    assert abs(float(loss) - float(z["ref_loss_f64"])) / float(z["ref_loss_f64"]) < 1e-12
    assert H.rel_l2(e.grad, torch.from_numpy(z["ref_dE_f64"])) < 1e-12
    if neg == "single_noself":
        lp, dE = O.scl_loss_pairs(z["embs"], z["seq_lens"], z["steps"], z["masks"])
        assert abs(lp - float(z["ref_loss_f64"])) / float(z["ref_loss_f64"]) < 1e-12
        assert np.linalg.norm(dE - z["ref_dE_f64"]) / np.linalg.norm(z["ref_dE_f64"]) < 1e-12


def test_sincos_table_properties():
    # even channels sin, odd channels cos, exponent uses the channel index itself (models/utils.py:113-126)
    tab = O.sincos_table(5, 8)
    assert np.allclose(tab[0, 0::2], 0.0) and np.allclose(tab[0, 1::2], 1.0)
    assert np.isclose(tab[3, 2], np.sin(3 / 10000 ** (2 / 8))) and np.isclose(tab[3, 5], np.cos(3 / 10000 ** (5 / 8)))
    t2 = O.sincos_table(7, 8, train_len=4)   # interpolated positions when S != train length
    assert np.isclose(t2[-1, 0], np.sin(3.0)) and np.isclose(t2[1, 0], np.sin(0.5))


def test_sampler_golden_bit_exact():
    """a13: both the oracle restatement and the product sampler reproduce the reference's draws bit for bit."""
    from video_rep_learning_b200.datasets import sample_frames
    with open(os.path.join(H.GOLDEN, "sampler.json")) as f:
        gold = json.load(f)
    assert set(gold) >= {"penn_action", "finegym", "pouring", "pouring_fix"}
    for variant, recs in gold.items():
        for r in recs:
            for fn in ("oracle", "product"):
                np.random.seed(r["seed"])
                torch.manual_seed(r["seed"])
                if fn == "oracle":
                    s0, c0, m0 = O.sample_frames_oracle(r["seq_len"], r["T"], None, variant=variant)
                    s1, c1, m1 = O.sample_frames_oracle(r["seq_len"], r["T"], c0, variant=variant)
                else:
                    s0, c0, m0 = sample_frames(r["seq_len"], r["T"], None, dataset=variant)
                    s1, c1, m1 = sample_frames(r["seq_len"], r["T"], c0, dataset=variant)
                assert s0.tolist() == r["steps0"] and c0.tolist() == r["chosen0"] and m0.tolist() == r["mask0"]
                assert s1.tolist() == r["steps1"] and c1.tolist() == r["chosen1"] and m1.tolist() == r["mask1"]
                assert s0.dtype == torch.int64 and m0.dtype == torch.float32


def test_sampler_edge_cases():
    from video_rep_learning_b200.datasets import sample_frames
    np.random.seed(0)
    torch.manual_seed(0)
    # seq_len < num_frames: padded with seq_len, clamped to seq_len-1, masked out
    s, c, m = sample_frames(5, 12)
    assert s.shape == (12,) and int(m.sum()) <= 5 and int(c.max()) <= 4
    assert all(int(m[i]) == 0 for i in range(12) if i >= 5)
    # offset_uniform branch
    s, c, m = sample_frames(50, 10, sampling_strategy="offset_uniform")
    assert torch.all(s[1:] >= s[:-1]) and float(m.sum()) == 10
    with pytest.raises(ValueError):
        sample_frames(10, 5, sampling_strategy="stride")


@pytest.mark.skipif(not R.available(), reason="reference tree not mounted (GPU box)")
def test_oracle_against_live_reference():
    """Build container only: re-run the reference modules and compare (the pinning itself)."""
    torch.set_num_threads(1)
    hc = O.HeadCfg(c_in=40, n_entities=2, pool_channels=24, fc_channels=(48, 48), hidden=32, d_ff=64, n_heads=4,
                   n_layers=1, emb=16, proj=16, train_frames=6, final="avg")
    P = O.init_params(hc, seed=77)
    tokens, seq_lens, steps, masks = O.synth_batch(2, 6, 9, hc.c_in, seed=78)
    cfg, head, proj, algo = R.build_reference_modules(hc, P)
    head.train(); proj.train()
    emb = head(R.tokens_to_nchw(tokens), video_masks=masks, cls_emb=None)
    e = torch.nn.functional.normalize(proj(emb), dim=-1)
    loss = algo.compute_sequence_loss(e.view(2, 2, 6, -1), seq_lens, steps, masks)["loss"]
    loss.backward()
    g = {"embed." + k: v.grad for k, v in head.named_parameters()}
    g.update({"ssl_projection." + k: v.grad for k, v in proj.named_parameters()})
    o = H.run_oracle(hc, P, O.init_bn_buffers(hc), tokens, masks, seq_lens, steps, dtype=torch.float32)
    assert float((o["emb"] - emb.detach()).abs().max()) < 2e-5
    assert abs(float(o["loss"]) - float(loss)) / float(loss) < 1e-5
    keys = list(P.keys())
    assert H.rel_l2(H.grad_vector(o["grads"], keys), H.grad_vector(g, keys)) < 1e-5


@pytest.mark.parametrize("name", ["f2_fwb", "f2_fwb_e5_avg"])
def test_oracle_matches_reference_golden_optional_branches(name):
    """SURVEY.md section 8f-2 branches: the oracle restatement against reference-generated fixtures (make_golden_f2.py)."""
    m, hc, z, P, G, B = H.load_case(name, "meta_f2.json")
    tokens, masks = torch.from_numpy(z["tokens"]), torch.from_numpy(z["masks"])
    seq_lens, steps = torch.from_numpy(z["seq_lens"]), torch.from_numpy(z["steps"])
    cls_emb = torch.from_numpy(z["cls_emb"]) if "cls_emb" in z.files else None
    o = H.run_oracle(hc, P, None, tokens, masks, seq_lens, steps, dtype=torch.float64, cls_emb=cls_emb)
    keys = list(P.keys())
    assert H.rel_l2(o["e"], torch.from_numpy(z["ref_e"])) < 1e-5
    assert abs(float(o["loss"]) - float(z["ref_loss"])) / float(z["ref_loss"]) < 1e-5
    assert H.rel_l2(H.grad_vector(o["grads"], keys), H.grad_vector(G, keys)) < 1e-5
