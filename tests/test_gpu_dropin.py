"""GPU: the drop-in boundary -- train.py's call sequence on the reference-shaped modules (build_model / get_algo /
algo.compute_loss(model, videos, seq_lens, chosen_steps, video_masks) / loss.backward / optimizer.step)."""
import pytest
import torch
import torch.nn as nn

from oracle import mvf_oracle as O
from tests import helpers as H
from tests.test_host_logic import DummyBackbone, small_cfg
from video_rep_learning_b200.algos import get_algo
from video_rep_learning_b200.models import build_model

pytestmark = pytest.mark.gpu


def _batch(Bv=2, T=8):
    g = torch.Generator().manual_seed(4)
    videos = torch.rand(Bv, 2, T, 3, 168, 168, generator=g)
    _, seq_lens, steps, masks = O.synth_batch(Bv, T, 1, 1, seed=21)
    return videos, seq_lens.view(-1)[::1].reshape(Bv, 2), steps, masks.view(Bv, 2, T)


def test_train_step_through_reference_call_sequence():
    torch.manual_seed(1)
    cfg = small_cfg(drop=0.0)
    model = build_model(cfg, backbone=DummyBackbone(c_out=48, patch=56)).cuda()
    algo = get_algo(cfg)
    optimizer = torch.optim.Adam([p for n, p in model.named_parameters() if "backbone" not in n], lr=1e-4, weight_decay=1e-5)
    videos, seq_lens, steps, masks = _batch()
    hooked = []
    model.embed.pooling.cross_att.attn_holder.register_forward_hook(lambda m, i, o: hooked.append(o))   # visualize_lstp.py:61
    model.train()
    optimizer.zero_grad()
    loss_dict = algo.compute_loss(model, videos.cuda(), seq_lens, steps, masks)
    loss = loss_dict["loss"]
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), cfg.OPTIMIZER.GRAD_CLIP)
    optimizer.step()
    assert torch.isfinite(loss) and len(hooked) == 1 and tuple(hooked[0].shape) == (8, 3, 9)
    assert torch.allclose(hooked[0].sum(-1), torch.ones(8, 3, device="cuda"), atol=1e-5)

    # same step on the oracle with the same parameters / tokens
    torch.manual_seed(1)
    model2 = build_model(small_cfg(drop=0.0), backbone=DummyBackbone(c_out=48, patch=56)).cuda()
    sd = {k: v.detach().cpu() for k, v in model2.state_dict().items() if not k.startswith("backbone")}
    with torch.no_grad():
        tokens = model2.backbone_tokens(videos.cuda().view(4, 8, 3, 168, 168)).cpu()
    hc = O.HeadCfg(c_in=48, n_entities=3, pool_channels=32, fc_channels=(64, 64), hidden=32, d_ff=64, n_heads=4,
                   n_layers=2, emb=16, proj=16, train_frames=8)
    P = {k: sd[k] for k in O.param_shapes(hc)}
    o = H.run_oracle(hc, P, None, tokens, masks.reshape(4, 1, 8), seq_lens, steps, dtype=torch.float64)
    assert abs(float(loss) - float(o["loss"])) / float(o["loss"]) < 1e-5
    model2.train()
    l2 = algo.compute_loss(model2, videos.cuda(), seq_lens, steps, masks)["loss"]
    l2.backward()
    got = {k: v.grad.cpu() for k, v in model2.named_parameters() if not k.startswith("backbone")}
    keys = list(P.keys())
    assert H.rel_l2(H.grad_vector(got, keys), H.grad_vector(o["grads"], keys)) < 2e-5
    # running statistics moved, eval path works and does not move them
    assert int(model2.embed.fc_layers[2].num_batches_tracked) == 1
    model2.eval()
    with torch.no_grad():
        emb = model2(videos.cuda().view(4, 8, 3, 168, 168)[:1, :5], 5)          # evaluate.py:58-62 style call
    assert tuple(emb.shape) == (1, 5, 16) and torch.allclose(emb.norm(dim=-1), torch.ones(1, 5, device="cuda"), atol=1e-5)
    assert int(model2.embed.fc_layers[2].num_batches_tracked) == 1


def test_separate_modules_equal_fused_node():
    """model.embed(x) -> model.ssl_projection(.) -> F.normalize, as transformer.py:222-228 chains them, equals the
    single fused node; also accepts the reference's NCHW hand-off layout."""
    torch.manual_seed(2)
    cfg = small_cfg(drop=0.0)
    model = build_model(cfg, backbone=DummyBackbone()).cuda().train()
    g = torch.Generator().manual_seed(5)
    tokens = torch.randn(4, 8, 9, 48, generator=g).cuda()
    masks = torch.ones(4, 1, 8).cuda(); masks[1, 0, 6:] = 0
    fused = model.forward_tokens(tokens, video_masks=masks, project=True)
    fused.square().sum().backward()          # any scalar
    gf = {k: v.grad.clone() for k, v in model.named_parameters() if v.grad is not None}
    model.zero_grad()
    for bn in (model.embed.fc_layers[2], model.embed.fc_layers[6], model.ssl_projection.net[1]):
        bn.reset_running_stats()
    nchw = tokens.transpose(2, 3).reshape(4, 8, 48, 3, 3).contiguous()
    emb = model.embed(nchw, video_masks=masks)
    sep = torch.nn.functional.normalize(model.ssl_projection(emb), dim=-1)
    assert float((sep - fused).abs().max()) < 1e-6
    w = torch.randn(4, 8, 16, generator=g).cuda()
    (sep * w).sum().backward()
    model.zero_grad()
    fused2 = model.forward_tokens(tokens, video_masks=masks, project=True)
    (fused2 * w).sum().backward()
    gs = {k: v.grad.clone() for k, v in model.named_parameters() if v.grad is not None}
    model.zero_grad()
    emb = model.embed(nchw, video_masks=masks)
    sep = torch.nn.functional.normalize(model.ssl_projection(emb), dim=-1)
    (sep * w).sum().backward()
    gmax = max(float(g.abs().max()) for g in gs.values())
    for k, v in model.named_parameters():
        if v.grad is not None and k in gs:
            # analytically-zero gradients (biases in front of BatchNorm) are rounding noise: absolute floor
            assert float((v.grad - gs[k]).abs().max()) <= 2e-5 * float(gs[k].abs().max()) + 1e-5 * gmax, k
