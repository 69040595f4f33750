"""GPU: CUDA-graph replay of the training step (video_rep_learning_b200/graph.py) against the eager step it was
captured from -- same kernels, same buffers, so the results agree to the re-association noise of the split-K
reductions; dropout masks must change from replay to replay (device-side seed counter) and be reproducible."""
import pytest
import torch

from oracle import mvf_oracle as O
from tests.test_host_logic import DummyBackbone, small_cfg
from video_rep_learning_b200 import engine
from video_rep_learning_b200.algos import get_algo
from video_rep_learning_b200.graph import GraphedTrainStep
from video_rep_learning_b200.models import build_model

pytestmark = pytest.mark.gpu

Bv, T, P, C = 4, 8, 9, 48


def _setup(drop, dtype):
    torch.manual_seed(3)
    cfg = small_cfg(drop=drop)
    model = build_model(cfg, backbone=DummyBackbone()).cuda().train()
    algo = get_algo(cfg)
    g = torch.Generator().manual_seed(7)
    tokens = torch.randn(2 * Bv, T, P, C, generator=g).to(dtype).cuda()
    _, seq_lens, steps, masks = O.synth_batch(Bv, T, 1, 1, seed=5)
    return cfg, model, algo, tokens, seq_lens.cuda(), steps.cuda(), masks.cuda()


def _eager(model, algo, tokens, seq_lens, steps, masks):
    for p in model.parameters():
        p.grad = None
    embs = model.forward_tokens(tokens, video_masks=masks, project=True)
    loss = algo.compute_sequence_loss(embs.view(Bv, 2, T, -1), seq_lens, steps, masks)["loss"]
    loss.backward()
    grads = torch.cat([p.grad.flatten() for n, p in model.named_parameters() if "backbone" not in n]).clone()
    return float(loss), grads


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_replay_equals_eager_without_dropout(dtype):
    cfg, model, algo, tokens, seq_lens, steps, masks = _setup(0.0, dtype)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    loss_e, grads_e = _eager(model, algo, tokens, seq_lens, steps, masks)
    model.load_state_dict(sd0)          # BatchNorm running statistics back to their start
    gs = GraphedTrainStep(model, algo, Bv, T, P, C, dtype=dtype)
    gs.capture()
    assert gs.launches_per_step > 20
    tol = 1e-6
    for rep in range(3):
        loss_g = gs(tokens, seq_lens, steps, masks)
        grads_g = torch.cat([p.grad.flatten() for p in gs.params])
        assert abs(float(loss_g) - loss_e) <= tol * abs(loss_e)
        assert float((grads_g - grads_e).norm()) <= 10 * tol * float(grads_e.norm())
    # new inputs through the static buffers: the replay follows them
    tokens2 = tokens.flip(0).contiguous()
    loss_e2, grads_e2 = _eager(model, algo, tokens2, seq_lens, steps, masks)
    loss_g2 = gs(tokens2)
    assert abs(loss_e2 - loss_e) > 1e-4 * abs(loss_e)
    assert abs(float(loss_g2) - loss_e2) <= tol * abs(loss_e2)
    gs.release()
    assert model.embed.seed_dev is None


def test_replay_draws_fresh_reproducible_dropout_masks(monkeypatch):
    cfg, model, algo, tokens, seq_lens, steps, masks = _setup(0.3, torch.float32)
    monkeypatch.setattr(engine, "new_seed", lambda: 1234)
    gs = GraphedTrainStep(model, algo, Bv, T, P, C, dtype=torch.float32, warmup=1)
    gs.capture()
    c0 = int(gs.seed_dev.item())
    l1 = float(gs(tokens, seq_lens, steps, masks))
    g1 = torch.cat([p.grad.flatten() for p in gs.params]).clone()
    l2 = float(gs())
    assert int(gs.seed_dev.item()) == c0 + 2
    assert abs(l1 - l2) > 1e-6 * abs(l1)                      # another mask
    # an eager step with the device counter set to the value replay #1 saw reproduces replay #1
    gs.seed_dev.fill_(c0 + 1)
    le, ge = _eager(model, algo, tokens, seq_lens, steps, masks)
    assert abs(le - l1) <= 1e-6 * abs(l1)
    assert float((ge - g1).norm()) <= 1e-6 * float(g1.norm())
    # and the host half of the seed still works on its own (no counter): seed s + counter c == seed (s + c)
    model.embed.seed_dev = None
    monkeypatch.setattr(engine, "new_seed", lambda: 1234 + c0 + 1)
    le2, _ = _eager(model, algo, tokens, seq_lens, steps, masks)
    assert abs(le2 - l1) <= 1e-6 * abs(l1)
    gs.release()


def test_optimizer_sees_replayed_gradients():
    cfg, model, algo, tokens, seq_lens, steps, masks = _setup(0.0, torch.float32)
    gs = GraphedTrainStep(model, algo, Bv, T, P, C, dtype=torch.float32).capture()
    opt = torch.optim.SGD(gs.params, lr=0.05)
    losses = []
    for _ in range(4):
        opt.zero_grad(set_to_none=True)
        losses.append(float(gs(tokens, seq_lens, steps, masks)))
        assert all(p.grad is not None for p in gs.params)
        opt.step()
    assert losses[-1] < losses[0]                              # parameters are read through pointers: updates are seen
    gs.release()
