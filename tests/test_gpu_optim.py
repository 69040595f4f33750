"""GPU: the fused optimizer tail (csrc/optim.cu, optim.FusedAdam) against torch.nn.utils.clip_grad_norm_ +
torch.optim.Adam / AdamW on the same tensors (train.py:124-133,151-155; utils/optimizer.py:60-73)."""
import pytest
import torch

from tests.test_host_logic import DummyBackbone, small_cfg
from video_rep_learning_b200.optim import FusedAdam

pytestmark = pytest.mark.gpu


def _tensors(seed, scale):
    g = torch.Generator(device="cuda").manual_seed(seed)
    shapes = [(37,), (64, 48), (3, 5, 7), (1,), (256, 300), (1024,)]
    ps = [torch.randn(s, generator=g, device="cuda") for s in shapes]
    gs = [[torch.randn(s, generator=g, device="cuda") * scale for s in shapes] for _ in range(4)]
    return ps, gs


@pytest.mark.parametrize("adamw", [False, True])
@pytest.mark.parametrize("scale,max_norm", [(1.0, 10.0), (0.001, 10.0), (1.0, 0.0)])
def test_fused_adam_matches_torch(adamw, scale, max_norm):
    ps, gs = _tensors(3, scale)
    ref_p = [torch.nn.Parameter(p.clone()) for p in ps]
    our_p = [torch.nn.Parameter(p.clone()) for p in ps]
    cls = torch.optim.AdamW if adamw else torch.optim.Adam
    ref = cls(ref_p, lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    ours = FusedAdam(our_p, lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=max_norm, adamw=adamw)
    for it in range(4):
        if it == 2:                                   # a scheduler changing the learning rate between steps
            for o in (ref, ours):
                for grp in o.param_groups:
                    grp["lr"] = 1e-3
        for p, q, g in zip(ref_p, our_p, gs[it]):
            p.grad = g.clone()
            q.grad = g.clone()
        norm_ref = None
        if max_norm > 0:
            norm_ref = torch.nn.utils.clip_grad_norm_(ref_p, max_norm)
        ref.step()
        ours.step()
        if norm_ref is not None:
            assert abs(float(ours.grad_norm) - float(norm_ref)) <= 1e-6 * float(norm_ref)
        for p, q in zip(ref_p, our_p):
            assert float((p - q).abs().max()) <= 2e-6 * max(1.0, float(p.abs().max())), it
    assert ours.step_count == 4
    for p, q in zip(ref_p, our_p):
        assert torch.allclose(ref.state[p]["exp_avg"], ours.state[q]["exp_avg"], rtol=1e-5, atol=1e-8)
        assert torch.allclose(ref.state[p]["exp_avg_sq"], ours.state[q]["exp_avg_sq"], rtol=1e-5, atol=1e-10)


def test_unscale_factor_and_state_dict_round_trip():
    ps, gs = _tensors(5, 1.0)
    a = [torch.nn.Parameter(p.clone()) for p in ps]
    b = [torch.nn.Parameter(p.clone()) for p in ps]
    oa = FusedAdam(a, lr=1e-3, max_grad_norm=5.0)
    ob = FusedAdam(b, lr=1e-3, max_grad_norm=5.0)
    for p, q, g in zip(a, b, gs[0]):
        p.grad = g.clone()
        q.grad = g.clone() * 1024.0                   # scaled loss (GradScaler)
    oa.step()
    ob.step(inv_scale=1.0 / 1024.0)
    for p, q in zip(a, b):
        assert float((p - q).abs().max()) <= 1e-6 * max(1.0, float(p.abs().max()))
    sd = oa.state_dict()
    assert len(sd["state"]) == len(ps) and "exp_avg" in sd["state"][0]


def test_optimizer_tail_inside_the_step_graph():
    """GraphedTrainStep(optimizer=FusedAdam): forward, SCL, backward and the parameter update in ONE graph launch."""
    from oracle import mvf_oracle as O
    from video_rep_learning_b200.algos import get_algo
    from video_rep_learning_b200.graph import GraphedTrainStep
    from video_rep_learning_b200.models import build_model
    from video_rep_learning_b200.optim import construct_optimizer
    Bv, T, P, Cc = 4, 8, 9, 48
    torch.manual_seed(3)
    cfg = small_cfg(drop=0.0)
    model = build_model(cfg, backbone=DummyBackbone()).cuda().train()
    algo = get_algo(cfg)
    g = torch.Generator().manual_seed(7)
    tokens = torch.randn(2 * Bv, T, P, Cc, generator=g).cuda()
    _, seq_lens, steps, masks = O.synth_batch(Bv, T, 1, 1, seed=5)
    opt = construct_optimizer(model, cfg)
    for grp in opt.param_groups:
        grp["lr"] = 2e-3
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    gs = GraphedTrainStep(model, algo, Bv, T, P, Cc, dtype=torch.float32, optimizer=opt, warmup=1)
    gs.capture()
    model.load_state_dict(sd0)                        # the warm-up steps of capture() already moved the parameters
    losses = [float(gs(tokens, seq_lens.cuda(), steps.cuda(), masks.cuda())) for _ in range(6)]
    assert losses[-1] < losses[0]
    assert opt.step_count >= 6 and float(opt.grad_norm) > 0
    gs.release()
