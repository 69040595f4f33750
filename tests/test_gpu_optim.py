"""GPU: the fused optimizer tail (csrc/optim.cu, optim.FusedAdam) against torch.nn.utils.clip_grad_norm_ +
torch.optim.Adam / AdamW on the same tensors (train.py:124-133,151-155; utils/optimizer.py:60-73)."""
import copy

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
    # the step count travels with the checkpoint (bias correction after a resume) ...
    assert float(sd["state"][0]["step"]) == 1.0
    c = [torch.nn.Parameter(p.detach().clone()) for p in a]
    oc = FusedAdam(c, lr=1e-3, max_grad_norm=5.0)
    oc.load_state_dict(copy.deepcopy(sd))      # as after torch.save / torch.load: load_state_dict itself aliases same-device tensors
    assert oc.step_count == 1
    for p, q, g in zip(a, c, gs[1]):
        p.grad = g.clone()
        q.grad = g.clone()
    oa.step()
    oc.step()
    assert oc.step_count == 2
    for p, q in zip(a, c):
        assert float((p - q).abs().max()) <= 1e-7 * max(1.0, float(p.abs().max()))
    # ... and a torch.optim.Adam checkpoint (per-parameter 'step' tensors) is accepted
    ref_p = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    ref = torch.optim.Adam(ref_p, lr=1e-3)
    for _ in range(3):
        for p, g in zip(ref_p, gs[2]):
            p.grad = g.clone()
        ref.step()
    od = FusedAdam([torch.nn.Parameter(p.detach().clone()) for p in ref_p], lr=1e-3)
    od.load_state_dict(ref.state_dict())
    assert od.step_count == 3


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
    gs = GraphedTrainStep(model, algo, Bv, T, P, Cc, dtype=torch.float32, optimizer=opt, warmup=2)
    gs.capture()
    # capturing (eager warm-up steps on whatever is in the static buffers) leaves no trace on the training state:
    # parameters, BatchNorm running statistics / counters, Adam moments and the step counter are as before
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd0[k]), k
    assert opt.step_count == 0
    assert all(float(st["exp_avg"].abs().max()) == 0.0 for st in opt.state.values())
    losses = [float(gs(tokens, seq_lens.cuda(), steps.cuda(), masks.cuda())) for _ in range(6)]
    assert losses[-1] < losses[0]
    assert opt.step_count == 6 and float(opt.grad_norm) > 0
    # a scheduler changing the learning rate between replays is honoured: lr = 0 freezes the parameters
    for grp in opt.param_groups:
        grp["lr"] = 0.0
    before = [p.detach().clone() for p in gs.params]
    gs()
    torch.cuda.synchronize()
    assert all(torch.equal(p.detach(), b) for p, b in zip(gs.params, before))
    for grp in opt.param_groups:
        grp["lr"] = 2e-3
    gs()
    torch.cuda.synchronize()
    assert any(not torch.equal(p.detach(), b) for p, b in zip(gs.params, before))
    gs.release()
