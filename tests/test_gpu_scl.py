"""GPU: fused SCL forward+gradient (csrc/scl.cu, csrc/scl_mma.cu) through the algos/ interface against the reference goldens,
plus size-independent properties at BASELINE-scale batches."""
import os

import numpy as np
import pytest
import torch

from oracle import mvf_oracle as O
from tests import helpers as H
from video_rep_learning_b200 import engine
from video_rep_learning_b200.algos import SCL
from video_rep_learning_b200.config import mvf_cfg

pytestmark = pytest.mark.gpu


def _run(embs, seq_lens, steps, masks, neg="single_noself", quirk=True):
    e = embs.cuda().float().requires_grad_(True)
    loss = engine.SCLFn.apply(e, seq_lens.cuda(), steps.cuda(), masks.cuda(), 0.1, 10.0, neg, quirk)
    loss.backward()
    torch.cuda.synchronize()
    return float(loss), e.grad.cpu()


@pytest.mark.parametrize("name", ["scl_T40_single", "scl_T20_batch", "scl_T80_single_nopad"])
def test_scl_matches_reference_golden(name):
    z = np.load(os.path.join(H.GOLDEN, name + ".npz"))
    neg = H.meta()["scl"][name]["negative_type"]
    loss, dE = _run(torch.from_numpy(z["embs"]), torch.from_numpy(z["seq_lens"]), torch.from_numpy(z["steps"]),
                    torch.from_numpy(z["masks"]), neg)
    # tolerance stated by the north star: 1e-5 relative in fp32 (the reference's own fp32-vs-fp64 gap is ~1e-7)
    le, ge = abs(loss - float(z["ref_loss_f64"])) / float(z["ref_loss_f64"]), H.rel_l2(dE, torch.from_numpy(z["ref_dE_f64"]))
    print(f"{name}: loss {le:.2e} gradient {ge:.2e} against the reference in fp64")
    assert le < 1e-5 and ge < 1e-5
    assert abs(loss - float(z["ref_loss_f32"])) / float(z["ref_loss_f32"]) < 1e-5


def test_scl_algos_interface_and_masked_gradient_quirk():
    """compute_sequence_loss keeps the reference signature/return type; masked frames receive the (tiny) 1e-6
    gradient of scl.py:80 when quirk is on and exactly zero when it is off."""
    z = np.load(os.path.join(H.GOLDEN, "scl_T40_single.npz"))
    algo = SCL(mvf_cfg(num_frames=40))
    e = torch.from_numpy(z["embs"]).cuda().requires_grad_(True)
    out = algo.compute_sequence_loss(e, torch.from_numpy(z["seq_lens"]).cuda(), torch.from_numpy(z["steps"]).cuda(),
                                     torch.from_numpy(z["masks"]).cuda())
    assert set(out.keys()) == {"loss"} and out["loss"].dim() == 0
    out["loss"].backward()
    m = torch.from_numpy(z["masks"]).reshape(-1) == 0
    g = e.grad.cpu().reshape(-1, e.shape[-1])
    assert m.any() and float(g[m].abs().max()) > 0 and float(g[m].abs().max()) < 1e-4 * float(g.abs().max())
    _, g0 = _run(torch.from_numpy(z["embs"]), torch.from_numpy(z["seq_lens"]), torch.from_numpy(z["steps"]),
                 torch.from_numpy(z["masks"]), quirk=False)
    assert float(g0.reshape(-1, e.shape[-1])[m].abs().max()) == 0.0
    lq, dq = O.scl_loss_pairs(z["embs"], z["seq_lens"], z["steps"], z["masks"], quirk=False)
    assert H.rel_l2(g0, torch.from_numpy(dq)) < 1e-5


@pytest.mark.parametrize("Bv,T,D", [(1, 1, 8), (2, 3, 16), (1, 33, 24), (2, 240, 128), (3, 80, 256)])
def test_scl_edge_shapes_against_closed_form(Bv, T, D):
    g = torch.Generator().manual_seed(Bv * 100 + T)
    e = torch.nn.functional.normalize(torch.randn(Bv, 2, T, D, generator=g), dim=-1)
    _, seq_lens, steps, masks = O.synth_batch(Bv, T, 1, 1, seed=T + D)
    loss, dE = _run(e, seq_lens, steps, masks)
    lp, dEp = O.scl_loss_pairs(e.numpy(), seq_lens.numpy(), steps.numpy(), masks.numpy())
    assert abs(loss - lp) <= 1e-5 * abs(lp) + 1e-7      # fp32 log(1 + 1e-6) itself carries ~5e-8 of rounding
    # 1e-5 relative, plus the fp32 floor of the cancelling O(1) terms p*g - y*r scaled by 1/tau = 10 (10 * eps per element):
    # in the degenerate T = 1 case the analytic gradient is exactly zero and only that rounding residue is left
    floor = 10 * 6e-8 * float(np.sqrt(dEp.size))
    assert float((dE.double() - torch.from_numpy(dEp)).norm()) <= 1e-5 * float(np.linalg.norm(dEp)) + floor


@pytest.mark.parametrize("Bv,T,D,neg", [
    (160, 40, 64, "single_noself"),     # one CTA per pair holding both views, S recomputed per pass (needs a batch >= the SM count)
    (150, 96, 32, "single_noself"),     # the same at its largest T (12 warps)
    (1, 256, 256, "single_noself"),     # largest shape: cluster of 8, partner view staged per pass (does not fit shared memory)
    (2, 32, 128, "single_noself"),      # exactly one tile
    (3, 64, 36, "single_noself"),       # channels padded to 64 inside the kernel
    (5, 20, 128, "batch_noself"),       # batch negatives through the cross passes, T <= 32 kernel
    (3, 48, 64, "batch_noself"),        # and through the cluster kernel
])
def test_scl_kernel_variants_against_dense_oracle(Bv, T, D, neg):
    """Every shape class of scl_mma.cu (kernel variant chosen by T, D and the batch size) against the dense N x N oracle in
    fp64 on a subsample of the videos' rows (loss exactly, gradient on all rows)."""
    g = torch.Generator().manual_seed(Bv * 7 + T + D)
    e = torch.nn.functional.normalize(torch.randn(Bv, 2, T, D, generator=g), dim=-1)
    _, seq_lens, steps, masks = O.synth_batch(Bv, T, 1, 1, seed=Bv + T)
    loss, dE = _run(e, seq_lens, steps, masks, neg=neg)
    if neg == "single_noself":
        lp, dEp = O.scl_loss_pairs(e.numpy(), seq_lens.numpy(), steps.numpy(), masks.numpy())
        ref_l, ref_g = lp, torch.from_numpy(dEp)
    else:
        e64 = e.double().requires_grad_(True)
        ref = O.scl_loss_dense(e64, seq_lens, steps, masks.double(), negative_type=neg)
        ref.backward()
        ref_l, ref_g = float(ref), e64.grad
    le, ge = abs(loss - ref_l) / abs(ref_l), H.rel_l2(dE, ref_g)
    print(f"SCL {Bv} x {T} x {D} {neg}: loss {le:.2e} gradient {ge:.2e}")
    assert le < 1e-5 and ge < 1e-5


def test_scl_all_frames_valid_and_fully_masked_video():
    Bv, T, D = 3, 20, 128
    g = torch.Generator().manual_seed(9)
    e = torch.nn.functional.normalize(torch.randn(Bv, 2, T, D, generator=g), dim=-1)
    steps = torch.sort(torch.randint(0, 100, (Bv, 2, T), generator=g), dim=-1)[0]
    seq_lens = torch.full((Bv, 2), 100)
    masks = torch.ones(Bv * 2, 1, T)
    loss, dE = _run(e, seq_lens, steps, masks)
    lp, dEp = O.scl_loss_pairs(e.numpy(), seq_lens.numpy(), steps.numpy(), masks.numpy())
    assert abs(loss - lp) < 1e-5 * lp and H.rel_l2(dE, torch.from_numpy(dEp)) < 1e-5
    masks[2:4] = 0           # video 1 entirely padded: contributes nothing, receives only quirk gradients
    loss2, dE2 = _run(e, seq_lens, steps, masks)
    lp2, dEp2 = O.scl_loss_pairs(e.numpy(), seq_lens.numpy(), steps.numpy(), masks.numpy())
    assert abs(loss2 - lp2) < 1e-5 * lp2 and H.rel_l2(dE2, torch.from_numpy(dEp2)) < 1e-5
    assert np.isfinite(loss2)


def test_scl_properties_at_scale():
    """BASELINE cfg2/cfg3-sized batches (and beyond), where the dense oracle is too slow: size-independent checks.
    (1) additivity: without padded frames pairs are independent, so loss*M adds over disjoint sets of videos;
    (2) permutation equivariance over videos; (3) the two views are interchangeable; (4) dE is tangent-free of
    nothing in particular but must be finite and reproducible run to run up to atomics on the scalar loss."""
    Bv, T, D = 256, 20, 128
    g = torch.Generator().manual_seed(11)
    e = torch.nn.functional.normalize(torch.randn(Bv, 2, T, D, generator=g), dim=-1)
    steps = torch.sort(torch.randint(0, 200, (Bv, 2, T), generator=g), dim=-1)[0]
    seq_lens = torch.full((Bv, 2), 200)
    masks = torch.ones(Bv * 2, 1, T)
    la, ga = _run(e, seq_lens, steps, masks)
    h = Bv // 2
    l1, g1 = _run(e[:h], seq_lens[:h], steps[:h], masks[:2 * h])
    l2, g2 = _run(e[h:], seq_lens[h:], steps[h:], masks[2 * h:])
    assert abs(la - 0.5 * (l1 + l2)) < 1e-5 * la
    assert H.rel_l2(ga, 0.5 * torch.cat([g1, g2])) < 2e-6
    perm = torch.randperm(Bv, generator=g)
    lp_, gp = _run(e[perm], seq_lens[perm], steps[perm], masks.view(Bv, 2, 1, T)[perm].reshape(Bv * 2, 1, T))
    assert abs(lp_ - la) < 1e-5 * la and H.rel_l2(gp, ga[perm]) < 1e-6
    ls, gs = _run(e.flip(1), seq_lens.flip(1), steps.flip(1), masks)
    assert abs(ls - la) < 1e-5 * la and H.rel_l2(gs, ga.flip(1)) < 1e-6
    assert torch.isfinite(ga).all()
    # a subsample of pairs against the float64 closed form
    idx = [0, 17, 255]
    lsub, gsub = O.scl_loss_pairs(e[idx].numpy(), seq_lens[idx].numpy(), steps[idx].numpy(), np.ones((3, 2, T)))
    assert H.rel_l2(ga[idx] * (Bv / 3.0), torch.from_numpy(gsub)) < 1e-5
