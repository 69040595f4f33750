"""CPU, world_size 2, gloo: the multi-GPU protocol of SURVEY.md section 8e on the host.

The kernels need a GPU, so what is exercised here is everything around them that makes N > 1 correct: video
sharding, the BatchNorm statistics exchange ([sum x, sum x^2] forward, [sum dy, sum dy*xhat] backward, float64,
n_global = rows * world, LOCAL d(gamma)/d(beta)), the single flat-gradient all-reduce with the 1/world scale,
and the per-rank loss normalisation.  Each rank runs the oracle on its shard with its BatchNorm replaced by an
implementation built from video_rep_learning_b200.parallel -- the same helpers engine.py calls between kernel
phases -- and the result must equal the single-process oracle with the reference's SyncBatchNorm + DDP
semantics: BatchNorm over the concatenated batch, loss = mean over ranks of the per-rank SCL means.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mvf_oracle as O
from video_rep_learning_b200 import parallel

HC = O.HeadCfg(c_in=24, n_entities=2, pool_channels=16, fc_channels=(32, 32), hidden=16, d_ff=32, n_heads=2, n_layers=1,
               emb=8, proj=8, train_frames=6)
BV_GLOBAL, T, PTOK = 4, 6, 4     # 4 videos -> 2 per rank


class _SyncBN(torch.autograd.Function):
    """What csrc/elementwise.cu does in three launches (bn_stats -> [all-reduce] -> bn_finalize/bn_apply and
    bn_bwd_stats -> [all-reduce] -> bn_bwd_apply), written with the package's collectives."""

    @staticmethod
    def forward(ctx, x, w, b, eps, world):
        stats = torch.cat([x.double().sum(0), (x.double() ** 2).sum(0)])
        parallel.sync_stats_(stats)
        n = parallel.bn_global_rows(x.shape[0], world)
        C = x.shape[1]
        mean = stats[:C] / n
        var = (stats[C:] / n - mean * mean).clamp_min(0)
        invstd = 1.0 / torch.sqrt(var + eps)
        xh = (x.double() - mean) * invstd
        ctx.save_for_backward(xh, invstd, w)
        ctx.n = n
        ctx.mark_non_differentiable(mean, var)
        return (xh * w.double() + b.double()).to(x.dtype), mean.to(x.dtype), var.to(x.dtype)

    @staticmethod
    def backward(ctx, dy, _dm, _dv):
        xh, invstd, w = ctx.saved_tensors
        dy = dy.double()
        local = torch.cat([dy.sum(0), (dy * xh).sum(0)])
        C = dy.shape[1]
        dgamma, dbeta = local[C:].clone(), local[:C].clone()       # local sums, as torch's SyncBatchNorm
        parallel.sync_stats_(local)
        dx = w.double() * invstd * (dy - local[:C] / ctx.n - xh * local[C:] / ctx.n)
        return dx.to(w.dtype), dgamma.to(w.dtype), dbeta.to(w.dtype), None, None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    P = O.init_params(HC, seed=5, dtype=torch.float64)
    tokens, seq_lens, steps, masks = O.synth_batch(BV_GLOBAL, T, PTOK, HC.c_in, seed=6)
    return P, tokens.double(), seq_lens, steps, masks.double()


def _rank_loss(P, tokens, masks, seq_lens, steps):
    emb, _ = O.head_forward(P, None, tokens, masks, HC, True)
    e, _ = O.proj_forward(P, None, emb, HC, True)
    Bv = tokens.shape[0] // 2
    return O.scl_loss_dense(e.view(Bv, 2, T, -1), seq_lens, steps, masks)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        P, tokens, seq_lens, steps, masks = _inputs()
        v0, v1 = parallel.shard_videos(BV_GLOBAL, rank, world)
        Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
        O.BN_TRAIN_HOOK = lambda x, w, b, eps: _SyncBN.apply(x, w, b, eps, world)
        loss = _rank_loss(Pr, tokens[2 * v0:2 * v1], masks[2 * v0:2 * v1], seq_lens[v0:v1], steps[v0:v1])
        loss.backward()
        flat = torch.cat([Pr[k].grad.reshape(-1) for k in P])          # the "gpack" of this rank
        # the engine sums the buffer in two pieces (the chain's gradients beside the pooling backward, the pooling gradients at
        # the end of the step, engine.py); off the symmetric-memory path both pieces are plain all-reduces
        assert parallel.flat_grad_buffer(8, torch.device("cpu")).tolist() == [0.0] * 8      # gloo: an ordinary zeroed buffer
        assert parallel.PeerFlatGrads.find(flat) is None
        split = (flat.numel() // 3 + 3) // 4 * 4
        scale = parallel.finish_flat_grads_(flat[split:], channel=1)
        assert parallel.finish_flat_grads_(flat[:split], channel=0) == scale
        flat *= scale
        # descriptor world sizes (engine.py): BatchNorm divides its statistics by local_rows * bn_world, so bn_world must be 1
        # when the statistics are NOT exchanged (sync_bn=False), whatever the size of the group; the gradient all-reduce is
        # gated by the group size alone
        from video_rep_learning_b200 import engine
        worlds = (engine._world(engine.RunOptions()), engine._bn_world(engine.RunOptions()),
                  engine._world(engine.RunOptions(sync_bn=False)), engine._bn_world(engine.RunOptions(sync_bn=False)))
        torch.save({"loss": loss.detach(), "flat": flat, "worlds": worlds}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        O.BN_TRAIN_HOOK = None
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_protocol_matches_syncbn_ddp_semantics(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(str(tmp_path), "rank0.pt"))
    r1 = torch.load(os.path.join(str(tmp_path), "rank1.pt"))
    assert torch.equal(r0["flat"], r1["flat"])                     # every rank ends with identical gradients
    assert r0["worlds"] == (2, 2, 2, 1) and r1["worlds"] == (2, 2, 2, 1)

    # single-process ground truth: BatchNorm over ALL rows, loss = mean of per-rank means
    P, tokens, seq_lens, steps, masks = _inputs()
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    emb, _ = O.head_forward(Pr, None, tokens, masks, HC, True)
    e, _ = O.proj_forward(Pr, None, emb, HC, True)
    e = e.view(BV_GLOBAL, 2, T, -1)
    losses = []
    for r in range(world):
        v0, v1 = parallel.shard_videos(BV_GLOBAL, r, world)
        losses.append(O.scl_loss_dense(e[v0:v1], seq_lens[v0:v1], steps[v0:v1], masks[2 * v0:2 * v1]))
    total = sum(losses) / world
    total.backward()
    want = torch.cat([Pr[k].grad.reshape(-1) for k in P])
    assert abs(float(r0["loss"]) - float(losses[0])) < 1e-12 and abs(float(r1["loss"]) - float(losses[1])) < 1e-12
    assert float((r0["flat"] - want).norm() / want.norm()) < 1e-10

    # and it is NOT what per-rank (unsynchronised) BatchNorm would give -- the exchange matters
    Pq = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    lq = sum(_rank_loss(Pq, tokens[4 * r:4 * r + 4], masks[4 * r:4 * r + 4], seq_lens[2 * r:2 * r + 2], steps[2 * r:2 * r + 2])
             for r in range(world)) / world
    lq.backward()
    other = torch.cat([Pq[k].grad.reshape(-1) for k in P])
    assert float((other - want).norm() / want.norm()) > 1e-3


def test_shard_videos_partition():
    assert [parallel.shard_videos(256, r, 8) for r in (0, 3, 7)] == [(0, 32), (96, 128), (224, 256)]
    with pytest.raises(ValueError):
        parallel.shard_videos(10, 0, 4)
    assert parallel.world_size() == 1 and parallel.rank() == 0
    t = torch.ones(4, dtype=torch.float64)
    assert parallel.finish_flat_grads_(t) == 1.0 and torch.equal(parallel.sync_stats_(t), torch.ones(4, dtype=torch.float64))
