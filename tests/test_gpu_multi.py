"""GPU x2 (NCCL): the sharded hot path with cross-rank BatchNorm statistics and the single flat-gradient all-reduce
against the single-process oracle with the reference's SyncBatchNorm + DDP semantics.
Skipped unless two GPUs are visible (`gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mvf_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

HC = O.HeadCfg(c_in=64, n_entities=3, pool_channels=32, fc_channels=(64, 64), hidden=32, d_ff=64, n_heads=4, n_layers=2,
               emb=16, proj=16, train_frames=8)
BV_GLOBAL, T, PTOK = 4, 8, 16


def _inputs():
    P = O.init_params(HC, seed=41)
    tokens, seq_lens, steps, masks = O.synth_batch(BV_GLOBAL, T, PTOK, HC.c_in, seed=42)
    return P, tokens, seq_lens, steps, masks


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from video_rep_learning_b200 import parallel
        P, tokens, seq_lens, steps, masks = _inputs()
        v0, v1 = parallel.shard_videos(BV_GLOBAL, rank, world)
        r = H.run_cuda(HC, P, None, tokens[2 * v0:2 * v1], masks[2 * v0:2 * v1], seq_lens[v0:v1], steps[v0:v1],
                       dtype=torch.float32, device=f"cuda:{rank}")
        torch.cuda.synchronize()
        torch.save({"loss": r["loss"], "grads": r["grads"], "bufs": r["bufs"]}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(600)
def test_two_gpu_step_matches_syncbn_ddp_oracle(tmp_path):
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(str(tmp_path), "rank0.pt"))
    r1 = torch.load(os.path.join(str(tmp_path), "rank1.pt"))
    keys = list(r0["grads"].keys())
    for k in keys:                                           # identical gradients on every rank after the all-reduce
        assert torch.equal(r0["grads"][k], r1["grads"][k]), k
    for k in r0["bufs"]:                                     # identical running statistics (global batch statistics)
        assert torch.equal(r0["bufs"][k], r1["bufs"][k]), k

    P, tokens, seq_lens, steps, masks = _inputs()
    Pr = {k: v.double().clone().requires_grad_(True) for k, v in P.items()}
    buf = {k: (v.double() if v.is_floating_point() else v) for k, v in O.init_bn_buffers(HC).items()}
    emb, nb = O.head_forward(Pr, buf, tokens.double(), masks.double(), HC, True)
    e, nb2 = O.proj_forward(Pr, buf, emb, HC, True)
    nb.update(nb2)
    e = e.view(BV_GLOBAL, 2, T, -1)
    losses = []
    for r in range(world):
        v0, v1 = 2 * r, 2 * r + 2
        losses.append(O.scl_loss_dense(e[v0:v1], seq_lens[v0:v1], steps[v0:v1], masks[2 * v0:2 * v1].double()))
    (sum(losses) / world).backward()
    want = {k: v.grad for k, v in Pr.items()}
    assert abs(float(r0["loss"]) - float(losses[0])) / float(losses[0]) < 1e-5
    assert abs(float(r1["loss"]) - float(losses[1])) / float(losses[1]) < 1e-5
    assert H.rel_l2(H.grad_vector(r0["grads"], keys), H.grad_vector(want, keys)) < 2e-5
    for k, v in nb.items():
        if v.is_floating_point():
            assert float((r0["bufs"][k].double() - v).abs().max()) < 1e-5, k


def _allreduce_worker(rank, world, port, out_dir, multicast):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["MVF_PEER_MULTICAST"] = "1" if multicast else "0"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from video_rep_learning_b200 import parallel
        dev = torch.device("cuda", rank)
        res = {}
        for n in (4, 1000, 4_800_003):
            g = torch.Generator(device="cpu").manual_seed(100 * rank + n % 97)
            x = torch.randn(n, generator=g)
            for rep in range(3):                                  # the buffer and its exchange counters are reused every step
                flat = parallel.flat_grad_buffer(n, dev)
                obj = [o for o in parallel.PeerFlatGrads._cache.values() if o.owns(flat)]
                assert obj, "symmetric-memory gradient buffer was not created"
                flat.copy_(x.to(dev) * (rep + 1))
                scale = parallel.finish_flat_grads_(flat)
                torch.cuda.synchronize()
                assert scale == 1.0 / world
                res[(n, rep)] = flat.cpu().clone()
            res[("mc", n)] = bool(obj[0].mc_ptr)
        torch.save(res, os.path.join(out_dir, f"ar{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(600)
@pytest.mark.parametrize("multicast", [True, False], ids=["multimem", "peer_loads"])
def test_two_gpu_flat_gradient_allreduce(tmp_path, multicast):
    """mvf_peer_allreduce_f32 (NVSwitch multimem path and the peer load/store path): the in-place sum of the symmetric flat
    buffer equals the sum of the ranks' inputs (one fp32 addition per element at world 2: exact) and is bitwise identical on
    both ranks, call after call on the same buffer."""
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_allreduce_worker, args=(world, port, str(tmp_path), multicast), nprocs=world, join=True)
    r = [torch.load(os.path.join(str(tmp_path), f"ar{k}.pt")) for k in range(world)]
    for n in (4, 1000, 4_800_003):
        xs = [torch.randn(n, generator=torch.Generator(device="cpu").manual_seed(100 * k + n % 97)) for k in range(world)]
        for rep in range(3):
            want = xs[0] * (rep + 1) + xs[1] * (rep + 1)
            assert torch.equal(r[0][(n, rep)], r[1][(n, rep)])
            assert torch.equal(r[0][(n, rep)], want), (n, rep, float((r[0][(n, rep)] - want).abs().max()))
        print(f"n = {n}: multicast mapping {'present' if r[0][('mc', n)] else 'absent'}")
