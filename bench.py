#!/usr/bin/env python
"""Benchmark of the MV-Former training hot path (head + projection + SCL, forward + backward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE.json configs[1]: Penn Action MV-Former shape, ViT-B/16 tokens x 3 feature layers (C_in 2304,
P 196), 32 videos x 20 frames x 2 views per GPU, bf16 operands / fp32 accumulate, penn_mvf.yml head sizes,
dropout 0.1 (training mode), synthetic tokens, deterministic synthetic parameters.  With N GPUs every rank runs
32 videos (weak scaling; N = 8 is BASELINE configs[2]'s global batch of 256), BatchNorm statistics are exchanged
across ranks and the flat head-gradient buffer is all-reduced once per step over NCCL.

One JSON line on stdout (rank 0).  `value` = videos/s with tokens already resident in HBM, the step (forward, SCL,
backward, cross-rank exchanges) captured once as a CUDA graph and replayed (`launch_mode`; `eager` = the same step issued
kernel by kernel through the drop-in Python API, `--eager` makes that the timed leg); `e2e` = videos/s
through the public Python API (model + algos.SCL) with the step's tokens copied from pinned host memory inside
the timed region and the loss read back; `roofline` = the dominant kernel timed with CUDA events on its launching
stream inside the timed steps (event-record nodes of the graph): with the default folded entity pooling that is the streaming pooling pass over the
tokens (HBM-bound, against the measured copy bandwidth); `dense_path` = the same step with the pooling evaluated as
written in the reference (K|V projection GEMM on tcgen05 + attention over K|V), with the GEMM's fraction of the
measured cuBLAS bf16 peak; `cpu_baseline` = the CPU oracle port of the reference algorithm on this host's cores on
a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = dict(name="penn_mvf_vitb16x3_bv32_T20_bf16", videos_per_gpu=32, T=20, P=196, c_in=2304, entities=3)
METRIC = "training videos/sec (head+SCL fwd/bwd)"
UNIT = "videos/s"


def workload_config(world: int) -> dict:
    """The `config` object of the JSON line (both arms print the same one)."""
    Bv = WORKLOAD["videos_per_gpu"]
    return dict(workload=WORKLOAD["name"], videos_per_gpu=Bv, global_videos=Bv * world, frames=WORKLOAD["T"], views=2,
                patch_tokens=WORKLOAD["P"], token_channels=WORKLOAD["c_in"], entities=WORKLOAD["entities"], dropout=0.1,
                l2="inputs (1.16 GB of tokens per step) larger than L2; no flush needed",
                parallelism=f"dp{world} (video shards; BN statistics + one flat gradient all-reduce)")


def flops_per_video(T=20, P=196, c_in=2304, E=3, SPC=384, FC=512, H=256, DFF=1024, L=3, D=128, PS=128):
    """SURVEY.md section 8d formula (as-written dense contractions, 2 FLOP/MAC)."""
    F2 = 2 * T
    kv = 2 * 2 * P * c_in * SPC * F2
    xatt = 2 * 2 * E * P * SPC * F2
    mlp = 2 * (F2 * E) * ((SPC + E) * FC + FC * FC + FC * H)
    S = E * T
    enc_lin = L * 2 * (8 * S * H * H + 4 * S * H * DFF)
    enc_att = L * 2 * (4 * S * S * H)
    tail = 2 * F2 * (H * D + 2 * D * PS)
    scl = 2 * T * T * D
    return dict(kv_fwd=kv, total=kv * 2 + (xatt + mlp + enc_lin + tail + scl) * 3 + enc_att * 3.5)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1399.9))),
                    tflops_burst=float(d.get("bf16_tflops", 1640.0)), hbm=float(d.get("hbm_gbs", 6538.9)), source="measured")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


def head_cfg():
    from oracle import mvf_oracle as O
    return O.HeadCfg(c_in=WORKLOAD["c_in"], train_frames=WORKLOAD["T"], drop_p=0.1)


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm (the reference is pure Python and /root/reference is not on the
# GPU box) on the host cores, bounded sample of the same workload
# ------------------------------------------------------------------------------------------------------------------
def cpu_step_time(sample_videos: int, iters: int, threads: int):
    import torch
    from oracle import mvf_oracle as O
    torch.set_num_threads(threads)
    hc = head_cfg()
    hc.drop_p = 0.0
    P = {k: v.requires_grad_(True) for k, v in O.init_params(hc, seed=1).items()}
    tokens, seq_lens, steps, masks = O.synth_batch(sample_videos, WORKLOAD["T"], WORKLOAD["P"], hc.c_in, seed=1)
    times = []
    for it in range(iters + 1):
        t0 = time.perf_counter()
        emb, _ = O.head_forward(P, None, tokens, masks, hc, True)
        e, _ = O.proj_forward(P, None, emb, hc, True)
        loss = O.scl_loss_dense(e.view(sample_videos, 2, WORKLOAD["T"], -1), seq_lens, steps, masks)
        loss.backward()
        for v in P.values():
            v.grad = None
        if it > 0:
            times.append(time.perf_counter() - t0)
    return times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = WORKLOAD["videos_per_gpu"]
    # warm-up + steps, each step = one fwd+bwd over the workload's batch (32 videos); at most 8 timed steps (~1.2 s each on
    # the GPU box's 16 cores) so that the run ends within a minute whatever --steps says
    times = cpu_step_time(sample, max(1, min(args.steps, 8)), threads)
    ms = 1e3 * statistics.mean(times)
    v = sample / (ms / 1e3)
    out = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=len(times), warmup=1,
               ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
               config=dict(workload_config(max(1, args.gpus)),
                           note="CPU arm: oracle port of the reference algorithm on the host cores of rank 0, fp32, dropout 0, "
                                "one step = the workload's 32-video batch"),
               cpu_baseline=dict(value=v, unit=UNIT, cores=threads, kind="port",
                                 sample=f"{sample} videos x 20 frames x 2 views of the cfg2 shape per step, {len(times)} steps"),
               e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
class TokenBackbone:
    """Stand-in for the frozen ViT (upstream producer, not part of the product): returns pre-computed tokens."""


def run_ours(args):
    import torch
    import torch.distributed as dist
    from oracle import mvf_oracle as O          # synthetic parameter / batch generators only (host side, untimed)
    from video_rep_learning_b200 import _lib as L
    from video_rep_learning_b200 import engine
    from video_rep_learning_b200.algos import get_algo
    from video_rep_learning_b200.config import mvf_cfg
    from video_rep_learning_b200.models import build_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # --overlap: the chain's gradient all-reduce runs beside the pooling backward, which leaves --reserve-sms SMs free
        # for it: keep NCCL's CTA count within that
        if args.overlap:
            os.environ.setdefault("NCCL_MAX_NCHANNELS", str(args.reserve_sms))
        dist.init_process_group("nccl", device_id=dev)
    Bv, T, P, C_in = WORKLOAD["videos_per_gpu"], WORKLOAD["T"], WORKLOAD["P"], WORKLOAD["c_in"]
    BV = 2 * Bv

    class _NoBackbone(torch.nn.Module):
        def forward(self, x):  # pragma: no cover - tokens are fed directly
            raise RuntimeError("bench feeds patch tokens directly (frozen ViT is upstream of the hot path)")

    cfg = mvf_cfg(c_in=C_in, num_frames=T)
    torch.manual_seed(1)
    model = build_model(cfg, backbone=_NoBackbone()).to(dev)
    hc = head_cfg()
    sd = O.init_params(hc, seed=1)
    model.load_state_dict({k: v for k, v in sd.items()}, strict=False)
    model.train()
    algo = get_algo(cfg)

    # synthetic step inputs: per-rank seed so ranks hold different videos
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    tokens_dev = torch.randn(BV, T, P, C_in, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    _, seq_lens, steps, masks = O.synth_batch(Bv, T, 1, 1, seed=1 + rank)
    seq_lens_d, steps_d, masks_d = seq_lens.to(dev), steps.to(dev), masks.to(dev)
    params = [p for n, p in model.named_parameters() if "backbone" not in n]

    def step(tok):
        for p in params:
            p.grad = None
        embs = model.forward_tokens(tok, video_masks=masks_d, project=True)
        loss = algo.compute_sequence_loss(embs.view(Bv, 2, T, -1), seq_lens_d, steps_d, masks_d)["loss"]
        loss.backward()
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    lib = L.lib()
    head_opts = model.run_options
    head_opts.overlap_grad_allreduce = bool(args.overlap)
    head_opts.pool_bwd_reserve_sms = args.reserve_sms

    def timed_region(pool_mode, steps, warmup, sample_clocks):
        """W untimed + K timed steps with the tokens resident in HBM; CUDA events; max over ranks."""
        head_opts.pool_mode = pool_mode
        for _ in range(warmup):
            step(tokens_dev)
        sync_all()
        clocks = ClockSampler(local_rank)
        if sample_clocks and rank == 0:
            clocks.start()
        lib.mvf_profile_enable(1)
        n0 = lib.mvf_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        ev0.record()
        for _ in range(steps):
            loss = step(tokens_dev)
        ev1.record()
        sync_all()
        ms_total = ev0.elapsed_time(ev1)
        launches = int(lib.mvf_launch_count() - n0)
        buf = (ctypes.c_float * 512)()
        n = ctypes.c_int(0)
        prof = {}
        for tag in range(6):
            lib.mvf_profile_read(tag, buf, 512, ctypes.byref(n))
            prof[tag] = [buf[i] for i in range(n.value)]
        lib.mvf_profile_enable(0)
        clk = clocks.stop() if (sample_clocks and rank == 0) else None
        t = torch.tensor([ms_total], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return dict(ms_step=float(t.item()) / steps, launches=launches, prof=prof, clocks=clk, loss=float(loss.item()))

    def read_prof():
        buf = (ctypes.c_float * 512)()
        n = ctypes.c_int(0)
        prof = {}
        for tag in range(6):
            lib.mvf_profile_read(tag, buf, 512, ctypes.byref(n))
            prof[tag] = [buf[i] for i in range(n.value)]
        return prof

    def graph_region(pool_mode, steps, warmup, sample_clocks):
        """The same step captured once as a CUDA graph (video_rep_learning_b200.graph.GraphedTrainStep) and replayed:
        W untimed + K timed replays, tokens resident in HBM, CUDA events, max over ranks.  The library's event brackets
        around the dominant kernels are part of the graph (external event-record nodes); they are read after the timed
        region from extra replays, one synchronisation per replay."""
        from video_rep_learning_b200.graph import GraphedTrainStep
        head_opts.pool_mode = pool_mode
        gs = GraphedTrainStep(model, algo, Bv, T, P, C_in, dtype=torch.bfloat16, device=dev, micro_batches=args.micro_batches)
        gs.adopt_tokens(tokens_dev)
        gs.set_inputs(seq_lens=seq_lens_d, steps=steps_d, masks=masks_d)
        gs.capture(profile=True)
        for _ in range(warmup):
            gs()
        sync_all()
        clocks = ClockSampler(local_rank)
        if sample_clocks and rank == 0:
            clocks.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        ev0.record()
        for _ in range(steps):
            loss = gs()
        ev1.record()
        sync_all()
        ms_total = ev0.elapsed_time(ev1)
        clk = clocks.stop() if (sample_clocks and rank == 0) else None
        prof = {tag: [] for tag in range(6)}
        for _ in range(min(steps, 20)):
            gs()
            torch.cuda.synchronize()
            for tag, vals in read_prof().items():
                prof[tag] += vals
        lib.mvf_profile_enable(0)
        final = float(loss.item())
        launches = gs.launches_per_step * steps
        gs.release()
        t = torch.tensor([ms_total], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return dict(ms_step=float(t.item()) / steps, launches=launches, prof=prof, clocks=clk, loss=final)

    pool_default = L.POOL_DENSE if args.pool == "dense" else L.POOL_FOLDED
    graph_note = None
    eager_run = None
    if args.eager:
        main_run = timed_region(pool_default, args.steps, max(args.warmup, 3), True)
    else:
        try:
            main_run = graph_region(pool_default, args.steps, max(args.warmup, 3), True)
            graph_note = "one cudaGraphLaunch per step (GraphedTrainStep)" + (
                f", views run as {args.micro_batches} slices on {args.micro_batches} streams inside the graph" if args.micro_batches > 1 else "")
            eager_run = timed_region(pool_default, max(3, min(args.steps, 20)), 3, False)
            if args.micro_batches > 1:
                # the pooling kernels of the slices queue behind each other on different streams, so an event pair around
                # one of them also times its wait for SMs: the roofline figures come from the un-split eager leg below
                # (same kernels, same tokens, whole batch per launch, events on the launching stream)
                main_run["prof"] = eager_run["prof"]
                main_run["prof_ms_step"] = eager_run["ms_step"]
        except Exception as e:  # capture refused (e.g. a collective that cannot be captured): fall back to eager launches
            graph_note = f"capture failed, eager launches timed instead: {type(e).__name__}: {str(e)[:200]}"
            torch.cuda.synchronize()
            main_run = timed_region(pool_default, args.steps, max(args.warmup, 3), True)
    ms_step, launches, prof, clk, final_loss = (main_run["ms_step"], main_run["launches"], main_run["prof"],
                                                main_run["clocks"], main_run["loss"])
    prof_ms_step = main_run.get("prof_ms_step", ms_step)
    value = world * Bv / (ms_step / 1e3)
    dense_run = None
    if args.pool == "folded" and not args.no_dense and world == 1:
        dense_run = timed_region(L.POOL_DENSE, max(3, min(args.steps, 10)), 3, False)
        head_opts.pool_mode = pool_default

    e2e_value = e2e_ms = None
    h2d = d2h = 0
    k2 = 0
    if not args.no_e2e:
        # ---- timed region 2: end to end from pinned host memory through the public API --------------------------
        n_host = 2
        host_tokens = [torch.empty(BV, T, P, C_in, dtype=torch.bfloat16).pin_memory() for _ in range(n_host)]
        for h in host_tokens:
            h.copy_(tokens_dev.cpu())
        host_meta = [(seq_lens.pin_memory(), steps.pin_memory(), masks.pin_memory()) for _ in range(n_host)]
        copy_stream = torch.cuda.Stream(device=dev)
        dev_tok = [torch.empty_like(tokens_dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            slot = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])
                dev_tok[slot].copy_(host_tokens[i % n_host], non_blocking=True)
                ready[slot].record(copy_stream)

        def e2e_loop(k):
            losses = []
            for c in consumed:
                c.record()
            prefetch(0)
            for i in range(k):
                slot = i % 2
                if i + 1 < k:
                    prefetch(i + 1)
                torch.cuda.current_stream().wait_event(ready[slot])
                sl, st_, mk = host_meta[i % n_host]
                sl_d, st_d, mk_d = sl.to(dev, non_blocking=True), st_.to(dev, non_blocking=True), mk.to(dev, non_blocking=True)
                for p in params:
                    p.grad = None
                embs = model.forward_tokens(dev_tok[slot], video_masks=mk_d, project=True)
                loss = algo.compute_sequence_loss(embs.view(Bv, 2, T, -1), sl_d, st_d, mk_d)["loss"]
                loss.backward()
                consumed[slot].record()
                losses.append(loss.item())            # D2H read of the step's result
            return losses

        e2e_loop(2)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k2 = max(3, min(args.steps, 10))
        e0.record()
        e2e_loop(k2)
        e1.record()
        sync_all()
        t2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_ms = float(t2.item()) / k2
        e2e_value = world * Bv / (e2e_ms / 1e3)
        h2d = tokens_dev.numel() * 2 + seq_lens.numel() * 8 + steps.numel() * 8 + masks.numel() * 4
        d2h = 4

    if rank == 0:
        fl = flops_per_video()
        peaks = measured_peaks()
        mean = lambda xs: statistics.mean(xs) if xs else None
        F_frames = BV * T

        def dense_roofline(pr, ms):
            kv_ms = mean(pr[0])
            if not kv_ms:
                return None
            kv_flops = fl["kv_fwd"] * Bv
            achieved = kv_flops / (kv_ms * 1e-3) / 1e12
            traffic = None
            tp = os.path.join(ROOT, "profiles", "kv_proj_fwd_traffic.json")
            if os.path.exists(tp):
                with open(tp) as f:
                    traffic = json.load(f).get("dram_bytes_per_launch")
            r = dict(bound="tensor", kernel="gemm_tc_kernel<256,4,2> (K|V projection, forward)", achieved=achieved,
                     peak=peaks["tflops"], unit="TFLOP/s", frac=achieved / peaks["tflops"], traffic=traffic,
                     peak_source=f"{peaks['source']} bf16_tflops_sustained", frac_of_nominal_2250=achieved / 2250.0,
                     ms_per_launch=kv_ms, launches_timed=len(pr[0]), flops_per_launch=kv_flops)
            if pr[1]:
                dw_ms = mean(pr[1])
                r["weight_grad_gemm"] = dict(ms_per_launch=dw_ms, achieved=kv_flops / (dw_ms * 1e-3) / 1e12,
                                             frac=kv_flops / (dw_ms * 1e-3) / 1e12 / peaks["tflops"])
            r["share_of_step"] = dict(kv_proj_fwd=kv_ms / ms, kv_proj_dw=(mean(pr[1]) / ms) if pr[1] else None,
                                      xattn_fwd=(mean(pr[2]) / ms) if pr[2] else None,
                                      xattn_bwd=(mean(pr[3]) / ms) if pr[3] else None)
            return r

        def folded_roofline(pr, ms):
            f_ms = mean(pr[0])
            if not f_ms:
                return None
            # algorithmic bytes of one launch: every token read once (bf16) + the pooled rows and the
            # attention maps written (fp32) + the folded query matrix read
            E = WORKLOAD["entities"]
            tok_bytes = F_frames * P * C_in * 2
            fwd_bytes = tok_bytes + F_frames * E * C_in * 4 + F_frames * E * P * 4 + E * C_in * 4
            achieved = fwd_bytes / (f_ms * 1e-3) / 1e9
            traffic = None
            tp = os.path.join(ROOT, "profiles", "pool_fold_fwd_traffic.json")
            if os.path.exists(tp):
                with open(tp) as f:
                    traffic = json.load(f).get("dram_bytes_per_launch")
            r = dict(bound="hbm", kernel="pool_foldw_fwd_kernel<9,true> (folded entity pooling: one warp-specialised streaming pass over the bf16 tokens, mma.sync)",
                     achieved=achieved, peak=peaks["hbm"], unit="GB/s", frac=achieved / peaks["hbm"], traffic=traffic,
                     peak_source=f"{peaks['source']} hbm_gbs (copy, read+write)", frac_of_nominal_8000=achieved / 8000.0,
                     ms_per_launch=f_ms, launches_timed=len(pr[0]), bytes_per_launch=fwd_bytes)
            if pr[1]:
                b_ms = mean(pr[1])
                bwd_bytes = tok_bytes + 2 * F_frames * E * C_in * 4 + F_frames * E * P * 4
                r["backward_pass"] = dict(kernel="pool_foldw_bwd_kernel<9,true>", ms_per_launch=b_ms, bytes_per_launch=bwd_bytes,
                                          achieved=bwd_bytes / (b_ms * 1e-3) / 1e9,
                                          frac=bwd_bytes / (b_ms * 1e-3) / 1e9 / peaks["hbm"])
            r["share_of_step"] = dict(pool_stream_fwd=f_ms / ms, pool_stream_bwd=(mean(pr[1]) / ms) if pr[1] else None,
                                      pool_rest_fwd=((mean(pr[2]) or 0) + (mean(pr[4]) or 0)) / ms if pr[4] else None,
                                      pool_rest_bwd=((mean(pr[3]) or 0) + (mean(pr[5]) or 0)) / ms if pr[3] else None)
            return r

        roof = dense_roofline(prof, prof_ms_step) if args.pool == "dense" else folded_roofline(prof, prof_ms_step)
        if roof is not None:
            roof["timed_in"] = ("eager leg of this run: the un-split step issued kernel by kernel (CUDA events on the launching stream, "
                                f"{prof_ms_step:.3f} ms/step); share_of_step is relative to that leg") if "prof_ms_step" in main_run \
                else "the timed region of `value`"
        dense = None
        if dense_run is not None:
            dms = dense_run["ms_step"]
            dense = dict(value=world * Bv / (dms / 1e3), unit=UNIT, ms_per_step=dms, gpu_launches=dense_run["launches"],
                         whole_step_tflops=fl["total"] * Bv * world / (dms * 1e-3) / 1e12,
                         note="same step, entity pooling evaluated as written (K|V GEMM on tcgen05 + attention over K|V)",
                         roofline=dense_roofline(dense_run["prof"], dms))
        whole = fl["total"] * Bv * world / (ms_step * 1e-3) / 1e12
        cpu = None
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            sample = 16       # half of the workload's batch per iteration: ~10 s of CPU work with the warm-up on 16 cores
            times = cpu_step_time(sample, 6, threads)
            cms = statistics.mean(times)
            cpu = dict(value=sample / cms, unit=UNIT, cores=threads, kind="port",
                       sample=f"{sample} videos of the cfg2 shape (fp32 oracle port), {len(times)} timed iterations after 1 warm-up")
        out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                   ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                   data="synthetic",
                   config=workload_config(world),
                   as_written_tflops=whole, gflop_per_video_as_written=fl["total"] / 1e9, loss=final_loss,
                   pooling=args.pool, launch_mode=("eager" if args.eager else "cuda_graph"), launch_note=graph_note,
                   eager=None if eager_run is None else dict(
                       value=world * Bv / (eager_run["ms_step"] / 1e3), unit=UNIT, ms_per_step=eager_run["ms_step"],
                       note="same step issued kernel by kernel through the drop-in Python API (no graph)"),
                   roofline=roof, dense_path=dense, cpu_baseline=cpu,
                   e2e=None if args.no_e2e else dict(
                       value=e2e_value, unit=UNIT, ms_per_step=e2e_ms, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=k2,
                       note="pinned host tokens -> HBM on a copy stream (double buffered) + loss.item() per step"),
                   gpu_launches=launches, clocks=clk)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the host-to-device leg")
    ap.add_argument("--no-cpu", action="store_true", help="profiling runs: skip the CPU baseline sample")
    ap.add_argument("--pool", default="folded", choices=["folded", "dense"],
                    help="entity pooling: folded (default product path) or dense (as written: K|V GEMM + attention)")
    ap.add_argument("--eager", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--micro-batches", type=int, default=1, help="slices of the captured step (1 = un-split)")
    ap.add_argument("--overlap", action="store_true",
                    help="multi-GPU: all-reduce the chain's gradients beside the pooling backward (measured slower; default off)")
    ap.add_argument("--reserve-sms", type=int, default=16, help="SMs the pooling backward leaves to the overlapped all-reduce")
    ap.add_argument("--no-dense", action="store_true", help="skip the short as-written (dense pooling) comparison leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
