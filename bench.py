#!/usr/bin/env python
"""Benchmark of the MV-Former training hot path (head + projection + SCL, forward + backward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload penn_cfg2|finegym_cfg4|long_cfg5]
                    [--impl ours|reference|reference-gpu] [--global-videos G] [--no-overlap]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (per-GPU shards of BASELINE.json configs; synthetic tokens, deterministic synthetic
parameters, bf16 tokens / fp32 accumulate, dropout 0.1, training mode):
  penn_cfg2     (default; the configuration the metric is quoted on)  32 videos x 2 views x 20 frames, ViT-B/16 x 3 feature
                layers (C_in 2304, P 196), penn_mvf.yml head (3 entities).  With N GPUs every rank runs 32 videos (weak
                scaling; N = 8 is configs[2]'s global batch of 256).
  finegym_cfg4  8 videos x 2 views x 80 frames per GPU (64 global on 8 GPUs), fg99_mvf.yml head: 6 entities, FC 1536,
                D 256, SMART_FINAL avg -> temporal attention over S = 480 tokens.
  long_cfg5     4 videos x 2 views x 240 frames per GPU (32 global on 8 GPUs), penn head with 16 entities -> S = 3840.

One JSON line on stdout (rank 0).  `value` = videos/s with tokens already resident in HBM, the step (forward, SCL,
backward, cross-rank exchanges) captured once as a CUDA graph and replayed (`launch_mode`; `eager` = the same step issued
kernel by kernel through the drop-in Python API, `--eager` makes that the timed leg); `e2e` = videos/s through the public
Python API (model + algos.SCL) with the step's tokens copied from pinned host memory inside the timed region and the loss
read back; `roofline` = the kernel that dominates THIS workload's step, timed with CUDA events on its launching stream
inside the timed steps (event-record nodes of the graph); `parity` = loss / embeddings / concatenated gradient of one
untimed dropout-free step on the bench's own inputs against the CPU oracle (and, across ranks, against a single-process
run of the concatenated batch); `reference_gpu` = the reference's own PyTorch modules run eagerly on the same GPU on the
same inputs (the like-for-like baseline a user of the reference runs today); `cpu_baseline` = the reference algorithm on
this host's cores on a bounded sample; `cross_rank` (N > 1) = what the gradient all-reduce and the BatchNorm exchanges cost
inside the replayed graph, by difference against the same step captured without them.  `--global-videos G` fixes the global
batch (strong scaling, `scaling: "strong"`); `--no-overlap` sums the whole gradient buffer at the end of the step.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "training videos/sec (head+SCL fwd/bwd)"
UNIT = "videos/s"

# per-GPU shards; `head` = keyword arguments of video_rep_learning_b200.config.mvf_cfg that differ from penn_mvf.yml
WORKLOADS = {
    "penn_cfg2": dict(name="penn_mvf_vitb16x3_bv32_T20_bf16", videos_per_gpu=32, T=20, P=196, c_in=2304,
                      head=dict(entities=3, capacity=2, emb=128, final="one", smart_feats="3,7,11"),
                      baseline_config="BASELINE.json configs[1] (configs[2] at 8 GPUs)", cpu_sample=16, ref_sample=32),
    "finegym_cfg4": dict(name="finegym_mvf_vitb16x3_bv8_T80_E6_bf16", videos_per_gpu=8, T=80, P=196, c_in=2304,
                         head=dict(entities=6, capacity=6, emb=256, final="avg", smart_feats="9,10,11"),
                         baseline_config="BASELINE.json configs[3] (64 videos over 8 GPUs)", cpu_sample=4, ref_sample=8),
    "long_cfg5": dict(name="long_mvf_vitb16x3_bv4_T240_E16_bf16", videos_per_gpu=4, T=240, P=196, c_in=2304,
                      head=dict(entities=16, capacity=2, emb=128, final="one", smart_feats="3,7,11"),
                      baseline_config="BASELINE.json configs[4] (32 videos over 8 GPUs)", cpu_sample=1, ref_sample=2),
}
WL = WORKLOADS["penn_cfg2"]      # selected in main()


def workload_config(world: int) -> dict:
    """The `config` object of the JSON line (every arm prints the same one)."""
    Bv, h = WL["videos_per_gpu"], WL["head"]
    tok_gb = 2 * Bv * WL["T"] * WL["P"] * WL["c_in"] * 2 / 1e9
    return dict(workload=WL["name"], baseline_config=WL["baseline_config"], videos_per_gpu=Bv, global_videos=Bv * world,
                frames=WL["T"], views=2, patch_tokens=WL["P"], token_channels=WL["c_in"], entities=h["entities"],
                fc_width=256 * h["capacity"], embedding=h["emb"], smart_final=h["final"],
                temporal_tokens=h["entities"] * WL["T"], dropout=0.1,
                l2=f"inputs ({tok_gb:.2f} GB of tokens per step) larger than L2; no flush needed",
                parallelism=f"dp{world} (video shards; BN statistics + one flat gradient all-reduce)")


def flops_per_video(T=None, P=None, c_in=None, E=None, SPC=384, FC=None, H=256, DFF=1024, L=3, D=None, PS=128):
    """SURVEY.md section 8d formula (as-written dense contractions, 2 FLOP/MAC)."""
    T = WL["T"] if T is None else T
    P = WL["P"] if P is None else P
    c_in = WL["c_in"] if c_in is None else c_in
    E = WL["head"]["entities"] if E is None else E
    FC = 256 * WL["head"]["capacity"] if FC is None else FC
    D = WL["head"]["emb"] if D is None else D
    F2 = 2 * T
    kv = 2 * 2 * P * c_in * SPC * F2
    xatt = 2 * 2 * E * P * SPC * F2
    mlp = 2 * (F2 * E) * ((SPC + E) * FC + FC * FC + FC * H)
    S = E * T
    enc_lin = 2 * L * 2 * (4 * S * H * H + 2 * S * H * DFF)
    enc_att = 2 * L * 2 * (2 * S * S * H)
    tail = 2 * F2 * (H * D + 2 * D * PS)
    scl = 2 * T * T * D
    return dict(kv_fwd=kv, enc_att_fwd=enc_att,
                total=kv * 2 + (xatt + mlp + enc_lin + tail + scl) * 3 + enc_att * 3.5)


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


def head_cfg(drop_p: float = 0.1):
    from oracle import mvf_oracle as O
    h = WL["head"]
    fc = 256 * h["capacity"]
    return O.HeadCfg(c_in=WL["c_in"], n_entities=h["entities"], fc_channels=(fc, fc), emb=h["emb"], final=h["final"],
                     train_frames=WL["T"], drop_p=drop_p)


def model_cfg(drop: float = 0.1):
    from video_rep_learning_b200.config import mvf_cfg
    h = WL["head"]
    return mvf_cfg(c_in=WL["c_in"], num_frames=WL["T"], entities=h["entities"], capacity=h["capacity"], emb=h["emb"],
                   final=h["final"], smart_feats=h["smart_feats"], drop=drop)


def bind_to_gpu_numa_node(local_rank: int):
    """Several ranks on one host: run this rank's host threads (and therefore its pinned staging buffers, first touch) on
    the CPUs local to its GPU, instead of all ranks sharing whatever node the launcher started them on."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return dict(pci=bdf, cpus=len(cpus))
    except Exception as e:   # no sysfs / not permitted: keep the launcher's affinity
        return dict(error=f"{type(e).__name__}: {e}"[:120])
    return None


# ------------------------------------------------------------------------------------------------------------------
# baseline arms: the reference's own PyTorch modules (baseline/_ref, placed there unmodified by build()) through
# oracle/ref_shim.py; when they are not available, the oracle port of the same algorithm
# ------------------------------------------------------------------------------------------------------------------
def reference_modules(hc, params, device, dtype):
    """(head, proj, algo, kind) of the reference on `device`; kind = "reference" (its own modules) or None."""
    try:
        from oracle import ref_shim as R
        if not R.available():
            return None
        yml = "fg99_mvf.yml" if WL is WORKLOADS["finegym_cfg4"] else "penn_mvf.yml"
        _cfg, head, proj, algo = R.build_reference_modules(hc, params, yml)
        return head.to(device=device, dtype=dtype).train(), proj.to(device=device, dtype=dtype).train(), algo, R
    except Exception as e:  # pragma: no cover - depends on the box
        sys.stderr.write(f"bench: reference modules unavailable ({type(e).__name__}: {e}); using the oracle port\n")
        return None


def cpu_step_time(sample_videos: int, iters: int, threads: int):
    """fwd+bwd of head + MLPHead + SCL on the host cores, fp32, dropout 0: the reference's modules when present
    (kind "reference"), else the oracle port.  Returns (times, kind)."""
    import torch
    from oracle import mvf_oracle as O
    torch.set_num_threads(threads)
    hc = head_cfg(0.0)
    T = WL["T"]
    Pm = O.init_params(hc, seed=1)
    tokens, seq_lens, steps, masks = O.synth_batch(sample_videos, T, WL["P"], hc.c_in, seed=1)
    ref = reference_modules(hc, Pm, torch.device("cpu"), torch.float32)
    times = []
    if ref is not None:
        head, proj, algo, R = ref
        x = R.tokens_to_nchw(tokens)
        for it in range(iters + 1):
            t0 = time.perf_counter()
            emb = head(x, video_masks=masks, cls_emb=None)
            e = torch.nn.functional.normalize(proj(emb), dim=-1)
            loss = algo.compute_sequence_loss(e.view(sample_videos, 2, T, -1), seq_lens, steps, masks)["loss"]
            loss.backward()
            for m in (head, proj):
                for p in m.parameters():
                    p.grad = None
            if it > 0:
                times.append(time.perf_counter() - t0)
        return times, "reference"
    P = {k: v.requires_grad_(True) for k, v in Pm.items()}
    for it in range(iters + 1):
        t0 = time.perf_counter()
        emb, _ = O.head_forward(P, None, tokens, masks, hc, True)
        e, _ = O.proj_forward(P, None, emb, hc, True)
        loss = O.scl_loss_dense(e.view(sample_videos, 2, T, -1), seq_lens, steps, masks)
        loss.backward()
        for v in P.values():
            v.grad = None
        if it > 0:
            times.append(time.perf_counter() - t0)
    return times, "port"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = WL["ref_sample"]
    # warm-up + steps, each step = one fwd+bwd over a bounded sample of the workload (for penn_cfg2 its whole 32-video
    # batch); at most 8 timed steps so that the run ends within a few minutes whatever --steps says
    times, kind = cpu_step_time(sample, max(1, min(args.steps, 8)), threads)
    ms = 1e3 * statistics.mean(times)
    v = sample / (ms / 1e3)
    what = ("the reference's own modules (CARL_MVF/models/mvformer.py, resnet_c2d.py:MLPHead, algos/scl.py, unmodified, "
            "via oracle/ref_shim.py)") if kind == "reference" else "oracle port of the reference algorithm"
    out = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=len(times), warmup=1,
               ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
               config=workload_config(max(1, args.gpus)),
               note=f"CPU arm: {what} on the host cores of rank 0, fp32, dropout 0, one step = {sample} videos of the workload",
               cpu_baseline=dict(value=v, unit=UNIT, cores=threads, kind=kind,
                                 sample=f"{sample} videos x {WL['T']} frames x 2 views of {WL['name']} per step, {len(times)} steps"),
               e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out), flush=True)


def reference_gpu_run(dev, tokens_dev, seq_lens, steps, masks, n_steps: int):
    """The reference's PyTorch modules, eager, on the same GPU and inputs: fp32 and bf16 autocast.  Tokens resident, the
    NCHW copy the reference expects made once outside the timed region."""
    import torch
    from oracle import mvf_oracle as O
    hc = head_cfg(0.1)
    Pm = O.init_params(hc, seed=1)
    ref = reference_modules(hc, Pm, dev, torch.float32)
    kind = "reference"
    Bv, T = WL["videos_per_gpu"], WL["T"]
    sl, st, mk = seq_lens.to(dev), steps.to(dev), masks.to(dev)
    out = {}
    if ref is not None:
        head, proj, algo, R = ref
        x32 = R.tokens_to_nchw(tokens_dev.float())

        def step(x):
            emb = head(x, video_masks=mk, cls_emb=None)
            e = torch.nn.functional.normalize(proj(emb), dim=-1)
            loss = algo.compute_sequence_loss(e.view(Bv, 2, T, -1), sl, st, mk)["loss"]
            loss.backward()
            for m in (head, proj):
                for p in m.parameters():
                    p.grad = None
            return loss
    else:
        kind = "port"
        P = {k: v.to(dev).requires_grad_(True) for k, v in Pm.items()}
        hc0 = head_cfg(0.0)
        x32 = tokens_dev.float()

        def step(x):
            emb, _ = O.head_forward(P, None, x, mk, hc0, True)
            e, _ = O.proj_forward(P, None, emb, hc0, True)
            loss = O.scl_loss_dense(e.view(Bv, 2, T, -1), sl, st, mk)
            loss.backward()
            for v in P.values():
                v.grad = None
            return loss

    def timed(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            loss = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, float(loss)

    try:
        ms, loss = timed(lambda: step(x32), n_steps)
        out["fp32"] = dict(value=Bv / (ms / 1e3), unit=UNIT, ms_per_step=ms, loss=loss, steps=n_steps)

        def step_amp():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return step(x32)
        ms, loss = timed(step_amp, n_steps)
        out["bf16_autocast"] = dict(value=Bv / (ms / 1e3), unit=UNIT, ms_per_step=ms, loss=loss, steps=n_steps)
    except Exception as e:  # pragma: no cover - e.g. out of memory on a small device
        out["error"] = f"{type(e).__name__}: {str(e)[:200]}"
    out["kind"] = kind
    out["note"] = ("reference modules (unmodified) in PyTorch eager on this GPU, same synthetic inputs and parameters, "
                   "tokens resident, dropout 0.1, NCHW copy made outside the timed region") if kind == "reference" else \
                  "oracle port in PyTorch eager on this GPU (reference files not present)"
    del x32
    torch.cuda.empty_cache()
    return out


def run_reference_gpu_arm(args):
    import torch
    from oracle import mvf_oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    Bv, T, P, C_in = WL["videos_per_gpu"], WL["T"], WL["P"], WL["c_in"]
    g = torch.Generator(device=dev).manual_seed(1)
    tokens_dev = torch.randn(2 * Bv, T, P, C_in, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    _, seq_lens, steps, masks = O.synth_batch(Bv, T, 1, 1, seed=1)
    r = reference_gpu_run(dev, tokens_dev, seq_lens, steps, masks, max(3, min(args.steps, 20)))
    best = r.get("bf16_autocast") or r.get("fp32") or {}
    out = dict(impl="reference-gpu", metric=METRIC, value=best.get("value"), unit=UNIT, n_gpus=1, steps=best.get("steps"),
               warmup=2, ms_per_step=best.get("ms_per_step"), higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="bf16 autocast", data="synthetic", config=workload_config(1), reference_gpu=r)
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
N_TAGS = 10


def run_ours(args):
    import torch
    import torch.distributed as dist
    from oracle import mvf_oracle as O          # synthetic parameter / batch generators + the untimed parity check
    from video_rep_learning_b200 import _lib as L
    from video_rep_learning_b200 import engine
    from video_rep_learning_b200.algos import get_algo
    from video_rep_learning_b200.models import build_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Bv, T, P, C_in = WL["videos_per_gpu"], WL["T"], WL["P"], WL["c_in"]
    E = WL["head"]["entities"]
    BV = 2 * Bv

    class _NoBackbone(torch.nn.Module):
        def forward(self, x):  # pragma: no cover - tokens are fed directly
            raise RuntimeError("bench feeds patch tokens directly (frozen ViT is upstream of the hot path)")

    def make_model(drop):
        torch.manual_seed(1)
        m = build_model(model_cfg(drop), backbone=_NoBackbone()).to(dev)
        m.load_state_dict({k: v for k, v in O.init_params(head_cfg(), seed=1).items()}, strict=False)
        return m.train()

    model = make_model(0.1)
    algo = get_algo(model_cfg(0.1))

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
    head_opts.overlap_grad_allreduce = not args.no_overlap
    head_opts.pool_bwd_reserve_sms = args.reserve_sms

    def read_prof():
        buf = (ctypes.c_float * 512)()
        n = ctypes.c_int(0)
        prof = {}
        for tag in range(N_TAGS):
            n.value = 0
            if lib.mvf_profile_read(tag, buf, 512, ctypes.byref(n)) != 0:
                prof[tag] = []
                continue
            prof[tag] = [buf[i] for i in range(n.value)]
        return prof

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_region(pool_mode, steps, warmup, sample_clocks):
        """W untimed + K timed steps with the tokens resident in HBM; CUDA events; max over ranks."""
        head_opts.pool_mode = pool_mode
        for _ in range(warmup):
            step(tokens_dev)
        sync_all()
        clocks = ClockSampler(local_rank)
        if sample_clocks and rank == 0:
            clocks.start()
        lib.mvf_profile_enable(2)
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
        prof = read_prof()
        lib.mvf_profile_enable(0)
        clk = clocks.stop() if (sample_clocks and rank == 0) else None
        return dict(ms_step=max_over_ranks(ms_total) / steps, launches=launches, prof=prof, prof_steps=steps, clocks=clk,
                    loss=float(loss.item()))

    def graph_region(pool_mode, steps, warmup, sample_clocks, extras=True):
        """The same step captured once as a CUDA graph (video_rep_learning_b200.graph.GraphedTrainStep) and replayed:
        W untimed + K timed replays, tokens resident in HBM, CUDA events, max over ranks.  The library's event brackets
        around the dominant kernels are part of the graph (external event-record nodes); they are read after the timed
        region from extra replays, one synchronisation per replay."""
        from video_rep_learning_b200.graph import GraphedTrainStep
        head_opts.pool_mode = pool_mode
        gs = GraphedTrainStep(model, algo, Bv, T, P, C_in, dtype=torch.bfloat16, device=dev)
        gs.adopt_tokens(tokens_dev)
        gs.set_inputs(seq_lens=seq_lens_d, steps=steps_d, masks=masks_d)
        # the timed graph brackets only the two streaming pooling kernels (every bracket is a pair of event-record nodes
        # that also cuts the launch overlap of its neighbours); the per-group shares come from the eager leg (level 2)
        gs.capture(profile=1)
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
        ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / steps
        # a K-step region of a ~1.5 ms step lasts only tens of ms: add a >= 1 s region of the same replays (reported as
        # `sustained`; `value` stays the exactly-K-steps figure of the contract)
        sustained = None
        n_sus = int(min(4000, math.ceil(1000.0 / max(ms_step, 1e-3))))
        if extras and n_sus > steps:
            sync_all()
            ev0.record()
            for _ in range(n_sus):
                gs()
            ev1.record()
            sync_all()
            sus_ms = max_over_ranks(ev0.elapsed_time(ev1)) / n_sus
            sustained = dict(steps=n_sus, ms_per_step=sus_ms, value=world * Bv / (sus_ms / 1e3), unit=UNIT)
        clk = clocks.stop() if (sample_clocks and rank == 0) else None
        prof = {tag: [] for tag in range(N_TAGS)}
        n_prof = min(steps, 20) if extras else 0
        for _ in range(n_prof):
            gs()
            torch.cuda.synchronize()
            for tag, vals in read_prof().items():
                prof[tag] += vals
        lib.mvf_profile_enable(0)
        final = float(loss.item())
        launches = gs.launches_per_step * steps
        gs.release()
        return dict(ms_step=ms_step, launches=launches, prof=prof, prof_steps=n_prof, clocks=clk, loss=final, sustained=sustained)

    def comm_breakdown(steps=50):
        """Several ranks: what the cross-rank steps cost inside the replayed graph, by difference -- the same step captured
        again (a) without the gradient all-reduce and (b) also with rank-local BatchNorm statistics.  Diagnostic legs after the
        timed region; their steps are not valid training steps and never enter `value`."""
        from video_rep_learning_b200 import parallel as par
        keep = (head_opts.allreduce_grads, head_opts.sync_bn)
        out = {}
        try:
            head_opts.allreduce_grads = False
            t_noar = graph_region(pool_default, steps, 5, False, extras=False)["ms_step"]
            head_opts.sync_bn = False
            t_local = graph_region(pool_default, steps, 5, False, extras=False)["ms_step"]
        finally:
            head_opts.allreduce_grads, head_opts.sync_bn = keep
        fg = list(par.PeerFlatGrads._cache.values())
        out["grad_allreduce"] = dict(us_per_step=1e3 * (ms_step - t_noar), calls_per_step=1,
                                     via=("symmetric memory, " + ("NVSwitch multimem" if fg[0].mc_ptr else "peer loads/stores"))
                                     if fg else "NCCL")
        out["bn_exchange"] = dict(us_per_step=1e3 * (t_noar - t_local),
                                  via="symmetric memory flags + peer loads" if par.PeerStats._cache else "NCCL")
        out["ms_per_step"] = dict(full=ms_step, no_grad_allreduce=t_noar, rank_local=t_local)
        out["note"] = ("graph replays of the same step without the gradient all-reduce / without any cross-rank call; "
                       "differences of max-over-ranks step times")
        return out

    pool_default = L.POOL_DENSE if args.pool == "dense" else L.POOL_FOLDED
    graph_note = None
    eager_run = None
    if args.eager:
        main_run = timed_region(pool_default, args.steps, max(args.warmup, 3), True)
    else:
        try:
            main_run = graph_region(pool_default, args.steps, max(args.warmup, 3), True)
            graph_note = "one cudaGraphLaunch per step (GraphedTrainStep)"
            eager_run = timed_region(pool_default, max(3, min(args.steps, 20)), 3, False)
            # attention / SCL / pooling-rest brackets (tags >= 2) are only recorded in the eager leg: per-step milliseconds
            # from there, normalised to this leg's step count when the shares are formed
            for tag in range(2, N_TAGS):
                main_run["prof"][tag] = list(eager_run["prof"].get(tag, []))
            main_run["prof_hi_steps"] = eager_run["prof_steps"]
        except Exception as e:  # capture refused (e.g. a collective that cannot be captured): fall back to eager launches
            graph_note = f"capture failed, eager launches timed instead: {type(e).__name__}: {str(e)[:200]}"
            torch.cuda.synchronize()
            main_run = timed_region(pool_default, args.steps, max(args.warmup, 3), True)
    ms_step, launches, prof, clk, final_loss = (main_run["ms_step"], main_run["launches"], main_run["prof"],
                                                main_run["clocks"], main_run["loss"])
    prof_ms_step = main_run.get("prof_ms_step", ms_step)
    prof_steps = max(1, main_run.get("prof_steps", 1))
    value = world * Bv / (ms_step / 1e3)
    dense_run = None
    if args.pool == "folded" and not args.no_dense and world == 1:
        dense_run = timed_region(L.POOL_DENSE, max(3, min(args.steps, 10)), 3, False)
        head_opts.pool_mode = pool_default

    # ---- untimed parity check on the bench's own inputs (dropout 0): CUDA path vs the CPU oracle -----------------------
    parity = None
    if not args.no_parity:
        parity = {}
        model0 = make_model(0.0)
        model0.run_options.pool_mode = pool_default
        names0 = [n for n, p in model0.named_parameters() if "backbone" not in n]
        params0 = [p for n, p in model0.named_parameters() if "backbone" not in n]

        def cuda_step0(tok, mk, sl, st, nv):
            for p in params0:
                p.grad = None
            e = model0.forward_tokens(tok, video_masks=mk, project=True)
            loss = algo.compute_sequence_loss(e.view(nv, 2, T, -1), sl, st, mk)["loss"]
            loss.backward()
            return e.detach(), loss.detach(), torch.cat([p.grad.reshape(-1) for p in params0]).double()

        e_d, loss_d, g_d = cuda_step0(tokens_dev, masks_d, seq_lens_d, steps_d, Bv)
        parity["cuda"] = dict(loss=float(loss_d), grad_norm=float(g_d.norm()))
        if world == 1:
            hc0 = head_cfg(0.0)
            t0 = time.perf_counter()
            torch.set_num_threads(os.cpu_count() or 1)
            Pr = {k: v.clone().requires_grad_(True) for k, v in O.init_params(hc0, seed=1).items()}
            tok_cpu = tokens_dev.float().cpu()
            emb, _ = O.head_forward(Pr, None, tok_cpu, masks, hc0, True)
            e_o, _ = O.proj_forward(Pr, None, emb, hc0, True)
            loss_o = O.scl_loss_dense(e_o.view(Bv, 2, T, -1), seq_lens, steps, masks)
            loss_o.backward()
            g_o = torch.cat([Pr[n].grad.reshape(-1) for n in names0]).double()
            rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())
            parity["oracle"] = dict(loss=float(loss_o), grad_norm=float(g_o.norm()),
                                    what="CPU oracle (fp32) on the same bf16-rounded tokens and parameters, dropout 0",
                                    seconds=time.perf_counter() - t0)
            parity["loss_rel"] = abs(float(loss_d) - float(loss_o)) / abs(float(loss_o))
            parity["emb_rel"] = rel(e_d, e_o.detach())
            parity["grad_rel"] = rel(g_d, g_o)
            parity["tolerance"] = 2e-2
            parity["ok"] = bool(parity["loss_rel"] < 2e-2 and parity["emb_rel"] < 2e-2 and parity["grad_rel"] < 2e-2)
            del tok_cpu, Pr
        else:
            # (1) every rank holds the same all-reduced gradient; (2) the distributed step (video shards, cross-rank
            # BatchNorm statistics, one flat gradient all-reduce) equals a single-process step on the concatenated batch
            g0 = g_d.clone()
            dist.broadcast(g0, src=0)
            dmax = (g_d - g0).abs().max()
            dist.all_reduce(dmax, op=dist.ReduceOp.MAX)
            parity["max_abs_grad_diff_across_ranks"] = float(dmax)
            lsum = loss_d.clone().double()
            dist.all_reduce(lsum)
            gather = lambda t: [torch.empty_like(t) for _ in range(world)]
            toks, mks, sls, sts = gather(tokens_dev), gather(masks_d), gather(seq_lens_d), gather(steps_d)
            dist.all_gather(toks, tokens_dev); dist.all_gather(mks, masks_d)
            dist.all_gather(sls, seq_lens_d); dist.all_gather(sts, steps_d)
            if rank == 0:
                model0.run_options.sync_bn = False
                model0.run_options.allreduce_grads = False
                # the per-rank loss is a mean over the rank's own valid frames and DDP averages the per-rank gradients
                # (mean of means): reproduce that by weighting each shard's loss with 1/world in one process
                for p in params0:
                    p.grad = None
                tok_all, mk_all = torch.cat(toks), torch.cat(mks)
                e_all = model0.forward_tokens(tok_all, video_masks=mk_all, project=True).view(world, Bv, 2, T, -1)
                losses = [algo.compute_sequence_loss(e_all[r], sls[r], sts[r], mks[r])["loss"] for r in range(world)]
                tot = sum(losses) / world
                tot.backward()
                g_one = torch.cat([p.grad.reshape(-1) for p in params0]).double()
                parity["single_process_concat_batch"] = dict(
                    loss_mean=float(tot), loss_mean_distributed=float(lsum) / world,
                    grad_rel=float((g_d - g_one).norm() / g_one.norm()),
                    what=f"rank 0 re-runs all {world * Bv} videos in one process (BatchNorm over the whole batch, no collectives)",
                    tolerance=2e-2,
                    note="two bf16-token / tf32-backward evaluations whose sums are split differently (one batch of "
                         f"{world * Bv} videos against {world} shards): 1e-3 at 2 ranks, 2.5e-3 at 8; the tolerance is the "
                         "path's stated 2e-2")
                # gradients must be bitwise identical on every rank; against the single-process run the path's own tolerance
                parity["ok"] = bool(parity["single_process_concat_batch"]["grad_rel"] < 2e-2 and float(dmax) == 0.0)
                del tok_all, e_all
            del toks
            torch.cuda.empty_cache()
            dist.barrier()
        del model0, params0
        torch.cuda.empty_cache()

    comm = None
    if world > 1 and not args.eager and graph_note and graph_note.startswith("one cudaGraphLaunch"):
        try:
            comm = comm_breakdown()
        except Exception as e:
            comm = dict(error=f"{type(e).__name__}: {str(e)[:200]}")
    ref_gpu = None
    if world == 1 and rank == 0 and not args.no_refgpu:
        ref_gpu = reference_gpu_run(dev, tokens_dev, seq_lens, steps, masks, max(3, min(args.steps, 10)))

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
        e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / k2
        e2e_value = world * Bv / (e2e_ms / 1e3)
        h2d = tokens_dev.numel() * 2 + seq_lens.numel() * 8 + steps.numel() * 8 + masks.numel() * 4
        d2h = 4

    if rank == 0:
        fl = flops_per_video()
        peaks = measured_peaks()
        mean = lambda xs: statistics.mean(xs) if xs else None
        prof_hi_steps = max(1, main_run.get("prof_hi_steps", prof_steps))
        per_step = lambda xs, n=None: (sum(xs) / (n or prof_steps)) if xs else 0.0      # ms per step spent inside that bracket
        F_frames = BV * T
        S_tok = E * T
        heads, Hh = 8, 256

        def traffic_of(fname):
            tp = os.path.join(ROOT, "profiles", fname)
            if os.path.exists(tp):
                with open(tp) as f:
                    d = json.load(f)
                if d.get("workload", "penn_cfg2") == args.workload:
                    return d.get("dram_bytes_per_launch")
            return None

        def dense_roofline(pr, ms):
            kv_ms = mean(pr[0])
            if not kv_ms:
                return None
            kv_flops = fl["kv_fwd"] * Bv
            achieved = kv_flops / (kv_ms * 1e-3) / 1e12
            r = dict(bound="tensor", kernel="gemm_tc_kernel<256,4,2> (K|V projection, forward)", achieved=achieved,
                     peak=peaks["tflops"], unit="TFLOP/s", frac=achieved / peaks["tflops"], traffic=traffic_of("kv_proj_fwd_traffic.json"),
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

        def pool_roofline(pr, ms):
            f_ms = mean(pr[0])
            if not f_ms:
                return None
            # algorithmic bytes of one launch: every token read once (bf16) + the pooled rows and the
            # attention maps written (fp32) + the folded query matrix read
            tok_bytes = F_frames * P * C_in * 2
            fwd_bytes = tok_bytes + F_frames * E * C_in * 4 + F_frames * E * P * 4 + E * C_in * 4
            achieved = fwd_bytes / (f_ms * 1e-3) / 1e9
            r = dict(bound="hbm", kernel="pool_foldw_fwd_kernel (folded entity pooling: one warp-specialised streaming pass over the bf16 tokens, mma.sync)",
                     achieved=achieved, peak=peaks["hbm"], unit="GB/s", frac=achieved / peaks["hbm"],
                     traffic=traffic_of("pool_fold_fwd_traffic.json"),
                     peak_source=f"{peaks['source']} hbm_gbs (copy, read+write)", frac_of_nominal_8000=achieved / 8000.0,
                     ms_per_launch=f_ms, launches_timed=len(pr[0]), bytes_per_launch=fwd_bytes,
                     entity_passes=(E + 7) // 8)
            if pr[1]:
                b_ms = mean(pr[1])
                bwd_bytes = tok_bytes + 2 * F_frames * E * C_in * 4 + F_frames * E * P * 4
                r["backward_pass"] = dict(kernel="pool_foldw_bwd_kernel", ms_per_launch=b_ms, bytes_per_launch=bwd_bytes,
                                          achieved=bwd_bytes / (b_ms * 1e-3) / 1e9,
                                          frac=bwd_bytes / (b_ms * 1e-3) / 1e9 / peaks["hbm"])
            return r

        def attn_roofline(pr, ms):
            f_ms = mean(pr[6])
            if not f_ms:
                return None
            # one launch = one encoder layer's attention over all views: S^2 scores per (view, head), two contractions of
            # d_k = 32 per score (QK^T, PV) forward, five in backward; one exp per score (forward) / two (backward: both
            # orientations are recomputed)
            scores = float(BV) * heads * S_tok * S_tok
            f_flops = scores * 2 * 2 * (Hh // heads)
            achieved = f_flops / (f_ms * 1e-3) / 1e12
            sm_mhz = (clk or {}).get("sm_mhz") or 1900.0
            mufu_peak = 148 * 16 * sm_mhz * 1e6        # ex2 per second: 16 / clk / SM
            r = dict(bound="tensor", kernel="temporal self-attention forward (one launch per encoder layer)", achieved=achieved,
                     peak=peaks["tflops_burst"], unit="TFLOP/s", frac=achieved / peaks["tflops_burst"], traffic=traffic_of("attention_fwd_traffic.json"),
                     peak_source=f"{peaks['source']} bf16_tflops (burst)", ms_per_launch=f_ms, launches_timed=len(pr[6]),
                     flops_per_launch=f_flops,
                     sfu=dict(exp_per_launch=scores, achieved_gexp_s=scores / (f_ms * 1e-3) / 1e9, peak_gexp_s=mufu_peak / 1e9,
                              frac=scores / (f_ms * 1e-3) / mufu_peak,
                              note="d_k = 32: 128 tensor FLOPs per exp -- the kernel is bound by MUFU ex2 (16/clk/SM), not by the tensor pipe"))
            if pr[7]:
                b_ms = mean(pr[7])
                b_flops = scores * 2 * 5 * (Hh // heads)
                r["backward_pass"] = dict(ms_per_launch=b_ms, flops_per_launch=b_flops, achieved=b_flops / (b_ms * 1e-3) / 1e12,
                                          frac=b_flops / (b_ms * 1e-3) / 1e12 / peaks["tflops_burst"],
                                          sfu_frac=2 * scores / (b_ms * 1e-3) / mufu_peak)
            return r

        def share(pr, ms):
            h = prof_hi_steps
            sh = dict(pool_stream_fwd=per_step(pr[0]) / ms, pool_stream_bwd=per_step(pr[1]) / ms,
                      pool_rest_fwd=(per_step(pr[2], h) + per_step(pr[4], h)) / ms,
                      pool_rest_bwd=(per_step(pr[3], h) + per_step(pr[5], h)) / ms,
                      attention_fwd=per_step(pr[6], h) / ms, attention_bwd=per_step(pr[7], h) / ms, scl=per_step(pr[8], h) / ms)
            sh["everything_else"] = max(0.0, 1.0 - sum(sh.values()))
            return sh

        if args.pool == "dense":
            roof = dense_roofline(prof, prof_ms_step)
        else:
            # the dominant kernel of THIS workload: the single launch with the largest share of the step
            pool_t = max(mean(prof[0]) or 0.0, mean(prof[1]) or 0.0)
            attn_t = max(mean(prof[6]) or 0.0, mean(prof[7]) or 0.0)
            roof = attn_roofline(prof, prof_ms_step) if attn_t > pool_t else pool_roofline(prof, prof_ms_step)
            if roof is not None:
                other = pool_roofline(prof, prof_ms_step) if attn_t > pool_t else attn_roofline(prof, prof_ms_step)
                roof["second_kernel"] = other
                roof["share_of_step"] = share(prof, prof_ms_step)
        if roof is not None:
            roof["timed_in"] = ("eager leg of this run: the un-split step issued kernel by kernel (CUDA events on the launching stream, "
                                f"{prof_ms_step:.3f} ms/step); share_of_step is relative to that leg") if "prof_ms_step" in main_run \
                else "the timed region of `value` (event-record nodes inside the replayed graph)"
        dense = None
        if dense_run is not None:
            dms = dense_run["ms_step"]
            dense = dict(value=world * Bv / (dms / 1e3), unit=UNIT, ms_per_step=dms, gpu_launches=dense_run["launches"],
                         whole_step_tflops=fl["total"] * Bv * world / (dms * 1e-3) / 1e12,
                         note="same step, entity pooling evaluated as written (K|V GEMM on tcgen05 + attention over K|V)",
                         roofline=dense_roofline(dense_run["prof"], dms))
        whole = fl["total"] * Bv * world / (ms_step * 1e-3) / 1e12
        cpu = None
        if not args.no_cpu and world == 1:
            threads = os.cpu_count() or 1
            sample = WL["cpu_sample"]     # a bounded part of the workload's batch per iteration (~10-20 s of CPU work)
            times, kind = cpu_step_time(sample, 4, threads)
            cms = statistics.mean(times)
            cpu = dict(value=sample / cms, unit=UNIT, cores=threads, kind=kind,
                       sample=f"{sample} videos of {WL['name']} (fp32, dropout 0), {len(times)} timed iterations after 1 warm-up")
        out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                   ms_per_step=ms_step, higher_is_better=True, scaling="strong" if WL.get("strong") else "weak", vs_baseline=None, dtype="bf16",
                   data="synthetic",
                   config=workload_config(world),
                   as_written_tflops=whole, gflop_per_video_as_written=fl["total"] / 1e9, loss=final_loss,
                   pooling=args.pool, launch_mode=("eager" if args.eager else "cuda_graph"), launch_note=graph_note,
                   sustained=main_run.get("sustained"),
                   eager=None if eager_run is None else dict(
                       value=world * Bv / (eager_run["ms_step"] / 1e3), unit=UNIT, ms_per_step=eager_run["ms_step"],
                       note="same step issued kernel by kernel through the drop-in Python API (no graph)"),
                   roofline=roof, parity=parity, dense_path=dense, reference_gpu=ref_gpu, cpu_baseline=cpu,
                   e2e=None if args.no_e2e else dict(
                       value=e2e_value, unit=UNIT, ms_per_step=e2e_ms, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=k2,
                       note="pinned host tokens -> HBM on a copy stream (double buffered) + loss.item() per step", numa=numa),
                   gpu_launches=launches, clocks=clk)
        if comm is not None:
            out["cross_rank"] = comm
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    global WL
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="penn_cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the host-to-device leg")
    ap.add_argument("--no-cpu", action="store_true", help="profiling runs: skip the CPU baseline sample")
    ap.add_argument("--no-parity", action="store_true", help="profiling runs: skip the untimed parity check against the oracle")
    ap.add_argument("--no-refgpu", action="store_true", help="skip the reference-modules-on-this-GPU baseline leg")
    ap.add_argument("--pool", default="folded", choices=["folded", "dense"],
                    help="entity pooling: folded (default product path) or dense (as written: K|V GEMM + attention)")
    ap.add_argument("--eager", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--no-overlap", action="store_true",
                    help="multi-GPU: one gradient all-reduce at the end instead of summing the chain's gradients beside the pooling backward")
    ap.add_argument("--reserve-sms", type=int, default=4, help="SMs the pooling backward leaves to the overlapped all-reduce")
    ap.add_argument("--global-videos", type=int, default=0,
                    help="strong scaling: total videos of the step, split evenly over the ranks (default: the workload's per-GPU shard on every rank)")
    ap.add_argument("--no-dense", action="store_true", help="skip the short as-written (dense pooling) comparison leg")
    args = ap.parse_args()
    WL = dict(WORKLOADS[args.workload])
    if args.global_videos:
        # strong scaling: a fixed global batch split over the ranks (BASELINE configs[2]: 256 Penn videos on 8 GPUs)
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.global_videos % world:
            raise SystemExit(f"--global-videos {args.global_videos} is not divisible by {world} ranks")
        WL["videos_per_gpu"] = args.global_videos // world
        WL["name"] = WL["name"].replace(f"bv{WORKLOADS[args.workload]['videos_per_gpu']}", f"bv{WL['videos_per_gpu']}")
        WL["baseline_config"] += f"; global batch fixed at {args.global_videos} videos (strong scaling)"
        WL["cpu_sample"] = min(WL["cpu_sample"], WL["videos_per_gpu"])
        WL["strong"] = True
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
