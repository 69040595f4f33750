"""CPU: the host-side mirror of the reference's models/ + algos/ interface (construction, names, init, ckpt)."""
import copy
import os

import pytest
import torch
import torch.nn as nn

from oracle import mvf_oracle as O
from oracle import ref_shim as R
from video_rep_learning_b200 import engine
from video_rep_learning_b200.algos import SCL, get_algo
from video_rep_learning_b200.config import Cfg, mvf_cfg
from video_rep_learning_b200.models import (MLPHead, MultiEntityTransformerEmbModel, TransformerModel, build_model,
                                            load_checkpoint, save_checkpoint)


class DummyBackbone(nn.Module):
    """frames [n,3,H,W] -> (tokens [n,1+P,C], cls [n,C]); stands in for the frozen timm ViT (upstream producer)."""

    def __init__(self, c_out=48, patch=56):
        super().__init__()
        self.proj = nn.Conv2d(3, c_out, patch, patch)

    def forward(self, x):
        t = self.proj(x).flatten(2).transpose(1, 2)
        cls = t.mean(1, keepdim=True)
        return torch.cat([cls, t], 1), cls[:, 0]


def small_cfg(**kw):
    base = dict(c_in=48, num_frames=8, entities=3, capacity=1, emb=16, hidden=32, d_ff=64, heads=4, layers=2,
                fc_layers=((64, True), (64, True)), projection_size=16, pool_channels=32)
    base.update(kw)
    return mvf_cfg(**base)


def test_state_dict_keys_match_reference_names():
    cfg = small_cfg()
    model = build_model(cfg, backbone=DummyBackbone())
    hc = O.HeadCfg(c_in=48, n_entities=3, pool_channels=32, fc_channels=(64, 64), hidden=32, d_ff=64, n_heads=4,
                   n_layers=2, emb=16, proj=16, train_frames=8)
    sd = {k: v for k, v in model.state_dict().items() if not k.startswith("backbone")}
    want = dict(O.param_shapes(hc))
    for pre, ch in zip(O.bn_buffer_names(hc), list(hc.fc_channels) + [hc.proj]):
        want[pre + ".running_mean"] = (ch,)
        want[pre + ".running_var"] = (ch,)
        want[pre + ".num_batches_tracked"] = ()
    assert set(sd.keys()) == set(want.keys())
    for k, shp in want.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    # canonical (C ABI) order == module lookup order
    assert model.embed.head_param_names() == [k[len("embed."):] for k in O.param_shapes(hc) if k.startswith("embed.")]
    assert [p.shape for p in model.ssl_projection.proj_params()] == [torch.Size(s) for k, s in O.param_shapes(hc).items()
                                                                     if k.startswith("ssl_projection.")]
    # optimizer contract (utils/optimizer.py:29-42 skips names containing 'backbone')
    assert all(not p.requires_grad for p in model.backbone.parameters())
    assert model.embedding_size == 16 and hasattr(model, "res_finetune")


def test_encoder_layers_start_identical_like_reference_clone():
    head = MultiEntityTransformerEmbModel(small_cfg(layers=3))
    l0 = head.video_encoder.enc_layers[0].state_dict()
    for l in (1, 2):
        for k, v in head.video_encoder.enc_layers[l].state_dict().items():
            assert torch.equal(v, l0[k])


@pytest.mark.skipif(not R.available(), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("yml_like", ["penn", "fg99"])
def test_seeded_init_identical_to_reference(yml_like):
    """Same constructor order -> same RNG stream -> bit-identical initial parameters (checkpoints interchange)."""
    ref = R.load_reference()
    yml = "penn_mvf.yml" if yml_like == "penn" else "fg99_mvf.yml"
    rcfg = R.reference_cfg(yml, c_in=96, T=8)
    if yml_like == "penn":
        ours = mvf_cfg(c_in=96, num_frames=8)
    else:
        ours = mvf_cfg(c_in=96, num_frames=8, entities=6, capacity=6, emb=256, final="avg", smart_feats="9,10,11")
    import contextlib, io
    torch.manual_seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        rhead = ref.MultiEntityTransformerEmbModel(rcfg)
        rproj = ref.MLPHead(rcfg)
    torch.manual_seed(1)
    head = MultiEntityTransformerEmbModel(ours)
    proj = MLPHead(ours)
    a, b = rhead.state_dict(), head.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k
    a, b = rproj.state_dict(), proj.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # and a reference state_dict loads strictly
    head.load_state_dict(rhead.state_dict(), strict=True)


def test_unsupported_branches_fail_loudly():
    for key, val in (("SMART_DYNAMIC_TOKENS", 2), ("VAL_PASS", True), ("SMART_DISJOINT", True), ("SMART_LN_KEYS", True)):
        cfg = small_cfg()
        cfg.MODEL.EMBEDDER_MODEL[key] = val
        with pytest.raises(NotImplementedError):
            MultiEntityTransformerEmbModel(cfg)
    # FIXED_WIDTH_BASELINE (configs_mvf/ablate_dinoB8_fwb{3,5}.yml) is built: FWBPooling.lin_conv replaces the cross-attention
    cfg = small_cfg()
    cfg.MODEL.EMBEDDER_MODEL.FIXED_WIDTH_BASELINE = True
    fwb = MultiEntityTransformerEmbModel(cfg)
    keys = set(fwb.state_dict().keys())
    assert "pooling.lin_conv.weight" in keys and not any("cross_att" in k for k in keys)
    assert fwb.spec.pool_kind == "fwb" and fwb.spec.cls_dim > 0
    assert tuple(fwb.pooling.lin_conv.weight.shape) == (fwb.spec.pool_channels * fwb.spec.n_entities, fwb.spec.cls_dim)
    assert fwb.head_param_names()[:2] == ["pooling.lin_conv.weight", "pooling.lin_conv.bias"]
    with pytest.raises(NotImplementedError):
        MultiEntityTransformerEmbModel(small_cfg(one_hot="enc"))
    cfg = small_cfg()
    cfg.MODEL.EMBEDDER_MODEL.FUSION_TYPE = "late"
    with pytest.raises(NotImplementedError):
        TransformerModel(cfg, backbone=DummyBackbone())
    cfg = small_cfg(negative_type="all")
    with pytest.raises(NotImplementedError):
        SCL(cfg)
    cfg = small_cfg()
    cfg.TRAINING_ALGO = "tcc"
    with pytest.raises(ValueError):
        get_algo(cfg)


def test_submodules_are_containers_not_silent_torch_paths():
    head = MultiEntityTransformerEmbModel(small_cfg())
    with pytest.raises(RuntimeError, match="parameter container"):
        head.video_encoder(torch.zeros(1, 4, 32))
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        head(torch.zeros(2, 8, 48, 3, 3))          # CPU tensor: must not fall back to PyTorch math


def test_head_spec_from_cfg_defaults():
    cfg = small_cfg()
    del cfg.MODEL.EMBEDDER_MODEL["SMART_POOL_CHANNELS"]
    head = MultiEntityTransformerEmbModel(cfg)
    assert head.spec.pool_channels == 384 and head.spec.fc_channels == (64, 64) and head.spec.final == "one"
    assert head.spec == engine.HeadSpec(c_in=48, n_entities=3, pool_channels=384, fc_channels=(64, 64), hidden=32, d_ff=64,
                                        n_heads=4, n_layers=2, emb=16, proj=16, one_hot="pool", final="one", train_frames=8,
                                        drop_p=0.1)


def test_to_token_major_is_the_inverse_of_the_reference_permute():
    x = torch.arange(2 * 3 * 5 * 2 * 2, dtype=torch.float32).view(2, 3, 5, 2, 2)        # [BV,T,C,h,w]
    t = MultiEntityTransformerEmbModel.to_token_major(x)
    assert t.shape == (2, 3, 4, 5)
    assert torch.equal(t[1, 2, 3], x[1, 2, :, 1, 1])
    assert torch.equal(R.tokens_to_nchw(t), x)


def test_checkpoint_roundtrip(tmp_path):
    cfg = small_cfg()
    cfg.LOGDIR = str(tmp_path)
    model = build_model(cfg, backbone=DummyBackbone())
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3)
    save_checkpoint(cfg, model, opt, 3)
    assert os.path.exists(os.path.join(str(tmp_path), "checkpoints", "checkpoint_epoch_00003.pth"))
    model2 = build_model(copy.deepcopy(cfg), backbone=DummyBackbone())
    opt2 = torch.optim.Adam([p for p in model2.parameters() if p.requires_grad], lr=1e-3)
    assert load_checkpoint(cfg, model2, opt2) == 4
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k


def test_cfg_container_semantics():
    cfg = Cfg({"A": {"B": 1}})
    assert cfg.A.B == 1 and "B" in cfg.A and "C" not in cfg.A
    cfg.A.C = {"D": 2}
    assert cfg.A.C.D == 2
    with pytest.raises(AttributeError):
        _ = cfg.missing


def test_optimizer_tail_and_graph_step_have_no_cpu_path():
    """optim.FusedAdam / graph.GraphedTrainStep: host-side contract only (the arithmetic is tested on the GPU box) -- CPU
    tensors are refused loudly, mixed hyper-parameters across groups are refused, construct_optimizer mirrors
    utils/optimizer.py:60-73."""
    import pytest
    import torch
    from video_rep_learning_b200.graph import GraphedTrainStep
    from video_rep_learning_b200.optim import FusedAdam, construct_optimizer

    cfg = small_cfg(drop=0.0)
    w = torch.nn.Parameter(torch.zeros(4, 4))
    w.grad = torch.ones(4, 4)
    opt = FusedAdam([w], lr=1e-3, max_grad_norm=10.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        opt.step()
    a, b = torch.nn.Parameter(torch.zeros(2)), torch.nn.Parameter(torch.zeros(2))
    a.grad, b.grad = torch.ones(2), torch.ones(2)
    mixed = FusedAdam([{"params": [a], "lr": 1e-3}, {"params": [b], "lr": 1e-2}])
    with pytest.raises(NotImplementedError):
        mixed.step()
    from video_rep_learning_b200.models import build_model
    m = build_model(cfg, backbone=DummyBackbone())
    o = construct_optimizer(m, cfg)
    n_head = sum(1 for n, p in m.named_parameters() if "backbone" not in n)
    assert sum(len(g["params"]) for g in o.param_groups) == n_head
    assert o.max_grad_norm == float(cfg.OPTIMIZER.GRAD_CLIP) and not o.adamw
    assert o.param_groups[0]["weight_decay"] == cfg.OPTIMIZER.WEIGHT_DECAY
    from video_rep_learning_b200.algos import get_algo
    with pytest.raises(RuntimeError, match="CUDA"):
        GraphedTrainStep(m, get_algo(cfg), 2, 8, 9, 48, device=torch.device("cpu"))


def test_eval_chunking_matches_reference_arithmetic():
    """evaluate.py:45-55 (INT): equal chunks of ceil(L / ceil(L / FRAMES_PER_BATCH)) frames, clamped context windows."""
    from video_rep_learning_b200.evaluate import chunk_steps
    import math
    for L, fpb in ((1500, 2000), (1500, 800), (2001, 1000), (7, 3), (1, 5)):
        chunks = chunk_steps(L, fpb)
        nb = int(math.ceil(float(L) / fpb))
        per = int(math.ceil(float(L) / nb))
        assert len(chunks) == nb and torch.equal(torch.cat(chunks), torch.arange(L))
        assert all(len(c) == min(L - i * per, per) for i, c in enumerate(chunks))
    c = chunk_steps(5, 10, num_contexts=2, context_stride=3)[0]
    assert c.tolist() == [0, 0, 0, 1, 0, 2, 0, 3, 1, 4]            # steps - 3 clamped at 0, interleaved with the steps


class _FakeViT(nn.Module):
    """Just enough of a timm ViT for FeatureExtractor: .blocks, .patch_embed.num_patches, .forward_features."""

    def __init__(self, width=8, depth=4, patches=9):
        super().__init__()
        self.embed = nn.Linear(3, width)
        self.blocks = nn.ModuleList([nn.Linear(width, width) for _ in range(depth)])
        self.patch_embed = nn.Module()
        self.patch_embed.num_patches = patches
        self.patches = patches

    def forward_features(self, x):                       # x [n, 3, H, W] -> tokens [n, 1 + P, width]
        n = x.shape[0]
        t = self.embed(x.flatten(2)[:, :, :1 + self.patches].transpose(1, 2))
        for b in self.blocks:
            t = torch.tanh(b(t))
        return t


def test_feature_extractor_writes_hooked_tokens_straight_into_the_token_buffer():
    """the hooked ViT block outputs land in the head's token buffer (CLS dropped, channel slices,
    dtype cast) without a concatenated intermediate; same values as the reference's cat + drop-CLS + reshape."""
    from video_rep_learning_b200.models.transformer import FeatureExtractor
    torch.manual_seed(0)
    vit = _FakeViT()
    fx = FeatureExtractor(vit, [1, 3])
    cfg = small_cfg(c_in=16)
    cfg.MODEL.BASE_MODEL.FRAMES_PER_BATCH = 4
    model = build_model(cfg, backbone=fx)
    x = torch.randn(1, 6, 3, 4, 4)                        # one video, 6 frames -> two chunks of 4 + 2 frames
    # reference-style path: cat of the hooked outputs, CLS dropped afterwards
    toks, _cls = fx(x[0])
    want = toks[:, 1:, :].reshape(1, 6, 9, 16)
    buf = torch.full((1, 6, 9, 16), float("nan"), dtype=torch.bfloat16)
    out = model.backbone_tokens(x, out=buf)
    assert out.data_ptr() == buf.data_ptr()               # the producer wrote where the consumer reads
    assert torch.equal(buf.float(), want.bfloat16().float())
    assert fx._dest is None and not fx._feats              # nothing retained, hooks unbound again
    # without a buffer it allocates one of the backbone's dtype and still goes through the hooks
    out2 = model.backbone_tokens(x)
    assert torch.equal(out2, want)


def test_symmetric_gradient_buffer_piece_arithmetic():
    """parallel.PeerFlatGrads.range_of: which tensors count as (16-byte aligned) pieces of the symmetric flat buffer and
    which float range the all-reduce kernel is given for them -- host arithmetic only, no symmetric memory needed."""
    from video_rep_learning_b200 import parallel
    obj = object.__new__(parallel.PeerFlatGrads)
    backing = torch.zeros(1024 + 4)
    obj.flat = backing[:1022]            # 1022 gradient floats, padded to 1024 inside the buffer
    obj.elems = 1024
    assert obj.range_of(obj.flat) == (0, 1024)                      # the whole buffer takes its zero padding along
    assert obj.range_of(obj.flat[:512]) == (0, 512)
    assert obj.range_of(obj.flat[512:]) == (512, 512)               # a tail piece reaches the padded end
    assert obj.range_of(obj.flat[2:514]) is None                    # not 16-byte aligned
    assert obj.range_of(obj.flat[:510]) is None                     # not a multiple of four floats
    assert obj.range_of(torch.zeros(512)) is None                   # somebody else's memory
    assert obj.range_of(obj.flat[:512].double()) is None
    assert obj.owns(obj.flat[512:]) and not obj.owns(torch.zeros(4))


def test_every_shipped_mvformer_config_constructs_unmodified():
    """configs_mvf/*.yml of the reference, read unchanged: the 15 files that select the MV-Former head (FUSION_TYPE: smart)
    construct with a caller-supplied producer -- OUT_CHANNEL derived from NETWORK as transformer.py:40-56,119-133 does -- and map
    to head shapes inside the supported envelope; the 5 late-fusion CARL ablations are rejected by name."""
    import glob
    from oracle import ref_shim
    from video_rep_learning_b200.config import load_yaml
    from video_rep_learning_b200.models import build_model
    from video_rep_learning_b200.models.mvformer import head_spec_from_cfg
    root = os.path.join(ref_shim.REF_ROOT, "configs_mvf")
    files = sorted(glob.glob(os.path.join(root, "*.yml")))
    if len(files) < 20:
        pytest.skip("reference configs not available (neither /root/reference nor baseline/_ref)")
    smart, late = {}, []
    for y in files:
        cfg = load_yaml(y)
        em = cfg.MODEL.EMBEDDER_MODEL
        if "FUSION_TYPE" in em and em.FUSION_TYPE == "smart":
            model = build_model(cfg, backbone=DummyBackbone(16))
            hs = head_spec_from_cfg(cfg)
            assert model.embed.embedding_size == hs.emb
            smart[os.path.basename(y)] = (hs.c_in, hs.n_entities, hs.fc_channels, hs.emb, hs.final, hs.one_hot, hs.pool_kind,
                                          cfg.TRAIN.NUM_FRAMES)
        else:
            with pytest.raises(NotImplementedError, match="FUSION_TYPE"):
                build_model(cfg, backbone=DummyBackbone(16))
            late.append(os.path.basename(y))
    assert len(smart) == 15 and len(late) == 5
    assert smart["penn_mvf.yml"] == (2304, 3, (512, 512), 128, "one", "pool", "lstp", 80)
    assert smart["fg99_mvf.yml"] == (2304, 6, (1536, 1536), 256, "avg", "pool", "lstp", 240)
    assert smart["pouring_mvf.yml"][:2] == (768, 3) and smart["ablate_rn50_lstp5.yml"][:2] == (2048, 5)
    assert smart["ablate_dinoB8_fwb5.yml"][6] == "fwb" and smart["ablate_dinoB8_fwb5.yml"][1] == 5
    for name, (c_in, E, fc, emb, final, one_hot, kind, T) in smart.items():
        assert c_in % 16 == 0 and 1 <= E <= 16 and emb in (128, 256) and T * E <= 16384, name


def test_get_embeddings_dataset_mirrors_reference_contract():
    """evaluate.get_embeddings_dataset (evaluate.py:27-81): same arguments and dictionary; chunks are concatenated in order and
    frames with a negative label dropped.  A stand-in model (frame index -> embedding) keeps this on the host."""
    from video_rep_learning_b200 import evaluate as E
    from video_rep_learning_b200.config import mvf_cfg

    class Stub(nn.Module):
        def forward(self, x, num_steps):          # x: [1, n, 1] frame indices -> [1, n, 2]
            assert x.shape[1] == num_steps
            return torch.cat([x.float(), 2 * x.float()], dim=-1)

    cfg = mvf_cfg(num_frames=8)
    cfg.EVAL = dict(FRAMES_PER_BATCH=4)
    cfg.DATA = dict(NUM_CONTEXTS=1, CONTEXT_STRIDE=1)
    L = 10
    video = torch.arange(L).view(1, L, 1)
    labels = torch.tensor([[0, 1, -1, 1, 2, 2, -1, 3, 3, 3]])
    loader = [(video, labels, torch.tensor([L]), torch.arange(L).view(1, L), None, ["vid0"])]
    ds = E.get_embeddings_dataset(cfg, Stub(), loader)
    assert set(ds) == {"embs", "labels", "seq_lens", "input_lens", "steps", "names"}
    keep = [0, 1, 3, 4, 5, 7, 8, 9]
    assert ds["embs"][0].shape == (8, 2) and ds["embs"][0][:, 0].tolist() == [float(k) for k in keep]
    assert ds["labels"][0].tolist() == [0, 1, 1, 2, 2, 3, 3, 3]
    assert ds["seq_lens"] == [L] and ds["input_lens"] == [L] and ds["names"] == ["vid0"] and ds["steps"][0].tolist() == list(range(L))
