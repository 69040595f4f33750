"""Import the reference's own hot-path modules.  TEST / BASELINE INFRASTRUCTURE ONLY.

Two locations are tried: /root/reference/CARL_MVF (build container) and, on the GPU box where that tree does not
exist, `baseline/_ref/CARL_MVF` -- an UNMODIFIED copy of the handful of reference files on the hot path that
`__graft_entry__.build()` places there (git-ignored, shipped by gpurun; never part of the repository history).
It is used by tests/golden/make_golden.py (to generate the committed golden vectors), by the `not gpu` test that
re-checks the oracle against the live reference, and by the baseline arms of bench.py (`--impl reference`,
`--impl reference-gpu`), which time the reference's own modules.  Nothing in the product path, the `-m gpu` parity
tests or smoke() imports this file.

A plain `import models` fails (timm / iopath / easydict are not installed), so
the four files on the hot path are loaded by path under stub parent packages.
"""
from __future__ import annotations

import importlib.util
import logging
import os
import sys
import types

import torch
import yaml

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VENDORED_ROOT = os.path.join(_REPO, "baseline", "_ref", "CARL_MVF")
# files of the reference that the hot path (and its samplers / configs) needs; copied verbatim by vendor_reference()
VENDOR_FILES = ("models/utils.py", "models/mvformer.py", "models/resnet_c2d.py", "algos/scl.py",
                "datasets/dataset_splits.py", "datasets/data_augment.py", "datasets/penn_action.py", "datasets/finegym.py",
                "datasets/pouring.py") + tuple(
    "configs_mvf/" + n for n in (
        "penn_mvf.yml", "pouring_mvf.yml", "fg99_mvf.yml", "fg288_mvf.yml", "k400_mvf.yml", "k400_penn_mvf.yml",
        "ablate_dinoB8_fwb3.yml", "ablate_dinoB8_fwb5.yml", "ablate_dinoB8_lstp1.yml", "ablate_dinoB8_lstp3.yml",
        "ablate_dinoB8_lstp5.yml", "ablate_dinoB8_multi_lstp1.yml", "ablate_dinoB8_multi_lstp5.yml", "ablate_dinoB8_avg.yml",
        "ablate_dinoB8_cls.yml", "ablate_dinoB8_max.yml", "ablate_rn50_lstp1.yml", "ablate_rn50_lstp3.yml",
        "ablate_rn50_lstp5.yml", "ablate_rn50_max.yml"))


def _default_root() -> str:
    env = os.environ.get("MVF_REFERENCE_ROOT")
    if env:
        return env
    live = "/root/reference/CARL_MVF"
    return live if os.path.isfile(os.path.join(live, "models", "mvformer.py")) else VENDORED_ROOT


REF_ROOT = _default_root()


def vendor_reference(src_root: str = "/root/reference/CARL_MVF") -> bool:
    """Copy the unmodified hot-path files of the reference into baseline/_ref (git-ignored) so that the baseline arms of
    bench.py can run the reference's own code on the GPU box.  No-op when the reference tree is absent."""
    import shutil
    if not os.path.isfile(os.path.join(src_root, "models", "mvformer.py")):
        return False
    for rel in VENDOR_FILES:
        src, dst = os.path.join(src_root, rel), os.path.join(VENDORED_ROOT, rel)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(dst) or os.path.getmtime(dst) < os.path.getmtime(src) or os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)
    return True


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "mvformer.py"))


class AttrDict(dict):
    """EasyDict stand-in: attribute access plus `'KEY' in cfg.X` (the reference probes keys that way)."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = __setitem__


_loaded = {}


def _stub_pkg(name, sub):
    m = types.ModuleType(name)
    m.__path__ = [os.path.join(REF_ROOT, sub)]
    sys.modules[name] = m
    return m


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load_reference():
    """Returns a namespace with the reference classes: MultiEntityTransformerEmbModel, MLPHead, SCL,
    PennAction/FineGym samplers (unbound functions), attention helpers."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    saved = {k: sys.modules.get(k) for k in ("models", "utils", "datasets", "utils.logging")}
    _stub_pkg("models", "models")
    u = _stub_pkg("utils", "utils")
    _stub_pkg("datasets", "datasets")
    ul = types.ModuleType("utils.logging")
    ul.get_logger = logging.getLogger
    sys.modules["utils.logging"] = ul
    u.logging = ul
    mu = _load("models.utils", "models/utils.py")
    mv = _load("models.mvformer", "models/mvformer.py")
    _load("datasets.dataset_splits", "datasets/dataset_splits.py")
    rc = _load("models.resnet_c2d", "models/resnet_c2d.py")
    scl = _load("ref_algos_scl", "algos/scl.py")
    # mvformer.py:145 calls torch.eye(device=x.get_device()); get_device() is -1 on CPU tensors -> hand back the device
    # object for CPU tensors (CUDA tensors keep their ordinal), so that the same modules run on either side
    if not getattr(torch.Tensor.get_device, "_mvf_patched", False):
        _orig_get_device = torch.Tensor.get_device

        def _get_device(self):
            return self.device if self.device.type == "cpu" else _orig_get_device(self)

        _get_device._mvf_patched = True
        torch.Tensor.get_device = _get_device
    _loaded.update(dict(utils=mu, mvformer=mv, resnet_c2d=rc, scl=scl,
                        MultiEntityTransformerEmbModel=mv.MultiEntityTransformerEmbModel,
                        MLPHead=rc.MLPHead, SCL=scl.SCL))
    # samplers: need torchvision.io.read_video to exist at import time
    try:
        import torchvision.io as tio
        if not hasattr(tio, "read_video"):
            tio.read_video = None
        import numpy as np
        if not hasattr(np, "int"):
            np.int = int  # penn_action.py:63 uses the removed alias at import-irrelevant sites
        # decord / cv2 are not installed; they are only used for video decode, never by sample_frames
        dl = types.ModuleType("utils.decord_loader")
        dl.decord_load = None
        sys.modules["utils.decord_loader"] = dl
        for missing in ("cv2", "psutil"):
            if missing not in sys.modules:
                try:
                    __import__(missing)
                except Exception:
                    sys.modules[missing] = types.ModuleType(missing)
        _load("datasets.data_augment", "datasets/data_augment.py")
        pa = _load("datasets.penn_action", "datasets/penn_action.py")
        _loaded["PennAction"] = pa.PennAction
        fg = _load("datasets.finegym", "datasets/finegym.py")
        _loaded["FineGym"] = fg.Finegym
        po = _load("datasets.pouring", "datasets/pouring.py")
        _loaded["Pouring"] = po.Pouring
    except Exception as e:  # pragma: no cover - sampler import is best effort
        _loaded["sampler_import_error"] = repr(e)
    return types.SimpleNamespace(**_loaded)


def reference_cfg(yml: str = "penn_mvf.yml", **over):
    """cfg = yaml of configs_mvf/<yml> as AttrDict with the runtime-set keys filled in."""
    with open(os.path.join(REF_ROOT, "configs_mvf", yml)) as f:
        cfg = AttrDict(yaml.safe_load(f))
    cfg.MODEL.BASE_MODEL.OUT_CHANNEL = over.pop("c_in", 2304)       # transformer.py:44-54,90 set this at run time
    cfg.TRAIN.NUM_FRAMES = over.pop("T", cfg.TRAIN.NUM_FRAMES)
    em = cfg.MODEL.EMBEDDER_MODEL
    for k, v in over.items():
        if k in ("PROJECTION_SIZE",):
            cfg.MODEL[k] = v
        elif k in ("NEGATIVE_TYPE", "SOFTMAX_TEMPERATURE", "LABEL_VARIENCE"):
            cfg.SCL[k] = v
        else:
            em[k] = v
    return cfg


def cfg_from_headcfg(hc, yml: str = "penn_mvf.yml"):
    """Build a reference cfg whose head has exactly the hyper-parameters of an oracle HeadCfg."""
    assert len(set(hc.fc_channels)) == 1
    cfg = reference_cfg(yml, c_in=hc.c_in, T=hc.train_frames,
                        SMART_TOKENS=hc.n_entities, SMART_POOL_CHANNELS=hc.pool_channels,
                        CAPACITY_SCALAR=1, FC_LAYERS=[[ch, True] for ch in hc.fc_channels],
                        HIDDEN_SIZE=hc.hidden, D_FF=hc.d_ff, NUM_HEADS=hc.n_heads, NUM_LAYERS=hc.n_layers,
                        EMBEDDING_SIZE=hc.emb, SMART_ONE_HOT=hc.one_hot, SMART_FINAL=hc.final,
                        FC_DROPOUT_RATE=hc.drop_p, PROJECTION_SIZE=hc.proj)
    # SMART_FEATS only matters for the width of the CLS embedding (dynamic tokens / FIXED_WIDTH_BASELINE): d_dyn_in =
    # OUT_CHANNEL / number of feature layers (mvformer.py:441-447); a single layer keeps d_dyn_in == c_in.
    cfg.MODEL.EMBEDDER_MODEL.SMART_FEATS = "11"
    if getattr(hc, "pool_kind", "lstp") == "fwb":
        cfg.MODEL.EMBEDDER_MODEL.FIXED_WIDTH_BASELINE = True
        n = hc.c_in // hc.cls_dim
        assert n * hc.cls_dim == hc.c_in
        cfg.MODEL.EMBEDDER_MODEL.SMART_FEATS = ",".join(str(11 - i) for i in range(n)) if n > 1 else "11"
    return cfg


def build_reference_modules(hc, params, yml: str = "penn_mvf.yml"):
    """Instantiate the reference head + MLPHead + SCL and load the oracle's parameter dict into them."""
    ref = load_reference()
    cfg = cfg_from_headcfg(hc, yml)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        head = ref.MultiEntityTransformerEmbModel(cfg)
        proj = ref.MLPHead(cfg)
    algo = ref.SCL(cfg)
    hs = {k[len("embed."):]: v.clone() for k, v in params.items() if k.startswith("embed.")}
    ps = {k[len("ssl_projection."):]: v.clone() for k, v in params.items() if k.startswith("ssl_projection.")}
    missing, unexpected = head.load_state_dict(hs, strict=False)
    assert not unexpected, unexpected
    assert all(("running_" in m or "num_batches" in m) for m in missing), missing
    missing, unexpected = proj.load_state_dict(ps, strict=False)
    assert not unexpected, unexpected
    return cfg, head, proj, algo


def tokens_to_nchw(tokens):
    """[BV,T,P,C] token-major -> the NCHW view [BV,T,C,h,w] the reference head expects (transformer.py:203-213)."""
    BV, T, P, C = tokens.shape
    h = int(round(P ** 0.5))
    assert h * h == P
    return tokens.transpose(2, 3).reshape(BV, T, C, h, h).contiguous()
