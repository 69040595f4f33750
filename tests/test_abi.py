"""CPU: the C-ABI library loads, exports every symbol include/mvf_b200.h declares, and its host-side
bookkeeping (parameter table, layouts, error reporting) behaves -- no compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest
import torch

from oracle import mvf_oracle as O
from tests import helpers as H
from video_rep_learning_b200 import _lib as L
from video_rep_learning_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mvf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/mvf_b200.h but not exported"
    assert set(declared) == set(L.EXPORTED_SYMBOLS), set(declared) ^ set(L.EXPORTED_SYMBOLS)
    assert lib.mvf_version() == 3


def test_struct_layout_matches_header():
    # 22 int32 (incl. fc[4]) + 4 float + 1 uint64, naturally aligned
    assert C.sizeof(L.HeadDesc) == 4 * 25 + 4 * 4 + 4 + 8 or C.sizeof(L.HeadDesc) % 8 == 0
    d = L.HeadDesc()
    d.seed = 2 ** 63 + 5
    assert d.seed == 2 ** 63 + 5


@pytest.mark.parametrize("kw", [dict(c_in=2304), dict(c_in=1152), dict(c_in=2304, n_entities=6, fc_channels=(1536, 1536), emb=256, final="avg"),
                                dict(c_in=48, final="lin", n_entities=2, one_hot="none"), dict(c_in=48, fc_channels=())])
def test_param_table_is_reference_state_dict_order(kw):
    hc = O.HeadCfg(**kw)
    plan = engine.Plan.get(H.spec_from_headcfg(hc), 4, 20, 196, L.MVF_F32, True, True, 1, 0)
    shapes = O.param_shapes(hc)
    assert plan.param_names == list(shapes.keys())
    for (r, c), shp in zip(plan.param_shapes, shapes.values()):
        n = 1
        for s in shp:
            n *= s
        assert r * c == n
    assert plan.bn_names == O.bn_buffer_names(hc)
    assert plan.gpack_elems >= sum(r * c for r, c in plan.param_shapes)
    assert plan.save_bytes > 0 and plan.ws_bytes > 0 and plan.proj_save_bytes > 0
    # the regions of the pooling parameters (Q_s, Q_s_b, W_k | W_v, b_k | b_v) lead the flat gradient buffer: what the last
    # backward phase writes, and what an overlapped all-reduce must leave for the end
    pool_params = hc.n_entities * hc.pool_channels + hc.pool_channels + 2 * hc.pool_channels * hc.c_in + 2 * hc.pool_channels
    assert pool_params <= plan.gpack_pool_elems < plan.gpack_elems
    assert plan.gpack_pool_elems - pool_params < 4 * 64 + 2 * hc.pool_channels * 8      # only alignment padding in between


def test_layout_lookup_and_bn_stat_regions():
    hc = O.HeadCfg(c_in=2304)
    plan = engine.Plan.get(H.spec_from_headcfg(hc), 64, 20, 196, L.MVF_BF16, True, True, 1, 0, L.POOL_DENSE)
    off, rows, cols, ld, dt = plan.lookup("kv")
    assert (rows, cols, ld, dt) == (64 * 20 * 196, 768, 768, 1)
    folded = engine.Plan.get(H.spec_from_headcfg(hc), 64, 20, 196, L.MVF_BF16, True, True, 1, 0)      # AUTO -> folded
    with pytest.raises(RuntimeError, match="no region"):
        folded.lookup("kv")                    # K|V are never materialised
    off, rows, cols, ld, dt = folded.lookup("px")
    assert (rows, cols, dt) == (64 * 20 * 3, 2304, 0)
    assert folded.save_bytes < plan.save_bytes - 64 * 20 * 196 * 768 * 2 + 64 * 20 * 3 * 2304 * 4 + (1 << 20)
    off, rows, cols, ld, dt = plan.lookup("h0")
    assert (cols, ld) == (387, 392)          # one-hot columns, padded to a 16-byte row for TMA
    off, rows, cols, ld, dt = plan.lookup("g.w.kv")
    assert (rows, cols) == (768, 2304) and dt == 0
    with pytest.raises(RuntimeError, match="no region"):
        plan.lookup("does.not.exist")
    buf = torch.zeros(plan.save_bytes, dtype=torch.uint8)
    s = plan.bn_stat(buf, 0, False)
    assert s.dtype == torch.float64 and s.numel() == 2 * 512


def test_errors_are_reported_not_swallowed():
    lib = L.lib()
    d = L.HeadDesc()            # all zeros: invalid
    assert lib.mvf_num_params(C.byref(d)) == -1
    assert "bad input shape" in L.last_error()
    hc = O.HeadCfg(c_in=50)      # not a multiple of 8 -> illegal for bf16/TMA
    with pytest.raises(RuntimeError, match="multiples of 8"):
        engine.Plan(H.spec_from_headcfg(hc), 2, 4, 9, L.MVF_BF16, True, False, 1, 0)
    with pytest.raises(NotImplementedError):
        engine.Plan(H.spec_from_headcfg(O.HeadCfg(c_in=48, fc_channels=(8, 8, 8, 8, 8))), 2, 4, 9, 0, True, False, 1, 0)
    # null operands are rejected before any launch
    st = lib.mvf_gemm(0, 0, 0, 1, 1, 4, 4, 4, None, 4, None, 4, None, 4, None, None, 0, 0, 1, None)
    assert st == 1 and "null operand" in L.last_error()


def test_product_refuses_cpu_tensors():
    """No CPU fallback: the autograd Functions raise on host tensors instead of computing elsewhere."""
    hc = O.HeadCfg(c_in=48, n_entities=3, pool_channels=32, fc_channels=(64, 64), hidden=32, d_ff=64, n_heads=4,
                   n_layers=1, emb=16, proj=16, train_frames=4)
    cs = engine.CallState(spec=H.spec_from_headcfg(hc), opts=engine.RunOptions(), training=True, bn_running=[], bn_tracked=[])
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        engine.HeadFn.apply(torch.zeros(2, 4, 9, 48), None, cs)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        engine.SCLFn.apply(torch.zeros(1, 2, 4, 16), torch.ones(1, 2), torch.zeros(1, 2, 4), torch.ones(2, 1, 4), 0.1, 10.0,
                           "single_noself", True)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libmvf_b200.so")
    with pytest.raises(RuntimeError, match="no PyTorch/CPU fallback"):
        L.lib()
