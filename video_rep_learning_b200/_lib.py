"""ctypes binding of libmvf_b200.so (include/mvf_b200.h).

There is no fallback: if the shared object is missing or a call fails, a RuntimeError carrying
mvf_last_error() is raised.  Nothing in this package computes the hot path in PyTorch or on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmvf_b200.so")

MVF_F32, MVF_BF16 = 0, 1
FINAL = {"max": 0, "one": 1, "avg": 2, "lin": 3}
ONEHOT = {"none": 0, "pool": 1, "enc": 2}
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = 0, 1, 2
NEG = {"single_noself": 0, "batch_noself": 1}
POOL_AUTO, POOL_DENSE, POOL_FOLDED = 0, 1, 2
POOLKIND = {"lstp": 0, "fwb": 1}
PHASE_ALL = 255
GEMM_RELU, GEMM_ACCUM, GEMM_RELUMASK, GEMM_SPLIT3, GEMM_B_PRESPLIT = 1, 2, 4, 8, 16
MAX_FC = 4
SITE_FC0, SITE_POS, SITE_ENC0 = 0, 8, 16


class HeadDesc(C.Structure):
    """mvf_head_desc"""
    _fields_ = [
        ("BV", C.c_int32), ("T", C.c_int32), ("P", C.c_int32), ("C_in", C.c_int32),
        ("E", C.c_int32), ("SPC", C.c_int32), ("n_fc", C.c_int32), ("fc", C.c_int32 * MAX_FC),
        ("H", C.c_int32), ("DFF", C.c_int32), ("heads", C.c_int32), ("L", C.c_int32),
        ("D", C.c_int32), ("PS", C.c_int32), ("one_hot", C.c_int32), ("final_mode", C.c_int32),
        ("train_frames", C.c_int32), ("dtype", C.c_int32), ("training", C.c_int32), ("has_mask", C.c_int32),
        ("gemm_backend", C.c_int32), ("world_size", C.c_int32), ("pool_mode", C.c_int32),
        ("drop_p", C.c_float), ("ln_eps", C.c_float), ("bn_eps", C.c_float), ("bn_momentum", C.c_float),
        ("seed", C.c_uint64), ("seed_dev", C.c_void_p),
        ("pool_kind", C.c_int32), ("cls_dim", C.c_int32), ("cls_emb", C.c_void_p),
    ]


_lib = None
_lock = threading.Lock()

_vp, _i32, _i64, _sz, _f32, _u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_float, C.c_uint64
_pd = C.POINTER(HeadDesc)

_PROTOS = {
    "mvf_version": (C.c_int, []),
    "mvf_last_error": (C.c_char_p, []),
    "mvf_has_tcgen05": (C.c_int, []),
    "mvf_launch_count": (C.c_uint64, []),
    "mvf_profile_enable": (C.c_int, [C.c_int]),
    "mvf_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]),
    "mvf_num_params": (C.c_int, [_pd]),
    "mvf_param_info": (C.c_int, [_pd, C.c_int, C.c_char_p, _sz, C.POINTER(_i64), C.POINTER(_i64)]),
    "mvf_num_bn": (C.c_int, [_pd]),
    "mvf_bn_info": (C.c_int, [_pd, C.c_int, C.c_char_p, _sz, C.POINTER(_i64)]),
    "mvf_save_bytes": (_sz, [_pd]),
    "mvf_ws_bytes": (_sz, [_pd]),
    "mvf_gpack_elems": (_sz, [_pd]),
    "mvf_gpack_pool_elems": (_sz, [_pd]),
    "mvf_proj_save_bytes": (_sz, [_pd]),
    "mvf_proj_ws_bytes": (_sz, [_pd]),
    "mvf_save_lookup": (C.c_int, [_pd, C.c_char_p, C.POINTER(_sz), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64),
                                  C.POINTER(_i32)]),
    "mvf_bn_stat_lookup": (C.c_int, [_pd, C.c_int, C.c_int, C.POINTER(_sz), C.POINTER(_i64)]),
    "mvf_head_forward": (C.c_int, [_pd, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _vp, C.c_int, C.c_int, _vp]),
    "mvf_head_backward": (C.c_int, [_pd, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp, C.c_int, C.c_int, _vp]),
    "mvf_proj_forward": (C.c_int, [_pd, _vp, _vp, _vp, _vp, C.c_int, _vp, _sz, _vp, _sz, _vp, C.c_int, C.c_int, _vp]),
    "mvf_proj_backward": (C.c_int, [_pd, _vp, _vp, C.c_int, _vp, _sz, _vp, _sz, _vp, _vp, C.c_int, C.c_int, _vp]),
    "mvf_unpack_grads": (C.c_int, [_pd, _vp, _vp, _f32, _vp]),
    "mvf_scl_ws_bytes": (_sz, [_i32, _i32, _i32]),
    "mvf_scl_fwd_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _f32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "mvf_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64,
                           _vp, _vp, _i64, C.c_int, C.c_int, _vp]),
    "mvf_xattn_pool_fwd": (C.c_int, [C.c_int, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _vp, C.c_int, _f32, _u64, _vp]),
    "mvf_xattn_pool_bwd": (C.c_int, [C.c_int, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _vp, C.c_int, _f32, _u64,
                                     _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvf_pool_fold_prep": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "mvf_pool_fold_fwd": (C.c_int, [C.c_int, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mvf_pool_fold_bwd": (C.c_int, [C.c_int, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvf_pool_fold_bwd_delta": (C.c_int, [C.c_int, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvf_pool_fold_finish": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _vp]),
    "mvf_opt_ws_bytes": (_sz, [_i32]),
    "mvf_opt_adam_step": (C.c_int, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, _f32, _f32, _i32, _f32, _f32, _vp, _vp,
                                    _sz, _vp]),
    "mvf_pool_bwd_reserve_sms": (C.c_int, [_i32]),
    "mvf_peer_buffer_bytes": (_sz, []),
    "mvf_peer_sum_f64": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _vp, _vp]),
    "mvf_peer_allreduce_flag_bytes": (_sz, []),
    "mvf_peer_allreduce_f32": (C.c_int, [_vp, _vp, _sz, _sz, _i64, _i32, _i32, _vp, _i32, _vp]),
    "mvf_attention_ws_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "mvf_attention_fwd": (C.c_int, [C.c_int, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mvf_attention_bwd": (C.c_int, [C.c_int, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mvf_dropout_mask": (C.c_int, [_u64, _i32, _i64, _i64, _f32, _vp, _vp]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS.keys())


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -m video_rep_learning_b200.build` "
                        "(there is no PyTorch/CPU fallback for the MV-Former hot path)")
                L = C.CDLL(LIB_PATH)
                for name, (res, args) in _PROTOS.items():
                    fn = getattr(L, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = L
    return _lib


def last_error() -> str:
    msg = lib().mvf_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str = "libmvf_b200") -> None:
    if status != 0:
        raise RuntimeError(f"{what} failed (status {status}): {last_error()}")


def ptr(t) -> int:
    """Device pointer of a tensor (or 0 for None)."""
    return 0 if t is None else t.data_ptr()


def ptr_array(tensors):
    """Host array of device pointers (void*[]) for a list of tensors / None."""
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr
