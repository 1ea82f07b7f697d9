"""ctypes binding of libsstg.so (C ABI declared in include/sstg.h).

There is deliberately no fallback: if the CUDA library is missing or a call
fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSTG_LIB") or os.path.join(_PKG, "libsstg.so")   # SSTG_LIB: A/B a differently built library

T_OBS, T_PRED, C_IN, C_OUT = 8, 12, 2, 5
N_PARAMS, N_BN, MAX_PEDS, MAX_RANKS = 7563, 30, 256, 8
XTILE_COLS, XTILE_MAX_SCENES, XTILE_MAX_PAIRS = 128, 32, 64
ST_PARAMS, TXP_GRAD_STRIDE, SAVE_FLOATS_PER_PED = 142, 7424, 600
GRAPH_NORM_LAP, GRAPH_FP64, GRAPH_V_ONLY = 1, 2, 4
EVAL_FP64_DIST, EVAL_VTR_IS_REL, EVAL_UNIFORM_N = 1, 2, 4

_p = C.c_void_p
_i32, _i64, _u64 = C.c_int32, C.c_int64, C.c_uint64
_SIGS = {
    "sstg_version": (C.c_int, []),
    "sstg_error_string": (C.c_char_p, [C.c_int]),
    "sstg_last_cuda_error": (C.c_char_p, []),
    "sstg_plan_host": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "sstg_graph_build": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _p, _p, _i32, _p]),
    "sstg_stgcn_forward": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _p, _p, _p, _i32, _p]),
    "sstg_txpcnn_forward": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "sstg_nll_forward": (C.c_int, [_p, _p, _p, _i32, _p, _p]),
    "sstg_sample_ade_fde": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _p, _u64, _p, _p, _p, _p, _i32, _p]),
}
_f32, _f64 = C.c_float, C.c_double
_SIGS.update({
    "sstg_train_partial_rows": (C.c_int, [_i32, _i32, _p, _p]),
    "sstg_forward_train": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p]),
    "sstg_bn_running_update": (C.c_int, [_p, _p, _i32, _f64, _p, _p]),
    "sstg_nll_backward": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p, _p]),
    "sstg_backward": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "sstg_rel_to_abs": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _p, _p]),
    "sstg_traj_error": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "sstg_train_step_workspace": (C.c_int64, [_i32, _i32, _i32, _i64]),
    "sstg_train_step": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i64, _p, _p, _p, _p, _i64, _f32, _f32,
                                  _f64, _i32, _p, _p, _p, _p]),
    "sstg_stgcn_forward_fused": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, _p, _p]),
    "sstg_decode_tracks": (C.c_int, [_p, _p, _i32, _p, _p, _p, _p]),
    "sstg_sgd_step": (C.c_int, [_p, _p, _i32, _f32, _f32, _p, _p]),
    "sstg_train_step_stats_offset": (C.c_int64, [_i32, _i32, _i32, _i64]),
    "sstg_peer_mailbox_bytes": (C.c_int64, [_i32, _i32]),
    "sstg_peer_alloc": (C.c_int, [_i64, C.POINTER(C.c_void_p), C.c_char_p]),
    "sstg_peer_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "sstg_peer_close": (C.c_int, [_p]),
    "sstg_peer_free": (C.c_int, [_p]),
    "sstg_peer_exchange_apply": (C.c_int, [C.POINTER(C.c_void_p), _i32, _i32, C.c_uint32, _i32, _p, _p, _p, _i32, _i32,
                                           _p, _i32, _p, _p, _f32, _f32, _f64, _p, _p, _p, _p, C.c_uint64, _p]),
})



class PeerRank(C.Structure):
    """sstg_peer_rank (include/sstg.h): one rank's arguments of sstg_peer_exchange_apply_local."""
    _fields_ = [("rank", _i32), ("scene_lo", _i32), ("s_local", _i32), ("reserved", _i32),
                ("grad_local", _p), ("loss_local", _p), ("stats_local", _p), ("seg_off_all", _p),
                ("params", _p), ("bn", _p), ("grad_out", _p), ("loss_out", _p), ("stats_out", _p), ("status", _p)]


_SIGS["sstg_peer_exchange_apply_local"] = (C.c_int, [C.POINTER(C.c_void_p), _i32, C.POINTER(PeerRank), _i32, C.c_uint32,
                                                     _i32, _i32, _f32, _f32, _f64, C.c_uint64, _p])

_lock = threading.Lock()
_lib = None


def load():
    """Load libsstg.so once; raises RuntimeError (never falls back) if absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA library has not been built. "
                "Run `python -m social_stgcnn_b200.build` (needs nvcc). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        lib = load()
        msg = lib.sstg_error_string(code).decode()
        cuda = lib.sstg_last_cuda_error().decode()
        raise RuntimeError(f"{what} failed: {msg} (code {code})" + (f" [{cuda}]" if cuda and code == -3 else ""))


def exported_symbols():
    return list(_SIGS)
