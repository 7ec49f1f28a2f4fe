"""ctypes binding of ``include/kge_b200.h`` (libkge_b200.so).

The product path has NO CPU or PyTorch fallback: if the CUDA library cannot be loaded, or a
tensor is not on a CUDA device, the call raises.  PyTorch is used only for device memory,
streams and autograd plumbing.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KGE_B200_LIB") or os.path.join(_HERE, "lib", "libkge_b200.so")

MODEL_IDS = {"TransE": 0, "DistMult": 1, "ComplEx": 2, "RotatE": 3, "pRotatE": 4}
ABI_VERSION = 2
TAIL_BATCH, HEAD_BATCH = 0, 1


class KgeTables(C.Structure):
    _fields_ = [
        ("entity", C.c_void_p),
        ("relation", C.c_void_p),
        ("n_entity", C.c_int64),
        ("n_relation", C.c_int64),
        ("hidden_dim", C.c_int32),
        ("model", C.c_int32),
        ("gamma", C.c_float),
        ("embedding_range", C.c_float),
        ("modulus", C.c_void_p),  # pRotatE: device scalar; NULL otherwise
    ]


class KgeFilterCsr(C.Structure):
    _fields_ = [
        ("keys", C.c_void_p),
        ("offsets", C.c_void_p),
        ("members", C.c_void_p),
        ("n_keys", C.c_int64),
    ]


MAX_SHARDS = 16  # KGE_MAX_SHARDS


class KgeShards(C.Structure):
    _fields_ = [
        ("entity", C.c_void_p * MAX_SHARDS),
        ("grad_entity", C.c_void_p * MAX_SHARDS),
        ("n_shards", C.c_int32),
        ("scalar_red", C.c_int32),
    ]


# name -> (restype, argtypes); mirrors include/kge_b200.h declaration by declaration
_P = C.c_void_p
_I64 = C.c_int64
PROTOTYPES = {
    "kge_abi_version": (C.c_int, []),
    "kge_strerror": (C.c_char_p, [C.c_int]),
    "kge_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "kge_score_fwd": (C.c_int, [C.POINTER(KgeTables), C.c_int, _P, _I64, _P, _I64, _P, _P]),
    "kge_score_bwd": (C.c_int, [C.POINTER(KgeTables), C.c_int, _P, _I64, _P, _I64, _P, _P, _P, _P]),
    "kge_loss_workspace_bytes": (C.c_size_t, [_I64]),
    "kge_adv_loss_fwd": (C.c_int, [_P, _P, _P, _I64, _I64, C.c_float, _P, _P, _P]),
    "kge_adv_loss_bwd": (C.c_int, [_P, _P, _P, _I64, _I64, C.c_float, _P, _P, _P, _P, _P]),
    "kge_modulus_grad": (C.c_int, [_P, _P, _I64, _P, _P, C.c_float, _P, _P, _P]),
    "kge_kl_div_fwd": (C.c_int, [_P, _P, _I64, _I64, C.c_float, _P, _P, _P]),
    "kge_kl_div_bwd": (C.c_int, [_P, _P, _I64, _I64, C.c_float, _P, _P, _P, _P]),
    "kge_topk_rows": (C.c_int, [_P, _I64, _I64, _I64, C.c_int32, _P, _P, _P]),
    "kge_fused_fwd": (C.c_int, [C.POINTER(KgeTables), C.c_int, _P, _I64, _P, _I64, _P, C.c_float,
                                _P, _P, _P, _P, _P, _P, _P]),
    "kge_fused_bwd": (C.c_int, [C.POINTER(KgeTables), C.c_int, _P, _I64, _P, _I64, _P, _P, _P, _P,
                                _P, _P, _P]),
    "kge_fused_bwd_chunk": (C.c_int, [C.POINTER(KgeTables), C.c_int, _P, _I64, _P, _I64, _P, _P, _P, _P,
                                      C.c_int32, C.c_int32, C.c_int32, _I64, _P, _P, _P]),
    "kge_byent_workspace_bytes": (C.c_size_t, [C.POINTER(KgeTables), _I64, _I64]),
    "kge_bwd_by_entity_adam": (C.c_int, [C.POINTER(KgeTables), C.c_int, _P, _I64, _P, _I64, _P, _P, _P, _P, _P, _P, _P,
                                         _P, _P, _P, _I64, C.c_float, C.c_float, C.c_float, C.c_float, _P, _P]),
    "kge_adam_step_chunk": (C.c_int, [_P, _P, _P, _P, _I64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, _I64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                      _P]),
    "kge_adam_slice_bcast": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, _P, _P, _P, _I64, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_int32, _I64, C.c_float, C.c_float,
                                       C.c_float, C.c_float, C.c_int, _P]),
    "kge_tma_fail_flag": (C.c_int, []),
    "kge_peer_copy": (C.c_int, [_P, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, _I64, _I64, _P]),
    "kge_peer_signal": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_uint32, _P]),
    "kge_peer_wait": (C.c_int, [_P, C.c_int32, C.c_uint32, _I64, _P, _P]),
    "kge_fused_fwd_sharded": (C.c_int, [C.POINTER(KgeTables), C.POINTER(KgeShards), C.c_int, _P, _I64, _P, _I64,
                                        _P, C.c_float, _P, _P, _P, _P, _P, _P, _P]),
    "kge_fused_bwd_sharded": (C.c_int, [C.POINTER(KgeTables), C.POINTER(KgeShards), C.c_int, _P, _I64, _P, _I64,
                                        _P, _P, _P, _P, _P, _P]),
    "kge_score_fwd_sharded": (C.c_int, [C.POINTER(KgeTables), C.POINTER(KgeShards), C.c_int, _P, _I64, _P, _I64,
                                        _P, _P]),
    "kge_rank_counts_sharded": (C.c_int, [C.POINTER(KgeTables), C.POINTER(KgeShards), C.c_int32, C.c_int, _P, _I64,
                                          C.POINTER(KgeFilterCsr), _P, _P, _P, _P]),
    "kge_sample_negatives": (C.c_int, [C.POINTER(KgeFilterCsr), C.c_int, _P, _I64, _I64, _I64,
                                       C.c_uint64, C.c_uint64, C.c_int, _P, _P, _P]),
    "kge_filter_pool": (C.c_int, [C.POINTER(KgeFilterCsr), C.c_int, _P, _I64, _I64, _I64, _P, _I64,
                                  _P, _P, _P]),
    "kge_filter_pool_positions": (C.c_int, [C.POINTER(KgeFilterCsr), C.c_int, _P, _I64, _I64, _I64, _P, _I64,
                                            _P, _P, _P, _P]),
    "kge_pooled_workspace_bytes": (C.c_size_t, [C.POINTER(KgeTables), _I64, _I64, _I64]),
    "kge_pooled_dot_fwd": (C.c_int, [C.POINTER(KgeTables), C.c_int, _P, _I64, _P, _I64, _P, _I64, _P, C.c_float,
                                     _P, _P, _P, _P, _P, _P, _P]),
    "kge_pooled_dot_bwd": (C.c_int, [C.POINTER(KgeTables), C.c_int, _P, _I64, _P, _I64, _I64, _P, _P, _P, _P, _P,
                                     _P, _P]),
    "kge_rank_workspace_bytes": (C.c_size_t, [C.POINTER(KgeTables), _I64]),
    "kge_rank_all": (C.c_int, [C.POINTER(KgeTables), C.c_int, _P, _I64, C.POINTER(KgeFilterCsr), _P,
                               _P, _P, _P]),
    "kge_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _I64, C.c_float, C.c_float, C.c_float,
                                C.c_float, C.c_int, _P]),
}

_lib = None
_lock = threading.Lock()
launches = 0  # number of kernel-launching ABI calls made through this module (bench.py reads it)


class KgeError(RuntimeError):
    """Raised for every non-zero return of the C ABI; ``code`` is that return value (negative: KGE_E_*,
    positive: cudaError_t)."""

    code = None


E_UNSUPPORTED = -6  # KGE_E_UNSUPPORTED


def load(build_if_missing: bool = True):
    """dlopen libkge_b200.so (building it with nvcc first when it is absent and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise KgeError(f"{LIB_PATH} is missing: run `python -m mkb_b200.build`")
            from . import build as _build

            _build.build()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError here = header and library disagree
            fn.restype = res
            fn.argtypes = args
        if lib.kge_abi_version() != ABI_VERSION:
            raise KgeError("libkge_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().kge_strerror(rc).decode()
        err = KgeError(f"{what or 'kge call'} failed: {msg} (code {rc})")
        err.code = rc
        raise err


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise KgeError(
                "mkb_b200 computes on CUDA only (no CPU fallback): got a tensor on "
                f"'{t.device}'. Move the model and the batch to a B200 with .to('cuda')."
            )


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def count_launch(n=1):
    global launches
    launches += n
