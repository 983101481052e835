"""ctypes binding of libcutdet_b200.so (the C ABI declared in include/cutdet_b200.h).

The library is the product: if it cannot be loaded every numeric entry point of this package raises
``RuntimeError`` -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import threading

from . import build as _build

_lock = threading.Lock()
_lib = None
LIB_OVERRIDE = None      # development aid (bench.py --lib): load this build of the library instead of the in-tree one

OK, EINVAL, ECUDA, EUNSUPPORTED, ECAPACITY, ELONE_ORPHAN = range(6)
ABI_VERSION = 2
NET_OPTIONS = {"conv1_acc32": 1, "sub_batch": 2, "group_frames": 3, "no_pdl": 4, "conv1_grid": 5, "conv1_variant": 6, "l2_persist": 7, "ring_cap": 8, "src_prefetch": 9}


class Frames(C.Structure):
    """cutdet_frames"""
    _fields_ = [("frames_dev", C.c_void_p), ("frame_stride", C.c_int64), ("row_pitch", C.c_int64),
                ("batch", C.c_int), ("row_map_compact", C.c_int)]


class NetConfig(C.Structure):
    """cutdet_net_config"""
    _fields_ = [(n, C.c_int) for n in ("input_channels", "hidden_channels", "n_conv_layers", "avg_pool_size",
                                         "n_fc_layers", "fc_input_size", "fc_hidden_size", "fc_output_size")]


class RunTable(C.Structure):
    """cutdet_run_table"""
    _fields_ = [("end_frames_dev", C.c_void_p), ("start_frames_dev", C.c_void_p), ("run_lengths_dev", C.c_void_p),
                ("frame_types_dev", C.c_void_p), ("score_means_dev", C.c_void_p), ("score_sums_dev", C.c_void_p),
                ("capacity", C.c_int64)]


_P = C.c_void_p
_PF = C.POINTER(C.c_float)
_SIGNATURES = {
    "cutdet_abi_version": (C.c_int, []),
    "cutdet_last_error": (C.c_char_p, []),
    "cutdet_device_check": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "cutdet_launch_count": (C.c_longlong, []),
    "cutdet_profile_begin": (C.c_int, []),
    "cutdet_profile_end": (C.c_int, [C.c_char_p, C.c_size_t]),
    "cutdet_upload_frames": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int64, _P, _P, C.POINTER(C.c_int64)]),
    "cutdet_target_size": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cutdet_resize_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "cutdet_resize_plan_destroy": (None, [_P]),
    "cutdet_resize_plan_rows": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cutdet_resize_rows": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cutdet_debug_k1_kernel": (C.c_int, [C.c_int]),
    "cutdet_preprocess_f32": (C.c_int, [_P, C.POINTER(Frames), _P, _P]),
    "cutdet_preprocess_u8": (C.c_int, [_P, C.POINTER(Frames), _P, _P]),
    "cutdet_net_create": (C.c_int, [C.POINTER(NetConfig), C.POINTER(_P)]),
    "cutdet_net_destroy": (None, [_P]),
    "cutdet_net_set_conv_layer": (C.c_int, [_P, C.c_int, _PF, _PF, _PF, _PF, _PF, _PF, C.c_float]),
    "cutdet_net_set_fc_layer": (C.c_int, [_P, C.c_int, _PF, _PF, _PF, _PF, _PF, _PF, C.c_float]),
    "cutdet_net_finalize": (C.c_int, [_P]),
    "cutdet_net_set_option": (C.c_int, [_P, C.c_int, C.c_int]),
    "cutdet_net_get_option": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int)]),
    "cutdet_net_forward_conv_layer": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P]),
    "cutdet_net_forward_fc_layer": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, _P]),
    "cutdet_net_debug_timeline": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "cutdet_net_uses_tensor_cores": (C.c_int, [_P, C.c_int, C.c_int]),
    "cutdet_net_workspace_bytes": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "cutdet_net_forward_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_size_t, _P]),
    "cutdet_net_forward_f32_batchstats": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_size_t, C.c_int, _P]),
    "cutdet_net_forward_frames": (C.c_int, [_P, _P, C.POINTER(Frames), _P, _P, C.c_size_t, _P]),
    "cutdet_contrastive_loss_workspace_bytes": (C.c_size_t, [C.c_int]),
    "cutdet_contrastive_loss": (C.c_int, [_P, C.c_int, C.c_int, C.c_float, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "cutdet_cross_entropy_workspace_bytes": (C.c_size_t, [C.c_int]),
    "cutdet_cross_entropy_sum": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, _P, _P, _P, C.c_size_t, C.POINTER(C.c_int), _P]),
    "cutdet_net_debug_conv_output": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "cutdet_argmax": (C.c_int, [_P, C.c_int64, C.c_int, _P, _P, _P]),
    "cutdet_rle_state_bytes": (C.c_size_t, []),
    "cutdet_rle_reset": (C.c_int, [_P, _P]),
    "cutdet_rle_append": (C.c_int, [_P, _P, _P, C.c_int64, C.POINTER(RunTable), _P]),
    "cutdet_rle_finish": (C.c_int, [_P, C.POINTER(RunTable), _P, _P]),
    "cutdet_rle_count": (C.c_int, [_P, C.POINTER(C.c_int64), _P]),
    "cutdet_glue_orphans_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "cutdet_glue_orphans": (C.c_int, [C.POINTER(RunTable), _P, C.c_int, C.c_int, _P, _P, C.c_size_t, _P]),
    "cutdet_combine_adjacent": (C.c_int, [C.POINTER(RunTable), _P, _P]),
    "cutdet_stitch_shards": (C.c_int, [C.POINTER(RunTable), C.c_int, C.c_int64, _P, _P, C.POINTER(RunTable), _P, _P]),
    "cutdet_shard_pack_bytes": (C.c_size_t, [C.c_int64]),
    "cutdet_shard_pack": (C.c_int, [C.POINTER(RunTable), _P, C.c_int64, C.c_int64, _P, _P]),
    "cutdet_stitch_packed": (C.c_int, [_P, C.c_int, C.c_int64, C.POINTER(RunTable), _P, _P, _P]),
}


def header_path() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                        "include", "cutdet_b200.h")


def declared_symbols() -> list[str]:
    """Every function name include/cutdet_b200.h declares."""
    text = open(header_path()).read()
    return sorted(set(re.findall(r"^CUTDET_API [^;(]*?\b(cutdet_[a-z0-9_]+)\(", text, flags=re.M)))


def library_path() -> str:
    return _build.LIB_PATH


def lib() -> C.CDLL:
    """The loaded library with argument types attached.  Raises if it is missing or stale ABI."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        # raises when the library is missing or stale and cannot be rebuilt
        path = LIB_OVERRIDE if LIB_OVERRIDE else _build.ensure_current()
        handle = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError here = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        if handle.cutdet_abi_version() != ABI_VERSION:
            raise RuntimeError("libcutdet_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int) -> None:
    """Map a non-zero return code to the exception the reference's Python would have raised."""
    if rc == OK:
        return
    msg = (lib().cutdet_last_error() or b"").decode("utf-8", "replace")
    if rc == EINVAL:
        raise ValueError(msg)
    if rc == ELONE_ORPHAN:
        raise IndexError(msg or "index 1 is out of bounds for dimension 0 with size 1")
    raise RuntimeError(f"libcutdet_b200 error {rc}: {msg}")


def fptr(array):
    """numpy float32 array -> float*  (None -> NULL)."""
    if array is None:
        return None
    return array.ctypes.data_as(_PF)
