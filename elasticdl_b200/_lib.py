"""ctypes binding of libb200ps.so (include/b200ps.h).

There is deliberately NO fallback: if the CUDA library is missing or a call
fails, an exception is raised.  The product path never routes through a CPU
implementation.
"""
import ctypes
import os

from elasticdl_b200 import build as _build

MAX_SHARDS = 16
MAX_SEGS = 96

OK, EINVAL, ECUDA, ENOTFOUND, EWIDTH, ERANGE, ESTATE = 0, -1, -2, -3, -4, -5, -6


class Seg(ctypes.Structure):
    _fields_ = [
        ("table", ctypes.c_int32),
        ("n", ctypes.c_int32),
        ("ids_dev", ctypes.c_void_p),
        ("n_dev", ctypes.c_void_p),
        ("rows_dev", ctypes.c_void_p),
    ]


class DeepFMArgs(ctypes.Structure):  # b200_deepfm_args_t (include/b200_deepfm.h)
    _fields_ = [
        ("G", ctypes.c_int32), ("B", ctypes.c_int32),
        ("inv", ctypes.c_void_p), ("n_unique", ctypes.c_void_p),
        ("bet_wide", ctypes.c_void_p), ("bet_deep", ctypes.c_void_p),
        ("dense", ctypes.c_void_p), ("labels", ctypes.c_void_p),
        ("params", ctypes.c_void_p), ("grads", ctypes.c_void_p),
        ("gsum_wide", ctypes.c_void_p), ("gsum_deep", ctypes.c_void_p),
        ("loss", ctypes.c_void_p), ("logits", ctypes.c_void_p), ("scratch", ctypes.c_void_p),
    ]


class PSError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200ps error %d: %s" % (code, msg))
        self.code = code


class PSNotFound(PSError, KeyError):
    pass


class PSRangeError(PSError, ValueError):
    pass


# every symbol include/b200ps.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
_segp = ctypes.POINTER(Seg)
SYMBOLS = {
    "b200ps_last_error": (ctypes.c_char_p, []),
    "b200ps_abi_version": (_i, []),
    "b200ps_create": (_i, [_i, _i, ctypes.c_char_p, ctypes.c_char_p, _i, ctypes.c_uint, ctypes.POINTER(_vp)]),
    "b200ps_destroy": (_i, [_vp]),
    "b200ps_clone_view": (_i, [_vp, _i, ctypes.POINTER(_vp)]),
    "b200ps_shard_create_local": (_i, [_vp, _i, _i]),
    "b200ps_shard_export": (_i, [_vp, _i, _vp, _sz, ctypes.POINTER(_sz)]),
    "b200ps_shard_import": (_i, [_vp, _i, _vp, _sz]),
    "b200ps_table_register": (_i, [_vp, ctypes.c_char_p, _i, ctypes.c_char_p, _i64, ctypes.c_uint64]),
    "b200ps_table_register_pair": (_i, [_vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, _i64, ctypes.c_uint64]),
    "b200ps_table_register_hashed": (_i, [_vp, ctypes.c_char_p, _i, ctypes.c_char_p, _i64, ctypes.c_uint64]),
    "b200ps_dense_register": (_i, [_vp, ctypes.c_char_p, _i, _i64, _i]),
    "b200ps_lookup": (_i, [_vp, ctypes.c_char_p]),
    "b200ps_commit": (_i, [_vp]),
    "b200ps_pull_rows": (_i, [_vp, _segp, _i, _vp]),
    "b200ps_set_rows": (_i, [_vp, _segp, _i, _vp]),
    "b200ps_pull_rows_pair": (_i, [_vp, _segp, ctypes.POINTER(_vp), _i, _vp]),
    "b200ps_push_rows_pair": (_i, [_vp, _segp, ctypes.POINTER(_vp), _i, _vp]),
    "b200ps_xchg_create": (_i, [_vp, _i, _i, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "b200ps_xchg_pull": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "b200ps_xchg_profile": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_float), _vp]),
    "b200ps_xchg_push": (_i, [_vp, _vp, _vp, _vp]),
    "b200ps_pull_dense": (_i, [_vp, _segp, _i, _vp]),
    "b200ps_set_dense": (_i, [_vp, _segp, _i, _vp]),
    "b200ps_slot_rows": (_i, [_vp, _i, _i, _segp, _i, _vp]),
    "b200ps_slot_dense": (_i, [_vp, _i, _i, _segp, _i, _vp]),
    "b200ps_push_begin": (_i, [_vp, _f, ctypes.POINTER(ctypes.c_int32), _vp]),
    "b200ps_push_rows": (_i, [_vp, _segp, _i, _vp]),
    "b200ps_push_dense": (_i, [_vp, _segp, _i, _vp]),
    "b200ps_push_dense_reduce": (_i, [_vp, _i, ctypes.POINTER(_vp), _i, _f, _vp]),
    "b200ps_push_end": (_i, [_vp, _vp, _vp]),
    "b200ps_bump_step": (_i, [_vp, _vp]),
    "b200ps_push_begin_shard": (_i, [_vp, _i, _f, ctypes.POINTER(ctypes.c_int32), _vp]),
    "b200ps_push_end_shard": (_i, [_vp, _i, _vp]),
    "b200ps_raw_register": (_i, [_vp, ctypes.c_char_p, _sz]),
    "b200ps_raw_ptr": (_i, [_vp, _i, _i, ctypes.POINTER(_vp), ctypes.POINTER(_sz)]),
    "b200ps_barrier": (_i, [_vp, _vp]),
    "b200ps_kernel_sgd": (_i, [_vp, _vp, _f, ctypes.c_longlong, _vp]),
    "b200ps_kernel_momentum": (_i, [_vp, _vp, _vp, _f, _i, _f, ctypes.c_longlong, _vp]),
    "b200ps_kernel_adam": (_i, [_vp, _vp, _vp, _vp, _f, ctypes.c_longlong, ctypes.c_longlong, _f, _f, _f, _vp, _vp]),
    "b200ps_kernel_adagrad": (_i, [_vp, _vp, _vp, _f, ctypes.c_longlong, _f, _vp]),
    "b200ps_unique_workspace": (_sz, [_i, _i64]),
    "b200ps_unique": (_i, [_vp, _vp, _i, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200ps_unique_bounded_workspace": (_sz, [_i, _i64, ctypes.POINTER(_i64)]),
    "b200ps_unique_bounded": (_i, [_vp, _vp, _i, _i64, ctypes.POINTER(_i64), _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200ps_unique_bounded_i32": (_i, [_vp, _vp, _i, _i64, ctypes.POINTER(_i64), _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200ps_unique_bounded_ex": (_i, [_vp, _vp, _i, _i, _i64, ctypes.POINTER(_i64), _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "b200ps_debug_buffer": (_i, [_vp, _sz]),
    "b200ps_unique_packed": (_i, [_vp, _vp, ctypes.POINTER(ctypes.c_int32), _i, _i64, ctypes.POINTER(_i64), _vp, _vp, _vp, _vp,
                                  _sz, _i, _vp]),
    "b200ps_packed_ids_bytes": (_sz, [ctypes.POINTER(ctypes.c_int32), _i, _i64]),
    "b200ps_segment_sum": (_i, [_vp, _vp, _vp, _i, _i64, _i, _vp, _vp]),
    "b200ps_gather_rows": (_i, [_vp, _vp, _vp, _i, _i64, _i, _vp, _vp]),
    "b200ps_shard_state": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(_i64), ctypes.POINTER(ctypes.c_int32)]),
    "b200ps_set_shard_state": (_i, [_vp, _i, ctypes.c_int32, _i64, ctypes.c_int32]),
    "b200ps_snapshot_state": (_i, [_vp, _vp, _vp]),
    "b200ps_try_init": (_i, [_vp, _i, ctypes.POINTER(_i)]),
    "b200ps_finish_init": (_i, [_vp, _i, ctypes.c_int32, _vp]),
    "b200ps_table_size": (_i, [_vp, _i, _i, ctypes.POINTER(_i64)]),
    "b200ps_table_ids": (_i, [_vp, _i, _i, _vp, _i64, ctypes.POINTER(_i64)]),
    "b200ps_check": (_i, [_vp]),
    "b200ps_launch_count": (_i64, [_vp]),
    # include/b200_deepfm.h
    "b200_deepfm_param_count": (_sz, [_i]),
    "b200_deepfm_fwd_bwd": (_i, [ctypes.POINTER(DeepFMArgs), _vp]),
    "b200_deepfm_forward": (_i, [ctypes.POINTER(DeepFMArgs), _vp]),
    "b200_deepfm_fwd_bwd_mma": (_i, [ctypes.POINTER(DeepFMArgs), _vp]),
    "b200_deepfm_mma_launch_count": (_i64, []),
    "b200_deepfm_launch_count": (_i64, []),
    "b200_deepfm_fwd_bwd_tile": (_i, [ctypes.POINTER(DeepFMArgs), _vp]),
    "b200_deepfm_forward_tile": (_i, [ctypes.POINTER(DeepFMArgs), _vp]),
    "b200_deepfm_publish_loss": (_i, [_vp, _vp, _i, _vp, _vp]),
    "b200_deepfm_tile_prologue": (_i, [ctypes.POINTER(DeepFMArgs), _vp]),
    "b200_deepfm_tile_main": (_i, [ctypes.POINTER(DeepFMArgs), _vp]),
    "b200_deepfm_tile_launch_count": (_i64, []),
    "b200_deepfm_tile_last_error": (ctypes.c_char_p, []),
    # include/b200_features.h
    "b200feat_last_error": (ctypes.c_char_p, []),
    "b200feat_transform": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _i, _i, _i64, _vp, _i, _vp, _vp]),
    "b200feat_hash_strings": (_i, [_vp, _i, _i64, _i64, _vp, _vp]),
    "b200feat_hash_ints": (_i, [_vp, _i64, _i64, _vp, _vp]),
    "b200feat_bucketize": (_i, [_vp, _i64, _vp, _i, _vp, _vp]),
    "b200feat_fingerprint64": (_i, [_vp, _i, _i64, _vp, _vp]),
    "b200feat_launch_count": (_i64, []),
}

_lib = None


def lib():
    """Load (building in-tree first if the sources are newer) the CUDA library."""
    global _lib
    if _lib is None:
        path = _build.LIB
        if not os.path.exists(path) or _build._stale():
            path = _build.build()
        handle = ctypes.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the ABI drifted: fail loudly
            fn.restype = res
            fn.argtypes = args
        if handle.b200ps_abi_version() != 1:
            raise ImportError("libb200ps.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc):
    """Turn a negative return code into the matching Python exception."""
    if rc >= 0:
        return rc
    msg = lib().b200ps_last_error().decode("utf-8", "replace")
    if rc == ENOTFOUND:
        raise PSNotFound(rc, msg)
    if rc == ERANGE:
        raise PSRangeError(rc, msg)
    if rc in (EINVAL, EWIDTH):
        raise ValueError("b200ps error %d: %s" % (rc, msg))
    raise PSError(rc, msg)
