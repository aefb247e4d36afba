"""ctypes binding of the C ABI declared in ``include/snfft_b200.h``.

The shared library is built in-tree by ``snfft_b200/build.py`` (nvcc, sm_100a) as
``snfft_b200/_lib/libsnfft_b200.so``.  There is no Python or CPU fallback for the transform: if
the library is missing, import of this module raises, and device entry points raise
``SnfftError`` when no B200 is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SNFFT_B200_LIB") or os.path.join(_HERE, "_lib", "libsnfft_b200.so")

SNFFT_OK = 0
ORDER_COSET = 0
ORDER_LEX = 1
MAX_N = 12

_ERR_NAMES = {1: "invalid argument", 2: "no device", 3: "CUDA error", 4: "allocation failure"}


class SnfftError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m snfft_b200.build` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). snfft_b200 has no CPU fallback.")
    return C.CDLL(LIB_PATH)


lib = _load()

_p = C.c_void_p
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol include/snfft_b200.h declares
SIGNATURES = {
    "snfft_abi_version": (C.c_int, []),
    "snfft_last_error": (C.c_char_p, []),
    "snfft_plan_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_p)]),
    "snfft_plan_destroy": (C.c_int, [_p]),
    "snfft_plan_n": (C.c_int, [_p]),
    "snfft_plan_device": (C.c_int, [_p]),
    "snfft_num_irreps": (C.c_int, [_p, C.c_int]),
    "snfft_layout": (C.c_int, [_p, C.c_int, _p, _p, _p]),
    "snfft_branch_down": (C.c_int, [_p, C.c_int, C.c_int, _p, _p, _p]),
    "snfft_yor_trans": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, _p]),
    "snfft_forward": (C.c_int, [_p, _p, _p, _p, C.c_int64, C.c_int, _p]),
    "snfft_inverse": (C.c_int, [_p, _p, _p, _p, C.c_int64, C.c_int, _p]),
    "snfft_level": (C.c_int, [_p, C.c_int, _p, _p, C.c_int64, C.c_int, _p]),
    "snfft_reorder": (C.c_int, [_p, _p, _p, C.c_int64, C.c_int, _p]),
    "snfft_coset_to_lex_host": (C.c_int, [_p, _p]),
    "snfft_yor": (C.c_int, [_p, C.c_int, _p, C.c_int64, _p, _p]),
    "snfft_ysemi": (C.c_int, [_p, C.c_int, _p, C.c_int64, _p, _p]),
    "snfft_inverse_at": (C.c_int, [_p, _p, _p, C.c_int64, _p, _p]),
    "snfft_dft_direct": (C.c_int, [_p, _p, _p, C.c_int64, _p]),
    "snfft_dft_release": (C.c_int, [_p]),
    "snfft_shard_create": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.POINTER(_p)]),
    "snfft_shard_destroy": (C.c_int, [_p]),
    "snfft_shard_blocks": (C.c_int, [_p, C.c_int, _p, _p]),
    "snfft_shard_size": (C.c_int, [_p, C.c_int, C.c_int, _p]),
    "snfft_shard_gather_index": (C.c_int, [_p, C.c_int, _p]),
    "snfft_shard_columns": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, _p, _p]),
    "snfft_shard_level": (C.c_int, [_p, C.c_int, _p, _p, C.c_int64, C.c_int, _p]),
    "snfft_shard_level_scratch": (C.c_int, [_p, C.c_int, _p, _p, C.c_int64, C.c_int, _p]),
    "snfft_shard_export_panel": (C.c_int, [_p, C.c_int, C.c_int, _p, _p, _p, _p, _p]),
    "snfft_shard_exchange": (C.c_int, [_p, _p, _p, C.c_int, _p]),
    "snfft_shard_ws_sizes": (C.c_int, [_p, _p]),
    "snfft_forward_sharded": (C.c_int, [_p, _p, _p, _p, _p]),
    "snfft_inverse_sharded": (C.c_int, [_p, _p, _p, _p, _p]),
    "snfft_wreath_create": (C.c_int, [C.c_int, _p, _p, C.c_int, C.c_int, C.POINTER(_p)]),
    "snfft_wreath_destroy": (C.c_int, [_p]),
    "snfft_wreath_dims": (C.c_int, [_p, _p]),
    "snfft_wreath_forward": (C.c_int, [_p, _p, _p, _p, _p]),
    "snfft_wreath_inverse": (C.c_int, [_p, _p, _p, _p, _p, _p]),
    "snfft_wreath_inverse_at": (C.c_int, [_p, _p, _p, _p, _p, C.c_int64, _p, _p, _p]),
    "snfft_dgemm": (C.c_int, [_p, _p, _p, C.c_int64, C.c_int64, C.c_int64, _p]),
    "snfft_c3_dft": (C.c_int, [C.c_int, _p, _p, _p, C.c_int, C.c_int64, _p, _p, C.c_int, _p]),
    "snfft_gather": (C.c_int, [_p, _p, _p, C.c_int64, C.c_int, _p]),
    "snfft_plan_device_bytes": (C.c_int64, [_p]),
    "snfft_level_macs_per_element": (C.c_double, [_p, C.c_int]),
    "snfft_launch_count": (C.c_int64, []),
    "snfft_debug_force_table_driven": (C.c_int, [C.c_int]),
    "snfft_debug_fused_tma": (C.c_int, [C.c_int]),
    "snfft_timing_enable": (C.c_int, [C.c_int]),
    "snfft_timing_read": (C.c_int, [_p, _p]),
    "snfft_export_panel": (C.c_int, [_p, C.c_int, C.c_int, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "snfft_export_phase": (C.c_int, [_p, C.c_int, C.c_int, _p, _p, _p, _p, _p]),
    "snfft_export_coef": (C.c_int, [_p, C.c_int, _p, _p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)       # AttributeError here = header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args


class ShardWorkspace(C.Structure):
    """``snfft_shard_ws`` of include/snfft_b200.h."""
    _fields_ = [("peer_cols", C.c_void_p * 16), ("peer_pads", C.c_void_p * 16),
                ("local_a", C.c_void_p), ("local_b", C.c_void_p),
                ("shard_a", C.c_void_p), ("shard_b", C.c_void_p), ("status", C.c_void_p)]


def check(status: int) -> None:
    if status != SNFFT_OK:
        msg = lib.snfft_last_error().decode("utf-8", "replace")
        raise SnfftError(f"snfft_b200: {_ERR_NAMES.get(status, status)}: {msg}")


def ptr(array_or_int):
    """A c_void_p from a numpy array, a torch tensor or an integer address (None -> NULL)."""
    if array_or_int is None:
        return None
    if isinstance(array_or_int, int):
        return C.c_void_p(array_or_int)
    if hasattr(array_or_int, "data_ptr"):
        return C.c_void_p(array_or_int.data_ptr())
    return C.c_void_p(array_or_int.ctypes.data)
