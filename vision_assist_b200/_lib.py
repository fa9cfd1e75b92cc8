"""ctypes binding of libva_sm100.so (include/vision_assist_b200.h).

The library is built in-tree by `vision_assist_b200/csrc/build.sh` (`__graft_entry__.build()`).
There is NO fallback: if the shared object is missing, or no CUDA device is usable, the calls
raise - the product path never routes through CPU code.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VA_LIB_PATH") or os.path.join(_HERE, "libva_sm100.so")   # VA_LIB_PATH: tuning builds

VA_OK, VA_ERR_INVALID, VA_ERR_CUDA, VA_ERR_CAPACITY, VA_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
VA_CFG_CHECK_SIMPLE, VA_CFG_NO_TENSOR_CORE = 1, 2
VA_IPC_HANDLE_BYTES = 64
VA_FLAG_EMPTY, VA_FLAG_CENTRE_OOB, VA_FLAG_LIST_OOB, VA_FLAG_NON_SIMPLE, VA_FLAG_OVERFLOW, VA_FLAG_NO_POLYGON = 1, 2, 4, 8, 16, 32

EXPORTS = ["va_abi_version", "va_create", "va_destroy", "va_last_error", "va_get_layout", "va_assemble_masks",
           "va_run_fused", "va_run_fused_host", "va_run_fused_host_f16", "va_mask_to_records", "va_grid_to_penalty_peaks", "va_nms", "va_scale_boxes",
           "va_last_launch_count", "va_uses_tensor_core", "va_profile_enable", "va_profile_read",
           "va_peer_alloc", "va_peer_open", "va_peer_close", "va_peer_free", "va_peer_put", "va_signal", "va_wait_flags"]


class VaConfig(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("device", "H", "W", "mh", "mw", "K", "max_n", "gs", "max_batch", "flags")]


class VaLayout(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("record_bytes", "rmax", "cmax", "pmax", "off_header", "off_row_y",
                                         "off_row_attr", "off_penalty", "off_peaks", "off_occ", "lat_rows",
                                         "lat_cols", "algorithmic_bytes_per_frame_n1", "off_goals", "off_lookup",
                                         "lookup_rows")]


class VaNmsParams(C.Structure):
    _fields_ = [("conf_thres", C.c_float), ("iou_thres", C.c_float), ("nc", C.c_int32), ("max_det", C.c_int32),
                ("agnostic", C.c_int32), ("max_wh", C.c_int32)]


class VaGridInput(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("x0", "n_cols", "n_rows", "n_plane", "use_easy")] + [("reserved", C.c_int32 * 3)]


class VaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libva_sm100 error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(vision_assist_b200/csrc/build.sh). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32 = C.c_void_p, C.c_int32
    lib.va_abi_version.restype = C.c_int
    lib.va_create.argtypes = [C.POINTER(vp), C.POINTER(VaConfig)]
    lib.va_destroy.argtypes = [vp]
    lib.va_destroy.restype = None
    lib.va_last_error.argtypes = [vp]
    lib.va_last_error.restype = C.c_char_p
    lib.va_get_layout.argtypes = [vp, C.POINTER(VaLayout)]
    lib.va_assemble_masks.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp]
    lib.va_run_fused.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp]
    lib.va_run_fused_host.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp]
    lib.va_run_fused_host_f16.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp]
    lib.va_mask_to_records.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp]
    lib.va_grid_to_penalty_peaks.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, vp, vp]
    lib.va_nms.argtypes = [vp, vp, i32, C.POINTER(VaNmsParams), i32, vp, vp, vp, vp, vp, vp]
    lib.va_scale_boxes.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]
    lib.va_last_launch_count.argtypes = [vp]
    lib.va_uses_tensor_core.argtypes = [vp]
    lib.va_profile_enable.argtypes = [vp, i32]
    lib.va_profile_read.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(i32)]
    lib.va_peer_alloc.argtypes = [vp, C.c_uint64, C.POINTER(vp), C.c_char_p]
    lib.va_peer_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    lib.va_peer_close.argtypes = [vp, vp]
    lib.va_peer_free.argtypes = [vp, vp]
    lib.va_peer_put.argtypes = [vp, vp, vp, C.c_uint64, vp]
    lib.va_signal.argtypes = [vp, vp, i32, vp]
    lib.va_wait_flags.argtypes = [vp, vp, i32, i32, vp]
    for name in EXPORTS:
        getattr(lib, name)  # raises AttributeError if a declared symbol is not exported
    _lib = lib
    return lib
