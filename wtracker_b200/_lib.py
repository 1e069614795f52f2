"""ctypes binding of the C ABI declared in ``include/wtracker_b200.h``.

The shared library is built in-tree by ``wtracker_b200.build`` (``__graft_entry__.build()``).
There is no CPU fallback: if the library is missing, :func:`lib` raises, and every compute entry
point returns an error on a machine without an sm_100 device.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# WTRACKER_B200_LIB points at another build of the same library (the tuning build of tools/gpu_r2c.sh)
LIB_PATH = Path(os.environ.get("WTRACKER_B200_LIB") or Path(__file__).resolve().parent / "_native" / "libwtracker_b200.so")

# ---- enums (mirror the header) --------------------------------------------------------------
WT_OP_CONV0, WT_OP_CONV, WT_OP_SPPF_POOL = 0, 1, 2
WT_ACT_NONE, WT_ACT_SILU = 0, 1
WT_DT_BF16, WT_DT_F32, WT_DT_U8 = 0, 1, 2


ABI_VERSION = 12  # WT_ABI_VERSION of include/wtracker_b200.h this binding was written for

class WtLetterbox(C.Structure):
    _fields_ = [
        ("src_w", C.c_int32), ("src_h", C.c_int32),
        ("dst_w", C.c_int32), ("dst_h", C.c_int32),
        ("new_w", C.c_int32), ("new_h", C.c_int32),
        ("pad_left", C.c_int32), ("pad_top", C.c_int32),
        ("xofs", C.c_void_p), ("xcoef", C.c_void_p),
        ("yofs", C.c_void_p), ("ycoef", C.c_void_p),
    ]


class WtBuf(C.Structure):
    _fields_ = [("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32), ("dtype", C.c_int32)]


class WtOp(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("src", C.c_int32), ("src_coff", C.c_int32),
        ("dst", C.c_int32), ("dst_coff", C.c_int32),
        ("res", C.c_int32), ("res_coff", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32),
        ("k", C.c_int32), ("stride", C.c_int32),
        ("act", C.c_int32),
        ("w_off", C.c_int64), ("b_off", C.c_int64), ("dot_off", C.c_int64),
        ("add_buf", C.c_int32), ("add_coff", C.c_int32),
        ("lane", C.c_int32),
        ("chain_act", C.c_int32),
        ("chain_w_off", C.c_int64), ("chain_b_off", C.c_int64),
        ("cat_buf", C.c_int32), ("cat_coff", C.c_int32), ("cat_c", C.c_int32), ("chain_cout", C.c_int32),
    ]


class WtHeadLevel(C.Structure):
    _fields_ = [
        ("box", C.c_void_p), ("cls_feat", C.c_void_p), ("cls_logit", C.c_void_p),
        ("h", C.c_int32), ("w", C.c_int32), ("stride", C.c_int32),
        ("box_dtype", C.c_int32), ("cls_c", C.c_int32),
        ("cls_w", C.c_void_p), ("cls_b", C.c_float),
        ("box_feat", C.c_void_p), ("box_w", C.c_void_p), ("box_b", C.c_void_p), ("box_c", C.c_int32),
    ]


class WtPostParams(C.Structure):
    _fields_ = [
        ("conf_thres", C.c_float), ("iou_thres", C.c_float), ("max_det", C.c_int32),
        ("net_w", C.c_int32), ("net_h", C.c_int32),
        ("img_w", C.c_int32), ("img_h", C.c_int32),
        ("gain", C.c_float), ("pad_x", C.c_float), ("pad_y", C.c_float),
    ]


class WtResmlpDesc(C.Structure):
    _fields_ = [
        ("in_dim", C.c_int32), ("hidden", C.c_int32), ("out_dim", C.c_int32),
        ("n_blocks", C.c_int32), ("block_len", C.c_int32),
        ("block_dims", C.c_int32 * 8),
        ("weights", C.c_void_p), ("n_weights", C.c_int32),
    ]


WT_TAIL_MAX_K = 16


class WtTailArgs(C.Structure):
    _fields_ = [
        ("boxes", C.c_void_p), ("count", C.c_void_p), ("max_det", C.c_int32),
        ("crop_x", C.c_void_p), ("crop_y", C.c_void_p),
        ("cam_w", C.c_int32), ("cam_h", C.c_int32), ("mic_w", C.c_int32), ("mic_h", C.c_int32),
        ("table", C.c_void_p), ("mic_table", C.c_void_p),
        ("table_rows", C.c_int64), ("first_row", C.c_int64), ("n", C.c_int64),
        ("k", C.c_int32), ("offsets", C.c_int32 * WT_TAIL_MAX_K),
        ("mlp", WtResmlpDesc),
        ("weights_t", C.c_void_p), ("x", C.c_void_p), ("valid", C.c_void_p), ("y", C.c_void_p), ("err", C.c_void_p),
    ]


# every symbol include/wtracker_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "wt_last_error": (C.c_char_p, []),
    "wt_abi_version": (C.c_int, []),
    "wt_launch_count": (C.c_uint64, []),
    "wt_device_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "wt_preprocess": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                C.POINTER(WtLetterbox), C.c_void_p, C.c_void_p, C.c_void_p]),
    "wt_engine_workspace_bytes": (C.c_int64, [C.POINTER(WtBuf), C.c_int, C.c_int]),
    "wt_engine_create": (C.c_int, [C.POINTER(WtBuf), C.c_int, C.POINTER(WtOp), C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                   C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "wt_engine_destroy": (None, [C.c_void_p]),
    "wt_engine_buffer": (C.c_void_p, [C.c_void_p, C.c_int]),
    "wt_engine_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wt_post_scratch_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "wt_decode_nms": (C.c_int, [C.POINTER(WtHeadLevel), C.c_int, C.c_int, C.POINTER(WtPostParams), C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "wt_track_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "wt_resmlp_forward": (C.c_int, [C.POINTER(WtResmlpDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "wt_mlp_gather": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_int64, C.c_void_p]),
    "wt_hot_tail": (C.c_int, [C.POINTER(WtTailArgs), C.c_void_p]),
    "wt_result_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "wt_bbox_error": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "wt_mse_error": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "wt_analysis_columns": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wt_analysis_masks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_double,
                                    C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wt_precise_error": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_double, C.c_void_p, C.c_int64, C.c_void_p]),
    "wt_log_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                              C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wt_selftest_conv": (C.c_int, [C.c_int] * 11 + [C.POINTER(C.c_double)]),
    "wt_selftest_conv_chain": (C.c_int, [C.c_int] * 8 + [C.POINTER(C.c_double)]),
    "wt_selftest_conv_cat": (C.c_int, [C.c_int] * 4 + [C.POINTER(C.c_double)]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads (once) and returns the native library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NativeLibraryError(
                f"{LIB_PATH} is missing: build it with `python -m wtracker_b200.build` "
                "(there is no CPU fallback for the wtracker_b200 hot path)"
            )
        handle = C.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here == header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.wt_abi_version() != ABI_VERSION:
            raise NativeLibraryError("ABI version mismatch between _lib.py and libwtracker_b200.so")
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().wt_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"wtracker_b200 native call failed{(' in ' + what) if what else ''}: {msg}")


def launch_count() -> int:
    return int(lib().wt_launch_count())
