"""ctypes binding of the C-ABI CUDA library (include/linetr_b200.h).

There is no fallback: if the shared library has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or tools/build_native.sh)
importing the compute entry points raises, and every call raises on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "liblinetr_b200.so")
_lib = None

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)

LAYOUT_ROWS = 0
LAYOUT_CHANNEL_FIRST = 1

# Every symbol include/linetr_b200.h declares (tests check the library exports all of them).
EXPORTED_SYMBOLS = (
    "ltr_abi_version", "ltr_last_error", "ltr_create", "ltr_destroy", "ltr_encode_workspace_bytes",
    "ltr_desc_tiles_bytes", "ltr_encode", "ltr_match_workspace_bytes", "ltr_match", "ltr_gather_wait", "ltr_match_distmat",
    "ltr_merge_sublines", "ltr_tokenize", "ltr_linear", "ltr_linear_img", "ltr_linear_img_norm",
    "ltr_launch_count", "ltr_reset_launch_count", "ltr_profile_begin", "ltr_profile_end",
)
# include/linetr_b200_debug.h (micro-benchmark / tracing hooks; not part of the drop-in boundary)
DEBUG_SYMBOLS = ("ltr_gemm_bench", "ltr_gemm_trace", "ltr_debug_trace_arm", "ltr_debug_trace_read")
ABI_VERSION = 3


class LtrError(RuntimeError):
    pass


class LtrTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64)]


class LtrConfig(C.Structure):
    _fields_ = [("d_model", C.c_int32), ("n_heads", C.c_int32), ("d_inner", C.c_int32),
                ("n_desc_layers", C.c_int32), ("n_sig_layers", C.c_int32)]


class LtrEncodeInput(C.Structure):
    _fields_ = [("sublines", C.c_void_p), ("resp", C.c_void_p), ("angle", C.c_void_p),
                ("pnt", C.c_void_p), ("desc", C.c_void_p), ("score", C.c_void_p),
                ("cu_lines_host", C.c_void_p), ("cu_lines_dev", C.c_void_p),
                ("n_images", C.c_int32), ("n_lines", C.c_int32), ("n_tokens", C.c_int32),
                ("lines_per_image", C.c_int32), ("image_width", C.c_float), ("image_height", C.c_float)]


class LtrEncodeOutput(C.Structure):
    _fields_ = [("desc_cf", C.c_void_p), ("desc_rows", C.c_void_p), ("desc_tiles", C.c_void_p)]


class LtrMatchInput(C.Structure):
    _fields_ = [("desc0", C.c_void_p), ("desc1", C.c_void_p), ("layout", C.c_int32), ("d", C.c_int32),
                ("n_pairs", C.c_int32), ("n0", C.c_int32), ("n1", C.c_int32),
                ("cu0", C.c_void_p), ("cu1", C.c_void_p), ("sub_off0", C.c_void_p), ("sub_off1", C.c_void_p),
                ("cuk0", C.c_void_p), ("cuk1", C.c_void_p),
                ("max_n0", C.c_int32), ("max_n1", C.c_int32), ("max_k0", C.c_int32), ("max_k1", C.c_int32),
                ("dist_pair_stride", C.c_int64), ("nn_thresh", C.c_float), ("mutual", C.c_int32),
                ("total_n0", C.c_int32), ("total_n1", C.c_int32), ("dist_mode", C.c_int32),
                ("tiles0", C.c_void_p), ("tiles1", C.c_void_p), ("tiles_lines0", C.c_int32), ("tiles_lines1", C.c_int32),
                ("tiles_row0_0", C.c_int32), ("tiles_row0_1", C.c_int32)]


GATHER_SLOTS = 8


class LtrPeerGather(C.Structure):
    _fields_ = [("mc_base", C.c_void_p), ("peer_bases", C.c_void_p), ("rank", C.c_int32), ("world", C.c_int32),
                ("slot", C.c_int32), ("epoch", C.c_int32)]


class LtrMatchOutput(C.Structure):
    _fields_ = [("matches0", C.c_void_p), ("scores0", C.c_void_p), ("nn1", C.c_void_p),
                ("counts", C.c_void_p), ("dist_key", C.c_void_p), ("dist_sub", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("gather", C.POINTER(LtrPeerGather))]


class LtrTokenizeInput(C.Structure):
    _fields_ = [("sp", C.c_void_p), ("ep", C.c_void_p), ("ep_clipped", C.c_void_p), ("length", C.c_void_p),
                ("angle", C.c_void_p), ("n_tok", C.c_void_p), ("sub0", C.c_void_p), ("sub2line", C.c_void_p),
                ("n_keylines", C.c_int32), ("n_sublines", C.c_int32), ("n_tokens", C.c_int32),
                ("token_distance", C.c_double), ("dense_desc", C.c_void_p), ("desc_channels", C.c_int32),
                ("desc_h", C.c_int32), ("desc_w", C.c_int32), ("dense_score", C.c_void_p), ("score_h", C.c_int32),
                ("score_w", C.c_int32), ("align_corners", C.c_int32)]


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load the shared library (once).  Raises LtrError if it is missing - no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise LtrError(
            f"linetr_b200: native CUDA library not built ({_LIB_PATH}); run tools/build_native.sh "
            "or __graft_entry__.build().  There is no CPU/PyTorch fallback.")
    lib = C.CDLL(_LIB_PATH)
    lib.ltr_abi_version.restype = C.c_int
    lib.ltr_last_error.restype = C.c_char_p
    lib.ltr_create.argtypes = [C.POINTER(LtrTensor), C.c_int32, C.POINTER(LtrConfig), C.c_int32,
                               C.POINTER(C.c_void_p)]
    lib.ltr_create.restype = C.c_int
    lib.ltr_destroy.argtypes = [C.c_void_p]
    lib.ltr_destroy.restype = None
    lib.ltr_encode_workspace_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.ltr_encode_workspace_bytes.restype = C.c_int64
    lib.ltr_desc_tiles_bytes.argtypes = [C.c_int32]
    lib.ltr_desc_tiles_bytes.restype = C.c_int64
    lib.ltr_encode.argtypes = [C.c_void_p, C.POINTER(LtrEncodeInput), C.POINTER(LtrEncodeOutput), C.c_void_p,
                               C.c_int64, C.c_void_p]
    lib.ltr_encode.restype = C.c_int
    lib.ltr_match_workspace_bytes.argtypes = [C.POINTER(LtrMatchInput)]
    lib.ltr_match_workspace_bytes.restype = C.c_int64
    lib.ltr_match.argtypes = [C.POINTER(LtrMatchInput), C.POINTER(LtrMatchOutput), C.c_int32, C.c_void_p]
    lib.ltr_match.restype = C.c_int
    lib.ltr_gather_wait.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    lib.ltr_gather_wait.restype = C.c_int
    lib.ltr_match_distmat.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_float,
                                      C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                      C.c_void_p]
    lib.ltr_match_distmat.restype = C.c_int
    lib.ltr_merge_sublines.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                       C.c_void_p]
    lib.ltr_merge_sublines.restype = C.c_int
    lib.ltr_tokenize.argtypes = [C.POINTER(LtrTokenizeInput)] + [C.c_void_p] * 7 + [C.c_int32, C.c_void_p]
    lib.ltr_tokenize.restype = C.c_int
    lib.ltr_linear.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                               C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.ltr_linear.restype = C.c_int
    lib.ltr_linear_img.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                   C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_void_p]
    lib.ltr_linear_img.restype = C.c_int
    lib.ltr_linear_img_norm.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                        C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                        C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.ltr_linear_img_norm.restype = C.c_int
    lib.ltr_gemm_bench.argtypes = [C.c_int32] * 7
    lib.ltr_gemm_bench.restype = C.c_float
    lib.ltr_gemm_trace.restype = C.POINTER(C.c_uint64)
    lib.ltr_debug_trace_arm.argtypes = [C.c_int32]
    lib.ltr_debug_trace_arm.restype = None
    lib.ltr_debug_trace_read.argtypes = [C.POINTER(C.c_uint64)]
    lib.ltr_debug_trace_read.restype = C.c_int
    lib.ltr_launch_count.restype = C.c_int64
    lib.ltr_reset_launch_count.restype = None
    lib.ltr_profile_begin.restype = None
    lib.ltr_profile_end.argtypes = [C.POINTER(C.c_char_p), c_float_p, c_int32_p, C.c_int32]
    lib.ltr_profile_end.restype = C.c_int
    if lib.ltr_abi_version() != ABI_VERSION:
        raise LtrError("linetr_b200: ABI version mismatch between _native.py and the shared library")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().ltr_last_error()
        raise LtrError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().ltr_launch_count())


def reset_launch_count():
    load().ltr_reset_launch_count()


def profile_begin():
    load().ltr_profile_begin()


def profile_end():
    """-> {kernel_class: (ms_total, launches)} since profile_begin(); synchronises the device."""
    n = 32
    names = (C.c_char_p * n)()
    ms = (C.c_float * n)()
    cnt = (C.c_int32 * n)()
    k = load().ltr_profile_end(names, ms, cnt, n)
    return {names[i].decode(): (float(ms[i]), int(cnt[i])) for i in range(k)}
