"""ctypes binding of libaad_b200.so -- the C ABI declared in include/aad.h.

The product path has NO fallback: if the shared library is missing or fails to
load, importing anything that computes raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AAD_LIB_PATH") or os.path.join(_PKG_DIR, "libaad_b200.so")  # override: dev builds only

# enums (mirror include/aad.h)
KIND_LOGMEL, KIND_MFCC, KIND_LFCC, KIND_GTCC = 0, 1, 2, 3
F32, I16 = 0, 1
WIN_HANN_PERIODIC, WIN_HAMMING_SYMMETRIC = 0, 1
FB_MEL_SLANEY, FB_LINEAR_INTBIN, FB_LINEAR_CONT, FB_CUSTOM, FB_GAMMATONE, FB_CUSTOM_DENSE = 0, 1, 2, 3, 4, 5
LOG_DB10, LOG_LN, LOG_CBRT = 0, 1, 2
SPEC_POWER, SPEC_MAGNITUDE = 0, 1
REF_ONE, REF_UTT_MAX = 0, 1
LAYOUT_CT, LAYOUT_TC = 0, 1
TABLE_WINDOW, TABLE_FILTERBANK, TABLE_DCT, TABLE_DELTA_TAPS = 0, 1, 2, 3

ITEM_OK = 0
ITEM_STATUS_NAMES = {
    0: "ok", 1: "empty", 2: "too short for one frame", 3: "too short for delta",
    4: "output too small", 5: "non-finite audio",
}


class AadParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("kind", C.c_int32), ("sample_rate", C.c_int32),
        ("n_fft", C.c_int32), ("win_length", C.c_int32), ("hop_length", C.c_int32),
        ("window", C.c_int32), ("center", C.c_int32), ("quantize_i16", C.c_int32),
        ("pre_emph", C.c_float), ("n_filt", C.c_int32), ("fb_type", C.c_int32),
        ("fmin", C.c_float), ("fmax", C.c_float), ("power_scale", C.c_float),
        ("log_type", C.c_int32), ("ref_type", C.c_int32), ("amin", C.c_float),
        ("top_db", C.c_float), ("n_ceps", C.c_int32), ("n_delta", C.c_int32),
        ("delta_width", C.c_int32), ("layout", C.c_int32), ("time_mean", C.c_int32),
        ("i16_scale", C.c_float), ("znorm", C.c_int32), ("custom_fb", C.POINTER(C.c_float)),
        ("spectrum", C.c_int32), ("reserved0", C.c_int32),
    ]


class AadFlacInfo(C.Structure):
    """mirrors `aad_flac_info_t` (include/aad.h)"""
    _fields_ = [("sample_rate", C.c_int32), ("channels", C.c_int32), ("bits_per_sample", C.c_int32),
                ("total_samples", C.c_int64), ("md5", C.c_uint8 * 16)]


class AadDetectorWeights(C.Structure):
    """mirrors `struct aad_detector_weights` (include/aad.h)"""
    _FP = C.POINTER(C.c_float)
    _fields_ = [("struct_size", C.c_int32), ("feature_dim", C.c_int32),
                ("conv_w", _FP), ("conv_b", _FP), ("bn_w", _FP), ("bn_b", _FP), ("bn_mean", _FP), ("bn_var", _FP),
                ("bn_eps", C.c_float),
                ("w_ih", _FP), ("w_hh", _FP), ("b_ih", _FP), ("b_hh", _FP),
                ("w_ih_r", _FP), ("w_hh_r", _FP), ("b_ih_r", _FP), ("b_hh_r", _FP),
                ("attn_w", _FP), ("attn_b", _FP), ("ln_w", _FP), ("ln_b", _FP),
                ("fc1_w", _FP), ("fc1_b", _FP), ("fc2_w", _FP), ("fc2_b", _FP)]


# every symbol include/aad.h declares: name -> (restype, argtypes)
_EXTRACT_ARGS = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int64,
                 C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                 C.c_void_p]
SYMBOLS = {
    "aad_version": (C.c_int, []),
    "aad_strerror": (C.c_char_p, [C.c_int]),
    "aad_last_error_detail": (C.c_char_p, []),
    "aad_params_default": (C.c_int, [C.POINTER(AadParams), C.c_int, C.c_int]),
    "aad_plan_create": (C.c_int, [C.POINTER(AadParams), C.c_int, C.POINTER(C.c_void_p)]),
    "aad_plan_destroy": (C.c_int, [C.c_void_p]),
    "aad_query": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_int32),
                            C.POINTER(C.c_int32), C.POINTER(C.c_size_t)]),
    "aad_extract": (C.c_int, _EXTRACT_ARGS),
    "aad_extract_indexed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                      C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_size_t, C.c_void_p]),
    "aad_extract_pair": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "aad_logmel": (C.c_int, _EXTRACT_ARGS),
    "aad_mfcc": (C.c_int, _EXTRACT_ARGS),
    "aad_lfcc": (C.c_int, _EXTRACT_ARGS),
    "aad_gtcc": (C.c_int, _EXTRACT_ARGS),
    "aad_delta": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int32, C.c_int, C.c_int,
                            C.c_void_p, C.c_void_p]),
    "aad_db_reference": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int32,
                                   C.c_int, C.c_float, C.c_void_p]),
    "aad_detector_create": (C.c_int, [C.POINTER(AadDetectorWeights), C.c_int, C.POINTER(C.c_void_p)]),
    "aad_detector_destroy": (C.c_int, [C.c_void_p]),
    "aad_detector_query": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]),
    "aad_detector_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_void_p]),
    "aad_scaler_accumulate": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "aad_scaler_accumulate_ragged": (C.c_int, [C.c_void_p, C.c_int, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p]),
    "aad_scaler_apply": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "aad_extract_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int,
                                   C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_int]),
    "aad_cqcc_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "aad_cqcc_plan_destroy": (C.c_int, [C.c_void_p]),
    "aad_cqcc_query": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_size_t)]),
    "aad_cqcc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                           C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "aad_flac_info": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(AadFlacInfo)]),
    "aad_flac_decode": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "aad_flac_decode_pcm16": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int64, C.POINTER(C.c_int64),
                                        C.POINTER(C.c_int32)]),
    "aad_host_reserve": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int32, C.c_int]),
    "aad_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t, C.c_int]),
    "aad_host_free": (C.c_int, [C.c_void_p]),
    "aad_plan_table": (C.c_int64, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
    "aad_plan_launches": (C.c_int, [C.c_void_p]),
    "aad_plan_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "aad_plan_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "aad_fp32_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
}

_lib = None


class AadError(RuntimeError):
    pass


def load():
    """Load libaad_b200.so (once) and declare every prototype.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AadError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -m audioanalysisdetector_b200.build` (needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "aad call"):
    if rc != 0:
        lib = load()
        msg = lib.aad_strerror(rc).decode()
        detail = lib.aad_last_error_detail().decode() if rc == -3 else ""
        raise AadError(f"{what} failed: {msg} ({rc})" + (f" [{detail}]" if detail else ""))
