"""ctypes binding of libdistilcodec_b200.so (include/distilcodec_b200.h).

The library is the product path; there is NO fallback: if it is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DC_LIB") or os.path.join(HERE, "libdistilcodec_b200.so")  # DC_LIB: experiment builds

DC_OK = 0
MODE_FP32, MODE_BF16 = 0, 1
STAGE_ENCODER, STAGE_QUANTIZER, STAGE_DECODE_CODES, STAGE_GENERATOR = 0, 1, 2, 3
ACT_NONE, ACT_GELU, ACT_SILU = 0, 1, 2


class DcConfig(C.Structure):
    """dc_config: the fields of configs/model_config.json the hot path depends on."""
    _fields_ = [("n_mels", C.c_int), ("enc_depths", C.c_int * 4), ("enc_dims", C.c_int * 4),
                ("codebook_size", C.c_int), ("codebook_dim", C.c_int), ("n_ups", C.c_int),
                ("up_rates", C.c_int * 8), ("up_kernels", C.c_int * 8), ("up_initial_channel", C.c_int),
                ("rb_kernels", C.c_int * 3), ("rb_dilations", C.c_int * 3), ("pre_kernel", C.c_int),
                ("post_kernel", C.c_int)]


class DcAudioInfo(C.Structure):
    """dc_audio_info"""
    _fields_ = [("sample_rate", C.c_int), ("channels", C.c_int), ("bits_per_sample", C.c_int), ("is_float", C.c_int),
                ("frames", C.c_int64)]


class DcProfileRow(C.Structure):
    """dc_profile_row: per-kernel-class device time and algorithmic work."""
    _fields_ = [("name", C.c_char * 64), ("launches", C.c_uint64), ("ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double)]


_vp, _i, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/distilcodec_b200.h declares
SIGNATURES = {
    "dc_version": (_i, []),
    "dc_last_error": (C.c_char_p, []),
    "dc_default_config": (_i, [C.POINTER(DcConfig)]),
    "dc_create": (_i, [_i, _i, C.POINTER(DcConfig), C.POINTER(_vp)]),
    "dc_destroy": (_i, [_vp]),
    "dc_set_option": (_i, [_vp, C.c_char_p, C.c_double]),
    "dc_set_tensor": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i]),
    "dc_finalize": (_i, [_vp, _vp]),
    "dc_workspace_bytes": (_i, [_vp, _i, _i, _i, C.POINTER(_sz)]),
    "dc_encoder_forward": (_i, [_vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "dc_quantizer_forward": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dc_quantizer_encode": (_i, [_vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "dc_vq_search": (_i, [_vp, _vp, _i, _vp, _i64, _vp, _vp, _sz, _vp, C.POINTER(_i)]),
    "dc_vq_workspace_bytes": (_i, [_vp, _i64, _i, C.POINTER(_sz)]),
    "dc_quantizer_decode": (_i, [_vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "dc_generator_forward": (_i, [_vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "dc_mel_forward": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "dc_copy2d_async": (_i, [_vp, _sz, _vp, _sz, _sz, _sz, _vp]),
    "dc_conv_post_toeplitz_weights": (_i, [_vp, _vp]),
    "dc_audio_probe": (_i, [C.c_char_p, C.POINTER(DcAudioInfo)]),
    "dc_audio_resampled_length": (_i, [_i64, _i, _i, C.POINTER(_i64)]),
    "dc_audio_resample": (_i, [_vp, _i64, _i, _i, _i, C.c_double, _vp, _i64, C.POINTER(_i64), _i]),
    "dc_audio_load": (_i, [C.c_char_p, _i, _i, C.c_double, _i64, _i64, _vp, _i64, C.POINTER(_i64), C.POINTER(_i)]),
    "dc_audio_load_batch": (_i, [C.POINTER(C.c_char_p), _i, _i, _i, C.c_double, _vp, _i64, _i64, _vp,
                                 C.POINTER(_i), _i]),
    "dc_audio_write_wav": (_i, [C.c_char_p, _vp, _i64, _i]),
    "dc_ncl_to_nlc": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "dc_nlc_to_ncl": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "dc_op_conv_gemm": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dc_op_dwconv_ln": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "dc_profile_enable": (_i, [_i]),
    "dc_profile_collect": (_i, [C.POINTER(DcProfileRow), _i, C.POINTER(_i)]),
    "dc_launch_count": (C.c_uint64, []),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and type every entry point.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m distilcodec_nabeel_b200.build` "
            "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != DC_OK:
        msg = load().dc_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libdistilcodec_b200 {what} failed (status {rc}): {msg}")
