"""ctypes binding of include/fs2_b200.h (libfs2_b200.so).  Plain pointers and sizes only.

The library is built in-tree by `build.py`; if it is missing this module raises -- there is no
Python/PyTorch fallback for the compute path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfs2_b200.so")

PREC_FP32 = 0
PREC_BF16 = 1
PREC_BF16X3 = 2
PREC_F16X2 = 3

FS2_OK = 0
ERR_NAMES = {-1: "FS2_ERR_INVALID", -2: "FS2_ERR_CUDA", -3: "FS2_ERR_STATE", -4: "FS2_ERR_UNSUPPORTED",
             -5: "FS2_ERR_MISSING_WEIGHT", -6: "FS2_ERR_NO_DEVICE"}


class Fs2Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "vocab", "d_model", "n_enc_layers", "n_dec_layers", "n_heads", "d_ffn", "ffn_k1", "ffn_k2", "vp_filter",
        "vp_kernel", "n_bins", "n_mel", "pn_dim", "pn_kernel", "pn_layers", "max_seq_len", "pitch_phoneme_level",
        "energy_phoneme_level")]


class WeightDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4),
                ("on_device", C.c_int32)]


class ProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int64), ("ms", C.c_double)]


_P = C.c_void_p
_I = C.c_int32
_F = C.c_float

# name -> (restype, argtypes): every symbol include/fs2_b200.h declares
SIGNATURES = {
    "fs2_create": (C.c_int, [C.POINTER(_P), C.POINTER(Dims), C.c_int]),
    "fs2_destroy": (None, [_P]),
    "fs2_last_error": (C.c_char_p, [_P]),
    "fs2_version": (C.c_char_p, []),
    "fs2_load_weights": (C.c_int, [_P, C.POINTER(WeightDesc), _I]),
    "fs2_share_weights": (C.c_int, [_P, _P]),
    "fs2_set_precision": (C.c_int, [_P, _I, _I]),
    "fs2_set_row_packing": (C.c_int, [_P, _I]),
    "fs2_set_mel_post_layout": (C.c_int, [_P, _I]),
    "fs2_set_upsampler": (C.c_int, [_P, _I]),
    "fs2_forward_stage1": (C.c_int, [_P, _P, _P, _I, _I, _F, _F, _F, _P, _P, _P, _P, _P, _P, C.POINTER(_I), _P]),
    "fs2_forward_stage1_async": (C.c_int, [_P, _P, _P, _I, _I, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fs2_forward_stage1_commit": (C.c_int, [_P, _I, _I]),
    "fs2_forward_stage2": (C.c_int, [_P, _I, _F, _F, _P, _P, _P, _P, _P, _P]),
    "fs2_forward_stage1_graph": (C.c_int, [_P, _P, _P, _I, _I, _I, _F, _F, _F, _P, _P, _P, _P, _P, _P, C.POINTER(_I), _P]),
    "fs2_forward_stage2_graph": (C.c_int, [_P, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P]),
    "fs2_graph_stats": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fs2_round_durations": (C.c_int, [_P, C.c_int64, _F, _P, _P]),
    "fs2_duration_scan": (C.c_int, [_P, _I, _I, _P, _P, C.POINTER(_I), _P]),
    "fs2_length_regulate": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "fs2_gaussian_upsample": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "fs2_mask_from_lengths": (C.c_int, [_P, _I, _I, _P, _P]),
    "fs2_pack_valid_rows": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "fs2_wav_to_int16": (C.c_int, [_P, _P, _I, C.c_int64, _F, _P, _P, _P]),
    "fs2_op_mel_encoder": (C.c_int, [_P, _I, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "fs2_op_sinusoid_table": (C.c_int, [_P, _I, _P, _P]),
    "fs2_op_embed_pe": (C.c_int, [_P, _P, _I, _I, _P, _P]),
    "fs2_op_fft_stack": (C.c_int, [_P, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P]),
    "fs2_op_variance_predictor": (C.c_int, [_P, _I, _P, _P, _I, _I, _P, _P]),
    "fs2_op_variance_embed": (C.c_int, [_P, _I, _P, _F, _P, _I, _I, _P, _P]),
    "fs2_op_mel_postnet": (C.c_int, [_P, _I, _P, _I, _I, _P, _P, _P]),
    "fs2_op_conv_gemm": (C.c_int, [_I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "fs2_op_attention": (C.c_int, [_I, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "fs2_profile_enable": (C.c_int, [_P, _I]),
    "fs2_profile_reset": (C.c_int, [_P]),
    "fs2_profile_read": (C.c_int, [_P, C.POINTER(ProfileEntry), _I, C.POINTER(_I)]),
    "fs2_launch_count": (C.c_int64, [_P]),
    "fs2_last_frame_count": (C.c_int64, [_P]),
}


class Fs2Library:
    """Loaded libfs2_b200.so with typed entry points."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it with `python smart-nar_fast_tts_b200/build.py` "
                "(there is no fallback implementation of the CUDA path)")
        self.path = path
        self.lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.lib, name)     # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def check(self, rc: int, handle: Optional[int] = None):
        if rc != FS2_OK:
            msg = self.fs2_last_error(handle)
            raise Fs2Error(rc, msg.decode() if msg else "")


_LIB: Optional[Fs2Library] = None


def load_library() -> Fs2Library:
    global _LIB
    if _LIB is None:
        _LIB = Fs2Library()
    return _LIB
