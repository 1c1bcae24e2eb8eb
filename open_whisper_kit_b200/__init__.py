"""open_whisper_kit_b200 -- B200-native (sm_100a) batched-transcription hot path behind the whisper.h C ABI.

The product is the shared library lib/libwhisper.so (C++ host + hand-written CUDA).  This package only
locates / builds it and binds the C ABI with ctypes.  There is no CPU fallback: if the library is missing or
no CUDA device is visible, loading or the first compute call fails loudly.
"""
import ctypes as _C
import os as _os

from . import capi

_HERE = _os.path.dirname(_os.path.abspath(__file__))
LIB_PATH = _os.path.join(_HERE, "lib", "libwhisper.so")

_FP = _C.POINTER(_C.c_float)
_U16P = _C.POINTER(_C.c_uint16)
_IP = _C.POINTER(_C.c_int)

# extension entry points of include/whisper_b200.h
EXT_PROTOTYPES = {
    "whisper_b200_vad_segments_from_probs": (_C.c_int, [_FP, _C.c_int, _C.c_int, capi.whisper_vad_params,
                                                        _C.POINTER(_C.c_longlong), _C.c_int]),
    "whisper_b200_vad_filter": (_C.c_int, [_C.c_void_p, capi.whisper_full_params, _FP, _C.c_int, _FP, _C.c_int,
                                           _C.POINTER(_C.c_longlong), _C.c_int, _IP]),
    "whisper_b200_vad_map_time": (_C.c_longlong, [_C.POINTER(_C.c_longlong), _C.c_int, _C.c_longlong]),
    "whisper_b200_device_count": (_C.c_int, []),
    "whisper_b200_kernel_log_mel": (_C.c_int, [_FP, _C.c_int, _FP, _C.c_int, _FP, _C.c_int, _IP, _IP]),
    "whisper_b200_kernel_log_mel_bench": (_C.c_double, [_C.c_int, _C.c_int, _FP, _C.c_int, _C.c_int, _C.c_int]),
    "whisper_b200_kernel_gemm": (_C.c_int, [_C.c_int, _C.c_int, _C.c_int, _C.c_int, _U16P, _U16P, _FP, _C.c_float,
                                            _C.c_int, _C.c_int, _FP, _C.c_int, _FP, _U16P, _FP]),
    "whisper_b200_kernel_skinny_gemm": (_C.c_int, [_C.c_int, _C.c_int, _C.c_int, _C.c_int, _U16P, _U16P, _FP, _C.c_float,
                                                   _C.c_int, _C.c_int, _FP, _U16P, _FP]),
    "whisper_b200_kernel_tc_skinny_gemm": (_C.c_int, [_C.c_int, _C.c_int, _C.c_int, _C.c_int, _U16P, _U16P, _FP, _C.c_float,
                                                      _C.c_int, _C.c_int, _FP, _U16P, _FP]),
    "whisper_b200_kernel_ln_gemm_pair": (_C.c_int, [_C.c_int] * 5 + [_U16P, _U16P, _FP, _FP, _FP, _FP, _C.c_float, _U16P, _FP, _FP]),
    "whisper_b200_kernel_gemm_bench": (_C.c_double, [_C.c_int, _C.c_int, _C.c_int, _C.c_int, _C.c_int, _C.c_int]),
    "whisper_b200_kernel_step_bench": (_C.c_double, [_C.c_int] * 6),
    "whisper_b200_kernel_self_attn": (_C.c_int, [_C.c_int] * 4 + [_IP, _U16P, _U16P, _C.c_int, _C.c_int, _U16P, _U16P]),
    "whisper_b200_full_device": (_C.c_int, [_C.c_void_p, capi.whisper_full_params, _C.c_void_p, _C.c_int, _C.c_int]),
    "whisper_b200_get_mel": (_C.c_int, [_C.c_void_p, _C.c_void_p, _FP, _C.c_int, _IP, _IP]),
    "whisper_b200_get_encoder_output": (_C.c_int, [_C.c_void_p, _FP, _C.c_int]),
    "whisper_b200_get_cross_kv": (_C.c_int, [_C.c_void_p, _C.c_int, _U16P, _C.c_int]),
    "whisper_b200_dtype": (_C.c_int, [_C.c_void_p]),
    "whisper_b200_kernel_launches": (_C.c_longlong, [_C.c_void_p]),
    "whisper_b200_profile_enable": (None, [_C.c_void_p, _C.c_int]),
    "whisper_b200_profile_read": (_C.c_int, [_C.c_void_p, _C.POINTER(_C.c_double), _C.c_int]),
    "whisper_b200_chain_geometry": (_C.c_int, [_C.c_int] * 6 + [_IP]),
    "whisper_b200_dequantize_blocks": (_C.c_longlong, [_C.c_int, _C.c_void_p, _C.c_longlong, _U16P]),
    "whisper_b200_pcm16_to_mel": (_C.c_int, [_C.c_void_p, _C.POINTER(_C.c_int16), _C.c_int]),
    "whisper_b200_full_parallel_i16": (_C.c_int, [_C.c_void_p, capi.whisper_full_params, _C.POINTER(_C.c_int16), _C.c_int, _C.c_int]),
    "whisper_b200_dtw_align": (_C.c_int, [_FP] + [_C.c_int] * 6 + [_IP]),
    "whisper_b200_kv_copy": (_C.c_int, [_C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_int]),
    "whisper_b200_partition_owner": (_C.c_int, [_C.c_int, _C.c_int, _C.c_int]),
    "whisper_b200_group_init_from_file": (_C.c_void_p, [_C.c_char_p, capi.whisper_context_params, _IP, _C.c_int]),
    "whisper_b200_group_free": (None, [_C.c_void_p]),
    "whisper_b200_group_size": (_C.c_int, [_C.c_void_p]),
    "whisper_b200_group_context": (_C.c_void_p, [_C.c_void_p, _C.c_int]),
    "whisper_b200_group_full_parallel": (_C.c_int, [_C.c_void_p, capi.whisper_full_params, _FP, _C.c_int, _C.c_int]),
    "whisper_b200_token_timestamps": (_C.c_int, [_C.c_void_p, _C.c_int, _C.c_int, _C.c_int, _FP, _C.c_int, _C.c_longlong, _C.c_longlong,
                                                 _C.c_void_p, _C.c_int, _C.c_float, _C.c_float, _C.POINTER(_C.c_longlong), _C.c_int, _C.c_int,
                                                 _C.POINTER(_C.c_longlong), _IP, _C.c_int]),
}

PROFILE_CLASSES = [("mel", "B"), ("im2col", "B"), ("gemm_conv", "flop"), ("layernorm", "B"), ("gemm_encoder", "flop"),
                   ("encoder_attention", "flop"), ("gemm_cross_kv", "flop"), ("decoder_misc", "B"), ("gemm_decoder", "B"),
                   ("self_attention", "B"), ("cross_attention", "B"), ("gemm_logits", "B"), ("sample", "B"),
                   ("layernorm_decoder", "B"), ("decoder_chain", "B")]

_lib = None


def load(strict_api=True):
    """Load lib/libwhisper.so and bind whisper.h + whisper_b200.h.  Raises OSError if it was not built."""
    global _lib
    if _lib is None:
        if not _os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA extension is required; there is no CPU fallback)")
        lib = _C.CDLL(LIB_PATH, mode=_C.RTLD_LOCAL)
        capi.bind(lib, EXT_PROTOTYPES, strict=True)
        capi.bind(lib, capi.PROTOTYPES, strict=strict_api)
        _lib = lib
    return _lib
