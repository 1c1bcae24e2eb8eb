"""TEST INFRASTRUCTURE ONLY -- loader for oracle/_ref/libwhisper_ref_<variant>.so.

The library is the UNMODIFIED reference CPU implementation (whisper.cpp + ggml-cpu) compiled by
oracle/Makefile.ref from the sources under /root/reference, plus the ref_* read-back hooks of
oracle/ref_harness.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product never does.
"""
import ctypes as C
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from open_whisper_kit_b200 import capi  # noqa: E402

_FP = C.POINTER(C.c_float)

REF_PROTOTYPES = {
    "ref_vad_segments_from_probs": (C.c_int, [_FP, C.c_int, C.c_int, capi.whisper_vad_params, C.POINTER(C.c_longlong), C.c_int]),
    "ref_vad_filter": (C.c_int, [C.c_void_p, capi.whisper_full_params, _FP, C.c_int, _FP, C.c_int, C.POINTER(C.c_longlong),
                                 C.c_int, C.POINTER(C.c_int)]),
    "ref_vad_map_time": (C.c_longlong, [C.POINTER(C.c_longlong), C.c_int, C.c_longlong]),
    "ref_mel_dims": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ref_mel_copy": (C.c_int, [C.c_void_p, _FP]),
    "ref_log_mel": (C.c_int, [_FP, C.c_int, C.c_int, _FP, C.c_int, _FP, C.c_int, C.POINTER(C.c_int),
                              C.POINTER(C.c_int)]),
    "ref_embd_enc_copy": (C.c_int, [C.c_void_p, _FP, C.c_int]),
    "ref_kv_cross_copy": (C.c_int64, [C.c_void_p, C.c_int, _FP, C.c_int64]),
    "ref_process_logits": (C.c_int, [C.c_void_p, capi.whisper_full_params, C.c_float, _FP,
                                     C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int, _FP, _FP, _FP,
                                     C.POINTER(capi.whisper_token_data)]),
    "ref_no_speech_prob": (C.c_float, [C.c_void_p]),
    "ref_sample_topk": (C.c_int, [C.c_void_p, C.c_int, C.c_uint, C.POINTER(capi.whisper_token_data)]),
    "ref_token_timestamps": (C.c_int, [C.c_void_p, _FP, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p, C.c_int, C.c_float,
                                       C.c_float, C.POINTER(C.c_longlong), C.c_int, C.c_int, C.POINTER(C.c_longlong),
                                       C.POINTER(C.c_int), C.c_int]),
}


def host_has_avx512():
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    need = ("avx512f", "avx512bw", "avx512vl", "avx512dq", "avx512cd", "avx512vbmi", "avx512_vnni")
    return all(n in flags for n in need)


def ref_library_path():
    variants = ["v4", "v3"] if host_has_avx512() else ["v3"]
    for v in variants:
        p = os.path.join(_HERE, "_ref", f"libwhisper_ref_{v}.so")
        if os.path.exists(p):
            return p, v
    return None, None


_cached = None


def load():
    """Return (lib, variant) or (None, None) when the reference library was not built."""
    global _cached
    if _cached is not None:
        return _cached
    path, variant = ref_library_path()
    if path is None:
        _cached = (None, None)
        return _cached
    lib = capi.load_library(path)
    capi.bind(lib, REF_PROTOTYPES)
    # silence the reference's INFO logging
    cb = capi.LOG_CB(lambda level, text, ud: None)
    lib._quiet_cb = cb
    if os.environ.get("REF_VERBOSE") is None:
        lib.whisper_log_set(C.cast(cb, C.c_void_p), None)
    _cached = (lib, variant)
    return _cached
