"""ctypes view of the C ABI in include/whisper.h (mirrors reference include/whisper.h:116-151, 487-591).

The same bindings are used against our libwhisper.so (the product) and, in tests / the CPU-baseline
leg of bench.py only, against oracle/_ref/libwhisper_ref_*.so (the unmodified reference built for
the host CPU).  Struct layouts follow the header field-for-field; tests/test_abi.py checks
sizeof/offsetof against a C program compiled with the header.
"""
import ctypes as C
import os

c_bool = C.c_bool
whisper_token = C.c_int32

GREEDY = 0
BEAM_SEARCH = 1


class whisper_ahead(C.Structure):
    _fields_ = [("n_text_layer", C.c_int), ("n_head", C.c_int)]


class whisper_aheads(C.Structure):
    _fields_ = [("n_heads", C.c_size_t), ("heads", C.POINTER(whisper_ahead))]


class whisper_context_params(C.Structure):
    _fields_ = [
        ("use_gpu", c_bool),
        ("flash_attn", c_bool),
        ("gpu_device", C.c_int),
        ("dtw_token_timestamps", c_bool),
        ("dtw_aheads_preset", C.c_int),
        ("dtw_n_top", C.c_int),
        ("dtw_aheads", whisper_aheads),
        ("dtw_mem_size", C.c_size_t),
    ]


class whisper_token_data(C.Structure):
    _fields_ = [
        ("id", whisper_token),
        ("tid", whisper_token),
        ("p", C.c_float),
        ("plog", C.c_float),
        ("pt", C.c_float),
        ("ptsum", C.c_float),
        ("t0", C.c_int64),
        ("t1", C.c_int64),
        ("t_dtw", C.c_int64),
        ("vlen", C.c_float),
    ]


class whisper_vad_params(C.Structure):
    _fields_ = [
        ("threshold", C.c_float),
        ("min_speech_duration_ms", C.c_int),
        ("min_silence_duration_ms", C.c_int),
        ("max_speech_duration_s", C.c_float),
        ("speech_pad_ms", C.c_int),
        ("samples_overlap", C.c_float),
    ]


class whisper_vad_context_params(C.Structure):
    _fields_ = [("n_threads", C.c_int), ("use_gpu", C.c_bool), ("gpu_device", C.c_int)]


class _greedy(C.Structure):
    _fields_ = [("best_of", C.c_int)]


class _beam(C.Structure):
    _fields_ = [("beam_size", C.c_int), ("patience", C.c_float)]


NEW_SEGMENT_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)
PROGRESS_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)
ENCODER_BEGIN_CB = C.CFUNCTYPE(c_bool, C.c_void_p, C.c_void_p, C.c_void_p)
ABORT_CB = C.CFUNCTYPE(c_bool, C.c_void_p)
LOGITS_FILTER_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.POINTER(whisper_token_data), C.c_int,
                               C.POINTER(C.c_float), C.c_void_p)


class whisper_full_params(C.Structure):
    _fields_ = [
        ("strategy", C.c_int),
        ("n_threads", C.c_int),
        ("n_max_text_ctx", C.c_int),
        ("offset_ms", C.c_int),
        ("duration_ms", C.c_int),
        ("translate", c_bool),
        ("no_context", c_bool),
        ("no_timestamps", c_bool),
        ("single_segment", c_bool),
        ("print_special", c_bool),
        ("print_progress", c_bool),
        ("print_realtime", c_bool),
        ("print_timestamps", c_bool),
        ("token_timestamps", c_bool),
        ("thold_pt", C.c_float),
        ("thold_ptsum", C.c_float),
        ("max_len", C.c_int),
        ("split_on_word", c_bool),
        ("max_tokens", C.c_int),
        ("debug_mode", c_bool),
        ("audio_ctx", C.c_int),
        ("tdrz_enable", c_bool),
        ("suppress_regex", C.c_char_p),
        ("initial_prompt", C.c_char_p),
        ("carry_initial_prompt", c_bool),
        ("prompt_tokens", C.POINTER(whisper_token)),
        ("prompt_n_tokens", C.c_int),
        ("language", C.c_char_p),
        ("detect_language", c_bool),
        ("suppress_blank", c_bool),
        ("suppress_nst", c_bool),
        ("temperature", C.c_float),
        ("max_initial_ts", C.c_float),
        ("length_penalty", C.c_float),
        ("temperature_inc", C.c_float),
        ("entropy_thold", C.c_float),
        ("logprob_thold", C.c_float),
        ("no_speech_thold", C.c_float),
        ("greedy", _greedy),
        ("beam_search", _beam),
        ("new_segment_callback", C.c_void_p),
        ("new_segment_callback_user_data", C.c_void_p),
        ("progress_callback", C.c_void_p),
        ("progress_callback_user_data", C.c_void_p),
        ("encoder_begin_callback", C.c_void_p),
        ("encoder_begin_callback_user_data", C.c_void_p),
        ("abort_callback", C.c_void_p),
        ("abort_callback_user_data", C.c_void_p),
        ("logits_filter_callback", C.c_void_p),
        ("logits_filter_callback_user_data", C.c_void_p),
        ("grammar_rules", C.c_void_p),
        ("n_grammar_rules", C.c_size_t),
        ("i_start_rule", C.c_size_t),
        ("grammar_penalty", C.c_float),
        ("vad", c_bool),
        ("vad_model_path", C.c_char_p),
        ("vad_params", whisper_vad_params),
    ]


class whisper_timings(C.Structure):
    _fields_ = [("sample_ms", C.c_float), ("encode_ms", C.c_float), ("decode_ms", C.c_float),
                ("batchd_ms", C.c_float), ("prompt_ms", C.c_float)]


LOG_CB = C.CFUNCTYPE(None, C.c_int, C.c_char_p, C.c_void_p)

_P = C.c_void_p
_FP = C.POINTER(C.c_float)
_I = C.c_int

# name -> (restype, argtypes).  Every WHISPER_API prototype of include/whisper.h.
PROTOTYPES = {
    "whisper_version": (C.c_char_p, []),
    "whisper_init_from_file_with_params": (_P, [C.c_char_p, whisper_context_params]),
    "whisper_init_from_buffer_with_params": (_P, [_P, C.c_size_t, whisper_context_params]),
    "whisper_init_with_params": (_P, [_P, whisper_context_params]),
    "whisper_init_from_file_with_params_no_state": (_P, [C.c_char_p, whisper_context_params]),
    "whisper_init_from_buffer_with_params_no_state": (_P, [_P, C.c_size_t, whisper_context_params]),
    "whisper_init_with_params_no_state": (_P, [_P, whisper_context_params]),
    "whisper_init_from_file": (_P, [C.c_char_p]),
    "whisper_init_from_buffer": (_P, [_P, C.c_size_t]),
    "whisper_init": (_P, [_P]),
    "whisper_init_from_file_no_state": (_P, [C.c_char_p]),
    "whisper_init_from_buffer_no_state": (_P, [_P, C.c_size_t]),
    "whisper_init_no_state": (_P, [_P]),
    "whisper_init_state": (_P, [_P]),
    "whisper_ctx_init_openvino_encoder_with_state": (_I, [_P, _P, C.c_char_p, C.c_char_p, C.c_char_p]),
    "whisper_ctx_init_openvino_encoder": (_I, [_P, C.c_char_p, C.c_char_p, C.c_char_p]),
    "whisper_free": (None, [_P]),
    "whisper_free_state": (None, [_P]),
    "whisper_free_params": (None, [_P]),
    "whisper_free_context_params": (None, [_P]),
    "whisper_pcm_to_mel": (_I, [_P, _FP, _I, _I]),
    "whisper_pcm_to_mel_with_state": (_I, [_P, _P, _FP, _I, _I]),
    "whisper_set_mel": (_I, [_P, _FP, _I, _I]),
    "whisper_set_mel_with_state": (_I, [_P, _P, _FP, _I, _I]),
    "whisper_encode": (_I, [_P, _I, _I]),
    "whisper_encode_with_state": (_I, [_P, _P, _I, _I]),
    "whisper_decode": (_I, [_P, C.POINTER(whisper_token), _I, _I, _I]),
    "whisper_decode_with_state": (_I, [_P, _P, C.POINTER(whisper_token), _I, _I, _I]),
    "whisper_tokenize": (_I, [_P, C.c_char_p, C.POINTER(whisper_token), _I]),
    "whisper_token_count": (_I, [_P, C.c_char_p]),
    "whisper_lang_max_id": (_I, []),
    "whisper_lang_id": (_I, [C.c_char_p]),
    "whisper_lang_str": (C.c_char_p, [_I]),
    "whisper_lang_str_full": (C.c_char_p, [_I]),
    "whisper_lang_auto_detect": (_I, [_P, _I, _I, _FP]),
    "whisper_lang_auto_detect_with_state": (_I, [_P, _P, _I, _I, _FP]),
    "whisper_n_len": (_I, [_P]),
    "whisper_n_len_from_state": (_I, [_P]),
    "whisper_n_vocab": (_I, [_P]),
    "whisper_n_text_ctx": (_I, [_P]),
    "whisper_n_audio_ctx": (_I, [_P]),
    "whisper_is_multilingual": (_I, [_P]),
    "whisper_model_n_vocab": (_I, [_P]),
    "whisper_model_n_audio_ctx": (_I, [_P]),
    "whisper_model_n_audio_state": (_I, [_P]),
    "whisper_model_n_audio_head": (_I, [_P]),
    "whisper_model_n_audio_layer": (_I, [_P]),
    "whisper_model_n_text_ctx": (_I, [_P]),
    "whisper_model_n_text_state": (_I, [_P]),
    "whisper_model_n_text_head": (_I, [_P]),
    "whisper_model_n_text_layer": (_I, [_P]),
    "whisper_model_n_mels": (_I, [_P]),
    "whisper_model_ftype": (_I, [_P]),
    "whisper_model_type": (_I, [_P]),
    "whisper_get_logits": (_FP, [_P]),
    "whisper_get_logits_from_state": (_FP, [_P]),
    "whisper_token_to_str": (C.c_char_p, [_P, whisper_token]),
    "whisper_model_type_readable": (C.c_char_p, [_P]),
    "whisper_token_eot": (whisper_token, [_P]),
    "whisper_token_sot": (whisper_token, [_P]),
    "whisper_token_solm": (whisper_token, [_P]),
    "whisper_token_prev": (whisper_token, [_P]),
    "whisper_token_nosp": (whisper_token, [_P]),
    "whisper_token_not": (whisper_token, [_P]),
    "whisper_token_beg": (whisper_token, [_P]),
    "whisper_token_lang": (whisper_token, [_P, _I]),
    "whisper_token_translate": (whisper_token, [_P]),
    "whisper_token_transcribe": (whisper_token, [_P]),
    "whisper_get_timings": (C.POINTER(whisper_timings), [_P]),
    "whisper_print_timings": (None, [_P]),
    "whisper_reset_timings": (None, [_P]),
    "whisper_print_system_info": (C.c_char_p, []),
    "whisper_context_default_params_by_ref": (C.POINTER(whisper_context_params), []),
    "whisper_context_default_params": (whisper_context_params, []),
    "whisper_full_default_params_by_ref": (C.POINTER(whisper_full_params), [_I]),
    "whisper_full_default_params": (whisper_full_params, [_I]),
    "whisper_full": (_I, [_P, whisper_full_params, _FP, _I]),
    "whisper_full_with_state": (_I, [_P, _P, whisper_full_params, _FP, _I]),
    "whisper_full_parallel": (_I, [_P, whisper_full_params, _FP, _I, _I]),
    "whisper_full_n_segments": (_I, [_P]),
    "whisper_full_n_segments_from_state": (_I, [_P]),
    "whisper_full_lang_id": (_I, [_P]),
    "whisper_full_lang_id_from_state": (_I, [_P]),
    "whisper_full_get_segment_t0": (C.c_int64, [_P, _I]),
    "whisper_full_get_segment_t0_from_state": (C.c_int64, [_P, _I]),
    "whisper_full_get_segment_t1": (C.c_int64, [_P, _I]),
    "whisper_full_get_segment_t1_from_state": (C.c_int64, [_P, _I]),
    "whisper_full_get_segment_speaker_turn_next": (c_bool, [_P, _I]),
    "whisper_full_get_segment_speaker_turn_next_from_state": (c_bool, [_P, _I]),
    "whisper_full_get_segment_text": (C.c_char_p, [_P, _I]),
    "whisper_full_get_segment_text_from_state": (C.c_char_p, [_P, _I]),
    "whisper_full_n_tokens": (_I, [_P, _I]),
    "whisper_full_n_tokens_from_state": (_I, [_P, _I]),
    "whisper_full_get_token_text": (C.c_char_p, [_P, _I, _I]),
    "whisper_full_get_token_text_from_state": (C.c_char_p, [_P, _P, _I, _I]),
    "whisper_full_get_token_id": (whisper_token, [_P, _I, _I]),
    "whisper_full_get_token_id_from_state": (whisper_token, [_P, _I, _I]),
    "whisper_full_get_token_data": (whisper_token_data, [_P, _I, _I]),
    "whisper_full_get_token_data_from_state": (whisper_token_data, [_P, _I, _I]),
    "whisper_full_get_token_p": (C.c_float, [_P, _I, _I]),
    "whisper_full_get_token_p_from_state": (C.c_float, [_P, _I, _I]),
    "whisper_vad_default_params": (whisper_vad_params, []),
    "whisper_vad_default_context_params": (whisper_vad_context_params, []),
    "whisper_vad_init_from_file_with_params": (_P, [C.c_char_p, whisper_vad_context_params]),
    "whisper_vad_init_with_params": (_P, [_P, whisper_vad_context_params]),
    "whisper_vad_detect_speech": (c_bool, [_P, _FP, _I]),
    "whisper_vad_detect_speech_stateful": (c_bool, [_P, _FP, _I]),
    "whisper_vad_reset_state": (None, [_P]),
    "whisper_vad_n_probs": (_I, [_P]),
    "whisper_vad_probs": (_FP, [_P]),
    "whisper_vad_segments_from_probs": (_P, [_P, whisper_vad_params]),
    "whisper_vad_segments_from_samples": (_P, [_P, whisper_vad_params, _FP, _I]),
    "whisper_vad_segments_n_segments": (_I, [_P]),
    "whisper_vad_segments_get_segment_t0": (C.c_float, [_P, _I]),
    "whisper_vad_segments_get_segment_t1": (C.c_float, [_P, _I]),
    "whisper_vad_free_segments": (None, [_P]),
    "whisper_vad_free": (None, [_P]),
    "whisper_bench_memcpy": (_I, [_I]),
    "whisper_bench_memcpy_str": (C.c_char_p, [_I]),
    "whisper_bench_ggml_mul_mat": (_I, [_I]),
    "whisper_bench_ggml_mul_mat_str": (C.c_char_p, [_I]),
    "whisper_log_set": (None, [_P, _P]),
    "whisper_full_get_segment_no_speech_prob": (C.c_float, [_P, _I]),
    "whisper_full_get_segment_no_speech_prob_from_state": (C.c_float, [_P, _I]),
}


def bind(lib, prototypes=PROTOTYPES, strict=True):
    """Attach restype/argtypes for every prototype; raise if a symbol is missing (strict)."""
    missing = []
    for name, (res, args) in prototypes.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        if args is not None:
            fn.argtypes = args
    if missing and strict:
        raise OSError("library is missing C-ABI symbols: " + ", ".join(missing))
    return missing


def load_library(path, prototypes=PROTOTYPES, strict=True):
    if not os.path.exists(path):
        raise OSError(f"shared library not found: {path}")
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    bind(lib, prototypes, strict)
    return lib


def as_float_ptr(arr):
    return arr.ctypes.data_as(_FP)
