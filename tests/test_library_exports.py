"""The C-ABI shared library loads without a GPU and exports every symbol include/*.h declares (no compute calls here)."""
import os
import re

import pytest

import open_whisper_kit_b200 as pkg
from open_whisper_kit_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, marker):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#"))      # drop the macro definitions
    return set(re.findall(marker + r"[^;(]*?\b([a-z_0-9]+)\s*\(", text, re.S))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(pkg.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return pkg.load()


def test_exports_every_declared_symbol(lib):
    names = _declared("whisper.h", "WHISPER_API") | _declared("whisper_b200.h", "WB200_API") | {"whisper_token_count", "ggml_backend_load_all"}
    assert len(names) > 120
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_defaults_match_reference_values(lib):
    """whisper_full_default_params / whisper_context_default_params (reference src/whisper.cpp:3606-3622, 5928-6034)."""
    g = lib.whisper_full_default_params(capi.GREEDY)
    assert (g.strategy, g.greedy.best_of, g.beam_search.beam_size) == (0, 5, -1)
    assert g.n_max_text_ctx == 16384 and g.no_context and not g.no_timestamps and g.suppress_blank and not g.suppress_nst
    assert abs(g.temperature_inc - 0.2) < 1e-7 and abs(g.entropy_thold - 2.4) < 1e-6 and g.logprob_thold == -1.0
    assert abs(g.no_speech_thold - 0.6) < 1e-6 and g.max_initial_ts == 1.0 and g.length_penalty == -1.0
    assert g.language == b"en" and g.print_progress and g.print_timestamps and abs(g.thold_pt - 0.01) < 1e-8
    assert abs(g.grammar_penalty - 100.0) < 1e-6 and abs(g.vad_params.threshold - 0.5) < 1e-7
    b = lib.whisper_full_default_params(capi.BEAM_SEARCH)
    assert (b.strategy, b.greedy.best_of, b.beam_search.beam_size) == (1, -1, 5)
    c = lib.whisper_context_default_params()
    assert c.use_gpu and c.flash_attn and c.gpu_device == 0 and c.dtw_n_top == -1 and c.dtw_mem_size == 128 * 1024 * 1024


def test_language_table(lib):
    assert lib.whisper_lang_max_id() == 99
    assert lib.whisper_lang_id(b"en") == 0 and lib.whisper_lang_id(b"german") == 2 and lib.whisper_lang_id(b"yue") == 99
    assert lib.whisper_lang_str(7) == b"ja" and lib.whisper_lang_str_full(80) == b"haitian creole"
    assert lib.whisper_lang_id(b"klingon") == -1 and lib.whisper_lang_str(1000) is None


def test_fails_loudly_without_a_gpu(lib, tmp_path):
    """No CPU fallback: without a CUDA device the context is refused (NULL), it is never emulated on the host."""
    if lib.whisper_b200_device_count() > 0:
        pytest.skip("a GPU is visible")
    from open_whisper_kit_b200 import modelgen
    p = tmp_path / "m.bin"
    modelgen.write_model(str(p), "micro.en")
    assert lib.whisper_init_from_file_with_params(str(p).encode(), lib.whisper_context_default_params()) is None
    import ctypes as C
    n_len, n_org = C.c_int(), C.c_int()
    import numpy as np
    pcm = np.zeros(1600, np.float32)
    filt = modelgen.mel_filters(80)
    out = np.zeros((80, 3010), np.float32)
    FP = C.POINTER(C.c_float)
    rc = lib.whisper_b200_kernel_log_mel(pcm.ctypes.data_as(FP), 1600, filt.ctypes.data_as(FP), 80, out.ctypes.data_as(FP),
                                         out.size, C.byref(n_len), C.byref(n_org))
    assert rc != 0
