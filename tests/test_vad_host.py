"""Host stages of the VAD path on the CPU: the probability -> segment state machine (csrc/vad.cu::vad_segments_from_probs through
whisper_b200_vad_segments_from_probs) and the processed -> original time map (csrc/vad_api.cu::vad_map_time), against
  * the committed golden vectors the live reference produced (tests/golden/golden_vad.json), and
  * the reference's own functions, when oracle/_ref is built (src/whisper.cpp:5209-5420, 7947-7989)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

import open_whisper_kit_b200 as pkg
from oracle import reflib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import vad_cases  # noqa: E402

FP = C.POINTER(C.c_float)
LLP = C.POINTER(C.c_longlong)
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden_vad.json")))


def ours_segments(lib, probs, n_window, prm, cap=4096):
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    out = (C.c_longlong * (2 * cap))()
    n = lib.whisper_b200_vad_segments_from_probs(probs.ctypes.data_as(FP), len(probs), n_window, prm, out, cap)
    assert 0 <= n <= cap
    return [[out[2 * i], out[2 * i + 1]] for i in range(n)]


def ref_segments(ref, probs, n_window, prm, cap=4096):
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    out = (C.c_longlong * (2 * cap))()
    n = ref.ref_vad_segments_from_probs(probs.ctypes.data_as(FP), len(probs), n_window, prm, out, cap)
    assert 0 <= n <= cap
    return [[out[2 * i], out[2 * i + 1]] for i in range(n)]


def test_reference_known_answer_is_in_the_golden_file():
    # tests/test-vad.cpp of the reference: 344 probabilities and 4 segments on jfk.wav with the default parameters
    case = GOLDEN["cases"]["jfk"]
    assert len(case["probs"]) == 344 and len(case["segments"]["default"]) == 4


@pytest.mark.parametrize("signal", sorted(GOLDEN["cases"]))
@pytest.mark.parametrize("pname", sorted(vad_cases.PARAM_SETS))
def test_segments_from_golden_probs(signal, pname):
    lib = pkg.load()
    case = GOLDEN["cases"][signal]
    prm = vad_cases.vad_params(lib, **vad_cases.PARAM_SETS[pname])
    got = ours_segments(lib, case["probs"], 512, prm)
    assert got == [[int(a), int(b)] for a, b in case["segments"][pname]]


def _random_probs(rng, n, kind):
    if kind == 0:                       # smooth random walk through the threshold band
        x = np.cumsum(rng.standard_normal(n) * 0.25)
        return (1 / (1 + np.exp(-x))).astype(np.float32)
    if kind == 1:                       # on/off blocks with noisy edges
        p = np.zeros(n, np.float32)
        i = 0
        while i < n:
            ln = int(rng.integers(1, 60))
            p[i:i + ln] = rng.uniform(0.6, 1.0) if rng.random() < 0.5 else rng.uniform(0.0, 0.4)
            i += ln
        return np.clip(p + 0.1 * rng.standard_normal(n), 0, 1).astype(np.float32)
    if kind == 2:
        return rng.random(n).astype(np.float32)
    return np.full(n, 0.9 if kind == 3 else 0.1, np.float32)


@pytest.mark.parametrize("seed", range(40))
def test_segments_match_reference_function(seed):
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    lib = pkg.load()
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.choice([0, 1, 2, 7, 31, 200, 344, 1000, 3000]))
    probs = _random_probs(rng, n, seed % 5)
    prm = lib.whisper_vad_default_params()
    prm.threshold = float(rng.choice([0.5, 0.2, 0.35, 0.8, 0.1]))
    prm.min_speech_duration_ms = int(rng.choice([250, 0, 64, 1000]))
    prm.min_silence_duration_ms = int(rng.choice([100, 0, 32, 500, 2000]))
    prm.max_speech_duration_s = float(rng.choice([3.4e38, 1.0, 0.5, 5.0, 30.0]))
    prm.speech_pad_ms = int(rng.choice([30, 0, 200, 1000]))
    prm.samples_overlap = float(rng.choice([0.1, 0.0, 0.5]))
    n_window = int(rng.choice([512, 512, 256, 1536]))
    assert ours_segments(lib, probs, n_window, prm) == ref_segments(ref, probs, n_window, prm)


@pytest.mark.parametrize("seed", range(10))
def test_time_map_matches_reference_function(seed):
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    lib = pkg.load()
    rng = np.random.default_rng(8100 + seed)
    n = int(rng.integers(0, 40))
    proc = np.sort(rng.choice(5000, size=n, replace=False)).astype(np.int64) if n else np.zeros(0, np.int64)
    orig = np.sort(rng.integers(0, 20000, size=n)).astype(np.int64)
    tab = (C.c_longlong * max(2 * n, 2))()
    for i in range(n):
        tab[2 * i], tab[2 * i + 1] = int(proc[i]), int(orig[i])
    ts = list(rng.integers(-50, 5200, size=200)) + [int(x) for x in proc]
    for t in ts:
        assert lib.whisper_b200_vad_map_time(tab, n, int(t)) == ref.ref_vad_map_time(tab, n, int(t)), (n, int(t))


def test_time_map_of_an_empty_table_is_the_identity():
    lib = pkg.load()
    tab = (C.c_longlong * 2)()
    assert lib.whisper_b200_vad_map_time(tab, 0, 123) == 123


def test_vad_defaults():
    lib = pkg.load()
    cp = lib.whisper_vad_default_context_params()
    assert (cp.n_threads, cp.use_gpu, cp.gpu_device) == (4, False, 0)           # tests/test-vad.cpp:19-23
    p = lib.whisper_vad_default_params()
    assert p.threshold == 0.5 and p.min_speech_duration_ms == 250 and p.min_silence_duration_ms == 100
    assert p.speech_pad_ms == 30 and abs(p.samples_overlap - 0.1) < 1e-8 and p.max_speech_duration_s > 3e38


def test_vad_init_fails_loudly_without_model_or_device():
    lib = pkg.load()
    cp = lib.whisper_vad_default_context_params()
    assert not lib.whisper_vad_init_from_file_with_params(b"/nonexistent/vad.bin", cp)
    if lib.whisper_b200_device_count() == 0:            # no CPU fallback: a valid model still does not load without a GPU
        path = os.path.join(HERE, "golden", "silero-v6.2.0-ggml.bin").encode()
        assert not lib.whisper_vad_init_from_file_with_params(path, cp)
