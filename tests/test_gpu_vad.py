"""Voice activity detection on the GPU (csrc/vad.cu, csrc/vad_api.cu) against the reference.

  * speech probabilities of the Silero model vs the golden vectors the live reference produced (tests/golden/golden_vad.json;
    jfk.wav is the reference's own known-answer case: 344 probabilities, 4 segments -- tests/test-vad.cpp) and vs the live
    reference on the box; the model is f16 weights with f32 accumulation on both sides, so probabilities agree to ~1e-3 and
    the segments (threshold crossings on 32 ms chunks) are identical;
  * stateful detection in pieces == one shot; reset;
  * the audio filter of whisper_full(vad=true) and its time table vs the reference's (ref_vad_filter), exact;
  * whisper_full(vad=true) == whisper_full on the filtered audio with the times mapped through the table."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

from open_whisper_kit_b200 import api, modelgen
from oracle import reflib

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import vad_cases  # noqa: E402

VAD_MODEL = os.path.join(HERE, "golden", "silero-v6.2.0-ggml.bin")
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden_vad.json")))
PROB_TOL = 5e-3
FP = C.POINTER(C.c_float)


def signal(name):
    return vad_cases.read_wav(os.path.join(HERE, "golden", "jfk.wav")) if name == "jfk" else vad_cases.synthetic_bursts()


@pytest.mark.parametrize("name", ["jfk", "bursts"])
def test_probabilities_and_segments_vs_golden(lib, name):
    case = GOLDEN["cases"][name]
    pcm = signal(name)
    assert len(pcm) == case["n_samples"]
    with api.Vad(lib, VAD_MODEL) as v:
        probs = v.detect(pcm)
        want = np.array(case["probs"], np.float32)
        assert len(probs) == len(want)
        d = np.abs(probs - want)
        print(f"{name}: {len(probs)} chunks, |p - reference| max {d.max():.2e} mean {d.mean():.2e}")
        assert d.max() < PROB_TOL
        for pname, kw in vad_cases.PARAM_SETS.items():
            got = v.segments_from_probs(vad_cases.vad_params(lib, **kw))
            assert [list(s) for s in got] == case["segments"][pname], pname
        # one call from samples == detect + from_probs
        assert [list(s) for s in v.segments_from_samples(pcm)] == case["segments"]["default"]


def test_probabilities_vs_live_reference(lib):
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    rng = np.random.default_rng(77)
    jfk = signal("jfk")
    cases = [jfk[:512], jfk[:513], jfk[:100], jfk[3000:3000 + 16000 * 3 + 77],
             np.concatenate([jfk[40000:90000], np.zeros(7000, np.float32), 0.3 * jfk[10000:50000]]),
             (0.05 * rng.standard_normal(20000)).astype(np.float32)]
    with api.Vad(lib, VAD_MODEL) as ours, api.Vad(ref, VAD_MODEL) as theirs:
        for pcm in cases:
            a, b = ours.detect(pcm), theirs.detect(pcm)
            assert len(a) == len(b) == (len(pcm) + 511) // 512
            assert np.abs(a - b).max() < PROB_TOL
            assert ours.segments_from_probs() == theirs.segments_from_probs()


def test_stateful_detection_in_pieces_equals_one_shot(lib):
    pcm = signal("bursts")
    with api.Vad(lib, VAD_MODEL) as v:
        whole = v.detect(pcm)
        v.reset()
        cuts = [0, 512 * 40, 512 * 41, 512 * 200, 512 * 450, len(pcm)]
        parts = [v.detect(pcm[a:b], stateful=True) for a, b in zip(cuts, cuts[1:])]
        got = np.concatenate(parts)
        assert len(got) == len(whole)
        # every chunk is padded by reflection of itself, so nothing but the LSTM state links chunks: cut at chunk multiples the
        # pieces reproduce the one-shot run bit for bit
        assert np.array_equal(got, whole)
        # the non-stateful call resets: same audio, same result, twice
        assert np.array_equal(v.detect(pcm), whole) and np.array_equal(v.detect(pcm), whole)
        assert len(v.detect(pcm[:0])) == 0


def test_stateful_pieces_vs_live_reference(lib):
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    pcm = signal("jfk")
    cuts = [0, 512 * 100, 512 * 101, 512 * 250, len(pcm)]
    with api.Vad(lib, VAD_MODEL) as ours, api.Vad(ref, VAD_MODEL) as theirs:
        ours.reset(), theirs.reset()
        for a, b in zip(cuts, cuts[1:]):
            pa, pb = ours.detect(pcm[a:b], stateful=True), theirs.detect(pcm[a:b], stateful=True)
            assert len(pa) == len(pb) and np.abs(pa - pb).max() < PROB_TOL


def _filter(L, fn, ctx, params, pcm):
    cap, cap_pairs = len(pcm) + 16000, 4096
    out = np.zeros(cap, np.float32)
    tab = (C.c_longlong * (2 * cap_pairs))()
    n_pairs = C.c_int(0)
    n = getattr(L, fn)(ctx, params, pcm.ctypes.data_as(FP), len(pcm), out.ctypes.data_as(FP), cap, tab, cap_pairs, C.byref(n_pairs))
    assert 0 <= n <= cap and n_pairs.value <= cap_pairs
    return out[:n].copy(), [(tab[2 * i], tab[2 * i + 1]) for i in range(n_pairs.value)]


@pytest.mark.parametrize("name,pname", [("jfk", "default"), ("bursts", "default"), ("bursts", "loose"), ("bursts", "split"), ("jfk", "nopad")])
def test_audio_filter_and_time_table_vs_live_reference(lib, model_dir, name, pname):
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    path = os.path.join(model_dir, "tiny.en-1.bin")
    if not os.path.exists(path):
        modelgen.write_model(path, "tiny.en", ftype=1)
    pcm = np.ascontiguousarray(signal(name))
    res = {}
    for tag, L, fn, kw in (("ours", lib, "whisper_b200_vad_filter", {}), ("ref", ref, "ref_vad_filter", {"use_gpu": False})):
        with api.Whisper(L, path, **kw) as w:
            p = w.greedy_params()
            p.vad = True
            p.vad_model_path = VAD_MODEL.encode()
            p.vad_params = vad_cases.vad_params(L, **vad_cases.PARAM_SETS[pname])
            res[tag] = _filter(L, fn, w.ctx, p, pcm)
    assert res["ours"][1] == res["ref"][1] and len(res["ours"][1]) > 0
    assert np.array_equal(res["ours"][0], res["ref"][0])


def test_full_with_vad_is_full_on_the_filtered_audio_with_mapped_times(lib, model_dir):
    path = os.path.join(model_dir, "tiny.en-1.bin")
    if not os.path.exists(path):
        modelgen.write_model(path, "tiny.en", ftype=1)
    pcm = np.ascontiguousarray(signal("bursts"))
    with api.Whisper(lib, path) as w:
        p = w.greedy_params()
        p.max_tokens = 24
        p.vad = True
        p.vad_model_path = VAD_MODEL.encode()
        rc, with_vad = w.full(p, pcm)
        assert rc == 0 and len(with_vad) > 0
        filtered, table = _filter(lib, "whisper_b200_vad_filter", w.ctx, p, pcm)
        assert 0 < len(filtered) < len(pcm)
        # parallel entry point, 2 processors: same filter in front
        rc, par = w.full(p, pcm, n_processors=2)
        assert rc == 0 and len(par) > 0
    # (a fresh context: like the reference, a state keeps its time table until the next VAD run, so a plain run on the same
    # state would read its times through the stale table)
    with api.Whisper(lib, path) as w:
        p = w.greedy_params()
        p.max_tokens = 24
        rc, plain = w.full(p, filtered)
        assert rc == 0
    assert [s.tokens for s in with_vad] == [s.tokens for s in plain]
    tab = (C.c_longlong * (2 * len(table)))()
    for i, (a, b) in enumerate(table):
        tab[2 * i], tab[2 * i + 1] = a, b
    for sv, sp in zip(with_vad, plain):
        t0 = lib.whisper_b200_vad_map_time(tab, len(table), sp.t0)
        t1 = lib.whisper_b200_vad_map_time(tab, len(table), sp.t1)
        assert sv.t0 == t0 and sv.t1 == (t0 + 10 if t1 - t0 < 10 else t1)
    assert with_vad[-1].t1 <= len(pcm) // 160 + 10


def test_full_with_vad_on_silence_returns_no_segments(lib, model_dir):
    path = os.path.join(model_dir, "tiny.en-1.bin")
    if not os.path.exists(path):
        modelgen.write_model(path, "tiny.en", ftype=1)
    with api.Whisper(lib, path) as w:
        p = w.greedy_params()
        p.vad = True
        p.vad_model_path = VAD_MODEL.encode()
        rc, segs = w.full(p, np.zeros(16000 * 3, np.float32))
        assert rc == 0 and segs == []
    with api.Whisper(lib, path) as w:               # (the detector is created once per state: a new state for the bad path)
        p = w.greedy_params()
        p.vad = True
        p.vad_model_path = b"/nonexistent/vad.bin"
        rc, _ = w.full(p, np.zeros(16000, np.float32))
        assert rc == -1
