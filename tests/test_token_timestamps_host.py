"""Host logic of token-level timestamps / max_len wrapping (csrc/full.cu) against the compiled reference, on the CPU.

Both sides get the same synthetic segment -- token ids, tid / pt / ptsum, segment times, PCM for the energy envelope, carried
{t_beg, t_last, tid_last} -- through host-only hooks: whisper_b200_token_timestamps of the product (no device needed) and
ref_token_timestamps of oracle/ref_harness.cpp, which calls the reference's own static functions
(whisper_exp_compute_token_level_timestamps, whisper_wrap_segment, get_signal_energy; src/whisper.cpp:8425-8660, 6077-6130).
"""
import ctypes as C
import os

import numpy as np
import pytest

import open_whisper_kit_b200 as pkg
from open_whisper_kit_b200 import api, capi, modelgen
from oracle import reflib

FP = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def ref_ctx(tmp_path_factory):
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    path = os.path.join(str(tmp_path_factory.mktemp("tt")), "tiny.en-header.bin")
    modelgen.write_model(path, "tiny.en", with_tensors=False)          # vocabulary only: "test model" of the reference loader
    w = api.Whisper(ref, path, use_gpu=False, flash_attn=False)
    yield ref, w
    w.close() if hasattr(w, "close") else None


def _segment(rng, texts, beg, eot, n_tok, t0, t1, with_lead_ts):
    toks = []
    if with_lead_ts:
        toks.append((beg + int(rng.integers(0, 5)), 1.0, 1.0))
    for _ in range(n_tok):
        tid = beg + int(rng.integers(0, (t1 - t0) // 2 + 40))
        toks.append((int(rng.integers(0, eot)), float(rng.uniform(0, 0.3)), float(rng.uniform(0, 0.3))) + (tid,))
    toks.append((beg + (t1 - t0) // 2, 0.9, 0.9))
    arr = (capi.whisper_token_data * len(toks))()
    for i, t in enumerate(toks):
        arr[i].id = t[0]
        arr[i].tid = t[3] if len(t) > 3 else t[0]
        arr[i].p, arr[i].plog, arr[i].pt, arr[i].ptsum = 0.5, -0.7, t[1], t[2]
        arr[i].t0 = arr[i].t1 = arr[i].t_dtw = -1
        arr[i].vlen = 0.0
    return arr


@pytest.mark.parametrize("case", range(12))
def test_token_timestamps_and_wrapping_match_reference(ref_ctx, case):
    ref, w = ref_ctx
    lib = pkg.load()
    n_vocab = ref.whisper_n_vocab(w.ctx)
    eot, beg = ref.whisper_token_eot(w.ctx), ref.whisper_token_beg(w.ctx)
    texts = [ref.whisper_token_to_str(w.ctx, i) or b"" for i in range(n_vocab)]
    text_arr = (C.c_char_p * n_vocab)(*texts)
    rng = np.random.default_rng(100 + case)
    n_samples = 16000 * int(rng.integers(3, 12))
    pcm = modelgen.synth_pcm(n_samples, seed=case, stream=case % 3)
    if case % 4 == 0:
        pcm[: n_samples // 3] = 0.0                                     # a silent stretch: the energy threshold logic
    t0 = int(rng.integers(0, 50))
    t1 = t0 + int(rng.integers(80, n_samples // 160))
    state0 = [int(rng.integers(0, 40)), int(rng.integers(0, 40)), beg + int(rng.integers(0, 10))]
    max_len, sow = [(0, 0), (10, 0), (14, 1), (1, 0)][case % 4]
    out = []
    for side in ("ours", "ref"):
        toks = _segment(np.random.default_rng(7 * case + 1), texts, beg, eot, int(4 + case * 2), t0, t1, case % 3 != 2)
        st = (C.c_longlong * 3)(*state0)
        seg_t = (C.c_longlong * 128)()
        seg_n = (C.c_int * 64)()
        if side == "ours":
            n = lib.whisper_b200_token_timestamps(C.cast(text_arr, C.c_void_p), n_vocab, eot, beg, pcm.ctypes.data_as(FP), n_samples, t0, t1,
                                                  C.cast(toks, C.c_void_p), len(toks), 0.01, 0.01, st, max_len, sow, seg_t, seg_n, 64)
        else:
            n = ref.ref_token_timestamps(w.ctx, pcm.ctypes.data_as(FP), n_samples, t0, t1, C.cast(toks, C.c_void_p), len(toks), 0.01, 0.01,
                                         st, max_len, sow, seg_t, seg_n, 64)
        assert n >= 1
        out.append((n, list(st), [(seg_t[2 * k], seg_t[2 * k + 1], seg_n[k]) for k in range(min(n, 64))],
                    [(t.id, t.t0, t.t1, t.vlen) for t in toks]))
    assert out[0] == out[1]
    assert any(t[1] >= 0 for t in out[0][3])
