"""Host logic of the sampling path (csrc/full.cu: whisper_process_logits + greedy whisper_sample_token restated) against the
compiled reference's own functions, on the CPU: same logits row, same decoder state, same parameters through
whisper_b200_process_logits (product, host-only hook) and ref_process_logits (oracle/ref_harness.cpp -> src/whisper.cpp:6177-6517).
"""
import ctypes as C
import os

import numpy as np
import pytest

import open_whisper_kit_b200 as pkg
from open_whisper_kit_b200 import api, capi, modelgen
from oracle import reflib

FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int32)


@pytest.fixture(scope="module", params=["tiny.en", "tiny"])
def ctxs(request, tmp_path_factory):
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    path = os.path.join(str(tmp_path_factory.mktemp("pl")), f"{request.param}-header.bin")
    modelgen.write_model(path, request.param, with_tensors=False)
    w = api.Whisper(ref, path, use_gpu=False, flash_attn=False)
    lib = pkg.load()
    lib.whisper_b200_process_logits.restype = C.c_int
    lib.whisper_b200_process_logits.argtypes = [C.c_void_p, C.c_int, IP, C.c_int, capi.whisper_full_params, C.c_float, FP, IP, C.c_int,
                                                C.c_int, C.c_int, FP, FP, FP, C.POINTER(capi.whisper_token_data), C.c_int, C.c_uint,
                                                C.POINTER(capi.whisper_token_data)]
    n_vocab = ref.whisper_n_vocab(w.ctx)
    texts = [ref.whisper_token_to_str(w.ctx, i) or b"" for i in range(n_vocab)]
    special = [ref.whisper_token_eot(w.ctx), ref.whisper_token_sot(w.ctx), ref.whisper_token_translate(w.ctx),
               ref.whisper_token_transcribe(w.ctx), ref.whisper_token_solm(w.ctx), ref.whisper_token_prev(w.ctx),
               ref.whisper_token_nosp(w.ctx), ref.whisper_token_not(w.ctx), ref.whisper_token_beg(w.ctx)]
    yield ref, w, lib, n_vocab, (C.c_char_p * n_vocab)(*texts), (C.c_int32 * 9)(*special), special


def _history(rng, kind, eot, beg):
    txt = lambda: int(rng.integers(0, eot))                      # noqa: E731
    ts = lambda: beg + int(rng.integers(1, 400))                 # noqa: E731
    return {"empty": [], "text": [txt(), txt(), txt()], "ts_last": [txt(), ts()], "ts_pair": [txt(), ts(), ts()],
            "only_ts": [ts()], "long": [ts()] + [txt() for _ in range(20)]}[kind]


@pytest.mark.parametrize("kind", ["empty", "text", "ts_last", "ts_pair", "only_ts", "long"])
@pytest.mark.parametrize("variant", ["default", "no_ts", "nst", "temp", "regex", "no_blank_rule", "ts_heavy"])
def test_process_logits_and_greedy_token_match_reference(ctxs, kind, variant):
    ref, w, lib, n_vocab, text_arr, special_arr, special = ctxs
    eot, beg = special[0], special[8]
    rng = np.random.default_rng(abs(hash((kind, variant, n_vocab))) % (2 ** 32))
    p = w.greedy_params(no_timestamps=False, n_threads=1)
    temperature = 0.0
    if variant == "no_ts":
        p.no_timestamps = True
    elif variant == "nst":
        p.suppress_nst = True
    elif variant == "temp":
        temperature = 0.6
    elif variant == "regex":
        p.suppress_regex = b".*[aeiou].*"
    elif variant == "no_blank_rule":
        p.suppress_blank = False
        p.max_initial_ts = 0.0
    logits = (3.0 * rng.standard_normal(n_vocab)).astype(np.float32)
    if variant == "ts_heavy":
        logits[beg:] += 6.0                                      # timestamp mass beats every text token
    hist = _history(rng, kind, eot, beg)
    has_ts = int(any(t >= beg for t in hist))
    seek_delta = 2 * (max([t - beg for t in hist if t >= beg] or [0]))
    h_arr = (C.c_int32 * max(1, len(hist)))(*hist)
    out = []
    for side in ("ours", "ref"):
        lo, lp, pr = (np.empty(n_vocab, np.float32) for _ in range(3))
        tok = capi.whisper_token_data()
        draws = (capi.whisper_token_data * 5)()
        if side == "ours":
            rc = lib.whisper_b200_process_logits(C.cast(text_arr, C.c_void_p), n_vocab, special_arr, 1500, p, temperature,
                                                 logits.ctypes.data_as(FP), h_arr, len(hist), has_ts, seek_delta, lo.ctypes.data_as(FP),
                                                 lp.ctypes.data_as(FP), pr.ctypes.data_as(FP), C.byref(tok), 5, 1234 + len(hist), draws)
        else:
            rc = ref.ref_process_logits(w.ctx, p, temperature, logits.ctypes.data_as(FP), h_arr, len(hist), has_ts, seek_delta,
                                        lo.ctypes.data_as(FP), lp.ctypes.data_as(FP), pr.ctypes.data_as(FP), C.byref(tok))
            assert ref.ref_sample_topk(w.ctx, 5, 1234 + len(hist), draws) == 0
        assert rc == 0
        out.append((lo, lp, pr, (tok.id, tok.tid, tok.p, tok.plog, tok.pt, tok.ptsum), [(d.id, d.tid) for d in draws]))
    (la, pa, qa, ta, da), (lb, pb, qb, tb, db) = out
    assert da == db, "mt19937 / discrete_distribution draws differ"
    assert np.array_equal(np.isneginf(la), np.isneginf(lb)), "different tokens suppressed"
    fin = ~np.isneginf(lb)
    assert np.array_equal(la[fin], lb[fin])
    assert np.array_equal(np.isneginf(pa), np.isneginf(pb))
    fin = ~np.isneginf(pb)
    assert np.abs(pa[fin] - pb[fin]).max() <= 2e-6 and np.abs(qa - qb).max() <= 1e-7
    assert ta[:2] == tb[:2] and np.allclose(ta[2:], tb[2:], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("case", range(8))
def test_sequence_score_matches_reference(ctxs, case):
    ref, w, lib, n_vocab, _, _, special = ctxs
    lib.whisper_b200_sequence_score.restype = C.c_int
    lib.whisper_b200_sequence_score.argtypes = [capi.whisper_full_params, FP, IP, C.c_int, C.c_int, C.POINTER(C.c_double)]
    ref.ref_sequence_score.restype = C.c_int
    ref.ref_sequence_score.argtypes = [capi.whisper_full_params, FP, IP, C.c_int, C.c_int, C.POINTER(C.c_double)]
    rng = np.random.default_rng(50 + case)
    n = int(rng.integers(1, 80))
    ids = rng.integers(0, 40 if case % 2 else special[0], n).astype(np.int32)          # few distinct ids: low entropy
    plog = (-rng.exponential(1.0, n)).astype(np.float32)
    p = w.greedy_params(no_timestamps=False, n_threads=1)
    p.length_penalty = [-1.0, 0.0, 0.6, 1.0][case % 4]
    result_len = int(rng.integers(0 if case == 0 else 1, n + 1))
    a, b = (C.c_double * 4)(), (C.c_double * 4)()
    assert lib.whisper_b200_sequence_score(p, plog.ctypes.data_as(FP), ids.ctypes.data_as(IP), n, result_len, a) == 0
    assert ref.ref_sequence_score(p, plog.ctypes.data_as(FP), ids.ctypes.data_as(IP), n, result_len, b) == 0
    assert np.allclose(list(a), list(b), rtol=1e-12, atol=0.0), (list(a), list(b))
