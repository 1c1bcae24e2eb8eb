"""GPU parity of the transcription path through the whisper.h C ABI against the reference.

Two anchors:
  * golden fixtures in tests/golden/ produced by the UNMODIFIED reference CPU build (tests/golden/make_golden.py);
  * when oracle/_ref/libwhisper_ref_*.so travelled to the box, the same reference run live on the host CPU.

Tolerances (BASELINE.json north_star): mel 1e-5 (see test_gpu_kernels.py for the fp32-FFT noise floor), encoder output
and logits <= 2e-2 max-abs, greedy token sequences identical.  The golden mode is flash_attn=false with an F16 model
file; with flash_attn=true the reference's CPU flash kernel accumulates P*V in F16 for short queries
(ggml/src/ggml-cpu/ops.cpp:8140-8208), so only the encoder (fp32 tiled kernel + 36 phantom keys) and the token
sequences are compared in that mode, and logits get the looser documented bound.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from open_whisper_kit_b200 import api, capi, modelgen
from oracle import reflib

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
FP = C.POINTER(C.c_float)
GOLD_TOK = json.load(open(os.path.join(HERE, "golden", "golden_tokens.json")))
GOLD_TEN = np.load(os.path.join(HERE, "golden", "golden_tensors.npz"))


def model_path(model_dir, arch, ftype=1):
    p = os.path.join(model_dir, f"{arch}-{ftype}.bin")
    if not os.path.exists(p):
        modelgen.write_model(p, arch, ftype=ftype)
    return p


def pcm_for(spec):
    if spec["kind"] == "jfk":
        return api.read_wav_f32(os.path.join(HERE, "golden", "jfk.wav"))
    return np.concatenate([modelgen.synth_pcm(480000, seed=spec["seed"], stream=i) for i in range(spec["windows"])])


def get_mel(lib, w):
    n_len, n_mel = C.c_int(), C.c_int()
    assert lib.whisper_b200_get_mel(w.ctx, None, None, 0, C.byref(n_len), C.byref(n_mel)) == 0
    mel = np.empty((n_mel.value, n_len.value), np.float32)
    assert lib.whisper_b200_get_mel(w.ctx, None, mel.ctypes.data_as(FP), mel.size, C.byref(n_len), C.byref(n_mel)) == 0
    return mel


def get_enc(lib, w, d):
    enc = np.empty((1500, d), np.float32)
    assert lib.whisper_b200_get_encoder_output(w.ctx, enc.ctypes.data_as(FP), enc.size) == 0
    return enc


@pytest.mark.parametrize("key", ["tiny.en/f1/fa0", "tiny.en/f1/fa1", "tiny/f1/fa0"])
def test_mel_encoder_logits_vs_golden(lib, model_dir, key):
    arch, _, fa = key.split("/")
    fa = fa == "fa1"
    d = modelgen.ARCHS[arch][2]
    pcm = pcm_for({"kind": "jfk"})
    with api.Whisper(lib, model_path(model_dir, arch), flash_attn=fa) as w:
        assert w.pcm_to_mel(pcm) == 0
        mel = get_mel(lib, w)
        g = GOLD_TEN[key + "/mel_sub"]
        dm = np.abs(mel[:, :1100:5] - g)
        print(f"{key}: mel max|d|={dm.max():.3e}")
        assert dm.max() <= 5e-5 and (dm <= 1e-5 * np.maximum(np.abs(g), 1.0)).mean() >= 0.999
        assert abs(mel.astype(np.float64).sum() - GOLD_TEN[key + "/mel_sum"][0]) <= 1e-6 * GOLD_TEN[key + "/mel_sum"][1] + 1e-2

        assert w.encode(0) == 0
        enc = get_enc(lib, w, d)
        ge = GOLD_TEN[key + "/enc_sub"]
        de = np.abs(enc[::25, ::3] - ge)
        print(f"{key}: embd_enc max|d|={de.max():.3e} mean|d|={de.mean():.3e}")
        assert de.max() <= 2e-2

        sot = lib.whisper_token_sot(w.ctx)
        rc, lg = w.decode([sot], 0)
        assert rc == 0
        gl = GOLD_TEN[key + "/logits_sub"]
        dl = np.abs(lg[::17] - gl)
        top = GOLD_TEN[key + "/logits_top_ids"]
        dtop = np.abs(lg[top] - GOLD_TEN[key + "/logits_top_vals"])
        print(f"{key}: logits max|d|={dl.max():.3e} top16 max|d|={dtop.max():.3e}")
        tol = 2e-2 if not fa else 0.25          # fa=1: reference accumulates P*V in F16 for 1-token queries
        assert dl.max() <= tol and dtop.max() <= tol
        assert int(lg.argmax()) == int(top[0])


def run_case(lib, model_dir, case):
    with api.Whisper(lib, model_path(model_dir, case["arch"], case["ftype"]), flash_attn=case["flash_attn"]) as w:
        p = w.greedy_params(no_timestamps=case["no_timestamps"])
        rc, segs = w.full(p, pcm_for(case["pcm"]), n_processors=case["n_processors"])
        return rc, [[int(s.t0), int(s.t1), [int(x) for x in s.tokens]] for s in segs], segs


MARGIN = 5e-3        # flips are accepted only below this top-2 logit margin -- an eighth of 2 x the spec's logits tolerance (2e-2);
                     # every flip measured so far sits below 1e-3 (profiles/r2_parity_report.json)


def by_chunk(segments):
    """segments: [(t0, t1, tokens, token_data|None)] -> {chunk index: [(token, data)]}; chunks are 30 s = 3000 units."""
    out = {}
    for t0, t1, toks, tds in segments:
        c = max(0, (int(t1) - 1) // 3000)
        out.setdefault(c, [])
        out[c] += [(tok, tds[i] if tds else None) for i, tok in enumerate(toks)]
    return out


def _report(name, entry):
    try:
        out = os.path.join(os.path.dirname(HERE), "gpurun_out")
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "parity_report.json")
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[name] = entry
        json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.mark.parametrize("name", sorted(GOLD_TOK.keys()))
def test_greedy_tokens_identical_to_reference_golden(lib, model_dir, name, monkeypatch):
    """Greedy tokens and segment times vs the reference CPU path (BASELINE.json configs 1 and 2).

    Bar: identical.  north_star also allows the logits to differ by 2e-2, and on random-init weights (Gaussian logits) the
    reference's own top-1 / top-2 margin is tiny at a fixed fraction of the steps, whatever the scale of the init -- margins and
    rounding errors scale together (1.3 % of the steps are below MARGIN = 5e-3); the reference's AVX2 and AVX-512 builds diverge
    from each other on the base.en case.  So identity is REQUIRED wherever the reference's recorded margin
    allows it and a difference is accepted only where it does not:
      * case with recorded per-step margins (golden "steps", from the reference alone): every token before a chunk's first
        sub-margin step must be identical; the first mismatch of a chunk, if any, must sit ON a sub-margin step and be the
        reference's recorded runner-up;
      * other cases (no margins recorded -- timestamp mode): a chunk's first mismatch must be a near-tie of OUR two best
        candidates (< MARGIN logits) with the reference's token as our runner-up; at most max(1, chunks / 4) chunks may have one.
    Every flip is written to gpurun_out/parity_report.json."""
    monkeypatch.setenv("WHISPER_B200_DEBUG_GAPS", "1")
    case = GOLD_TOK[name]
    rc, _, segs = run_case(lib, model_dir, case)
    assert rc == case["rc"] == 0
    ours = by_chunk([(s.t0, s.t1, s.tokens, s.token_data) for s in segs])
    ref = by_chunk([(s[0], s[1], s[2], None) for s in case["segments"]])
    n_chunks = case["n_processors"]
    steps = case.get("steps")
    n_ident, flips = 0, []
    for c in range(n_chunks):
        a = [t for t, _ in ours.get(c, [])]
        b = [t for t, _ in ref.get(c, [])]
        if a == b:
            n_ident += 1
            continue
        k = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), None)
        assert k is not None, f"chunk {c}: one sequence is a strict prefix of the other ({len(a)} vs {len(b)} tokens)"
        td = ours[c][k][1]
        if steps:
            gaps, runner = steps[c]["gaps"], steps[c]["runner_up"]
            strict = next((i for i, g in enumerate(gaps) if g < MARGIN), len(gaps))
            flips.append({"chunk": c, "step": k, "ours": a[k], "reference": b[k], "reference_margin": gaps[k], "strict_prefix": strict})
            assert k >= strict and gaps[k] < MARGIN and a[k] == runner[k], \
                f"chunk {c} step {k}: ours {a[k]} vs reference {b[k]}; reference margin {gaps[k]}, its runner-up {runner[k]}"
        else:
            flips.append({"chunk": c, "step": k, "ours": a[k], "reference": b[k], "our_margin": round(td.vlen, 5)})
            assert td.vlen < MARGIN and td.t_dtw == b[k], \
                f"chunk {c} step {k}: ours {a[k]} vs reference {b[k]}, our runner-up {td.t_dtw} at distance {td.vlen}"
    n_ref_tok = sum(len(v) for v in ref.values())
    _report("golden_tokens/" + name, {"reference_tokens": n_ref_tok, "chunks": n_chunks, "chunks_identical": n_ident, "flips": flips})
    print(f"{name}: {n_ref_tok} reference tokens, {n_ident}/{n_chunks} chunks identical, flips: {flips}")
    if not flips:
        assert [(s.t0, s.t1) for s in segs] == [(s[0], s[1]) for s in case["segments"]]
    if not steps:
        assert len(flips) <= max(1, n_chunks // 4)


def test_token_data_fields_match_live_reference(lib, model_dir):
    """p / plog / pt / ptsum of whisper_token_data against the reference run live on this host's CPU."""
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    case = GOLD_TOK["tiny.en/synth4/ts/fa0"]
    rc, _, segs = run_case(lib, model_dir, case)
    with api.Whisper(ref, model_path(model_dir, case["arch"]), use_gpu=False, flash_attn=False) as r:
        p = r.greedy_params(no_timestamps=False, n_threads=4)
        rc2, rsegs = r.full(p, pcm_for(case["pcm"]), n_processors=case["n_processors"])
    assert rc == 0 and rc2 == 0
    worst, n_cmp = 0.0, 0
    for a, b in zip(segs, rsegs):
        if a.tokens != b.tokens:          # after a near-tie flip the sequences are different sequences (see above)
            break
        assert a.text == b.text and (a.t0, a.t1) == (b.t0, b.t1)
        assert abs(a.no_speech_prob - b.no_speech_prob) <= 1e-3
        for ta, tb in zip(a.token_data, b.token_data):
            assert ta.id == tb.id and ta.tid == tb.tid
            worst = max(worst, abs(ta.plog - tb.plog), abs(ta.p - tb.p), abs(ta.pt - tb.pt), abs(ta.ptsum - tb.ptsum))
            n_cmp += 1
    print(f"worst token_data float deviation over {n_cmp} tokens:", worst)
    assert n_cmp >= 40 and worst <= 2e-2


@pytest.mark.parametrize("max_len,sow", [(0, False), (12, False), (16, True)])
def test_token_level_timestamps_and_max_len_vs_live_reference(lib, model_dir, max_len, sow):
    """params.token_timestamps / max_len / split_on_word (the cli's -ml / -sow / word-level output): per-token t0 / t1 / vlen
    and the re-wrapped segments against the reference run live on this host's CPU (src/whisper.cpp:8455-8660, 6077-6130)."""
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    path = model_path(model_dir, "tiny.en")
    pcm = pcm_for({"kind": "synth", "seed": 7, "windows": 2})

    def run(lib_, **kw):
        with api.Whisper(lib_, path, flash_attn=False, **kw) as w:
            p = w.greedy_params(no_timestamps=False, n_threads=8)
            p.token_timestamps = True
            p.max_len = max_len
            p.split_on_word = sow
            rc, segs = w.full(p, pcm)
            assert rc == 0
            return segs

    segs, rsegs = run(lib), run(ref, use_gpu=False)
    n_cmp = 0
    for a, b in zip(segs, rsegs):
        if a.tokens != b.tokens:          # after a near-tie flip the sequences are different sequences
            break
        assert a.text == b.text and (a.t0, a.t1) == (b.t0, b.t1)
        for ta, tb in zip(a.token_data, b.token_data):
            assert (ta.t0, ta.t1) == (tb.t0, tb.t1) and ta.vlen == tb.vlen
            n_cmp += 1
    print(f"max_len={max_len} split_on_word={sow}: {len(segs)} segments (reference {len(rsegs)}), {n_cmp} tokens compared")
    assert n_cmp >= 20 and len(rsegs) >= (3 if max_len else 1)
    if max_len:
        assert all(len(s.text) <= max_len or len(s.tokens) <= 2 or sow for s in segs)


@pytest.mark.parametrize("audio_ctx,fa", [(512, False), (750, False), (768, True)])
def test_audio_ctx_vs_live_reference(lib, model_dir, audio_ctx, fa):
    """whisper_full_params::audio_ctx (the cli's -ac): the encoder runs over the first 2 * audio_ctx mel frames of every window,
    the cross K/V and the decoder's cross-attention over audio_ctx positions (src/whisper.cpp:1982, 2044, 2278, 2383, 2479);
    with flash attention the phantom keys are pad256(audio_ctx) - audio_ctx.  Tokens, segment times and token probabilities
    against the reference run live on this host's CPU."""
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    path = model_path(model_dir, "tiny.en")
    pcm = pcm_for({"kind": "synth", "seed": 7, "windows": 2})

    def run(lib_, **kw):
        with api.Whisper(lib_, path, flash_attn=fa, **kw) as w:
            p = w.greedy_params(no_timestamps=False, n_threads=8)
            p.audio_ctx = audio_ctx
            rc, segs = w.full(p, pcm)
            assert rc == 0
            return segs

    segs, rsegs = run(lib), run(ref, use_gpu=False)
    worst, n_cmp = 0.0, 0
    for a, b in zip(segs, rsegs):
        if a.tokens != b.tokens:          # after a near-tie flip the sequences are different sequences
            break
        assert a.text == b.text and (a.t0, a.t1) == (b.t0, b.t1)
        for ta, tb in zip(a.token_data, b.token_data):
            worst = max(worst, abs(ta.plog - tb.plog), abs(ta.p - tb.p))
            n_cmp += 1
    print(f"audio_ctx={audio_ctx} flash={fa}: {n_cmp} tokens compared, worst p / plog deviation {worst:.3e}")
    assert n_cmp >= 20 and worst <= (2e-2 if not fa else 5e-2)     # measured: 1.2e-3 / 1.1e-3 / 8.6e-3


def test_language_auto_detect_vs_live_reference(lib, model_dir):
    """whisper_lang_auto_detect (encode the window at the offset, one decoder step on <|sot|>, softmax over the language
    tokens; src/whisper.cpp:4021-4094) and whisper_full with language = "auto" against the reference run live on this CPU."""
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    path = model_path(model_dir, "tiny")            # multilingual vocabulary (99 languages)
    pcm = pcm_for({"kind": "synth", "seed": 7, "windows": 2})
    n_lang = lib.whisper_lang_max_id() + 1
    out = {}
    for name, lib_, kw in (("ours", lib, {}), ("ref", ref, {"use_gpu": False})):
        with api.Whisper(lib_, path, flash_attn=False, **kw) as w:
            assert w.pcm_to_mel(pcm, 8) == 0
            per_offset = []
            for offset_ms in (0, 12000):
                probs = np.zeros(n_lang, np.float32)
                lid = lib_.whisper_lang_auto_detect(w.ctx, offset_ms, 8, probs.ctypes.data_as(FP))
                per_offset.append((lid, probs))
            assert lib_.whisper_lang_auto_detect(w.ctx, 10 ** 7, 8, None) == -2          # past the end of the audio
            p = w.greedy_params(no_timestamps=False, n_threads=8, language=b"auto")
            rc, segs = w.full(p, pcm)
            assert rc == 0
            out[name] = (per_offset, lib_.whisper_full_lang_id(w.ctx), [t for s_ in segs for t in s_.tokens])
    for (la, pa), (lb, pb) in zip(out["ours"][0], out["ref"][0]):
        top2 = np.sort(pb)[-2:]
        print(f"language id ours {la} reference {lb}; probs max|d| = {np.abs(pa - pb).max():.3e}; reference top-2 {top2[1]:.4f} / {top2[0]:.4f}")
        assert abs(pa.sum() - 1.0) < 1e-4 and np.abs(pa - pb).max() <= 2e-3
        if top2[1] - top2[0] > 4e-3:
            assert la == lb
    if out["ours"][1] == out["ref"][1]:
        n = min(len(out["ours"][2]), len(out["ref"][2]), 12)
        assert n >= 4 and out["ours"][2][:n] == out["ref"][2][:n]


def test_batched_language_detection_matches_per_stream_and_reference(lib, model_dir):
    """whisper_full_parallel with language = "auto": the chunks' languages are detected in one encoder batch + one decoder
    step (csrc/full.cu) -- same ids and same transcription as the reference, whose worker states detect one by one."""
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    path = model_path(model_dir, "tiny")
    pcm = pcm_for({"kind": "synth", "seed": 7, "windows": 3})
    res = {}
    for name, lib_, kw in (("ours", lib, {}), ("ref", ref, {"use_gpu": False})):
        with api.Whisper(lib_, path, flash_attn=False, **kw) as w:
            p = w.greedy_params(no_timestamps=False, n_threads=4, language=b"auto")
            rc, segs = w.full(p, pcm, n_processors=3)
            assert rc == 0
            res[name] = (lib_.whisper_full_lang_id(w.ctx), [(s_.t0, s_.t1, s_.tokens) for s_ in segs])
    print("language ids:", res["ours"][0], res["ref"][0], "segments:", len(res["ours"][1]), len(res["ref"][1]))
    assert res["ours"][0] == res["ref"][0]
    n = min(len(res["ours"][1]), len(res["ref"][1]), 4)
    assert n >= 2 and res["ours"][1][:n] == res["ref"][1][:n]


@pytest.mark.parametrize("n_samples", [800, 16000 * 3 + 123, 480000 + 16000 * 7])
def test_short_and_ragged_inputs_vs_live_reference(lib, model_dir, n_samples):
    """Edge lengths through whisper_full: less audio than one FFT hop budget (no window at all), a few seconds, and one full
    window plus a ragged tail (second window mostly padding) -- return code, segment times and tokens against the live reference;
    and the same ragged signal split over 3 chunks by whisper_full_parallel."""
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    path = model_path(model_dir, "tiny.en")
    pcm = modelgen.synth_pcm(480000 + 16000 * 7, seed=3, stream=1)[:n_samples]
    out = {}
    for name, lib_, kw in (("ours", lib, {}), ("ref", ref, {"use_gpu": False})):
        with api.Whisper(lib_, path, flash_attn=False, **kw) as w:
            p = w.greedy_params(no_timestamps=False, n_threads=8)
            rc, segs = w.full(p, pcm)
            res = [(rc, [(s_.t0, s_.t1, s_.tokens) for s_ in segs])]
            if n_samples > 480000:
                rc, segs = w.full(p, pcm, n_processors=3)
                res.append((rc, [(s_.t0, s_.t1, s_.tokens) for s_ in segs]))
            out[name] = res
    for (rca, sa), (rcb, sb) in zip(out["ours"], out["ref"]):
        assert rca == rcb == 0
        print(f"{n_samples} samples: {len(sa)} segments (reference {len(sb)})")
        if n_samples <= 16000 * 4:
            assert sa == sb
        else:                               # long enough for a near-tie flip on random weights: compare the common head
            n = min(len(sa), len(sb), 3)
            assert len(sb) == 0 or (n >= 1 and sa[:n] == sb[:n])


def test_callbacks_vs_live_reference(lib, model_dir):
    """new_segment / progress / encoder_begin / abort / logits_filter callbacks of whisper_full: same call sequence, same
    arguments and same effect as in the reference (src/whisper.cpp:7035-7052, 7504-7538, 7685-7694)."""
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    path = model_path(model_dir, "tiny.en")
    pcm = pcm_for({"kind": "synth", "seed": 7, "windows": 2})
    NEW_SEG = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)
    PROGRESS = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)
    ENC_BEGIN = C.CFUNCTYPE(C.c_bool, C.c_void_p, C.c_void_p, C.c_void_p)
    ABORT = C.CFUNCTYPE(C.c_bool, C.c_void_p)
    FILTER = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_void_p)

    def run(lib_, mode, **kw):
        ev = []
        with api.Whisper(lib_, path, flash_attn=False, **kw) as w:
            n_vocab = lib_.whisper_n_vocab(w.ctx)
            eot = lib_.whisper_token_eot(w.ctx)
            p = w.greedy_params(no_timestamps=False, n_threads=8)
            cbs = [NEW_SEG(lambda c, st, n_new, ud: ev.append(("seg", n_new, lib_.whisper_full_n_segments(w.ctx)))),
                   PROGRESS(lambda c, st, prog, ud: ev.append(("progress", prog))),
                   ENC_BEGIN(lambda c, st, ud: (ev.append(("enc",)), mode != "enc_false" or len([e for e in ev if e[0] == "enc"]) < 2)[1]),
                   ABORT(lambda ud: mode == "abort")]            # always true: the first encoder pass is abandoned (-6)

            def filt(c, st, toks, n_toks, logits, ud):
                a = np.ctypeslib.as_array(logits, shape=(n_vocab,))
                a[1:eot:2] = -np.inf                                   # odd text tokens are forbidden
            cbs.append(FILTER(filt))
            p.new_segment_callback = C.cast(cbs[0], C.c_void_p)
            p.progress_callback = C.cast(cbs[1], C.c_void_p)
            p.encoder_begin_callback = C.cast(cbs[2], C.c_void_p)
            if mode == "abort":
                p.abort_callback = C.cast(cbs[3], C.c_void_p)
            if mode == "filter":
                p.logits_filter_callback = C.cast(cbs[4], C.c_void_p)
            rc, segs = w.full(p, pcm)
            toks = [t for s_ in segs for t in s_.tokens]
        return rc, [e for e in ev if e[0] != "abort?"], toks, eot

    for mode in ("plain", "filter", "enc_false", "abort"):
        rc_a, ev_a, tok_a, eot = run(lib, mode)
        rc_b, ev_b, tok_b, _ = run(ref, mode, use_gpu=False)
        print(f"{mode}: rc {rc_a}/{rc_b}, {len(ev_a)}/{len(ev_b)} callback events, {len(tok_a)}/{len(tok_b)} tokens")
        assert rc_a == rc_b
        if mode in ("plain", "filter", "enc_false"):
            n = min(len(tok_a), len(tok_b), 16)
            assert tok_a[:n] == tok_b[:n] and (mode == "enc_false" or n >= 8)
            if tok_a == tok_b:
                assert ev_a == ev_b
        if mode == "filter":
            assert all(t % 2 == 0 or t >= eot for t in tok_a)


@pytest.mark.parametrize("arch", ["tiny.en", "tiny", "large-v3-turbo"])
def test_vocabulary_and_model_info_vs_live_reference(lib, model_dir, arch, tmp_path):
    """whisper_model_load's header / vocabulary handling (src/whisper.cpp:1485-1672): every token string incl. the synthesised
    specials ([_EOT_], [_TT_n], [_LANG_xx] ...), every special-token id and every model getter equal the reference's, for the
    51864-, 51865- and 51866-entry vocabularies (header-only files: the reference's own "test model" convention)."""
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    path = os.path.join(str(tmp_path), f"{arch}-header.bin")
    modelgen.write_model(path, arch, with_tensors=False)
    getters = ["whisper_n_vocab", "whisper_n_text_ctx", "whisper_n_audio_ctx", "whisper_is_multilingual", "whisper_model_n_vocab",
               "whisper_model_n_audio_ctx", "whisper_model_n_audio_state", "whisper_model_n_audio_head", "whisper_model_n_audio_layer",
               "whisper_model_n_text_ctx", "whisper_model_n_text_state", "whisper_model_n_text_head", "whisper_model_n_text_layer",
               "whisper_model_n_mels", "whisper_model_ftype", "whisper_model_type", "whisper_token_eot", "whisper_token_sot",
               "whisper_token_solm", "whisper_token_prev", "whisper_token_nosp", "whisper_token_not", "whisper_token_beg",
               "whisper_token_translate", "whisper_token_transcribe"]
    info = {}
    for name, lib_, kw in (("ours", lib, {}), ("ref", ref, {"use_gpu": False})):
        with api.Whisper(lib_, path, flash_attn=False, **kw) as w:
            vals = {g: getattr(lib_, g)(w.ctx) for g in getters}
            n = vals["whisper_n_vocab"]
            vals["strings"] = [lib_.whisper_token_to_str(w.ctx, i) for i in range(n)]
            vals["lang_tokens"] = [lib_.whisper_token_lang(w.ctx, i) for i in range(lib_.whisper_lang_max_id() + 1)]
            vals["model_type_str"] = lib_.whisper_model_type_readable(w.ctx)
            info[name] = vals
    for k in info["ref"]:
        assert info["ours"][k] == info["ref"][k], k


def test_low_level_api_vs_live_reference(lib, model_dir):
    """whisper_pcm_to_mel -> whisper_encode -> whisper_decode (bench.cpp's call sequence) against the live reference."""
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    path = model_path(model_dir, "base.en")
    pcm = modelgen.synth_pcm(480000, stream=9)
    with api.Whisper(lib, path, flash_attn=False) as w, api.Whisper(ref, path, use_gpu=False, flash_attn=False) as r:
        assert w.pcm_to_mel(pcm) == 0 and r.pcm_to_mel(pcm, 4) == 0
        assert lib.whisper_n_len(w.ctx) == ref.whisper_n_len(r.ctx)
        assert w.encode(0) == 0 and r.encode(0, 8) == 0
        enc = get_enc(lib, w, 512)
        renc = np.empty((1500, 512), np.float32)
        assert ref.ref_embd_enc_copy(r.ctx, renc.ctypes.data_as(FP), renc.size) == 0
        print("base.en embd_enc max|d| =", np.abs(enc - renc).max())
        assert np.abs(enc - renc).max() <= 2e-2
        toks = [lib.whisper_token_sot(w.ctx), lib.whisper_token_not(w.ctx)]
        rc1, lg = w.decode(toks, 0)
        rc2, rlg = r.decode(toks, 0, 8)
        assert rc1 == 0 and rc2 == 0
        worst = np.abs(lg - rlg).max()
        for step in range(6):               # teacher-forced single-token steps over the KV cache
            nxt = int(rlg.argmax())
            rc1, lg = w.decode([nxt], len(toks))
            rc2, rlg = r.decode([nxt], len(toks), 8)
            toks.append(nxt)
            assert rc1 == 0 and rc2 == 0
            worst = max(worst, np.abs(lg - rlg).max())
            assert int(lg.argmax()) == int(rlg.argmax())
        print("base.en logits max|d| over prompt + 6 steps =", worst)
        assert worst <= 2e-2


def test_large_v3_turbo_geometry_vs_live_reference(lib, model_dir):
    """The geometry of BASELINE.json configs 3/4 (128 mel bins, d = 1280, 20 heads, 32 encoder layers; turbo = 4 text layers so
    the CPU reference stays affordable): mel -> encoder -> prompt + teacher-forced single-token steps against the live reference."""
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    path = model_path(model_dir, "large-v3-turbo")
    pcm = pcm_for({"kind": "jfk"})
    n_thr = min(os.cpu_count() or 4, 32)
    with api.Whisper(lib, path, flash_attn=False) as w, api.Whisper(ref, path, use_gpu=False, flash_attn=False) as r:
        assert w.pcm_to_mel(pcm) == 0 and r.pcm_to_mel(pcm, n_thr) == 0
        mel = get_mel(lib, w)
        n_len, n_len_org, n_mel = C.c_int(), C.c_int(), C.c_int()
        assert ref.ref_mel_dims(r.ctx, C.byref(n_len), C.byref(n_len_org), C.byref(n_mel)) == 0
        assert (n_mel.value, n_len.value) == mel.shape == (128, mel.shape[1])
        rmel = np.empty(mel.shape, np.float32)
        assert ref.ref_mel_copy(r.ctx, rmel.ctypes.data_as(FP)) == 0
        dm = np.abs(mel - rmel)
        print(f"128-bin mel max|d| = {dm.max():.3e}")
        assert dm.max() <= 5e-5 and (dm <= 1e-5 * np.maximum(np.abs(rmel), 1.0)).mean() >= 0.999
        assert w.encode(0) == 0 and r.encode(0, n_thr) == 0
        enc = get_enc(lib, w, 1280)
        renc = np.empty((1500, 1280), np.float32)
        assert ref.ref_embd_enc_copy(r.ctx, renc.ctypes.data_as(FP), renc.size) == 0
        de = np.abs(enc - renc)
        print(f"large-v3-turbo embd_enc max|d| = {de.max():.3e} mean|d| = {de.mean():.3e} (rms {np.sqrt((renc ** 2).mean()):.3f})")
        assert de.max() <= 2e-2
        toks = [lib.whisper_token_sot(w.ctx), lib.whisper_token_lang(w.ctx, 0), lib.whisper_token_transcribe(w.ctx)]
        rc1, lg = w.decode(toks, 0)
        rc2, rlg = r.decode(toks, 0, n_thr)
        assert rc1 == 0 and rc2 == 0
        worst = np.abs(lg - rlg).max()
        for step in range(4):
            nxt = int(rlg.argmax())
            rc1, lg = w.decode([nxt], len(toks))
            rc2, rlg = r.decode([nxt], len(toks), n_thr)
            toks.append(nxt)
            assert rc1 == 0 and rc2 == 0
            worst = max(worst, np.abs(lg - rlg).max())
        print(f"large-v3-turbo logits max|d| over prompt + 4 steps = {worst:.3e}")
        assert worst <= 2e-2


def test_cross_kv_pool_matches_live_reference(lib, model_dir):
    """Cross-attention K/V (src/whisper.cpp:2272-2346) read back from the head-major device pool ([head][K|V][1500][64] per
    window and layer, written by the GEMM epilogue) and handed out in the reference's [1500][K | V] order, against the
    reference's own kv_cross (K = [layer][1500][d], V = [layer][d][1500] without flash attention)."""
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    path = model_path(model_dir, "tiny.en")
    pcm = pcm_for({"kind": "jfk"})
    n_thr = min(os.cpu_count() or 4, 32)
    d, n_layer = 384, 4
    with api.Whisper(lib, path, flash_attn=False) as w, api.Whisper(ref, path, use_gpu=False, flash_attn=False) as r:
        assert w.pcm_to_mel(pcm) == 0 and r.pcm_to_mel(pcm, n_thr) == 0
        assert w.encode(0) == 0 and r.encode(0, n_thr) == 0
        n = ref.ref_kv_cross_copy(r.ctx, 0, None, 0)
        used = n_layer * 1500 * d          # the buffer is allocated for a padded context; layers are packed at 1500 rows
        assert n >= used
        rk, rv = np.empty(n, np.float32), np.empty(n, np.float32)
        assert ref.ref_kv_cross_copy(r.ctx, 0, rk.ctypes.data_as(FP), n) == n
        assert ref.ref_kv_cross_copy(r.ctx, 1, rv.ctypes.data_as(FP), n) == n
        rk = rk[:used].reshape(n_layer, 1500, d)
        rv = rv[:used].reshape(n_layer, d, 1500).transpose(0, 2, 1)
        for il in range(n_layer):
            buf = np.zeros(1500 * 2 * d, np.uint16)
            assert lib.whisper_b200_get_cross_kv(w.ctx, il, buf.ctypes.data_as(C.POINTER(C.c_uint16)), buf.size) == 0
            kv = buf.view(np.float16).astype(np.float32).reshape(1500, 2 * d)
            dk, dv = np.abs(kv[:, :d] - rk[il]).max(), np.abs(kv[:, d:] - rv[il]).max()
            print(f"layer {il}: cross K max|d| = {dk:.3e}, V max|d| = {dv:.3e} (rms K {np.sqrt((rk[il] ** 2).mean()):.3f})")
            assert dk <= 2e-2 and dv <= 2e-2


def _quant_paths(model_dir, arch, qtype):
    q = os.path.join(model_dir, f"{arch}-{qtype}.bin")
    t = os.path.join(model_dir, f"{arch}-{qtype}-expanded.bin")
    if not os.path.exists(q):
        modelgen.write_model(q, arch, qtype=qtype)
        modelgen.write_model(t, arch, qtype=qtype, dequantized=True)
    return q, t


def _enc_and_logits(lib_, path, pcm, toks, **kw):
    with api.Whisper(lib_, path, flash_attn=False, **kw) as w:
        n_thr = min(os.cpu_count() or 4, 32)
        assert w.pcm_to_mel(pcm, n_thr) == 0 and w.encode(0, n_thr) == 0
        rc, lg = w.decode(toks, 0, n_thr)
        assert rc == 0
        return lg.copy()


@pytest.mark.parametrize("qtype", sorted(modelgen.QUANT_TYPES))
def test_quantised_file_loads_as_its_expansion(lib, model_dir, qtype):
    """A block-quantised GGML file (what the reference's `quantize` tool writes) is expanded to 16-bit weights
    at load: the run must be BIT-identical to the run on an F16 file holding the oracle's expansion of the same blocks
    (modelgen.dequantize_blocks, pinned against ggml's dequantize_row_* in tests/test_oracle_pinning.py)."""
    q, t = _quant_paths(model_dir, "tiny.en", qtype)
    pcm = pcm_for({"kind": "jfk"})
    toks = [50257]          # <|startoftranscript|> of the *.en vocabularies
    a = _enc_and_logits(lib, q, pcm, toks)
    b = _enc_and_logits(lib, t, pcm, toks)
    assert np.isfinite(a).all()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def _ref_kquantizer(ref):
    def fn(x, qtype):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
        raw = np.zeros(len(x) // 256 * modelgen.KQUANT_TYPES[qtype][2], np.uint8)
        getattr(ref, f"quantize_row_{qtype}_ref")(x.ctypes.data_as(FP), raw.ctypes.data_as(C.c_void_p), C.c_int64(len(x)))
        return raw.tobytes()
    return fn


@pytest.mark.parametrize("qtype", ["q4_K", "q6_K"])
def test_kquant_file_loads_as_its_expansion(lib, model_dir, qtype):
    """A K-quant file (256-element super-blocks, quantised here by the reference's own quantize_row_*_ref) against the F16 file
    holding the oracle's expansion of the same blocks: bit-identical logits; and against the reference's CPU run of the same
    file.  base.en: the smallest geometry whose rows are multiples of 256."""
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    q = os.path.join(model_dir, f"base.en-{qtype}.bin")
    t = os.path.join(model_dir, f"base.en-{qtype}-expanded.bin")
    if not os.path.exists(q):
        modelgen.write_model(q, "base.en", qtype=qtype, quantize_fn=_ref_kquantizer(ref))
        modelgen.write_model(t, "base.en", qtype=qtype, quantize_fn=_ref_kquantizer(ref), dequantized=True)
    pcm = pcm_for({"kind": "jfk"})
    a = _enc_and_logits(lib, q, pcm, [50257])
    b = _enc_and_logits(lib, t, pcm, [50257])
    assert np.isfinite(a).all() and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    r = _enc_and_logits(ref, q, pcm, [50257], use_gpu=False)
    dl = np.abs(a - r)
    print(f"{qtype}: vs reference CPU run logits max|d| = {dl.max():.3e} mean|d| = {dl.mean():.3e}")
    assert dl.max() <= 0.2 and dl.mean() <= 0.03        # measured on B200: 7.2e-2 / 1.4e-2 (q4_K), 7.9e-2 / 1.4e-2 (q6_K)


@pytest.mark.parametrize("qtype", ["q8_0", "q5_0"])
def test_quantised_model_vs_live_reference(lib, model_dir, qtype):
    """Same quantised file through the reference CPU path.  The reference multiplies quantised weights with activations it
    quantises to Q8 blocks on the fly (ggml vec_dot_q*_q8_0), ours multiplies the expanded weights with f16 activations, so
    the two agree to the activation-quantisation noise, not to the 2e-2 of the F16 files; the greedy token must agree."""
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    q, _ = _quant_paths(model_dir, "tiny.en", qtype)
    pcm = pcm_for({"kind": "jfk"})
    toks = [50257]
    a = _enc_and_logits(lib, q, pcm, toks)
    b = _enc_and_logits(ref, q, pcm, toks, use_gpu=False)
    dl = np.abs(a - b)
    top2 = np.sort(b[:50256])[-2:]
    print(f"{qtype}: logits max|d| = {dl.max():.3e} mean|d| = {dl.mean():.3e}; reference top-2 gap {top2[1] - top2[0]:.3f}")
    assert dl.max() <= 0.15 and dl.mean() <= 0.02       # measured on B200: 6.6e-2 / 1.1e-2 (q8_0), 6.2e-2 / 1.1e-2 (q5_0)
    if top2[1] - top2[0] > 2 * dl.max():
        assert int(a[:50256].argmax()) == int(b[:50256].argmax())


@pytest.mark.parametrize("n_tok", [130, 200, 384])
def test_short_prompt_row_blocks_vs_live_reference(lib, model_dir, n_tok):
    """129..512 decoder rows in one call (short prompts of many windows; here one window with a long forced prompt) run the
    weight-streaming GEMM once per block of 128 rows instead of 128 x 256 tensor-core tiles: last-row logits against the
    live reference."""
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    path = model_path(model_dir, "tiny.en")
    pcm = pcm_for({"kind": "jfk"})
    toks = [50257] + [int(t) for t in np.random.default_rng(n_tok).integers(0, 50000, n_tok - 1)]
    n_thr = min(os.cpu_count() or 4, 32)
    a = _enc_and_logits(lib, path, pcm, toks)
    b = _enc_and_logits(ref, path, pcm, toks, use_gpu=False)
    print(f"{n_tok} prompt rows: last-row logits max|d| = {np.abs(a - b).max():.3e}")
    assert np.abs(a - b).max() <= 2e-2


def test_bf16_operands_track_f16(lib, model_dir, monkeypatch):
    """WHISPER_B200_DTYPE=bf16 (the operand type BASELINE.json names) runs every kernel of the path -- tcgen05 GEMMs and
    attention, decoder step -- and stays close to the f16 run (bf16 has 3 fewer mantissa bits: ~8x the rounding error)."""
    pcm = pcm_for({"kind": "jfk"})

    def run(dtype):
        monkeypatch.setenv("WHISPER_B200_DTYPE", dtype)
        with api.Whisper(lib, model_path(model_dir, "base.en"), flash_attn=False) as w:
            assert lib.whisper_b200_dtype(w.ctx) == (1 if dtype == "bf16" else 0)
            assert w.pcm_to_mel(pcm) == 0 and w.encode(0) == 0
            enc = get_enc(lib, w, 512)
            toks, out = [lib.whisper_token_sot(w.ctx)], []
            for step in range(6):
                rc, lg = w.decode(toks[-1:] if step else toks, 0 if step == 0 else len(toks) - 1)
                assert rc == 0
                out.append(lg)
                toks.append(int(lg[:50000].argmax()))
            return enc, np.stack(out)

    e16, l16 = run("f16")
    eb, lb = run("bf16")
    de, dl = np.abs(e16 - eb).max(), np.abs(l16[0] - lb[0]).max()
    print(f"bf16 vs f16: embd_enc max|d| = {de:.3e}, first-step logits max|d| = {dl:.3e}")
    assert np.isfinite(lb).all() and de <= 6e-2 and dl <= 1.5e-1


def test_set_mel_empty_runs_like_whisper_bench(lib, model_dir):
    """examples/bench/bench.cpp:84 feeds an empty mel (encoder input all zeros) and times encode/decode."""
    with api.Whisper(lib, model_path(model_dir, "tiny.en")) as w:
        assert lib.whisper_set_mel(w.ctx, None, 0, 80) == 0
        assert w.encode(0) == 0
        rc, lg = w.decode([lib.whisper_token_sot(w.ctx)] * 5, 0)
        assert rc == 0 and np.isfinite(lg).all()
        assert lib.whisper_set_mel(w.ctx, None, 0, 64) == -1          # wrong number of mel bands


def test_error_conventions(lib, model_dir, tmp_path):
    cp = lib.whisper_context_default_params()
    assert lib.whisper_init_from_file_with_params(b"/nonexistent/model.bin", cp) is None
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"\x00" * 64)
    assert lib.whisper_init_from_file_with_params(str(bad).encode(), cp) is None
    cp.use_gpu = False                                               # no CPU fallback exists: refuse loudly
    assert lib.whisper_init_from_file_with_params(model_path(model_dir, "tiny.en").encode(), cp) is None
    with api.Whisper(lib, model_path(model_dir, "tiny.en")) as w:
        p = w.greedy_params()
        p.greedy.best_of = 9
        rc, _ = w.full(p, modelgen.synth_pcm(32000))
        assert rc == -4                                              # too many decoders (src/whisper.cpp:6912-6915)
        p = w.greedy_params()
        rc, segs = w.full(p, modelgen.synth_pcm(800))                # < 100 ms: returns 0 with no segments
        assert rc == 0 and segs == []
        assert lib.whisper_tokenize(w.ctx, b" hello world", None, 0) < 0


def test_header_only_test_model_completes(lib, model_dir):
    """The reference's ctest smoke runs on header-only model files (models/for-tests-*.bin): exit 0, no segments."""
    p = os.path.join(model_dir, "header-only.bin")
    modelgen.write_model(p, "tiny.en", with_tensors=False)
    with api.Whisper(lib, p) as w:
        rc, segs = w.full(w.greedy_params(), pcm_for({"kind": "jfk"}))
        assert rc == 0 and segs == []


def test_beam_and_temperature_fallback_host_path_runs(lib, model_dir):
    """cli default is beam 5 with temperature fallback: the host selection path (mt19937 draws) must complete."""
    with api.Whisper(lib, model_path(model_dir, "tiny.en")) as w:
        p = w.default_params(1)
        p.print_progress = False
        rc, segs = w.full(p, modelgen.synth_pcm(160000, stream=3))
        assert rc == 0
        p = w.default_params(0)
        p.print_progress = False
        rc, segs = w.full(p, modelgen.synth_pcm(160000, stream=3))
        assert rc == 0


def test_reference_cli_and_bench_run_on_our_library(lib, model_dir):
    """BASELINE.json config 1: the reference's UNMODIFIED whisper-cli (examples/cli/cli.cpp), compiled against our headers and
    linked to our library, transcribes samples/jfk.wav greedily and prints what the same cli linked to the reference build
    prints; whisper-bench (examples/bench/bench.cpp) runs its encode/decode timing loop on our library."""
    import subprocess
    bindir = os.path.join(os.path.dirname(HERE), "oracle", "_ref")
    cli, ref_cli, bench = (os.path.join(bindir, n) for n in ("whisper-cli-b200", "whisper-cli-ref", "whisper-bench-b200"))
    if not (os.path.exists(cli) and os.path.exists(bench)):
        pytest.skip("reference callers were not built (needs /root/reference at build time)")
    model = model_path(model_dir, "tiny.en")
    wav = os.path.join(HERE, "golden", "jfk.wav")
    args = ["-m", model, "-f", wav, "-bs", "1", "-bo", "1", "-nf", "-ps"]
    ours = subprocess.run([cli] + args, capture_output=True, text=True, timeout=300)
    assert ours.returncode == 0, ours.stderr[-2000:]
    lines = [l for l in ours.stdout.splitlines() if l.startswith("[")]
    assert lines, ours.stdout
    if os.path.exists(ref_cli):
        theirs = subprocess.run([ref_cli] + args + ["-ng", "-nfa", "-t", "4"], capture_output=True, text=True, timeout=600)
        assert theirs.returncode == 0, theirs.stderr[-2000:]
        ref_lines = [l for l in theirs.stdout.splitlines() if l.startswith("[")]
        print("cli output:", lines, "| reference:", ref_lines)
        # our cli ran with the default flash_attn=true; token-identical on this fixture in both modes (golden_tokens.json)
        assert lines == ref_lines
    b = subprocess.run([bench, "-m", model, "-t", "1"], capture_output=True, text=True, timeout=300)
    assert b.returncode == 0, b.stderr[-2000:]
    assert "encode time" in b.stderr and "decode time" in b.stderr
