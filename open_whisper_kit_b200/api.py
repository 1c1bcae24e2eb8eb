"""Thin object wrapper over the whisper.h C ABI (works on any library that exports it).

Mirrors how the reference's own callers drive the API (examples/cli/cli.cpp:1039-1276,
examples/bench/bench.cpp:63-165): init -> [pcm_to_mel -> encode -> decode] or full -> read segments.
"""
import ctypes as C
import wave

import numpy as np

from . import capi


def read_wav_f32(path):
    """16-bit PCM mono/stereo 16 kHz WAV -> float32 mono in [-1, 1) (what examples/common-whisper.cpp:42-134 yields)."""
    with wave.open(path, "rb") as w:
        assert w.getsampwidth() == 2 and w.getframerate() == 16000
        raw = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
        if w.getnchannels() == 2:
            raw = raw.reshape(-1, 2).astype(np.int32).sum(axis=1) // 2
        return (raw.astype(np.float32) / 32768.0).astype(np.float32)


class Segment:
    def __init__(self, t0, t1, text, tokens, token_data, no_speech_prob):
        self.t0, self.t1, self.text = t0, t1, text
        self.tokens, self.token_data, self.no_speech_prob = tokens, token_data, no_speech_prob

    def __repr__(self):
        return f"Segment({self.t0},{self.t1},{self.text!r},{self.tokens})"


class Whisper:
    def __init__(self, lib, model_path, use_gpu=True, flash_attn=True, gpu_device=0, dtw_preset=None, dtw_heads=None, dtw_n_top=-1):
        """dtw_preset: a whisper_alignment_heads_preset value (1 = N_TOP_MOST with dtw_n_top, 2 = CUSTOM with dtw_heads =
        [(text layer, head), ...]) switches DTW token timestamps on (the reference refuses them with flash_attn)."""
        self.lib = lib
        cp = lib.whisper_context_default_params()
        cp.use_gpu = use_gpu
        cp.flash_attn = flash_attn
        cp.gpu_device = gpu_device
        if dtw_preset is not None:
            cp.dtw_token_timestamps = True
            cp.dtw_aheads_preset = dtw_preset
            cp.dtw_n_top = dtw_n_top
            if dtw_heads:
                self._heads = (capi.whisper_ahead * len(dtw_heads))(*[capi.whisper_ahead(l, h) for l, h in dtw_heads])
                cp.dtw_aheads.n_heads = len(dtw_heads)
                cp.dtw_aheads.heads = self._heads
        self.ctx = lib.whisper_init_from_file_with_params(model_path.encode(), cp)
        if not self.ctx:
            raise RuntimeError(f"whisper_init_from_file_with_params failed for {model_path}")

    def close(self):
        if self.ctx:
            self.lib.whisper_free(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- low-level trio -------------------------------------------------------------------
    def pcm_to_mel(self, pcm, n_threads=1):
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        return self.lib.whisper_pcm_to_mel(self.ctx, capi.as_float_ptr(pcm), len(pcm), n_threads)

    def set_mel(self, mel):
        """mel: [n_mel][n_len] float32."""
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        return self.lib.whisper_set_mel(self.ctx, capi.as_float_ptr(mel), mel.shape[1], mel.shape[0])

    def encode(self, offset=0, n_threads=1):
        return self.lib.whisper_encode(self.ctx, offset, n_threads)

    def decode(self, tokens, n_past, n_threads=1):
        arr = (C.c_int32 * len(tokens))(*tokens)
        rc = self.lib.whisper_decode(self.ctx, arr, len(tokens), n_past, n_threads)
        if rc != 0:
            return rc, None
        n_vocab = self.lib.whisper_n_vocab(self.ctx)
        p = self.lib.whisper_get_logits(self.ctx)
        # the reference exposes [n_tokens][n_vocab] with only the last row valid (src/whisper.cpp:2949-2955)
        full = np.ctypeslib.as_array(p, shape=(len(tokens) * n_vocab,))
        return rc, full[(len(tokens) - 1) * n_vocab:].copy()

    # ---- whisper_full ---------------------------------------------------------------------
    def default_params(self, strategy=capi.GREEDY):
        return self.lib.whisper_full_default_params(strategy)

    def greedy_params(self, no_timestamps=False, n_threads=1, language=b"en"):
        """whisper-cli -bs 1 -bo 1 -nf: greedy, one decoder, no temperature fallback."""
        p = self.default_params(capi.GREEDY)
        p.greedy.best_of = 1
        p.temperature_inc = 0.0
        p.n_threads = n_threads
        p.no_timestamps = no_timestamps
        p.print_progress = False
        p.language = language
        return p

    def full(self, params, pcm, n_processors=1):
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        if n_processors == 1:
            rc = self.lib.whisper_full(self.ctx, params, capi.as_float_ptr(pcm), len(pcm))
        else:
            rc = self.lib.whisper_full_parallel(self.ctx, params, capi.as_float_ptr(pcm), len(pcm), n_processors)
        return rc, self.segments() if rc == 0 else []

    def segments(self):
        lib, ctx = self.lib, self.ctx
        out = []
        for i in range(lib.whisper_full_n_segments(ctx)):
            n_tok = lib.whisper_full_n_tokens(ctx, i)
            tds = [lib.whisper_full_get_token_data(ctx, i, j) for j in range(n_tok)]
            out.append(Segment(lib.whisper_full_get_segment_t0(ctx, i), lib.whisper_full_get_segment_t1(ctx, i),
                               lib.whisper_full_get_segment_text(ctx, i),
                               [t.id for t in tds], tds,
                               lib.whisper_full_get_segment_no_speech_prob(ctx, i)))
        return out

    def all_tokens(self):
        return [t for s in self.segments() for t in s.tokens]


class Vad:
    """whisper_vad_* of include/whisper.h: Silero voice activity detection (speech probability per 512-sample chunk, and the
    speech segments derived from them).  Works on either library; the product runs the model on the GPU and has no CPU path."""

    def __init__(self, lib, model_path, gpu_device=0):
        self.lib = lib
        cp = lib.whisper_vad_default_context_params()
        cp.gpu_device = gpu_device
        self.vctx = lib.whisper_vad_init_from_file_with_params(model_path.encode(), cp)
        if not self.vctx:
            raise RuntimeError(f"whisper_vad_init_from_file_with_params failed for {model_path}")

    def close(self):
        if self.vctx:
            self.lib.whisper_vad_free(self.vctx)
            self.vctx = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def default_params(self):
        return self.lib.whisper_vad_default_params()

    def probs(self):
        n = self.lib.whisper_vad_n_probs(self.vctx)
        if n == 0:
            return np.zeros(0, np.float32)
        return np.ctypeslib.as_array(self.lib.whisper_vad_probs(self.vctx), shape=(n,)).copy()

    def detect(self, pcm, stateful=False):
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        fn = self.lib.whisper_vad_detect_speech_stateful if stateful else self.lib.whisper_vad_detect_speech
        if not fn(self.vctx, capi.as_float_ptr(pcm), len(pcm)):
            raise RuntimeError("whisper_vad_detect_speech failed")
        return self.probs()

    def reset(self):
        self.lib.whisper_vad_reset_state(self.vctx)

    def _take(self, segs):
        if not segs:
            raise RuntimeError("whisper_vad_segments_* failed")
        n = self.lib.whisper_vad_segments_n_segments(segs)
        out = [(self.lib.whisper_vad_segments_get_segment_t0(segs, i), self.lib.whisper_vad_segments_get_segment_t1(segs, i))
               for i in range(n)]
        self.lib.whisper_vad_free_segments(segs)
        return out

    def segments_from_probs(self, params=None):
        return self._take(self.lib.whisper_vad_segments_from_probs(self.vctx, params or self.default_params()))

    def segments_from_samples(self, pcm, params=None):
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        return self._take(self.lib.whisper_vad_segments_from_samples(self.vctx, params or self.default_params(),
                                                                     capi.as_float_ptr(pcm), len(pcm)))
