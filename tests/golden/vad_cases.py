"""Inputs shared by make_golden_vad.py and the VAD tests."""
import wave

import numpy as np

PARAM_SETS = {
    "default": {},
    "strict": {"threshold": 0.8, "min_speech_duration_ms": 500, "min_silence_duration_ms": 300},
    "loose": {"threshold": 0.3, "min_speech_duration_ms": 100, "min_silence_duration_ms": 50, "speech_pad_ms": 100},
    "capped": {"max_speech_duration_s": 2.0, "min_silence_duration_ms": 20},
    "split": {"max_speech_duration_s": 1.0},
    "split_pad": {"max_speech_duration_s": 1.2, "min_silence_duration_ms": 200, "speech_pad_ms": 60},
    "nopad": {"speech_pad_ms": 0, "samples_overlap": 0.0},
}


def vad_params(lib, **kw):
    p = lib.whisper_vad_default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def read_wav(path):
    with wave.open(path, "rb") as w:
        assert w.getframerate() == 16000 and w.getnchannels() == 1 and w.getsampwidth() == 2
        raw = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
    return raw.astype(np.float32) / 32768.0


def synthetic_bursts(seed=5):
    """~21 s: pieces of speech-like audio (the jfk recording, cut and rescaled) separated by silences and low noise, ending in
    a ragged tail that is not a multiple of the 512-sample chunk."""
    import os
    rng = np.random.default_rng(seed)
    jfk = read_wav(os.path.join(os.path.dirname(os.path.abspath(__file__)), "jfk.wav"))
    parts = [np.zeros(9000, np.float32), jfk[5000:45000], (1e-3 * rng.standard_normal(24000)).astype(np.float32),
             0.5 * jfk[60000:150000], np.zeros(16000, np.float32), jfk[100000:106000], np.zeros(30000, np.float32),
             1.5 * jfk[20000:100000], (3e-3 * rng.standard_normal(20333)).astype(np.float32)]
    return np.clip(np.concatenate(parts), -1.0, 1.0).astype(np.float32)
