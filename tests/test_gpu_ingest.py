"""16-bit PCM ingest fused into the mel kernel's load (include/whisper_b200.h).

The reference's callers decode audio to f32 on the host before the API (examples/common-whisper.cpp:42-134; for 16-bit
PCM miniaudio's conversion is x = s / 32768, exact in f32).  Passing the int16 samples straight to the product must give
bit-identical mel values and the same transcription as passing the converted floats -- and, through that, the reference's."""
import ctypes as C
import os

import numpy as np
import pytest

from open_whisper_kit_b200 import api, modelgen

pytestmark = pytest.mark.gpu
FP = C.POINTER(C.c_float)
HERE = os.path.dirname(os.path.abspath(__file__))


def _get_mel(lib, ctx):
    n_len, n_mel = C.c_int(), C.c_int()
    assert lib.whisper_b200_get_mel(ctx, None, None, 0, C.byref(n_len), C.byref(n_mel)) == 0
    mel = np.empty((n_mel.value, n_len.value), np.float32)
    assert lib.whisper_b200_get_mel(ctx, None, mel.ctypes.data_as(FP), mel.size, C.byref(n_len), C.byref(n_mel)) == 0
    return mel


@pytest.mark.parametrize("n_samples", [176000, 480000 * 2 + 12345, 799])
def test_int16_ingest_is_bit_identical_to_float_ingest(lib, model_dir, n_samples):
    path = os.path.join(model_dir, "tiny.en-1.bin")
    if not os.path.exists(path):
        modelgen.write_model(path, "tiny.en", ftype=1)
    rng = np.random.default_rng(n_samples)
    pcm16 = np.clip(rng.normal(0, 3000, n_samples) + 8000 * np.sin(np.arange(n_samples) * 0.05), -32768, 32767).astype(np.int16)
    pcm16[:3] = [-32768, 32767, 0]
    pcmf = (pcm16.astype(np.float32) / np.float32(32768.0)).astype(np.float32)
    with api.Whisper(lib, path, flash_attn=False) as w:
        assert w.pcm_to_mel(pcmf) == 0
        mel_f = _get_mel(lib, w.ctx)
        assert lib.whisper_b200_pcm16_to_mel(w.ctx, pcm16.ctypes.data_as(C.POINTER(C.c_int16)), n_samples) == 0
        mel_i = _get_mel(lib, w.ctx)
        assert mel_f.shape == mel_i.shape and np.array_equal(mel_f, mel_i)
        if n_samples > 16000:
            p = w.greedy_params(no_timestamps=False)
            p.token_timestamps = True
            nproc = 2 if n_samples > 480000 else 1
            rc, segs_f = w.full(p, pcmf, n_processors=nproc)
            assert rc == 0
            rc = lib.whisper_b200_full_parallel_i16(w.ctx, p, pcm16.ctypes.data_as(C.POINTER(C.c_int16)), n_samples, nproc)
            assert rc == 0
            segs_i = w.segments()
            key = lambda ss: [(s.t0, s.t1, s.tokens, [(t.t0, t.t1, t.p) for t in s.token_data]) for s in ss]     # noqa: E731
            assert len(segs_f) > 0 and key(segs_f) == key(segs_i)
