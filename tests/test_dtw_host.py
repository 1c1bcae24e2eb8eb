"""Host stage of the DTW token timestamps (csrc/dtw.cu::dtw_align through the host-only hook whisper_b200_dtw_align) against
the numpy restatement of the reference's algorithm (oracle/dtw_oracle.py), on the CPU."""
import ctypes as C

import numpy as np
import pytest

import open_whisper_kit_b200 as pkg
from oracle import dtw_oracle

FP = C.POINTER(C.c_float)


@pytest.mark.parametrize("case", range(6))
def test_dtw_alignment_matches_oracle(case):
    lib = pkg.load()
    rng = np.random.default_rng(900 + case)
    n_heads, n_tokens, T = [(6, 12, 80), (1, 5, 64), (8, 40, 300), (3, 9, 1500), (2, 30, 50), (5, 20, 128)][case]
    n_audio = T if case % 2 == 0 else T - 13
    skip = 1 + case % 2
    # a noisy monotone alignment, like real cross-attention: token i attends around position i * n_audio / n_tokens
    centre = (np.arange(n_tokens)[:, None] + 0.5) * n_audio / n_tokens
    logits = -((np.arange(T)[None, :] - centre) ** 2) / (2 * (0.08 * n_audio + 1) ** 2) + 1.5 * rng.standard_normal((n_heads, n_tokens, T))
    probs = np.exp(logits - logits.max(axis=-1, keepdims=True))
    probs = (probs / probs.sum(axis=-1, keepdims=True)).astype(np.float32)
    want = dtw_oracle.dtw_first_positions(probs, n_audio, skip)
    out = (C.c_int * n_tokens)()
    n = lib.whisper_b200_dtw_align(np.ascontiguousarray(probs).ctypes.data_as(FP), n_heads, n_tokens, T, n_audio, skip, 7, out)
    assert n == n_tokens - skip - 1 == len(want)
    got = list(out)[:n]
    assert got == want
    assert got[0] == 0 and all(b >= a for a, b in zip(got, got[1:]))          # a monotone path that starts at the first frame
