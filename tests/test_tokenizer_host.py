"""Host logic of whisper_tokenize (regex word split + greedy longest match, src/whisper.cpp:3272-3320) through the host-only hook
whisper_b200_tokenize against the compiled reference's public whisper_tokenize on the same vocabulary.  No device needed."""
import ctypes as C
import os

import numpy as np
import pytest

import open_whisper_kit_b200 as pkg
from open_whisper_kit_b200 import api, modelgen
from oracle import reflib

IP = C.POINTER(C.c_int32)

TEXTS = [" hello world", "Hello, world!", "it's 12:30 -- we'll go", "  leading  and   trailing   ", "numbers 1234567890 and symbols #$%^&*()",
         "don't can't I'm they've she'd", "UPPER lower MiXeD", "tab\tand\nnewline", "café naïve 中文 \U0001F600", "a", " ", "",
         "x" * 300, "the quick brown fox jumps over the lazy dog " * 5]


@pytest.mark.parametrize("arch", ["tiny.en", "tiny"])
def test_tokenizer_matches_reference(arch, tmp_path):
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    lib = pkg.load()
    lib.whisper_b200_tokenize.restype = C.c_int
    lib.whisper_b200_tokenize.argtypes = [C.c_void_p, C.c_int, C.c_char_p, IP, C.c_int]
    path = os.path.join(str(tmp_path), f"{arch}-header.bin")
    modelgen.write_model(path, arch, with_tensors=False)
    with api.Whisper(ref, path, use_gpu=False, flash_attn=False) as w:
        n_vocab = ref.whisper_n_vocab(w.ctx)
        texts = [ref.whisper_token_to_str(w.ctx, i) or b"" for i in range(n_vocab)]
        arr = (C.c_char_p * n_vocab)(*texts)
        rng = np.random.default_rng(3)
        cases = list(TEXTS)
        for _ in range(40):      # random concatenations of real vocabulary entries: every longest-match decision is exercised
            ids = rng.integers(0, ref.whisper_token_eot(w.ctx), int(rng.integers(1, 12)))
            cases.append(b"".join(texts[i] for i in ids).decode("utf-8", "ignore"))
        n_nonempty = 0
        for text in cases:
            raw = text.encode("utf-8")
            if b"\x00" in raw:
                continue
            cap = 1024
            a, b = (C.c_int32 * cap)(), (C.c_int32 * cap)()
            na = lib.whisper_b200_tokenize(C.cast(arr, C.c_void_p), n_vocab, raw, a, cap)
            nb = ref.whisper_tokenize(w.ctx, raw, b, cap)
            assert na == nb, (text, na, nb)
            assert list(a[:max(na, 0)]) == list(b[:max(nb, 0)]), text
            n_nonempty += na > 0
            # the "buffer too small" convention
            if na > 1:
                assert lib.whisper_b200_tokenize(C.cast(arr, C.c_void_p), n_vocab, raw, a, na - 1) == -na
                assert ref.whisper_tokenize(w.ctx, raw, b, na - 1) == -na
        assert n_nonempty >= 40
