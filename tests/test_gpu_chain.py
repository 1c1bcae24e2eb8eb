"""Persistent chain kernel of the single-token decoder step (csrc/dec_chain.cu) against the unfused launch sequence.

Both run the same arithmetic (16-bit operands, f32 accumulation, identical rounding points); only the f32 summation
order of the stream-K partial tiles differs, so logits must agree far inside the 2e-2 gate that either path has to meet
against the reference (tests/test_gpu_model.py runs the golden fixtures through the chain path, which is the default).
"""
import os

import numpy as np
import pytest

from open_whisper_kit_b200 import api, modelgen

from test_gpu_model import model_path, pcm_for

pytestmark = pytest.mark.gpu


def _with_chain(mode, fn):
    """mode 0: separate launches, 1: one chain launch per layer half, 2: hybrid (see Engine::chain_usable)."""
    old = os.environ.get("WHISPER_B200_CHAIN")
    os.environ["WHISPER_B200_CHAIN"] = str(int(mode))
    try:
        return fn()
    finally:
        if old is None:
            os.environ.pop("WHISPER_B200_CHAIN", None)
        else:
            os.environ["WHISPER_B200_CHAIN"] = old


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("arch,fa", [("tiny.en", False), ("base.en", True), ("tiny", False)])
def test_chain_step_logits_match_unfused(lib, model_dir, arch, fa, mode):
    pcm = pcm_for({"kind": "jfk"})

    def run(on):
        def body():
            with api.Whisper(lib, model_path(model_dir, arch), flash_attn=fa) as w:
                assert w.pcm_to_mel(pcm) == 0 and w.encode(0) == 0
                tok, out = lib.whisper_token_sot(w.ctx), []
                for n_past in range(24):                      # one token per call: the single-token step
                    rc, lg = w.decode([tok], n_past)
                    assert rc == 0
                    out.append(lg)
                    tok = int(lg[:50000].argmax())
                return np.stack(out)
        return _with_chain(on, body)

    a, b = run(mode), run(0)
    d = np.abs(a - b).max(axis=1)
    print(f"{arch} fa={fa} mode={mode}: chain vs unfused logits max|d| per step: {d.max():.3e}")
    assert d.max() <= 4e-3
    assert (a[:, :50000].argmax(axis=1) == b[:, :50000].argmax(axis=1)).all()


@pytest.mark.parametrize("mode", [1, 2])
def test_chain_batched_more_than_64_rows(lib, model_dir, mode):
    """70 sequences in one decoder batch: more than 64 live accumulator rows (the 16 KB activation-tile geometry)."""
    n_win = 70
    pcm = np.concatenate([modelgen.synth_pcm(480000, seed=11, stream=i) for i in range(n_win)])

    def run(on):
        def body():
            with api.Whisper(lib, model_path(model_dir, "tiny.en"), flash_attn=False) as w:
                p = w.greedy_params(no_timestamps=True)
                p.max_tokens = 12
                rc, segs = w.full(p, pcm, n_processors=n_win)
                assert rc == 0
                return [tuple(s.tokens) for s in segs]
        return _with_chain(on, body)

    a, b = run(mode), run(0)
    assert len(a) == len(b) == n_win
    same = sum(x == y for x, y in zip(a, b))
    print(f"windows with identical token sequences: {same}/{n_win}")
    assert same >= n_win - 2          # near-tie flips of a random-init model are the only allowed cause
