"""DTW token timestamps (whisper_context_params::dtw_token_timestamps -> whisper_token_data::t_dtw) against the live reference.

Reference: after each window the decoder runs once more over [sot, (lang), notimestamps, text tokens, eot], the cross-attention
probabilities of the alignment heads are normalised, median-filtered, averaged and dynamic-time-warped against the audio
positions (src/whisper.cpp:2721-2737, 8683-8998).  Product: the same extra pass with the alignment heads' probabilities
recomputed on the device (csrc/dtw.cu::cross_align_kernel) and the alignment on the host.

The path is a chain of discrete arg-min decisions over costs that differ between the implementations at the 1e-3 level
(16-bit operands), so single path cells may flip near ties; the bar is: identical greedy tokens / segments, t_dtw set on exactly
the same tokens, monotone, and equal to the reference's within one audio position (2 ticks = 20 ms) on at least 90 % of the
tokens with no token further than 0.2 s away."""
import os

import numpy as np
import pytest

from open_whisper_kit_b200 import api, modelgen
from oracle import reflib

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("arch,preset,heads", [("tiny.en", 3, None), ("tiny", 2, [(1, 2), (2, 0), (3, 5)]), ("base.en", 1, None)])
def test_dtw_token_timestamps_vs_live_reference(lib, model_dir, arch, preset, heads):
    ref, _ = reflib.load()
    if ref is None:
        pytest.skip("oracle/_ref was not built / did not travel")
    path = os.path.join(model_dir, f"{arch}-1.bin")
    if not os.path.exists(path):
        modelgen.write_model(path, arch, ftype=1)
    pcm = api.read_wav_f32(os.path.join(HERE, "golden", "jfk.wav"))
    res = {}
    for name, L, kw in (("ours", lib, {}), ("ref", ref, {"use_gpu": False})):
        with api.Whisper(L, path, flash_attn=False, dtw_preset=preset, dtw_heads=heads, dtw_n_top=2, **kw) as w:
            p = w.greedy_params(no_timestamps=True, n_threads=min(os.cpu_count() or 4, 16))
            p.max_tokens = 60
            rc, segs = w.full(p, pcm)
            assert rc == 0
            res[name] = [(s.t0, s.t1, [(t.id, t.t_dtw) for t in s.token_data]) for s in segs]
    ours = [x for s in res["ours"] for x in s[2]]
    theirs = [x for s in res["ref"] for x in s[2]]
    assert [(s[0], s[1]) for s in res["ours"]] == [(s[0], s[1]) for s in res["ref"]]
    assert [a for a, _ in ours] == [a for a, _ in theirs], "greedy tokens differ"
    set_ours = [t >= 0 for _, t in ours]
    assert set_ours == [t >= 0 for _, t in theirs] and any(set_ours)
    a = np.array([t for _, t in ours if t >= 0])
    b = np.array([t for _, t in theirs if t >= 0])
    assert np.all(np.diff(a) >= 0)
    d = np.abs(a - b)
    print(f"{arch}: {len(a)} aligned tokens, |t_dtw - reference| max {d.max()} ticks, identical on {(d == 0).mean():.2%}")
    assert (d <= 2).mean() >= 0.9 and d.max() <= 20
