"""Product-level multi-GPU partition (include/whisper_b200.h: whisper_b200_group_*): the chunks of a whisper_full_parallel
call are dealt out to one model replica per GPU in contiguous blocks, decoded without any collective, and gathered on the
host in chunk order with the reference's timestamp fix-up (src/whisper.cpp:7879-7889).  The gathered result must be exactly
what the single-GPU whisper_full_parallel call produces -- same segments, times, tokens, probabilities.

Needs >= 2 visible GPUs for the real thing (`gpurun --gpus 2`); on a one-GPU box the group degenerates to one replica and the
same equality is checked there.
"""
import ctypes as C
import os

import numpy as np
import pytest

from open_whisper_kit_b200 import api, capi, modelgen

pytestmark = pytest.mark.gpu


def _segments(lib, ctx):
    w = api.Whisper.__new__(api.Whisper)
    w.lib, w.ctx = lib, ctx
    return [(int(s.t0), int(s.t1), [int(t) for t in s.tokens], [round(float(t.p), 7) for t in s.token_data]) for s in w.segments()]


@pytest.mark.parametrize("n_chunks,mode", [(6, "greedy"), (5, "beam")])
def test_group_result_equals_single_gpu(lib, model_dir, n_chunks, mode):
    n_gpu = lib.whisper_b200_device_count()
    path = os.path.join(model_dir, "tiny-1.bin")
    if not os.path.exists(path):
        modelgen.write_model(path, "tiny", ftype=1)
    pcm = np.concatenate([modelgen.synth_pcm(480000, seed=5, stream=i) for i in range(n_chunks)])
    with api.Whisper(lib, path, flash_attn=False) as w:
        if mode == "greedy":
            p = w.greedy_params(no_timestamps=False)
        else:
            p = w.default_params(capi.BEAM_SEARCH)
            p.temperature_inc = 0.0
            p.print_progress = False
        rc, _ = w.full(p, pcm, n_processors=n_chunks)
        assert rc == 0
        single = _segments(lib, w.ctx)
        cp = lib.whisper_context_default_params()
        cp.flash_attn = False
        g = lib.whisper_b200_group_init_from_file(path.encode(), cp, None, min(n_gpu, 2))
        assert g
        try:
            assert lib.whisper_b200_group_size(g) == min(n_gpu, 2)
            rc = lib.whisper_b200_group_full_parallel(g, p, capi.as_float_ptr(pcm), len(pcm), n_chunks)
            assert rc == 0
            multi = _segments(lib, lib.whisper_b200_group_context(g, 0))
        finally:
            lib.whisper_b200_group_free(g)
    print(f"{n_gpu} GPU(s) visible, group of {min(n_gpu, 2)}: {len(single)} segments, {sum(len(s[2]) for s in single)} tokens")
    assert len(single) > 0 and multi == single
