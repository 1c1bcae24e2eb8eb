"""Small end-to-end pass over every kernel added in round 2, meant to run under `compute-sanitizer --tool memcheck`:
tiny model, 3 windows -- greedy + timestamps (float and int16 PCM), "beam search" 5 (device draws, batched history copies,
grouped cross-attention), temperature fallback, DTW token timestamps, the LayerNorm-folded decoder GEMMs; base model, 3 windows --
the 2-CTA encoder GEMM with its epilogues."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import open_whisper_kit_b200 as pkg  # noqa: E402
from open_whisper_kit_b200 import api, capi, modelgen  # noqa: E402

lib = pkg.load()
os.makedirs("/tmp/models", exist_ok=True)
path = "/tmp/models/tiny-1.bin"
if not os.path.exists(path):
    modelgen.write_model(path, "tiny")
cb = capi.LOG_CB(lambda level, text, ud: None)
lib.whisper_log_set(C.cast(cb, C.c_void_p), None)
pcm = np.concatenate([modelgen.synth_pcm(480000, seed=7, stream=i) for i in range(3)])
pcm16 = np.clip(pcm * 32768.0, -32768, 32767).astype(np.int16)


def run(tag, w, p, n_proc=3, i16=False):
    if i16:
        rc = lib.whisper_b200_full_parallel_i16(w.ctx, p, pcm16.ctypes.data_as(C.POINTER(C.c_int16)), len(pcm16), n_proc)
        segs = w.segments()
    else:
        rc, segs = w.full(p, pcm, n_processors=n_proc)
    print(f"{tag}: rc {rc}, {len(segs)} segments, {sum(len(s.tokens) for s in segs)} tokens", flush=True)
    assert rc == 0


with api.Whisper(lib, path, flash_attn=False, dtw_preset=4) as w:          # WHISPER_AHEADS_TINY
    p = w.greedy_params(no_timestamps=False)
    p.max_tokens = 24
    run("greedy+ts+dtw f32", w, p)
    run("greedy+ts+dtw i16", w, p, i16=True)
    p.token_timestamps = True
    p.max_len = 12
    run("token timestamps", w, p, n_proc=1)
with api.Whisper(lib, path, flash_attn=True) as w:
    p = w.default_params(capi.BEAM_SEARCH)
    p.print_progress = False
    p.temperature_inc = 0.0
    p.max_tokens = 24
    run("beam 5", w, p)
    p = w.default_params(capi.GREEDY)
    p.print_progress = False
    p.greedy.best_of = 3
    p.temperature_inc = 0.4
    p.logprob_thold = -0.5
    p.max_tokens = 12
    run("temperature ladder", w, p)
# base geometry (d = 512): three windows = 4 500 rows, so the encoder GEMMs run on the 2-CTA kernel (cta_group::2 tiles, the two-wide
# GELU epilogue, the V^T epilogue); the eight-warp attention kernel with P in tensor memory and the mma.sync self-attention run in
# every pass above as well
path_b = "/tmp/models/base-1.bin"
if not os.path.exists(path_b):
    modelgen.write_model(path_b, "base")
with api.Whisper(lib, path_b, flash_attn=True) as w:
    p = w.greedy_params(no_timestamps=True)
    p.max_tokens = 8
    run("base greedy", w, p)
print("sanitize probe done")
